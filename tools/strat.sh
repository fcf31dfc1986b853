python -m pytest tests/test_gpu_sampler.py tests/test_gpu_fuzz.py -m gpu -x -q 2>&1 | tail -3
for f in "" "--no-emit" "--no-emit --graph" "--graph"; do echo "== $f"; python tools/bench_strategies.py $f 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['strategy'], '%.4f ms' % d['ms_per_step'], '%.3e kept lists/s' % d['kept_lists_per_s'], d['launches_per_step'])"; done
