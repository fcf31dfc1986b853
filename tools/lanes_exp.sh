for l in 2 4; do python bench.py --steps 60 --warmup 5 --no-cpu-baseline --lanes $l 2>gpurun_out/lanes_$l.err | tail -1 > gpurun_out/lanes_$l.json; done
python bench.py --steps 60 --warmup 5 --no-cpu-baseline --lanes 3 --sets 6 2>gpurun_out/lanes_3.err | tail -1 > gpurun_out/lanes_3.json
python bench.py --steps 60 --warmup 5 --no-cpu-baseline --lanes 6 --sets 6 2>gpurun_out/lanes_6.err | tail -1 > gpurun_out/lanes_6.json
python - <<'PY'
import json
for f in ["lanes_2","lanes_3","lanes_4","lanes_6"]:
    try:
        d=json.load(open("gpurun_out/%s.json"%f)); print(f, "value %.3e"%d["value"], "ms %.4f"%d["ms_per_step"], "kernel_ms %.4f"%d["roofline"]["kernel_ms"], d["config"]["lanes"][:20])
    except Exception as e: print(f, "ERR", e, open("gpurun_out/%s.err"%f).read()[-600:])
PY
