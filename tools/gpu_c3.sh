#!/bin/bash
# config-3 development loop on the GPU box: parity tests of the long-list paths, bench rows, one full ncu capture
TAG=${1:-c3}
python -m pytest tests/test_gpu_listmle.py tests/test_gpu_fuzz.py tests/test_gpu_guards.py -x -q -m gpu 2>&1 | tail -3
for l in 1 3; do
  python bench.py --workload C3 --steps 20 --no-cpu-baseline --lanes $l > gpurun_out/r02_${TAG}_lanes$l.json 2> gpurun_out/r02_${TAG}_lanes$l.err
  python tools/bench_line.py gpurun_out/r02_${TAG}_lanes$l.json
done
python bench.py --workload C3 --steps 20 --no-cpu-baseline --lanes 1 --no-emit | python tools/bench_line.py
ncu --set full --clock-control none --import-source on -k regex:lists_tab -s 3 -c 1 -o gpurun_out/r02_${TAG} -f \
  python bench.py --workload C3 --steps 3 --warmup 3 --no-cpu-baseline --lanes 1 > gpurun_out/ncu_${TAG}.log 2>&1
