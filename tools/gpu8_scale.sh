#!/bin/bash
# weak-scaling bench lines only (N = 1 and 8, optionally 2 and 4) on one 8-GPU box
for g in ${@:-1 8}; do
  if [ $g = 1 ]; then
    python bench.py --gpus 1 --steps 30 --warmup 5 --no-cpu-baseline --no-secondary > gpurun_out/r02b_scale_n$g.json 2> gpurun_out/scaleb_n$g.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $g --steps 30 --warmup 5 --no-cpu-baseline --no-secondary > gpurun_out/r02b_scale_n$g.json 2> gpurun_out/scaleb_n$g.err
  fi
  tail -n 1 gpurun_out/r02b_scale_n$g.json | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['n_gpus'], d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['config']['host_enqueue_ms_per_step'], d['e2e']['ms_per_step'], d['clocks'])"
done
