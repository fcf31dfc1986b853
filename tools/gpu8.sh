#!/bin/bash
# 8-GPU box: host->device fabric check, weak-scaling bench lines (N = 1, 2, 4, 8), C4 training step on 8 GPUs
N=${1:-8}
nvidia-smi topo -m > gpurun_out/r02_topo_n$N.txt 2>&1
python tools/h2d_scaling.py > gpurun_out/r02_h2d_scaling_n$N.txt 2> gpurun_out/h2d.err
cat gpurun_out/r02_h2d_scaling_n$N.txt
for g in 1 2 4 8; do
  if [ $g -gt $N ]; then continue; fi
  if [ $g = 1 ]; then
    python bench.py --gpus 1 --steps 30 --warmup 5 --no-cpu-baseline --no-secondary > gpurun_out/r02_scale_n$g.json 2> gpurun_out/scale_n$g.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $g --steps 30 --warmup 5 --no-cpu-baseline --no-secondary > gpurun_out/r02_scale_n$g.json 2> gpurun_out/scale_n$g.err
  fi
  tail -n 1 gpurun_out/r02_scale_n$g.json | python tools/bench_line.py
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/c4_train_step.py --batch 64 --rankings 1000 --steps 10 > gpurun_out/r02_c4_train_step_n$N.json 2> gpurun_out/c4_n$N.err
cat gpurun_out/r02_c4_train_step_n$N.json
