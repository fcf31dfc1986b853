nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
lscpu | grep -i -E "numa|socket|model name|^CPU\(s\)" >> gpurun_out/topo.txt
for bind in 0 1; do
  PLD_NUMA_BIND=$bind python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 --no-cpu-baseline 2>gpurun_out/numa_$bind.err | tail -1 > gpurun_out/numa_$bind.json
done
PLD_NUMA_BIND=1 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>gpurun_out/numa_n1.err | tail -1 > gpurun_out/numa_n1.json
cat gpurun_out/topo.txt
python - <<'PY'
import json
for f in ["numa_0","numa_1","numa_n1"]:
    try:
        d=json.load(open("gpurun_out/%s.json"%f)); print(f, "%.3e"%d["value"], "e2e %.3e"%d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["host_numa"])
    except Exception as e: print(f, "ERR", e)
PY
