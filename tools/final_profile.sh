# round-end artefacts: tests, bench lines, launch list and one full capture of the dominant kernel
TAG=${1:-r02}
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/pytest.log
python bench.py --steps 50 --warmup 5 > gpurun_out/${TAG}_bench_default.json 2> gpurun_out/bench_default.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/bench_reference.err
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_final.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lists_small -c 2 -o gpurun_out/${TAG}_fused_final -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/ncu2.log 2>&1
for st in thresholded information; do for em in emit no-emit; do
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_scored_${st}_${em}.csv python tools/profile_scored_step.py $st $em > gpurun_out/ncu_scored.log 2>&1
done; done
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
tail -n 2 gpurun_out/pytest.log; tail -n 2 gpurun_out/smoke.log
