#!/bin/bash
# A/B of library variants (pldepth_b200/variants/NAME.so) on bench workloads: tools/ab_bench.sh "WL [flags];WL [flags]" v1 v2 ...
IFS=';' read -ra RUNS <<< "$1"; shift
for lib in default "$@"; do
  if [ "$lib" = default ]; then unset PLDEPTH_B200_LIB; else export PLDEPTH_B200_LIB=$PWD/pldepth_b200/variants/$lib.so; fi
  for run in "${RUNS[@]}"; do
    echo "== $lib | $run"
    python bench.py --workload $run --steps 30 --no-cpu-baseline --no-secondary --lanes 1 2>/dev/null | python tools/bench_line.py
  done
done
