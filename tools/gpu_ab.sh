#!/bin/bash
# A/B of library variants (pldepth_b200/variants/*.so, selected through PLDEPTH_B200_LIB) on one bench workload
WL=${1:-C3}; shift
for lib in default "$@"; do
  if [ "$lib" = default ]; then unset PLDEPTH_B200_LIB; else export PLDEPTH_B200_LIB=$PWD/pldepth_b200/variants/$lib; fi
  for extra in "" "--no-emit"; do
    echo "== $lib $extra"
    python bench.py --workload $WL --steps 20 --no-cpu-baseline --lanes 1 $extra | python tools/bench_line.py
  done
done
