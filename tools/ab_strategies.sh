#!/bin/bash
# A/B of library variants (pldepth_b200/variants/*.so via PLDEPTH_B200_LIB) on the scored steps without rankings
for lib in default "$@"; do
  if [ "$lib" = default ]; then unset PLDEPTH_B200_LIB; else export PLDEPTH_B200_LIB=$PWD/pldepth_b200/variants/$lib.so; fi
  echo "== $lib"
  python tools/bench_strategies.py --no-emit --graph --steps 20 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('  ', d['strategy'], '%.4f ms' % d['ms_per_step'])"
done
