#!/bin/bash
for m in 8 16 4 3 2; do
  PLD_GRID_MULT=$m ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/exp_gm${m}.csv python tools/profile_scored_step.py information no-emit > gpurun_out/exp_ncu.log 2>&1
  echo "== PLD_GRID_MULT=$m"; python tools/ncu_summary.py gpurun_out/exp_gm${m}.csv 2>/dev/null | sed -n 2,3p
done
