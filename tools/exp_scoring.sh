#!/bin/bash
# scoring-pass experiments: ncu launch list of one scored step per library variant
for lib in default "$@"; do
  if [ "$lib" = default ]; then unset PLDEPTH_B200_LIB; else export PLDEPTH_B200_LIB=$PWD/pldepth_b200/variants/$lib.so; fi
  for st in information thresholded; do
    ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/exp_${lib}_${st}.csv python tools/profile_scored_step.py $st no-emit > gpurun_out/exp_ncu.log 2>&1
    echo "== $lib $st"; python tools/ncu_summary.py gpurun_out/exp_${lib}_${st}.csv | head -5
  done
done
