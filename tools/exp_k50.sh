#!/bin/bash
# scoring pass at config-3 shapes per library variant: kernel time from an ncu launch list
export PS_B=16 PS_K=50 PS_R=50000
for lib in default "$@"; do
  if [ "$lib" = default ]; then unset PLDEPTH_B200_LIB; else export PLDEPTH_B200_LIB=$PWD/pldepth_b200/variants/$lib.so; fi
  for st in information thresholded; do
    ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/exp_k50_${lib}_${st}.csv python tools/profile_scored_step.py $st no-emit > gpurun_out/exp_ncu.log 2>&1
    echo "== $lib $st"; python tools/ncu_summary.py gpurun_out/exp_k50_${lib}_${st}.csv 2>/dev/null | grep -E "score_reg|lists_tab_kernel<8, 8, 256, 0, 1>"
  done
done
