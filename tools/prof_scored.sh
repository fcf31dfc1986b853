for st in thresholded information; do for em in emit no-emit; do
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_scored_${st}_${em}.csv python tools/profile_scored_step.py $st $em > gpurun_out/ncu_scored.log 2>&1
done; done
