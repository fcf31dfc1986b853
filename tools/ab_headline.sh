#!/bin/bash
# same-box A/B of an environment switch on the headline line (3 lanes): tools/ab_headline.sh VAR
for rep in 1 2 3; do for e in 1 0; do
  echo -n "$1=$e  "; env $1=$e python bench.py --steps 50 --no-cpu-baseline --no-secondary 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%.4f ms/step  %.4e lists/s  kernel %.4f' % (d['ms_per_step'], d['value'], d['roofline']['kernel_ms']))"
done; done
