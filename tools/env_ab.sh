#!/bin/bash
# A/B of an environment switch of the library on one box: tools/env_ab.sh VAR "0 1" "<op names>" [batch]
B=${4:-4096}
for val in $2; do
  for op in $3; do
    env $1=$val python tools/op_sweep.py --only $op --batch $B --iters 20 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('$1=$val %-16s B=%d %8.4f ms  %5.1f%%' % (d['case'], d['batch'], d['ms'], 100 * d['frac_of_measured_peak']))
"
  done
  env $1=$val python bench.py --steps 300 --warmup 20 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('$1=$val bench B=256 %.4f ms' % d['ms_per_step'])
"
done
