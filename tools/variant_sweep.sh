#!/bin/bash
# A/B of library variants on the same box: tools/variant_sweep.sh "<variant names>" "<op names>" [batch]
mkdir -p gpurun_out
B=${3:-4096}
for v in base $1; do
  lib=$PWD/chambers_b200/libchambers_aug.so
  [ "$v" != base ] && lib=$PWD/chambers_b200/libchambers_aug_$v.so
  for op in $2; do
    CHB_LIB=$lib python tools/op_sweep.py --only $op --batch $B --iters 20 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('%-8s %-14s B=%d %8.4f ms  %5.1f%%' % ('$v', d['case'], d['batch'], d['ms'], 100 * d['frac_of_measured_peak']))
"
  done
done
