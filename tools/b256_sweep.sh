#!/bin/bash
# B=256 headline under several split thresholds (one box).
for pct in 0 60 80 100 120 150; do
  if [ $pct = 0 ]; then export CHB_SPLIT=0; else export CHB_SPLIT=1; export CHB_SPLIT_PCT=$pct; fi
  python bench.py --steps 300 --warmup 20 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('split_pct $pct: ms/step %.4f frac %.3f' % (d['ms_per_step'], d['roofline']['frac']))"
done
for pol in autoaugment; do
for pct in 0 90; do
  if [ $pct = 0 ]; then export CHB_SPLIT=0; else export CHB_SPLIT=1; export CHB_SPLIT_PCT=$pct; fi
  python bench.py --steps 300 --warmup 20 --no-cpu-baseline --no-e2e --policy $pol 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('$pol split_pct $pct: ms/step %.4f frac %.3f' % (d['ms_per_step'], d['roofline']['frac']))"
done; done
