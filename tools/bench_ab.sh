for v in ${VARIANTS:-base}; do
  lib=$PWD/chambers_b200/libchambers_aug.so; [ "$v" != base ] && lib=$PWD/chambers_b200/libchambers_aug_$v.so
  CHB_LIB=$lib python bench.py --steps 300 --warmup 20 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('$v bench B=256 %.4f ms  frac %.3f' % (d['ms_per_step'], d['roofline']['frac']))
"
  CHB_LIB=$lib python bench.py --policy autoaugment --steps 300 --warmup 20 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('$v autoaugment B=256 %.4f ms' % (d['ms_per_step']))
"
done
