#!/bin/bash
# Host-path (chb_policy_apply_host) throughput for a few chunk sizes / stream counts.
for st in 3 4 6; do for kb in 1200 2560 5000 10000; do
  CHB_E2E_STREAMS=$st CHB_E2E_CHUNK_KB=$kb python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('streams $st chunk_kb $kb  e2e %.0f img/s  device %.4f ms/step' % (d['e2e']['value'], d['ms_per_step']))
"
done; done
