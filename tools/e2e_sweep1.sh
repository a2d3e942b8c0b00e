#!/bin/bash
# e2e (host-buffer) throughput at 1 GPU for several chunk sizes / stream counts of chb_policy_apply_host.
for cfg in "2560 4" "1280 4" "1280 8" "640 8" "3840 4" "5120 3" "2560 6"; do
  set -- $cfg
  echo -n "chunk $1 KiB, $2 streams: "
  CHB_E2E_CHUNK_KB=$1 CHB_E2E_STREAMS=$2 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.readline()); print('e2e %.0f img/s = %.1f GB/s each way' % (d['e2e']['value'], d['e2e']['value'] * 150528 / 1e9))"
done
