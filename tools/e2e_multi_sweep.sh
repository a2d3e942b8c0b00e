#!/bin/bash
# e2e (host-buffer) throughput at 8 GPUs for several chunk sizes / stream counts of chb_policy_apply_host.
mkdir -p gpurun_out
for cfg in "2560 4" "9700 4" "19300 2" "38600 1" "9700 2"; do
  set -- $cfg
  (CHB_E2E_CHUNK_KB=$1 CHB_E2E_STREAMS=$2 timeout 200 python bench.py --gpus ${N:-8} --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/e2e_$1_$2.log 2>&1)
  python - $1 $2 <<'PY'
import json, sys
for line in open("gpurun_out/e2e_%s_%s.log" % (sys.argv[1], sys.argv[2])):
    if line.startswith("{"):
        d = json.loads(line)
        print("chunk %s KiB, %s streams: e2e %.0f img/s = %.1f GB/s each way over %d GPUs" % (sys.argv[1], sys.argv[2], d["e2e"]["value"], d["e2e"]["value"] * 150528 / 1e9, d["n_gpus"]))
PY
done
