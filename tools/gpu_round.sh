#!/bin/bash
# One GPU visit: full parity tests, headline bench, per-op sweep, timelines.
# Usage: tools/gpu_round.sh [sweep-batch]
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > gpurun_out/gpu.txt 2>&1
(timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 -p no:cacheprovider > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest.log)
grep -E "^(FAILED|ERROR)|passed|failed|exit" gpurun_out/pytest.log | cut -c1-300 | head -20
if ! grep -q "pytest exit 0" gpurun_out/pytest.log; then
  grep -E "^(FAILED|ERROR)|Error|error" gpurun_out/pytest.log | head -20   # (compute-sanitizer is closed on this pool)
  exit 1
fi
(timeout 400 python bench.py --steps 200 --warmup 10 > gpurun_out/bench.log 2>&1; echo "bench exit $?" >> gpurun_out/bench.log)
python - <<PY
import json
for line in open("gpurun_out/bench.log"):
    if line.startswith("{"):
        d = json.loads(line)
        print("BENCH value %.0f img/s  ms/step %.4f  roofline %.3f  e2e %.0f  cpu %.0f (%d cores) launches %d clocks %s" % (
            d["value"], d["ms_per_step"], d["roofline"]["frac"], (d.get("e2e") or {}).get("value", 0),
            (d.get("cpu_baseline") or {}).get("value", 0), (d.get("cpu_baseline") or {}).get("cores", 0), d["gpu_launches"], d["clocks"]))
    elif "exit" in line or "Error" in line:
        print(line.strip())
PY
(timeout 400 python tools/op_sweep.py --batch ${1:-4096} --iters 10 > gpurun_out/sweep.log 2>&1; echo "sweep exit $?" >> gpurun_out/sweep.log)
tail -1 gpurun_out/sweep.log | cut -c1-300
python - <<PY
import json
try:
    for r in json.load(open("gpurun_out/op_sweep.json")): print("%-45s %8.3f ms %6.2f Mimg/s  %5.1f%%" % (r["case"], r["ms"], r["images_per_s"]/1e6, 100*r["frac_of_measured_peak"]))
except Exception as e: print("no sweep", e)
PY
if [ -f chambers_b200/libchambers_aug_timeline.so ]; then
  for spec in ${TIMELINES:-256,randaugment 4096,Equalize 4096,randaugment}; do
    set -- ${spec//,/ }
    echo "=== timeline B=$1 $2"
    CHB_LIB=$PWD/chambers_b200/libchambers_aug_timeline.so timeout 300 python tools/timeline.py --batch $1 --policy $2 --out gpurun_out/timeline_$1_$2.json 2>&1 | tail -40
  done
fi
