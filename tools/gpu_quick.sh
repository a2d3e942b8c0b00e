#!/bin/bash
# Quick GPU visit: resident-engine parity tests, then the per-op sweep (and the headline bench).
# Usage: tools/gpu_quick.sh [sweep-batch] [pytest -k expression]
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_resident.py -x -q --timeout 200 -p no:cacheprovider ${2:+-k "$2"} > gpurun_out/pytest_res.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_res.log)
tail -4 gpurun_out/pytest_res.log | cut -c1-400
if ! grep -q "pytest exit 0" gpurun_out/pytest_res.log; then grep -E "^(FAILED|ERROR|E  )" gpurun_out/pytest_res.log | head -30 | cut -c1-300; exit 1; fi
(timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/bench.log 2>&1; echo "bench exit $?" >> gpurun_out/bench.log)
python - <<PY
import json
for line in open("gpurun_out/bench.log"):
    if line.startswith("{"):
        d = json.loads(line)
        print("BENCH value %.0f img/s  ms/step %.4f  roofline %.3f  e2e %.0f launches %d" % (
            d["value"], d["ms_per_step"], d["roofline"]["frac"], (d.get("e2e") or {}).get("value", 0), d["gpu_launches"]))
    elif "exit" in line or "Error" in line:
        print(line.strip())
PY
(timeout 400 python tools/op_sweep.py --batch ${1:-8192} --iters 10 > gpurun_out/sweep.log 2>&1; echo "sweep exit $?" >> gpurun_out/sweep.log)
tail -1 gpurun_out/sweep.log | cut -c1-300
python - <<PY
import json
try:
    for r in json.load(open("gpurun_out/op_sweep.json")): print("%-45s %8.3f ms %6.2f Mimg/s  %5.1f%%" % (r["case"], r["ms"], r["images_per_s"]/1e6, 100*r["frac_of_measured_peak"]))
except Exception as e: print("no sweep", e)
PY
