mkdir -p gpurun_out
(timeout 400 python tools/op_sweep.py --batch ${1:-4096} --iters 10 > gpurun_out/sweep.log 2>&1; echo "sweep exit $?" >> gpurun_out/sweep.log)
tail -2 gpurun_out/sweep.log | cut -c1-300
python - <<PY
import json
try:
    for r in json.load(open("gpurun_out/op_sweep.json")): print("%-45s %8.3f ms %6.2f Mimg/s  %5.1f%%" % (r["case"], r["ms"], r["images_per_s"]/1e6, 100*r["frac_of_measured_peak"]))
except Exception as e: print("no sweep", e)
PY
