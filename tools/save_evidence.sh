#!/bin/bash
# Copy the artefacts of the last tools/gpu_final.sh run from gpurun_out/ into profiles/ under a version tag:
#   tools/save_evidence.sh r01_v17 [old-tag-to-remove]
tag=$1; old=$2
[ -n "$old" ] && { git rm -q --cached profiles/${old}_* 2>/dev/null; rm -f profiles/${old}_*; }
cp gpurun_out/r01_traffic.json profiles/r01_traffic.json
cp gpurun_out/launches.csv profiles/${tag}_launches_b256.csv
grep "^{" gpurun_out/bench.log > profiles/${tag}_bench.json
grep "^{" gpurun_out/bench_autoaugment_4096.log > profiles/${tag}_bench_autoaugment_b4096.json
grep "^{" gpurun_out/bench_reference.log > profiles/${tag}_bench_reference_arm.json
cp gpurun_out/op_sweep_8192.json profiles/${tag}_op_sweep_b8192.json
cp gpurun_out/op_sweep.json profiles/${tag}_op_sweep_b4096.json
for op in Rotate Invert Equalize Sharpness; do python tools/ncu_summary.py gpurun_out/prof_op_$op.ncu-rep --sass 20 > profiles/${tag}_ncu_op_$op.txt 2>&1; done
python tools/ncu_summary.py gpurun_out/prof_mixed.ncu-rep --sass 25 > profiles/${tag}_ncu_randaugment_b256.txt 2>&1
for f in gpurun_out/timeline_*.txt; do cp $f profiles/${tag}_$(basename $f); done
python - <<PY
import json
d=json.loads(open("profiles/${tag}_bench.json").read())
print("bench", d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["traffic"], d["e2e"]["value"], d["cpu_baseline"]["value"], d["gpu_launches"], d["clocks"])
d=json.loads(open("profiles/${tag}_bench_autoaugment_b4096.json").read())
print("auto4096", d["value"], d["ms_per_step"], d["roofline"]["frac"])
d=json.loads(open("profiles/${tag}_bench_reference_arm.json").read())
print("ref", d["value"], d["cpu_baseline"]["cores"])
for r in json.load(open("profiles/${tag}_op_sweep_b8192.json")): print("%-45s B=%5d %8.3f ms %6.2f Mimg/s  %5.1f%%" % (r["case"], r["batch"], r["ms"], r["images_per_s"]/1e6, 100*r["frac_of_measured_peak"]))
PY
