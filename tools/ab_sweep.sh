#!/bin/bash
# A/B on ONE box: the full per-op sweep for several builds of the library (CHB_LIB), side by side.
# Usage: tools/ab_sweep.sh <batch> <name=path> [<name=path> ...]      (path relative to the repo root)
mkdir -p gpurun_out
B=$1; shift
names=()
for spec in "$@"; do
  name=${spec%%=*}; path=${spec#*=}
  names+=($name)
  if [ -n "$AB_TEST" ]; then
    (CHB_LIB=$PWD/$path timeout 600 python -m pytest tests/test_gpu_resident.py -x -q --timeout 200 -p no:cacheprovider > gpurun_out/pytest_$name.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_$name.log)
    tail -3 gpurun_out/pytest_$name.log | cut -c1-300
    grep -E "^(FAILED|ERROR|E  )" gpurun_out/pytest_$name.log | head -20 | cut -c1-300
  fi
  (CHB_LIB=$PWD/$path timeout 400 python tools/op_sweep.py --batch $B --iters 10 --out gpurun_out/op_sweep_$name.json > gpurun_out/sweep_$name.log 2>&1; echo "sweep $name exit $?")
done
python - "${names[@]}" <<'PY'
import json, sys
names = sys.argv[1:]
tabs = {}
for n in names:
    try:
        tabs[n] = {r["case"]: r for r in json.load(open("gpurun_out/op_sweep_%s.json" % n))}
    except Exception as e:
        print("no sweep for", n, e)
        tabs[n] = {}
cases = []
for n in names:
    for c in tabs[n]:
        if c not in cases:
            cases.append(c)
print("%-45s" % "case (ms, % of measured HBM peak)" + "".join("%22s" % n for n in names))
for c in cases:
    print("%-45s" % c + "".join(("%14.3f %6.1f%%" % (tabs[n][c]["ms"], 100 * tabs[n][c]["frac_of_measured_peak"])) if c in tabs[n] else "%22s" % "-" for n in names))
PY
