#!/bin/bash
# Round-2 evidence of the resident engine on ONE GPU (one gpurun call, every ncu run after a plain run of the same command):
#   1. launch list of the default bench command          -> gpurun_out/r02_launches_b256.csv
#   2. DRAM bytes of the bench's kernels (traffic)       -> gpurun_out/r02_traffic.json
#   3. ncu --set full of the bench kernel (B = 256 mix)  -> gpurun_out/prof_bench256.ncu-rep
#   4. ncu --set full per op class at 2048 images        -> gpurun_out/prof_op_<Op>.ncu-rep
mkdir -p gpurun_out
BENCH="python bench.py --steps 6 --warmup 4 --no-cpu-baseline --no-e2e"
$BENCH > gpurun_out/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_launches_b256.csv $BENCH > gpurun_out/ncu_bench.log 2>&1
$BENCH > gpurun_out/plain_bench2.log 2>&1 &&
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"resident_kernel|plan_kernel|pass_kernel" \
    --csv --log-file gpurun_out/traffic.csv $BENCH > gpurun_out/ncu_traffic.log 2>&1
python - <<'PY'
import csv, json, collections
rows = [r for r in csv.reader(open("gpurun_out/traffic.csv")) if len(r) > 10]
hdr = rows[0]; body = rows[1:]
kn, mn, mu, mv = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value")
idc = hdr.index("ID")
per = collections.OrderedDict()
for r in body:
    d = per.setdefault(r[idc], {"kernel": r[kn]})
    v = float(r[mv].replace(",", ""))
    u = r[mu].lower()
    if "byte" in u:
        v *= {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
    d[r[mn]] = v
launches = list(per.values())
steps = len(launches)
rd = sum(l.get("dram__bytes_read.sum", 0) for l in launches); wr = sum(l.get("dram__bytes_write.sum", 0) for l in launches)
out = {"randaugment_b256": {"dram_bytes_per_step": (rd + wr) / steps, "read": rd / steps, "write": wr / steps, "steps_captured": steps,
                            "algorithmic_bytes_per_step": 2 * 256 * 224 * 224 * 3,
                            "kernel": launches[0]["kernel"][:80] if launches else None,
                            "how": "ncu dram__bytes_{read,write}.sum per resident_kernel launch of bench.py --steps 6 --warmup 4 (one launch per step)"}}
json.dump(out, open("gpurun_out/r02_traffic.json", "w"), indent=1)
print(json.dumps(out))
PY
$BENCH > gpurun_out/plain_bench3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:resident_kernel -s 6 -c 1 -f -o gpurun_out/prof_bench256 $BENCH > gpurun_out/ncu_bench256.log 2>&1
for op in ${OPS:-Equalize Rotate Sharpness Brightness TranslateX RandAugment}; do
  python tools/op_sweep.py --only $op --batch 2048 --iters 3 > gpurun_out/plain_$op.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:resident_kernel -s 2 -c 1 -f -o gpurun_out/prof_op_$op \
      python tools/op_sweep.py --only $op --batch 2048 --iters 3 > gpurun_out/ncu_$op.log 2>&1
  tail -1 gpurun_out/plain_$op.log | cut -c1-160
done
ls -la gpurun_out/*.ncu-rep | tail -8
