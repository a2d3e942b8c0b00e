#!/bin/bash
# DRAM traffic of one bench step: ncu dram bytes of every kernel of a few steps of the default bench
# command (after a plain run), summed per step -> profiles/r01_traffic.json (read by bench.py).
mkdir -p gpurun_out
python bench.py --steps 6 --warmup 4 --no-cpu-baseline --no-e2e > gpurun_out/plain_traffic.log 2>&1 &&
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"plan_kernel|pass_kernel" \
    --csv --log-file gpurun_out/traffic.csv python bench.py --steps 6 --warmup 4 --no-cpu-baseline --no-e2e > gpurun_out/ncu_traffic.log 2>&1
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
steps = sum(1 for l in launches if "plan_kernel" in l["kernel"])
tot = sum(l.get("dram__bytes_read.sum", 0) + l.get("dram__bytes_write.sum", 0) for l in launches)
rd = sum(l.get("dram__bytes_read.sum", 0) for l in launches); wr = sum(l.get("dram__bytes_write.sum", 0) for l in launches)
out = {"randaugment_b256": {"dram_bytes_per_step": tot / steps, "read": rd / steps, "write": wr / steps, "steps_captured": steps,
                            "algorithmic_bytes_per_step": 2 * 256 * 224 * 224 * 3, "how": "ncu dram__bytes_{read,write}.sum over plan_kernel + pass_kernel launches of bench.py --steps 6 --warmup 4"}}
json.dump(out, open("gpurun_out/r01_traffic.json", "w"), indent=1)
print(json.dumps(out))
PY
