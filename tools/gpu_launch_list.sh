#!/bin/bash
# ncu launch list (per-launch device time) of the default bench command, after a plain run.
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/ncu_bench.log 2>&1
tail -1 gpurun_out/plain_bench.log | cut -c1-300
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/launches.csv")) if len(r) > 10]
hdr = rows[0]; body = rows[1:]
kn = hdr.index("Kernel Name"); mv = hdr.index("Metric Value")
agg = collections.defaultdict(list)
for r in body:
    try: agg[r[kn][:60]].append(float(r[mv].replace(",", "")))
    except ValueError: pass
for k, v in agg.items(): print("%-62s n=%4d  mean %9.1f ns  total %10.1f us" % (k, len(v), sum(v) / len(v), sum(v) / 1e3))
# last 12 launches in order
for r in body[-12:]: print(r[kn][:50], r[mv])
PY
