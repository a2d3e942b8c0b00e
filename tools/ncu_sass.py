#!/usr/bin/env python
"""Dump SASS rows [a, b) of an .ncu-rep with executed counts and stall samples:  python tools/ncu_sass.py rep a b"""
import csv, io, subprocess, sys
rep, a, b = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
while rows and (len(rows[0]) < 6 or rows[0][0] != "Address"):
    rows.pop(0)
h = rows[0]
body = [r for r in rows[1:] if len(r) == len(h)]
ci = h.index("Source"); ce = h.index("Instructions Executed")
cs = [i for i, k in enumerate(h) if k.strip().startswith("Warp Stall Sampling (All")][0]
if a < 0:  # histogram of executed counts
    import collections
    cnt = collections.Counter(r[ce] for r in body)
    for k, v in sorted(cnt.items(), key=lambda kv: -kv[1])[:25]:
        print(k, v)
    sys.exit(0)
for i in range(a, min(b, len(body))):
    r = body[i]
    print("%5d %9s %5s  %s" % (i, r[ce], r[cs], r[ci][:110]))
