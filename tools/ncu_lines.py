#!/usr/bin/env python
"""Static / executed SASS instruction counts per source function region of an .ncu-rep (needs -lineinfo + --import-source on).
   python tools/ncu_lines.py rep"""
import csv, io, re, subprocess, sys, collections
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur_file = None
stat = collections.defaultdict(lambda: [0, 0.0, 0.0])   # (file, line) -> [static, executed, samples]
src = collections.defaultdict(dict)
hdr = None
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]; hdr = None; continue
    if len(r) > 6 and r[0] == "Line No":
        hdr = r; ia = r.index("Address"); ie = r.index("Instructions Executed")
        isamp = [i for i, k in enumerate(r) if k.startswith("Warp Stall Sampling (All")][0]
        continue
    if hdr is None or len(r) != len(hdr):
        continue
    line = int(r[0]) if r[0].isdigit() else None
    if line is None: continue
    if r[1]: src[cur_file][line] = r[1]
    if r[ia]:
        s = stat[(cur_file, line)]
        s[0] += 1
        try: s[1] += float(r[ie].replace(",", "") or 0)
        except ValueError: pass
        try: s[2] += float(r[isamp].replace(",", "") or 0)
        except ValueError: pass
# function regions: lines starting with "__device__" / "template" followed by name -- use a simple regex on the embedded source
regions = collections.defaultdict(lambda: [0, 0.0, 0.0])
for f in src:
    names = {}
    cur = "?"
    for ln in sorted(src[f]):
        t = src[f][ln]
        m = re.match(r"^(?:__device__|__global__|static|template|inline).*?([A-Za-z_0-9]+)\s*\(", t)
        if t.startswith("__device__") or t.startswith("__global__") or t.startswith("pass_kernel"):
            m2 = re.search(r"([A-Za-z_0-9]+)\s*\(", t)
            if m2: cur = m2.group(1)
        names[ln] = cur
    for (ff, ln), s in stat.items():
        if ff != f: continue
        # nearest known line
        k = ln
        while k not in names and k > 0: k -= 1
        fn = names.get(k, "?")
        a = regions[(f, fn)]
        a[0] += s[0]; a[1] += s[1]; a[2] += s[2]
tot = [sum(v[i] for v in regions.values()) for i in range(3)]
print("total static %d executed %.0f samples %.0f" % tuple(tot))
for (f, fn), v in sorted(regions.items(), key=lambda kv: -kv[1][0])[:40]:
    print("%-22s %-28s static %6d (%4.1f%%)  executed %5.1f%%  samples %5.1f%%" % (f, fn, v[0], 100 * v[0] / tot[0], 100 * v[1] / max(tot[1], 1), 100 * v[2] / max(tot[2], 1)))
