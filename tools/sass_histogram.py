#!/usr/bin/env python
"""SASS instruction histogram of the built library (no GPU needed):  python tools/sass_histogram.py [lib.so] > profiles/r02_sass_histogram.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "chambers_b200", "libchambers_aug.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cur, hist = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        hist[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and cur:
        hist[cur][m.group(1)] += 1
print("SASS instruction histogram of %s (cuobjdump -sass, sm_100a), per kernel: total, then the opcodes" % os.path.relpath(lib, ROOT))
print("that show what the kernel is made of (UBLKCP / UTMALDG = TMA bulk / tensor copies, SYNCS = mbarrier, ATOMS / REDUX = shared atomics /")
print("warp reductions; no HMMA / UTC*MMA: nothing on this path is a contraction) and the 25 most frequent opcodes.\n")
KEYS = ("UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "ATOMS", "REDUX", "ATOMG", "RED", "HMMA", "UTC", "LDL", "STL", "MATCH", "SHFL", "BAR",
        "FFMA", "FADD", "PRMT", "LDS", "STS", "LDG", "STG")
for k, c in hist.items():
    if not any(s in k for s in ("resident_kernel", "pass_kernel", "plan_kernel", "normalize", "resize", "optab")):
        continue
    if "pass_kernel" in k and "ILi3E" not in k:
        continue  # one channel count of the tile engine is enough
    short = re.sub(r"_ZN3chb\d+_GLOBAL__N__[0-9a-f_]+?_cu_[0-9a-f]+", "chb::", k)
    key = {}
    for o, n in c.items():
        for p in KEYS:
            if o.startswith(p):
                key[p] = key.get(p, 0) + n
    print("%s: %d instructions" % (short[:120], sum(c.values())))
    print("   " + "  ".join("%s=%d" % kv for kv in sorted(key.items())))
    print("   top: " + "  ".join("%s=%d" % kv for kv in c.most_common(25)))
    print()
