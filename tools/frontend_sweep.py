#!/usr/bin/env python
"""Throughput of the front-end layers on one B200: ImageNetNormalization (1 B read + 4 B written per value) and
the policy -> normalisation pipeline (separate kernels)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from chambers_b200 import augmentations as A

def peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0

def timeit(fn, iters=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randint(0, 256, (B, 224, 224, 3), dtype=torch.uint8, device="cuda", generator=g)
n = x.numel()
rows = []
for mode in ("tf", "torch", "caffe"):
    layer = A.ImageNetNormalization(mode=mode)
    ms = timeit(lambda: layer(x))
    gbs = 5.0 * n / (ms * 1e-3) / 1e9
    rows.append({"case": "ImageNetNormalization(%s) B=%d" % (mode, B), "ms": ms, "GBs": gbs, "frac_of_measured_peak": gbs / peak()})
ra = A.RandAugment(2, 10, elementwise=True)
norm = A.ImageNetNormalization(mode="tf")
cnt = [0]
def pipeline():
    cnt[0] += 1
    return norm(ra(x, training=True, seed=0, call_counter=cnt[0]))
ms = timeit(pipeline)
rows.append({"case": "RandAugment(2,10) -> ImageNetNormalization(tf), two kernels, B=%d" % B, "ms": ms, "images_per_s": B / (ms * 1e-3)})
def fused():
    cnt[0] += 1
    return ra(x, training=True, seed=0, call_counter=cnt[0], normalize="tf")
ms = timeit(fused)
rows.append({"case": "RandAugment(2,10, normalize='tf'): normalisation as the write epilogue of the policy's last pass, B=%d" % B, "ms": ms,
             "images_per_s": B / (ms * 1e-3)})
ms = timeit(lambda: ra(x, training=True, seed=0, call_counter=1))
rows.append({"case": "RandAugment(2,10) alone, B=%d" % B, "ms": ms, "images_per_s": B / (ms * 1e-3)})
for r in rows: print(json.dumps(r))
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "frontend_sweep.json"), "w"), indent=1)
