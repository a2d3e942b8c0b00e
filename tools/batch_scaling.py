#!/usr/bin/env python
"""Device time of one RandAugment(2,10) call as a function of the batch size (fixed per-call cost)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from chambers_b200 import build
build.build_library()
from chambers_b200 import augmentations as A
for name, layer in (("RandAugment(2,10)", A.RandAugment(2, 10, elementwise=True)._transform), ("Invert", A.RandomChoice([A.Invert()], 1))):
    for B in (1, 16, 64, 256, 1024, 4096):
        g = torch.Generator(device="cuda").manual_seed(0)
        n = max(2, min(16, (512 << 20) // (B * 150528)))
        xs = [torch.randint(0, 256, (B, 224, 224, 3), dtype=torch.uint8, device="cuda", generator=g) for _ in range(n)]
        ys = [torch.empty_like(x) for x in xs]
        for k in range(5): layer(xs[k % n], seed=0, call_counter=k, out=ys[k % n])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        it = 100
        import time
        e0.record()
        t0 = time.perf_counter()
        for k in range(it): layer(xs[k % n], seed=0, call_counter=k, out=ys[k % n])
        t1 = time.perf_counter()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / it
        print("%-18s B=%5d  %8.1f us/call  %7.3f us/image   host issue %6.1f us/call" % (name, B, ms * 1e3, ms * 1e3 / B, (t1 - t0) / it * 1e6), flush=True)
