#!/usr/bin/env python
"""Host time to enqueue one policy call (Python layer -> ctypes -> two launches) against the device time of the call."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from chambers_b200 import augmentations as A
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
x = torch.randint(0, 256, (B, 224, 224, 3), dtype=torch.uint8, device="cuda")
bufs = [(x.clone(), torch.empty_like(x)) for _ in range(8)]
layer = A.RandAugment(2, 10, elementwise=True)._transform
for i in range(20):
    layer(bufs[i % 8][0], seed=0, call_counter=i, out=bufs[i % 8][1])
torch.cuda.synchronize()
n = 2000
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
for i in range(n):
    layer(bufs[i % 8][0], seed=0, call_counter=i, out=bufs[i % 8][1])
e1.record()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print("B=%d  host enqueue %.1f us/call   device %.1f us/call   wall %.1f us/call" % (B, 1e6 * (t1 - t0) / n, 1e3 * e0.elapsed_time(e1) / n, 1e6 * (t2 - t0) / n))
