#!/usr/bin/env python
"""Per-image timeline of one call on the image-resident engine (debug build, -DCHB_TIMELINE).

    CHB_LIB=chambers_b200/libchambers_aug_timeline.so python tools/res_timeline.py [--batch 256] [--policy randaugment]

Prints the kernel span, how busy the SMs were, the mean time per op-chain class (so the cost model of
plan_order can be checked against reality), and the images on the critical path.  Not a benchmark.
"""
import argparse
import collections
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from chambers_b200 import _lib  # noqa: E402
from chambers_b200 import augmentations as A  # noqa: E402

NAMES = ["AutoContrast", "Equalize", "Invert", "Brightness", "Contrast", "Color", "Sharpness", "ShearX", "ShearY",
         "TranslateX", "TranslateY", "Posterize", "Solarize", "SolarizeAdd", "CutOut", "Rotate"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--policy", default="randaugment")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "res_timeline.json"))
    args = ap.parse_args()
    B, S = args.batch, 224
    g = torch.Generator(device="cuda").manual_seed(0)
    n = max(2, min(8, (640 << 20) // (2 * B * S * S * 3)))
    bufs = [(torch.randint(0, 256, (B, S, S, 3), dtype=torch.uint8, device="cuda", generator=g),) for _ in range(n)]
    bufs = [(x[0], torch.empty_like(x[0])) for x in bufs]
    pol = A.RandAugment(2, 10, elementwise=True) if args.policy == "randaugment" else A.AutoAugment(elementwise=True)
    layer = pol._transform
    for i in range(400):  # long warm-up: the SM clock has to be up before a single call is looked at
        layer(bufs[i % n][0], seed=0, call_counter=i, out=bufs[i % n][1])
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(100):
        layer(bufs[i % n][0], seed=0, call_counter=i, out=bufs[i % n][1])
    ev1.record()
    torch.cuda.synchronize()
    print("steady state: %.1f us per call over 100 back-to-back calls" % (ev0.elapsed_time(ev1) * 10))
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(0)
        print("SM clock now %d MHz" % pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
    except Exception as e:
        print("no nvml", e)
    lib = _lib.load()
    ctx = _lib.context(0)
    _lib.check(ctx, lib.chb_debug_timeline(ctx, None, 0))
    torch.cuda.synchronize()
    words = 1024 * 2 * 16
    host = np.zeros(words, dtype=np.uint64)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for i in range(50):
        layer(bufs[i % n][0], seed=0, call_counter=i, out=bufs[i % n][1])
    lib.chb_debug_timeline(ctx, host.ctypes.data, words)  # drop what the warm-up recorded (returns the word count)
    e0.record()
    layer(bufs[6 % n][0], seed=0, call_counter=6, out=bufs[6 % n][1], record=True)
    e1.record()
    torch.cuda.synchronize()
    sched = layer.last_schedule
    got = lib.chb_debug_timeline(ctx, host.ctypes.data, words)
    assert got == words, got
    nrec = int(host[0])
    t_entry = int(host[1])
    rec = host[4:4 + 4 * nrec].reshape(nrec, 4).astype(np.int64)
    cta = rec[:, 0] >> 32
    img = rec[:, 0] & 0xFFFFFFFF
    t0, t1, t2 = rec[:, 1] - t_entry, rec[:, 2] - t_entry, rec[:, 3] - t_entry
    span = int(t2.max())
    print("records %d  kernel span (first CTA entry -> last image done) %.1f us   CUDA-event time of the call (with record) %.1f us"
          % (nrec, span / 1e3, e0.elapsed_time(e1) * 1e3))
    print("first image start: min %.1f us, max %.1f us after entry" % (t0.min() / 1e3, np.array([t0[cta == c].min() for c in np.unique(cta)]).max() / 1e3))
    busy = np.array([(t2[cta == c] - t0[cta == c]).sum() for c in np.unique(cta)])
    last = np.array([t2[cta == c].max() for c in np.unique(cta)])
    print("CTAs %d  busy/span mean %.2f  last-finish: min %.1f median %.1f max %.1f us" % (
        len(busy), busy.mean() / span, last.min() / 1e3, np.median(last) / 1e3, last.max() / 1e3))
    dur = (t2 - t0) / 1e3
    book = (t1 - t0) / 1e3
    print("per image: mean %.1f us (load issue + bookkeeping before the first pass %.2f us), sum %.0f us = %.1f us per SM over 148 SMs" % (
        dur.mean(), book.mean(), dur.sum(), dur.sum() / 148))
    cls = collections.defaultdict(list)
    for k in range(nrec):
        ops = [NAMES[o] for o in sched[img[k], :, 0, 0]] if args.policy == "randaugment" else ["sub%d" % sched[img[k], 0, 0, 0]] + \
            ["+".join(str(int(a)) for a in sched[img[k], 0, :, 1])]
        cls["->".join(ops)].append(dur[k])
    rows = sorted(((np.mean(v), len(v), k) for k, v in cls.items()), reverse=True)
    print("slowest chain classes (mean us, count):")
    for m, cnt, k in rows[:25]:
        print("   %6.1f  x%-3d %s" % (m, cnt, k))
    print("fastest:")
    for m, cnt, k in rows[-6:]:
        print("   %6.1f  x%-3d %s" % (m, cnt, k))
    worst = np.argsort(-last)[:5]
    ctas = np.unique(cta)
    print("critical CTAs:")
    for w in worst:
        c = ctas[w]
        sel = np.nonzero(cta == c)[0]
        sel = sel[np.argsort(t0[sel])]
        print("   cta %3d: " % c + "  ".join("[%.1f-%.1f img %d %s]" % (t0[k] / 1e3, t2[k] / 1e3, img[k], "->".join(NAMES[o] for o in sched[img[k], :, 0, 0]) if args.policy == "randaugment" else "") for k in sel))
    json.dump({"span_us": span / 1e3, "records": nrec, "classes": [(float(m), int(cnt), k) for m, cnt, k in rows]}, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
