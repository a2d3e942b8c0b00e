#!/usr/bin/env python
"""Per-op and per-config throughput sweep on one B200 (BASELINE.json configs[2..4]).

    python tools/op_sweep.py [--batch 8192] [--iters 20] [--out gpurun_out/op_sweep.json]

Every line: images/s, achieved algorithmic GB/s (2*H*W*C bytes per image) and its fraction of
the measured HBM copy peak.  Inputs are resident in HBM; buffers are far larger than L2 at the
default batch, smaller configs rotate through a pool.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from chambers_b200 import build  # noqa: E402

build.build_library()
from chambers_b200 import augmentations as A  # noqa: E402


def peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


def time_layer(fn, bufs, iters, warm=3):
    n = len(bufs)
    for i in range(warm):
        fn(i, *bufs[i % n])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(warm + i, *bufs[(warm + i) % n])
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def make_bufs(B, H, W, C, pool_bytes=768 << 20, kind="uniform"):
    n = max(2, min(8, pool_bytes // (2 * B * H * W * C)))
    g = torch.Generator(device="cuda").manual_seed(0)
    bufs = []
    for _ in range(n):
        if kind == "uniform":
            x = torch.randint(0, 256, (B, H, W, C), dtype=torch.uint8, device="cuda", generator=g)
        else:
            x = torch.full((B, H, W, C), 77, dtype=torch.uint8, device="cuda")
        bufs.append((x, torch.empty_like(x)))
    return bufs


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8192)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "op_sweep.json"))
    ap.add_argument("--only", default=None, help="run just this op (e.g. Rotate) -- for ncu captures")
    args = ap.parse_args()
    pk = peak()
    rows = []

    def report(name, B, H, W, ms, extra=None):
        gbs = 2.0 * B * H * W * 3 / (ms * 1e-3) / 1e9
        row = {"case": name, "batch": B, "image": [H, W, 3], "ms": ms, "images_per_s": B / (ms * 1e-3),
               "GBs": gbs, "frac_of_measured_peak": gbs / pk}
        if extra:
            row.update(extra)
        rows.append(row)
        print(json.dumps(row), flush=True)

    B = args.batch
    bufs = make_bufs(B, 224, 224, 3)
    from oracle.policy import magnitude_kwargs  # parameters only (no pixels): same table as the layers use
    names = ["AutoContrast", "Equalize", "Invert", "Brightness", "Contrast", "Color", "Sharpness", "ShearX", "ShearY",
             "TranslateX", "TranslateY", "Posterize", "Solarize", "SolarizeAdd", "CutOut", "Rotate"]
    if args.only == "RandAugment":
        layer = A.RandAugment(2, 10, elementwise=True)._transform
        ms = time_layer(lambda i, x, y: layer(x, seed=0, call_counter=i, out=y), bufs, args.iters, warm=1)
        report("RandAugment(2,10)", B, 224, 224, ms)
        return
    if args.only == "config3":  # BASELINE.json configs[3]
        b = make_bufs(512, 512, 512, 3, pool_bytes=2 << 30)
        ra = A.RandAugment(3, 15, elementwise=True)._transform
        ms = time_layer(lambda i, x, y: ra(x, seed=0, call_counter=i, out=y), b, args.iters, warm=2)
        report("config: RandAugment(3,15) 512x512 B=512", 512, 512, 512, ms)
        return
    if args.only in ("EqualizeConst", "AutoContrastConst"):  # adversarial histogram input: one value everywhere
        cb = make_bufs(B, 224, 224, 3, kind="constant")
        layer = A.RandomChoice([getattr(A, args.only[:-5])()], 1)
        ms = time_layer(lambda i, x, y: layer(x, seed=0, call_counter=i, out=y), cb, args.iters, warm=1)
        report("op:%s(constant image)" % args.only[:-5], B, 224, 224, ms)
        return
    if args.only == "identity":
        layer = A.RandomChoice([A.RandomChance(A.Invert(), 0.0)], 1)
        ms = time_layer(lambda i, x, y: layer(x, seed=0, call_counter=i, out=y), bufs, args.iters, warm=1)
        report("identity(copy)", B, 224, 224, ms)
        return
    if args.only:
        layer = A.RandomChoice([getattr(A, args.only)(**magnitude_kwargs(args.only, 10))], 1)
        ms = time_layer(lambda i, x, y: layer(x, seed=0, call_counter=i, out=y), bufs, args.iters, warm=1)
        report("op:" + args.only, B, 224, 224, ms)
        return
    ident = A.RandomChoice([A.RandomChance(A.Invert(), 0.0)], 1)
    ms = time_layer(lambda i, x, y: ident(x, seed=0, call_counter=i, out=y), bufs, args.iters)
    report("identity(copy)", B, 224, 224, ms)
    for n in names:
        layer = A.RandomChoice([getattr(A, n)(**magnitude_kwargs(n, 10))], 1)
        ms = time_layer(lambda i, x, y: layer(x, seed=0, call_counter=i, out=y), bufs, args.iters)
        report("op:" + n, B, 224, 224, ms)
    # representative ordered pairs (replayed schedule: every image gets the same two ops)
    import numpy as np
    ra2 = A.RandAugment(2, 10, elementwise=True)._transform
    for a, b in [("Rotate", "ShearX"), ("CutOut", "Rotate"), ("Rotate", "Sharpness"), ("Sharpness", "Rotate"),
                 ("Rotate", "Equalize"), ("Equalize", "Rotate"), ("Color", "Rotate"), ("Brightness", "Rotate"),
                 ("Brightness", "Color"), ("Equalize", "Sharpness"), ("Color", "Color")]:
        sch = np.zeros((B, 2, 1, 5), np.int32)
        sch[:, 0, 0, 0], sch[:, 1, 0, 0] = names.index(a), names.index(b)
        sch[..., 1] = 1
        sch[..., 3] = 100
        sch[..., 4] = 120
        rep = torch.from_numpy(sch).cuda()
        ms = time_layer(lambda i, x, y: ra2(x, seed=0, call_counter=i, out=y, replay=rep), bufs, args.iters)
        report("pair:%s->%s" % (a, b), B, 224, 224, ms)
    cbufs = make_bufs(B, 224, 224, 3, kind="constant")
    for n in ("Equalize", "AutoContrast"):
        layer = A.RandomChoice([getattr(A, n)()], 1)
        ms = time_layer(lambda i, x, y: layer(x, seed=0, call_counter=i, out=y), cbufs, args.iters)
        report("op:%s(constant image)" % n, B, 224, 224, ms)
    del cbufs
    # natural images (SURVEY.md 8d distribution B): the reference's sample_data batch tiled to B -- peaked histograms
    try:
        import numpy as np
        nat = np.load(os.path.join(ROOT, "tests", "golden", "c1_batch.npz"))["images"]
        reps = (B + nat.shape[0] - 1) // nat.shape[0]
        xn = torch.from_numpy(np.tile(nat, (reps, 1, 1, 1))[:B]).cuda()
        nb = max(2, min(8, (768 << 20) // (2 * xn.numel())))
        nbufs = [(xn.roll(k, 0).contiguous(), torch.empty_like(xn)) for k in range(nb)]
        for n in ("Equalize", "AutoContrast", "Sharpness", "Rotate"):
            layer = A.RandomChoice([getattr(A, n)(**magnitude_kwargs(n, 10))], 1)
            ms = time_layer(lambda i, x, y: layer(x, seed=0, call_counter=i, out=y), nbufs, args.iters)
            report("op:%s(natural images)" % n, B, 224, 224, ms)
        ran = A.RandAugment(2, 10, elementwise=True)._transform
        ms = time_layer(lambda i, x, y: ran(x, seed=0, call_counter=i, out=y), nbufs, args.iters)
        report("RandAugment(2,10) elementwise=True (natural images)", B, 224, 224, ms)
        del nbufs, xn
    except Exception as e:  # the fixture is optional for a throughput sweep
        print("natural-image cases skipped: %r" % (e,))
    for ew in (True, False):
        ra = A.RandAugment(2, 10, elementwise=ew)._transform
        ms = time_layer(lambda i, x, y: ra(x, seed=0, call_counter=i, out=y), bufs, args.iters)
        report("RandAugment(2,10) elementwise=%s" % ew, B, 224, 224, ms)
    del bufs
    torch.cuda.empty_cache()
    for Bc in (256, 4096):
        b = make_bufs(Bc, 224, 224, 3)
        ra = A.RandAugment(2, 10, elementwise=True)._transform
        ms = time_layer(lambda i, x, y: ra(x, seed=0, call_counter=i, out=y), b, max(args.iters, 50))
        report("config: RandAugment(2,10) B=%d" % Bc, Bc, 224, 224, ms, {"pool": len(b)})
        aa = A.AutoAugment(elementwise=True)._transform
        ms = time_layer(lambda i, x, y: aa(x, seed=0, call_counter=i, out=y), b, max(args.iters, 50))
        report("config: AutoAugment B=%d" % Bc, Bc, 224, 224, ms, {"pool": len(b)})
        del b
    torch.cuda.empty_cache()
    b = make_bufs(512, 512, 512, 3, pool_bytes=2 << 30)
    ra = A.RandAugment(3, 15, elementwise=True)._transform
    ms = time_layer(lambda i, x, y: ra(x, seed=0, call_counter=i, out=y), b, args.iters)
    report("config: RandAugment(3,15) 512x512 B=512", 512, 512, 512, ms, {"pool": len(b)})
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(rows, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
