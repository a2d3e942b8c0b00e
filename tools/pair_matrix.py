#!/usr/bin/env python
"""Time every ordered pair of the 16 RandAugment ops (replayed schedule, every image the same pair).
    python tools/pair_matrix.py [--batch 2048] [--magnitude 10]"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from chambers_b200 import build
build.build_library()
from chambers_b200 import augmentations as A

NAMES = ["AutoContrast", "Equalize", "Invert", "Brightness", "Contrast", "Color", "Sharpness", "ShearX", "ShearY",
         "TranslateX", "TranslateY", "Posterize", "Solarize", "SolarizeAdd", "CutOut", "Rotate"]

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=2048)
    ap.add_argument("--magnitude", type=float, default=10)
    ap.add_argument("--size", type=int, default=224)
    a = ap.parse_args()
    B, S = a.batch, a.size
    g = torch.Generator(device="cuda").manual_seed(0)
    xs = [torch.randint(0, 256, (B, S, S, 3), dtype=torch.uint8, device="cuda", generator=g) for _ in range(2)]
    ys = [torch.empty_like(x) for x in xs]
    layer = A.RandAugment(2, a.magnitude, elementwise=True)._transform
    M = np.zeros((16, 16))
    for i in range(16):
        for j in range(16):
            sch = np.zeros((B, 2, 1, 5), np.int32)
            sch[:, 0, 0, 0], sch[:, 1, 0, 0] = i, j
            sch[..., 1] = 1; sch[..., 3] = 100; sch[..., 4] = 120
            rep = torch.from_numpy(sch).cuda()
            for k in range(2): layer(xs[k], seed=0, call_counter=k, out=ys[k], replay=rep)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for k in range(4): layer(xs[k & 1], seed=0, call_counter=k, out=ys[k & 1], replay=rep)
            e1.record(); torch.cuda.synchronize()
            M[i, j] = e0.elapsed_time(e1) / 4
    ideal = 2.0 * B * S * S * 3 / 6549.8e9 * 1e3
    print("ms per call, batch %d; 100%% of the measured copy peak = %.3f ms; mean %.3f ms" % (B, ideal, M.mean()))
    print("%-13s" % "first\\second" + "".join("%7s" % n[:6] for n in NAMES))
    for i in range(16):
        print("%-13s" % NAMES[i] + "".join("%7.2f" % M[i, j] for j in range(16)))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump({"batch": B, "size": S, "magnitude": a.magnitude, "names": NAMES, "ms": M.tolist(), "ideal_ms": ideal},
              open(os.path.join(ROOT, "gpurun_out", "pair_matrix.json"), "w"))

if __name__ == "__main__":
    main()
