#!/usr/bin/env python
"""Per-CTA timeline of one policy call (debug build of the library, -DCHB_TIMELINE).

    CHB_LIB=chambers_b200/libchambers_aug_timeline.so python tools/timeline.py [--batch 256] [--policy randaugment|autoaugment|<OpName>]

Prints, per pass level: when the first / last CTA entered, started its first tile, finished its
last tile and left (ns after the first entry of level 0), and where producer and consumer cycles
went (waiting on each other, planning, per-executor run time).  Not a benchmark: the debug build
reads clocks around every wait.
"""
import argparse
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from chambers_b200 import _lib  # noqa: E402
from chambers_b200 import augmentations as A  # noqa: E402

CLS = {8: "flat", 9: "gather", 10: "sharp", 11: "generic", 12: "gather_sharp"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--size", type=int, default=224)
    ap.add_argument("--policy", default="randaugment")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "timeline.json"))
    args = ap.parse_args()
    B, S = args.batch, args.size
    g = torch.Generator(device="cuda").manual_seed(0)
    n = max(2, min(8, (640 << 20) // (2 * B * S * S * 3)))
    bufs = []
    for _ in range(n):
        x = torch.randint(0, 256, (B, S, S, 3), dtype=torch.uint8, device="cuda", generator=g)
        bufs.append((x, torch.empty_like(x)))
    if args.policy == "randaugment":
        layer = A.RandAugment(2, 10, elementwise=True)._transform
    elif args.policy == "autoaugment":
        layer = A.AutoAugment(elementwise=True)._transform
    else:
        from oracle.policy import magnitude_kwargs
        layer = A.RandomChoice([getattr(A, args.policy)(**magnitude_kwargs(args.policy, 10))], 1)
    for i in range(6):
        layer(bufs[i % n][0], seed=0, call_counter=i, out=bufs[i % n][1])
    torch.cuda.synchronize()
    lib = _lib.load()
    ctx = _lib.context(0)
    _lib.check(ctx, lib.chb_debug_timeline(ctx, None, 0))
    words = 1024 * 2 * 16
    host = np.zeros(words, dtype=np.uint64)
    # one recorded call on buffers that have not been touched for n - 1 calls (L2-cold like the bench)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    layer(bufs[6 % n][0], seed=0, call_counter=6, out=bufs[6 % n][1])
    e1.record()
    torch.cuda.synchronize()
    got = lib.chb_debug_timeline(ctx, host.ctypes.data_as(ctypes.c_void_p), words)
    assert got == words, got
    rec = host.reshape(1, 1024, 2, 16).astype(np.int64)
    t_base = None
    report = {"batch": B, "size": S, "policy": args.policy, "call_ms": e0.elapsed_time(e1), "levels": []}
    print("call %.1f us (debug build)" % (1e3 * e0.elapsed_time(e1)))
    for L in range(1):
        cons = rec[L, :, 1, :]
        prod = rec[L, :, 0, :]
        live = cons[:, 0] > 0
        if not live.any():
            continue
        cons, prod = cons[live], prod[live]
        if t_base is None:
            t_base = cons[:, 0].min()
        rel = lambda a: (float(a.min() - t_base) / 1e3, float(a.max() - t_base) / 1e3)  # noqa: E731
        worked = cons[:, 5] > 0
        lv = {"level": L, "ctas": int(live.sum()), "ctas_with_tiles": int(worked.sum()), "tiles": int(cons[:, 5].sum()),
              "enter_us": rel(cons[:, 0]), "dep_wait_done_us": rel(cons[:, 1]), "exit_us": rel(cons[:, 4])}
        if worked.any():
            lv["first_tile_us"] = rel(cons[worked, 2])
            lv["last_tile_done_us"] = rel(cons[worked, 3])
        for k, nm in ((13, "prod_counters_read_us"), (5, "prod_first_claim_us"), (6, "prod_first_state_us"), (7, "prod_first_issue_us")):
            ok = prod[:, k] > 0
            if ok.any():
                lv[nm] = rel(prod[ok, k])
        tot = lambda a: float(a.sum())  # noqa: E731
        lv["consumer_cycles"] = {"wait_full": tot(cons[:, 6]), "finalise": tot(cons[:, 13]), "finalised_images": int(cons[:, 14].sum())}
        for k, name in CLS.items():
            lv["consumer_cycles"][name] = tot(cons[:, k])
        lv["producer_cycles"] = {"wait_empty": tot(prod[:, 8]), "wait_state": tot(prod[:, 9]), "wait_claim": tot(prod[:, 10]),
                                 "plan": tot(prod[:, 11]), "quiesce": tot(prod[:, 12])}
        nct = max(1, int(live.sum()))
        print("level %d: %d CTAs (%d with tiles), %d tiles" % (L, lv["ctas"], lv["ctas_with_tiles"], lv["tiles"]))
        for k in ("enter_us", "dep_wait_done_us", "prod_counters_read_us", "prod_first_claim_us", "prod_first_state_us", "prod_first_issue_us",
                  "first_tile_us", "last_tile_done_us", "exit_us"):
            if k in lv:
                print("   %-18s first %8.1f  last %8.1f" % (k, lv[k][0], lv[k][1]))
        print("   consumer cycles per CTA: " + "  ".join("%s %.0f" % (k, v / nct) for k, v in lv["consumer_cycles"].items() if k != "finalised_images"))
        print("   producer cycles per CTA: " + "  ".join("%s %.0f" % (k, v / nct) for k, v in lv["producer_cycles"].items()))
        # the slowest CTA
        i = int(np.argmax(cons[:, 3]))
        print("   last CTA to finish: tiles %d, wait_full %d, run %s, finalise %d | producer wait_empty %d state %d claim %d plan %d" % (
            cons[i, 5], cons[i, 6], {CLS[k]: int(cons[i, k]) for k in CLS if cons[i, k]}, cons[i, 13],
            prod[i, 8], prod[i, 9], prod[i, 10], prod[i, 11]))
        report["levels"].append(lv)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(report, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
