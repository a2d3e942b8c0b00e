#!/usr/bin/env python
"""Headline benchmark: augmented images/sec of RandAugment(N=2, M=10) on 256x224x224x3 uint8
batches (BASELINE.json configs[1]) per B200, weak-scaled over --gpus N with no data-path collective.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

One JSON line on rank 0.  A "step" is one pass of the policy over one batch:
  value     whole-job images/s with the inputs already resident in HBM (one plan kernel + one pass
            kernel per level per step, CUDA-event timed, max over ranks); input/output buffers rotate through a pool larger
            than L2 so no step re-reads a cached batch
  e2e       the same metric through the public layer API with HOST (pinned) buffers: H2D copy,
            kernels and D2H copy inside the timed region (chb_policy_apply_host)
  roofline  HBM: algorithmic bytes (2*H*W*C per image) / average duration of one step's
            kernels, against MEASURED_PEAKS.json's measured copy bandwidth; traffic = DRAM bytes of one
            step from the ncu capture recorded in profiles/r01_traffic.json
  cpu_baseline  the oracle port (numpy restatement of the reference; TensorFlow is not installable
            here) timed on this box's host cores over a bounded sample -- a reported baseline only
--impl reference times that same CPU port as the reference arm (rank 0 only).
"""

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H = W = 224
C = 3
BATCH = 256           # per GPU (weak scaling)
N_TRANSFORMS = 2
MAGNITUDE = 10
SEED = 0
POOL_BYTES = 640 << 20  # rotating in+out pool per GPU, >> 126 MB L2


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--elementwise", type=int, default=1)
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--policy", default="randaugment", choices=["randaugment", "autoaugment"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


# ----------------------------------------------------------------------------- CPU baseline
def _cpu_worker(job):
    import numpy as np
    import oracle
    seed, start, n, policy_name, elementwise, batch_total = job
    rng = np.random.default_rng(seed + start)
    x = rng.integers(0, 256, size=(n, H, W, C), dtype=np.uint8)
    pol = oracle.randaugment_policy(N_TRANSFORMS, MAGNITUDE) if policy_name == "randaugment" else oracle.autoaugment_policy()
    sched = oracle.decode_schedule(pol, SEED, 0, start, n, H, W, bool(elementwise))
    t0 = time.perf_counter()
    if elementwise:
        oracle.apply_schedule(x, pol, sched, True)
    else:
        # batch mode shares one schedule row; Contrast's constant needs the whole batch size, which
        # the oracle derives from the array it is given -- fine for a throughput baseline.
        oracle.apply_schedule(x, pol, sched, False)
    return n, time.perf_counter() - t0


def cpu_port_throughput(batch, policy_name, elementwise, min_seconds=8.0, max_batches=64):
    """images/s of the oracle port over all host cores (process pool over batch shards)."""
    import multiprocessing as mp
    cores = len(os.sched_getaffinity(0))
    ctx = mp.get_context("spawn")
    done = 0
    steps = []
    with ctx.Pool(cores) as pool:
        pool.map(_cpu_worker, [(1, 0, 1, policy_name, elementwise, batch)] * cores)  # warm the workers
        t_start = time.perf_counter()
        for step in range(max_batches):
            per = (batch + cores - 1) // cores
            jobs = [(SEED, s, min(per, batch - s), policy_name, elementwise, batch) for s in range(0, batch, per)]
            t0 = time.perf_counter()
            pool.map(_cpu_worker, jobs)
            steps.append(time.perf_counter() - t0)
            done += batch
            if time.perf_counter() - t_start >= min_seconds:
                break
        total = time.perf_counter() - t_start
    return done / total, cores, done, steps


# --------------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    REASONS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.sm, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        self._h = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            # CUDA_VISIBLE_DEVICES may renumber devices; go through the PCI bus id of the torch device.
            import torch
            bus = torch.cuda.get_device_properties(gpu_index).pci_bus_id if hasattr(
                torch.cuda.get_device_properties(gpu_index), "pci_bus_id") else None
            self._h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index) if bus is None else self._by_bus(bus, gpu_index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._h = None

    def _by_bus(self, bus, fallback):
        nv = self._nv
        try:
            for i in range(nv.nvmlDeviceGetCount()):
                h = nv.nvmlDeviceGetHandleByIndex(i)
                if int(nv.nvmlDeviceGetPciInfo(h).bus) == int(bus):
                    return h
        except Exception:
            pass
        return nv.nvmlDeviceGetHandleByIndex(fallback)

    def _sample(self):
        nv = self._nv
        try:
            self.sm.append(float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
            mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)) if hasattr(
                nv, "nvmlDeviceGetCurrentClocksEventReasons") else int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h))
            for name, bit in self.REASONS.items():
                if mask & bit:
                    self.reasons.add(name)
        except Exception:
            pass

    def _run(self):
        while not self._stop.is_set():
            self._sample()
            time.sleep(0.005)

    def start(self):
        if self._h is None:
            return
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._thread.start()

    def stop(self):
        if self._thread is not None:
            self._sample()
            self._stop.set()
            self._thread.join(timeout=1.0)
        return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.sm)}


def measured_traffic(args):
    """DRAM bytes (read + write) of one step, from an `ncu` capture of this command (tools/gpu_traffic.sh);
    None if no capture of this workload is committed."""
    try:
        rec = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
        key = "%s_b%d" % (args.policy, args.batch)
        return rec[key]["dram_bytes_per_step"]
    except Exception:
        return None


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------ reference arm
def run_reference(args, rank):
    if rank != 0:
        return
    per_step = []
    cores = len(os.sched_getaffinity(0))
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    sample = min(args.batch, max(8 * cores, 64))  # images per step: a bounded sample of the 256-image batch
    with ctx.Pool(cores) as pool:
        per = (sample + cores - 1) // cores
        jobs = [(SEED, s, min(per, sample - s), args.policy, args.elementwise, args.batch) for s in range(0, sample, per)]
        for _ in range(max(args.warmup, 1)):
            pool.map(_cpu_worker, jobs)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            t1 = time.perf_counter()
            pool.map(_cpu_worker, jobs)
            per_step.append(time.perf_counter() - t1)
        total = time.perf_counter() - t0
    value = sample * args.steps / total
    line = {
        "impl": "reference", "metric": "augmented images/sec", "value": value, "unit": "images/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": workload_config(args, sample_note="each step = a %d-image sample of the batch" % sample),
        "cpu_baseline": {"value": value, "unit": "images/s", "cores": cores, "kind": "port",
                         "sample": "%d steps x %d images of the workload batch; numpy restatement of the reference "
                                   "(TensorFlow 2.6 / tensorflow-addons are not installable here), process pool over "
                                   "all host cores" % (args.steps, sample)},
        "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(args, sample_note=None):
    cfg = {
        "workload": "%s %s on %dx%dx%dx%d uint8 per GPU (BASELINE.json configs[1])" % (
            "RandAugment(N=%d,M=%d)" % (N_TRANSFORMS, MAGNITUDE) if args.policy == "randaugment" else "AutoAugment(V0)",
            "elementwise" if args.elementwise else "batchwise", args.batch, H, W, C),
        "batch_per_gpu": args.batch, "global_batch": args.batch * args.gpus, "image": [H, W, C],
        "elementwise": bool(args.elementwise), "parallelism": "batch-sharded x%d, no collective" % args.gpus,
        "l2": "in/out buffers rotate through a %d MiB pool per GPU (> 126 MB L2)" % (POOL_BYTES >> 20),
        "input": "i.i.d. uniform uint8, torch.Generator seed 0",
    }
    if sample_note:
        cfg["sample"] = sample_note
    return cfg


# ---------------------------------------------------------------------------------- native arm
def run_native(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    from chambers_b200 import build, _lib
    build.build_library()
    from chambers_b200 import augmentations as A
    from chambers_b200.sharding import shard_kwargs

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the native arm has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=dev)

    B = args.batch
    img_bytes = H * W * C
    n_pool = max(2, POOL_BYTES // (2 * B * img_bytes))
    g = torch.Generator(device="cpu").manual_seed(SEED + rank)
    host_pool = torch.randint(0, 256, (B, H, W, C), dtype=torch.uint8, generator=g).pin_memory()
    ins = []
    for i in range(n_pool):
        t = host_pool.to(dev, non_blocking=True)
        if i:
            t = t.roll(i, 0).contiguous()
        ins.append(t)
    outs = [torch.empty_like(ins[0]) for _ in range(n_pool)]
    if args.policy == "randaugment":
        layer = A.RandAugment(N_TRANSFORMS, MAGNITUDE, elementwise=bool(args.elementwise))
    else:
        layer = A.AutoAugment(elementwise=bool(args.elementwise))
    choice = layer._transform
    skw = shard_kwargs(B * world, rank, world)

    def step(i):
        choice(ins[i % n_pool], seed=SEED, call_counter=i, out=outs[i % n_pool], **skw)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    for i in range(args.warmup):
        step(i)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = _lib.kernel_launches(local_rank)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for i in range(args.steps):
        step(args.warmup + i)
    ev1.record()
    barrier()
    elapsed_ms = ev0.elapsed_time(ev1)
    launches = _lib.kernel_launches(local_rank) - launches0
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t)

    # ---- e2e: host buffers through the public layer API
    e2e = None
    if not args.no_e2e:
        host_out = torch.empty_like(host_pool).pin_memory()
        e2e_steps = max(3, min(args.steps, 20))
        for i in range(2):
            choice(host_pool, seed=SEED, call_counter=i, out=host_out, **skw)
        barrier()
        t0 = time.perf_counter()
        for i in range(e2e_steps):
            choice(host_pool, seed=SEED, call_counter=i, out=host_out, **skw)  # synchronous: returns with host_out complete
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e = {"value": world * B * e2e_steps / float(tt), "unit": "images/s",
               "h2d_bytes_per_step": B * img_bytes, "d2h_bytes_per_step": B * img_bytes,
               "steps": e2e_steps, "api": "RandomChoice.__call__(pinned host tensor) -> chb_policy_apply_host"}

    if rank != 0:
        return
    value = world * B * args.steps / (elapsed_ms * 1e-3)
    ms_per_step = elapsed_ms / args.steps
    peak, peak_src = measured_peak()
    alg_bytes = 2.0 * B * img_bytes  # per launch: one read + one write of every image
    achieved = alg_bytes / (ms_per_step * 1e-3) / 1e9
    line = {
        "metric": "augmented images/sec", "value": value, "unit": "images/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": workload_config(args),
        "clocks": clocks,
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": measured_traffic(args), "peak_source": peak_src,
                     "kernel": "chb::plan_kernel + chb::pass_kernel<3> (one step)",
                     "algorithmic_bytes_per_launch": alg_bytes,
                     "frac_of_spec_8TBs": achieved / 8000.0},
        "e2e": e2e,
    }
    if world == 1 and not args.no_cpu_baseline:
        v, cores, n_done, _ = cpu_port_throughput(B, args.policy, args.elementwise)
        line["cpu_baseline"] = {
            "value": v, "unit": "images/s", "cores": cores, "kind": "port",
            "sample": "%d images (%d passes over the %d-image batch) through the numpy restatement of the reference, "
                      "process pool over %d host cores; TensorFlow itself is not installable here"
                      % (n_done, n_done // B, B, cores)}
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        # launched without torchrun: re-exec under it
        port = 29500 + os.getpid() % 2000
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", str(port)] + sys.argv
        raise SystemExit(subprocess.call(cmd))
    try:
        run_native(args, rank, local_rank, world)
    finally:
        try:
            import torch.distributed as dist
            if dist.is_initialized():
                dist.destroy_process_group()
        except Exception:
            pass


if __name__ == "__main__":
    main()
