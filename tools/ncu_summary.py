#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU): headline metrics, stall reasons, and the hottest
source lines / SASS instructions.   python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--sass N]"""
import csv
import io
import subprocess
import sys


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
    "sm__cycles_active.avg", "sm__cycles_elapsed.max", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__lsu_writeback_active_mem_lg.sum.pct_of_peak_sustained_elapsed",
    "launch__registers_per_thread", "smsp__warps_eligible.avg.per_cycle_active", "smsp__warps_active.avg.per_cycle_active",
    "sm__inst_executed_pipe_xu.sum", "smsp__inst_executed_pipe_xu.sum", "l1tex__data_pipe_lsu_wavefronts.sum",
]


def main():
    rep = sys.argv[1]
    nsass = int(sys.argv[sys.argv.index("--sass") + 1]) if "--sass" in sys.argv else 25
    hdr, units, data = raw(rep)
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print("%-88s %-10s %s" % (k, units[i], [r[i] for r in data]))
    print("-- stall reasons (warps per issue-active cycle), launch 0")
    st = []
    for i, k in enumerate(hdr):
        if "issue_stalled" in k and k.endswith("per_issue_active.ratio") and "not_issued" not in k:
            try:
                st.append((float(data[0][i]), k.split("issue_stalled_")[1].split("_per_issue")[0]))
            except ValueError:
                pass
    for v, k in sorted(st, reverse=True)[:10]:
        print("   %-28s %.3f" % (k, v))
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    while rows and (len(rows[0]) < 6 or rows[0][0] != "Address"):
        rows.pop(0)  # leading "Kernel Name" line(s)
    if len(rows) < 3:
        print("no source page")
        return
    h = rows[0]
    def col(name):
        for i, k in enumerate(h):
            if k.strip() == name:
                return i
        return None
    ci = col("Source"); cs = col("Warp Stall Sampling (All Samples)") or col("Warp Stall Sampling (All Cycles)")
    ce = col("Instructions Executed")
    cx = col("Thread Instructions Executed")
    body = [r for r in rows[1:] if len(r) == len(h)]
    def num(x):
        try:
            return float(x.replace(",", ""))
        except Exception:
            return 0.0
    tot = sum(num(r[cs]) for r in body) or 1.0
    print("-- hottest SASS by stall samples (total %.0f), %d instructions in kernel" % (tot, len(body)))
    ranked = sorted(range(len(body)), key=lambda i: -num(body[i][cs]))[:nsass]
    for i in ranked:
        r = body[i]
        reasons = sorted(((num(r[j]), h[j]) for j in range(len(h)) if h[j].startswith("stall_") and "(" not in h[j]), reverse=True)[:2]
        print("   %5.2f%%  exec %10s  #%5d  %-70s %s" % (100 * num(r[cs]) / tot, r[ce] if ce is not None else "", i, r[ci][:70],
              " ".join("%s=%d" % (n.replace("stall_", ""), v) for v, n in reasons if v > 0)))
    # cumulative by contiguous hot regions
    print("-- regions (64-instruction buckets) by stall share")
    buckets = {}
    for i, r in enumerate(body):
        buckets[i // 64] = buckets.get(i // 64, 0.0) + num(r[cs])
    ex = {}
    for i, r in enumerate(body):
        ex[i // 64] = ex.get(i // 64, 0.0) + num(r[ce])
    totex = sum(ex.values()) or 1.0
    for b, v in sorted(buckets.items(), key=lambda kv: -kv[1])[:14]:
        print("   instr %5d-%5d  stall %5.1f%%   executed %5.1f%%" % (b * 64, b * 64 + 63, 100 * v / tot, 100 * ex[b] / totex))


if __name__ == "__main__":
    main()
