#!/usr/bin/env python
"""Turn ncu output brought back in gpurun_out/ into the small text summaries committed under profiles/.

    python profiles/summarize.py launches gpurun_out/launches_r1.csv > profiles/r1_launches.txt
    python profiles/summarize.py full gpurun_out/prof_r1c_kernels.ncu-rep 0 > profiles/r1c_front_kernel_full.txt   # 0 = first launch in the report
"""
import csv
import subprocess
import sys
from collections import OrderedDict


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit()]
    per = OrderedDict()
    order = []
    for r in rows:
        name = r[4].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
        if "duration" not in r[12]:
            continue
        t = float(r[14].replace(",", ""))
        unit = r[13]
        t_us = t / 1e3 if unit in ("ns", "nsecond") else (t if unit in ("us", "usecond") else t * 1e3)
        per.setdefault(name, []).append(t_us)
        order.append((name, r[7], r[8], t_us))
    tot = sum(sum(v) for v in per.values())
    print("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)")
    print("# %d launches, total %.1f us" % (len(order), tot))
    print("%-70s %6s %12s %12s %7s" % ("kernel", "count", "total_us", "mean_us", "share"))
    for k, v in sorted(per.items(), key=lambda kv: -sum(kv[1])):
        print("%-70s %6d %12.1f %12.1f %6.1f%%" % (k[:70], len(v), sum(v), sum(v) / len(v), 100 * sum(v) / tot))
    print("\n# launch list (in order)")
    for name, blk, grid, t in order:
        print("%-60s block %-14s grid %-16s %10.1f us" % (name[:60], blk, grid, t))


KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_warps", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__average_warp_latency_per_inst_issued.ratio",
    "smsp__sass_average_branch_targets_threads_uniform.pct", "smsp__sass_branch_targets_threads_divergent.sum",
    "smsp__sass_branch_targets.sum",
]


def full(path, index=0):
    """index: which launch of a multi-launch report to summarise."""
    pick = ["--launch-skip", str(int(index)), "--launch-count", "1"]
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"] + pick, capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = OrderedDict((h, (v, u)) for h, u, v in zip(hdr, units, vals))
    print("# ncu --set full --clock-control none, kernel: %s" % d.get("Kernel Name", ("?", ""))[0][:90])
    for k in KEYS:
        if k in d:
            print("%-75s %18s %s" % (k, d[k][0], d[k][1]))
    print("\n# warp stall reasons (warps stalled per issue-active cycle)")
    st = [(h, float(v[0])) for h, v in d.items() if "issue_stalled" in h and h.endswith("per_issue_active.ratio")]
    for h, v in sorted(st, key=lambda x: -x[1]):
        if v > 0.005:
            print("  %-40s %8.3f" % (h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v))
    src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"] + pick, capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    h = rows[1]
    data = []
    for r in rows[2:]:  # the page repeats per kernel section: keep the first section's instruction rows
        if r and r[0] == "Kernel Name":
            break
        if r and r[0].startswith("0x"):
            data.append(r)
    isrc, iex, ismp = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
    tot = sum(int(r[iex]) for r in data)
    totsmp = max(1, sum(int(r[ismp]) for r in data))
    print("\n# SASS: %d instructions in the kernel, %d warp-instructions executed" % (len(data), tot))
    blocks, cur = [], None
    for idx, r in enumerate(data):
        ex, sm = int(r[iex]), int(r[ismp])
        if cur and cur["ex"] == ex:
            cur["n"] += 1
            cur["smp"] += sm
            cur["end"] = idx
        else:
            cur = {"ex": ex, "n": 1, "smp": sm, "start": idx, "end": idx}
            blocks.append(cur)
    print("# hottest straight-line blocks (by stall samples)")
    for b in sorted(blocks, key=lambda b: -b["smp"])[:10]:
        ops = {}
        for r in data[b["start"]:b["end"] + 1]:
            t = r[isrc].split()
            op = t[1] if t and t[0].startswith("@") else (t[0] if t else "?")
            ops[op] = ops.get(op, 0) + 1
        top = ", ".join("%s x%d" % kv for kv in sorted(ops.items(), key=lambda x: -x[1])[:8])
        print("  sass[%5d..%5d] n=%4d exec_share %5.1f%% sample_share %5.1f%% : %s" % (
            b["start"], b["end"], b["n"], 100.0 * b["ex"] * b["n"] / tot, 100.0 * b["smp"] / totsmp, top))
    div = [(int(r[h.index("Divergent Branches")]), r[isrc]) for r in data if r[h.index("Divergent Branches")].isdigit()]
    print("\n# divergent branches (sum over instructions): %d" % sum(x for x, _ in div))


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](*sys.argv[2:])
