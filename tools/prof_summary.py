#!/usr/bin/env python
"""Compact summaries of ncu output (run here on the CPU box on files brought back in gpurun_out/).

  python tools/prof_summary.py launches <launches.csv>      per-kernel count / total / share from the
                                                            `--metrics gpu__time_duration.sum` launch list
  python tools/prof_summary.py kernel <report.ncu-rep>      key raw metrics of each captured launch
  python tools/prof_summary.py sass <report.ncu-rep> [N]    opcode histogram + the N hottest SASS lines
"""
import collections
import csv
import io
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct"]


def launches(path):
    rows = list(csv.reader(open(path)))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    H = rows[hdr]
    ki, vi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[hdr + 1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v = {"ns": v / 1e3, "us": v, "ms": v * 1e3, "s": v * 1e6}.get(r[ui], v)
        agg.setdefault(re.sub(r"^void ", "", r[ki].split("(")[0]), []).append(v)
    tot = sum(sum(v) for v in agg.values())
    print(f"{'kernel':44s} {'n':>5s} {'total us':>12s} {'avg us':>10s} {'share':>7s}")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"{k[:44]:44s} {len(v):5d} {sum(v):12.1f} {sum(v) / len(v):10.1f} {100 * sum(v) / tot:6.1f}%")


def _ncu(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def kernel(rep):
    rows = _ncu(rep, "raw")
    H = rows[0]
    names = [r[H.index("Kernel Name")][:40] for r in rows[2:]]
    print("launches:", names)
    for k in KEYS:
        if k in H:
            i = H.index(k)
            print(f"{k:85s} {rows[1][i]:>10s} ", "  ".join(r[i] for r in rows[2:]))


def sass(rep, top=25):
    rows = _ncu(rep, "source")
    start = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    H = rows[start]
    ie, si = H.index("Instructions Executed"), H.index("Source")
    body = []
    for r in rows[start + 1:]:
        if r and r[0] == "Kernel Name":
            break                                      # first captured launch only
        if len(r) > ie and r[ie].isdigit():
            body.append(r)
    tot = sum(int(r[ie]) for r in body)
    ops = collections.Counter()
    for r in body:
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[si])
        ops[m.group(2) if m else "?"] += int(r[ie])
    print("warp instructions executed:", tot, " SASS lines:", len(body))
    print("  ".join(f"{o}:{100 * c / tot:.1f}%" for o, c in ops.most_common(24)))
    stall = H.index("Warp Stall Sampling (All Samples)") if "Warp Stall Sampling (All Samples)" in H else None
    if stall is not None:
        hot = sorted(body, key=lambda r: -int(r[stall] or 0))[:top]
        ts = sum(int(r[stall] or 0) for r in body) or 1
        print("hottest by stall samples:")
        for r in hot:
            print(f"  {100 * int(r[stall] or 0) / ts:5.1f}%  exec={int(r[ie]):>10d}  {r[si].strip()[:90]}")


if __name__ == "__main__":
    mode = sys.argv[1]
    if mode == "launches":
        launches(sys.argv[2])
    elif mode == "kernel":
        kernel(sys.argv[2])
    else:
        sass(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 25)
