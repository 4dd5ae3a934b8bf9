#!/usr/bin/env python3
"""Summarises an .ncu-rep captured with --import-source on: key raw metrics, warp instructions per source file / line and
per SASS opcode (with lanes per instruction). Usage: python tools/ncu_lines.py REPORT.ncu-rep [--top N] [--file b2_bvh.h]"""
import argparse
import collections
import csv
import io
import re
import subprocess
import sys

RAW = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__block_size", "launch__grid_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
       "smsp__issue_active.avg.pct", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "dram__bytes_read.sum", "dram__bytes_write.sum",
       "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__icc_request_hit_rate.pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
       "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
       "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum",
       "sm__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_lsu.sum"]


def run(args):
    return subprocess.run(args, capture_output=True, text=True).stdout


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--top", type=int, default=40)
    ap.add_argument("--file", default=None, help="list every line of this source file (substring match)")
    a = ap.parse_args()

    raw = list(csv.reader(io.StringIO(run(["ncu", "-i", a.report, "--page", "raw", "--csv"]))))
    if len(raw) >= 3:
        names, values = raw[0], raw[-1]
        units = raw[1]
        col = {n: i for i, n in enumerate(names)}
        print("kernel:", values[col.get("Kernel Name", 4)][:140])
        for m in RAW:
            if m in col:
                print(f"  {m:95s} {values[col[m]]:>16s} {units[col[m]]}")

    text = run(["ncu", "-i", a.report, "--page", "source", "--csv", "--print-source", "cuda,sass"])
    cur, hdr = None, None
    per_line = collections.defaultdict(lambda: [0, 0, ""])
    per_file = collections.defaultdict(lambda: [0, 0])
    per_op = collections.defaultdict(lambda: [0, 0])
    stall = collections.defaultdict(lambda: collections.defaultdict(int))
    total = 0
    for r in csv.reader(io.StringIO(text)):
        if not r:
            continue
        if r[0] == "File Path":
            cur, hdr = r[1].split("/")[-1], None
            continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            hdr = r
            i_line, i_src, i_sass = 0, 1, 3
            i_inst, i_thr = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
            i_samples = hdr.index("# Samples")
            continue
        if hdr is None or cur is None:
            continue
        try:
            inst, thr = int(r[i_inst] or 0), int(r[i_thr] or 0)
        except ValueError:
            continue
        key = (cur, int(r[i_line] or 0))
        per_line[key][0] += inst
        per_line[key][1] += thr
        per_line[key][2] = r[i_src].strip()[:110]
        per_file[cur][0] += inst
        per_file[cur][1] += thr
        op = r[i_sass].strip().split()
        if op and op[0].startswith("@"):
            op = op[1:]
        opn = op[0].split(".")[0] if op else "?"
        if opn in ("MUFU", "CALL", "LDL", "STL", "LDG", "LDS", "STS", "BAR", "DFMA", "DMUL", "DADD") and op:
            opn = ".".join(op[0].split(".")[:2])
        per_op[opn][0] += inst
        per_op[opn][1] += thr
        total += inst
    print(f"\ntotal warp instructions: {total / 1e6:.1f} M")
    print("\nper file (warp instr M, %, lanes):")
    for f, (i, t) in sorted(per_file.items(), key=lambda kv: -kv[1][0]):
        if i:
            print(f"  {f:32s} {i / 1e6:10.1f} {100 * i / total:6.1f}% {t / i:6.1f}")
    print("\nper opcode:")
    for o, (i, t) in sorted(per_op.items(), key=lambda kv: -kv[1][0])[:45]:
        if i:
            print(f"  {o:16s} {i / 1e6:10.1f} {100 * i / total:6.1f}% {t / i:6.1f}")
    print(f"\ntop {a.top} lines:")
    for (f, l), (i, t, s) in sorted(per_line.items(), key=lambda kv: -kv[1][0])[:a.top]:
        if i:
            print(f"  {f}:{l:<5d} {i / 1e6:8.1f} {100 * i / total:5.1f}% {t / i:5.1f}  {s}")
    if a.file:
        print(f"\nlines of {a.file}:")
        for (f, l), (i, t, s) in sorted(per_line.items()):
            if a.file in f and i:
                print(f"  {f}:{l:<5d} {i / 1e6:8.1f} {100 * i / total:5.1f}% {t / i:5.1f}  {s}")


if __name__ == "__main__":
    main()
