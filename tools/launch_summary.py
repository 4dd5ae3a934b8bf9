#!/usr/bin/env python3
"""Per-kernel summary of an `ncu --metrics gpu__time_duration.sum,... --csv` launch list (tools/gpu_runs/gpu_r02m.sh):
    python tools/launch_summary.py gpurun_out/r02m/launches_c4s.csv > profiles/r02_launches_c4s.md
Times are cold-cache and serialised (ncu replays every kernel): the SHARE of each kernel is what compares with the live bench."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[hi]
ki, mi, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
per = collections.defaultdict(dict)
for r in rows[hi + 1:]:
    if len(r) <= vi:
        continue
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    name = r[ki].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
    name = name.replace("B2Variant<0, 5, 1, 3, 2, 0>", "c4").replace("B2Variant<0, 5, 1, 3, 2, 1>", "c4+counters").replace("B2Variant<0, 4, 1, 3, 2, 0>", "c3")
    per[(int(r[0]), name[:64])][r[mi]] = (v, r[ui])
TIME = {"ns": 1e-6, "nsecond": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3}
BYTES = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
agg = collections.defaultdict(lambda: collections.defaultdict(float))
for (_, name), m in per.items():
    a = agg[name]
    t, u = m["gpu__time_duration.sum"]
    t *= TIME[u]
    a["n"] += 1
    a["ms"] += t
    for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        if k in m:
            a[k] += m[k][0] * BYTES[m[k][1]]
    if "smsp__inst_executed.sum" in m:
        inst = m["smsp__inst_executed.sum"][0]
        a["inst"] += inst
        a["lanes_w"] += m["smsp__thread_inst_executed_per_inst_executed.ratio"][0] * inst
        a["warps_w"] += m["sm__warps_active.avg.pct_of_peak_sustained_active"][0] * t
        a["issue_w"] += m["sm__inst_issued.avg.pct_of_peak_sustained_active"][0] * t
total = sum(a["ms"] for a in agg.values())
print(f"launch list: {sys.argv[1]}  ({int(sum(a['n'] for a in agg.values()))} launches, {total:.1f} ms summed, cold-cache / serialised)\n")
print("| kernel | launches | ms | share | DRAM read GB | DRAM write GB | DRAM GB/s | warp-inst G | lanes / inst | warps active % | issue slots busy % |")
print("|---|---|---|---|---|---|---|---|---|---|---|")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["ms"]):
    ms = max(a["ms"], 1e-9)
    rd, wr = a["dram__bytes_read.sum"], a["dram__bytes_write.sum"]
    print(f"| `{k}` | {int(a['n'])} | {a['ms']:.2f} | {a['ms'] / total:.3f} | {rd / 1e9:.2f} | {wr / 1e9:.2f} | {(rd + wr) / 1e6 / ms:.0f} | {a['inst'] / 1e9:.2f} | "
          f"{a['lanes_w'] / max(a['inst'], 1):.1f} | {a['warps_w'] / ms:.1f} | {a['issue_w'] / ms:.1f} |")
