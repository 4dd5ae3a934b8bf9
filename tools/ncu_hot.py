#!/usr/bin/env python3
"""Stall-sample hot spots of one kernel of an ncu report captured with --import-source on:
    python tools/ncu_hot.py REPORT.ncu-rep KERNEL_REGEX [--top N]
Prints samples per source file, per function-ish source line and per SASS opcode."""
import collections
import csv
import io
import subprocess
import sys

rep, rx = sys.argv[1], sys.argv[2]
top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 40
text = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "-k", f"regex:{rx}", "-c", "1"], capture_output=True, text=True).stdout
cur = hdr = None
per_line = collections.defaultdict(lambda: [0, 0, ""])
per_file = collections.defaultdict(lambda: [0, 0])
per_op = collections.defaultdict(lambda: [0, 0])
tot = 0
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
        i_inst, i_s = hdr.index("Instructions Executed"), hdr.index("# Samples")
        continue
    if hdr is None or cur is None:
        continue
    try:
        inst, smp = int(r[i_inst] or 0), int(r[i_s] or 0)
    except ValueError:
        continue
    k = (cur, int(r[0] or 0))
    per_line[k][0] += inst
    per_line[k][1] += smp
    per_line[k][2] = r[1].strip()[:110]
    per_file[cur][0] += inst
    per_file[cur][1] += smp
    op = r[3].strip().split()
    if op and op[0].startswith("@"):
        op = op[1:]
    if op:
        name = ".".join(op[0].split(".")[:2]) if op[0].split(".")[0] in ("LDG", "STG", "LDL", "STL", "LDS", "STS", "MUFU", "CALL", "BAR") else op[0].split(".")[0]
        per_op[name][0] += inst
        per_op[name][1] += smp
    tot += smp
print(f"{rx}: {tot} stall samples")
for f, (i, s) in sorted(per_file.items(), key=lambda kv: -kv[1][1])[:14]:
    print(f"  {f:30s} inst {i / 1e6:8.1f}M  samples {100 * s / max(tot, 1):5.1f}%")
print("per opcode (by samples):")
for o, (i, s) in sorted(per_op.items(), key=lambda kv: -kv[1][1])[:16]:
    print(f"  {o:12s} inst {i / 1e6:8.1f}M  samples {100 * s / max(tot, 1):5.1f}%")
print("per source line:")
for (f, l), (i, s, src) in sorted(per_line.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"  {f}:{l:<5d} inst {i / 1e6:7.1f}M {100 * s / max(tot, 1):5.1f}%  {src}")
