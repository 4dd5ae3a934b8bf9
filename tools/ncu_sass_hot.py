#!/usr/bin/env python3
"""Where one kernel's warps wait, from the SASS page of an ncu report: stall samples per opcode the warp was stalled AT, split by
stall reason, and the top instructions.  python tools/ncu_sass_hot.py REPORT.ncu-rep KERNEL_REGEX [--top N]"""
import collections
import csv
import io
import subprocess
import sys

rep, rx = sys.argv[1], sys.argv[2]
top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 25
text = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "-k", f"regex:{rx}", "-c", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(text)))
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
per_op = collections.defaultdict(lambda: collections.Counter())
insts = []
total = 0
for r in rows[2:]:
    if len(r) < len(hdr) or r[col["# Samples"]] == "# Samples":
        continue
    src = r[col["Source"]].strip()
    toks = src.split()
    if toks and toks[0].startswith("@"):
        toks = toks[1:]
    op = toks[0] if toks else "?"
    op = ".".join(op.split(".")[:2]) if op.split(".")[0] in ("LDG", "STG", "LDL", "STL", "LDS", "STS", "MUFU", "LD", "ST", "ATOMG", "RED") else op.split(".")[0]
    smp = int(r[col["# Samples"]] or 0)
    total += smp
    per_op[op]["samples"] += smp
    per_op[op]["inst"] += int(r[col["Instructions Executed"]] or 0)
    for k in reasons:
        per_op[op][k] += int(r[col[k]] or 0)
    insts.append((smp, r[col["Address"]], src[:90], {k: int(r[col[k]] or 0) for k in reasons}))
print(f"{rx}: {total} samples")
by_reason = collections.Counter()
for op, c in per_op.items():
    for k in reasons:
        by_reason[k] += c[k]
print("by reason:", ", ".join(f"{k[6:]} {100 * v / total:.1f}%" for k, v in by_reason.most_common(9)))
print("\nstalled AT opcode (an instruction waits for its operands: a load's latency shows on its first consumer):")
for op, c in sorted(per_op.items(), key=lambda kv: -kv[1]["samples"])[:top]:
    main = ", ".join(f"{k[6:]} {100 * c[k] / max(c['samples'], 1):.0f}%" for k in sorted(reasons, key=lambda k: -c[k])[:3])
    print(f"  {op:14s} {100 * c['samples'] / total:5.1f}%  inst {c['inst'] / 1e6:8.1f}M   {main}")
print("\ntop instructions:")
for smp, addr, src, rs in sorted(insts, key=lambda t: -t[0])[:top]:
    main = ", ".join(f"{k[6:]} {v}" for k, v in sorted(rs.items(), key=lambda kv: -kv[1])[:2])
    print(f"  {100 * smp / total:5.2f}%  {src:90s} {main}")
