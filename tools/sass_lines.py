#!/usr/bin/env python3
"""Static SASS size per source file / line / inlined-at function of one kernel of an object file (needs -lineinfo).
Usage: python tools/sass_lines.py OBJECT.o KERNEL_SUBSTRING [--top N]"""
import argparse
import collections
import re
import subprocess
import tempfile
from pathlib import Path


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("obj")
    ap.add_argument("kernel")
    ap.add_argument("--top", type=int, default=40)
    a = ap.parse_args()
    with tempfile.TemporaryDirectory() as d:
        subprocess.run(["cuobjdump", "-xelf", "all", str(Path(a.obj).resolve())], cwd=d, check=True, capture_output=True)
        cubin = next(Path(d).glob("*.cubin"))
        text = subprocess.run(["nvdisasm", "-g", "-c", str(cubin)], capture_output=True, text=True).stdout
    per_file, per_line, per_fn = collections.Counter(), collections.Counter(), collections.Counter()
    in_kernel, cur, total = False, ("?", 0), 0
    for line in text.splitlines():
        m = re.match(r"\s*\.text\.(\S+):", line)
        if m:
            in_kernel = a.kernel in m.group(1)
            continue
        if not in_kernel:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', line)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+", line):
            per_file[cur[0]] += 1
            per_line[cur] += 1
            total += 1
    print(f"kernel *{a.kernel}*: {total} instructions = {total * 16 / 1024:.0f} KB")
    for f, n in per_file.most_common(20):
        print(f"  {f:30s} {n:7d} {100 * n / total:5.1f}%")
    print("top lines:")
    for (f, l), n in per_line.most_common(a.top):
        print(f"  {f}:{l:<5d} {n:6d} {100 * n / total:5.1f}%")


if __name__ == "__main__":
    main()
