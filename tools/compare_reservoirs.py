#!/usr/bin/env python3
"""Compare per-pass ReSTIR DI reservoir dumps (36-byte records: M, weight_sum, UCW, point xyz, triangle, target, flags).
    python tools/compare_reservoirs.py <prefix_a> <prefix_b> <pass> <frame> [width]"""
import sys
import numpy as np

FIELDS = ["M", "weight_sum", "UCW", "px", "py", "pz", "tri", "target", "flags"]


def load(prefix, name, frame):
    a = np.fromfile(f"{prefix}restir_{name}_f{frame}.bin", np.uint32).reshape(-1, 9).copy()
    a[:, 8] &= 0xFF
    return a


def compare(pa, pb, name, frame, width=None, verbose=True):
    a, b = load(pa, name, frame), load(pb, name, frame)
    # only active reservoirs matter: inactive pixels keep whatever the buffer held
    neq = (a != b)
    # NaN payloads compare unequal bitwise only if the bits differ; treat NaN == NaN
    fa, fb = a.view(np.float32), b.view(np.float32)
    for col in (1, 2, 3, 4, 5, 7):
        both_nan = np.isnan(fa[:, col]) & np.isnan(fb[:, col])
        neq[:, col] &= ~both_nan
    bad = neq.any(axis=1)
    out = dict(records=len(a), mismatching=int(bad.sum()), per_field={f: int(neq[:, i].sum()) for i, f in enumerate(FIELDS)})
    if verbose:
        print(name, frame, out)
        idx = np.nonzero(bad)[0][:5]
        for i in idx:
            loc = f"({i % width},{i // width})" if width else str(i)
            print(" pixel", loc)
            print("   a:", a[i, 0].view(np.int32) if False else int(a[i, 0]), fa[i, 1:6], int(a[i, 6].astype(np.int32)), fa[i, 7], hex(a[i, 8]))
            print("   b:", int(b[i, 0]), fb[i, 1:6], int(b[i, 6].astype(np.int32)), fb[i, 7], hex(b[i, 8]))
    return out


if __name__ == "__main__":
    compare(sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]), int(sys.argv[5]) if len(sys.argv) > 5 else None)
