#!/usr/bin/env python3
"""Compare two dump prefixes written by the oracle harnesses / the CUDA renderer (G-buffer + framebuffer)."""
import sys
import numpy as np


def load(prefix, name, dtype):
    try:
        return np.fromfile(prefix + name + ".bin", dtype)
    except FileNotFoundError:
        return None


def compare(pa, pb, width, height, rel_tol=1e-3, verbose=True, window=None):
    """window = (x0, y0, x1, y1): compare only those pixels (the oracle rendered a crop of the frame)."""
    out = {}
    n = width * height
    sel = np.ones(n, bool)
    if window is not None:
        x0, y0, x1, y1 = window
        m = np.zeros((height, width), bool)
        m[y0:y1, x0:x1] = True
        sel = m.ravel()
    prim_a, prim_b = load(pa, "gbuffer_prim", np.int32), load(pb, "gbuffer_prim", np.int32)
    if prim_a is not None and prim_b is not None:
        out["prim_mismatch"] = int((prim_a != prim_b)[sel].sum())
        pos_a = load(pa, "gbuffer_position", np.float32).reshape(n, 3)
        pos_b = load(pb, "gbuffer_position", np.float32).reshape(n, 3)
        same = (prim_a == prim_b) & sel
        out["position_bit_mismatch"] = int((pos_a[same].view(np.uint32) != pos_b[same].view(np.uint32)).any(axis=1).sum())
        for key in ("gbuffer_geometric_normal", "gbuffer_shading_normal"):
            a, b = load(pa, key, np.uint32), load(pb, key, np.uint32)
            out[key + "_mismatch"] = int(((a & 0xFFFFFF) != (b & 0xFFFFFF))[same].sum())
        ma = load(pa, "gbuffer_material", np.uint32).reshape(n, 27).copy()
        mb = load(pb, "gbuffer_material", np.uint32).reshape(n, 27).copy()
        ma[:, 0] &= 0xFF  # byte 0 = flags, bytes 1..3 = struct padding (indeterminate in the reference)
        mb[:, 0] &= 0xFF
        out["material_mismatch"] = int((ma != mb).any(axis=1)[same].sum())
    fa, fb = load(pa, "framebuffer", np.float32), load(pb, "framebuffer", np.float32)
    if fa is not None and fb is not None:
        fa, fb = fa.reshape(n, 3)[sel], fb.reshape(n, 3)[sel]
        denom = np.maximum(np.abs(fa), np.abs(fb))
        rel = np.where(denom > 1e-6, np.abs(fa - fb) / np.maximum(denom, 1e-30), 0.0).max(axis=1)
        out["pixels"] = int(sel.sum())
        out["pixels_over_tol"] = int((rel > rel_tol).sum())
        out["pixels_bit_exact"] = int((fa.view(np.uint32) == fb.view(np.uint32)).all(axis=1).sum())
        out["max_rel"] = float(rel.max())
        out["mean_a"] = float(fa.mean())
        out["mean_b"] = float(fb.mean())
        out["rmse"] = float(np.sqrt(((fa - fb) ** 2).mean()))
    if verbose:
        for k, v in out.items():
            print(f"{k}: {v}")
    return out


if __name__ == "__main__":
    compare(sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4]))
