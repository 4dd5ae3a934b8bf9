#!/usr/bin/env python3
"""Seeded procedural HDR sky (lat-long, RGBA fp32) for BASELINE.json configs[3]: a horizon-to-zenith gradient, a small
very bright sun disc (so luminance-proportional alias sampling differs strongly from uniform) and low-frequency
"cloud" modulation. Stands in for the reference's data/Skyspheres/evening_road_01_puresky_2k.hdr, which is not copied
into this repository. Writes the "B2ENVMP1" blob both the CUDA path and the oracles read.

    python tools/make_envmap.py out.b2env [--width 2048 --height 1024 --seed 7]
"""
import argparse
import struct

import numpy as np


def build(width, height, seed):
    rng = np.random.Generator(np.random.PCG64(seed))
    v = (np.arange(height, dtype=np.float64) + 0.5) / height          # 0 = top row
    u = (np.arange(width, dtype=np.float64) + 0.5) / width
    theta = v[:, None] * np.pi                                          # polar angle from +Y
    phi = u[None, :] * 2.0 * np.pi
    up = np.cos(theta) * np.ones_like(phi)
    d = np.stack([np.sin(theta) * np.cos(phi), up, np.sin(theta) * np.sin(phi)], -1)
    t = np.clip(up, 0.0, 1.0)[..., None]
    horizon, zenith, ground = np.array([1.0, 0.62, 0.38]), np.array([0.12, 0.25, 0.65]), np.array([0.08, 0.07, 0.06])
    sky = horizon * (1.0 - t) ** 3 + zenith * (1.0 - (1.0 - t) ** 3)
    img = np.where(up[..., None] >= 0.0, sky, ground)
    # low-frequency clouds: a few random cosine lobes on the sphere
    for _ in range(12):
        axis = rng.normal(size=3)
        axis /= np.linalg.norm(axis)
        axis[1] = abs(axis[1])
        lobe = np.clip(d @ axis, 0.0, 1.0) ** rng.uniform(8.0, 40.0)
        img = img + lobe[..., None] * rng.uniform(0.05, 0.35) * np.array([1.0, 0.95, 0.9])
    sun_dir = np.array([0.35, 0.45, -0.82])
    sun_dir /= np.linalg.norm(sun_dir)
    cos_sun = d @ sun_dir
    img = img + (cos_sun > np.cos(np.radians(1.2)))[..., None] * np.array([9000.0, 7800.0, 6000.0])
    img = img + (np.clip(cos_sun, 0.0, 1.0) ** 256)[..., None] * np.array([6.0, 4.5, 3.0])
    rgba = np.concatenate([img, np.ones((height, width, 1))], -1).astype(np.float32)
    return rgba


def write(path, rgba):
    h, w, _ = rgba.shape
    with open(path, "wb") as f:
        f.write(b"B2ENVMP1")
        f.write(struct.pack("<2i", w, h))
        f.write(np.ascontiguousarray(rgba, np.float32).tobytes())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("out")
    ap.add_argument("--width", type=int, default=2048)
    ap.add_argument("--height", type=int, default=1024)
    ap.add_argument("--seed", type=int, default=7)
    a = ap.parse_args()
    write(a.out, build(a.width, a.height, a.seed))


if __name__ == "__main__":
    main()
