#!/usr/bin/env python3
"""Writes a Radiance .hdr (RGBE, run-length coded scanlines) from a B2ENVMP1 blob -- the file format the reference loads its
environment maps from (Image32Bit::read_image_hdr). Used to produce data/sky_64x32.hdr, the fixture of the .hdr reader test.

    python tools/make_hdr.py data/sky_128x64.b2env data/sky_64x32.hdr --downsample 2 [--flat]
"""
import argparse
import struct

import numpy as np


def read_b2env(path):
    raw = open(path, "rb").read()
    assert raw[:8] == b"B2ENVMP1"
    w, h = struct.unpack("<ii", raw[8:16])
    return np.frombuffer(raw, np.float32, offset=16).reshape(h, w, 4)


def to_rgbe(rgb):
    """float RGB -> uint8 RGBE (shared exponent of the largest channel, mantissas truncated)."""
    rgb = np.maximum(np.asarray(rgb, np.float64), 0.0)
    m = rgb.max(axis=-1)
    e = np.zeros(m.shape, np.int32)
    nz = m > 1e-32
    mant, ex = np.frexp(m[nz])
    e[nz] = ex + 128
    scale = np.zeros(m.shape)
    scale[nz] = mant * 256.0 / m[nz]
    out = np.zeros(rgb.shape[:-1] + (4,), np.uint8)
    out[..., :3] = np.clip(rgb * scale[..., None], 0, 255).astype(np.uint8)
    out[..., 3] = np.where(nz, e, 0).astype(np.uint8)
    return out


def rle_channel(values):
    out = bytearray()
    i, n = 0, len(values)
    while i < n:
        run = 1
        while i + run < n and run < 127 and values[i + run] == values[i]:
            run += 1
        if run >= 4:
            out += bytes([128 + run, values[i]])
            i += run
            continue
        j = i
        while j < n and j - i < 128:
            run = 1
            while j + run < n and run < 4 and values[j + run] == values[j]:
                run += 1
            if run >= 4:
                break
            j += 1
        out += bytes([j - i]) + bytes(values[i:j])
        i = j
    return bytes(out)


def write_hdr(path, rgbe, flat=False):
    h, w, _ = rgbe.shape
    with open(path, "wb") as f:
        f.write(b"#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n")
        f.write(f"-Y {h} +X {w}\n".encode())
        for y in range(h):
            if flat or w < 8 or w >= 32768:
                f.write(rgbe[y].tobytes())
            else:
                f.write(bytes([2, 2, w >> 8, w & 255]))
                for c in range(4):
                    f.write(rle_channel(rgbe[y, :, c].tolist()))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("blob")
    ap.add_argument("out")
    ap.add_argument("--downsample", type=int, default=1)
    ap.add_argument("--flat", action="store_true")
    a = ap.parse_args()
    img = read_b2env(a.blob)[::a.downsample, ::a.downsample, :3]
    write_hdr(a.out, to_rgbe(img), a.flat)
    print(f"wrote {a.out}: {img.shape[1]}x{img.shape[0]}")
