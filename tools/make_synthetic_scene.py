#!/usr/bin/env python3
"""Seeded procedural many-light scene for BASELINE.json configs[2..4] (SURVEY.md 8d "C3"): a closed box filled with a
jittered 3-D grid of tessellated spheres (opaque principled materials) and many small spherical emitters whose power
is log-uniform over three decades, so power-proportional light sampling differs from uniform. Writes the B2SCENE1 blob
both the CUDA path and the reference oracle read.

    python tools/make_synthetic_scene.py --triangles 10000000 --emissive 100000 --seed 45568 --aspect 1.7777778 out.b2scene

Default = the full-size scene: exactly 10,000,000 triangles, 100,000 of them emissive (1250 emitters x 80 triangles),
16 non-emissive + 32 emissive materials, vertices SHARED inside each sphere (smooth normals), camera inside the box.
Smaller sizes (tests) scale the number of spheres / emitters down; the layout recipe is the same.
Deterministic: numpy Generator(PCG64(seed)); no file or clock input.
"""
import argparse
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent))
import b2scene  # noqa: E402


def icosphere(level):
    """Unit icosphere with shared vertices: 20 * 4^level triangles, 10 * 4^level + 2 vertices."""
    t = (1.0 + 5.0 ** 0.5) / 2.0
    v = np.array([[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0], [0, -1, t], [0, 1, t], [0, -1, -t], [0, 1, -t],
                  [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]], np.float64)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    f = np.array([[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11], [1, 5, 9], [5, 11, 4], [11, 10, 2], [10, 7, 6], [7, 1, 8],
                  [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9], [4, 9, 5], [2, 4, 11], [6, 2, 10], [8, 6, 7], [9, 8, 1]], np.int64)
    for _ in range(level):
        edges = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]])
        edges.sort(axis=1)
        uniq, inverse = np.unique(edges, axis=0, return_inverse=True)
        mid = v[uniq[:, 0]] + v[uniq[:, 1]]
        mid /= np.linalg.norm(mid, axis=1, keepdims=True)
        base = len(v)
        v = np.concatenate([v, mid])
        n = len(f)
        ab, bc, ca = base + inverse[:n], base + inverse[n:2 * n], base + inverse[2 * n:]
        a, b, c = f[:, 0], f[:, 1], f[:, 2]
        f = np.concatenate([np.stack([a, ab, ca], 1), np.stack([b, bc, ab], 1), np.stack([c, ca, bc], 1), np.stack([ab, bc, ca], 1)])
    return v.astype(np.float32), f.astype(np.int32)


def plan(n_triangles, n_emissive):
    """How many spheres of each tessellation level (+ box 12 + filler quads) give exactly n_triangles."""
    emitter_level = 1
    emitter_tris = 20 * 4 ** emitter_level
    if n_emissive % emitter_tris:
        raise SystemExit(f"--emissive must be a multiple of {emitter_tris}")
    n_emitters = n_emissive // emitter_tris
    remaining = n_triangles - n_emissive - 12
    if remaining < 0 or remaining % 2:
        raise SystemExit("triangle budget must leave an even, non-negative count after the box and the emitters")
    big_level = 4 if remaining >= 200 * 5120 else (3 if remaining >= 50 * 1280 else 2)
    big_tris = 20 * 4 ** big_level
    n_big = remaining // big_tris
    filler_quads = (remaining - n_big * big_tris) // 2
    return dict(emitter_level=emitter_level, n_emitters=n_emitters, big_level=big_level, n_big=n_big, filler_quads=filler_quads)


def build(n_triangles, n_emissive, seed, aspect):
    rng = np.random.Generator(np.random.PCG64(seed))
    p = plan(n_triangles, n_emissive)
    cells = 1
    while cells ** 3 < int(1.15 * p["n_big"]) + 30:
        cells += 1
    cell = 2.0
    half = cells * cell / 2.0

    # camera: inside the box near the +z wall, looking down -z; the cells around it stay empty
    cam_pos = np.array([0.0, 0.0, half - 1.2])
    centers = (np.stack(np.meshgrid(*[np.arange(cells)] * 3, indexing="ij"), -1).reshape(-1, 3) + 0.5) * cell - half
    keep = np.linalg.norm(centers - cam_pos, axis=1) > 3.2
    centers = centers[keep]
    if len(centers) < p["n_big"]:
        raise SystemExit("internal: not enough cells")
    centers = centers[rng.permutation(len(centers))[:p["n_big"]]]
    centers = centers + rng.uniform(-0.25, 0.25, centers.shape)
    radii = rng.uniform(0.55, 0.72, len(centers))

    corners = np.stack(np.meshgrid(*[np.arange(1, cells)] * 3, indexing="ij"), -1).reshape(-1, 3) * cell - half
    corners = corners[np.linalg.norm(corners - cam_pos, axis=1) > 2.0]
    if len(corners) < p["n_emitters"]:
        raise SystemExit("internal: not enough emitter sites")
    e_centers = corners[rng.permutation(len(corners))[:p["n_emitters"]]] + rng.uniform(-0.15, 0.15, (p["n_emitters"], 3))
    e_radii = rng.uniform(0.04, 0.08, p["n_emitters"])

    # materials: 16 opaque principled surfaces + 32 emissive ones with log-spaced power over 3 decades
    mats = []
    for i in range(16):
        hue = rng.uniform(0.15, 0.95, 3)
        kind = i % 4
        kw = dict(base_color=hue, roughness=float(rng.uniform(0.15, 0.8)))
        if kind == 1:
            kw.update(metallic=1.0, roughness=float(rng.uniform(0.1, 0.5)))
        elif kind == 2:
            kw.update(coat=1.0, coat_roughness=float(rng.uniform(0.0, 0.2)))
        elif kind == 3:
            kw.update(specular=float(rng.uniform(0.3, 1.0)), sheen=float(rng.uniform(0.0, 0.5)))
        mats.append(b2scene.default_material(**kw))
    n_emissive_mats = 32
    for i in range(n_emissive_mats):
        power = 10.0 ** (3.0 * i / (n_emissive_mats - 1))  # 1 .. 1000
        tint = rng.uniform(0.6, 1.0, 3)
        mats.append(b2scene.default_material(base_color=[0.0, 0.0, 0.0], emission=tint, emission_strength=float(power * 4.0)))
    wall_material = 0

    big_v, big_f = icosphere(p["big_level"])
    em_v, em_f = icosphere(p["emitter_level"])

    positions, normals, indices, mat_idx = [], [], [], []
    base = 0

    def add(v, n, f, m):
        nonlocal base
        positions.append(v.astype(np.float32))
        normals.append(n.astype(np.float32))
        indices.append((f + base).astype(np.int32))
        mat_idx.append(np.asarray(m, np.int32))
        base += len(v)

    # closed box, normals facing inwards (2 triangles per face)
    h = half
    box_v = np.array([[-h, -h, -h], [h, -h, -h], [h, h, -h], [-h, h, -h], [-h, -h, h], [h, -h, h], [h, h, h], [-h, h, h]], np.float32)
    box_f = np.array([[0, 1, 2], [0, 2, 3], [5, 4, 7], [5, 7, 6], [4, 0, 3], [4, 3, 7], [1, 5, 6], [1, 6, 2], [3, 2, 6], [3, 6, 7], [4, 5, 1], [4, 1, 0]], np.int32)
    add(box_v, np.zeros_like(box_v), box_f, np.full(12, wall_material))
    box_has_normals = 8

    # filler: a strip of quads just above the floor so that the total is exactly n_triangles
    fq = p["filler_quads"]
    if fq:
        xs = np.linspace(-h * 0.9, h * 0.9, fq + 1)
        y = -h + 0.01
        fv = np.concatenate([np.stack([xs, np.full_like(xs, y), np.full_like(xs, -h * 0.9)], 1), np.stack([xs, np.full_like(xs, y), np.full_like(xs, h * 0.9)], 1)])
        i0 = np.arange(fq)
        ff = np.concatenate([np.stack([i0, i0 + fq + 1, i0 + 1], 1), np.stack([i0 + 1, i0 + fq + 1, i0 + fq + 2], 1)])
        add(fv, np.zeros_like(fv), ff, np.full(2 * fq, 1))

    n_flat_vertices = base

    # the big spheres, vectorised: (n, V, 3)
    n_big = p["n_big"]
    sphere_mats = rng.integers(0, 16, n_big)
    if n_big:
        chunk = 256
        for s in range(0, n_big, chunk):
            c = centers[s:s + chunk].astype(np.float32)
            r = radii[s:s + chunk].astype(np.float32)
            k = len(c)
            v = (big_v[None] * r[:, None, None] + c[:, None, :]).reshape(-1, 3)
            n = np.broadcast_to(big_v[None], (k, len(big_v), 3)).reshape(-1, 3)
            f = (big_f[None] + (np.arange(k) * len(big_v))[:, None, None]).reshape(-1, 3)
            add(v, n, f, np.repeat(sphere_mats[s:s + chunk], len(big_f)))

    emitter_mats = 16 + rng.integers(0, n_emissive_mats, p["n_emitters"])
    if p["n_emitters"]:
        c = e_centers.astype(np.float32)
        r = e_radii.astype(np.float32)
        k = len(c)
        v = (em_v[None] * r[:, None, None] + c[:, None, :]).reshape(-1, 3)
        n = np.broadcast_to(em_v[None], (k, len(em_v), 3)).reshape(-1, 3)
        f = (em_f[None] + (np.arange(k) * len(em_v))[:, None, None]).reshape(-1, 3)
        add(v, n, f, np.repeat(emitter_mats, len(em_f)))

    positions = np.concatenate(positions)
    normals = np.concatenate(normals)
    indices = np.concatenate(indices)
    mat_idx = np.concatenate(mat_idx)
    has_normals = np.ones(len(positions), np.uint8)
    has_normals[:n_flat_vertices] = 0
    texcoords = np.zeros((len(positions), 2), np.float32)
    assert len(indices) == n_triangles, (len(indices), n_triangles)

    hfov = 2.0 * np.arctan(np.tan(np.float32(0.4039) * 0.5) * aspect)  # the Cornell file's yfov, same +0.425 rad quirk downstream
    camera = b2scene.make_camera(cam_pos, [0.0, 0.0, 0.0, 1.0], hfov, aspect, clip_near=0.1, clip_far=1000.0)
    info = dict(p, cells=cells, half_extent=half, n_vertices=int(len(positions)), n_triangles=int(len(indices)), n_materials=len(mats),
                emissive_triangles=int(p["n_emitters"] * len(em_f)), vertices="shared per sphere", seed=seed)
    return dict(camera=camera, positions=positions, normals=normals, texcoords=texcoords, indices=indices, material_indices=mat_idx,
                has_normals=has_normals, materials=np.stack(mats)), info


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("out")
    ap.add_argument("--triangles", type=int, default=10_000_000)
    ap.add_argument("--emissive", type=int, default=100_000)
    ap.add_argument("--seed", type=lambda s: int(s, 0), default=0xB200)
    ap.add_argument("--aspect", type=float, default=16.0 / 9.0)
    a = ap.parse_args()
    scene, info = build(a.triangles, a.emissive, a.seed, a.aspect)
    b2scene.write_scene(a.out, scene["camera"], scene["positions"], scene["normals"], scene["texcoords"], scene["indices"],
                        scene["material_indices"], scene["has_normals"], scene["materials"])
    import json
    print(json.dumps(info))


if __name__ == "__main__":
    main()
