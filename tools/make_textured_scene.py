#!/usr/bin/env python3
"""data/cornell_textured_16x9.b2scene: the Cornell box of data/cornell_pbr_16x9.b2scene with procedural 8-bit textures on its
materials -- sRGB base colour, normal map, packed roughness/metallic, single-channel metallic / anisotropy / specular / coat /
sheen / transmission maps, a non-constant emission texture (triangles that emit but are not importance sampled) and the
constant-emission marker -- so that every branch of the reference's get_intersection_material (Material.h:60-150),
normal_mapping (Intersect.h:77-112) and the emissive-texture reads (Intersect.h:596-623) is exercised. Seeded, deterministic.

    python tools/make_textured_scene.py data/cornell_pbr_16x9.b2scene data/cornell_textured_16x9.b2scene
"""
import sys

import numpy as np

import b2scene


def textures(rng, n=48):
    yy, xx = np.mgrid[0:n, 0:n].astype(np.float32) / n
    out = []
    checker = ((np.floor(xx * 6) + np.floor(yy * 6)) % 2).astype(np.float32)
    base = np.stack([0.25 + 0.7 * checker, 0.3 + 0.5 * xx, 0.9 - 0.6 * checker * yy, np.ones_like(xx)], -1)
    out.append(base)                                                                                     # 0 base colour (sRGB), RGBA opaque
    nx, ny = 0.35 * np.sin(xx * 12.0), 0.35 * np.cos(yy * 9.0)
    out.append(np.stack([0.5 + 0.5 * nx, 0.5 + 0.5 * ny, 0.5 + 0.5 * np.sqrt(np.maximum(0.0, 1 - nx * nx - ny * ny)), np.ones_like(xx)], -1))  # 1 normal map
    out.append(np.stack([np.zeros_like(xx), 0.05 + 0.9 * xx, (yy > 0.5).astype(np.float32) * 0.9, np.ones_like(xx)], -1))  # 2 roughness (g) / metallic (b)
    out.append((0.2 + 0.8 * checker)[..., None])                                                        # 3 metallic
    out.append((0.1 + 0.8 * np.abs(np.sin(xx * 7.0)))[..., None])                                       # 4 anisotropic (the reference reads it as roughness too)
    out.append((0.3 + 0.7 * yy)[..., None])                                                             # 5 specular
    out.append((checker * 0.9)[..., None])                                                              # 6 coat
    out.append((0.8 * xx * yy)[..., None])                                                              # 7 sheen
    out.append(((xx + yy) % 1.0 > 0.6).astype(np.float32)[..., None] * 0.85)                            # 8 specular transmission
    spots = np.exp(-60.0 * ((xx - 0.3) ** 2 + (yy - 0.6) ** 2)) + np.exp(-90.0 * ((xx - 0.75) ** 2 + (yy - 0.25) ** 2))
    out.append(np.stack([spots, 0.6 * spots, 0.2 * spots + 0.02 * rng.random(xx.shape, dtype=np.float32), np.ones_like(xx)], -1))  # 9 emission
    return [np.clip(t * 255.0, 0, 255).astype(np.uint8) for t in out]


def main(src, dst, alpha=False):
    rng = np.random.default_rng(0xB200)
    s = b2scene.read_scene(src)
    pos = s["positions"]
    # one affine map for the whole scene: never degenerate on an axis-aligned face, a few repeats across the box
    uv = np.stack([0.37 * pos[:, 0] + 0.21 * pos[:, 2] + 0.11 * pos[:, 1], 0.41 * pos[:, 1] + 0.29 * pos[:, 2] - 0.07 * pos[:, 0]], -1).astype(np.float32)
    n_mat = len(s["materials"])
    idx = np.full((n_mat, len(b2scene.TEX_SLOTS)), b2scene.TEX_NONE, np.int32)
    slot = {n: i for i, n in enumerate(b2scene.TEX_SLOTS)}
    mats = np.array(s["materials"], np.float32).copy()
    M = b2scene.MAT_INDEX
    idx[0, slot["roughness_metallic"]] = 2; idx[0, slot["normal_map"]] = 1
    idx[1, slot["base_color"]] = 0; idx[1, slot["normal_map"]] = 1
    idx[3, slot["base_color"]] = 0; idx[3, slot["coat"]] = 6; idx[3, slot["specular"]] = 5; idx[3, slot["sheen"]] = 7
    mats[3, M["coat_roughness"]] = 0.25; mats[3, M["coat_roughening"]] = 1.0
    idx[4, slot["emission"]] = 9
    idx[5, slot["anisotropic"]] = 4; idx[5, slot["metallic"]] = 3
    idx[6, slot["emission"]] = b2scene.TEX_CONSTANT_EMISSIVE
    mats[6, M["emission_r"]:M["emission_r"] + 3] = (0.25, 0.2, 0.05)
    idx[7, slot["specular_transmission"]] = 8
    idx[8, slot["base_color"]] = 0
    tex = textures(rng)
    if alpha:
        # alpha-tested variant (data/cornell_alpha_16x9.b2scene): a base colour texture whose alpha channel has holes (a lattice), and two
        # materials with alpha_opacity < 1 -- every branch of the reference's any-hit filter (FilterFunction.h:47-74)
        n = 48
        yy, xx = np.mgrid[0:n, 0:n].astype(np.float32) / n
        holes = (((np.floor(xx * 8) + np.floor(yy * 8)) % 2) > 0).astype(np.float32)
        lattice = np.stack([0.8 - 0.5 * xx, 0.35 + 0.5 * yy, 0.3 + 0.4 * holes, 0.15 + 0.85 * holes], -1)
        tex.append(np.clip(lattice * 255.0, 0, 255).astype(np.uint8))                                    # 10 base colour with alpha holes
        idx[1, slot["base_color"]] = 10
        idx[8, slot["base_color"]] = 10
        mats[2, M["alpha_opacity"]] = 0.5
        mats[5, M["alpha_opacity"]] = 0.25
    cam = s["camera"]
    camera = b2scene.camera_block(cam["inverse_view"], cam["inverse_projection"], cam["view_projection"], cam["position"], cam["vertical_fov"])
    b2scene.write_scene(dst, camera, pos, s["normals"], uv, s["indices"], s["material_indices"], s["has_normals"], mats, textures=tex, texture_indices=idx)
    print(f"wrote {dst}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], alpha="--alpha" in sys.argv[3:])
