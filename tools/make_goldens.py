#!/usr/bin/env python3
"""Generates tests/golden/*.npz from the REFERENCE oracle (oracle/_ref, i.e. the reference's own CPU kernels compiled
in place). Only runs where /root/reference was available to build oracle/_ref; the fixtures are committed."""
import json
import subprocess
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
GOLD = ROOT / "tests" / "golden"

CASES = [
    # name, variant, scene, w, h, spp, bounces, extra args
    ("c1_128x72", "c1", "cornell_pbr_16x9", 128, 72, 1, 3, []),
    ("c1_64x64_4spp", "c1", "cornell_pbr_1x1", 64, 64, 4, 5, []),
    ("c2_128x72", "c2", "cornell_pbr_16x9", 128, 72, 2, 8, []),
    ("c3_128x72", "c3", "cornell_pbr_16x9", 128, 72, 2, 8, []),
    ("c4_96x54_fused", "c4", "cornell_pbr_16x9", 96, 54, 3, 4, []),
    ("c4_96x54_split", "c4", "cornell_pbr_16x9", 96, 54, 3, 4, ["--no-fused"]),
    # many-light synthetic scene (tools/make_synthetic_scene.py --triangles 20000 --emissive 1600: 57 spheres with metallic /
    # coated / sheen materials, 20 emitters over 3 decades of power) and the procedural HDR sky (tools/make_envmap.py):
    # RIS over many lights, envmap alias-table sampling, ReSTIR DI with envmap candidates
    ("c2_syn_96x54", "c2", "synthetic_20k", 96, 54, 2, 4, []),
    ("c3_syn_96x54", "c3", "synthetic_20k", 96, 54, 2, 4, []),
    ("c3_env_96x54", "c3", "cornell_pbr_16x9", 96, 54, 2, 4, ["--envmap", "data/sky_128x64.b2env"]),
    ("c4_synenv_96x54_fused", "c4", "synthetic_20k", 96, 54, 3, 4, ["--envmap", "data/sky_128x64.b2env"]),
    # (Cornell + envmap + split temporal/spatial passes is not a usable case: at the 2nd frame the REFERENCE produces a NaN
    #  reservoir at pixel [54, 47], indexes emissive_triangles[-1] and segfaults; the fused variant below is fine)
    ("c4_env_96x54_fused", "c4", "cornell_pbr_16x9", 96, 54, 3, 4, ["--envmap", "data/sky_128x64.b2env"]),
    # ReSTIR_DI_DoLightPresampling: the ReSTIR_DI_LightsPresampling kernel + presampled initial light candidates (variant "c4p")
    ("c4p_synenv_96x54_fused", "c4p", "synthetic_20k", 96, 54, 3, 4, ["--envmap", "data/sky_128x64.b2env"]),
    ("c4p_cornell_96x54_split", "c4p", "cornell_pbr_16x9", 96, 54, 2, 4, ["--no-fused"]),
    # adaptive-directional spatial reuse: the ReSTIR_Directional_Reuse_Compute prepass (per-pixel radius + 64-bit direction
    # mask) and the spatial neighbour sampling restricted to the allowed directions
    ("c4_syn_96x54_directional_fused", "c4", "synthetic_20k", 96, 54, 3, 4, ["--directional-reuse"]),
    ("c4_syn_96x54_directional_split", "c4", "synthetic_20k", 96, 54, 2, 4, ["--directional-reuse", "--no-fused"]),
    # textured Cornell box (tools/make_textured_scene.py): sRGB base colour, normal map, packed roughness/metallic, single-channel
    # parameter maps, emission texture + constant-emission marker. Rendered by oracle/_ref/ref_extras_* (the harness binds the blob's
    # textures as the reference's Image8Bit scene textures only under -DB2_REF_EXTRAS, so that ref_render_* stays what made the other
    # fixtures). ReSTIR DI needs the envmap here: without one the reference dereferences a null envmap through its uninitialised
    # sampled_lobe_info on sheen-lobe samples and segfaults.
    ("c2_tex_96x54", "c2", "cornell_textured_16x9", 96, 54, 2, 4, []),
    ("c3_tex_96x54", "c3", "cornell_textured_16x9", 96, 54, 2, 4, []),
    ("c4_texenv_96x54_fused", "c4", "cornell_textured_16x9", 96, 54, 3, 4, ["--envmap", "data/sky_128x64.b2env"]),
    # spectral dispersion (tools/make_dispersion_scene.py): wavelength sampling, wavelength-dependent IOR in the glass lobe, the
    # wavelength's colour filter on the throughput; in the fused ReSTIR DI pass also the wavelength draws of every surface rebuild
    ("c2_disp_96x54", "c2", "cornell_dispersion_16x9", 96, 54, 3, 6, []),
    ("c3_disp_96x54", "c3", "cornell_dispersion_16x9", 96, 54, 3, 6, []),
    ("c4_disp_96x54_fused", "c4", "cornell_dispersion_16x9", 96, 54, 3, 6, []),
    # ReSTIR_DI_BiasCorrectionWeights other than pairwise MIS (variants "c4b<N>": 1 = 1/Z, 2 = MIS-like, 3 = generalised balance
    # heuristic, 5 = defensive pairwise MIS, 6 / 7 below), temporal + spatial passes, with and without confidence weights, one and two spatial passes
    ("c4b1_96x54_split", "c4b1", "cornell_pbr_16x9", 96, 54, 3, 4, ["--no-fused"]),
    ("c4b2_synenv_96x54_split", "c4b2", "synthetic_20k", 96, 54, 3, 4, ["--no-fused", "--envmap", "data/sky_128x64.b2env"]),
    ("c4b3_96x54_split_noconf", "c4b3", "cornell_pbr_16x9", 96, 54, 3, 4, ["--no-fused", "--no-confidence-weights"]),
    ("c4b3_syn_96x54_split", "c4b3", "synthetic_20k", 96, 54, 2, 4, ["--no-fused"]),
    ("c4b5_96x54_split2", "c4b5", "cornell_pbr_16x9", 96, 54, 3, 4, ["--no-fused", "--spatial-passes", "2"]),
    ("c4b5_syn_96x54_split_noconf", "c4b5", "synthetic_20k", 96, 54, 2, 4, ["--no-fused", "--no-confidence-weights"]),
    # 6 = symmetric ratio, 7 = asymmetric ratio
    ("c4b6_96x54_split", "c4b6", "cornell_pbr_16x9", 96, 54, 3, 4, ["--no-fused"]),
    ("c4b6_synenv_96x54_split_noconf", "c4b6", "synthetic_20k", 96, 54, 2, 4, ["--no-fused", "--no-confidence-weights", "--envmap", "data/sky_128x64.b2env"]),
    ("c4b7_96x54_split2_noconf", "c4b7", "cornell_pbr_16x9", 96, 54, 3, 4, ["--no-fused", "--no-confidence-weights", "--spatial-passes", "2"]),
    ("c4b7_syn_96x54_split", "c4b7", "synthetic_20k", 96, 54, 2, 4, ["--no-fused"]),
    # the fused spatiotemporal pass with those weights, and with more neighbours than its phase-cut form holds (both run the pass in
    # one piece, restir_di_spatiotemporal_inplace_pixel)
    ("c4b1_96x54_fused", "c4b1", "cornell_pbr_16x9", 96, 54, 3, 4, []),
    ("c4b2_synenv_96x54_fused", "c4b2", "synthetic_20k", 96, 54, 3, 4, ["--envmap", "data/sky_128x64.b2env"]),
    ("c4b3_96x54_fused_noconf", "c4b3", "cornell_pbr_16x9", 96, 54, 3, 4, ["--no-confidence-weights"]),
    ("c4b5_syn_96x54_directional_fused", "c4b5", "synthetic_20k", 96, 54, 3, 4, ["--directional-reuse"]),
    ("c4b6_96x54_fused", "c4b6", "cornell_pbr_16x9", 96, 54, 3, 4, []),
    ("c4_96x54_fused_9neighbors", "c4", "cornell_pbr_16x9", 96, 54, 3, 4, ["--neighbors", "9"]),
    ("c4b1_synenv_96x54_fused_12neighbors", "c4b1", "synthetic_20k", 96, 54, 2, 4, ["--neighbors", "12", "--envmap", "data/sky_128x64.b2env"]),
    # the four ReSTIR DI visibility macros (ReSTIRDIOptions.h:50-84) away from their defaults, variants "c4v<N>": 1 = InitialTargetFunctionVisibility
    # on, 2 = SpatialTargetFunctionVisibility on, 3 = DoVisibilityReuse off, 4 = BiasCorrectionUseVisibility off, 5 = 1 + 2 + 3.
    # (DoVisibilityReuse off on the Cornell box: the reference puts 2e30 into the pixels that look at the lamp; the port follows bit for
    #  bit, but an image mean of that says nothing about the fast flavour, so 3 and 5 use the many-light scene)
    ("c4v1_96x54_fused", "c4v1", "cornell_pbr_16x9", 96, 54, 3, 4, []),
    ("c4v1_syn_96x54_split", "c4v1", "synthetic_20k", 96, 54, 2, 4, ["--no-fused"]),
    ("c4v2_96x54_fused", "c4v2", "cornell_pbr_16x9", 96, 54, 3, 4, []),
    ("c4v2_96x54_split", "c4v2", "cornell_pbr_16x9", 96, 54, 3, 4, ["--no-fused"]),
    ("c4v3_synenv_96x54_fused", "c4v3", "synthetic_20k", 96, 54, 3, 4, ["--envmap", "data/sky_128x64.b2env"]),
    ("c4v3_syn_96x54_split", "c4v3", "synthetic_20k", 96, 54, 2, 4, ["--no-fused"]),
    ("c4v4_96x54_fused", "c4v4", "cornell_pbr_16x9", 96, 54, 3, 4, []),
    ("c4v4_synenv_96x54_split", "c4v4", "synthetic_20k", 96, 54, 2, 4, ["--no-fused", "--envmap", "data/sky_128x64.b2env"]),
    ("c4v5_synenv_96x54_fused", "c4v5", "synthetic_20k", 96, 54, 3, 4, ["--envmap", "data/sky_128x64.b2env"]),
    # ReSTIR_DI_LaterBouncesSamplingStrategy (ReSTIRDIOptions.h:18-21,140), variants "c4l<N>": 0 = uniform one light, 1 = BSDF only,
    # 2 = light + BSDF MIS at the bounces after the first (3 = RIS, the default, is c4)
    ("c4l0_96x54_fused", "c4l0", "cornell_pbr_16x9", 96, 54, 3, 4, []),
    ("c4l1_synenv_96x54_fused", "c4l1", "synthetic_20k", 96, 54, 3, 4, ["--envmap", "data/sky_128x64.b2env"]),
    ("c4l2_96x54_split", "c4l2", "cornell_pbr_16x9", 96, 54, 3, 4, ["--no-fused"]),
    ("c4l2_synenv_96x54_fused", "c4l2", "synthetic_20k", 96, 54, 3, 4, ["--envmap", "data/sky_128x64.b2env"]),
    # ReSTIR_DI_LightPresamplingStrategy = LSS_BASE_UNIFORM (variant "c4pu" = c4p with uniform presampling)
    ("c4pu_synenv_96x54_fused", "c4pu", "synthetic_20k", 96, 54, 3, 4, ["--envmap", "data/sky_128x64.b2env"]),
    ("c4pu_96x54_split", "c4pu", "cornell_pbr_16x9", 96, 54, 2, 4, ["--no-fused"]),
    # several options at once (they interact: which passes end with visibility reuse, which flags a sample keeps): "c4x1" = balance
    # heuristic + no bias-correction visibility + spatial target-function visibility + MIS at later bounces; "c4x2" = symmetric ratio +
    # no visibility reuse + initial target-function visibility + BSDF sampling at later bounces
    ("c4x1_96x54_fused", "c4x1", "cornell_pbr_16x9", 96, 54, 3, 4, []),
    ("c4x1_syn_96x54_split_directional", "c4x1", "synthetic_20k", 96, 54, 2, 4, ["--no-fused", "--no-confidence-weights", "--directional-reuse"]),
    ("c4x2_synenv_96x54_fused_8neighbors", "c4x2", "synthetic_20k", 96, 54, 3, 4, ["--envmap", "data/sky_128x64.b2env", "--neighbors", "8"]),
    ("c4x2_syn_96x54_split2", "c4x2", "synthetic_20k", 96, 54, 2, 4, ["--no-fused", "--spatial-passes", "2"]),
    # ReSTIR_DI_DoOptimalVisibilitySampling (variant "c4o"; the spatial pass reads it). The reference needs -Dlast_hit_primitive_index=
    # primitive_index to compile with it (oracle/Makefile: c4o, and c4r = the rename alone, which reproduces c4's fixtures)
    ("c4o_96x54_split", "c4o", "cornell_pbr_16x9", 96, 54, 3, 4, ["--no-fused"]),
    ("c4o_synenv_96x54_split", "c4o", "synthetic_20k", 96, 54, 3, 4, ["--no-fused", "--envmap", "data/sky_128x64.b2env"]),
    ("c4o_96x54_split2_3neighbors", "c4o", "cornell_pbr_16x9", 96, 54, 3, 4, ["--no-fused", "--spatial-passes", "2", "--neighbors", "3"]),
    # 1/M weights (variant "c4b0"): the reference's spatial reuse kernel does not compile with them (SpatialReuse.h:237), its fused and
    # temporal kernels do -- ref_render_c4b0 is built without the spatial kernel and runs the fused pass or temporal reuse alone
    ("c4b0_96x54_fused", "c4b0", "cornell_pbr_16x9", 96, 54, 3, 4, []),
    ("c4b0_synenv_96x54_fused_8neighbors", "c4b0", "synthetic_20k", 96, 54, 3, 4, ["--envmap", "data/sky_128x64.b2env", "--neighbors", "8"]),
    ("c4b0_96x54_temporal_only", "c4b0", "cornell_pbr_16x9", 96, 54, 3, 4, ["--no-fused", "--no-spatial"]),
    ("c4b0_syn_96x54_temporal_only_noconf", "c4b0", "synthetic_20k", 96, 54, 3, 4, ["--no-fused", "--no-spatial", "--no-confidence-weights"]),
]
EXTRAS_SCENES = {"cornell_textured_16x9"}  # cases rendered by ref_extras_* instead of ref_render_*
# Cases rendered by ref_zero_* (the extras harness compiled with -ftrivial-auto-var-init=zero, oracle/Makefile): on the textured Cornell box
# 17 pixels of the plain ReSTIR DI build follow stack variables the reference reads before writing them, and WHICH 17 changes with every
# edit of the harness (the stack moves). The zero-initialised build is stable, and it is the one the product agrees with in every pixel.
ZERO_CASES = {"c4_texenv_96x54_fused"}
# fixtures of kernel options added after the round's GPU budget was spent live in tests/golden/options/ and are tested by
# tests/test_restir_di_options.py, which runs after the suites whose CUDA side has been seen green on a B200
OPTIONS_PREFIXES = ("c4v", "c4l", "c4pu", "c4x", "c4o", "c4b0")


def run_case(name, variant, scene, w, h, spp, bounces, extra):
    binary = "zero" if name in ZERO_CASES else "extras" if scene in EXTRAS_SCENES else "render"
    exe = ROOT / "oracle" / "_ref" / f"ref_{binary}_{variant}"
    if not exe.exists():
        print(f"skip {name}: {exe} not built")
        return
    with tempfile.TemporaryDirectory() as tmp:
        prefix = str(Path(tmp) / "d_")
        out = subprocess.run([str(exe), "--scene", str(ROOT / "data" / f"{scene}.b2scene"), "--luts", str(ROOT / "data" / "brdf_luts.b2luts"),
                              "--width", str(w), "--height", str(h), "--spp", str(spp), "--bounces", str(bounces), "--dump", prefix, *extra],
                             check=True, capture_output=True, text=True, cwd=ROOT)
        meta = json.loads(out.stdout.strip().splitlines()[-1])
        load = lambda n, dt: np.fromfile(prefix + n + ".bin", dt)
        mat = load("gbuffer_material", np.uint32).reshape(-1, 27).copy()
        mat[:, 0] &= 0xFF  # struct padding bytes are indeterminate in the reference
        arrays = dict(
            gbuffer_prim=load("gbuffer_prim", np.int32), gbuffer_position=load("gbuffer_position", np.float32),
            gbuffer_geometric_normal=load("gbuffer_geometric_normal", np.uint32) & 0xFFFFFF,
            gbuffer_shading_normal=load("gbuffer_shading_normal", np.uint32) & 0xFFFFFF,
            gbuffer_material=mat, framebuffer=load("framebuffer", np.float32),
            packed_materials=load("packed_materials", np.uint32), triangle_areas=load("triangle_areas", np.float32),
            emissive_prims=load("emissive_prims", np.int32), power_alias=load("power_alias", np.int32), power_probas=load("power_probas", np.float32),
            meta=np.frombuffer(json.dumps(dict(meta, name=name, scene=scene, extra=extra, oracle_binary=binary)).encode(), np.uint8))
        out_dir = GOLD / "options" if name.startswith(OPTIONS_PREFIXES) else GOLD
        out_dir.mkdir(parents=True, exist_ok=True)
        np.savez_compressed(out_dir / f"{name}.npz", **arrays)
        print(f"wrote {out_dir.name}/{name}.npz ({(out_dir / (name + '.npz')).stat().st_size} bytes)")


if __name__ == "__main__":
    GOLD.mkdir(parents=True, exist_ok=True)
    only = set(sys.argv[1:])
    for case in CASES:
        if not only or case[0] in only:
            run_case(*case)
