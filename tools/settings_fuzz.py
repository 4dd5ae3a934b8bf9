#!/usr/bin/env python3
"""Settings fuzzing on the host: random values for the run-time settings the ReSTIR DI / path-tracing code reads (fields of
HIPRTRenderData, not macros) are written at the same byte offsets into the reference's and the port's render data (--poke, both structs
are byte-identical), and the two framebuffers must agree bit for bit. The reference is the zero-initialised build (oracle/_ref/ref_zero_*),
so that its reads of unwritten stack variables do not show up as noise.

    python tools/settings_fuzz.py [--variant c4] [--cases 24] [--seed 1] [--scene synthetic_20k]

Prints one line per case and a JSON summary; exit code 1 if any case differs."""
import argparse
import json
import random
import subprocess
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent

# member of b2pt_render_data -> (kind, candidate values); offsets come from offsetof() on include/b2pt_render_data.h
RS = "render_settings."
DI = RS + "restir_di_settings."
FIELDS = {
    RS + "nb_bounces": ("i", [0, 1, 2, 3, 6]),
    RS + "do_russian_roulette": ("b", [0, 1]),
    RS + "russian_roulette_min_depth": ("i", [0, 1, 3]),
    RS + "russian_roulette_throughput_clamp": ("f", [1.0, 4.0, 10.0]),
    RS + "path_russian_roulette_method": ("i", [0, 1]),
    RS + "enable_direct": ("b", [0, 1, 1]),
    RS + "direct_contribution_clamp": ("f", [0.0, 5.0]),
    RS + "envmap_contribution_clamp": ("f", [0.0, 3.0]),
    RS + "indirect_contribution_clamp": ("f", [0.0, 5.0]),
    RS + "minimum_light_contribution": ("f", [0.0, 0.01, 0.2]),
    RS + "ris_settings.number_of_light_candidates": ("i", [1, 4, 7]),
    RS + "ris_settings.number_of_bsdf_candidates": ("i", [0, 1, 2]),
    DI + "common_temporal_pass.do_temporal_reuse_pass": ("b", [0, 1, 1]),
    DI + "common_temporal_pass.use_permutation_sampling": ("b", [0, 1]),
    DI + "common_temporal_pass.max_neighbor_search_count": ("i", [0, 4, 8]),
    DI + "common_temporal_pass.neighbor_search_radius": ("i", [1, 4]),
    DI + "common_spatial_pass.do_spatial_reuse_pass": ("b", [0, 1, 1]),
    DI + "common_spatial_pass.number_of_passes": ("i", [1, 2, 3]),
    DI + "common_spatial_pass.reuse_radius": ("i", [0, 4, 16, 40]),
    DI + "common_spatial_pass.reuse_neighbor_count": ("i", [0, 1, 5, 9]),
    DI + "common_spatial_pass.do_disocclusion_reuse_boost": ("b", [0, 1]),
    DI + "common_spatial_pass.disocclusion_reuse_count": ("i", [3, 5, 8]),
    DI + "common_spatial_pass.debug_neighbor_location": ("b", [0, 0, 1]),
    DI + "common_spatial_pass.debug_neighbor_location_direction": ("i", [0, 1, 2]),
    DI + "common_spatial_pass.do_neighbor_rotation": ("b", [0, 1]),
    DI + "common_spatial_pass.coalesced_spatial_reuse": ("b", [0, 1]),
    DI + "common_spatial_pass.do_visibility_only_last_pass": ("b", [0, 1]),
    DI + "common_spatial_pass.neighbor_visibility_count": ("i", [0, 2, 5]),
    DI + "neighbor_similarity_settings.use_normal_similarity_heuristic": ("b", [0, 1]),
    DI + "neighbor_similarity_settings.normal_similarity_angle_precomp": ("f", [0.5, 0.9, 0.99]),
    DI + "neighbor_similarity_settings.reject_using_geometric_normals": ("b", [0, 1]),
    DI + "neighbor_similarity_settings.use_plane_distance_heuristic": ("b", [0, 1]),
    DI + "neighbor_similarity_settings.plane_distance_threshold": ("f", [0.01, 0.1, 1.0]),
    DI + "neighbor_similarity_settings.use_roughness_similarity_heuristic": ("b", [0, 1]),
    DI + "neighbor_similarity_settings.roughness_similarity_threshold": ("f", [0.05, 0.25]),
    DI + "m_cap": ("i", [0, 3, 20]),
    DI + "use_confidence_weights": ("b", [0, 1]),
    DI + "symmetric_ratio_mis_weights_beta_exponent": ("f", [2.0, 3.0, 4.0, 2.5]),
    DI + "initial_candidates.number_of_initial_light_candidates": ("i", [0, 1, 4, 6]),
    DI + "initial_candidates.number_of_initial_bsdf_candidates": ("i", [0, 1, 2]),
    DI + "initial_candidates.envmap_candidate_probability": ("f", [0.0, 0.25, 0.7]),
    DI + "do_fused_spatiotemporal": ("b", [0, 1]),
    DI + "do_final_shading_visibility": ("b", [0, 1]),
}


def offsets(tmp):
    names = sorted(FIELDS)
    lines = "\n".join(f'    std::printf("{n} %zu\\n", offsetof(b2pt_render_data, {n}));' for n in names)
    src = Path(tmp) / "off.cpp"
    src.write_text(f'#include <cstdio>\n#include <cstddef>\n#include "b2pt_render_data.h"\nint main()\n{{\n{lines}\n    return 0;\n}}\n')
    subprocess.run(["g++", "-w", f"-I{ROOT / 'include'}", str(src), "-o", str(Path(tmp) / "off")], check=True)
    out = subprocess.run([str(Path(tmp) / "off")], check=True, capture_output=True, text=True).stdout
    return {l.split()[0]: int(l.split()[1]) for l in out.splitlines()}


def render(exe, variant, scene, pokes, prefix, extra):
    cmd = [*exe, "--scene", str(ROOT / "data" / f"{scene}.b2scene"), "--luts", str(ROOT / "data" / "brdf_luts.b2luts"), "--width", "96", "--height", "54",
           "--spp", "3", "--bounces", "4", "--dump", prefix, *extra]
    for p in pokes:
        cmd += ["--poke", p]
    res = subprocess.run(cmd, capture_output=True, text=True, cwd=ROOT)
    if res.returncode != 0:
        return None
    return np.fromfile(prefix + "framebuffer.bin", np.uint32).reshape(-1, 3)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--variant", default="c4")
    ap.add_argument("--cases", type=int, default=24)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--scene", default="synthetic_20k")
    ap.add_argument("--fields", type=int, default=8, help="settings changed per case")
    ap.add_argument("--ref-exe", default="", help="path of a reference binary to use instead (e.g. a one-off zero-initialised build of another variant)")
    ap.add_argument("--reference", default="zero", choices=["zero", "render"], help="ref_zero_<variant> (ReSTIR DI variants) or the plain ref_render_<variant>")
    a = ap.parse_args()
    rng = random.Random(a.seed)
    ref_exe = Path(a.ref_exe) if a.ref_exe else ROOT / "oracle" / "_ref" / f"ref_{a.reference}_{a.variant}"
    port_exe = ROOT / "oracle" / "_port" / "port_render"
    if not ref_exe.exists() or not port_exe.exists():
        print(f"{ref_exe} or {port_exe} not built")
        return 2
    extra = ["--envmap", "data/sky_128x64.b2env"] if a.scene == "synthetic_20k" else []
    failures, crashed = [], 0
    with tempfile.TemporaryDirectory() as tmp:
        off = offsets(tmp)
        for case in range(a.cases):
            chosen = rng.sample(sorted(FIELDS), a.fields)
            pokes, described = [], {}
            for name in chosen:
                kind, values = FIELDS[name]
                value = rng.choice(values)
                pokes.append(f"{off[name]}:{kind}:{value}")
                described[name.split(".")[-1]] = value
            ref = render([str(ref_exe)], a.variant, a.scene, pokes, str(Path(tmp) / "r_"), extra)
            port = render([str(port_exe), "--variant", a.variant], a.variant, a.scene, pokes, str(Path(tmp) / "p_"), extra)
            if ref is None:
                crashed += 1  # the reference itself does not survive these settings: not a case
                print(f"case {case}: reference crashed {described}")
                continue
            differing = -1 if port is None else int((ref != port).any(axis=1).sum())
            print(f"case {case}: differing {differing} {described}", flush=True)
            if differing != 0:
                failures.append(dict(case=case, differing=differing, settings=described, pokes=pokes))
    print(json.dumps(dict(settings_fuzz=True, variant=a.variant, cases=a.cases, reference_crashed=crashed, failures=failures)))
    return 1 if failures else 0


if __name__ == "__main__":
    sys.exit(main())
