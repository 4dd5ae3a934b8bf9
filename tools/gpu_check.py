#!/usr/bin/env python3
"""GPU-box helper: render a config with the CUDA path, run the reference oracle (oracle/_ref) and the port oracle
(oracle/_port) on the same blob on the box's host cores, and report parity + timing. Writes gpurun_out/gpu_check_*.json.

    python tools/gpu_check.py --variant c1 --scene cornell_pbr_1x1 --width 512 --height 512 --spp 1 --bounces 3
"""
import argparse
import json
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))
import __graft_entry__ as entry  # noqa: E402
from compare_dumps import compare  # noqa: E402


def scene_path(a):
    return Path(a.scene) if a.scene.endswith(".b2scene") else ROOT / "data" / f"{a.scene}.b2scene"


def run_oracle(exe, a, prefix, extra=()):
    extra = list(extra)
    if a.window:
        extra += ["--window", *[str(v) for v in a.window]]
    if a.envmap:
        extra += ["--envmap", a.envmap]
    cmd = [str(exe), "--scene", str(scene_path(a)), "--luts", str(ROOT / "data" / "brdf_luts.b2luts"),
           "--width", str(a.width), "--height", str(a.height), "--spp", str(a.spp), "--bounces", str(a.bounces), "--dump", prefix, *extra]
    t0 = time.time()
    out = subprocess.run(cmd, check=True, capture_output=True, text=True)
    return json.loads(out.stdout.strip().splitlines()[-1]), time.time() - t0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--variant", default="c1")
    ap.add_argument("--scene", default="cornell_pbr_1x1")
    ap.add_argument("--width", type=int, default=512)
    ap.add_argument("--height", type=int, default=512)
    ap.add_argument("--spp", type=int, default=1)
    ap.add_argument("--bounces", type=int, default=3)
    ap.add_argument("--no-oracle", action="store_true")
    ap.add_argument("--fused", type=int, default=1)
    ap.add_argument("--directional", action="store_true", help="ReSTIR DI adaptive directional spatial reuse (+ its prepass)")
    ap.add_argument("--tag", default="")
    ap.add_argument("--window", type=int, nargs=4, default=None, help="x0 y0 x1 y1: the oracles render (and parity is checked on) this crop only")
    ap.add_argument("--envmap", default=None, help="B2ENVMP1 blob")
    ap.add_argument("--margin", type=int, default=0, help="compare only the window shrunk by this many pixels: ReSTIR reuse reads neighbours (radius 16 per frame) "
                    "that a windowed oracle run does not have outside its window")
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()

    entry.ensure_data()
    b2pt = entry.load_b2pt()
    ctx = b2pt.Context(0)
    props = ctx.props()
    report = dict(args=vars(a), device=props.name.decode(), sms=props.sm_count)
    tmp = Path(tempfile.mkdtemp())
    n = a.width * a.height

    for flavour in ("precise", "fast"):
        r = b2pt.Renderer(ctx, a.width, a.height, a.variant, precise=(flavour == "precise"))
        r.set_scene_file(scene_path(a))
        if a.envmap:
            r.set_envmap_file(a.envmap)
        r.set_luts_file(ROOT / "data" / "brdf_luts.b2luts")
        r.set(nb_bounces=a.bounces, restir_di_do_fused_spatiotemporal=a.fused)
        if a.directional:
            r.set(restir_di_use_adaptive_directional_spatial_reuse=1)
        if flavour == "precise":
            report["scene"] = r.scene_info()
        # counters pass (separate instrumented kernels)
        r.set_counters(True)
        ctx.counters_reset()
        r.render(a.spp)
        report[f"{flavour}_counters"] = ctx.counters_read()
        r.set_counters(False)
        # timed passes: warm-up then 5 repeats of the whole spp loop from a reset state
        times = []
        for rep in range(a.reps + 1):
            r.reset()
            r.set_host_seed(42)
            r.render(a.spp)
            times.append(r.last_render_ms())
        report[f"{flavour}_ms"] = times[1:]
        report[f"{flavour}_mpaths_per_s"] = n * a.spp / (min(times[1:]) * 1e-3) / 1e6
        # parity dump: the oracles dump the G-buffer of the FIRST sample, the framebuffer after the last one
        r.reset()
        r.set_host_seed(42)
        r.render(1)
        g = r.gbuffer()
        if a.spp > 1:
            r.render(a.spp - 1)
        fb = r.framebuffer()
        report[f"{flavour}_faults"] = ctx.faults()
        prefix = str(tmp / f"cuda_{flavour}_")
        g["prim"].tofile(prefix + "gbuffer_prim.bin")
        g["position"].tofile(prefix + "gbuffer_position.bin")
        g["geometric_normal"].tofile(prefix + "gbuffer_geometric_normal.bin")
        g["shading_normal"].tofile(prefix + "gbuffer_shading_normal.bin")
        g["material"].tofile(prefix + "gbuffer_material.bin")
        fb.tofile(prefix + "framebuffer.bin")
        r.close()

    cmp_window = a.window
    if a.window and a.margin:
        cmp_window = [a.window[0] + a.margin, a.window[1] + a.margin, a.window[2] - a.margin, a.window[3] - a.margin]
    if not a.no_oracle:
        extra = [] if a.fused else ["--no-fused"]
        if a.directional:
            extra.append("--directional-reuse")
        ref_meta, _ = run_oracle(ROOT / "oracle" / "_ref" / f"ref_render_{a.variant}", a, str(tmp / "ref_"), extra)
        port_meta, _ = run_oracle(ROOT / "oracle" / "_port" / "port_render", a, str(tmp / "port_"), ["--variant", a.variant, *extra])
        report["ref_oracle"] = ref_meta
        report["port_oracle"] = port_meta
        report["cuda_precise_vs_ref"] = compare(str(tmp / "ref_"), str(tmp / "cuda_precise_"), a.width, a.height, verbose=False, window=cmp_window)
        report["cuda_fast_vs_ref"] = compare(str(tmp / "ref_"), str(tmp / "cuda_fast_"), a.width, a.height, verbose=False, window=cmp_window)
        report["cuda_precise_vs_port"] = compare(str(tmp / "port_"), str(tmp / "cuda_precise_"), a.width, a.height, verbose=False, window=cmp_window)
        report["port_vs_ref"] = compare(str(tmp / "ref_"), str(tmp / "port_"), a.width, a.height, verbose=False, window=cmp_window)

    out_dir = ROOT / "gpurun_out"
    out_dir.mkdir(exist_ok=True)
    name = f"gpu_check_{a.variant}_{a.width}x{a.height}_{a.spp}spp{a.tag}.json"
    (out_dir / name).write_text(json.dumps(report, indent=1))
    print(json.dumps(report, indent=1))
    ctx.close()


if __name__ == "__main__":
    main()
