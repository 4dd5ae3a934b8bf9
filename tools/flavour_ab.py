#!/usr/bin/env python3
"""A/B timing of one library build (B2PT_LIB=...) on bench.py's workloads: device ms per sample of the MegaKernel and of the whole
step, for the flavours the library holds. Used to attribute the precise/fast gap to compiler flags (tools/gpu_runs/gpu_flavours.sh)."""
import argparse
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
import __graft_entry__ as entry  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workloads", default="c2,c3s")
    ap.add_argument("--flavours", default="precise")
    ap.add_argument("--spp", type=int, default=2)
    ap.add_argument("--tag", default="")
    a = ap.parse_args()
    entry.ensure_data()
    b2pt = entry.load_b2pt()
    ctx = b2pt.Context(0)
    out = dict(tag=a.tag, lib=str(b2pt.LIB_PATH))
    for name in a.workloads.split(","):
        w = bench.WORKLOADS[name]
        for flavour in a.flavours.split(","):
            try:
                r = b2pt.Renderer(ctx, w["width"], w["height"], w["variant"], precise=(flavour == "precise"))
            except b2pt.B2ptError as e:
                out[f"{name}_{flavour}"] = f"unavailable: {e}"
                continue
            r.set_scene_file(bench.scene_file(w))
            r.set_luts_file(ROOT / "data" / "brdf_luts.b2luts")
            r.set(nb_bounces=w["bounces"])
            if w["variant"] == "c4":
                r.set(restir_di_do_fused_spatiotemporal=int(w.get("fused", 0)))
            if bench.envmap_file(w):
                r.set_envmap_file(bench.envmap_file(w))
            ms = []
            for i in range(3):
                r.reset()
                r.set_host_seed(42)
                r.render(a.spp, sync=False)
                r.synchronize()
                ms.append(r.last_render_ms())
            r.reset()
            r.set_host_seed(42)
            r.set_profile_passes(True)
            before = r.pass_times()
            r.render(a.spp)
            after = r.pass_times()
            pass_ms = {k: (after[k]["ms"] - before.get(k, dict(ms=0.0))["ms"]) / a.spp for k in after}
            step_ms = min(ms[1:]) / a.spp
            # per-kernel times of the wavefront sequences (libraries older than b2pt_wavefront_profile: skipped)
            wf_ms = None
            try:
                r.set_profile_passes(False)
                r.reset()
                r.set_host_seed(42)
                ctx.wavefront_profile(True)
                r.render(a.spp)
                wf_ms = {k: round(v["ms"] / a.spp, 3) for k, v in ctx.wavefront_profile_read().items() if v["groups"]}
                ctx.wavefront_profile(False)
            except AttributeError:
                pass
            out[f"{name}_{flavour}"] = dict(ms_per_sample=step_ms, mpaths=w["width"] * w["height"] / step_ms / 1e3, pass_ms_per_sample=pass_ms, wavefront_kernel_ms_per_sample=wf_ms,
                                            checksum=float(r.framebuffer().astype("float64").sum()))
            r.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
