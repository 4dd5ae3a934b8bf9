#!/usr/bin/env python3
"""Minimal GPU check of the ReSTIR DI bias-correction variants (c4b<N> fixtures) and of the fused pass with more than 6 neighbours:
CUDA precise vs golden bit-identity, fast flavour image mean. No torch import: ctypes + numpy only (a few seconds of box time)."""
import json
import sys
import traceback
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import __graft_entry__ as entry  # noqa: E402

b2pt = entry.load_b2pt()
ctx = b2pt.Context(0)
out = {}
golden = ROOT / "tests" / "golden"
for path in sorted(list(golden.glob("c4b*.npz")) + list(golden.glob("c4_*neighbors.npz"))):
    g = np.load(path)
    meta = json.loads(bytes(g["meta"]).decode())
    extra = meta["extra"]
    for precise in (True, False):
        key = f"{meta['name']}:{'precise' if precise else 'fast'}"
        try:
            r = b2pt.Renderer(ctx, meta["width"], meta["height"], meta["name"].split("_")[0], precise=precise)
            r.set_scene_file(ROOT / "data" / f"{meta['scene']}.b2scene")
            r.set_luts_file(ROOT / "data" / "brdf_luts.b2luts")
            r.set(nb_bounces=meta["bounces"], restir_di_do_fused_spatiotemporal=0 if "--no-fused" in extra else 1)
            if "--directional-reuse" in extra:
                r.set(restir_di_use_adaptive_directional_spatial_reuse=1)
            if "--no-confidence-weights" in extra:
                r.set(restir_di_use_confidence_weights=0)
            if "--neighbors" in extra:
                n = int(extra[extra.index("--neighbors") + 1])
                r.set(restir_di_reuse_neighbor_count=n, restir_di_disocclusion_reuse_count=n)
            if "--spatial-passes" in extra:
                r.set(restir_di_number_of_passes=int(extra[extra.index("--spatial-passes") + 1]))
            if "--envmap" in extra:
                r.set_envmap_file(ROOT / extra[extra.index("--envmap") + 1])
            r.render(meta["spp"])
            fb = np.asarray(r.framebuffer(), np.float32).reshape(-1, 3)
            r.close()
            ref = g["framebuffer"].reshape(-1, 3)
            differing = int((fb.view(np.uint32) != ref.view(np.uint32)).any(axis=1).sum())
            out[key] = dict(differing_pixels=differing, pixels=int(len(ref)), mean=float(fb.mean()), reference_mean=float(ref.mean()))
        except Exception as e:  # noqa: BLE001
            out[key] = dict(error=str(e), trace=traceback.format_exc())
        print(key, out[key], flush=True)
(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / "bias_check.json").write_text(json.dumps(out, indent=1))
