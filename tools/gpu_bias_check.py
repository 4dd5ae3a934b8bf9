#!/usr/bin/env python3
"""Minimal GPU check of the ReSTIR DI bias-correction variants (c4b<N> fixtures): CUDA precise vs golden bit-identity, fast
flavour image mean. No torch import: ctypes + numpy only, so it fits in a few seconds of box time."""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import __graft_entry__ as entry  # noqa: E402

b2pt = entry.load_b2pt()
ctx = b2pt.Context(0)
out = {}
for path in sorted((ROOT / "tests" / "golden").glob("c4b*.npz")):
    g = np.load(path)
    meta = json.loads(bytes(g["meta"]).decode())
    extra = meta["extra"]
    for precise in (True, False):
        r = b2pt.Renderer(ctx, meta["width"], meta["height"], meta["name"].split("_")[0], precise=precise)
        r.set_scene_file(ROOT / "data" / f"{meta['scene']}.b2scene")
        r.set_luts_file(ROOT / "data" / "brdf_luts.b2luts")
        r.set(nb_bounces=meta["bounces"], restir_di_do_fused_spatiotemporal=0)
        if "--no-confidence-weights" in extra:
            r.set(restir_di_use_confidence_weights=0)
        if "--spatial-passes" in extra:
            r.set(restir_di_number_of_passes=int(extra[extra.index("--spatial-passes") + 1]))
        if "--envmap" in extra:
            r.set_envmap_file(ROOT / extra[extra.index("--envmap") + 1])
        r.render(meta["spp"])
        fb = np.asarray(r.framebuffer(), np.float32).reshape(-1, 3)
        r.close()
        ref = g["framebuffer"].reshape(-1, 3)
        differing = int((fb.view(np.uint32) != ref.view(np.uint32)).any(axis=1).sum())
        out[f"{meta['name']}:{'precise' if precise else 'fast'}"] = dict(differing=differing, mean=float(fb.mean()), ref_mean=float(ref.mean()))
        print(meta["name"], "precise" if precise else "fast", differing, float(fb.mean()), float(ref.mean()), flush=True)
(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / "bias_check.json").write_text(json.dumps(out, indent=1))
