#!/usr/bin/env python3
"""Step-by-step GPU bring-up check with a watchdog: prints each stage (unbuffered) and dumps the Python stack if a stage
hangs, so a stuck CUDA call or kernel is located in one short gpurun call.   python -u tools/gpu_diag.py [variant] [w] [h] [spp]"""
import faulthandler
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
faulthandler.enable()
faulthandler.dump_traceback_later(90, repeat=True, exit=False)


def say(*a):
    print(f"[{time.time() - T0:7.2f}s]", *a, flush=True)


T0 = time.time()
variant = sys.argv[1] if len(sys.argv) > 1 else "c1"
w = int(sys.argv[2]) if len(sys.argv) > 2 else 64
h = int(sys.argv[3]) if len(sys.argv) > 3 else 64
spp = int(sys.argv[4]) if len(sys.argv) > 4 else 1
import __graft_entry__ as entry  # noqa: E402

say("ensure_data")
entry.ensure_data()
say("load binding")
b2pt = entry.load_b2pt()
say("load library", b2pt.lib().b2pt_version_string())
ctx = b2pt.Context(0)
p = ctx.props()
say("context", p.name.decode(), "SMs", p.sm_count)
import ctypes as C  # noqa: E402

# stage A: the two kernels one by one through the device-layer ABI (b2pt_kernel_get / b2pt_kernel_launch)
r = b2pt.Renderer(ctx, w, h, variant, precise=True)
r.set_scene_file(ROOT / "data" / "cornell_pbr_16x9.b2scene")
r.set_luts_file(ROOT / "data" / "brdf_luts.b2luts")
r.set(nb_bounces=3)
lib = b2pt.lib()
vkey = dict(c1=(1, 1, 0), c2=(0, 3, 1), c3=(0, 4, 1), c4=(0, 5, 1))[variant]
for entry_name in ("CameraRays", "MegaKernel"):
    opts = (b2pt.Option * 4)((b"BSDFOverride", vkey[0]), (b"DirectLightSamplingStrategy", vkey[1]), (b"DirectLightSamplingBaseStrategy", vkey[2]), (b"B2PT_PRECISE", 1))
    k = C.c_void_p()
    b2pt.check(lib.b2pt_kernel_get(ctx.handle, entry_name.encode(), opts, 4, C.byref(k)))
    args = (C.c_void_p * 1)(r._rd)
    say("launch", entry_name)
    b2pt.check(lib.b2pt_kernel_launch(k, 8, 8, w, h, args, None))
    ms = C.c_float()
    b2pt.check(lib.b2pt_kernel_last_ms(k, C.byref(ms)))
    say(entry_name, "done", ms.value, "ms; faults", ctx.faults())
r.close()

for precise in (True, False):
    r = b2pt.Renderer(ctx, w, h, variant, precise=precise)
    say("renderer created precise=", precise)
    r.set_scene_file(ROOT / "data" / "cornell_pbr_16x9.b2scene")
    say("scene set", r.scene_info())
    r.set_luts_file(ROOT / "data" / "brdf_luts.b2luts")
    say("luts set")
    r.set(nb_bounces=3)
    r.render(spp, sync=False)
    say("render enqueued")
    r.synchronize()
    say("render done", r.last_render_ms(), "ms")
    fb = r.framebuffer()
    say("framebuffer mean", float(fb.mean()), "max", float(fb.max()))
    g = r.gbuffer()
    say("gbuffer hits", int((g["prim"] >= 0).sum()), "of", w * h, "faults", ctx.faults())
    r.close()
ctx.close()
say("OK")
