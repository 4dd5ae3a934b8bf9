"""The drop-in boundary driven the way the reference's host would drive it (minimum slice):
oracle/_ref/boundary_check_<variant> is compiled against the REFERENCE's HostDeviceCommon/RenderData.h, fills HIPRTRenderData with
the reference's own host-side packing, and talks to libb2pt.so through the device-layer C ABI only (b2pt_malloc / memcpy,
b2pt_lut_create, b2pt_bvh_build, b2pt_kernel_get with the reference's macro names, b2pt_kernel_launch(8, 8, w, h, {&render_data})).
Its GPU framebuffer must be bit-identical to the reference CPU path's, rendered in the same process from the same render data."""
import json
import subprocess

import pytest

from conftest import DATA, REF_DIR, ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("variant,scene,extra", [("c1", "cornell_pbr_16x9", []), ("c2", "cornell_pbr_16x9", []), ("c3", "synthetic_20k", []),
                                                  ("c3", "cornell_pbr_16x9", ["--envmap", "data/sky_128x64.b2env"]), ("c4", "cornell_pbr_16x9", []),
                                                  ("c4", "cornell_pbr_16x9", ["--no-fused"])])
def test_reference_typed_host_drives_the_device_layer(built, variant, scene, extra):
    exe = REF_DIR / f"boundary_check_{variant}"
    if not exe.exists():
        pytest.skip(f"{exe} not built (needs /root/reference at build time)")
    out = subprocess.run([str(exe), "--scene", str(DATA / f"{scene}.b2scene"), "--luts", str(DATA / "brdf_luts.b2luts"), "--width", "256", "--height", "144",
                          "--spp", "3", "--bounces", "4", *extra], capture_output=True, text=True, timeout=300, cwd=ROOT)
    lines = [l for l in out.stdout.splitlines() if l.startswith('{"boundary_check"')]
    assert lines, out.stdout[-1500:] + out.stderr[-1500:]
    report = json.loads(lines[-1])
    assert report["render_data_bytes"] == 2784
    assert report["pixels_not_bit_identical"] <= (2 if variant == "c4" else 0), report
    assert out.returncode == 0 or variant == "c4"
    assert report["mean"] > 0
