"""GPU edge cases of the hot path, through the C ABI: ragged and tiny frames (tiles are 8x4 pixels, bands 16 rows), a
scene without a single emissive triangle, zero samples, and the error returns. The checker is the reference oracle
(oracle/_ref, prebuilt) run on the box's host cores on the same inputs."""
import subprocess
import sys

import numpy as np
import pytest

from conftest import DATA, ROOT, pixels_not_bit_identical, rel_error, run_oracle

pytestmark = pytest.mark.gpu

RAGGED = [(1, 1), (7, 3), (97, 53), (130, 66)]


def render(b2pt, ctx, variant, scene, width, height, spp, bounces, envmap=None):
    r = b2pt.Renderer(ctx, width, height, variant, precise=True)
    r.set_scene_file(scene)
    r.set_luts_file(DATA / "brdf_luts.b2luts")
    if envmap:
        r.set_envmap_file(envmap)
    r.set(nb_bounces=bounces)
    if variant.startswith("c4"):
        r.set(restir_di_do_fused_spatiotemporal=1)  # what the oracle harness runs unless given --no-fused
    r.render(1)
    gb = r.gbuffer()
    if spp > 1:
        r.render(spp - 1)
    fb = r.framebuffer()
    r.close()
    return gb, fb


@pytest.mark.parametrize("variant", ["c2", "c4"])
@pytest.mark.parametrize("width,height", RAGGED)
def test_ragged_and_tiny_frames(b2pt, gpu_ctx, tmp_path, variant, width, height):
    """Frames that are not a multiple of the 8x4 tile, smaller than one tile, and smaller than a warp."""
    ref = run_oracle("zero" if variant == "c4" else "ref", variant, "cornell_pbr_16x9", width, height, 2, 4, tmp_path)
    gb, fb = render(b2pt, gpu_ctx, variant, DATA / "cornell_pbr_16x9.b2scene", width, height, 2, 4)
    assert fb.reshape(-1, 3).shape[0] == width * height
    assert np.array_equal(gb["prim"], ref["prim"])
    assert np.array_equal(gb["position"].ravel().view(np.uint32), ref["position"].ravel().view(np.uint32))
    assert (rel_error(fb, ref["framebuffer"]) > 1e-3).sum() == 0
    assert pixels_not_bit_identical(fb, ref["framebuffer"]) <= max(1, 1e-3 * width * height)


@pytest.mark.parametrize("variant,envmap", [("c2", None), ("c3", "data/sky_128x64.b2env"), ("c4", "data/sky_128x64.b2env")])
def test_scene_without_emissive_triangles(b2pt, gpu_ctx, tmp_path, variant, envmap):
    """No lights at all: the light samplers have nothing to draw from (power table of size 0, envmap candidate probability
    forced to 1 in ReSTIR DI, InitialCandidates.h:457-463); a closed box, so the frame is (almost) black in both."""
    scene = tmp_path / "no_emissive.b2scene"
    subprocess.run([sys.executable, str(ROOT / "tools" / "make_synthetic_scene.py"), str(scene), "--triangles", "6000", "--emissive", "0", "--seed", "7"],
                   check=True, capture_output=True)
    extra = ["--envmap", envmap] if envmap else []
    exe = ROOT / "oracle" / "_ref" / f"ref_render_{variant}"
    if not exe.exists():
        pytest.skip(f"{exe} not built")
    prefix = str(tmp_path / "ref_")
    subprocess.run([str(exe), "--scene", str(scene), "--luts", str(DATA / "brdf_luts.b2luts"), "--width", "64", "--height", "36", "--spp", "2", "--bounces", "4",
                    "--dump", prefix, *extra], check=True, capture_output=True, cwd=ROOT)
    ref_fb = np.fromfile(prefix + "framebuffer.bin", np.float32).reshape(-1, 3)
    ref_prim = np.fromfile(prefix + "gbuffer_prim.bin", np.int32)
    gb, fb = render(b2pt, gpu_ctx, variant, scene, 64, 36, 2, 4, envmap=(ROOT / envmap) if envmap else None)
    assert np.array_equal(gb["prim"], ref_prim)
    assert np.isfinite(fb).all()
    assert (rel_error(fb, ref_fb) > 1e-3).sum() == 0


def test_zero_samples_and_errors(b2pt, gpu_ctx):
    """render(0) is a no-op that leaves a zero framebuffer; a render without a scene, an unknown variant and a missing
    scene file are reported as errors (never a crash, never a silent CPU path)."""
    r = b2pt.Renderer(gpu_ctx, 32, 16, "c2", precise=True)
    with pytest.raises(b2pt.B2ptError):
        r.render(1)  # no scene yet
    with pytest.raises(b2pt.B2ptError):
        r.set_scene_file(DATA / "does_not_exist.b2scene")
    r.set_scene_file(DATA / "cornell_pbr_16x9.b2scene")
    r.set_luts_file(DATA / "brdf_luts.b2luts")
    r.render(0)
    assert not np.asarray(r.framebuffer()).any()
    r.render(1)
    assert np.asarray(r.framebuffer()).any()
    r.close()
    with pytest.raises(b2pt.B2ptError):
        b2pt.Renderer(gpu_ctx, 32, 16, "no_such_variant")
