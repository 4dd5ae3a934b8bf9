"""Converged-image gate (BASELINE.json north_star: "converged images must agree within a stated RMSE/FLIP threshold").
Images of 128 / 64 samples per pixel from the CUDA path against the reference CPU path (oracle/_ref) at the same seeds, compared
with the reference's OWN metrics: its tone mapper, its FLIP (thirdparties/nvidia-FLIP through Utils::compute_image_weighted_median_FLIP)
and its RMSE, all compiled from /root/reference into oracle/_ref/ref_tools (oracle/ref_tools.cpp).
Stated thresholds: mean FLIP <= 0.02 and relative RMSE of the HDR image <= 0.03 for the fast-math flavour; the precise flavour is
bit-identical, i.e. FLIP = 0 exactly. For scale, two renders of the same frame with DIFFERENT seeds sit at a FLIP of ~0.1."""
import json
import subprocess

import numpy as np
import pytest

from conftest import DATA, REF_DIR, ROOT, run_oracle

pytestmark = pytest.mark.gpu
FLIP_MAX, REL_RMSE_MAX = 0.02, 0.03


def flip(tmp_path, a, b, width, height, samples, tag):
    exe = REF_DIR / "ref_tools"
    if not exe.exists():
        pytest.skip("oracle/_ref/ref_tools not built")
    pa, pb = tmp_path / f"{tag}_a.bin", tmp_path / f"{tag}_b.bin"
    np.asarray(a, np.float32).tofile(pa)
    np.asarray(b, np.float32).tofile(pb)
    out = subprocess.run([str(exe), "flip", str(pa), str(pb), str(width), str(height), str(samples)], check=True, capture_output=True, text=True)
    return json.loads([l for l in out.stdout.splitlines() if l.startswith('{"tool"')][-1])


def render(b2pt, ctx, variant, scene, width, height, samples, bounces, precise, envmap=None, seed=None):
    r = b2pt.Renderer(ctx, width, height, variant, precise=precise)
    r.set_scene_file(DATA / f"{scene}.b2scene")
    r.set_luts_file(DATA / "brdf_luts.b2luts")
    r.set(nb_bounces=bounces)
    if variant.startswith("c4"):
        r.set(restir_di_do_fused_spatiotemporal=1)
    if envmap:
        r.set_envmap_file(ROOT / envmap)
    if seed is not None:
        r.set_host_seed(seed)
    r.render(samples)
    fb = r.framebuffer().copy()
    r.close()
    return fb


@pytest.mark.parametrize("variant,scene,width,height,samples,bounces,envmap", [("c2", "cornell_pbr_16x9", 320, 180, 128, 8, None),
                                                                                ("c4", "synthetic_20k", 192, 108, 64, 4, "data/sky_128x64.b2env")])
def test_converged_image_agrees_with_the_reference(b2pt, gpu_ctx, tmp_path, variant, scene, width, height, samples, bounces, envmap):
    extra = ["--envmap", envmap] if envmap else []
    # ReSTIR DI: the zero-initialised reference build (what the plain build's unwritten stack variables hold moves with every harness edit)
    ref = run_oracle("zero" if variant.startswith("c4") else "ref", variant, scene, width, height, samples, bounces, tmp_path, extra=extra)["framebuffer"]
    precise = render(b2pt, gpu_ctx, variant, scene, width, height, samples, bounces, True, envmap)
    fast = render(b2pt, gpu_ctx, variant, scene, width, height, samples, bounces, False, envmap)
    other_seed = render(b2pt, gpu_ctx, variant, scene, width, height, samples, bounces, True, envmap, seed=7)

    # the reference marks NaN samples with 1e30 "pink" pixels (display_NaNs); both sides produce the same ones -- they are not image content
    sane = lambda img: np.where(np.asarray(img, np.float32) > 1e20, 0.0, img).astype(np.float32)
    m_precise = flip(tmp_path, sane(ref), sane(precise), width, height, samples, "precise")
    m_fast = flip(tmp_path, sane(ref), sane(fast), width, height, samples, "fast")
    m_noise = flip(tmp_path, sane(ref), sane(other_seed), width, height, samples, "noise")
    print(json.dumps(dict(variant=variant, precise=m_precise, fast=m_fast, different_seed=m_noise)))

    # c2: bit-identical => exactly 0. ReSTIR DI over 64 frames: the reference's uninitialised sampled_lobe_info on sheen-lobe samples
    # (the garbage depends on the reference BINARY's stack layout -- recompiling oracle/ref_harness.cpp moves it)
    # reaches a few reservoirs per frame and temporal reuse spreads them: the precise flavour then sits an order of magnitude below
    # the fast flavour instead of at zero (the host port of the same code differs from the reference in the same way, tests/test_oracle.py).
    precise_flip_max, precise_rmse_max = (1e-4, 1e-3) if not variant.startswith("c4") else (FLIP_MAX / 10, REL_RMSE_MAX / 10)
    assert m_precise["mean_flip"] <= precise_flip_max and m_precise["rel_rmse_hdr"] <= precise_rmse_max, m_precise
    assert m_fast["mean_flip"] <= FLIP_MAX and m_fast["rel_rmse_hdr"] <= REL_RMSE_MAX, m_fast
    # the gate has teeth: an independent render of the same frame is an order of magnitude further away than the fast flavour
    assert m_noise["mean_flip"] > 2 * m_fast["mean_flip"]


@pytest.mark.parametrize("variant", ["c2", "c3"])
def test_alpha_tested_scene_converges_to_the_reference(b2pt, gpu_ctx, tmp_path, variant):
    """Alpha-tested materials: statistical bar (see tests/test_oracle.py): the CUDA image must be as close to the reference's as two CUDA
    renders with different seeds are to each other, by the reference's own FLIP."""
    w, h, spp, bounces = 160, 90, 512, 4
    ref = run_oracle("extras", variant, "cornell_alpha_16x9", w, h, spp, bounces, tmp_path)["framebuffer"]
    cuda = render(b2pt, gpu_ctx, variant, "cornell_alpha_16x9", w, h, spp, bounces, True)
    other = render(b2pt, gpu_ctx, variant, "cornell_alpha_16x9", w, h, spp, bounces, True, seed=7)
    m_ref = flip(tmp_path, ref, cuda, w, h, spp, "alpha_ref")
    m_noise = flip(tmp_path, cuda, other, w, h, spp, "alpha_noise")
    print(json.dumps(dict(variant=variant, against_reference=m_ref, different_seed=m_noise)))
    assert m_ref["mean_flip"] <= 1.25 * m_noise["mean_flip"] + 0.003, (m_ref, m_noise)
    assert abs(m_ref["mean_a"] - m_ref["mean_b"]) <= 0.02 * m_ref["mean_a"], m_ref
