"""Function-level known-answer vectors: the reference's own bsdf_dispatcher_eval / bsdf_dispatcher_sample
(oracle/_ref/ref_kat = ref_harness.cpp built with -DB2_REF_KAT) against the product's bsdf_eval / bsdf_sample (oracle/_port) on
every material of the synthetic scene x 64 seeded (normal, view, light, RNG seed) tuples. The bar is BIT-EXACT: colour, pdf and
sampled direction. Coated materials are left out of the comparison: with the bare RayVolumeState of this harness the
reference's coat lobe reads ior[-1] and returns NaN (in a render the interior stack is initialised by the primary hit)."""
import subprocess
import sys

import numpy as np
import pytest

from conftest import DATA, PORT_EXE, REF_DIR, ROOT

sys.path.insert(0, str(ROOT / "tools"))


def test_bsdf_eval_and_sample_bit_exact(built, tmp_path):
    import b2scene
    ref_exe = REF_DIR / "ref_kat"
    if not ref_exe.exists() or not PORT_EXE.exists():
        pytest.skip("oracle binaries not built")
    scene = DATA / "synthetic_20k.b2scene"
    common = ["--scene", str(scene), "--luts", str(DATA / "brdf_luts.b2luts"), "--kat"]
    subprocess.run([str(ref_exe), *common, "--dump", str(tmp_path / "ref_")], check=True, capture_output=True)
    subprocess.run([str(PORT_EXE), "--variant", "c2", *common, "--dump", str(tmp_path / "port_")], check=True, capture_output=True)
    materials = b2scene.read_scene(scene)["materials"]
    uncoated = materials[:, b2scene.MAT_INDEX["coat"]] == 0.0
    assert uncoated.sum() >= 40
    for name, width in (("eval", 4), ("sample", 7)):
        ref = np.fromfile(tmp_path / f"ref_kat_bsdf_{name}.bin", np.float32).reshape(len(materials), 64, width)[uncoated]
        port = np.fromfile(tmp_path / f"port_kat_bsdf_{name}.bin", np.float32).reshape(len(materials), 64, width)[uncoated]
        assert np.isfinite(ref).all()
        assert (np.abs(ref[..., :3]).sum(-1) > 0).mean() > 0.3, "the vectors must exercise non-trivial evaluations"
        assert np.array_equal(ref.view(np.uint32), port.view(np.uint32)), f"bsdf_{name}: {(ref.view(np.uint32) != port.view(np.uint32)).any(-1).sum()} records differ"


def test_light_samplers_and_envmap_bit_exact(built, tmp_path):
    """sample_one_emissive_triangle_uniform / _power (LightUtils.h:248-309), envmap_sample (Envmap.h:77-133) and envmap_eval
    (:139-168) of the reference called directly on 4096 seeds each, against the product's samplers: triangle index, point,
    normal, area, emission, pdf, direction -- every word identical."""
    ref_exe = REF_DIR / "ref_kat"
    if not ref_exe.exists() or not PORT_EXE.exists():
        pytest.skip("oracle binaries not built")
    common = ["--scene", str(DATA / "synthetic_20k.b2scene"), "--luts", str(DATA / "brdf_luts.b2luts"), "--envmap", "data/sky_128x64.b2env", "--kat"]
    subprocess.run([str(ref_exe), *common, "--dump", str(tmp_path / "ref_")], check=True, capture_output=True, cwd=ROOT)
    subprocess.run([str(PORT_EXE), "--variant", "c2", *common, "--dump", str(tmp_path / "port_")], check=True, capture_output=True, cwd=ROOT)
    for name, width in (("light_uniform", 12), ("light_power", 12), ("envmap_sample", 7), ("envmap_eval", 4)):
        ref = np.fromfile(tmp_path / f"ref_kat_{name}.bin", np.uint32).reshape(4096, width)
        port = np.fromfile(tmp_path / f"port_kat_{name}.bin", np.uint32).reshape(4096, width)
        assert len(np.unique(ref[:, 0])) > 100, "the vectors must spread over many lights / texels"
        assert np.isfinite(ref.view(np.float32)[:, 1:]).all()
        assert np.array_equal(ref, port), f"{name}: {(ref != port).any(-1).sum()} of 4096 records differ"
