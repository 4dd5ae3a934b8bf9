"""Shared fixtures. `gpu` marks tests that need a CUDA device (run on the B200 box: `pytest -m gpu`); everything else
runs on CPU (`pytest -m "not gpu"`): the oracles against the committed golden fixtures, host logic, and the C-ABI
export check. The oracle (oracle/_ref = the reference's own CPU kernels compiled in place, oracle/_port = host build
of the product's per-pixel headers) is TEST INFRASTRUCTURE: only tests/, smoke() and bench.py's cpu_baseline use it."""
import json
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))

import __graft_entry__ as entry  # noqa: E402

GOLDEN = ROOT / "tests" / "golden"
DATA = ROOT / "data"
REF_DIR = ROOT / "oracle" / "_ref"
PORT_EXE = ROOT / "oracle" / "_port" / "port_render"


# Fixtures on which the REFERENCE reads uninitialised memory often enough to matter: principled_bsdf_sample does not write
# incident_light_info when it picks the sheen lobe (Principled.h:1875-1884), so a ReSTIR DI BSDF candidate's flags are stack garbage
# (DESIGN.md section 3). The textured Cornell box has a sheen TEXTURE on a large wall: 17 of 5184 pixels follow the garbage.
UNINITIALISED_READ_PIXELS = {"c4_texenv_96x54_fused": 20}


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def built():
    """Library + oracles built (no GPU needed: nvcc cross-compiles)."""
    lib = ROOT / "hiprt-path-tracer_b200" / "libb2pt.so"
    if not lib.exists() or not PORT_EXE.exists() or not (REF_DIR / "ref_render_c1").exists():
        entry.build()
    entry.ensure_data()
    return True


@pytest.fixture(scope="session")
def b2pt(built):
    return entry.load_b2pt()


@pytest.fixture(scope="session")
def gpu_ctx(b2pt):
    ctx = b2pt.Context(0)  # raises if there is no device: there is no CPU fallback
    yield ctx
    ctx.close()


def load_golden(name):
    """name: a fixture of tests/golden/, or "options/<fixture>" for one of tests/golden/options/."""
    g = np.load(GOLDEN / f"{name}.npz")
    meta = json.loads(bytes(g["meta"]).decode())
    return g, meta


def run_oracle(kind, variant, scene, width, height, spp, bounces, tmp_path, extra=(), threads=None):
    """kind: 'ref' (reference sources compiled in place), 'extras' (the same harness + --cameras / --repeat, a separate binary)
    or 'port' (host build of the product headers). scene: a name under data/ or the path of a .b2scene blob."""
    prefix = str(tmp_path / f"{kind}_")
    if kind in ("ref", "extras", "zero"):  # zero: the extras harness with every automatic variable zero-initialised (oracle/Makefile)
        exe = REF_DIR / f"ref_{ {'ref': 'render', 'extras': 'extras', 'zero': 'zero'}[kind] }_{variant}"
        cmd = [str(exe)]
    else:
        exe = PORT_EXE
        cmd = [str(exe), "--variant", variant]
    if not exe.exists():
        pytest.skip(f"{exe} not built")
    scene_path = Path(scene) if str(scene).endswith(".b2scene") else DATA / f"{scene}.b2scene"
    cmd += ["--scene", str(scene_path), "--luts", str(DATA / "brdf_luts.b2luts"), "--width", str(width), "--height", str(height),
            "--spp", str(spp), "--bounces", str(bounces), "--dump", prefix, *extra]
    if threads:
        cmd += ["--threads", str(threads)]
    out = subprocess.run(cmd, check=True, capture_output=True, text=True, cwd=ROOT)  # envmap paths in `extra` are repo-relative
    meta = json.loads([l for l in out.stdout.splitlines() if l.startswith('{"oracle"')][-1])
    n = width * height
    load = lambda nm, dt: np.fromfile(prefix + nm + ".bin", dt)
    mat = load("gbuffer_material", np.uint32).reshape(n, 27).copy()
    mat[:, 0] &= 0xFF
    return dict(meta=meta, prim=load("gbuffer_prim", np.int32), position=load("gbuffer_position", np.float32).reshape(n, 3),
                geometric_normal=load("gbuffer_geometric_normal", np.uint32) & 0xFFFFFF, shading_normal=load("gbuffer_shading_normal", np.uint32) & 0xFFFFFF,
                material=mat, framebuffer=load("framebuffer", np.float32).reshape(n, 3))


def envmap_of(meta):
    """Repo-relative envmap blob named in a fixture's extra args, or None."""
    extra = list(meta.get("extra", []))
    return ROOT / extra[extra.index("--envmap") + 1] if "--envmap" in extra else None


def pixels_not_bit_identical(a, b):
    """Number of pixels whose float32 RGB differs in any bit."""
    a = np.ascontiguousarray(np.asarray(a, np.float32).reshape(-1, 3)).view(np.uint32)
    b = np.ascontiguousarray(np.asarray(b, np.float32).reshape(-1, 3)).view(np.uint32)
    return int((a != b).any(axis=1).sum())


def rel_error(a, b):
    """Per-pixel max-over-channels relative error, |a-b| / max(|a|,|b|), 0 where both are ~0."""
    a = np.asarray(a, np.float64).reshape(-1, 3)
    b = np.asarray(b, np.float64).reshape(-1, 3)
    denom = np.maximum(np.abs(a), np.abs(b))
    return np.where(denom > 1e-6, np.abs(a - b) / np.maximum(denom, 1e-30), 0.0).max(axis=1)


def window_mask(width, height, window, margin=0):
    """Boolean per-pixel mask of the window [x0, y0, x1, y1) shrunk by `margin` pixels."""
    x0, y0, x1, y1 = window
    m = np.zeros((height, width), bool)
    m[y0 + margin:y1 - margin, x0 + margin:x1 - margin] = True
    return m.ravel()


def moved_camera(camera, offset):
    """The scene camera (dict of tools/b2scene.read_scene) translated by `offset` in world space, as the 52 floats of a
    B2SceneCamera: inverse_view gets the translation, view_projection = VP * T(-offset), all in float32."""
    d = np.asarray(offset, np.float32)
    inv_view = np.array(camera["inverse_view"], np.float32).reshape(4, 4).copy()
    vp = np.array(camera["view_projection"], np.float32).reshape(4, 4).copy()
    inv_view[:3, 3] = inv_view[:3, 3] + d
    vp[:, 3] = vp[:, 3] - (vp[:, 0] * d[0] + vp[:, 1] * d[1] + vp[:, 2] * d[2])
    position = np.array(camera["position"], np.float32) + d
    return np.concatenate([inv_view.ravel(), np.array(camera["inverse_projection"], np.float32).ravel(), vp.ravel(), position,
                           np.array([camera["vertical_fov"]], np.float32)]).astype(np.float32)
