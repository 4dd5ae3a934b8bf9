"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol include/b2pt.h declares,
the by-value render data has the reference's size, and the product fails loudly without a device (no CPU fallback)."""
import ctypes
import re

import pytest

from conftest import ROOT


def test_library_exports_every_declared_symbol(b2pt):
    lib = b2pt.load_library()
    names = b2pt.declared_symbols()
    assert len(names) >= 45
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_render_data_layout_constants(b2pt):
    text = (ROOT / "include" / "b2pt_render_data.h").read_text()
    assert "sizeof(b2pt_render_data) == 2784" in text
    # offsets used by the Python binding are the ones oracle/layout_check.cpp static_asserts against the reference
    check = (ROOT / "oracle" / "layout_check.cpp").read_text()
    for field in ("render_settings", "world_settings", "current_camera", "bsdfs_data"):
        assert field in check
    m = re.search(r"offsetof\(b2pt_render_data,\s*render_settings\)\s*==\s*(\d+)", check)
    if m:
        assert int(m.group(1)) == 728


def test_version_and_error_string(b2pt):
    lib = b2pt.load_library()
    assert lib.b2pt_version_string().decode().startswith("b2pt")
    assert isinstance(lib.b2pt_last_error_string(), bytes)


def test_no_cpu_fallback_without_device(b2pt):
    """Without a CUDA device context creation must FAIL (error code + message), never silently run on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    lib = b2pt.load_library()
    handle = ctypes.c_void_p()
    rc = lib.b2pt_ctx_create(0, ctypes.byref(handle))
    assert rc != 0
    assert lib.b2pt_last_error_string()
    with pytest.raises(b2pt.B2ptError):
        b2pt.Context(0)


def test_pass_names_follow_the_pass_enumeration(b2pt):
    """b2pt_renderer_pass_ms takes a B2_PASS_* index; the binding's name table must list them in enumeration order
    (a pass added to the enum but not to the table shifted every later name -- and zeroed bench.py's roofline)."""
    text = (ROOT / "hiprt-path-tracer_b200" / "csrc" / "b2_frame_loop.h").read_text()
    body = re.search(r"enum\s*\w*\s*\{([^}]*B2_PASS_COUNT[^}]*)\}", text).group(1)
    enumerators = [e.split("=")[0].strip() for e in re.sub(r"/\*.*?\*/", "", body, flags=re.S).split(",") if e.strip()]
    enumerators = [e for e in enumerators if e != "B2_PASS_COUNT"]
    assert len(enumerators) == len(b2pt.PASS_NAMES)
    assert enumerators[0] == "B2_PASS_CAMERA_RAYS" and b2pt.PASS_NAMES[0] == "CameraRays"
    assert enumerators[-1] == "B2_PASS_MEGAKERNEL" and b2pt.PASS_NAMES[-1] == "MegaKernel"
    key = lambda s: re.sub(r"[^a-z]", "", s.lower())
    for enumerator, name in zip(enumerators, b2pt.PASS_NAMES):
        stem = key(enumerator.replace("B2_PASS_", "")).replace("restirdi", "").replace("restir", "")
        assert stem[:6] in key(name), (enumerator, name)
