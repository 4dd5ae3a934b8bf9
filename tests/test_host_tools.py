"""CPU-side checks of host code against the reference's own functions (oracle/_ref/ref_tools: the reference's real
src/Utils/Utils.cpp and src/Image/Image.cpp compiled from where they lie)."""
import json
import subprocess

import numpy as np
import pytest

from conftest import DATA, REF_DIR


def tool(*args):
    exe = REF_DIR / "ref_tools"
    if not exe.exists():
        pytest.skip("oracle/_ref/ref_tools not built")
    out = subprocess.run([str(exe), *map(str, args)], capture_output=True, text=True)
    lines = [l for l in out.stdout.splitlines() if l.startswith('{"tool"')]
    assert lines, out.stdout[-500:] + out.stderr[-500:]
    return out.returncode, json.loads(lines[-1])


def test_alias_table_matches_the_reference_function(built):
    """include/b2pt_alias_table.h (used by the product AND by ref_render_*) against the reference's real Utils::compute_alias_table
    (src/Utils/Utils.cpp:63-131) on 3000 random weight lists of six families: every probability and alias bit-identical."""
    rc, report = tool("alias-check", 3000, 7)
    assert rc == 0 and report["mismatches"] == 0 and report["entries"] > 1_000_000, report


@pytest.mark.parametrize("flat", [False, True])
def test_hdr_decoder_matches_the_reference_loader(b2pt, tmp_path, flat):
    """b2pt_image_read_hdr (csrc/b2_host_scene.cpp) against Image32Bit::read_image_hdr(path, 4, flipY = true) -- stb_image -- on a
    run-length coded and on a flat Radiance file."""
    import make_hdr
    path = DATA / "sky_64x32.hdr"
    if flat:
        path = tmp_path / "flat.hdr"
        make_hdr.write_hdr(path, make_hdr.to_rgbe(make_hdr.read_b2env(DATA / "sky_128x64.b2env")[::2, ::2, :3]), flat=True)
    rc, report = tool("read-hdr", path, tmp_path / "ref.bin")
    mine = b2pt.read_hdr(path)
    ref = np.fromfile(tmp_path / "ref.bin", np.float32).reshape(report["height"], report["width"], 4)
    assert mine.shape == ref.shape and np.array_equal(mine.view(np.uint32), ref.view(np.uint32))
    assert mine[..., :3].max() > 100.0 and np.all(mine[..., 3] == 1.0)
