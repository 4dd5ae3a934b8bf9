"""N > 1 on real GPUs (needs >= 2 devices; skipped otherwise): the partitioned render (interleaved row bands, NCCL reservoir
exchange + framebuffer all-gather) must be bit-identical to the single-GPU render of the same frame."""
import json
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _gpu_count():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.parametrize("variant,extra", [("c2", []), ("c4", ["--fused", "1"]), ("c4", ["--fused", "0"]),
                                           ("c4", ["--fused", "1", "--scene", "synthetic_20k", "--envmap", "data/sky_128x64.b2env"]),
                                           ("c4p", ["--fused", "1", "--scene", "synthetic_20k", "--envmap", "data/sky_128x64.b2env"])])
def test_partitioned_render_is_bit_identical(built, variant, extra):
    n = min(_gpu_count(), 4)
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1", "--master-port", "29517",
           str(ROOT / "tools" / "multigpu_check.py"), "--variant", variant, *extra]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=240, cwd=ROOT)
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert out.returncode == 0 and lines, out.stdout[-2000:] + out.stderr[-2000:]
    report = json.loads(lines[-1])
    assert report["bit_identical_on_all_ranks"], report
    if variant.startswith("c4"):
        assert report["reservoir_exchanges"] > 0
