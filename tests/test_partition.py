"""Multi-GPU host logic on CPU (world_size 2 and 3, gloo): band ownership and the row all-gather that completes the
framebuffer / ReSTIR reservoir planes across ranks (hiprt-path-tracer_b200/b2pt/partition.py; the same code runs under
nccl on the GPU box)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT

sys.path.insert(0, str(ROOT / "hiprt-path-tracer_b200"))


def _load_partition():
    import importlib.util
    spec = importlib.util.spec_from_file_location("b2pt_partition", ROOT / "hiprt-path-tracer_b200" / "b2pt" / "partition.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_band_ownership_covers_every_row_once():
    p = _load_partition()
    for height in (1, 63, 64, 65, 1080, 2160):
        for world in (1, 2, 3, 4, 8):
            rows = [p.owned_rows(height, r, world) for r in range(world)]
            allrows = np.sort(np.concatenate(rows))
            assert np.array_equal(allrows, np.arange(height))
            for r in range(world):
                assert np.all((rows[r] // p.BAND_ROWS) % world == r)
    # 1080p on 8 GPUs: 67 full bands + an 8-row tail band -> ranks own 8 or 9 bands, within 7 % of each other
    counts = [len(p.owned_rows(1080, r, 8)) for r in range(8)]
    assert sum(counts) == 1080 and max(counts) <= 144 and min(counts) >= 128


def _worker(rank, world, port, height, width, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    p = _load_partition()
    g = torch.Generator().manual_seed(1234)
    full_fb = torch.rand((height, width, 3), generator=g)
    full_planes = torch.randint(-2**31, 2**31 - 1, (9, height, width), generator=g, dtype=torch.int64).to(torch.int32)
    rows = torch.as_tensor(p.owned_rows(height, rank, world), dtype=torch.long)
    # every rank starts with ONLY its own rows valid (the rest poisoned), like a partitioned render
    fb = torch.full_like(full_fb, float("nan"))
    fb[rows] = full_fb[rows]
    planes = torch.full_like(full_planes, -1)
    planes[:, rows] = full_planes[:, rows]
    p.allgather_rows(fb, 0, rank, world, dist, torch)
    p.allgather_rows(planes, 1, rank, world, dist, torch)
    ok = torch.equal(fb, full_fb) and torch.equal(planes, full_planes)
    # the prepared form used by bench.py / the reservoir exchange: same result, and reusable (second call = no-op change)
    fb2 = torch.full_like(full_fb, float("nan"))
    fb2[rows] = full_fb[rows]
    gatherer = p.RowGatherer(fb2, 0, rank, world, dist, torch)
    gatherer()
    gatherer()
    ok = ok and torch.equal(fb2, full_fb)
    with open(os.path.join(out_dir, f"rank{rank}.txt"), "w") as f:
        f.write("ok" if ok else "mismatch")
    dist.destroy_process_group()


@pytest.mark.parametrize("world,height", [(2, 200), (3, 130), (2, 40), (3, 20)])
def test_allgather_rows_gloo(world, height, tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(world, port, height, 48, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert (tmp_path / f"rank{r}.txt").read_text() == "ok"


def test_c_partition_rows_match_the_python_mirror(b2pt):
    """b2pt_partition_rows (csrc/b2_comm.cu, pure host logic) against b2pt/partition.py, halo 0 (owned rows) and ReSTIR DI halos."""
    p = _load_partition()
    for height in (1, 31, 64, 97, 360, 1080, 2160):
        for world in (1, 2, 3, 4, 8):
            for band_rows, halo in ((16, 0), (32, 0), (32, 22), (32, 32), (16, 7), (32, height)):
                for rank in range(world):
                    got = b2pt.partition_rows(height, rank, world, band_rows, halo)
                    want = p.needed_rows(height, rank, world, band_rows, halo)
                    assert np.array_equal(got, want), (height, world, band_rows, halo, rank)
    # 4K on 8 GPUs with the ReSTIR DI halo: CameraRays covers the rank's bands + 22 rows either side = at most 76 / 32 of its rows
    own = len(b2pt.partition_rows(2160, 3, 8, 32, 0))
    need = len(b2pt.partition_rows(2160, 3, 8, 32, 22))
    assert own in (256, 272, 288) and own < need <= own * 76 // 32


def test_halo_plan_is_symmetric():
    """What rank a sends to rank b is exactly what b expects from a, and together with its own rows every rank ends up holding
    every row within the halo of its bands."""
    p = _load_partition()
    for height, world, band_rows, halo in ((360, 2, 32, 22), (360, 3, 32, 22), (1080, 4, 32, 22), (2160, 8, 32, 22), (130, 3, 16, 130)):
        plans = [p.halo_plan(height, r, world, band_rows, halo) for r in range(world)]
        for a in range(world):
            have = set(p.owned_rows(height, a, world, band_rows).tolist())
            for b, (send, recv) in plans[a].items():
                assert np.array_equal(send, plans[b][a][1]) and np.array_equal(recv, plans[a][b][1])
                assert all((y // band_rows) % world == a for y in send) and all((y // band_rows) % world == b for y in recv)
                have |= set(recv.tolist())
            assert have == set(p.needed_rows(height, a, world, band_rows, halo).tolist())
        if world >= 3 and halo < band_rows:
            assert all(set(plans[r]) <= {(r - 1) % world, (r + 1) % world} for r in range(world)), "a halo within one band only involves the two neighbours"


def _halo_worker(rank, world, port, height, width, band_rows, halo, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    p = _load_partition()
    g = torch.Generator().manual_seed(4321)
    full = torch.randint(-2**31, 2**31 - 1, (9, height, width), generator=g, dtype=torch.int64).to(torch.int32)
    own = torch.as_tensor(p.owned_rows(height, rank, world, band_rows), dtype=torch.long)
    need = torch.as_tensor(p.needed_rows(height, rank, world, band_rows, halo), dtype=torch.long)
    planes = torch.full_like(full, -1)
    planes[:, own] = full[:, own]
    p.exchange_halo(planes, 1, rank, world, band_rows, halo, dist, torch)
    ok = torch.equal(planes[:, need], full[:, need])
    outside = torch.ones(height, dtype=torch.bool)
    outside[need] = False
    ok = ok and bool((planes[:, outside] == -1).all())   # nothing but the halo travelled
    with open(os.path.join(out_dir, f"rank{rank}.txt"), "w") as f:
        f.write("ok" if ok else "mismatch")
    dist.destroy_process_group()


@pytest.mark.parametrize("world,height,band_rows,halo", [(2, 200, 32, 22), (3, 330, 32, 22), (3, 130, 16, 7), (2, 64, 32, 64)])
def test_halo_exchange_gloo(world, height, band_rows, halo, tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_halo_worker, args=(world, port, height, 24, band_rows, halo, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert (tmp_path / f"rank{r}.txt").read_text() == "ok"
