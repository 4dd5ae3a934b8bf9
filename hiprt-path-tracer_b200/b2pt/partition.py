"""Multi-GPU partition of the frame: interleaved bands of BAND_ROWS pixel rows, band b -> rank b % world.

Pure host logic over torch tensors (CPU or CUDA), so the same code runs under `gloo` on CPU in tests/ and under `nccl`
on the B200 box. The reference has no multi-GPU path; the per-pixel passes are independent (CameraRays.h:95,
Megakernel.h:35) and only the ReSTIR reuse passes read neighbouring pixels' reservoirs
(kernels/ReSTIR/DI/SpatialReuse.h:142, FusedSpatiotemporalReuse.h:90,380,397), which is what `allgather_rows` completes.
"""
import numpy as np

BAND_ROWS = 16  # == B2PT_BAND_ROWS in include/b2pt.h (fine bands: 1080 rows / 16 = 67.5 bands keep 8 ranks within 7 % of each other)


RESTIR_BAND_ROWS = 32  # ReSTIR DI partitions: a band is at least as tall as the reuse halo (reuse_radius 16 + search radius 4 + 2)


def owned_rows(height, rank, world, band_rows=BAND_ROWS):
    """Row indices (ascending) of the bands rank owns."""
    rows = np.arange(height)
    return rows[(rows // band_rows) % world == rank]


def needed_rows(height, rank, world, band_rows, halo):
    """Rows within `halo` rows of a row rank owns (what CameraRays covers and what the reuse passes may read under a ReSTIR DI
    partition). Mirrors b2pt_partition_rows of the C layer (csrc/b2_comm.cu); tests/test_partition.py keeps the two in step."""
    own = np.zeros(height, bool)
    own[owned_rows(height, rank, world, band_rows)] = True
    reach = np.zeros(height, bool)
    for d in range(-halo, halo + 1):
        lo, hi = max(0, d), min(height, height + d)
        if lo < hi:
            reach[lo - d:hi - d] |= own[lo:hi]
    return np.nonzero(reach)[0]


def halo_plan(height, rank, world, band_rows, halo):
    """{peer: (rows rank sends to peer, rows rank receives from peer)}, ascending -- the halo exchange of csrc/b2_comm.cu."""
    plan = {}
    mine = needed_rows(height, rank, world, band_rows, halo)
    for q in range(world):
        if q == rank:
            continue
        theirs = needed_rows(height, q, world, band_rows, halo)
        send = theirs[(theirs // band_rows) % world == rank]
        recv = mine[(mine // band_rows) % world == q]
        if len(send) or len(recv):
            plan[q] = (send, recv)
    return plan


def exchange_halo(t, row_dim, rank, world, band_rows, halo, dist, torch):
    """Completes the halo rows of `t` in place: point-to-point sends / receives between the ranks whose bands touch (the host-logic
    twin of b2pt_comm_exchange_halo, run under gloo in tests)."""
    ops, recvs = [], []
    for q, (send, recv) in sorted(halo_plan(t.shape[row_dim], rank, world, band_rows, halo).items()):
        if len(send):
            ops.append(dist.P2POp(dist.isend, t.index_select(row_dim, torch.as_tensor(send, dtype=torch.long)).contiguous(), q))
        if len(recv):
            shape = list(t.shape)
            shape[row_dim] = len(recv)
            buf = torch.empty(shape, dtype=t.dtype)
            ops.append(dist.P2POp(dist.irecv, buf, q))
            recvs.append((recv, buf))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    for recv, buf in recvs:
        t.index_copy_(row_dim, torch.as_tensor(recv, dtype=torch.long), buf)
    return t


def allgather_rows(t, row_dim, rank, world, dist, torch, group=None):
    """Completes `t` in place across ranks: every rank contributes the rows it owns along `row_dim` (size = frame
    height) and receives everybody else's. Rows are packed to a contiguous send buffer (padded to the largest owner),
    all-gathered, and scattered back -- one collective per call."""
    if world == 1:
        return t
    height = t.shape[row_dim]
    rows = [torch.as_tensor(owned_rows(height, k, world), device=t.device, dtype=torch.long) for k in range(world)]
    max_rows = max(len(r) for r in rows)
    if max_rows == 0:
        return t
    mine = t.index_select(row_dim, rows[rank])
    pad_shape = list(mine.shape)
    pad_shape[row_dim] = max_rows
    send = torch.zeros(pad_shape, dtype=t.dtype, device=t.device)
    send.narrow(row_dim, 0, len(rows[rank])).copy_(mine)
    recv = torch.empty([world] + pad_shape, dtype=t.dtype, device=t.device)
    if dist.get_backend(group) == "nccl":
        dist.all_gather_into_tensor(recv, send, group=group)
    else:
        parts = [recv[k] for k in range(world)]
        dist.all_gather(parts, send, group=group)
    for k in range(world):
        if k != rank and len(rows[k]):
            t.index_copy_(row_dim, rows[k], recv[k].narrow(row_dim, 0, len(rows[k])))
    return t


class RowGatherer:
    """allgather_rows with everything that does not change between calls prepared once: row index tensors, the padded
    send buffer and the receive buffer. One gather = index_select into the send buffer, ONE all-gather, world-1 index_copy_."""

    def __init__(self, t, row_dim, rank, world, dist, torch, group=None):
        self.t, self.row_dim, self.rank, self.world, self.dist, self.torch, self.group = t, row_dim, rank, world, dist, torch, group
        height = t.shape[row_dim]
        self.rows = [torch.as_tensor(owned_rows(height, k, world), device=t.device, dtype=torch.long) for k in range(world)]
        self.max_rows = max(len(r) for r in self.rows) if world > 1 else 0
        if world > 1 and self.max_rows:
            pad_shape = list(t.shape)
            pad_shape[row_dim] = self.max_rows
            self.send = torch.zeros(pad_shape, dtype=t.dtype, device=t.device)
            self.recv = torch.empty([world] + pad_shape, dtype=t.dtype, device=t.device)
            self.nccl = dist.get_backend(group) == "nccl"

    def __call__(self):
        if self.world == 1 or not self.max_rows:
            return self.t
        torch, rd = self.torch, self.row_dim
        mine = self.rows[self.rank]
        if len(mine):
            torch.index_select(self.t, rd, mine, out=self.send.narrow(rd, 0, len(mine)))
        if self.nccl:
            self.dist.all_gather_into_tensor(self.recv, self.send, group=self.group)
        else:
            self.dist.all_gather([self.recv[k] for k in range(self.world)], self.send, group=self.group)
        for k in range(self.world):
            if k != self.rank and len(self.rows[k]):
                self.t.index_copy_(rd, self.rows[k], self.recv[k].narrow(rd, 0, len(self.rows[k])))
        return self.t


class _DevicePointer:
    """Exposes a raw CUDA device pointer through __cuda_array_interface__ so torch can wrap it without a copy."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = dict(shape=tuple(shape), typestr=typestr, data=(int(ptr), False), version=3, strides=None)


def device_tensor(torch, ptr, shape, typestr):
    return torch.as_tensor(_DevicePointer(ptr, shape, typestr), device="cuda")
