#!/usr/bin/env python3
"""Run under torchrun with N ranks (one per GPU): renders a frame partitioned across the ranks (interleaved row bands,
ReSTIR DI reservoir halo exchange, framebuffer all-gather -- NCCL from C) and checks on every rank that the gathered frame is BIT-IDENTICAL to
the same frame rendered by that rank alone -- results must not depend on the number of GPUs.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/multigpu_check.py --variant c4
"""
import argparse
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import __graft_entry__ as entry  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--variant", default="c4")
    ap.add_argument("--scene", default="cornell_pbr_16x9")
    ap.add_argument("--envmap", default=None)
    ap.add_argument("--width", type=int, default=640)
    ap.add_argument("--height", type=int, default=360)
    ap.add_argument("--spp", type=int, default=3)
    ap.add_argument("--bounces", type=int, default=4)
    ap.add_argument("--fused", type=int, default=1)
    ap.add_argument("--precise", type=int, default=1)
    a = ap.parse_args()

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    entry.ensure_data()
    b2pt = entry.load_b2pt()
    ctx = b2pt.Context(local)

    def make():
        r = b2pt.Renderer(ctx, a.width, a.height, a.variant, precise=bool(a.precise))
        r.set_scene_file(ROOT / "data" / f"{a.scene}.b2scene")
        r.set_luts_file(ROOT / "data" / "brdf_luts.b2luts")
        r.set(nb_bounces=a.bounces)
        if a.variant.startswith("c4"):
            r.set(restir_di_do_fused_spatiotemporal=a.fused)
        if a.envmap:
            r.set_envmap_file(ROOT / a.envmap)
        return r

    whole = make()
    whole.render(a.spp)
    ref = whole.framebuffer().copy()
    ref_res = whole.reservoirs(3).copy() if a.variant.startswith("c4") else None
    whole.close()

    part = make()
    part.set_partition(rank, world, torch, dist)
    gather = part.make_gather(torch, dist)
    part.render(a.spp, sync=False)
    gather()
    part.synchronize()
    gather_ms = part.last_render_ms()
    got = part.framebuffer()
    exchanges = part.exchange_count
    info = part.partition_info()
    part.close()

    same = bool(np.array_equal(ref.view(np.uint32), got.view(np.uint32)))
    mismatching = int((ref.view(np.uint32) != got.view(np.uint32)).any(axis=-1).sum())
    flags = torch.tensor([1 if same else 0], device="cuda")
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps(dict(multigpu_check=a.variant, world=world, width=a.width, height=a.height, spp=a.spp, fused=a.fused, bit_identical_on_all_ranks=bool(flags.item()),
                              rank0_mismatching_pixels=mismatching, reservoir_exchanges=exchanges, render_plus_gather_ms=gather_ms, partition=info)))
    ctx.close()
    dist.destroy_process_group()
    sys.exit(0 if flags.item() else 1)


if __name__ == "__main__":
    main()
