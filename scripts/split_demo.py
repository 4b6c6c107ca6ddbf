#!/usr/bin/env python
"""The single-stream feature split over GPUs, product path: one frame pair, its 300 features split over the
ranks, each rank ONE kernel for the whole coarse-to-fine loop, the per-iteration sums crossing between the GPUs through peer
memory over NVLink (svo_b200_sia_split_*).  Beside it, the round-1 formulation of the same thing -- a host loop with one NCCL
all-reduce of 44 doubles per Gauss-Newton iteration (rpg_svo_b200/split_align.py) -- and the undivided one-GPU kernel.  Launch:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 scripts/split_demo.py
"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rpg_svo_b200 import capi, shard, split_align, synth  # noqa: E402


def main():
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    d = synth.make_frame_pair(1000)
    ctx = capi.Context(local)
    ref, cur = ctx.frame(d["ref_pyr"]), ctx.frame(d["cur_pyr"])
    T0 = synth.se3_identity()
    full = lambda c: c.sparse_img_align(ref, cur, d["cam"], T0, d["px"], d["f"], d["pos"], d["has_point"], d["ref_pos"], 4, 0)  # noqa: E731
    g = full(ctx)
    t1 = time.perf_counter()
    for _ in range(50):
        full(ctx)
    t_single = (time.perf_counter() - t1) / 50

    # ---- round-1 formulation: host loop, one NCCL all-reduce per iteration
    ev = split_align.make_gpu_evaluator(ctx, ref, cur, d["cam"], d["px"], d["f"], d["pos"], d["has_point"], d["ref_pos"])
    kw = dict(dist=dist if world > 1 else None, device=dev, rank=rank, world=world)
    r_nccl = split_align.sparse_img_align_split(ev, T0, 300, 4, 0, **kw)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        r_nccl = split_align.sparse_img_align_split(ev, T0, 300, 4, 0, **kw)
    torch.cuda.synchronize()
    t_nccl = (time.perf_counter() - t0) / 5

    # ---- product path: the kernels exchange through peer memory
    t_p2p, r_p2p = None, None
    if world > 1:
        handle, _ = ctx.sia_split_create(rank, world, 1)
        handles = [None] * world
        dist.all_gather_object(handles, handle)
        ctx.sia_split_connect(ipc_handles=handles)
        dist.barrier()
        lo, hi = shard.shard_range(300, rank, world)
        part = lambda: ctx.sparse_img_align(ref, cur, d["cam"], T0, d["px"][lo:hi], d["f"][lo:hi], d["pos"][lo:hi],  # noqa: E731
                                            d["has_point"][lo:hi], d["ref_pos"], 4, 0)
        r_p2p = part()
        dist.barrier()
        t0 = time.perf_counter()
        for _ in range(50):
            r_p2p = part()
        t_p2p = (time.perf_counter() - t0) / 50
        Ts = [None] * world
        dist.all_gather_object(Ts, r_p2p["T"].tobytes())
        same = all(t == Ts[0] for t in Ts)
        dist.barrier()
        ctx.sia_split_destroy()
    if rank == 0:
        e_n = synth.pose_error(r_nccl["T"], g["T"])
        msg = (f"split-feature mode, world={world}: undivided one-GPU call {t_single*1e6:.0f} us/pair (incl. H2D/D2H); "
               f"host loop + NCCL all-reduce per iteration: {r_nccl['n_allreduce']} all-reduces, {t_nccl*1e3:.2f} ms/pair, "
               f"|dt|={e_n[0]:.1e} m vs undivided")
        if r_p2p is not None:
            e_p = synth.pose_error(r_p2p["T"], g["T"])
            msg += (f"; in-kernel exchange through peer memory: {t_p2p*1e6:.0f} us/pair (incl. H2D/D2H), |dt|={e_p[0]:.1e} m "
                    f"dR={e_p[1]:.1e} rad vs undivided, tracked {r_p2p['n_tracked']} vs {g['n_tracked']}, all ranks bit-identical: {same}")
        print(msg, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
