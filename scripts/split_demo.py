#!/usr/bin/env python
"""NCCL demonstration of the split-feature mode: one frame pair, 300 features split over the
ranks, one all-reduce of 44 doubles per Gauss-Newton iteration.  Launch:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 scripts/split_demo.py
Prints the pose difference against the one-CTA kernel and the time per pair (expected: much slower than
the single-GPU kernel -- the mode is latency-bound by construction)."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rpg_svo_b200 import capi, split_align, synth  # noqa: E402


def main():
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    d = synth.make_frame_pair(1000)
    ctx = capi.Context(local)
    ref, cur = ctx.frame(d["ref_pyr"]), ctx.frame(d["cur_pyr"])
    ev = split_align.make_gpu_evaluator(ctx, ref, cur, d["cam"], d["px"], d["f"], d["pos"], d["has_point"], d["ref_pos"])
    kw = dict(dist=dist if world > 1 else None, device=torch.device("cuda", local), rank=rank, world=world)
    r = split_align.sparse_img_align_split(ev, synth.se3_identity(), 300, 4, 0, **kw)  # warm-up
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 5
    for _ in range(reps):
        r = split_align.sparse_img_align_split(ev, synth.se3_identity(), 300, 4, 0, **kw)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    if rank == 0:
        g = ctx.sparse_img_align(ref, cur, d["cam"], synth.se3_identity(), d["px"], d["f"], d["pos"], d["has_point"], d["ref_pos"], 4, 0)
        t1 = time.perf_counter()
        for _ in range(20):
            ctx.sparse_img_align(ref, cur, d["cam"], synth.se3_identity(), d["px"], d["f"], d["pos"], d["has_point"], d["ref_pos"], 4, 0)
        t_single = (time.perf_counter() - t1) / 20
        dtr, drot = synth.pose_error(r["T"], g["T"])
        print(f"split-feature mode: world={world} all-reduces/pair={r['n_allreduce']} time/pair={dt*1e3:.2f} ms  "
              f"vs one-CTA kernel {t_single*1e3:.3f} ms (incl. H2D/D2H);  |dt|={dtr:.2e} m dR={drot:.2e} rad, "
              f"tracked {r['n_tracked']} vs {g['n_tracked']}", flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
