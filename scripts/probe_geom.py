"""Full-batch probe (run on the GPU box): device time of the alignment kernel per launch geometry at one batch size.
   python scripts/probe_geom.py B [ctas_per_pair:features_per_thread ...]     e.g.  3552 -1:0 1:2 1:1"""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from rpg_svo_b200 import capi


def main():
    B = int(sys.argv[1])
    geoms = [tuple(int(x) for x in a.split(":")) for a in sys.argv[2:]] or [(-1, 0)]
    inp = bench.make_inputs(0, B, "cuda:0")
    ctx = capi.Context(0)
    stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", 0))
    pool = capi.FramePool(ctx, bench.W, bench.H, bench.NLEVELS, B + 1)
    host_l0 = inp["level0"].cpu().pin_memory()
    pool.upload(0, B + 1, host_l0.data_ptr(), bench.W * bench.H)
    ctx.synchronize()
    fr = pool.frames
    n = B * bench.NFEAT
    out = {}
    ref_T = None
    reps = int(os.environ.get("PROBE_REPS", "10"))
    for g in geoms:
        ctx.sia_config(*g)
        ctx.sia_batch_stage(fr[:B], fr[1:B + 1], inp["cam"], inp["T0"][:B], inp["off"][:B + 1], inp["px"][:n], inp["f"][:n],
                            inp["pos"][:n], inp["hp"][:n], inp["ref_pos"][:B], bench.MAX_LEVEL, bench.MIN_LEVEL, bench.NITER)
        for _ in range(2):
            ctx.sia_batch_run()
        ctx.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            ctx.sia_batch_run()
        e1.record(stream)
        ctx.synchronize()
        us = 1e3 * e0.elapsed_time(e1) / reps
        r = ctx.sia_batch_fetch()
        if ref_T is None:
            ref_T = r["T"].copy()
        out[f"{g[0]}:{g[1]}"] = {"device_us_per_launch": us, "frames_per_s": B / (us * 1e-6),
                                 "max_abs_dT_vs_first": float(np.abs(r["T"] - ref_T).max())}
    print(json.dumps(out))


main()
