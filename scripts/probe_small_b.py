"""Small-batch probe (run on the GPU box): device / end-to-end latency of one frame pair (C1 as a live stream,
pair k+1 issued after pair k returned), and throughput at 32 pairs per GPU (C4), next to the full window."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from rpg_svo_b200 import capi

def main():
    Bs = [int(a) for a in sys.argv[1:]] or [1, 8, 32, 148, 296, 2368]
    Bmax = max(Bs)
    inp = bench.make_inputs(0, Bmax, "cuda:0")
    ctx = capi.Context(0)
    stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", 0))
    pool = capi.FramePool(ctx, bench.W, bench.H, bench.NLEVELS, Bmax + 1)
    host_l0 = inp["level0"].cpu().pin_memory()
    pool.upload(0, Bmax + 1, host_l0.data_ptr(), bench.W * bench.H)
    ctx.synchronize()
    fr = pool.frames
    out = {}
    for B in Bs:
        n = B * bench.NFEAT
        ctx.sia_batch_stage(fr[:B], fr[1:B + 1], inp["cam"], inp["T0"][:B], inp["off"][:B + 1], inp["px"][:n], inp["f"][:n],
                            inp["pos"][:n], inp["hp"][:n], inp["ref_pos"][:B], bench.MAX_LEVEL, bench.MIN_LEVEL, bench.NITER)
        for _ in range(5):
            ctx.sia_batch_run()
        ctx.synchronize()
        reps = 200 if B <= 296 else 20
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            ctx.sia_batch_run()
        e1.record(stream)
        ctx.synchronize()
        us = 1e3 * e0.elapsed_time(e1) / reps
        st = ctx.sia_batch_fetch()["stats"]
        out[f"B{B}"] = {"device_us_per_launch": us, "frames_per_s": B / (us * 1e-6), "mean_iters": float(st["n_iters"].mean())}
    # live stream: one ABI call per pair with host buffers (features H2D, kernel, results D2H), pair k+1 after pair k
    nlive = min(64, Bmax)
    def live():
        for k in range(nlive):
            s = slice(k * bench.NFEAT, (k + 1) * bench.NFEAT)
            ctx.sparse_img_align(fr[k], fr[k + 1], inp["cam"], inp["T0"][k], inp["px"][s], inp["f"][s], inp["pos"][s],
                                 inp["hp"][s], inp["ref_pos"][k], bench.MAX_LEVEL, bench.MIN_LEVEL, bench.NITER)
    live()
    t0 = time.perf_counter(); live(); t1 = time.perf_counter()
    out["live_stream_e2e_us_per_pair"] = 1e6 * (t1 - t0) / nlive
    print(json.dumps(out))

main()
