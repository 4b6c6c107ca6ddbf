#!/usr/bin/env python
"""bench.py -- SparseImgAlign frames/sec on the BASELINE.json workload (C1), one process per GPU.

Workload (config.workload): one synthetic 640x480 camera stream per GPU, 300 features per frame,
pyramid levels 4..0, <=30 Gauss-Newton iterations per level (BASELINE.json configs[1]).  A *step* is
one pass of svo::SparseImgAlign::run over a window of `pairs_per_gpu` consecutive frame pairs of the
stream (frame k is the reference of pair k and the current frame of pair k-1); every pair starts from
the identity relative pose, so the pairs of a window are independent and map to one CTA each.

Two timed legs per run (CUDA events on the library's stream, barrier + sync on both sides, max over
ranks):
  value : pyramids + feature records already resident in HBM; K launches of the alignment kernel.
  e2e   : through the C ABI with HOST (pinned) buffers: per step every frame's level-0 image is
          copied host->device and its pyramid built on the GPU, features are packed + copied, the
          kernel runs, poses / masks / counters are copied back.
`--impl reference` times oracle/_ref (the reference's own sparse_img_align.cpp compiled in place; its build system cannot run here: Eigen, OpenCV,
Sophus, vikit, Boost are absent) on all host cores, on a bounded sample of the same pairs.

L2 policy: inputs larger than L2 (2368 pairs, 2369 distinct pyramids ~ 970 MB of images per step vs
126 MB of L2); no explicit flush.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "SparseImgAlign frames/sec at 640x480/300 feats/5 lvls; pose RMSE vs ref"
UNIT = "frames/s"
W, H, NFEAT, NLEVELS, MAX_LEVEL, MIN_LEVEL, NITER = 640, 480, 300, 5, 4, 0, 30


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs-per-gpu", type=int, default=2368)  # 8 full waves of 2 CTAs x 148 SMs
    ap.add_argument("--cpu-sample", type=int, default=256, help="pairs timed by the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-chunks", type=int, default=4, help="chunks of the window in the e2e leg (copy/compute overlap)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the C2/C3 side measurements")
    return ap.parse_args()


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for nme, val in zip(names, r[3:7]):
                    if val.lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def algorithmic_bytes(stats, n_feat: int) -> float:
    """Compulsory bytes of one frame pair (DESIGN.md 'algorithmic bytes'; SURVEY.md 8d adapted to
    this kernel's records): 49 B ref footprint per visible patch per level, 65 B feature record once,
    25 B current-image footprint per in-image patch per residual pass, pose in/out + H + counters +
    visibility mask out."""
    return (49.0 * float(stats["sum_visible"]) + 65.0 * n_feat + 25.0 * float(stats["sum_in_image"]) +
            96 + 96 + 288 + 16 + n_feat)


def make_inputs(rank: int, B: int, device: str):
    from rpg_svo_b200 import synth

    st = synth.make_stream_fast(1000 + rank, B + 1, W, H, NFEAT, NLEVELS, device=device)
    feats = st["feats"]
    px = np.concatenate([feats[k]["px"] for k in range(B)])
    f = np.concatenate([feats[k]["f"] for k in range(B)])
    pos = np.concatenate([feats[k]["pos"] for k in range(B)])
    hp = np.concatenate([feats[k]["has_point"] for k in range(B)])
    off = np.arange(B + 1, dtype=np.int32) * NFEAT
    ref_pos = np.stack([synth.se3_inv(st["poses"][k])[:, 3] for k in range(B)])
    T0 = np.tile(synth.se3_identity()[None], (B, 1, 1))
    T_gt = np.stack([synth.se3_mul(st["poses"][k + 1], synth.se3_inv(st["poses"][k])) for k in range(B)])
    return dict(cam=st["cam"], level0=st["level0"], px=px, f=f, pos=pos, hp=hp, off=off, ref_pos=ref_pos,
                T0=T0, T_gt=T_gt, poses=np.stack(st["poses"]))


def cpu_runner(ob, synth, inp, n: int):
    """The CPU implementation of the step's first n frame pairs: returns (kind, note, run(n_threads) -> seconds).
    kind "reference" = oracle/_ref, svo::SparseImgAlign::run of the reference's own sparse_img_align.cpp/frame.cpp
    compiled in place (stand-in third-party headers); "port" = the oracle restatement when
    oracle/_ref was never built."""
    sl = slice(0, n * NFEAT)
    if ob.ref_lib() is not None:
        rs = ob.RefStream(inp["level0"][:n + 1].cpu().numpy(), inp["cam"], NLEVELS, inp["poses"][:n + 1], inp["off"][:n + 1],
                          inp["px"][sl], inp["f"][sl], inp["pos"][sl], inp["hp"][sl])
        note = ("svo::SparseImgAlign::run from the reference's own svo/src/sparse_img_align.cpp + frame.cpp, compiled in "
                "place with g++ -O3 -mfma -mavx2 against stand-in Eigen/Sophus/vikit/OpenCV headers (oracle/shim); "
                "a fresh SparseImgAlign per frame as in FrameHandlerMono::processFrame; pyramids prebuilt")
        return "reference", note, lambda n_threads: rs.run(n_threads, MAX_LEVEL, MIN_LEVEL, NITER)["seconds"]
    pyrs = [synth.build_pyramid(inp["level0"][i].cpu().numpy(), NLEVELS) for i in range(n + 1)]

    def run(n_threads):
        t0 = time.perf_counter()
        ob.sparse_img_align_batch(pyrs[:n], pyrs[1:n + 1], inp["cam"], inp["T0"][:n], inp["off"][:n + 1], inp["px"][sl],
                                  inp["f"][sl], inp["pos"][sl], inp["hp"][sl], inp["ref_pos"][:n], MAX_LEVEL, MIN_LEVEL,
                                  NITER, n_threads=n_threads)
        return time.perf_counter() - t0

    return "port", "CPU oracle port of svo::SparseImgAlign::run (oracle/_ref not built on this box)", run


def oracle_pair(ob, synth, inp, pyr_cache, k):
    def pyr(i):
        if i not in pyr_cache:
            pyr_cache[i] = synth.build_pyramid(inp["level0"][i].numpy(), NLEVELS)
        return pyr_cache[i]

    s = slice(k * NFEAT, (k + 1) * NFEAT)
    return pyr(k), pyr(k + 1), s


def run_reference(args, rank: int, world: int):
    """CPU arm (rank 0 only): the reference's own SparseImgAlign (oracle/_ref) on all host cores, else the oracle port."""
    if rank != 0:
        return
    from oracle import binding as ob
    from rpg_svo_b200 import synth

    ob.build()
    cores = os.cpu_count() or 1
    sample = args.pairs_per_gpu  # the whole window of the step (bounded: ~1 s of CPU work on 1 core per 500 pairs)
    inp = make_inputs(0, sample, "cuda" if _has_cuda() else "cpu")
    kind, note, run = cpu_runner(ob, synth, inp, sample)
    for _ in range(args.warmup):
        run(cores)
    dt = 0.0
    for _ in range(args.steps):
        dt += run(cores)
    fps = sample * args.steps / dt
    line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32/f64",
            "data": "synthetic",
            "config": {"workload": "C1 single stream 640x480 / 300 feats / levels 4..0 / 30 GN iters",
                       "pairs_per_step": sample, "note": note},
            "cpu_baseline": {"value": fps, "unit": UNIT, "cores": cores, "kind": kind,
                             "sample": f"{sample} frame pairs of the step x {args.steps} steps, {cores} threads"},
            "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def _has_cuda() -> bool:
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False



def measure_other_paths(ctx, rank: int) -> dict:
    """BASELINE configs C2 (DepthFilter, 2000 seeds, 752x480) and C3 (HD 1920x1080, 1000 features: align2D +
    pose_optimizer) through the C ABI with host buffers (each call = H2D + one kernel + D2H + sync), next to
    the single-thread CPU oracle on the same inputs.  Reported as extra keys; not the headline metric."""
    import time as _t

    from oracle import binding as ob  # checker / CPU baseline only
    from rpg_svo_b200 import synth

    out = {}

    def timeit(fn, reps):
        fn()
        t0 = _t.perf_counter()
        for _ in range(reps):
            fn()
        return (_t.perf_counter() - t0) / reps

    # ---- C2: DepthFilter::updateSeeds --------------------------------------------------------
    c = synth.make_depth_case(2031, 2000, baseline=0.3)
    ref, cur = ctx.frame(c["ref_pyr"]), ctx.frame(c["cur_pyr"])
    args = ([ref], [c["T_ref_w"]], cur, c["T_cur_w"], c["cam"], c["ref_index"], c["ftr_px"], c["ftr_f"], c["ftr_level"],
            c["ftr_type"], c["ftr_grad"], c["batch_id"], c["batch_counter"], c["seeds"])
    g = ctx.depth_filter_update(*args)
    t_gpu = timeit(lambda: ctx.depth_filter_update(*args), 20)
    oargs = ([c["ref_pyr"]], [c["T_ref_w"]], c["cur_pyr"], c["T_cur_w"], c["cam"], c["ref_index"], c["ftr_px"], c["ftr_f"],
             c["ftr_level"], c["ftr_type"], c["ftr_grad"], c["batch_id"], c["batch_counter"], c["seeds"])
    o = ob.depth_filter_update(*oargs)
    t_cpu = timeit(lambda: ob.depth_filter_update(*oargs), 5)
    evals = int(g["n_zmssd"].sum())
    alg = 2000 * (48 + 64 + 400) + 64 * evals + 81 * 10 * int((g["status"] >= 4).sum())
    out["depth_filter_C2"] = {"seeds": 2000, "gpu_seeds_per_s_e2e": 2000 / t_gpu, "cpu_port_seeds_per_s_1thread": 2000 / t_cpu,
                              "ms_per_call_e2e": 1e3 * t_gpu, "zmssd_evals": evals, "algorithmic_bytes": alg,
                              "status_bit_exact_vs_oracle": bool(np.array_equal(g["status"], o["status"])),
                              "updated": int((g["status"] >= 5).sum())}
    ref.destroy(); cur.destroy()

    # ---- C3: HD align2D (1000 features, 10 iterations) + pose_optimizer (1000 observations) -----
    a = synth.make_align_case(3001, 1000, 1920, 1080, n_levels=6)
    fr = ctx.frame(a["pyr"])
    conv, px = ctx.align2d_batch(fr, a["level"], a["pwb"], a["patch"], 10, a["px_start"])
    t_gpu = timeit(lambda: ctx.align2d_batch(fr, a["level"], a["pwb"], a["patch"], 10, a["px_start"]), 20)

    def cpu_align():
        for i in range(1000):
            ob.align2d(a["pyr"][a["level"][i]], a["pwb"][i], a["patch"][i], 10, a["px_start"][i])

    t_cpu = timeit(cpu_align, 2)
    out["align2d_C3"] = {"features": 1000, "gpu_features_per_s_e2e": 1000 / t_gpu, "cpu_port_features_per_s_1thread": 1000 / t_cpu,
                         "ms_per_call_e2e": 1e3 * t_gpu, "converged": int(conv.sum()),
                         "note": "CPU figure includes ~1 us/feature of ctypes call overhead"}
    fr.destroy()
    pcase = synth.make_pose_opt_case(1005, 1000, 1920, 1080)
    pargs = (2.0, 10, pcase["cam"].fx, pcase["T_init"], pcase["f"], pcase["pos"], pcase["level"], pcase["has_point"])
    gp = ctx.pose_optimize(*pargs)
    t_gpu = timeit(lambda: ctx.pose_optimize(*pargs), 20)
    op = ob.pose_optimize(*pargs)
    t_cpu = timeit(lambda: ob.pose_optimize(*pargs), 10)
    out["pose_optimizer_C3"] = {"observations": 1000, "gpu_frames_per_s_e2e": 1 / t_gpu, "cpu_port_frames_per_s_1thread": 1 / t_cpu,
                                "ms_per_call_e2e": 1e3 * t_gpu, "iterations": int(gp["n_iter_done"]),
                                "pose_diff_vs_oracle_m": float(synth.pose_error(gp["T"], op["T"])[0]),
                                "algorithmic_bytes": 52 * 1000 * (int(gp["n_iter_done"]) + 2)}

    # ---- row f2: Reprojector::reprojectMap on a 10-keyframe map (speculative device alignment + host policy replay) ----
    m = synth.make_map_case(4001, n_kfs=10, n_points=1200, n_candidates=150)
    kfs, curf = [ctx.frame(p) for p in m["kf_pyr"]], ctx.frame(m["cur_pyr"])
    rargs = (m["view"], kfs, curf, m["cur_T_f_w"], m["cam"], m["options"], m["cell_order"], m["pt_type"], m["pt_n_failed"],
             m["pt_n_succeeded"])
    gr = ctx.reproject_map(*rargs)
    t_gpu = timeit(lambda: ctx.reproject_map(*rargs), 20)
    orp = ob.reproject_map(m)
    t_cpu = timeit(lambda: ob.reproject_map(m), 10)
    t_ref = timeit(lambda: ob.ref_reproject_map(m), 3) if ob.ref_lib() is not None else None
    out["reprojector_f2"] = {"map_points": int(m["view"]["n_points"]), "keyframes": 10, "matches": int(gr["n_matches"]),
                             "trials_sequential": int(gr["n_trials"]), "aligned_speculatively": int(gr["n_speculative"]),
                             "ms_per_call_e2e": 1e3 * t_gpu, "cpu_port_ms_1thread": 1e3 * t_cpu,
                             "cpu_reference_ms_incl_graph_build": None if t_ref is None else 1e3 * t_ref,
                             "same_features_as_oracle": bool(np.array_equal(gr["new_point"], orp["new_point"]) and
                                                             np.array_equal(gr["pt_type"], orp["pt_type"])),
                             "note": "the CPU matches only the ~130 candidates the cell policy reaches; the GPU aligns every "
                                     "in-frame point in one launch and replays the policy on the host"}
    # ---- row f4: FastDetector::detect on a 752x480 keyframe, 3 levels (svo/test/test_feature_detection.cpp:47-55) ----
    det_args = (curf, 30, 3, 20.0)
    gd = ctx.fast_detect(*det_args)
    t_gpu = timeit(lambda: ctx.fast_detect(*det_args), 50)
    od = ob.fast_detect(m["cur_pyr"], 3, 30, 20.0)
    t_cpu = timeit(lambda: ob.fast_detect(m["cur_pyr"], 3, 30, 20.0), 5)
    out["fast_detector_f4"] = {"image": "752x480, levels 0..2, 30 px cells", "corners": int(gd["n"]), "ms_per_call_e2e": 1e3 * t_gpu,
                               "cpu_port_ms_1thread": 1e3 * t_cpu, "reference_published_ms": 7.17,
                               "reference_published_source": "svo/test/test_feature_detection.cpp:55 (i7-W520, SSE2 fast library)",
                               "same_corners_as_oracle": bool(np.array_equal(gd["x"], od["x"]) and np.array_equal(gd["y"], od["y"])
                                                              and np.array_equal(gd["level"], od["level"])),
                               "algorithmic_bytes": int(sum(im.size for im in m["cur_pyr"][:3]) + 8 * 26 * 16)}
    for fr_ in kfs:
        fr_.destroy()
    curf.destroy()
    return out


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    from rpg_svo_b200 import capi
    from rpg_svo_b200.shard import shard_range

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the CUDA path is the only path (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    B, K, Wm = args.pairs_per_gpu, args.steps, args.warmup

    inp = make_inputs(rank, B, f"cuda:{local_rank}")
    ctx = capi.Context(local_rank)
    stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local_rank))
    host_l0 = inp["level0"].pin_memory()  # [B+1, H, W] uint8, the HOST buffers of the e2e leg
    pool = capi.FramePool(ctx, W, H, NLEVELS, B + 1)  # one slab: one strided H2D copy per window
    frames = pool.frames
    frame_bytes = W * H

    def upload_all():
        pool.upload(0, B + 1, host_l0.data_ptr(), frame_bytes)

    def stage():
        ctx.sia_batch_stage(frames[:B], frames[1:], inp["cam"], inp["T0"], inp["off"], inp["px"], inp["f"],
                            inp["pos"], inp["hp"], inp["ref_pos"], MAX_LEVEL, MIN_LEVEL, NITER)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ctx.synchronize()

    def max_over_ranks(ms: float) -> float:
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---------------- leg 1: device-resident (value) ----------------
    # clocks / throttle reasons are sampled from here to the end of the e2e leg (both timed regions)
    sampler = ClockSampler(local_rank)
    sampler.start()
    upload_all()
    stage()
    for _ in range(max(Wm, 3)):
        ctx.sia_batch_run()
    barrier()
    l0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(K):
        ctx.sia_batch_run()
    e1.record(stream)
    barrier()
    launches = ctx.launch_count() - l0
    ms_dev = max_over_ranks(e0.elapsed_time(e1))
    res = ctx.sia_batch_fetch()
    stats = res["stats"]
    alg_bytes = sum(algorithmic_bytes(stats[b], NFEAT) for b in range(B))
    kernel_ms = ms_dev / K
    value = world * B * K / (ms_dev * 1e-3)

    # ---------------- pose RMSE vs the oracle on a sample (not timed) ----------------
    rmse = None
    if rank == 0:
        from oracle import binding as ob  # checker only
        from rpg_svo_b200 import synth

        ob.build()
        idx = list(range(0, B, max(1, B // 16)))[:16]
        et, er, gt_t = [], [], []
        for k in idx:
            s = slice(k * NFEAT, (k + 1) * NFEAT)
            pr = synth.build_pyramid(inp["level0"][k].numpy(), NLEVELS)
            pc = synth.build_pyramid(inp["level0"][k + 1].numpy(), NLEVELS)
            o = ob.sparse_img_align(pr, pc, inp["cam"], inp["T0"][k], inp["px"][s], inp["f"][s], inp["pos"][s],
                                    inp["hp"][s], inp["ref_pos"][k], MAX_LEVEL, MIN_LEVEL, NITER, want_trace=False)
            a, b_ = synth.pose_error(res["T"][k], o["T"])
            et.append(a); er.append(b_)
            gt_t.append(synth.pose_error(res["T"][k], inp["T_gt"][k])[0])
        rmse = {"trans_m": float(np.sqrt(np.mean(np.square(et)))), "rot_rad": float(np.sqrt(np.mean(np.square(er)))),
                "pairs": len(idx), "gpu_vs_ground_truth_trans_m": float(np.sqrt(np.mean(np.square(gt_t))))}

    # ---------------- leg 2: end to end through the C ABI with host buffers ----------------
    e2e = None
    if not args.no_e2e:
        # The window is cut into chunks that alternate between two contexts (= two CUDA streams): while
        # one stream runs the alignment kernel of chunk c, the other copies the images of chunk c+1 over
        # PCIe.  Each chunk owns a private frame pool (pairs k..k+n-1 need frames k..k+n), so nothing
        # is shared between the streams; the boundary frame of a chunk is simply uploaded twice.
        n_chunks = max(1, min(args.e2e_chunks, B))
        ctx2 = capi.Context(local_rank)
        ctxs = [ctx, ctx2]
        bounds = [shard_range(B, c, n_chunks) for c in range(n_chunks)]
        pools = [capi.FramePool(ctxs[c % 2], W, H, NLEVELS, e - b + 1) for c, (b, e) in enumerate(bounds)]
        base = host_l0.data_ptr()

        def e2e_step():
            for c, (b, e) in enumerate(bounds):
                cx, pl, n = ctxs[c % 2], pools[c], e - b
                fo = inp["off"][b:e + 1] - inp["off"][b]  # offsets index the (sliced) arrays passed below
                sl = slice(b * NFEAT, e * NFEAT)
                if c >= 2:  # the context's staging buffers are about to be reused: drain its previous chunk
                    results[c - 2] = cx.sia_batch_fetch()
                # features first (host packing + small H2D), then the images: the packing of the NEXT chunk
                # (other context) overlaps this chunk's image copy; the kernel is ordered after both
                cx.sia_batch_stage(pl.frames[:n], pl.frames[1:n + 1], inp["cam"], inp["T0"][b:e], fo, inp["px"][sl],
                                   inp["f"][sl], inp["pos"][sl], inp["hp"][sl], inp["ref_pos"][b:e], MAX_LEVEL,
                                   MIN_LEVEL, NITER)
                pl.upload(0, n + 1, base + b * frame_bytes, frame_bytes)
                cx.sia_batch_run()
            for c in range(max(0, n_chunks - 2), n_chunks):
                results[c] = ctxs[c % 2].sia_batch_fetch()
            return results

        results = [None] * n_chunks
        for _ in range(max(Wm, 3)):
            e2e_step()
        # the chunked path must give the same poses as the one-launch path
        T_chunks = np.concatenate([r["T"] for r in results])
        assert np.array_equal(T_chunks, res["T"]), "chunked e2e path disagrees with the device-resident run"
        barrier()
        ctx2.synchronize()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record(stream)
        for _ in range(K):
            out = e2e_step()  # every fetch synchronises its stream, so all work of the step is complete here
        s1.record(stream)
        barrier()
        ctx2.synchronize()
        ms_e2e = max_over_ranks(s0.elapsed_time(s1))
        for pl in pools:
            pl.destroy()
        ctx2.close()
        # level-0 images + per-pair job descriptor (272 B) + packed feature blob (65 B x padded N)
        h2d = (B + n_chunks) * frame_bytes + B * (272 + ((NFEAT + 15) // 16 * 16) * 65)
        d2h = B * (96 + 288 + 16) + B * NFEAT
        e2e = {"value": world * B * K / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "ms_per_step": ms_e2e / K}

    clocks = sampler.stop()
    other = None
    if rank == 0 and world == 1 and not args.no_extras:
        other = measure_other_paths(ctx, rank)

    # ---------------- CPU baseline (rank 0, N == 1 only) ----------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        from oracle import binding as ob
        from rpg_svo_b200 import synth

        n = min(args.cpu_sample, B)
        kind, note, run = cpu_runner(ob, synth, inp, n)
        run(1)
        reps, dt = 0, 0.0
        while dt < 10.0 and reps < 40:
            dt += run(1)
            reps += 1
        cpu = {"value": n * reps / dt, "unit": UNIT, "cores": 1, "kind": kind,
               "sample": f"first {n} frame pairs of the step x {reps} passes, single thread, {os.cpu_count()} host cores present",
               "note": note}

    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
        traffic = None
        prof = os.path.join(ROOT, "profiles", "sia_kernel_dram.json")
        if os.path.exists(prof):
            pj = json.load(open(prof))
            if pj.get("pairs_per_launch") == B:
                traffic = pj.get("dram_bytes_per_launch")
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": max(Wm, 3),
                "ms_per_step": kernel_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32 interpolation/residuals + f64 geometry/normal equations (as the reference)",
                "data": "synthetic",
                "config": {"workload": "C1 single stream 640x480 / 300 feats / levels 4..0 / 30 GN iters",
                           "pairs_per_step_per_gpu": B, "parallelism": f"{world} independent streams, no collective",
                           "l2_policy": "inputs larger than L2 (distinct pyramids per pair, ~%d MB/step/GPU)" % ((B + 1) * 409200 // 1000000)},
                "pose_rmse_vs_ref": rmse,
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": traffic, "peak_source": peak_src, "kernel": "svo::sia_kernel<1,false>",
                             "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": kernel_ms,
                             "mean_gn_iterations_per_pair": float(np.mean(stats["n_iters"]))},
                "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
                "other_paths": other}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    pool.destroy()
    ctx.close()


if __name__ == "__main__":
    main()
