"""Split-feature SparseImgAlign (one all-reduce of H/Jres/chi2/n_meas per GN iteration).

CPU / gloo / world_size 2: the distributed driver with the CPU oracle as the per-rank evaluator must land on
the single-process result.  GPU: the same driver over `svo_b200_sparse_residuals` (two "ranks" evaluated in
one process) must match the one-CTA kernel."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from rpg_svo_b200 import split_align, synth


def _oracle_evaluator(d):
    from oracle import binding as ob

    def evaluate(level, T, lo, hi, visible):
        return ob.sparse_residuals(d["ref_pyr"][level], d["cur_pyr"][level], level, d["cam"], T, d["px"][lo:hi],
                                   d["f"][lo:hi], d["pos"][lo:hi], d["has_point"][lo:hi], d["ref_pos"], visible_in=visible)

    return evaluate


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    d = synth.make_frame_pair(1000, n_feat=120, n_levels=4)
    r = split_align.sparse_img_align_split(_oracle_evaluator(d), synth.se3_identity(), 120, 3, 1, dist=dist, rank=rank,
                                           world=world)
    q.put((rank, r["T"], r["n_tracked"], r["n_allreduce"]))
    dist.destroy_process_group()


def test_split_two_ranks_gloo_matches_single_process(oracle):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    d = synth.make_frame_pair(1000, n_feat=120, n_levels=4)
    o = oracle.sparse_img_align(d["ref_pyr"], d["cur_pyr"], d["cam"], synth.se3_identity(), d["px"], d["f"], d["pos"],
                                d["has_point"], d["ref_pos"], 3, 1)
    assert np.array_equal(res[0][1], res[1][1])            # both ranks hold the same pose, no broadcast needed
    dt, dr = synth.pose_error(res[0][1], o["T"])
    assert dt < 1e-5 and dr < 1e-5, (dt, dr)               # summation order differs, nothing else
    assert res[0][2] == o["n_tracked"] and res[0][3] == len(o["trace"])  # one all-reduce per GN iteration


@pytest.mark.gpu
def test_split_on_gpu_matches_one_cta_kernel(ctx, pair300):
    d = pair300
    ref, cur = ctx.frame(d["ref_pyr"]), ctx.frame(d["cur_pyr"])
    ev = split_align.make_gpu_evaluator(ctx, ref, cur, d["cam"], d["px"], d["f"], d["pos"], d["has_point"], d["ref_pos"])
    # emulate a 2-way split in one process: evaluate both shards and add them (what the all-reduce does)
    from rpg_svo_b200 import shard

    def both(level, T, lo, hi, visible):
        out = None
        vis_parts = []
        for r in range(2):
            a, b = shard.shard_range(300, r, 2)
            p = ev(level, T, a, b, visible[a:b])
            vis_parts.append(p["visible"])
            chi2_sum = float(np.float32(p["chi2"]) * np.float32(p["n_meas"])) if p["n_meas"] else 0.0
            if out is None:
                out = dict(H=p["H"].copy(), Jres=p["Jres"].copy(), chi2_sum=chi2_sum, n_meas=p["n_meas"])
            else:
                out["H"] += p["H"]; out["Jres"] += p["Jres"]; out["chi2_sum"] += chi2_sum; out["n_meas"] += p["n_meas"]
        out["visible"] = np.concatenate(vis_parts)
        out["chi2"] = out["chi2_sum"] / out["n_meas"] if out["n_meas"] else float("nan")
        return out

    r = split_align.sparse_img_align_split(both, synth.se3_identity(), 300, 4, 0)
    g = ctx.sparse_img_align(ref, cur, d["cam"], synth.se3_identity(), d["px"], d["f"], d["pos"], d["has_point"],
                             d["ref_pos"], 4, 0)
    dt, dr = synth.pose_error(r["T"], g["T"])
    assert dt < 1e-5 and dr < 1e-5 and r["n_tracked"] == g["n_tracked"]
    ref.destroy(); cur.destroy()


# ---- the product path of the split: the kernels of the ranks exchange their sums through peer memory ---------------------
def _run_split_in_process(contexts, d, max_level, min_level, frames=None):
    """`world` contexts = `world` ranks sharing this process (and, on the one-GPU test box, the GPU): each rank aligns its
    contiguous slice of the pair's features from a host thread of its own; the kernels meet in the exchange buffers."""
    import threading

    from rpg_svo_b200 import shard

    world = len(contexts)
    if frames is None:
        frames = [(c.frame(d["ref_pyr"]), c.frame(d["cur_pyr"])) for c in contexts]
    out, errs = [None] * world, [None] * world

    def work(r):
        lo, hi = shard.shard_range(len(d["px"]), r, world)
        try:
            out[r] = contexts[r].sparse_img_align(frames[r][0], frames[r][1], d["cam"], synth.se3_identity(), d["px"][lo:hi],
                                                  d["f"][lo:hi], d["pos"][lo:hi], d["has_point"][lo:hi], d["ref_pos"], max_level,
                                                  min_level)
        except Exception as e:  # noqa: BLE001
            errs[r] = e

    th = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    return out, errs, frames


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 3])
def test_split_through_peer_memory_matches_the_single_kernel(ctx, oracle, pair300, world):
    """svo_b200_sia_split_*: every rank runs ONE kernel for the whole coarse-to-fine loop; the per-iteration sums cross
    between the ranks' kernels through the exchange buffers (system-scope stores + flags).  All ranks must end bit-identical,
    equal to the undivided kernel up to summation order, and the set-only visibility masks of the slices must concatenate to
    the undivided mask."""
    from rpg_svo_b200 import capi

    d = pair300
    contexts = [capi.Context(0) for _ in range(world)]
    try:
        ptrs = [c.sia_split_create(r, world, 1)[1] for r, c in enumerate(contexts)]
        for c in contexts:
            c.sia_split_connect(in_process_ptrs=ptrs)
        ref, cur = ctx.frame(d["ref_pyr"]), ctx.frame(d["cur_pyr"])
        g = ctx.sparse_img_align(ref, cur, d["cam"], synth.se3_identity(), d["px"], d["f"], d["pos"], d["has_point"], d["ref_pos"], 4, 0)
        frames = None
        for rep in range(2):  # the exchange sequence numbers carry over from launch to launch
            out, errs, frames = _run_split_in_process(contexts, d, 4, 0, frames)
            assert errs == [None] * world, errs
            for r in range(1, world):
                assert np.array_equal(out[r]["T"], out[0]["T"]) and np.array_equal(out[r]["H"], out[0]["H"])
                assert out[r]["n_tracked"] == out[0]["n_tracked"]
            dt, dr = synth.pose_error(out[0]["T"], g["T"])
            assert dt < 1e-7 and dr < 1e-7, (dt, dr)
            assert out[0]["n_tracked"] == g["n_tracked"]
            assert np.array_equal(np.concatenate([o["visible"] for o in out]), g["visible"])
            assert np.allclose(out[0]["H"], g["H"], rtol=1e-9, atol=1e-6)
        o = oracle.sparse_img_align(d["ref_pyr"], d["cur_pyr"], d["cam"], synth.se3_identity(), d["px"], d["f"], d["pos"],
                                    d["has_point"], d["ref_pos"], 4, 0)
        dt, dr = synth.pose_error(out[0]["T"], o["T"])
        assert dt < 1e-4 and dr < 1e-4
        for fr in frames:
            fr[0].destroy(); fr[1].destroy()
        ref.destroy(); cur.destroy()
    finally:
        for c in contexts:
            c.sia_split_destroy()
            c.close()


@pytest.mark.gpu
def test_split_peer_that_never_arrives_times_out_instead_of_hanging(pair300):
    """Only rank 0 of a 2-way split launches: its kernel waits ~2 s at the first exchange, gives up, and the call reports the
    error -- the GPU is never left spinning."""
    import time

    from rpg_svo_b200 import capi

    d = pair300
    contexts = [capi.Context(0) for _ in range(2)]
    try:
        ptrs = [c.sia_split_create(r, 2, 1)[1] for r, c in enumerate(contexts)]
        for c in contexts:
            c.sia_split_connect(in_process_ptrs=ptrs)
        ref, cur = contexts[0].frame(d["ref_pyr"]), contexts[0].frame(d["cur_pyr"])
        t0 = time.perf_counter()
        with pytest.raises(capi.SvoB200Error, match="did not arrive"):
            contexts[0].sparse_img_align(ref, cur, d["cam"], synth.se3_identity(), d["px"][:150], d["f"][:150], d["pos"][:150],
                                         d["has_point"][:150], d["ref_pos"], 4, 0)
        assert time.perf_counter() - t0 < 10.0   # one ~2 s timeout, not one per exchange
        ref.destroy(); cur.destroy()
    finally:
        for c in contexts:
            c.sia_split_destroy()
            c.close()
