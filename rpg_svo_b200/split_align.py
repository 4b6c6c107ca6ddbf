"""Split-feature SparseImgAlign: one frame pair, its features partitioned over the ranks of a process
group, the Gauss-Newton normal equations summed with ONE all-reduce per iteration (SURVEY.md 8e:
"optional single-stream split ... 29 doubles per iteration over NCCL").

This is the demonstration mode the north star asks for, not the throughput path: a pair needs ~17
strictly sequential iterations and each now carries a kernel launch, a device->host copy and a
latency-bound all-reduce of 44 doubles, so it is slower than running the whole pair in the one-CTA
kernel (`Context.sparse_img_align`).  It exists to show that the per-iteration state really is just
(H, Jres, chi2, n_meas) and that summing it across devices reproduces the single-device result.

Every rank evaluates `computeResiduals` for ITS features with `svo_b200_sparse_residuals` (one launch of
the same CUDA kernel, EVAL mode), then the host runs the reference's GN control flow
([EXT] vk::NLLSSolver::optimizeGaussNewton as used by svo/src/sparse_img_align.cpp:61-69,245-258)
identically on every rank, so all ranks hold the same pose at every step without a broadcast.
"""
from __future__ import annotations

import numpy as np

from . import shard, synth


def _all_reduce_sum(vec: np.ndarray, dist, device) -> np.ndarray:
    if dist is None or not dist.is_available() or not dist.is_initialized():
        return vec
    import torch

    t = torch.from_numpy(vec.copy())
    if device is not None:
        t = t.to(device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()


def sparse_img_align_split(evaluate, T_init, n_features: int, max_level: int, min_level: int, n_iter: int = 30,
                           eps: float = 1e-6, dist=None, device=None, rank: int = 0, world: int = 1):
    """`evaluate(level, T, lo, hi, visible)` -> dict(H 6x6, Jres 6, chi2 (mean), n_meas, visible) for the
    features [lo, hi) owned by this rank (e.g. a closure over `Context.sparse_residuals`).
    Returns dict(T, n_tracked, n_iters, n_allreduce)."""
    lo, hi = shard.shard_range(n_features, rank, world)
    visible = np.zeros(hi - lo, np.uint8)  # set-only across levels, like visible_fts_
    T = np.array(T_init, dtype=np.float64).reshape(3, 4)
    chi2_prev, stop = 1e10, False
    n_iters = n_allreduce = 0
    n_meas_last = 0
    for level in range(max_level, min_level - 1, -1):
        T_old = T.copy()
        for it in range(n_iter):
            r = evaluate(level, T, lo, hi, visible)
            visible = r["visible"]
            n_loc = float(r["n_meas"])
            chi2_sum_loc = float(np.float32(r["chi2"]) * np.float32(n_loc)) if n_loc > 0 else 0.0
            packed = np.concatenate([np.asarray(r["H"], np.float64).ravel(), np.asarray(r["Jres"], np.float64),
                                     [chi2_sum_loc, n_loc]])
            tot = _all_reduce_sum(packed, dist, device)  # 36 + 6 + 2 doubles: the only exchange of the iteration
            n_allreduce += 1
            n_iters += 1
            H, Jres, chi2_sum, n_meas = tot[:36].reshape(6, 6), tot[36:42], tot[42], tot[43]
            n_meas_last = int(n_meas)
            new_chi2 = float(np.float32(chi2_sum) / np.float32(n_meas)) if n_meas > 0 else float("nan")
            if n_meas == 0:
                x = np.zeros(6)  # Eigen's LDLT of the zero matrix solves to 0
            else:
                try:
                    x = np.linalg.solve(H, Jres)
                except np.linalg.LinAlgError:
                    x = np.full(6, np.nan)
            if np.isnan(x[0]):
                stop = True
            if (it > 0 and new_chi2 > chi2_prev) or stop:
                T = T_old  # rollback
                break
            T_new = synth.se3_mul(T, synth.se3_exp(-x))  # T * exp(-x)
            T_old, T, chi2_prev = T, T_new, new_chi2
            if np.max(np.abs(x)) <= eps:
                break
    return dict(T=T, n_tracked=n_meas_last // 16, n_iters=n_iters, n_allreduce=n_allreduce)


def make_gpu_evaluator(ctx, ref, cur, cam, px, f, pos, has_point, ref_pos):
    """Per-rank evaluator on top of the C ABI (`svo_b200_sparse_residuals`)."""

    def evaluate(level, T, lo, hi, visible):
        r = ctx.sparse_residuals(ref, cur, cam, level, T, px[lo:hi], f[lo:hi], pos[lo:hi], has_point[lo:hi], ref_pos,
                                 visible_in=visible)
        return r

    return evaluate
