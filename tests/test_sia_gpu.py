"""GPU parity: svo_b200_sparse_img_align / svo_b200_sparse_residuals vs the CPU oracle.

Tolerances are the ones BASELINE.json's north_star states: 1e-4 on the final SE3 and on per-patch
residuals, bit-exact visibility masks.
"""
import numpy as np
import pytest

from rpg_svo_b200 import synth

pytestmark = pytest.mark.gpu

POSE_TOL = 1e-4
RES_TOL = 1e-4


def _run_both(ctx, oracle, d, max_level, min_level, n_iter=30, T0=None, trace=True):
    T0 = synth.se3_identity() if T0 is None else T0
    ref = ctx.frame(d["ref_pyr"])
    cur = ctx.frame(d["cur_pyr"])
    g = ctx.sparse_img_align(ref, cur, d["cam"], T0, d["px"], d["f"], d["pos"], d["has_point"],
                             d["ref_pos"], max_level, min_level, n_iter, want_trace=trace)
    o = oracle.sparse_img_align(d["ref_pyr"], d["cur_pyr"], d["cam"], T0, d["px"], d["f"], d["pos"],
                                d["has_point"], d["ref_pos"], max_level, min_level, n_iter)
    ref.destroy()
    cur.destroy()
    return g, o


@pytest.mark.parametrize("levels", [(4, 0), (4, 2), (2, 0), (0, 0)])
def test_final_pose_and_mask_c1(ctx, oracle, pair300, levels):
    g, o = _run_both(ctx, oracle, pair300, *levels)
    dt, dr = synth.pose_error(g["T"], o["T"])
    assert dt <= POSE_TOL and dr <= POSE_TOL, (dt, dr)
    assert np.array_equal(g["visible"], o["visible"])  # bit-exact mask
    assert g["n_tracked"] == o["n_tracked"]
    # both must actually have tracked the motion
    assert synth.pose_error(g["T"], pair300["T_cur_ref_gt"])[0] < 1e-3


def test_iteration_trace_matches(ctx, oracle, pair300):
    g, o = _run_both(ctx, oracle, pair300, 4, 0)
    assert len(g["trace"]) == len(o["trace"])
    for a, b in zip(g["trace"], o["trace"]):
        assert (a["level"], a["iter"], a["accepted"], a["n_meas"]) == (b["level"], b["iter"], b["accepted"], b["n_meas"])
        assert abs(a["chi2"] - b["chi2"]) <= 1e-4 * max(1.0, abs(b["chi2"]))
        assert np.allclose(a["x"], b["x"], rtol=1e-4, atol=1e-7)
        assert np.allclose(a["T"], b["T"], atol=1e-6)


@pytest.mark.parametrize("level", [0, 1, 2, 3, 4])
def test_residual_pass_matches(ctx, oracle, pair300, level):
    d = pair300
    T = synth.se3_exp(np.array([0.004, -0.003, 0.002, 0.001, -0.002, 0.0015]))
    ref = ctx.frame(d["ref_pyr"])
    cur = ctx.frame(d["cur_pyr"])
    g = ctx.sparse_residuals(ref, cur, d["cam"], level, T, d["px"], d["f"], d["pos"], d["has_point"], d["ref_pos"])
    o = oracle.sparse_residuals(d["ref_pyr"][level], d["cur_pyr"][level], level, d["cam"], T, d["px"],
                                d["f"], d["pos"], d["has_point"], d["ref_pos"])
    assert np.array_equal(g["visible"], o["visible"])
    assert np.array_equal(g["in_image"], o["in_image"])
    v = o["visible"].astype(bool)
    assert np.array_equal(g["ref_patch"][v], o["ref_patch"][v])  # f32 stage: bit-exact by construction
    m = o["in_image"].astype(bool)
    assert np.max(np.abs(g["residuals"][m] - o["residuals"][m])) <= RES_TOL
    assert np.all(np.isnan(g["residuals"][~m]))
    assert g["n_meas"] == o["n_meas"]
    assert abs(g["chi2"] - o["chi2"]) <= 1e-5 * abs(o["chi2"])
    assert np.allclose(g["H"], o["H"], rtol=1e-9, atol=1e-6)
    assert np.allclose(g["Jres"], o["Jres"], rtol=1e-5, atol=1e-3)


def test_no_features_returns_zero(ctx, pair300):
    d = pair300
    ref = ctx.frame(d["ref_pyr"])
    cur = ctx.frame(d["cur_pyr"])
    e = np.zeros((0, 3))
    g = ctx.sparse_img_align(ref, cur, d["cam"], synth.se3_identity(), np.zeros((0, 2)), e, e,
                             np.zeros(0, np.uint8), d["ref_pos"], 4, 0)
    assert g["n_tracked"] == 0 and np.allclose(g["T"], synth.se3_identity())
