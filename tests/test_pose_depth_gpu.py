"""GPU parity: pose_optimizer and DepthFilter kernels vs the CPU oracle."""
import numpy as np
import pytest

from rpg_svo_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,size", [(1000, (1920, 1080)), (120, (752, 480)), (7, (640, 480))])
def test_pose_optimize_matches_oracle(ctx, oracle, n, size):
    c = synth.make_pose_opt_case(5 + n, n, *size)
    g = ctx.pose_optimize(2.0, 10, c["cam"].fx, c["T_init"], c["f"], c["pos"], c["level"], c["has_point"])
    o = oracle.pose_optimize(2.0, 10, c["cam"].fx, c["T_init"], c["f"], c["pos"], c["level"], c["has_point"])
    assert np.array_equal(g["has_point"], o["has_point"])  # culling mask bit-exact
    assert g["num_obs"] == o["num_obs"] and g["n_iter_done"] == o["n_iter_done"]
    dt, dr = synth.pose_error(g["T"], o["T"])
    assert dt < 1e-8 and dr < 1e-8, (dt, dr)
    for k in ("estimated_scale", "error_init", "error_final"):
        assert abs(g[k] - o[k]) <= 1e-9 * max(1.0, abs(o[k])), k
    assert np.allclose(g["cov"], o["cov"], rtol=1e-6, atol=1e-12)
    if n >= 100:
        assert synth.pose_error(g["T"], c["T_true"])[0] < 5e-3  # the optimiser recovers the pose


def test_pose_optimize_no_observations(ctx):
    c = synth.make_pose_opt_case(3, 16, 640, 480)
    hp = np.zeros(16, np.uint8)
    g = ctx.pose_optimize(2.0, 10, c["cam"].fx, c["T_init"], c["f"], c["pos"], c["level"], hp)
    assert g["num_obs"] == 0 and np.allclose(g["T"], c["T_init"])


@pytest.mark.parametrize("n_seeds,baseline", [(2000, 0.3), (300, 0.05), (300, 0.6)])
def test_depth_filter_update_matches_oracle(ctx, oracle, n_seeds, baseline):
    c = synth.make_depth_case(31 + n_seeds, n_seeds, baseline=baseline)
    ref, cur = ctx.frame(c["ref_pyr"]), ctx.frame(c["cur_pyr"])
    g = ctx.depth_filter_update([ref], [c["T_ref_w"]], cur, c["T_cur_w"], c["cam"], c["ref_index"], c["ftr_px"],
                                c["ftr_f"], c["ftr_level"], c["ftr_type"], c["ftr_grad"], c["batch_id"],
                                c["batch_counter"], c["seeds"])
    o = oracle.depth_filter_update([c["ref_pyr"]], [c["T_ref_w"]], c["cur_pyr"], c["T_cur_w"], c["cam"],
                                   c["ref_index"], c["ftr_px"], c["ftr_f"], c["ftr_level"], c["ftr_type"],
                                   c["ftr_grad"], c["batch_id"], c["batch_counter"], c["seeds"])
    assert np.array_equal(g["status"], o["status"])      # per-seed outcome bit-exact
    assert np.array_equal(g["n_zmssd"], o["n_zmssd"])    # same pixels scored along the epipolar line
    upd = o["status"] >= 5
    assert upd.sum() > 0.3 * n_seeds
    assert np.max(np.abs(g["px_cur"][upd] - o["px_cur"][upd])) <= 1e-4
    assert np.allclose(g["z"][upd], o["z"][upd], rtol=1e-6)
    for k in ("a", "b", "mu", "sigma2"):
        assert np.allclose(g[k], o[k], rtol=2e-5, atol=1e-7), k
    # the measurements are real: triangulated depth close to the plane depth
    assert np.median(np.abs(g["z"][upd] - c["depth_gt"][upd])) < 0.05
    ref.destroy(); cur.destroy()
