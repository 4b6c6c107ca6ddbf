"""GPU: the CUDA kernels against the committed golden fixtures (tests/golden/*.npz)."""
import os

import numpy as np
import pytest

from rpg_svo_b200 import synth

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(G, name), allow_pickle=False)


@pytest.mark.parametrize("name", ["sia_c0.npz", "sia_c1.npz"])
def test_sia_golden(ctx, name):
    from tests.golden.make_golden import digest

    g = load(name)
    d = synth.make_frame_pair(int(g["seed"]), n_feat=int(g["n_feat"]), n_levels=int(g["n_levels"]))
    if digest(*d["ref_pyr"], *d["cur_pyr"], d["px"], d["f"], d["pos"], d["has_point"], d["ref_pos"]) != str(g["input_sha256"]):
        pytest.skip("synthetic inputs differ on this machine; fixture not comparable")
    ref, cur = ctx.frame(d["ref_pyr"]), ctx.frame(d["cur_pyr"])
    r = ctx.sparse_img_align(ref, cur, d["cam"], synth.se3_identity(), d["px"], d["f"], d["pos"], d["has_point"],
                             d["ref_pos"], int(g["max_level"]), int(g["min_level"]), want_trace=True)
    dt, dr = synth.pose_error(r["T"], g["T"])
    assert dt <= 1e-4 and dr <= 1e-4
    assert np.array_equal(r["visible"], g["visible"]) and r["n_tracked"] == int(g["n_tracked"])
    assert [t["accepted"] for t in r["trace"]] == list(g["trace_accepted"])
    assert np.allclose(r["H"], g["H"], rtol=1e-9)
    ref.destroy(); cur.destroy()


def test_align_golden(ctx):
    g = load("align.npz")
    fr = ctx.frame([g["img"]])
    m = len(g["px_start"])
    conv, px = ctx.align2d_batch(fr, np.zeros(m, np.int32), g["pwb"], g["patch"], 10, g["px_start"])
    assert np.array_equal(conv, g["conv2d"]) and np.array_equal(px, g["px2d"])
    conv, px, h = ctx.align1d_batch(fr, np.zeros(m, np.int32), g["dir"], g["pwb"], g["patch"], 10, g["px_start"])
    assert np.array_equal(conv, g["conv1d"]) and np.array_equal(px, g["px1d"]) and np.array_equal(h, g["h_inv"])
    fr.destroy()


def test_pose_opt_golden(ctx):
    g = load("pose_opt.npz")
    o = ctx.pose_optimize(2.0, 10, float(g["fx"]), g["T_init"], g["f"], g["pos"], g["level"], g["has_point"])
    dt, dr = synth.pose_error(o["T"], g["T"])
    assert dt < 1e-8 and dr < 1e-8
    assert np.array_equal(o["has_point"], g["has_point_out"]) and o["num_obs"] == int(g["num_obs"])
    assert np.isclose(o["error_final"], g["error_final"], rtol=1e-9)
