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


def test_detect_golden(ctx):
    g = load("detect.npz")
    fr = ctx.frame(synth.build_pyramid(g["img"], 3))
    r = ctx.fast_detect(fr, 30, 3, 20.0, g["occupancy"])
    assert all(np.array_equal(r[k], g[k]) for k in ("x", "y", "level"))
    assert np.array_equal(r["score"].view(np.uint32), g["score"].view(np.uint32))
    fr.destroy()


def test_reproject_golden(ctx):
    from tests.test_golden_cpu import _map_case_for

    g = load("reproject.npz")
    c = _map_case_for(g)
    kfs, cur = [ctx.frame(p) for p in c["kf_pyr"]], ctx.frame(c["cur_pyr"])
    r = ctx.reproject_map(c["view"], kfs, cur, c["cur_T_f_w"], c["cam"], c["options"], c["cell_order"], c["pt_type"],
                          c["pt_n_failed"], c["pt_n_succeeded"])
    assert r["n_matches"] == int(g["n_matches"]) and r["n_trials"] == int(g["n_trials"])
    for k in ("new_point", "new_level", "new_type", "pt_type", "pt_n_failed", "pt_n_succeeded", "pt_action", "overlap_kf", "overlap_count"):
        assert np.array_equal(r[k], g[k]), k
    assert np.max(np.abs(r["new_px"] - g["new_px"]), initial=0.0) <= 1e-4
    for f in kfs + [cur]:
        f.destroy()
