"""GPU: the C++ host classes (rpg_svo_b200/host/svo_host.h: svo::Frame / Feature / Point /
SparseImgAlign / pose_optimizer with the reference's signatures) driven by a small C++ program, checked
against the CPU oracle."""
import os
import struct
import subprocess

import numpy as np
import pytest

from rpg_svo_b200 import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "rpg_svo_b200", "host")


def build_demo(name="host_demo"):
    exe = os.path.join(HOST, name)
    src = os.path.join(HOST, name + ".cpp")
    if not os.path.exists(exe) or os.path.getmtime(exe) < max(os.path.getmtime(src), os.path.getmtime(os.path.join(HOST, "svo_host.h"))):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-o", exe, src, "-L" + os.path.join(ROOT, "rpg_svo_b200"),
                               "-lsvo_b200", "-Wl,-rpath,$ORIGIN/.."])
    return exe


def test_cpp_host_surface(tmp_path, oracle):
    d = synth.make_frame_pair(1003, n_feat=200, n_levels=5)
    cam, N = d["cam"], 200
    rng = np.random.default_rng(1)
    Tc = d["T_cur_w"]
    pc = d["pos"] @ Tc[:, :3].T + Tc[:, 3]
    px_cur = cam.world2cam(pc) + rng.normal(0, 0.5, (N, 2))
    f_cur = cam.cam2world(px_cur)
    level = rng.integers(0, 3, N).astype(np.int32)
    inp, outp = tmp_path / "in.bin", tmp_path / "out.bin"
    with open(inp, "wb") as fh:
        fh.write(struct.pack("6i", cam.width, cam.height, 5, N, 4, 0))
        fh.write(struct.pack("4d", cam.fx, cam.fy, cam.cx, cam.cy))
        fh.write(d["ref_pyr"][0].tobytes()); fh.write(d["cur_pyr"][0].tobytes())
        fh.write(np.ascontiguousarray(d["T_ref_w"]).tobytes())
        fh.write(np.ascontiguousarray(d["T_ref_w"]).tobytes())  # cur starts at the reference pose
        for a in (d["px"], d["f"], d["pos"]):
            fh.write(np.ascontiguousarray(a, np.float64).tobytes())
        fh.write(d["has_point"].tobytes()); fh.write(np.ascontiguousarray(f_cur).tobytes()); fh.write(level.tobytes())
    subprocess.check_call([build_demo(), str(inp), str(outp)])
    raw = open(outp, "rb").read()
    T_align = np.frombuffer(raw, np.float64, 12, 0).reshape(3, 4)
    n_tracked = struct.unpack_from("q", raw, 96)[0]
    fisher = np.frombuffer(raw, np.float64, 36, 104).reshape(6, 6)
    T_opt = np.frombuffer(raw, np.float64, 12, 392).reshape(3, 4)
    est_scale, e_init, e_final = np.frombuffer(raw, np.float64, 3, 488)
    num_obs = struct.unpack_from("q", raw, 512)[0]
    hp_after = np.frombuffer(raw, np.uint8, N, 520)

    o = oracle.sparse_img_align(d["ref_pyr"], d["cur_pyr"], cam, synth.se3_identity(), d["px"], d["f"], d["pos"],
                                d["has_point"], d["ref_pos"], 4, 0)
    T_cur_w_oracle = oracle.se3_mul(o["T"], d["T_ref_w"])  # cur.T_f_w_ = T_cur_from_ref * ref.T_f_w_
    dt, dr = synth.pose_error(T_align, T_cur_w_oracle)
    assert dt < 1e-4 and dr < 1e-4 and n_tracked == o["n_tracked"]
    assert np.allclose(fisher, o["H"] / (5e-4 * 255 * 255), rtol=1e-8)
    po = oracle.pose_optimize(2.0, 10, cam.fx, T_cur_w_oracle, f_cur, d["pos"], level, d["has_point"])
    dt, dr = synth.pose_error(T_opt, po["T"])
    assert dt < 1e-6 and dr < 1e-6
    assert num_obs == po["num_obs"] and np.array_equal(hp_after, po["has_point"])
    assert np.isclose(e_final, po["error_final"], rtol=1e-6) and np.isclose(est_scale, po["estimated_scale"], rtol=1e-6)
    # FastDetector over the frame's device pyramid, cells of the existing features excluded
    n_new = struct.unpack_from("i", raw, 520 + N)[0]
    new = np.frombuffer(raw, np.int32, 3 * n_new, 524 + N).reshape(n_new, 3)
    n_cols = int(np.ceil(cam.width / 30))
    occ = np.zeros(n_cols * int(np.ceil(cam.height / 30)), np.uint8)
    occ[(d["px"][:, 1] / 30).astype(int) * n_cols + (d["px"][:, 0] / 30).astype(int)] = 1
    fo = oracle.fast_detect(d["ref_pyr"], 3, 30, 20.0, occ)
    assert n_new == len(fo["x"]) and n_new > 20
    assert np.array_equal(new[:, 0], fo["x"]) and np.array_equal(new[:, 1], fo["y"]) and np.array_equal(new[:, 2], fo["level"])


def test_cpp_host_reprojector(tmp_path, oracle):
    """svo::Reprojector of svo_host.h (pointer graph -> flat view -> svo_b200_reproject_map -> side effects applied to
    the Frame / Point / Map objects) against the sequential oracle."""
    c = synth.make_map_case(12, n_kfs=6, n_points=500)
    v, cam, opt = c["view"], c["cam"], c["options"]
    inp, outp = tmp_path / "in.bin", tmp_path / "out.bin"
    with open(inp, "wb") as fh:
        fh.write(struct.pack("14i", cam.width, cam.height, c["n_levels"], v["n_kfs"], v["n_ftrs"], v["n_points"], v["n_candidates"],
                             len(v["kf_fts"]), len(v["pt_obs"]), len(c["cell_order"]), opt["grid_size"], opt["max_fts"],
                             opt["max_n_kfs"], opt["max_search_level"] + 1))
        fh.write(struct.pack("4d", cam.fx, cam.fy, cam.cx, cam.cy))
        for pyr in c["kf_pyr"]:
            fh.write(pyr[0].tobytes())
        fh.write(c["cur_pyr"][0].tobytes())
        fh.write(np.ascontiguousarray(v["kf_T_f_w"], np.float64).tobytes())
        fh.write(np.ascontiguousarray(c["cur_T_f_w"], np.float64).tobytes())
        fh.write(np.ascontiguousarray(c["keypt_ftr"], np.int32).tobytes())
        for k, dt in (("kf_fts_offset", np.int32), ("kf_fts", np.int32), ("ftr_kf", np.int32), ("ftr_px", np.float64),
                      ("ftr_f", np.float64), ("ftr_level", np.int32), ("ftr_type", np.int32), ("ftr_grad", np.float64),
                      ("ftr_point", np.int32), ("pt_pos", np.float64), ("pt_obs_offset", np.int32), ("pt_obs", np.int32),
                      ("cand_point", np.int32)):
            fh.write(np.ascontiguousarray(v[k], dt).tobytes())
        for k in ("pt_type", "pt_n_failed", "pt_n_succeeded", "cell_order"):
            fh.write(np.ascontiguousarray(c[k], np.int32).tobytes())
    subprocess.check_call([build_demo("host_reproject_demo"), str(inp), str(outp)])
    raw = open(outp, "rb").read()
    n_matches, n_trials, n_new, n_ov = struct.unpack_from("4q", raw, 0)
    off = 32
    ov = np.frombuffer(raw, np.int64, 2 * n_ov, off).reshape(n_ov, 2); off += 16 * n_ov
    rec = np.dtype([("pt", "<i4"), ("level", "<i4"), ("type", "<i4"), ("px", "<f8", 2), ("grad", "<f8", 2)])
    new = np.frombuffer(raw, rec, n_new, off); off += rec.itemsize * n_new
    pst = np.frombuffer(raw, np.int32, 3 * v["n_points"], off).reshape(-1, 3)

    o = oracle.reproject_map(c)
    assert (n_matches, n_trials, n_new, n_ov) == (o["n_matches"], o["n_trials"], o["n_new"], o["n_overlap"])
    assert np.array_equal(ov[:, 0], o["overlap_kf"]) and np.array_equal(ov[:, 1], o["overlap_count"])
    assert np.array_equal(new["pt"], o["new_point"]) and np.array_equal(new["level"], o["new_level"])
    assert np.array_equal(new["type"], o["new_type"])
    assert np.max(np.abs(new["px"] - o["new_px"])) <= 1e-4
    edge = o["new_type"] == 1
    assert np.allclose(new["grad"][edge], o["new_grad"][edge], atol=1e-9)
    assert np.array_equal(pst[:, 0], o["pt_type"]) and np.array_equal(pst[:, 1], o["pt_n_failed"])
    assert np.array_equal(pst[:, 2], o["pt_n_succeeded"])
