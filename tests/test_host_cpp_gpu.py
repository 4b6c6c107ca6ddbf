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


def build_demo():
    exe = os.path.join(HOST, "host_demo")
    src = os.path.join(HOST, "host_demo.cpp")
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
