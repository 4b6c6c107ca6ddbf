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
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-pthread", "-o", exe, src, "-L" + os.path.join(ROOT, "rpg_svo_b200"),
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


def _write_camera(fh, cam):
    d = (list(getattr(cam, "d", ())) + [0.0] * 5)[:5]
    fh.write(struct.pack("3i", cam.width, cam.height, int(getattr(cam, "model", 0))))
    fh.write(struct.pack("9d", cam.fx, cam.fy, cam.cx, cam.cy, *d))


@pytest.mark.parametrize("kind", ["pinhole", "atan"])
def test_cpp_host_units_matcher_alignment_depth_filter(tmp_path, oracle, kind):
    """host_pipeline_demo `units`: svo::Matcher (findMatchDirect / findEpipolarMatchDirect + scratch members),
    feature_alignment::align2D / align1D with the reference's argument lists, and DepthFilter::updateSeeds driven over
    several frames until seeds converge into MapPointCandidates -- each against the CPU oracle."""
    cam = synth.camera_for(752, 480) if kind == "pinhole" else synth.reference_param_camera("atan")
    tv = synth.make_depth_case(91, 60, baseline=0.25, cam=cam)
    rng = np.random.default_rng(5)
    n_levels = tv["n_levels"]
    T_ref_w, T_cur_w = tv["T_ref_w"], tv["T_cur_w"]
    # findMatchDirect candidates on the same two frames
    M = 40
    m_level = rng.integers(0, 3, M).astype(np.int32)
    m_px = np.stack([rng.uniform(60, cam.width - 60, M), rng.uniform(60, cam.height - 60, M)], axis=1)
    m_px = np.round(m_px / (1 << m_level)[:, None]) * (1 << m_level)[:, None]
    m_f = cam.cam2world(m_px)
    m_pos = synth.intersect(tv["plane"], T_ref_w, m_f)
    m_pxcur = cam.world2cam(m_pos @ T_cur_w[:, :3].T + T_cur_w[:, 3]) + rng.uniform(-1.5, 1.5, (M, 2))
    m_type = (rng.uniform(size=M) < 0.2).astype(np.int32)
    ang = rng.uniform(0, 2 * np.pi, M)
    m_grad = np.stack([np.cos(ang), np.sin(ang)], axis=1)
    # epipolar candidates = the depth case's seeds
    E = tv["M"]
    mu, sig = 0.5, np.sqrt(tv["seeds"]["sigma2"].astype(np.float64))
    e_d = np.stack([np.full(E, 1.0 / mu), 1.0 / (mu + sig), 1.0 / np.maximum(mu - sig, 1e-7)], axis=1)
    # alignment problems on the current frame's levels
    A = 30
    a_level = rng.integers(0, 3, A).astype(np.int32)
    a_pwb, a_patch, a_px0, a_dir = np.zeros((A, 100), np.uint8), np.zeros((A, 64), np.uint8), np.zeros((A, 2)), np.zeros((A, 2), np.float32)
    for i in range(A):
        im = tv["cur_pyr"][a_level[i]]
        pt = np.array([rng.uniform(14, im.shape[1] - 14), rng.uniform(14, im.shape[0] - 14)])
        p = synth.patch_with_border(im, pt)
        a_pwb[i], a_patch[i] = p.ravel(), p[1:9, 1:9].ravel()
        off = rng.uniform(-1.2, 1.2, 2)
        a_px0[i] = pt - off
        a_dir[i] = (off / np.linalg.norm(off)).astype(np.float32)
    # depth filter: S seeds on the keyframe, F further frames
    S, F = 120, 14
    s_px = np.floor(synth.jittered_features(rng, cam, S, margin=40.0))
    s_level = np.zeros(S, np.int32)
    depth_mean, depth_min = 2.0, 1.0
    plane, tex = tv["plane"], synth.make_texture(7)
    f_T, f_pyr = [], []
    for k in range(F):
        dvec = rng.normal(size=3); dvec[2] *= 0.2; dvec *= 0.5 / np.linalg.norm(dvec)
        T = synth.se3_mul(synth.se3_exp(np.concatenate([dvec, np.deg2rad(rng.uniform(-2, 2, 3))])), T_ref_w)
        f_T.append(T)
        f_pyr.append(synth.build_pyramid(synth.render(cam, T, plane, tex), n_levels))
    inp, outp = tmp_path / "in.bin", tmp_path / "out.bin"
    with open(inp, "wb") as fh:
        _write_camera(fh, cam)
        fh.write(struct.pack("6i", n_levels, M, E, A, S, F))
        fh.write(tv["ref_pyr"][0].tobytes()); fh.write(tv["cur_pyr"][0].tobytes())
        fh.write(np.ascontiguousarray(T_ref_w).tobytes()); fh.write(np.ascontiguousarray(T_cur_w).tobytes())
        for a in (m_px, m_f, m_level, m_type, m_grad, m_pos, m_pxcur):
            fh.write(np.ascontiguousarray(a).tobytes())
        for a in (tv["ftr_px"], tv["ftr_f"], tv["ftr_level"].astype(np.int32), tv["ftr_type"].astype(np.int32), tv["ftr_grad"], e_d):
            fh.write(np.ascontiguousarray(a).tobytes())
        for a in (a_level, a_pwb, a_patch, a_px0, a_dir):
            fh.write(np.ascontiguousarray(a).tobytes())
        fh.write(np.ascontiguousarray(s_px).tobytes()); fh.write(s_level.tobytes()); fh.write(struct.pack("2d", depth_mean, depth_min))
        for k in range(F):
            fh.write(f_pyr[k][0].tobytes()); fh.write(np.ascontiguousarray(f_T[k]).tobytes())
    subprocess.check_call([build_demo("host_pipeline_demo"), "units", str(inp), str(outp)])
    raw = np.fromfile(outp, np.uint8)
    off = 0

    def take(dtype, n):
        nonlocal off
        a = np.frombuffer(raw.tobytes(), dtype, n, off)
        off += a.nbytes
        return a

    T_cur_ref = synth.se3_mul(T_cur_w, synth.se3_inv(T_ref_w))
    ref_pos = synth.se3_inv(T_ref_w)[:, 3]
    mrec = take(np.float64, 8 * M).reshape(M, 8)
    n_ok = 0
    for i in range(M):
        o = oracle.find_match_direct(tv["ref_pyr"], tv["cur_pyr"], cam, T_cur_ref, m_px[i], m_f[i], int(m_level[i]), int(m_type[i]),
                                     m_grad[i], np.linalg.norm(m_pos[i] - ref_pos), 2, 10, m_pxcur[i])
        assert bool(mrec[i, 0]) == bool(o["success"]) and int(mrec[i, 3]) == o["search_level"], i
        assert np.allclose(mrec[i, 4:8].reshape(2, 2), o["A_cur_ref"], rtol=1e-7, atol=1e-9), i
        if o["success"]:
            n_ok += 1
            assert np.max(np.abs(mrec[i, 1:3] - o["px_cur"])) <= 1e-4, i
    assert n_ok > M // 2
    erec = take(np.float64, 7 * E).reshape(E, 7)
    n_ok = 0
    for i in range(E):
        o = oracle.find_epipolar_match_direct(tv["ref_pyr"], tv["cur_pyr"], cam, T_cur_ref, tv["ftr_px"][i], tv["ftr_f"][i],
                                              int(tv["ftr_level"][i]), int(tv["ftr_type"][i]), tv["ftr_grad"][i], *e_d[i], 2)
        assert bool(erec[i, 0]) == bool(o["success"]) and bool(erec[i, 6]) == bool(o["reject"]), i
        if o["success"]:
            n_ok += 1
            assert np.isclose(erec[i, 1], o["depth"], rtol=1e-6) and np.max(np.abs(erec[i, 2:4] - o["px_cur"])) <= 1e-4, i
            assert int(erec[i, 4]) == o["search_level"] and np.isclose(erec[i, 5], o["epi_length"], rtol=1e-9), i
    assert n_ok > E // 4
    arec = take(np.float64, 7 * A).reshape(A, 7)
    for i in range(A):
        ok2, p2 = oracle.align2d(tv["cur_pyr"][a_level[i]], a_pwb[i], a_patch[i], 10, a_px0[i])
        ok1, p1, hinv = oracle.align1d(tv["cur_pyr"][a_level[i]], a_dir[i], a_pwb[i], a_patch[i], 10, a_px0[i])
        assert bool(arec[i, 0]) == bool(ok2) and np.array_equal(arec[i, 1:3], p2), i      # bit-exact, as the batch kernels
        assert bool(arec[i, 3]) == bool(ok1) and np.array_equal(arec[i, 4:6], p1) and arec[i, 6] == hinv, i
    # depth filter chain replayed with the oracle: same list semantics (erase converged / NaN seeds, keep the rest)
    s_f = cam.cam2world(s_px)
    z_range = np.float32(1.0 / depth_min)
    seeds = dict(a=np.full(S, 10, np.float32), b=np.full(S, 10, np.float32), mu=np.full(S, np.float32(1.0 / depth_mean), np.float32),
                 z_range=np.full(S, z_range, np.float32), sigma2=np.full(S, z_range * z_range / np.float32(36), np.float32))
    alive = np.arange(S)
    cands = []
    T_w_ref = synth.se3_inv(T_ref_w)
    counts = take(np.int64, 3 * F).reshape(F, 3)
    n_upd = 0
    for k in range(F):
        n = len(alive)
        o = oracle.depth_filter_update([tv["ref_pyr"]], [T_ref_w], f_pyr[k], f_T[k], cam, np.zeros(n, np.int32), s_px[alive], s_f[alive],
                                       s_level[alive], np.zeros(n, np.int32), np.tile([1.0, 0.0], (n, 1)), np.ones(n, np.int32), 1,
                                       {key: v[alive] for key, v in seeds.items()})
        for key in ("a", "b", "mu", "sigma2"):
            seeds[key][alive] = o[key]
        st = o["status"]
        n_upd += int((st >= 5).sum())
        for j in np.nonzero(st == 6)[0]:
            i = alive[j]
            cands.append((s_px[i], T_w_ref[:, :3] @ (s_f[i] * (1.0 / np.float64(seeds["mu"][i]))) + T_w_ref[:, 3]))
        alive = alive[(st != 6) & (st != 7) & (st != 1)]
        assert counts[k, 0] == len(alive) and counts[k, 1] == len(cands) and counts[k, 2] == n_upd, k
    ns = take(np.int64, 1)[0]
    srec = take(np.float64, 7 * ns).reshape(ns, 7)
    assert ns == len(alive) and np.array_equal(srec[:, :2], s_px[alive])
    for c, key in enumerate(("a", "b", "mu", "z_range", "sigma2"), start=2):
        assert np.allclose(srec[:, c], seeds[key][alive], rtol=5e-5, atol=1e-7), key
    nc = take(np.int64, 1)[0]
    crec = take(np.float64, 5 * nc).reshape(nc, 5)
    assert nc == len(cands) > S // 10
    for r, (px, pos) in zip(crec, cands):
        assert np.array_equal(r[:2], px) and np.allclose(r[2:], pos, rtol=0, atol=1e-4)
    # the converged points lie on the rendered plane
    d_plane = np.abs(crec[:, 2:] @ plane.n - plane.d)
    assert np.median(d_plane) < 0.02


@pytest.mark.parametrize("mapper_thread", [1, 0])
def test_cpp_frame_handler_pipeline_two_threads(tmp_path, mapper_thread):
    """host_pipeline_demo `pipeline`: svo::FrameHandlerMono::addImage chains SparseImgAlign -> Reprojector -> pose optimizer
    -> Point::optimize on the tracking context while the DepthFilter thread updates seeds on its own context.  Every stage
    has its own parity test; here the chain must track the synthetic stream and the mapper must make progress."""
    K, N = 7, 220
    st = synth.make_stream(71, K, n_feat=N, n_levels=5, trans=0.012, rot_deg=0.25)
    cam = st["cam"]
    inp, outp = tmp_path / "in.bin", tmp_path / "out.bin"
    with open(inp, "wb") as fh:
        _write_camera(fh, cam)
        fh.write(struct.pack("3i", K, N, mapper_thread))
        fh.write(st["frames"][0][0].tobytes()); fh.write(np.ascontiguousarray(st["poses"][0]).tobytes())
        fh.write(np.ascontiguousarray(st["feats"][0]["px"]).tobytes()); fh.write(np.ascontiguousarray(st["feats"][0]["pos"]).tobytes())
        for k in range(1, K):
            fh.write(st["frames"][k][0].tobytes())
    subprocess.check_call([build_demo("host_pipeline_demo"), "pipeline", str(inp), str(outp)])
    raw = open(outp, "rb").read()
    rec = np.frombuffer(raw, np.float64, 18 * (K - 1), 0).reshape(K - 1, 18)
    tail = np.frombuffer(raw, np.int64, 4, 18 * 8 * (K - 1))
    for k in range(1, K):
        r = rec[k - 1]
        assert int(r[0]) != 2, f"frame {k}: RESULT_FAILURE"
        assert r[1] > 100 and r[2] >= 50 and r[3] >= 50, (k, r[:6])        # tracked patches (<= max_fts+1 after frame 1), matches, pose-opt observations
        dt, dr = synth.pose_error(r[6:].reshape(3, 4), st["poses"][k])
        assert dt < 3e-3 and dr < 2e-3, (k, dt, dr)
        assert r[5] <= r[4] + 1e-9 and r[5] < 1.0                            # reprojection error (px) after <= before
    seeds_left, n_candidates, n_updates, n_kfs = tail
    assert n_kfs >= 1 and n_updates > 100 and seeds_left + n_candidates > 50  # the mapper thread really updated seeds
