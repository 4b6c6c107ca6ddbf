#!/usr/bin/env python
"""Generates the golden fixtures in this directory from the CPU oracle (oracle/) and, in the build container where
/root/reference exists, CHECKS every fixture against oracle/_ref -- the reference's own svo/src sources compiled in place
against stand-in third-party headers (oracle/Makefile target `ref`) -- before writing it (`ref_checked` = 1 in the file).
The reference ships no golden vectors for the hot path (its tests print numbers computed on an external dataset), so
these files are what pins the oracle and the GPU kernels box-independently; what remains unpinned is the arithmetic inside
the un-vendored third-party libraries.

    python tests/golden/make_golden.py          # rewrites *.npz next to this script
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import binding as ob  # noqa: E402
from rpg_svo_b200 import synth  # noqa: E402


def digest(*arrays) -> str:
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def sia_case(seed, n_feat, n_levels, max_level, min_level):
    d = synth.make_frame_pair(seed, n_feat=n_feat, n_levels=n_levels)
    r = ob.sparse_img_align(d["ref_pyr"], d["cur_pyr"], d["cam"], synth.se3_identity(), d["px"], d["f"], d["pos"],
                            d["has_point"], d["ref_pos"], max_level, min_level)
    tr = r["trace"]
    ref_checked = 0
    if ob.ref_lib() is not None:  # the reference's own SparseImgAlign on the same pair
        rr = ob.ref_sparse_img_align(d["ref_pyr"][0], d["cur_pyr"][0], n_levels, d["cam"], d["T_ref_w"], d["T_ref_w"], d["px"],
                                     d["f"], d["pos"], d["has_point"], max_level, min_level)
        assert np.array_equal(rr["visible"], r["visible"]) and rr["n_tracked"] == r["n_tracked"]
        assert np.allclose(rr["T_cur_w"], synth.se3_mul(r["T"], d["T_ref_w"]), rtol=0, atol=1e-9)
        ref_checked = 1
    return dict(seed=seed, ref_checked=ref_checked, n_feat=n_feat, n_levels=n_levels, max_level=max_level, min_level=min_level,
                input_sha256=digest(*d["ref_pyr"], *d["cur_pyr"], d["px"], d["f"], d["pos"], d["has_point"], d["ref_pos"]),
                T=r["T"], visible=r["visible"], n_tracked=r["n_tracked"], H=r["H"],
                trace_level=np.array([t["level"] for t in tr]), trace_iter=np.array([t["iter"] for t in tr]),
                trace_accepted=np.array([t["accepted"] for t in tr]), trace_n_meas=np.array([t["n_meas"] for t in tr]),
                trace_chi2=np.array([t["chi2"] for t in tr]), trace_x=np.array([t["x"] for t in tr]),
                T_gt=d["T_cur_ref_gt"])


def main():
    # C0: "2-frame 640x480, 100 feats, 3 pyramid lvls" and C1: 300 feats, 5 levels
    np.savez_compressed(os.path.join(HERE, "sia_c0.npz"), **sia_case(1000, 100, 3, 2, 0))
    np.savez_compressed(os.path.join(HERE, "sia_c1.npz"), **sia_case(1000, 300, 5, 4, 0))

    # align2D / align1D on a stored 128x96 image
    rng = np.random.default_rng(77)
    cam = synth.camera_for(640, 480)
    img = synth.render(cam, synth.base_pose(), synth.Plane.tilted(), synth.make_texture(7))[200:296, 300:428].copy()
    m = 48
    px_true = np.stack([rng.uniform(12, 116, m), rng.uniform(12, 84, m)], axis=1)
    off = rng.uniform(-1.4, 1.4, (m, 2))
    px_start = px_true - off
    direction = (off / np.linalg.norm(off, axis=1, keepdims=True)).astype(np.float32)
    pwb = np.stack([synth.patch_with_border(img, p).ravel() for p in px_true])
    patch = np.stack([p.reshape(10, 10)[1:9, 1:9].ravel() for p in pwb])
    r2 = [ob.align2d(img, pwb[i], patch[i], 10, px_start[i]) for i in range(m)]
    r1 = [ob.align1d(img, direction[i], pwb[i], patch[i], 10, px_start[i]) for i in range(m)]
    ref_checked = 0
    if ob.ref_lib() is not None:
        for i in range(m):
            ok, p = ob.ref_align2d(img, pwb[i], patch[i], 10, px_start[i])
            assert ok == r2[i][0] and np.array_equal(p, r2[i][1])
            ok, p, h = ob.ref_align1d(img, direction[i], pwb[i], patch[i], 10, px_start[i])
            assert ok == r1[i][0] and np.array_equal(p, r1[i][1]) and h == r1[i][2]
        ref_checked = 1
    np.savez_compressed(os.path.join(HERE, "align.npz"), ref_checked=ref_checked, img=img, px_true=px_true, px_start=px_start, dir=direction,
                        pwb=pwb, patch=patch, conv2d=np.array([r[0] for r in r2]), px2d=np.array([r[1] for r in r2]),
                        conv1d=np.array([r[0] for r in r1]), px1d=np.array([r[1] for r in r1]),
                        h_inv=np.array([r[2] for r in r1]))

    # updateSeed / computeTau known-answer tables
    n = 64
    a, b = rng.uniform(5, 30, n).astype(np.float32), rng.uniform(5, 30, n).astype(np.float32)
    mu = rng.uniform(0.2, 1.0, n).astype(np.float32)
    sigma2 = rng.uniform(1e-3, 0.2, n).astype(np.float32)
    x = (mu + rng.normal(size=n).astype(np.float32) * 0.05).astype(np.float32)
    tau2 = rng.uniform(1e-5, 1e-2, n).astype(np.float32)
    out = np.stack([ob.update_seed(x[i], tau2[i], a[i], b[i], mu[i], 2.0, sigma2[i]) for i in range(n)])
    T = np.stack([ob.se3_exp(np.concatenate([rng.normal(size=3) * 0.3, rng.normal(size=3) * 0.1])) for _ in range(n)])
    f = rng.normal(size=(n, 3)) * 0.2 + [0, 0, 1]
    f /= np.linalg.norm(f, axis=1, keepdims=True)
    z = rng.uniform(0.5, 5, n)
    ang = 2 * np.arctan(1 / (2 * 315.5))
    tau = np.array([ob.compute_tau(T[i], f[i], z[i], ang) for i in range(n)])
    ref_checked = 0
    if ob.ref_lib() is not None:
        for i in range(n):
            assert np.array_equal(ob.ref_update_seed(x[i], tau2[i], a[i], b[i], mu[i], 2.0, sigma2[i]).view(np.uint32), out[i].view(np.uint32))
            assert np.isclose(ob.ref_compute_tau(T[i], f[i], z[i], ang), tau[i], rtol=1e-10)
        ref_checked = 1
    np.savez_compressed(os.path.join(HERE, "depth_kat.npz"), ref_checked=ref_checked, a=a, b=b, mu=mu, sigma2=sigma2, x=x, tau2=tau2, seed_out=out,
                        T=T, f=f, z=z, px_error_angle=ang, tau=tau)

    # pose optimizer, inputs stored in full
    c = synth.make_pose_opt_case(125, 120, 752, 480)
    o = ob.pose_optimize(2.0, 10, c["cam"].fx, c["T_init"], c["f"], c["pos"], c["level"], c["has_point"])
    ref_checked = 0
    if ob.ref_lib() is not None:
        rr = ob.ref_pose_optimize(2.0, 10, c["cam"], c["T_init"], c["f"], c["pos"], c["level"], c["has_point"])
        assert np.array_equal(rr["has_point"], o["has_point"]) and np.allclose(rr["T"], o["T"], rtol=0, atol=1e-10)
        ref_checked = 1
    np.savez_compressed(os.path.join(HERE, "pose_opt.npz"), ref_checked=ref_checked, fx=c["cam"].fx, T_init=c["T_init"], f=c["f"], pos=c["pos"],
                        level=c["level"], has_point=c["has_point"], T=o["T"], has_point_out=o["has_point"],
                        estimated_scale=o["estimated_scale"], error_init=o["error_init"], error_final=o["error_final"],
                        num_obs=o["num_obs"], n_iter_done=o["n_iter_done"], cov=o["cov"])
    # FastDetector::detect on a stored 192x144 image (3 levels built by the oracle's halfSample rule)
    img = synth.render(cam, synth.base_pose(), synth.Plane.tilted(), synth.make_texture(7))[150:294, 250:442].copy()
    pyr = synth.build_pyramid(img, 3)
    occ = (rng.uniform(size=7 * 5) < 0.25).astype(np.uint8)
    det = ob.fast_detect(pyr, 3, 30, 20.0, occ)
    ref_checked = 0
    if ob.ref_lib() is not None:
        rr = ob.ref_fast_detect(img, 3, 3, 30, 20.0, occ)
        assert all(np.array_equal(rr[k], det[k]) for k in ("x", "y", "level"))
        ref_checked = 1
    np.savez_compressed(os.path.join(HERE, "detect.npz"), ref_checked=ref_checked, img=img, occupancy=occ, x=det["x"], y=det["y"],
                        level=det["level"], score=det["score"])

    # Reprojector::reprojectMap on a small synthetic map (inputs regenerated from the seed, digest stored)
    mc = synth.make_map_case(31, n_kfs=5, n_points=300, n_candidates=40)
    ro = ob.reproject_map(mc)
    ref_checked = 0
    if ob.ref_lib() is not None:
        rr = ob.ref_reproject_map(mc)
        for k in ("n_matches", "n_trials", "n_new"):
            assert rr[k] == ro[k]
        for k in ("new_point", "new_level", "new_type", "pt_type", "pt_n_failed", "pt_n_succeeded", "overlap_kf", "overlap_count"):
            assert np.array_equal(rr[k], ro[k]), k
        assert np.array_equal(rr["new_px"], ro["new_px"])
        ref_checked = 1
    v = mc["view"]
    np.savez_compressed(os.path.join(HERE, "reproject.npz"), ref_checked=ref_checked, seed=31,
                        input_sha256=digest(*[p[0] for p in mc["kf_pyr"]], mc["cur_pyr"][0], v["pt_pos"], v["ftr_px"], v["pt_obs"],
                                            mc["cell_order"], mc["pt_type"]),
                        n_matches=ro["n_matches"], n_trials=ro["n_trials"], new_point=ro["new_point"], new_px=ro["new_px"],
                        new_level=ro["new_level"], new_type=ro["new_type"], new_grad=ro["new_grad"], pt_type=ro["pt_type"],
                        pt_n_failed=ro["pt_n_failed"], pt_n_succeeded=ro["pt_n_succeeded"], pt_action=ro["pt_action"],
                        overlap_kf=ro["overlap_kf"], overlap_count=ro["overlap_count"])
    for fn in sorted(os.listdir(HERE)):
        if fn.endswith(".npz"):
            print(fn, os.path.getsize(os.path.join(HERE, fn)), "bytes")


if __name__ == "__main__":
    main()
