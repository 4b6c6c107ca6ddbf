#!/usr/bin/env python
"""Generates the golden fixtures in this directory FROM THE CPU ORACLE (oracle/), because the reference
itself cannot be built or imported in this environment (C++ with Eigen/OpenCV/Sophus/vikit/Boost, none
present) and ships no golden vectors for the hot path (its tests print numbers computed on an external
dataset).  The fixtures therefore pin the oracle against regressions and give the GPU tests
size-independent, box-independent expected values; they are NOT outputs of the reference binary.

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
    return dict(seed=seed, n_feat=n_feat, n_levels=n_levels, max_level=max_level, min_level=min_level,
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
    np.savez_compressed(os.path.join(HERE, "align.npz"), img=img, px_true=px_true, px_start=px_start, dir=direction,
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
    np.savez_compressed(os.path.join(HERE, "depth_kat.npz"), a=a, b=b, mu=mu, sigma2=sigma2, x=x, tau2=tau2, seed_out=out,
                        T=T, f=f, z=z, px_error_angle=ang, tau=tau)

    # pose optimizer, inputs stored in full
    c = synth.make_pose_opt_case(125, 120, 752, 480)
    o = ob.pose_optimize(2.0, 10, c["cam"].fx, c["T_init"], c["f"], c["pos"], c["level"], c["has_point"])
    np.savez_compressed(os.path.join(HERE, "pose_opt.npz"), fx=c["cam"].fx, T_init=c["T_init"], f=c["f"], pos=c["pos"],
                        level=c["level"], has_point=c["has_point"], T=o["T"], has_point_out=o["has_point"],
                        estimated_scale=o["estimated_scale"], error_init=o["error_init"], error_final=o["error_final"],
                        num_obs=o["num_obs"], n_iter_done=o["n_iter_done"], cov=o["cov"])
    for fn in sorted(os.listdir(HERE)):
        if fn.endswith(".npz"):
            print(fn, os.path.getsize(os.path.join(HERE, fn)), "bytes")


if __name__ == "__main__":
    main()
