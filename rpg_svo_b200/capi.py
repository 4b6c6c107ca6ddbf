"""ctypes binding of libsvo_b200.so -- the thin Python face of the C ABI in include/svo_b200.h.

The product path is the CUDA library.  There is no CPU fallback: `load()` raises if the shared
library has not been built, and `Context()` raises if no CUDA device is usable.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsvo_b200.so")
MAX_LEVELS = 8


class SvoB200Error(RuntimeError):
    pass


class Camera(C.Structure):
    _fields_ = [("fx", C.c_double), ("fy", C.c_double), ("cx", C.c_double), ("cy", C.c_double),
                ("width", C.c_int), ("height", C.c_int), ("model", C.c_int), ("reserved_", C.c_int),
                ("d", C.c_double * 5)]


class SiaOptions(C.Structure):
    _fields_ = [("max_level", C.c_int), ("min_level", C.c_int), ("n_iter", C.c_int), ("eps", C.c_double)]


class SiaIter(C.Structure):
    _fields_ = [("level", C.c_int), ("iter", C.c_int), ("accepted", C.c_int), ("n_meas", C.c_int),
                ("chi2", C.c_double), ("x", C.c_double * 6), ("T", C.c_double * 12)]


class SiaStats(C.Structure):
    _fields_ = [("n_iters", C.c_int32), ("sum_visible", C.c_int32), ("sum_in_image", C.c_int32),
                ("n_tracked", C.c_int32)]


class MatchOptions(C.Structure):
    _fields_ = [("max_search_level", C.c_int), ("align_max_iter", C.c_int)]


class PoseOptResult(C.Structure):
    _fields_ = [("estimated_scale", C.c_double), ("error_init", C.c_double), ("error_final", C.c_double),
                ("num_obs", C.c_int64), ("n_iter_done", C.c_int), ("cov", C.c_double * 36)]


class DepthOptions(C.Structure):
    _fields_ = [("max_n_kfs", C.c_int), ("seed_convergence_sigma2_thresh", C.c_double),
                ("max_search_level", C.c_int), ("align_max_iter", C.c_int), ("max_epi_search_steps", C.c_int)]


SIA_STATS_DTYPE = np.dtype([("n_iters", np.int32), ("sum_visible", np.int32),
                            ("sum_in_image", np.int32), ("n_tracked", np.int32)])

_lib = None


def load() -> C.CDLL:
    """Load the CUDA library; fail loudly if it is missing (no fallback of any kind)."""
    global _lib
    if _lib is None:
        path = os.environ.get("SVO_B200_LIB", LIB_PATH)  # override: instrumented builds of the same library
        if not os.path.exists(path):
            raise SvoB200Error(f"{path} is missing: build it with `python -m rpg_svo_b200.build` "
                               "(there is no CPU fallback)")
        _lib = C.CDLL(path)
        _lib.svo_b200_last_error.restype = C.c_char_p
        _lib.svo_b200_version.restype = C.c_char_p
        _lib.svo_b200_stream.restype = C.c_void_p
        _lib.svo_b200_launch_count.restype = C.c_uint64
        for name in ("svo_b200_last_error", "svo_b200_stream", "svo_b200_launch_count",
                     "svo_b200_synchronize", "svo_b200_destroy"):
            getattr(_lib, name).argtypes = [C.c_void_p]
        _lib.svo_b200_frame_destroy.argtypes = [C.c_void_p, C.c_void_p]
        _lib.svo_b200_frame_destroy.restype = None
        _lib.svo_b200_frame_pool_destroy.argtypes = [C.c_void_p, C.c_void_p]
        _lib.svo_b200_frame_pool_destroy.restype = None
        _lib.svo_b200_frame_pool_get.argtypes = [C.c_void_p, C.c_int]
        _lib.svo_b200_frame_pool_get.restype = C.c_void_p
        _lib.svo_b200_destroy.restype = None
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def c64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


CAM_PINHOLE, CAM_ATAN = 0, 1


def cam_struct(cam) -> Camera:
    """`cam` needs fx, fy, cx, cy, width, height; optional `model` (CAM_*) and `d` (up to 5 coefficients)."""
    d = (C.c_double * 5)(*([float(x) for x in getattr(cam, "d", ())] + [0.0] * 5)[:5])
    return Camera(cam.fx, cam.fy, cam.cx, cam.cy, cam.width, cam.height, int(getattr(cam, "model", 0)), 0, d)


class Frame:
    """An image pyramid resident in HBM (the image side of svo::Frame)."""

    def __init__(self, ctx: "Context", width: int, height: int, n_levels: int):
        self.ctx, self.width, self.height, self.n_levels = ctx, width, height, n_levels
        h = C.c_void_p()
        ctx._check(ctx.lib.svo_b200_frame_create(ctx.h, width, height, n_levels, C.byref(h)))
        self.h = h

    def upload(self, levels) -> "Frame":
        """levels: list of >=1 contiguous uint8 arrays (level 0 first); missing levels are built on the GPU."""
        arr = (C.c_void_p * len(levels))()
        keep = []
        for i, im in enumerate(levels):
            im = np.ascontiguousarray(im, dtype=np.uint8)
            assert im.shape == (self.height >> i, self.width >> i), (im.shape, i)
            keep.append(im)
            arr[i] = im.ctypes.data
        self.ctx._check(self.ctx.lib.svo_b200_frame_upload(self.ctx.h, self.h, arr, len(levels)))
        self.ctx.synchronize()  # host arrays may be freed by the caller right after
        return self

    def upload_ptrs(self, ptrs) -> None:
        """Asynchronous upload from raw (pinned) host pointers; the caller keeps the memory alive."""
        arr = (C.c_void_p * len(ptrs))(*ptrs)
        self.ctx._check(self.ctx.lib.svo_b200_frame_upload(self.ctx.h, self.h, arr, len(ptrs)))

    def upload_device(self, dev_ptr: int) -> None:
        self.ctx._check(self.ctx.lib.svo_b200_frame_upload_device(self.ctx.h, self.h, C.c_void_p(dev_ptr)))

    def download_level(self, level: int) -> np.ndarray:
        out = np.zeros((self.height >> level, self.width >> level), np.uint8)
        self.ctx._check(self.ctx.lib.svo_b200_frame_download_level(self.ctx.h, self.h, level, _p(out)))
        return out

    def destroy(self):
        if self.h and not getattr(self, "borrowed", False):
            self.ctx.lib.svo_b200_frame_destroy(self.ctx.h, self.h)
        self.h = None

    def __del__(self):
        try:
            if self.h and self.ctx.h:
                self.destroy()
        except Exception:
            pass


class FramePool:
    """`count` frames of one geometry in one device slab: one strided H2D copy + one fused pyramid
    kernel per upload (svo_b200_frame_pool_*)."""

    def __init__(self, ctx: "Context", width: int, height: int, n_levels: int, count: int):
        self.ctx, self.width, self.height, self.n_levels, self.count = ctx, width, height, n_levels, count
        h = C.c_void_p()
        ctx._check(ctx.lib.svo_b200_frame_pool_create(ctx.h, width, height, n_levels, count, C.byref(h)))
        self.h = h
        self.frames = []
        for i in range(count):
            f = Frame.__new__(Frame)
            f.ctx, f.width, f.height, f.n_levels = ctx, width, height, n_levels
            f.h = C.c_void_p(ctx.lib.svo_b200_frame_pool_get(self.h, i))
            f.borrowed = True
            self.frames.append(f)

    def upload(self, first: int, count: int, host_ptr: int, host_stride: int) -> None:
        """Asynchronous when `host_ptr` is pinned memory; the caller keeps it alive."""
        self.ctx._check(self.ctx.lib.svo_b200_frame_pool_upload(self.ctx.h, self.h, first, count,
                                                                C.c_void_p(host_ptr), C.c_size_t(host_stride)))

    def upload_array(self, imgs: np.ndarray, first: int = 0) -> None:
        imgs = np.ascontiguousarray(imgs, dtype=np.uint8)
        assert imgs.shape[1:] == (self.height, self.width)
        self.upload(first, imgs.shape[0], imgs.ctypes.data, self.height * self.width)
        self.ctx.synchronize()

    def destroy(self):
        if self.h:
            for f in self.frames:
                f.h = None
            self.ctx.lib.svo_b200_frame_pool_destroy(self.ctx.h, self.h)
            self.h = None


class Context:
    """One GPU + one CUDA stream (use one per calling host thread)."""

    def __init__(self, device: int = 0):
        self.lib = load()
        h = C.c_void_p()
        rc = self.lib.svo_b200_create(C.byref(h), device)
        if rc != 0:
            raise SvoB200Error(f"svo_b200_create(device={device}) failed with {rc}: no usable CUDA device "
                               "(the CUDA path is the only path)")
        self.h = h
        self.device = device

    def close(self):
        if self.h:
            self.lib.svo_b200_destroy(self.h)
            self.h = None

    def _check(self, rc: int):
        if rc != 0:
            raise SvoB200Error(f"svo_b200 error {rc}: {self.lib.svo_b200_last_error(self.h).decode()}")

    @property
    def stream(self) -> int:
        return int(self.lib.svo_b200_stream(self.h))

    def synchronize(self):
        self._check(self.lib.svo_b200_synchronize(self.h))

    def last_kernel_ms(self) -> float:
        """Device time of the kernel(s) of the last entry point (CUDA events inside the library, no copies)."""
        ms = C.c_float(0)
        self._check(self.lib.svo_b200_last_kernel_ms(self.h, C.byref(ms)))
        return float(ms.value)

    def launch_count(self) -> int:
        return int(self.lib.svo_b200_launch_count(self.h))

    def frame(self, pyr) -> Frame:
        """Create + upload a frame from a full host pyramid (list of uint8 arrays)."""
        f = Frame(self, pyr[0].shape[1], pyr[0].shape[0], len(pyr))
        return f.upload(pyr)

    def frame_from_level0(self, img, n_levels: int) -> Frame:
        f = Frame(self, img.shape[1], img.shape[0], n_levels)
        return f.upload([img])

    # ------------------------------------------------------------------ SparseImgAlign
    def sparse_img_align(self, ref: Frame, cur: Frame, cam, T_init, px, f, pos, has_point, ref_pos,
                         max_level, min_level, n_iter=30, eps=1e-6, want_trace=False):
        n = int(np.asarray(px).shape[0])
        T = c64(T_init).copy().reshape(12)
        px, f, pos, rpos = c64(px), c64(f), c64(pos), c64(ref_pos)
        hp = np.ascontiguousarray(has_point, dtype=np.uint8)
        visible = np.zeros(max(n, 1), np.uint8)
        H = np.zeros(36)
        stats = SiaStats()
        cap = ((max_level - min_level + 1) * max(n_iter, 1) + 8) if want_trace else 0
        trace = (SiaIter * cap)() if cap else None
        ntr = C.c_int(0)
        cs = cam_struct(cam)
        opt = SiaOptions(max_level, min_level, n_iter, eps)
        self._check(self.lib.svo_b200_sparse_img_align(
            self.h, ref.h, cur.h, C.byref(cs), C.byref(opt), _p(T), _p(px), _p(f), _p(pos), _p(hp),
            _p(rpos), n, _p(visible), _p(H), C.byref(stats), trace, cap, C.byref(ntr)))
        tr = []
        for k in range(min(ntr.value, cap)):
            r = trace[k]
            tr.append(dict(level=r.level, iter=r.iter, accepted=r.accepted, n_meas=r.n_meas, chi2=r.chi2,
                           x=np.array(r.x[:]), T=np.array(r.T[:]).reshape(3, 4)))
        return dict(T=T.reshape(3, 4), n_tracked=int(stats.n_tracked), visible=visible[:n],
                    H=H.reshape(6, 6), trace=tr,
                    stats=dict(n_iters=stats.n_iters, sum_visible=stats.sum_visible,
                               sum_in_image=stats.sum_in_image, n_tracked=stats.n_tracked))

    def sia_batch_stage(self, refs, curs, cam, T_init, feat_offset, px, f, pos, has_point, ref_pos,
                        max_level, min_level, n_iter=30, eps=1e-6):
        B = len(refs)
        ra = (C.c_void_p * B)(*[r.h.value for r in refs])
        ca = (C.c_void_p * B)(*[c.h.value for c in curs])
        cs = cam_struct(cam)
        opt = SiaOptions(max_level, min_level, n_iter, eps)
        self._batch = dict(B=B, n=int(feat_offset[-1] - feat_offset[0]))
        fo = np.ascontiguousarray(feat_offset, np.int32)
        T, px, f, pos, rpos = c64(T_init), c64(px), c64(f), c64(pos), c64(ref_pos)
        hp = np.ascontiguousarray(has_point, np.uint8)
        self._check(self.lib.svo_b200_sia_batch_stage(self.h, B, ra, ca, C.byref(cs), C.byref(opt), _p(T),
                                                      _p(fo), _p(px), _p(f), _p(pos), _p(hp), _p(rpos)))

    def set_pyramid_rule(self, rule: int):
        """1 = PYR_X86 (vikit's SSE2 rounding where an x86 build of the reference takes it; default), 0 = PYR_SCALAR."""
        self._check(self.lib.svo_b200_set_pyramid_rule(self.h, int(rule)))

    # ---- feature split of single pairs over GPUs (svo_b200_sia_split_*)
    def sia_split_create(self, rank: int, world: int, max_pairs: int = 1):
        """Returns (ipc_handle: bytes[64], device_ptr: int) of this rank's exchange buffer."""
        h = (C.c_ubyte * 64)()
        ptr = C.c_void_p()
        self._check(self.lib.svo_b200_sia_split_create(self.h, int(rank), int(world), int(max_pairs), h, C.byref(ptr)))
        return bytes(h), int(ptr.value)

    def sia_split_connect(self, ipc_handles=None, in_process_ptrs=None):
        """ipc_handles: list of `world` 64-byte handles (other processes) or in_process_ptrs: list of `world` device pointers."""
        if in_process_ptrs is not None:
            arr = (C.c_void_p * len(in_process_ptrs))(*[C.c_void_p(p) for p in in_process_ptrs])
            self._check(self.lib.svo_b200_sia_split_connect(self.h, None, arr))
        else:
            blob = b"".join(ipc_handles)
            buf = (C.c_ubyte * len(blob)).from_buffer_copy(blob)
            self._check(self.lib.svo_b200_sia_split_connect(self.h, buf, None))

    def sia_split_destroy(self):
        self._check(self.lib.svo_b200_sia_split_destroy(self.h))

    def sia_upfront(self, mode=-1):
        """Small-batch cluster geometry: all levels prepared before the first iteration (svo_b200_sia_upfront): -1 auto, 0 off."""
        self._check(self.lib.svo_b200_sia_upfront(self.h, int(mode)))

    def sia_config(self, ctas_per_pair=-1, features_per_thread=0):
        """Launch geometry of the alignment kernel (svo_b200_sia_config): -1 / 0 = automatic."""
        self._check(self.lib.svo_b200_sia_config(self.h, int(ctas_per_pair), int(features_per_thread)))

    def sia_batch_run(self):
        self._check(self.lib.svo_b200_sia_batch_run(self.h))

    def sia_batch_fetch(self, want_H=False):
        B, n = self._batch["B"], self._batch["n"]
        T = np.zeros((B, 3, 4))
        vis = np.zeros(max(n, 1), np.uint8)
        H = np.zeros((B, 6, 6)) if want_H else None
        stats = np.zeros(B, SIA_STATS_DTYPE)
        self._check(self.lib.svo_b200_sia_batch_fetch(self.h, _p(T), _p(vis), _p(H), _p(stats)))
        return dict(T=T, visible=vis[:n], H=H, stats=stats)

    def sparse_residuals(self, ref: Frame, cur: Frame, cam, level, T, px, f, pos, has_point, ref_pos,
                         visible_in=None):
        n = int(np.asarray(px).shape[0])
        vis = np.zeros(n, np.uint8) if visible_in is None else np.ascontiguousarray(visible_in, np.uint8).copy()
        ref_patch = np.zeros((n, 16), np.float32)
        res = np.zeros((n, 16), np.float32)
        inimg = np.zeros(n, np.uint8)
        H, Jres = np.zeros(36), np.zeros(6)
        chi2, nm = C.c_double(0), C.c_int64(0)
        cs = cam_struct(cam)
        T, px, f, pos, rpos = c64(T).reshape(12), c64(px), c64(f), c64(pos), c64(ref_pos)
        hp = np.ascontiguousarray(has_point, np.uint8)
        self._check(self.lib.svo_b200_sparse_residuals(
            self.h, ref.h, cur.h, C.byref(cs), level, _p(T), _p(px), _p(f), _p(pos), _p(hp), _p(rpos), n,
            _p(vis), _p(ref_patch), _p(res), _p(inimg), _p(H), _p(Jres), C.byref(chi2), C.byref(nm)))
        return dict(visible=vis, ref_patch=ref_patch, residuals=res, in_image=inimg, H=H.reshape(6, 6),
                    Jres=Jres, chi2=chi2.value, n_meas=nm.value)


# --------------------------------------------------------------------------------------------
# feature alignment / matcher / pose optimizer / depth filter entry points (methods of Context)
# --------------------------------------------------------------------------------------------
def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _u8(a):
    return np.ascontiguousarray(a, dtype=np.uint8)


def _frame_array(frames):
    return (C.c_void_p * len(frames))(*[f.h.value for f in frames])


def _align2d_batch(self, cur: Frame, level, pwb, patch, n_iter, px):
    """feature_alignment::align2D for M features; returns (converged[M] bool, px[M,2])."""
    level = _i32(level)
    M = len(level)
    px = c64(px).copy().reshape(M, 2)
    conv = np.zeros(max(M, 1), np.uint8)
    pwb, patch = _u8(pwb).reshape(M, 100), _u8(patch).reshape(M, 64)
    self._check(self.lib.svo_b200_align2d_batch(self.h, cur.h, M, _p(level), _p(pwb), _p(patch), int(n_iter),
                                                _p(px), _p(conv)))
    return conv[:M].astype(bool), px


def _align1d_batch(self, cur: Frame, level, direction, pwb, patch, n_iter, px):
    level = _i32(level)
    M = len(level)
    px = c64(px).copy().reshape(M, 2)
    conv = np.zeros(max(M, 1), np.uint8)
    h_inv = np.zeros(max(M, 1))
    d = np.ascontiguousarray(direction, np.float32).reshape(M, 2)
    pwb, patch = _u8(pwb).reshape(M, 100), _u8(patch).reshape(M, 64)
    self._check(self.lib.svo_b200_align1d_batch(self.h, cur.h, M, _p(level), _p(d), _p(pwb), _p(patch), int(n_iter),
                                                _p(px), _p(conv), _p(h_inv)))
    return conv[:M].astype(bool), px, h_inv[:M]


def _find_match_direct(self, ref_frames, ref_T_f_w, cur: Frame, cur_T_f_w, cam, ref_index, ref_px, ref_f, ref_level,
                       ftr_type, ref_grad, point_pos, px_cur, max_search_level, align_max_iter=10):
    M = len(ref_index)
    ra = _frame_array(ref_frames)
    refT = c64(np.asarray(ref_T_f_w)).reshape(-1)
    px = c64(px_cur).copy().reshape(M, 2)
    succ = np.zeros(max(M, 1), np.uint8)
    sl = np.zeros(max(M, 1), np.int32)
    A = np.zeros((max(M, 1), 4))
    hinv = np.zeros(max(M, 1))
    cs = cam_struct(cam)
    opt = MatchOptions(max_search_level, align_max_iter)
    ri, lv, ty = _i32(ref_index), _i32(ref_level), _i32(ftr_type)
    rpx, rf, rg, pp, cT = c64(ref_px), c64(ref_f), c64(ref_grad), c64(point_pos), c64(cur_T_f_w).reshape(12)
    self._check(self.lib.svo_b200_find_match_direct(self.h, ra, _p(refT), len(ref_frames), cur.h, _p(cT), C.byref(cs),
                                                    C.byref(opt), M, _p(ri), _p(rpx), _p(rf), _p(lv), _p(ty), _p(rg),
                                                    _p(pp), _p(px), _p(succ), _p(sl), _p(A), _p(hinv)))
    return dict(success=succ[:M].astype(bool), px_cur=px, search_level=sl[:M], A_cur_ref=A[:M].reshape(M, 2, 2),
                h_inv=hinv[:M])


def _pose_optimize(self, reproj_thresh, n_iter, fx, T_f_w, f, pos, level, has_point):
    """pose_optimizer::optimizeGaussNewton; returns dict like the oracle's."""
    T = c64(T_f_w).copy().reshape(12)
    hp = _u8(has_point).copy()
    out = PoseOptResult()
    lv = _i32(level)
    f, pos = c64(f), c64(pos)
    self._check(self.lib.svo_b200_pose_optimize(self.h, C.c_double(reproj_thresh), int(n_iter), C.c_double(fx), _p(T),
                                                _p(f), _p(pos), _p(lv), _p(hp), len(hp), C.byref(out)))
    return dict(T=T.reshape(3, 4), has_point=hp, estimated_scale=out.estimated_scale, error_init=out.error_init,
                error_final=out.error_final, num_obs=out.num_obs, n_iter_done=out.n_iter_done,
                cov=np.array(out.cov[:]).reshape(6, 6))


def _pose_optimize_batch(self, reproj_thresh, n_iter, fx, T_f_w, obs_offset, f, pos, level, has_point):
    """B frames in one launch (svo_b200_pose_optimize_batch); returns a list of dicts like pose_optimize."""
    off = _i32(obs_offset)
    B = len(off) - 1
    T = c64(T_f_w).copy().reshape(B, 12)
    hp = _u8(has_point).copy()
    out = (PoseOptResult * B)()
    fxa = c64(np.broadcast_to(np.asarray(fx, np.float64), (B,)))
    self._check(self.lib.svo_b200_pose_optimize_batch(self.h, B, C.c_double(reproj_thresh), int(n_iter), _p(fxa), _p(T), _p(off),
                                                      _p(c64(f)), _p(c64(pos)), _p(_i32(level)), _p(hp), out))
    res = []
    for b in range(B):
        o = out[b]
        res.append(dict(T=T[b].reshape(3, 4), has_point=hp[off[b]:off[b + 1]], estimated_scale=o.estimated_scale,
                        error_init=o.error_init, error_final=o.error_final, num_obs=o.num_obs, n_iter_done=o.n_iter_done,
                        cov=np.array(o.cov[:]).reshape(6, 6)))
    return res


def _depth_filter_update(self, ref_frames, ref_T_f_w, cur: Frame, cur_T_f_w, cam, ref_index, ftr_px, ftr_f, ftr_level,
                         ftr_type, ftr_grad, batch_id, batch_counter, seeds, max_n_kfs=3, sigma2_thresh=200.0,
                         max_search_level=2, align_max_iter=10, max_epi_search_steps=1000):
    """DepthFilter::updateSeeds; `seeds` = dict of float32 arrays a,b,mu,z_range,sigma2 (copies are updated)."""
    M = len(ref_index)
    ra = _frame_array(ref_frames)
    refT = c64(np.asarray(ref_T_f_w)).reshape(-1)
    out = {k: np.ascontiguousarray(seeds[k], np.float32).copy() for k in ("a", "b", "mu", "z_range", "sigma2")}
    status = np.zeros(max(M, 1), np.uint8)
    pxc = np.zeros((max(M, 1), 2))
    z = np.zeros(max(M, 1))
    nz = np.zeros(max(M, 1), np.int32)
    cs = cam_struct(cam)
    opt = DepthOptions(max_n_kfs, sigma2_thresh, max_search_level, align_max_iter, max_epi_search_steps)
    ri, fl, ft, bi = _i32(ref_index), _i32(ftr_level), _i32(ftr_type), _i32(batch_id)
    fpx, ff, fg, cT = c64(ftr_px), c64(ftr_f), c64(ftr_grad), c64(cur_T_f_w).reshape(12)
    self._check(self.lib.svo_b200_depth_filter_update(
        self.h, ra, _p(refT), len(ref_frames), cur.h, _p(cT), C.byref(cs), C.byref(opt), M, _p(ri), _p(fpx), _p(ff),
        _p(fl), _p(ft), _p(fg), _p(bi), int(batch_counter), _p(out["a"]), _p(out["b"]), _p(out["mu"]),
        _p(out["z_range"]), _p(out["sigma2"]), _p(status), _p(pxc), _p(z), _p(nz)))
    out.update(status=status[:M], px_cur=pxc[:M], z=z[:M], n_zmssd=nz[:M])
    return out


def _find_epipolar_match_direct(self, ref_frames, ref_T_f_w, cur: Frame, cur_T_f_w, cam, ref_index, ftr_px, ftr_f, ftr_level,
                                ftr_type, ftr_grad, d_est, d_min, d_max, max_search_level=2, align_max_iter=10,
                                max_epi_search_steps=1000):
    """Matcher::findEpipolarMatchDirect for M candidates (svo_b200_find_epipolar_match_direct)."""
    M = len(ref_index)
    ra = _frame_array(ref_frames)
    refT = c64(np.asarray(ref_T_f_w)).reshape(-1)
    succ, rej = np.zeros(max(M, 1), np.uint8), np.zeros(max(M, 1), np.uint8)
    depth, epi = np.zeros(max(M, 1)), np.zeros(max(M, 1))
    pxc, A = np.zeros((max(M, 1), 2)), np.zeros((max(M, 1), 4))
    sl, nz = np.zeros(max(M, 1), np.int32), np.zeros(max(M, 1), np.int32)
    cs = cam_struct(cam)
    opt = DepthOptions(3, 200.0, max_search_level, align_max_iter, max_epi_search_steps)
    self._check(self.lib.svo_b200_find_epipolar_match_direct(
        self.h, ra, _p(refT), len(ref_frames), cur.h, _p(c64(cur_T_f_w).reshape(12)), C.byref(cs), C.byref(opt), M,
        _p(_i32(ref_index)), _p(c64(ftr_px)), _p(c64(ftr_f)), _p(_i32(ftr_level)), _p(_i32(ftr_type)), _p(c64(ftr_grad)),
        _p(c64(d_est)), _p(c64(d_min)), _p(c64(d_max)), _p(succ), _p(depth), _p(pxc), _p(sl), _p(epi), _p(rej), _p(A), _p(nz)))
    return dict(success=succ[:M].astype(bool), depth=depth[:M], px_cur=pxc[:M], search_level=sl[:M], epi_length=epi[:M],
                reject=rej[:M].astype(bool), A_cur_ref=A[:M].reshape(M, 2, 2), n_zmssd=nz[:M])


Context.find_epipolar_match_direct = _find_epipolar_match_direct
Context.align2d_batch = _align2d_batch
Context.align1d_batch = _align1d_batch
Context.find_match_direct = _find_match_direct
Context.pose_optimize = _pose_optimize
Context.pose_optimize_batch = _pose_optimize_batch
Context.depth_filter_update = _depth_filter_update


def _point_optimize_batch(self, n_iter, pos, obs_offset, obs_frame, obs_f, frame_T_f_w):
    """Point::optimize for P points; returns the refined positions [P,3]."""
    p = c64(pos).copy().reshape(-1, 3)
    off, fr = _i32(obs_offset), _i32(obs_frame)
    f, T = c64(obs_f), c64(np.asarray(frame_T_f_w)).reshape(-1)
    self._check(self.lib.svo_b200_point_optimize_batch(self.h, len(p), int(n_iter), _p(off), _p(fr), _p(f), _p(T),
                                                       len(T) // 12, _p(p)))
    return p


Context.point_optimize_batch = _point_optimize_batch


# ---- Reprojector::reprojectMap on a flat map view (svo_b200_reproject_map) ----
class MapView(C.Structure):
    _fields_ = [("n_kfs", C.c_int), ("kf_T_f_w", C.c_void_p), ("kf_keypt_pos", C.c_void_p), ("kf_keypt_valid", C.c_void_p),
                ("kf_fts_offset", C.c_void_p), ("kf_fts", C.c_void_p), ("n_ftrs", C.c_int), ("ftr_kf", C.c_void_p),
                ("ftr_px", C.c_void_p), ("ftr_f", C.c_void_p), ("ftr_level", C.c_void_p), ("ftr_type", C.c_void_p),
                ("ftr_grad", C.c_void_p), ("ftr_point", C.c_void_p), ("n_points", C.c_int), ("pt_pos", C.c_void_p),
                ("pt_obs_offset", C.c_void_p), ("pt_obs", C.c_void_p), ("n_candidates", C.c_int), ("cand_point", C.c_void_p)]


class ReprojectOptions(C.Structure):
    _fields_ = [("grid_size", C.c_int), ("max_fts", C.c_int), ("max_n_kfs", C.c_int), ("find_match_direct", C.c_int),
                ("max_search_level", C.c_int), ("align_max_iter", C.c_int)]


class ReprojectStats(C.Structure):
    _fields_ = [("n_matches", C.c_int64), ("n_trials", C.c_int64), ("n_new", C.c_int), ("n_overlap", C.c_int),
                ("n_projected", C.c_int), ("n_speculative", C.c_int)]


_MV_DTYPES = dict(kf_T_f_w=np.float64, kf_keypt_pos=np.float64, kf_keypt_valid=np.uint8, kf_fts_offset=np.int32,
                  kf_fts=np.int32, ftr_kf=np.int32, ftr_px=np.float64, ftr_f=np.float64, ftr_level=np.int32,
                  ftr_type=np.int32, ftr_grad=np.float64, ftr_point=np.int32, pt_pos=np.float64, pt_obs_offset=np.int32,
                  pt_obs=np.int32, cand_point=np.int32)


def _reproject_map(self, view: dict, kf_frames, cur: Frame, cur_T_f_w, cam, options: dict, cell_order, pt_type, pt_n_failed,
                   pt_n_succeeded):
    """Reprojector::reprojectMap: `view` holds the svo_b200_map_view arrays by field name.  Returns the features the
    reference would add to the frame (new_*), the updated point state, per-point actions and the overlap keyframes."""
    mv, keep = MapView(), []
    for k, v in view.items():
        if k in _MV_DTYPES:
            a = np.ascontiguousarray(v, _MV_DTYPES[k])
            keep.append(a)
            setattr(mv, k, a.ctypes.data)
        else:
            setattr(mv, k, int(v))
    opt = ReprojectOptions(**options)
    P, cap, nk = int(view["n_points"]), int(options["max_fts"]) + 1, int(options["max_n_kfs"])
    o = dict(pt_type=_i32(pt_type).copy(), pt_n_failed=_i32(pt_n_failed).copy(), pt_n_succeeded=_i32(pt_n_succeeded).copy(),
             pt_action=np.zeros(P, np.uint8), overlap_kf=np.full(nk, -1, np.int32), overlap_count=np.zeros(nk, np.int64),
             new_point=np.full(cap, -1, np.int32), new_px=np.zeros((cap, 2)), new_level=np.zeros(cap, np.int32),
             new_type=np.zeros(cap, np.int32), new_grad=np.zeros((cap, 2)))
    st = ReprojectStats()
    fr = _frame_array(kf_frames)
    cs = cam_struct(cam)
    co = _i32(cell_order)
    self._check(self.lib.svo_b200_reproject_map(self.h, C.byref(mv), fr, cur.h, _p(c64(cur_T_f_w).reshape(12)), C.byref(cs),
                                                C.byref(opt), _p(co), _p(o["pt_type"]), _p(o["pt_n_failed"]),
                                                _p(o["pt_n_succeeded"]), _p(o["pt_action"]), _p(o["overlap_kf"]),
                                                _p(o["overlap_count"]), _p(o["new_point"]), _p(o["new_px"]),
                                                _p(o["new_level"]), _p(o["new_type"]), _p(o["new_grad"]), C.byref(st)))
    n, k = st.n_new, st.n_overlap
    for key in ("new_point", "new_px", "new_level", "new_type", "new_grad"):
        o[key] = o[key][:n]
    o["overlap_kf"], o["overlap_count"] = o["overlap_kf"][:k], o["overlap_count"][:k]
    o.update(n_matches=st.n_matches, n_trials=st.n_trials, n_new=n, n_overlap=k, n_projected=st.n_projected,
             n_speculative=st.n_speculative)
    return o


Context.reproject_map = _reproject_map


class DetectOptions(C.Structure):
    _fields_ = [("cell_size", C.c_int), ("n_pyr_levels", C.c_int), ("fast_threshold", C.c_int),
                ("nonmax_ties_suppress", C.c_int), ("detection_threshold", C.c_double)]


def _fast_detect(self, frame: Frame, cell_size, n_pyr_levels, detection_threshold, grid_occupancy=None, fast_threshold=20,
                 nonmax_ties_suppress=0, cap=8192):
    """FastDetector::detect: dict(x, y, level, score) of the best corner per free grid cell, cell order."""
    opt = DetectOptions(int(cell_size), int(n_pyr_levels), int(fast_threshold), int(nonmax_ties_suppress), float(detection_threshold))
    x, y, lv = np.zeros(cap, np.int32), np.zeros(cap, np.int32), np.zeros(cap, np.int32)
    sc = np.zeros(cap, np.float32)
    occ = None if grid_occupancy is None else _u8(grid_occupancy)
    n = C.c_int(0)
    self._check(self.lib.svo_b200_fast_detect(self.h, frame.h, C.byref(opt), _p(occ) if occ is not None else None, cap, _p(x),
                                              _p(y), _p(lv), _p(sc), C.byref(n)))
    k = min(n.value, cap)
    return dict(x=x[:k], y=y[:k], level=lv[:k], score=sc[:k], n=n.value)


Context.fast_detect = _fast_detect
