// oracle/svo_oracle.cpp -- TEST INFRASTRUCTURE ONLY (see svo_oracle.h).
//
// CPU restatement of the rpg_svo direct-tracking hot path.  Every function cites the reference
// file:line it follows (paths relative to /root/reference).  Per-quantity precision is the
// reference's (SURVEY.md 8a "Precision summary"): u8 images, f32 interpolation / residuals /
// chi2, f64 geometry / Jacobians / normal equations.
//
// Floating-point contraction: the reference is built with -O3 -march=native and GCC's default
// -ffp-contract=fast (svo/CMakeLists.txt:34-45), i.e. a*b+c is fused where the multiply feeds one
// add.  This file is compiled with -ffp-contract=off and spells the fusions out with fma()/fmaf()
// in the canonical "first product fuses into the add" pattern, so the CUDA kernels (compiled with
// -fmad=false and the same explicit fma calls) can be compared bit-for-bit on the f32 stages.
//
// PARITY UNPINNED at the vikit/Sophus/Eigen boundary ([EXT] tags): see svo_oracle.h.
#include "svo_oracle.h"

#include <algorithm>
#include <atomic>
#include <thread>
#include <cmath>
#include <cstdio>
#include <list>
#include <vector>

#include "oracle_math.h"

using namespace orc;

namespace {

// ------------------------------------------------------------------------------------------
// [EXT] vk::AbstractCamera models (rpg_vikit, not vendored, no version pinned), restated from the published
// pinhole_camera.cpp / atan_camera.cpp:
//   model 0  vk::PinholeCamera(width, height, fx, fy, cx, cy, d0..d4): radial-tangential distortion when
//            fabs(d0) > 1e-7; cam2world then goes through cv::undistortPoints on ONE CV_32FC2 point [EXT OpenCV]
//            (float in, 5 fixed-point iterations in double, float out);
//   model 1  vk::ATANCamera (PTAM's FOV model), d0 = s; fx, fy, cx, cy are the pixel values its constructor derives.
// Explicit fma() spells the contraction GCC applies to the reference build (px = fx*u + cx).
// ------------------------------------------------------------------------------------------
struct Cam {
  double fx, fy, cx, cy;
  int width, height;
  int model;
  double d[5];
  bool distorted;
  double s_inv, tans, tans_inv;
  // world2cam(uv on unit plane)
  V2 world2cam_uv(V2 uv) const {
    const double x = uv.x, y = uv.y;
    if (!distorted) return {std::fma(fx, x, cx), std::fma(fy, y, cy)};
    if (model == 0) {
      const double r2 = std::fma(x, x, y * y), r4 = r2 * r2, r6 = r4 * r2;
      const double a1 = 2.0 * x * y, a2 = std::fma(2.0 * x, x, r2), a3 = std::fma(2.0 * y, y, r2);
      const double cdist = std::fma(d[4], r6, std::fma(d[1], r4, std::fma(d[0], r2, 1.0)));
      const double xd = std::fma(d[3], a2, std::fma(d[2], a1, x * cdist));
      const double yd = std::fma(d[3], a1, std::fma(d[2], a3, y * cdist));
      return {std::fma(xd, fx, cx), std::fma(yd, fy, cy)};
    }
    const double r = std::sqrt(std::fma(x, x, y * y));
    const double factor = r < 0.001 ? 1.0 : s_inv * std::atan(r * tans) / r;  // rtrans_factor
    return {std::fma(fx * factor, x, cx), std::fma(fy * factor, y, cy)};
  }
  // world2cam(xyz) = world2cam(project2d(xyz)), project2d = xyz.head2 / z
  V2 world2cam(V3 p) const { return world2cam_uv(V2{p.x / p.z, p.y / p.z}); }
  // cam2world(u,v): normalised bearing vector
  V3 cam2world(double u, double v) const {
    double x, y;
    if (model == 0) {
      if (!distorted) {
        x = (u - cx) / fx;
        y = (v - cy) / fy;
      } else {
        const double uf = double(float(u)), vf = double(float(v));
        const double ifx = 1.0 / fx, ify = 1.0 / fy;
        const double x0 = (uf - cx) * ifx, y0 = (vf - cy) * ify;
        x = x0; y = y0;
        for (int j = 0; j < 5; ++j) {
          const double r2 = x * x + y * y;
          const double icdist = 1.0 / (1.0 + ((d[4] * r2 + d[1]) * r2 + d[0]) * r2);
          const double dX = 2.0 * d[2] * x * y + d[3] * (r2 + 2.0 * x * x);
          const double dY = d[2] * (r2 + 2.0 * y * y) + 2.0 * d[3] * x * y;
          x = (x0 - dX) * icdist;
          y = (y0 - dY) * icdist;
        }
        x = double(float(x));
        y = double(float(y));
      }
    } else {
      const double ifx = 1.0 / fx, ify = 1.0 / fy;
      const double dx = (u - cx) * ifx, dy = (v - cy) * ify;
      const double dist_r = std::sqrt(dx * dx + dy * dy);
      const double r = distorted ? std::tan(dist_r * d[0]) * tans_inv : dist_r;  // invrtrans
      const double d_factor = dist_r > 0.01 ? r / dist_r : 1.0;
      x = d_factor * dx;
      y = d_factor * dy;
    }
    return normalized(V3{x, y, 1.0});
  }
  double errorMultiplier2() const { return std::fabs(fx); }
  bool isInFrame(int x, int y, int boundary) const {
    return x >= boundary && x < width - boundary && y >= boundary && y < height - boundary;
  }
  bool isInFrame(int x, int y, int boundary, int level) const {
    return x >= boundary && x < width / (1 << level) - boundary && y >= boundary &&
           y < height / (1 << level) - boundary;
  }
};
inline Cam make_cam(const orc_camera* c) {
  Cam k{};
  k.fx = c->fx; k.fy = c->fy; k.cx = c->cx; k.cy = c->cy; k.width = c->width; k.height = c->height;
  k.model = c->model;
  for (int i = 0; i < 5; ++i) k.d[i] = c->d[i];
  if (c->model == 0) {
    k.distorted = std::fabs(c->d[0]) > 0.0000001;
  } else if (c->d[0] != 0.0) {
    k.tans = 2.0 * std::tan(c->d[0] / 2.0);
    k.tans_inv = 1.0 / k.tans;
    k.s_inv = 1.0 / c->d[0];
    k.distorted = true;
  }
  return k;
}

struct Img {
  const uint8_t* data;
  int cols, rows, step;
};

// Bilinear blend as GCC contracts `wtl*a + wtr*b + wbl*c + wbr*d` (left-assoc sums, the first
// product of each add fused): fma(wbr,d, fma(wbl,c, fma(wtl,a, wtr*b))).
inline float bilin(float wtl, float wtr, float wbl, float wbr, float a, float b, float c,
                   float d) {
  return std::fmaf(wbr, d, std::fmaf(wbl, c, std::fmaf(wtl, a, wtr * b)));
}

// svo/include/svo/frame.h:116-138  Frame::jacobian_xyz2uv
inline void jacobian_xyz2uv(V3 p, double J[2][6]) {
  const double x = p.x, y = p.y;
  const double z_inv = 1. / p.z;
  const double z_inv_2 = z_inv * z_inv;
  J[0][0] = -z_inv;
  J[0][1] = 0.0;
  J[0][2] = x * z_inv_2;
  J[0][3] = y * J[0][2];
  J[0][4] = -(1.0 + x * J[0][2]);
  J[0][5] = y * z_inv;
  J[1][0] = 0.0;
  J[1][1] = -z_inv;
  J[1][2] = y * z_inv_2;
  J[1][3] = 1.0 + y * J[1][2];
  J[1][4] = -J[0][3];
  J[1][5] = -x * z_inv;
}

inline double norm_max6(const double x[6]) {  // [EXT] vk::norm_max
  double m = 0;
  for (int i = 0; i < 6; ++i) m = std::max(m, std::fabs(x[i]));
  return m;
}

// ------------------------------------------------------------------------------------------
// svo::SparseImgAlign (svo/src/sparse_img_align.cpp, svo/include/svo/sparse_img_align.h)
// on top of [EXT] vk::NLLSSolver<6,SE3> (GN driver restated in optimizeGaussNewton below).
// ------------------------------------------------------------------------------------------
struct SparseImgAlign {
  static const int patch_halfsize_ = 2;               // sparse_img_align.h:35
  static const int patch_size_ = 2 * patch_halfsize_;  // :36
  static const int patch_area_ = patch_size_ * patch_size_;  // :37

  // inputs (the Frame/Feature/Point graph flattened)
  const Img* ref_pyr;
  const Img* cur_pyr;
  Cam cam;
  int N;
  const double* px;
  const double* f;
  const double* pos;
  const uint8_t* has_point;
  V3 ref_pos;

  // NLLSSolver state [EXT]
  double H_[6][6];
  double Jres_[6];
  double x_[6];
  double chi2_;
  size_t n_meas_;
  size_t n_iter_, n_iter_init_;
  size_t iter_;
  bool stop_;
  double eps_;

  // SparseImgAlign state
  int level_, max_level_, min_level_;
  std::vector<float> ref_patch_cache_;   // N x 16
  std::vector<double> jacobian_cache_;   // 6 x (16 N), column-major
  std::vector<uint8_t> visible_fts_;
  bool have_ref_patch_cache_;

  // instrumentation
  std::vector<float> last_res_;       // N x 16, NaN where not evaluated in the last pass
  std::vector<uint8_t> last_in_img_;  // N
  orc_sia_iter* trace = nullptr;
  int trace_cap = 0, n_trace = 0;

  // sparse_img_align.cpp:29-41 (ctor) -- eps_ = 0.000001
  SparseImgAlign(int max_level, int min_level, int n_iter, double eps)
      : max_level_(max_level), min_level_(min_level) {
    n_iter_ = n_iter;
    n_iter_init_ = n_iter_;
    eps_ = eps;
    reset();
  }

  // [EXT] NLLSSolver::reset
  void reset() {
    chi2_ = 1e10;
    n_meas_ = 0;
    n_iter_ = n_iter_init_;
    iter_ = 0;
    stop_ = false;
  }

  // sparse_img_align.cpp:84-145
  void precomputeReferencePatches() {
    const int border = patch_halfsize_ + 1;
    const Img& ref_img = ref_pyr[level_];
    const int stride = ref_img.cols;
    const float scale = 1.0f / (1 << level_);
    const double focal_length = cam.errorMultiplier2();
    for (int i = 0; i < N; ++i) {
      const float u_ref = px[2 * i] * scale;  // f64 * f32 -> f64 -> f32
      const float v_ref = px[2 * i + 1] * scale;
      const int u_ref_i = floorf(u_ref);
      const int v_ref_i = floorf(v_ref);
      if (!has_point[i] || u_ref_i - border < 0 || v_ref_i - border < 0 ||
          u_ref_i + border >= ref_img.cols || v_ref_i + border >= ref_img.rows)
        continue;
      visible_fts_[i] = 1;

      const V3 p{pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]};
      const double depth = norm(p - ref_pos);
      const V3 xyz_ref = V3{f[3 * i], f[3 * i + 1], f[3 * i + 2]} * depth;

      double frame_jac[2][6];
      jacobian_xyz2uv(xyz_ref, frame_jac);

      const float subpix_u_ref = u_ref - u_ref_i;
      const float subpix_v_ref = v_ref - v_ref_i;
      const float w_ref_tl = (1.0 - subpix_u_ref) * (1.0 - subpix_v_ref);
      const float w_ref_tr = subpix_u_ref * (1.0 - subpix_v_ref);
      const float w_ref_bl = (1.0 - subpix_u_ref) * subpix_v_ref;
      const float w_ref_br = subpix_u_ref * subpix_v_ref;
      size_t pixel_counter = 0;
      float* cache_ptr = ref_patch_cache_.data() + patch_area_ * i;
      const double jscale = focal_length / (1 << level_);
      for (int y = 0; y < patch_size_; ++y) {
        const uint8_t* p0 =
            ref_img.data + (v_ref_i + y - patch_halfsize_) * stride + (u_ref_i - patch_halfsize_);
        for (int x = 0; x < patch_size_; ++x, ++p0, ++cache_ptr, ++pixel_counter) {
          *cache_ptr = bilin(w_ref_tl, w_ref_tr, w_ref_bl, w_ref_br, p0[0], p0[1], p0[stride],
                             p0[stride + 1]);
          const float dx =
              0.5f * (bilin(w_ref_tl, w_ref_tr, w_ref_bl, w_ref_br, p0[1], p0[2], p0[stride + 1],
                            p0[stride + 2]) -
                      bilin(w_ref_tl, w_ref_tr, w_ref_bl, w_ref_br, p0[-1], p0[0], p0[stride - 1],
                            p0[stride]));
          const float dy =
              0.5f * (bilin(w_ref_tl, w_ref_tr, w_ref_bl, w_ref_br, p0[stride], p0[1 + stride],
                            p0[stride * 2], p0[stride * 2 + 1]) -
                      bilin(w_ref_tl, w_ref_tr, w_ref_bl, w_ref_br, p0[-stride], p0[1 - stride],
                            p0[0], p0[1]));
          double* col = jacobian_cache_.data() + 6 * (size_t(i) * patch_area_ + pixel_counter);
          for (int k = 0; k < 6; ++k)
            col[k] = std::fma(double(dx), frame_jac[0][k], double(dy) * frame_jac[1][k]) * jscale;
        }
      }
    }
    have_ref_patch_cache_ = true;
  }

  // sparse_img_align.cpp:147-243 (use_weights_ == false, display_ == false: the defaults)
  double computeResiduals(const SE3& T_cur_from_ref, bool linearize_system) {
    const Img& cur_img = cur_pyr[level_];
    if (!have_ref_patch_cache_) precomputeReferencePatches();
    const int stride = cur_img.cols;
    const int border = patch_halfsize_ + 1;
    const float scale = 1.0f / (1 << level_);
    float chi2 = 0.0;
    std::fill(last_res_.begin(), last_res_.end(), std::numeric_limits<float>::quiet_NaN());
    std::fill(last_in_img_.begin(), last_in_img_.end(), 0);
    for (int i = 0; i < N; ++i) {
      if (!visible_fts_[i]) continue;
      const V3 p{pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]};
      const double depth = norm(p - ref_pos);
      const V3 xyz_ref = V3{f[3 * i], f[3 * i + 1], f[3 * i + 2]} * depth;
      const V3 xyz_cur = T_cur_from_ref * xyz_ref;
      const V2 uvd = cam.world2cam(xyz_cur);
      const float u_cur = float(uvd.x) * scale;  // .cast<float>() * scale
      const float v_cur = float(uvd.y) * scale;
      const int u_cur_i = floorf(u_cur);
      const int v_cur_i = floorf(v_cur);
      if (u_cur_i < 0 || v_cur_i < 0 || u_cur_i - border < 0 || v_cur_i - border < 0 ||
          u_cur_i + border >= cur_img.cols || v_cur_i + border >= cur_img.rows)
        continue;
      last_in_img_[i] = 1;

      const float subpix_u_cur = u_cur - u_cur_i;
      const float subpix_v_cur = v_cur - v_cur_i;
      const float w_cur_tl = (1.0 - subpix_u_cur) * (1.0 - subpix_v_cur);
      const float w_cur_tr = subpix_u_cur * (1.0 - subpix_v_cur);
      const float w_cur_bl = (1.0 - subpix_u_cur) * subpix_v_cur;
      const float w_cur_br = subpix_u_cur * subpix_v_cur;
      const float* ref_patch_cache_ptr = ref_patch_cache_.data() + patch_area_ * i;
      size_t pixel_counter = 0;
      for (int y = 0; y < patch_size_; ++y) {
        const uint8_t* c0 =
            cur_img.data + (v_cur_i + y - patch_halfsize_) * stride + (u_cur_i - patch_halfsize_);
        for (int x = 0; x < patch_size_; ++x, ++pixel_counter, ++c0, ++ref_patch_cache_ptr) {
          const float intensity_cur =
              bilin(w_cur_tl, w_cur_tr, w_cur_bl, w_cur_br, c0[0], c0[1], c0[stride], c0[stride + 1]);
          const float res = intensity_cur - (*ref_patch_cache_ptr);
          last_res_[size_t(i) * patch_area_ + pixel_counter] = res;
          const float weight = 1.0;
          chi2 = std::fmaf(res * res, weight, chi2);  // chi2 += res*res*weight
          n_meas_++;
          if (linearize_system) {
            const double* J = jacobian_cache_.data() + 6 * (size_t(i) * patch_area_ + pixel_counter);
            const double w = weight, r = res;
            for (int a = 0; a < 6; ++a) {
              for (int b = 0; b < 6; ++b) H_[a][b] += J[a] * J[b] * w;
              Jres_[a] -= J[a] * r * w;
            }
          }
        }
      }
    }
    return chi2 / n_meas_;  // float / size_t -> float; NaN when n_meas_ == 0
  }

  // sparse_img_align.cpp:245-251
  int solve() {
    LDLT<6> ldlt;
    ldlt.compute(H_);
    ldlt.solve(Jres_, x_);
    if (std::isnan(x_[0])) return 0;
    return 1;
  }

  // sparse_img_align.cpp:253-258
  void update(const SE3& T_curold_from_ref, SE3& T_curnew_from_ref) {
    double mx[6];
    for (int i = 0; i < 6; ++i) mx[i] = -x_[i];
    T_curnew_from_ref = T_curold_from_ref * se3_exp(mx);
  }

  // [EXT] vk::NLLSSolver<6,SE3>::optimizeGaussNewton (rpg_vikit nlls_solver_impl.hpp),
  // restated from the published source: see SURVEY.md row a6 for the control flow.
  void optimizeGaussNewton(SE3& model) {
    SE3 old_model = model;
    for (iter_ = 0; iter_ < n_iter_; ++iter_) {
      std::memset(H_, 0, sizeof(H_));
      std::memset(Jres_, 0, sizeof(Jres_));
      n_meas_ = 0;
      const double new_chi2 = computeResiduals(model, true);
      if (!solve()) stop_ = true;
      const bool reject = (iter_ > 0 && new_chi2 > chi2_) || stop_;
      if (reject) {
        model = old_model;  // rollback
        record(new_chi2, 0, model);
        break;
      }
      SE3 new_model;
      update(model, new_model);
      old_model = model;
      model = new_model;
      chi2_ = new_chi2;
      record(new_chi2, 1, model);
      if (norm_max6(x_) <= eps_) break;
    }
  }

  void record(double chi2, int accepted, const SE3& model) {
    if (trace && n_trace < trace_cap) {
      orc_sia_iter& r = trace[n_trace];
      r.level = level_;
      r.iter = int(iter_);
      r.accepted = accepted;
      r.n_meas = int(n_meas_);
      r.chi2 = chi2;
      for (int i = 0; i < 6; ++i) r.x[i] = x_[i];
      se3_to_rt12(model, r.T);
    }
    ++n_trace;
  }

  // sparse_img_align.cpp:43-75 with T_cur_from_ref passed in/out instead of the two T_f_w_
  size_t run(SE3& T_cur_from_ref) {
    reset();
    if (N == 0) return 0;  // :47-51 "no features to track"
    ref_patch_cache_.assign(size_t(N) * patch_area_, 0.f);
    jacobian_cache_.assign(size_t(N) * patch_area_ * 6, 0.0);
    visible_fts_.assign(N, 0);  // resized once, never cleared per level (:57)
    last_res_.assign(size_t(N) * patch_area_, 0.f);
    last_in_img_.assign(N, 0);
    for (level_ = max_level_; level_ >= min_level_; --level_) {
      std::fill(jacobian_cache_.begin(), jacobian_cache_.end(), 0.0);
      have_ref_patch_cache_ = false;
      optimizeGaussNewton(T_cur_from_ref);
    }
    return n_meas_ / patch_area_;
  }
};

}  // namespace

// ==========================================================================================
// C interface
// ==========================================================================================
extern "C" {

int64_t orc_sparse_img_align_run(const uint8_t* const* ref_levels,
                                 const uint8_t* const* cur_levels, const int* cols,
                                 const int* rows, int n_levels, const orc_camera* cam,
                                 double* T_io, const double* px, const double* f,
                                 const double* point_pos, const uint8_t* has_point,
                                 const double* ref_pos, int N, int max_level, int min_level,
                                 int n_iter, double eps, uint8_t* visible_out, double* H_out,
                                 float* residuals_out, orc_sia_iter* trace, int trace_cap,
                                 int* n_trace) {
  Img rp[ORC_MAX_LEVELS], cp[ORC_MAX_LEVELS];
  for (int l = 0; l < n_levels && l < ORC_MAX_LEVELS; ++l) {
    rp[l] = Img{ref_levels[l], cols[l], rows[l], cols[l]};
    cp[l] = Img{cur_levels[l], cols[l], rows[l], cols[l]};
  }
  SparseImgAlign sia(max_level, min_level, n_iter, eps);
  sia.ref_pyr = rp;
  sia.cur_pyr = cp;
  sia.cam = make_cam(cam);
  sia.N = N;
  sia.px = px;
  sia.f = f;
  sia.pos = point_pos;
  sia.has_point = has_point;
  sia.ref_pos = V3{ref_pos[0], ref_pos[1], ref_pos[2]};
  sia.trace = trace;
  sia.trace_cap = trace_cap;
  std::memset(sia.H_, 0, sizeof(sia.H_));
  SE3 T = se3_from_rt12(T_io);
  const size_t ret = sia.run(T);
  se3_to_rt12(T, T_io);
  if (visible_out)
    for (int i = 0; i < N; ++i) visible_out[i] = N ? sia.visible_fts_[i] : 0;
  if (H_out)
    for (int a = 0; a < 6; ++a)
      for (int b = 0; b < 6; ++b) H_out[a * 6 + b] = sia.H_[a][b];
  if (residuals_out && N)
    std::memcpy(residuals_out, sia.last_res_.data(), sizeof(float) * size_t(N) * 16);
  if (n_trace) *n_trace = sia.n_trace;
  return int64_t(ret);
}

void orc_sparse_img_align_batch(int B, const uint8_t* const* ref_levels,
                                const uint8_t* const* cur_levels, const int* cols, const int* rows,
                                int n_levels, const orc_camera* cam, double* T_io,
                                const int* feat_offset, const double* px, const double* f,
                                const double* point_pos, const uint8_t* has_point,
                                const double* ref_pos, int max_level, int min_level, int n_iter,
                                double eps, int64_t* n_tracked_out, int n_threads) {
  std::atomic<int> next(0);
  auto worker = [&]() {
    // one solver object per worker: its caches are allocated once and reused for every pair this
    // thread processes (a fresh 230 KB Jacobian cache per pair would go through mmap/munmap and
    // serialise the threads in the kernel).
    SparseImgAlign sia(max_level, min_level, n_iter, eps);
    sia.cam = make_cam(cam);
    Img rp[ORC_MAX_LEVELS], cp[ORC_MAX_LEVELS];
    for (;;) {
      const int b = next.fetch_add(1);
      if (b >= B) break;
      const int o = feat_offset[b], n = feat_offset[b + 1] - o;
      for (int l = 0; l < n_levels && l < ORC_MAX_LEVELS; ++l) {
        rp[l] = Img{ref_levels[size_t(b) * n_levels + l], cols[l], rows[l], cols[l]};
        cp[l] = Img{cur_levels[size_t(b) * n_levels + l], cols[l], rows[l], cols[l]};
      }
      sia.ref_pyr = rp;
      sia.cur_pyr = cp;
      sia.N = n;
      sia.px = px + 2 * size_t(o);
      sia.f = f + 3 * size_t(o);
      sia.pos = point_pos + 3 * size_t(o);
      sia.has_point = has_point + o;
      sia.ref_pos = V3{ref_pos[3 * size_t(b)], ref_pos[3 * size_t(b) + 1], ref_pos[3 * size_t(b) + 2]};
      sia.n_trace = 0;
      SE3 T = se3_from_rt12(T_io + 12 * size_t(b));
      const size_t r = sia.run(T);
      se3_to_rt12(T, T_io + 12 * size_t(b));
      if (n_tracked_out) n_tracked_out[b] = int64_t(r);
    }
  };
  if (n_threads <= 1) {
    worker();
    return;
  }
  std::vector<std::thread> pool;
  for (int t = 0; t < n_threads; ++t) pool.emplace_back(worker);
  for (auto& th : pool) th.join();
}

int orc_sparse_residuals(const uint8_t* ref_img, const uint8_t* cur_img, int cols, int rows,
                         int level, const orc_camera* cam, const double* T_cur_from_ref,
                         const double* px, const double* f, const double* point_pos,
                         const uint8_t* has_point, const double* ref_pos, int N,
                         uint8_t* visible_io, float* ref_patch_out, double* jac_out,
                         float* residuals_out, uint8_t* in_image_out, double* H_out,
                         double* Jres_out, double* chi2_out, int64_t* n_meas_out) {
  Img rp[ORC_MAX_LEVELS], cp[ORC_MAX_LEVELS];
  for (int l = 0; l < ORC_MAX_LEVELS; ++l) rp[l] = cp[l] = Img{nullptr, 0, 0, 0};
  rp[level] = Img{ref_img, cols, rows, cols};
  cp[level] = Img{cur_img, cols, rows, cols};
  SparseImgAlign sia(level, level, 1, 1e-6);
  sia.ref_pyr = rp;
  sia.cur_pyr = cp;
  sia.cam = make_cam(cam);
  sia.N = N;
  sia.px = px;
  sia.f = f;
  sia.pos = point_pos;
  sia.has_point = has_point;
  sia.ref_pos = V3{ref_pos[0], ref_pos[1], ref_pos[2]};
  sia.ref_patch_cache_.assign(size_t(N) * 16, 0.f);
  sia.jacobian_cache_.assign(size_t(N) * 16 * 6, 0.0);
  sia.visible_fts_.assign(visible_io, visible_io + N);
  sia.last_res_.assign(size_t(N) * 16, 0.f);
  sia.last_in_img_.assign(N, 0);
  sia.level_ = level;
  sia.have_ref_patch_cache_ = false;
  std::memset(sia.H_, 0, sizeof(sia.H_));
  std::memset(sia.Jres_, 0, sizeof(sia.Jres_));
  sia.n_meas_ = 0;
  const SE3 T = se3_from_rt12(T_cur_from_ref);
  const double chi2 = sia.computeResiduals(T, true);
  for (int i = 0; i < N; ++i) visible_io[i] = sia.visible_fts_[i];
  if (ref_patch_out) std::memcpy(ref_patch_out, sia.ref_patch_cache_.data(), sizeof(float) * N * 16);
  if (jac_out) std::memcpy(jac_out, sia.jacobian_cache_.data(), sizeof(double) * N * 96);
  if (residuals_out) std::memcpy(residuals_out, sia.last_res_.data(), sizeof(float) * N * 16);
  if (in_image_out) std::memcpy(in_image_out, sia.last_in_img_.data(), N);
  if (H_out)
    for (int a = 0; a < 6; ++a)
      for (int b = 0; b < 6; ++b) H_out[a * 6 + b] = sia.H_[a][b];
  if (Jres_out)
    for (int a = 0; a < 6; ++a) Jres_out[a] = sia.Jres_[a];
  if (chi2_out) *chi2_out = chi2;
  if (n_meas_out) *n_meas_out = int64_t(sia.n_meas_);
  return 0;
}

// [EXT] vk::halfSample (rpg_vikit vision.cpp); called from svo/src/frame.cpp:156-165 (createImgPyramid; rows/2, cols/2).
//   rule 0 (SCALAR): (tl + tr + bl + br) / 4, integer division -- the non-SIMD branch.
//   rule 1 (X86):    what an x86 build of the reference computes: when in_cols % 16 == 0 (the buffers are 16-byte
//                    aligned cv::Mat allocations) the SSE2 branch runs -- vertical rounded average (a+c+1)>>1 of
//                    _mm_avg_epu8, then rounded average of horizontally adjacent results (_mm_avg_epu16) -- else rule 0.
void orc_half_sample_rule(const uint8_t* in, int in_cols, int in_rows, uint8_t* out, int rule) {
  const int oc = in_cols / 2, orows = in_rows / 2;
  const bool sse2 = rule == 1 && (in_cols % 16) == 0;
  for (int y = 0; y < orows; ++y) {
    const uint8_t* top = in + size_t(2 * y) * in_cols;
    const uint8_t* bot = top + in_cols;
    for (int x = 0; x < oc; ++x) {
      if (sse2) {
        const unsigned v0 = (unsigned(top[2 * x]) + bot[2 * x] + 1u) >> 1, v1 = (unsigned(top[2 * x + 1]) + bot[2 * x + 1] + 1u) >> 1;
        out[size_t(y) * oc + x] = uint8_t((v0 + v1 + 1u) >> 1);
      } else {
        out[size_t(y) * oc + x] = uint8_t((uint16_t(top[2 * x]) + top[2 * x + 1] + bot[2 * x] + bot[2 * x + 1]) / 4);
      }
    }
  }
}
void orc_half_sample(const uint8_t* in, int in_cols, int in_rows, uint8_t* out) {
  orc_half_sample_rule(in, in_cols, in_rows, out, 0);
}

// [EXT] camera models, exposed for the tests: n points, xyz (n*3) -> px (n*2) and px (n*2) -> unit bearing (n*3)
void orc_camera_world2cam(const orc_camera* cam, const double* xyz, int n, double* px_out) {
  const Cam c = make_cam(cam);
  for (int i = 0; i < n; ++i) {
    const V2 p = c.world2cam(V3{xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]});
    px_out[2 * i] = p.x; px_out[2 * i + 1] = p.y;
  }
}
void orc_camera_cam2world(const orc_camera* cam, const double* px, int n, double* f_out) {
  const Cam c = make_cam(cam);
  for (int i = 0; i < n; ++i) {
    const V3 f = c.cam2world(px[2 * i], px[2 * i + 1]);
    f_out[3 * i] = f.x; f_out[3 * i + 1] = f.y; f_out[3 * i + 2] = f.z;
  }
}

void orc_se3_exp(const double* x6, double* T12_out) { se3_to_rt12(se3_exp(x6), T12_out); }
void orc_se3_mul(const double* A12, const double* B12, double* C12_out) {
  se3_to_rt12(se3_from_rt12(A12) * se3_from_rt12(B12), C12_out);
}
void orc_se3_inv(const double* A12, double* C12_out) {
  se3_to_rt12(inverse(se3_from_rt12(A12)), C12_out);
}
void orc_ldlt6_solve(const double* H36, const double* b6, double* x6_out) {
  double H[6][6];
  for (int a = 0; a < 6; ++a)
    for (int b = 0; b < 6; ++b) H[a][b] = H36[a * 6 + b];
  LDLT<6> l;
  l.compute(H);
  l.solve(b6, x6_out);
}

}  // extern "C"

#include "svo_oracle_align.inc"
#include "svo_oracle_depth.inc"
#include "svo_oracle_pose.inc"
#include "svo_oracle_reproject.inc"
#include "fast_ext.h"
#include "svo_oracle_detect.inc"
