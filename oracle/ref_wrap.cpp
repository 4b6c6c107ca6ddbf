// oracle/ref_wrap.cpp -- TEST INFRASTRUCTURE ONLY.
// C entry points around the reference's OWN classes (svo::SparseImgAlign, svo::Matcher, svo::pose_optimizer,
// svo::Point, svo::feature_alignment), compiled from the sources where they lie in /root/reference/svo/src against
// the stand-in third-party headers of oracle/shim/ (Eigen, Sophus, vikit, OpenCV, Boost are absent here).  Used by the
// tests to validate the oracle's restatement of everything that IS in the reference tree; the third-party arithmetic
// inside the stand-ins remains the unpinned [EXT] boundary.
#include <svo/config.h>
#include <svo/depth_filter.h>
#include <svo/feature.h>
#include <svo/feature_alignment.h>
#include <svo/feature_detection.h>
#include <svo/frame.h>
#include <svo/matcher.h>
#include <svo/point.h>
#include <svo/pose_optimizer.h>
#include <svo/sparse_img_align.h>
#include <vikit/abstract_camera.h>

#include <chrono>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

using namespace svo;

// camera from the flat parameter block of oracle/binding.py (_cam4): [fx fy cx cy model d0..d4]; model 0 =
// vk::PinholeCamera(w, h, fx, fy, cx, cy, d0..d4), model 1 = vk::ATANCamera with the pixel parameters converted back to
// the normalised ones its constructor expects
vk::AbstractCamera* ref_make_camera(int w, int h, const double* c) {
  if ((int)c[4] == 1) return new vk::ATANCamera(w, h, c[0] / w, c[1] / h, (c[2] + 0.5) / w, (c[3] + 0.5) / h, c[5]);
  return new vk::PinholeCamera(w, h, c[0], c[1], c[2], c[3], c[5], c[6], c[7], c[8], c[9]);
}

// wall-clock seconds of the reference algorithm inside the most recent ref_* call (object construction, pyramid building
// and result copying excluded): what bench.py's CPU legs report
double g_ref_last_seconds = 0.0;
struct RefTimer {
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  ~RefTimer() { g_ref_last_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); }
};

namespace {
SE3 se3_from12(const double* T) {
  Matrix3d R;
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) R(i, j) = T[i * 4 + j];
  return SE3(R, Vector3d(T[3], T[7], T[11]));
}
void se3_to12(const SE3& S, double* T) {
  const Matrix3d R = S.rotation_matrix();
  const Vector3d t = S.translation();
  for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) T[i * 4 + j] = R(i, j); T[i * 4 + 3] = t[i]; }
}
struct SIA : public SparseImgAlign {
  SIA(int a, int b, int c) : SparseImgAlign(a, b, c, GaussNewton, false, false) {}
  const std::vector<bool>& visible() const { return visible_fts_; }
  const cv::Mat& patch_cache() const { return ref_patch_cache_; }
  const Matrix<double, 6, 6>& H() const { return H_; }
};
FramePtr make_frame(vk::AbstractCamera* cam, const uint8_t* img, int w, int h, int n_levels, const double* T_f_w) {
  Config::nPyrLevels() = 1;
  Config::kltMaxLevel() = n_levels - 1;  // Frame::initFrame builds max(nPyrLevels, kltMaxLevel+1) levels (frame.cpp:58)
  // the image goes into an aligned cv::Mat of its own, as the reference's callers hand it over (cv::imread /
  // cv_bridge buffers are 16-byte aligned): vk::halfSample's SSE2 branch tests the alignment of level 0
  cv::Mat m(h, w, CV_8U);
  std::memcpy(m.data, img, (size_t)w * h);
  FramePtr f(new Frame(cam, m, 0.0));
  f->T_f_w_ = se3_from12(T_f_w);
  return f;
}
}  // namespace

// [EXT] the `fast` library and vk::shiTomasiScore behind the reference's feature_detection.cpp: the restatement of
// oracle/fast_ext.h (PARITY UNPINNED there; what oracle/_ref pins is FastDetector::detect itself).
#include <fast/fast.h>
#include <vikit/vision.h>
#include "fast_ext.h"
namespace fast {
static void detect_any(const fast_byte* img, int w, int h, int stride, short b, std::vector<fast_xy>& corners) {
  std::vector<fast_ext::xy> c;
  fast_ext::detect10(img, w, h, stride, b, c);
  corners.clear();
  for (auto& q : c) corners.push_back(fast_xy(q.x, q.y));
}
void fast_corner_detect_10(const fast_byte* img, int w, int h, int stride, short b, std::vector<fast_xy>& c) { detect_any(img, w, h, stride, b, c); }
void fast_corner_detect_10_sse2(const fast_byte* img, int w, int h, int stride, short b, std::vector<fast_xy>& c) { detect_any(img, w, h, stride, b, c); }
void fast_corner_detect_10_neon(const fast_byte* img, int w, int h, int stride, short b, std::vector<fast_xy>& c) { detect_any(img, w, h, stride, b, c); }
void fast_corner_score_10(const fast_byte* img, int stride, const std::vector<fast_xy>& corners, int b, std::vector<int>& scores) {
  std::vector<fast_ext::xy> c;
  for (auto& q : corners) c.push_back(fast_ext::xy{q.x, q.y});
  fast_ext::score10(img, stride, c, b, scores);
}
void fast_nonmax_3x3(const std::vector<fast_xy>& corners, const std::vector<int>& scores, std::vector<int>& nonmax) {
  std::vector<fast_ext::xy> c;
  for (auto& q : corners) c.push_back(fast_ext::xy{q.x, q.y});
  fast_ext::nonmax3x3(c, scores, nonmax);
}
}  // namespace fast
namespace vk {
float shiTomasiScore(const cv::Mat& img, int u, int v) { return fast_ext::shiTomasiScore(img.data, img.cols, img.rows, (int)img.step.p[0], u, v); }
}

extern "C" {

// frame_utils::createImgPyramid of the reference's frame.cpp (through Frame::initFrame): levels 0..n_levels-1 written
// back to back into `out` (level l has (w >> l) x (h >> l) pixels).
void ref_image_pyramid(const uint8_t* img, int w, int h, int n_levels, uint8_t* out) {
  vk::PinholeCamera cam(w, h, 300.0, 300.0, w / 2.0, h / 2.0);
  const double T[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
  FramePtr f = make_frame(&cam, img, w, h, n_levels, T);
  size_t o = 0;
  for (int l = 0; l < n_levels; ++l) {
    const cv::Mat& m = f->img_pyr_[l];
    for (int y = 0; y < m.rows; ++y) std::memcpy(out + o + (size_t)y * m.cols, m.data + (size_t)y * m.step.p[0], m.cols);
    o += (size_t)m.rows * m.cols;
  }
}

// svo::SparseImgAlign::run on two frames built from level-0 images (pyramids by the reference's createImgPyramid).
long long ref_sparse_img_align(const uint8_t* ref_l0, const uint8_t* cur_l0, int w, int h, int n_levels, const double* cam4,
                               const double* T_ref_w, double* T_cur_w_io, const double* px, const double* f, const double* pos,
                               const uint8_t* has_point, int N, int max_level, int min_level, int n_iter,
                               uint8_t* visible_out, double* H_out, float* patch_cache_out) {
  std::unique_ptr<vk::AbstractCamera> cam_owner(ref_make_camera(w, h, cam4));
  vk::AbstractCamera& cam = *cam_owner;
  FramePtr ref = make_frame(&cam, ref_l0, w, h, n_levels, T_ref_w);
  FramePtr cur = make_frame(&cam, cur_l0, w, h, n_levels, T_cur_w_io);
  std::vector<std::unique_ptr<Point>> pts;
  for (int i = 0; i < N; ++i) {
    Point* p = nullptr;
    if (has_point[i]) { pts.emplace_back(new Point(Vector3d(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]))); p = pts.back().get(); }
    ref->addFeature(new Feature(ref.get(), p, Vector2d(px[2 * i], px[2 * i + 1]), Vector3d(f[3 * i], f[3 * i + 1], f[3 * i + 2]), 0));
  }
  SIA sia(max_level, min_level, n_iter);
  const size_t ret = sia.run(ref, cur);
  se3_to12(cur->T_f_w_, T_cur_w_io);
  if (visible_out) for (int i = 0; i < N && i < (int)sia.visible().size(); ++i) visible_out[i] = sia.visible()[i];
  if (H_out) for (int a = 0; a < 6; ++a) for (int b = 0; b < 6; ++b) H_out[a * 6 + b] = sia.H()(a, b);
  if (patch_cache_out && N) memcpy(patch_cache_out, sia.patch_cache().data, sizeof(float) * 16 * (size_t)N);
  return (long long)ret;
}

// svo::SparseImgAlign::computeResiduals (sparse_img_align.cpp:147-243) of the compiled reference at ONE level and ONE
// pose, called directly (it is protected: the derived class below reaches it), with everything it leaves behind:
// the visibility mask and patch cache of precomputeReferencePatches, H_, Jres_, chi2, n_meas_ -- and the per-pixel
// residual magnitudes, captured by handing the solver a scale estimator that copies the `errors` vector the reference
// fills when compute_weight_scale is set (|res| of every pixel of every in-image patch, in feature order).
namespace {
struct CapturingScale : public vk::robust_cost::ScaleEstimator {
  mutable std::vector<float> got;
  float compute(std::vector<float>& errors) const override { got = errors; return 1.0f; }
};
struct SIAResiduals : public SparseImgAlign {
  static const int kArea = 16;  // patch_size_ * patch_size_ (sparse_img_align.h:35-37, private there)
  std::shared_ptr<CapturingScale> cap{new CapturingScale};
  SIAResiduals(int level) : SparseImgAlign(level, level, 1, GaussNewton, false, false) { scale_estimator_ = cap; }
  double eval(FramePtr ref, FramePtr cur, int level, const SE3& T_cur_from_ref) {
    ref_frame_ = ref; cur_frame_ = cur; level_ = level;
    ref_patch_cache_ = cv::Mat(ref->fts_.size(), kArea, CV_32F);   // as run() sizes them (:55-57); patch_area_ is private
    jacobian_cache_.resize(Eigen::NoChange, ref_patch_cache_.rows * kArea);
    visible_fts_.assign(ref_patch_cache_.rows, false);
    jacobian_cache_.setZero();
    have_ref_patch_cache_ = false;
    H_.setZero(); Jres_.setZero(); n_meas_ = 0; iter_ = 0;
    return computeResiduals(T_cur_from_ref, true, true);
  }
  const std::vector<bool>& visible() const { return visible_fts_; }
  const cv::Mat& patch_cache() const { return ref_patch_cache_; }
  const Matrix<double, 6, 6>& H() const { return H_; }
  const Matrix<double, 6, 1>& Jres() const { return Jres_; }
  size_t n_meas() const { return n_meas_; }
};
}  // namespace
long long ref_sparse_residuals(const uint8_t* ref_l0, const uint8_t* cur_l0, int w, int h, int n_levels, const double* cam4,
                               const double* T_ref_w, const double* T_cur_w, const double* px, const double* f, const double* pos,
                               const uint8_t* has_point, int N, int level, uint8_t* visible_out, float* patch_cache_out,
                               double* H_out, double* Jres_out, double* chi2_out, float* abs_res_out, long long abs_res_cap) {
  std::unique_ptr<vk::AbstractCamera> cam_owner(ref_make_camera(w, h, cam4));
  vk::AbstractCamera& cam = *cam_owner;
  FramePtr ref = make_frame(&cam, ref_l0, w, h, n_levels, T_ref_w);
  FramePtr cur = make_frame(&cam, cur_l0, w, h, n_levels, T_cur_w);
  std::vector<std::unique_ptr<Point>> pts;
  for (int i = 0; i < N; ++i) {
    Point* p = nullptr;
    if (has_point[i]) { pts.emplace_back(new Point(Vector3d(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]))); p = pts.back().get(); }
    ref->addFeature(new Feature(ref.get(), p, Vector2d(px[2 * i], px[2 * i + 1]), Vector3d(f[3 * i], f[3 * i + 1], f[3 * i + 2]), 0));
  }
  SIAResiduals sia(level);
  const SE3 T_cur_from_ref(cur->T_f_w_ * ref->T_f_w_.inverse());  // as run() forms it (:66)
  const double chi2 = sia.eval(ref, cur, level, T_cur_from_ref);
  if (visible_out) for (int i = 0; i < N; ++i) visible_out[i] = sia.visible()[i];
  if (patch_cache_out && N) memcpy(patch_cache_out, sia.patch_cache().data, sizeof(float) * 16 * (size_t)N);
  if (H_out) for (int a = 0; a < 6; ++a) for (int b = 0; b < 6; ++b) H_out[a * 6 + b] = sia.H()(a, b);
  if (Jres_out) for (int a = 0; a < 6; ++a) Jres_out[a] = sia.Jres()(a);
  if (chi2_out) *chi2_out = chi2;
  const long long n_abs = (long long)sia.cap->got.size();
  if (abs_res_out) for (long long k = 0; k < n_abs && k < abs_res_cap; ++k) abs_res_out[k] = sia.cap->got[k];
  return (long long)sia.n_meas();
}

// ---- a stream of frame pairs kept alive between calls, for timing svo::SparseImgAlign::run alone (bench.py's
// --impl reference arm and cpu_baseline leg): pair k = (frame k, frame k+1); pyramids are built at create time.
struct RefStream {
  std::unique_ptr<vk::AbstractCamera> cam;
  std::vector<FramePtr> ref, cur;  // separate objects: run() writes cur->T_f_w_
  std::vector<std::unique_ptr<Point>> pts;
};
void* ref_stream_create(const uint8_t* level0s /*(B+1) images*/, int B, int w, int h, int n_levels, const double* cam4,
                        const double* T_f_w /*(B+1)x12*/, const int* feat_offset, const double* px, const double* f,
                        const double* pos, const uint8_t* has_point) {
  RefStream* s = new RefStream;
  s->cam.reset(ref_make_camera(w, h, cam4));
  for (int k = 0; k < B; ++k) {
    FramePtr ref = make_frame(s->cam.get(), level0s + (size_t)k * w * h, w, h, n_levels, T_f_w + 12 * k);
    for (int i = feat_offset[k]; i < feat_offset[k + 1]; ++i) {
      Point* p = nullptr;
      if (has_point[i]) { s->pts.emplace_back(new Point(Vector3d(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]))); p = s->pts.back().get(); }
      ref->addFeature(new Feature(ref.get(), p, Vector2d(px[2 * i], px[2 * i + 1]), Vector3d(f[3 * i], f[3 * i + 1], f[3 * i + 2]), 0));
    }
    s->ref.push_back(ref);
    s->cur.push_back(make_frame(s->cam.get(), level0s + (size_t)(k + 1) * w * h, w, h, n_levels, T_f_w + 12 * k));
  }
  return s;
}
// Runs SparseImgAlign on every pair from the frame handler's initial guess (cur.T_f_w_ = ref.T_f_w_,
// frame_handler_mono.cpp:129) with n_threads workers.  Returns the
// wall-clock seconds of the alignment alone; T_cur_from_ref_out (Bx12) may be NULL.
double ref_stream_run(void* handle, int n_threads, int max_level, int min_level, int n_iter, double* T_cur_from_ref_out,
                      long long* n_tracked_out) {
  RefStream* s = (RefStream*)handle;
  const int B = (int)s->ref.size();
  if (n_threads < 1) n_threads = 1;
  if (n_threads > B) n_threads = B;
  auto work = [&](int t) {
    for (int k = t; k < B; k += n_threads) {
      // a fresh object per frame, as FrameHandlerMono::processFrame does (frame_handler_mono.cpp:131): visible_fts_ is
      // only ever resize()d, so a reused object would carry stale visibility flags into the next pair
      SparseImgAlign sia(max_level, min_level, n_iter, SparseImgAlign::GaussNewton, false, false);
      s->cur[k]->T_f_w_ = s->ref[k]->T_f_w_;
      const size_t n = sia.run(s->ref[k], s->cur[k]);
      if (n_tracked_out) n_tracked_out[k] = (long long)n;
      if (T_cur_from_ref_out) se3_to12(s->cur[k]->T_f_w_ * s->ref[k]->T_f_w_.inverse(), T_cur_from_ref_out + 12 * k);
    }
  };
  const auto t0 = std::chrono::steady_clock::now();
  if (n_threads == 1) work(0);
  else {
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; ++t) th.emplace_back(work, t);
    for (auto& x : th) x.join();
  }
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}
void ref_stream_destroy(void* handle) { delete (RefStream*)handle; }

// svo::pose_optimizer::optimizeGaussNewton on a frame built from flat arrays.
void ref_pose_optimize(double reproj_thresh, int n_iter, const double* cam4, int w, int h, double* T_f_w_io, const double* f,
                       const double* pos, const int* level, uint8_t* has_point_io, int N, double* scalars4 /*scale,init,final,num_obs*/,
                       double* cov36) {
  std::unique_ptr<vk::AbstractCamera> cam_owner(ref_make_camera(w, h, cam4));
  vk::AbstractCamera& cam = *cam_owner;
  std::vector<uint8_t> img((size_t)w * h, 0);
  FramePtr fr = make_frame(&cam, img.data(), w, h, 1, T_f_w_io);
  std::vector<std::unique_ptr<Point>> pts;
  for (int i = 0; i < N; ++i) {
    Point* p = nullptr;
    if (has_point_io[i]) { pts.emplace_back(new Point(Vector3d(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]))); p = pts.back().get(); }
    fr->addFeature(new Feature(fr.get(), p, Vector2d(0, 0), Vector3d(f[3 * i], f[3 * i + 1], f[3 * i + 2]), level[i]));
  }
  double scale = 0, e_init = 0, e_final = 0;
  size_t num_obs = 0;
  { RefTimer tm; pose_optimizer::optimizeGaussNewton(reproj_thresh, (size_t)n_iter, false, fr, scale, e_init, e_final, num_obs); }
  se3_to12(fr->T_f_w_, T_f_w_io);
  int i = 0;
  for (auto it = fr->fts_.begin(); it != fr->fts_.end(); ++it, ++i) has_point_io[i] = (*it)->point != NULL;
  scalars4[0] = scale; scalars4[1] = e_init; scalars4[2] = e_final; scalars4[3] = (double)num_obs;
  for (int a = 0; a < 6; ++a) for (int b = 0; b < 6; ++b) cov36[a * 6 + b] = fr->Cov_(a, b);
}

// svo::Point::optimize with n_obs observing frames.
void ref_point_optimize(int n_iter, double* pos_io, int n_obs, const double* obs_T_f_w, const double* obs_f) {
  vk::PinholeCamera cam(16, 16, 10, 10, 8, 8);
  std::vector<uint8_t> img(256, 0);
  std::vector<FramePtr> frames;
  Point pt(Vector3d(pos_io[0], pos_io[1], pos_io[2]));
  for (int o = 0; o < n_obs; ++o) {
    frames.push_back(make_frame(&cam, img.data(), 16, 16, 1, obs_T_f_w + 12 * o));
    Feature* ft = new Feature(frames.back().get(), &pt, Vector2d(0, 0), Vector3d(obs_f[3 * o], obs_f[3 * o + 1], obs_f[3 * o + 2]), 0);
    frames.back()->addFeature(ft);
    pt.obs_.push_back(ft);  // list order = observation order (addFrameRef pushes front; order is irrelevant to the sums' terms)
  }
  pt.optimize((size_t)n_iter);
  for (int k = 0; k < 3; ++k) pos_io[k] = pt.pos_[k];
}

// svo::Matcher::findMatchDirect / findEpipolarMatchDirect on two frames built from level-0 images.
struct ref_match_out { int success, search_level, reject; double px_cur[2], A[4], h_inv, epi_length, depth; };
void ref_matcher(int mode /*0 direct, 1 epipolar*/, const uint8_t* ref_l0, const uint8_t* cur_l0, int w, int h, int n_levels,
                 const double* cam4, const double* T_ref_w, const double* T_cur_w, const double* ref_px, const double* ref_f,
                 int ref_level, int ftr_type, const double* ref_grad, const double* point_pos, const double* px_cur_in,
                 double d_est, double d_min, double d_max, int n_pyr_levels, ref_match_out* out) {
  std::unique_ptr<vk::AbstractCamera> cam_owner(ref_make_camera(w, h, cam4));
  vk::AbstractCamera& cam = *cam_owner;
  FramePtr ref = make_frame(&cam, ref_l0, w, h, n_levels, T_ref_w);
  FramePtr cur = make_frame(&cam, cur_l0, w, h, n_levels, T_cur_w);
  Config::nPyrLevels() = n_pyr_levels;  // max search level = nPyrLevels()-1 (matcher.cpp:153,214)
  Point pt(Vector3d(point_pos[0], point_pos[1], point_pos[2]));
  Feature* ft = new Feature(ref.get(), &pt, Vector2d(ref_px[0], ref_px[1]), Vector3d(ref_f[0], ref_f[1], ref_f[2]), ref_level);
  ft->type = ftr_type ? Feature::EDGELET : Feature::CORNER;
  ft->grad = Vector2d(ref_grad[0], ref_grad[1]);
  ref->addFeature(ft);
  pt.addFrameRef(ft);
  Matcher m;
  memset(m.patch_, 0, sizeof(m.patch_));
  memset(m.patch_with_border_, 0, sizeof(m.patch_with_border_));
  m.search_level_ = 0; m.h_inv_ = 0; m.epi_length_ = 0; m.reject_ = false;
  m.px_cur_ = Vector2d(0, 0); m.A_cur_ref_.setZero();
  memset(out, 0, sizeof(*out));
  if (mode == 0) {
    Vector2d px(px_cur_in[0], px_cur_in[1]);
    out->success = m.findMatchDirect(pt, *cur, px);
    out->px_cur[0] = px[0]; out->px_cur[1] = px[1];
  } else {
    double depth = 0;
    out->success = m.findEpipolarMatchDirect(*ref, *cur, *ft, d_est, d_min, d_max, depth);
    out->px_cur[0] = m.px_cur_[0]; out->px_cur[1] = m.px_cur_[1];
    out->depth = depth;
  }
  out->search_level = m.search_level_; out->reject = m.reject_;
  out->A[0] = m.A_cur_ref_(0, 0); out->A[1] = m.A_cur_ref_(0, 1); out->A[2] = m.A_cur_ref_(1, 0); out->A[3] = m.A_cur_ref_(1, 1);
  out->h_inv = m.h_inv_; out->epi_length = m.epi_length_;
}

int ref_align2d(const uint8_t* img, int cols, int rows, int step, uint8_t* pwb, uint8_t* patch, int n_iter, double* px_io) {
  cv::Mat m(rows, cols, const_cast<uint8_t*>(img), (size_t)step);
  Eigen::Vector2d px(px_io[0], px_io[1]);
  const bool ok = svo::feature_alignment::align2D(m, pwb, patch, n_iter, px, true);
  px_io[0] = px[0]; px_io[1] = px[1];
  return ok ? 1 : 0;
}
double ref_last_seconds() { return g_ref_last_seconds; }

// M align2D calls in one C loop (no per-call FFI overhead): problem i runs on level[i] of the pyramid built from l0
void ref_align2d_batch(const uint8_t* l0, int w, int h, int n_levels, int M, const int* level, const uint8_t* pwb, const uint8_t* patch,
                       int n_iter, double* px_io, uint8_t* conv_out) {
  vk::PinholeCamera cam(w, h, 300, 300, w / 2.0, h / 2.0);
  const double I[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
  FramePtr fr = make_frame(&cam, l0, w, h, n_levels, I);
  std::vector<uint8_t> a(100), b(64);
  RefTimer tm;
  for (int i = 0; i < M; ++i) {
    std::memcpy(a.data(), pwb + 100 * (size_t)i, 100);
    std::memcpy(b.data(), patch + 64 * (size_t)i, 64);
    Eigen::Vector2d px(px_io[2 * i], px_io[2 * i + 1]);
    conv_out[i] = svo::feature_alignment::align2D(fr->img_pyr_[level[i]], a.data(), b.data(), n_iter, px, true) ? 1 : 0;
    px_io[2 * i] = px[0]; px_io[2 * i + 1] = px[1];
  }
}

int ref_align1d(const uint8_t* img, int cols, int rows, int step, const float* dir, uint8_t* pwb, uint8_t* patch, int n_iter,
                double* px_io, double* h_inv) {
  cv::Mat m(rows, cols, const_cast<uint8_t*>(img), (size_t)step);
  Eigen::Vector2d px(px_io[0], px_io[1]);
  Eigen::Vector2f d(dir[0], dir[1]);
  const bool ok = svo::feature_alignment::align1D(m, d, pwb, patch, n_iter, px, *h_inv);
  px_io[0] = px[0]; px_io[1] = px[1];
  return ok ? 1 : 0;
}

// svo::DepthFilter::updateSeeds (depth_filter.cpp:197-291) on seeds built from flat arrays; the filter thread is never
// started.  status_out: 0 = still in the list, 1 = converged (callback fired), 2 = erased without converging.
struct RefDF : public DepthFilter {
  RefDF(callback_t cb) : DepthFilter(feature_detection::DetectorPtr(), cb) {}
  void update(FramePtr f) { updateSeeds(f); }
};
void ref_depth_filter_update(const uint8_t* ref_l0s /*n_ref images*/, const double* ref_T_f_w, int n_ref, const uint8_t* cur_l0,
                             const double* cur_T_f_w, int w, int h, int n_levels, const double* cam4, int M, const int* ref_index,
                             const double* ftr_px, const double* ftr_f, const int* ftr_level, const int* ftr_type,
                             const double* ftr_grad, const int* batch_id, int batch_counter, int n_pyr_levels, float* a, float* b,
                             float* mu, float* z_range, float* sigma2, uint8_t* status_out, double* xyz_world_out) {
  std::unique_ptr<vk::AbstractCamera> cam_owner(ref_make_camera(w, h, cam4));
  vk::AbstractCamera& cam = *cam_owner;
  std::vector<FramePtr> refs;
  for (int r = 0; r < n_ref; ++r) refs.push_back(make_frame(&cam, ref_l0s + (size_t)r * w * h, w, h, n_levels, ref_T_f_w + 12 * r));
  FramePtr cur = make_frame(&cam, cur_l0, w, h, n_levels, cur_T_f_w);
  Config::nPyrLevels() = n_pyr_levels;
  std::vector<Feature*> fts(M);
  std::vector<Point*> made;
  std::vector<double> conv_sigma2;
  RefDF df([&](Point* p, double s2) { made.push_back(p); conv_sigma2.push_back(s2); });
  for (int i = 0; i < M; ++i) {
    Frame* fr = refs[ref_index[i]].get();
    Feature* ft = new Feature(fr, Vector2d(ftr_px[2 * i], ftr_px[2 * i + 1]), Vector3d(ftr_f[3 * i], ftr_f[3 * i + 1], ftr_f[3 * i + 2]), ftr_level[i]);
    ft->type = ftr_type[i] ? Feature::EDGELET : Feature::CORNER;
    ft->grad = Vector2d(ftr_grad[2 * i], ftr_grad[2 * i + 1]);
    fr->addFeature(ft);
    fts[i] = ft;
    Seed s(ft, 1.0f, 0.5f);
    s.id = i; s.batch_id = batch_id[i];
    s.a = a[i]; s.b = b[i]; s.mu = mu[i]; s.z_range = z_range[i]; s.sigma2 = sigma2[i];
    df.getSeeds().push_back(s);
  }
  Seed::batch_counter = batch_counter;
  { RefTimer tm; df.update(cur); }
  for (int i = 0; i < M; ++i) status_out[i] = 2;
  for (auto& s : df.getSeeds()) {
    const int i = s.id;
    status_out[i] = 0;
    a[i] = s.a; b[i] = s.b; mu[i] = s.mu; z_range[i] = s.z_range; sigma2[i] = s.sigma2;
  }
  for (int i = 0; i < M; ++i) {
    if (fts[i]->point == NULL) continue;
    status_out[i] = 1;
    for (size_t k = 0; k < made.size(); ++k)
      if (made[k] == fts[i]->point) sigma2[i] = (float)conv_sigma2[k];
    for (int k = 0; k < 3; ++k) xyz_world_out[3 * i + k] = fts[i]->point->pos_[k];
  }
  for (Point* p : made) delete p;
}

// feature_detection::FastDetector::detect on a frame built from a level-0 image (pyramid by createImgPyramid).
int ref_fast_detect(const uint8_t* l0, int w, int h, int n_levels, int n_pyr_levels, int cell_size, const uint8_t* grid_occupancy,
                    double detection_threshold, int* out_x, int* out_y, int* out_level, int cap) {
  vk::PinholeCamera cam(w, h, 300, 300, w / 2.0, h / 2.0);
  const double I[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
  FramePtr fr = make_frame(&cam, l0, w, h, n_levels, I);
  feature_detection::FastDetector det(w, h, cell_size, n_pyr_levels);
  const int n_cols = (int)ceil((double)w / cell_size), n_rows = (int)ceil((double)h / cell_size);
  if (grid_occupancy)
    for (int r = 0; r < n_rows; ++r)
      for (int c = 0; c < n_cols; ++c)
        if (grid_occupancy[r * n_cols + c]) det.setGridOccpuancy(Vector2d(c * cell_size + 0.5, r * cell_size + 0.5));
  Features fts;
  { RefTimer tm; det.detect(fr.get(), fr->img_pyr_, detection_threshold, fts); }
  int n = 0;
  for (Feature* f : fts) {
    if (n < cap) { out_x[n] = (int)f->px[0]; out_y[n] = (int)f->px[1]; out_level[n] = f->level; }
    ++n;
    delete f;
  }
  return n;
}

void ref_update_seed(float x, float tau2, float* a, float* b, float* mu, float* z_range, float* sigma2) {
  Seed s(NULL, 1.0f, 0.5f);
  s.a = *a; s.b = *b; s.mu = *mu; s.z_range = *z_range; s.sigma2 = *sigma2;
  DepthFilter::updateSeed(x, tau2, &s);
  *a = s.a; *b = s.b; *mu = s.mu; *z_range = s.z_range; *sigma2 = s.sigma2;
}
double ref_compute_tau(const double* T_ref_cur, const double* f, double z, double px_error_angle) {
  return DepthFilter::computeTau(se3_from12(T_ref_cur), Vector3d(f[0], f[1], f[2]), z, px_error_angle);
}

}  // extern "C"
