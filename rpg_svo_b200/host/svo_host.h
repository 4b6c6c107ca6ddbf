// rpg_svo_b200/host/svo_host.h -- C++ host classes that keep the reference's call surface for the
// hot path and forward to the C ABI (include/svo_b200.h).  Header-only, depends on nothing but the
// C++17 standard library and libsvo_b200.so.
//
// What is mirrored (names, argument meaning, error behaviour):
//   svo::SparseImgAlign(max_level, min_level, n_iter, method, display, verbose)::run(ref, cur)
//       + getFisherInformation()                      svo/include/svo/sparse_img_align.h:43-57
//   svo::pose_optimizer::optimizeGaussNewton(...)     svo/include/svo/pose_optimizer.h:37-45
//   svo::feature_alignment::align2D / align1D         svo/include/svo/feature_alignment.h:29-44
//   svo::DepthFilter::{addFrame, addKeyframe(seeds), removeKeyframe, reset, getSeeds, updateSeeds}
//       + static updateSeed/computeTau stay host-side in the reference and are not re-exported
//                                                     svo/include/svo/depth_filter.h:101-158
//   svo::Frame / Feature / Point / Seed               svo/include/svo/{frame,feature,point,depth_filter}.h
// The data model is the reference's pointer graph (std::list<Feature*>, Point*); the wrappers gather it
// into the flat arrays the C ABI takes -- that gather is the cost SURVEY.md row a18 says must be
// counted end to end.  Differences from the reference, all forced by the missing third-party types:
//   * Eigen/Sophus/cv::Mat are replaced by the minimal Vector2d/Vector3d/SE3/Image below;
//   * FramePtr is std::shared_ptr (reference: boost::shared_ptr);
//   * the camera is the pinhole-without-distortion model only;
//   * DepthFilter has no detector: addKeyframe takes the new features explicitly, and the mapper
//     thread / halt flag are the caller's (updateSeeds is synchronous, as in the reference when
//     thread_ == NULL, depth_filter.cpp:95-96).
#pragma once
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <list>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/svo_b200.h"

namespace svo {

using Vector2d = std::array<double, 2>;
using Vector3d = std::array<double, 3>;
using Matrix6d = std::array<double, 36>;  // row-major

// Minimal rigid transform, row-major [R|t]; only what the wrappers need on the host.
struct SE3 {
  double m[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
  SE3 operator*(const SE3& o) const {
    SE3 r;
    for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < 3; ++j)
        r.m[i * 4 + j] = m[i * 4] * o.m[j] + m[i * 4 + 1] * o.m[4 + j] + m[i * 4 + 2] * o.m[8 + j];
      r.m[i * 4 + 3] = m[i * 4] * o.m[3] + m[i * 4 + 1] * o.m[7] + m[i * 4 + 2] * o.m[11] + m[i * 4 + 3];
    }
    return r;
  }
  SE3 inverse() const {
    SE3 r;
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) r.m[i * 4 + j] = m[j * 4 + i];
    for (int i = 0; i < 3; ++i) r.m[i * 4 + 3] = -(r.m[i * 4] * m[3] + r.m[i * 4 + 1] * m[7] + r.m[i * 4 + 2] * m[11]);
    return r;
  }
  Vector3d translation() const { return {m[3], m[7], m[11]}; }
};

struct PinholeCamera {  // [EXT] vk::PinholeCamera without distortion
  int width_, height_;
  double fx_, fy_, cx_, cy_;
  PinholeCamera(int w, int h, double fx, double fy, double cx, double cy) : width_(w), height_(h), fx_(fx), fy_(fy), cx_(cx), cy_(cy) {}
  Vector3d cam2world(const Vector2d& px) const {
    double x = (px[0] - cx_) / fx_, y = (px[1] - cy_) / fy_, n = std::sqrt(x * x + y * y + 1.0);
    return {x / n, y / n, 1.0 / n};
  }
  double errorMultiplier2() const { return std::fabs(fx_); }
  svo_b200_camera c_abi() const { return svo_b200_camera{fx_, fy_, cx_, cy_, width_, height_}; }
};

// One CUDA context per calling thread, as include/svo_b200.h asks.
class Context {
 public:
  explicit Context(int device = 0) {
    if (svo_b200_create(&ctx_, device) != 0)
      throw std::runtime_error("svo_b200_create failed: no usable CUDA device (there is no CPU fallback)");
  }
  ~Context() { svo_b200_destroy(ctx_); }
  Context(const Context&) = delete;
  Context& operator=(const Context&) = delete;
  svo_b200_ctx* get() const { return ctx_; }
  void check(int rc) const {
    if (rc != 0) throw std::runtime_error(std::string("svo_b200: ") + svo_b200_last_error(ctx_));
  }

 private:
  svo_b200_ctx* ctx_ = nullptr;
};

class Frame;
struct Feature;
struct Point {  // svo/include/svo/point.h:35-106 (the fields the hot path reads)
  enum PointType { TYPE_DELETED, TYPE_CANDIDATE, TYPE_UNKNOWN, TYPE_GOOD };  // point.h:40-45
  Vector3d pos_;
  std::list<Feature*> obs_;        // references to keyframes which observe the point (point.h:52)
  PointType type_ = TYPE_UNKNOWN;  // point.h:58
  int n_failed_reproj_ = 0;        // point.h:59
  int n_succeeded_reproj_ = 0;     // point.h:60
  explicit Point(const Vector3d& pos) : pos_(pos) {}
  Point(const Vector3d& pos, Feature* ftr) : pos_(pos) { obs_.push_front(ftr); }  // point.cpp:38-50
  void addFrameRef(Feature* ftr) { obs_.push_front(ftr); }                        // point.cpp:55-59
};

struct Feature {  // svo/include/svo/feature.h:25-71
  enum FeatureType { CORNER, EDGELET };
  FeatureType type = CORNER;
  Frame* frame;
  Vector2d px;
  Vector3d f;
  int level;
  Point* point = nullptr;
  Vector2d grad{1.0, 0.0};
  Feature(Frame* _frame, const Vector2d& _px, int _level);
  Feature(Frame* _frame, Point* _point, const Vector2d& _px, const Vector3d& _f, int _level)
      : frame(_frame), px(_px), f(_f), level(_level), point(_point) {}
};
typedef std::list<Feature*> Features;

// svo/include/svo/frame.h:40-139.  The image pyramid lives in HBM (level 0 uploaded once, the other
// levels built on the device with the scalar vk::halfSample rule).
class Frame {
 public:
  PinholeCamera* cam_;
  SE3 T_f_w_;
  Matrix6d Cov_{};
  Features fts_;
  std::vector<Feature*> key_pts_ = std::vector<Feature*>(5, nullptr);  // frame.h:53 (maintained by the caller: setKeyPoints)
  bool is_keyframe_ = false;
  Frame(Context& ctx, PinholeCamera* cam, const uint8_t* img, int n_levels, double /*timestamp*/) : cam_(cam), ctx_(ctx) {
    if (!img) throw std::runtime_error("Frame: provided image is empty");  // frame.cpp:51-52
    ctx_.check(svo_b200_frame_create(ctx_.get(), cam->width_, cam->height_, n_levels, &dev_));
    const uint8_t* lv[1] = {img};
    ctx_.check(svo_b200_frame_upload(ctx_.get(), dev_, lv, 1));
    ctx_.check(svo_b200_synchronize(ctx_.get()));
  }
  ~Frame() {
    for (Feature* f : fts_) delete f;  // frame.cpp:43-46
    svo_b200_frame_destroy(ctx_.get(), dev_);
  }
  Frame(const Frame&) = delete;
  void addFeature(Feature* ftr) { fts_.push_back(ftr); }
  void setKeyframe() { is_keyframe_ = true; }
  bool isKeyframe() const { return is_keyframe_; }
  Vector3d pos() const { return T_f_w_.inverse().translation(); }  // frame.h:112
  size_t nObs() const { return fts_.size(); }
  svo_b200_frame* device() const { return dev_; }
  Context& context() const { return ctx_; }

 private:
  Context& ctx_;
  svo_b200_frame* dev_ = nullptr;
};
typedef std::shared_ptr<Frame> FramePtr;

inline Feature::Feature(Frame* _frame, const Vector2d& _px, int _level)
    : frame(_frame), px(_px), f(_frame->cam_->cam2world(_px)), level(_level) {}

// ------------------------------------------------------------------------------------------------
// svo::SparseImgAlign (svo/include/svo/sparse_img_align.h:33-81)
// ------------------------------------------------------------------------------------------------
class SparseImgAlign {
 public:
  enum Method { GaussNewton, LevenbergMarquardt };  // [EXT] vk::NLLSSolver::Method; only GaussNewton is used
  SparseImgAlign(int max_level, int min_level, int n_iter, Method method, bool /*display*/, bool /*verbose*/)
      : max_level_(max_level), min_level_(min_level), n_iter_(n_iter) {
    if (method != GaussNewton) throw std::invalid_argument("SparseImgAlign: only GaussNewton is implemented");
    H_.fill(0.0);
  }
  // sparse_img_align.cpp:43-75
  size_t run(FramePtr ref_frame, FramePtr cur_frame) {
    if (ref_frame->fts_.empty()) return 0;  // "SparseImgAlign: no features to track!"
    const size_t n = ref_frame->fts_.size();
    std::vector<double> px(2 * n), f(3 * n), pos(3 * n, 0.0);
    std::vector<uint8_t> has_point(n);
    size_t i = 0;
    for (Feature* ft : ref_frame->fts_) {
      px[2 * i] = ft->px[0]; px[2 * i + 1] = ft->px[1];
      for (int k = 0; k < 3; ++k) f[3 * i + k] = ft->f[k];
      has_point[i] = ft->point != nullptr;
      if (ft->point) for (int k = 0; k < 3; ++k) pos[3 * i + k] = ft->point->pos_[k];
      ++i;
    }
    SE3 T_cur_from_ref = cur_frame->T_f_w_ * ref_frame->T_f_w_.inverse();  // :59
    const Vector3d ref_pos = ref_frame->pos();
    const svo_b200_camera cam = ref_frame->cam_->c_abi();
    const svo_b200_sia_options opt = {max_level_, min_level_, n_iter_, 0.000001};  // eps_ (:40)
    svo_b200_sia_stats st;
    visible_fts_.assign(n, 0);
    Context& c = ref_frame->context();
    c.check(svo_b200_sparse_img_align(c.get(), ref_frame->device(), cur_frame->device(), &cam, &opt, T_cur_from_ref.m,
                                      px.data(), f.data(), pos.data(), has_point.data(), ref_pos.data(), (int)n,
                                      visible_fts_.data(), H_.data(), &st, nullptr, 0, nullptr));
    cur_frame->T_f_w_ = T_cur_from_ref * ref_frame->T_f_w_;  // :70
    return (size_t)st.n_tracked;                              // n_meas_/patch_area_ (:74)
  }
  // sparse_img_align.cpp:77-82
  Matrix6d getFisherInformation() const {
    const double sigma_i_sq = 5e-4 * 255 * 255;
    Matrix6d I;
    for (int k = 0; k < 36; ++k) I[k] = H_[k] / sigma_i_sq;
    return I;
  }
  const std::vector<uint8_t>& visibleFeatures() const { return visible_fts_; }

 private:
  int max_level_, min_level_, n_iter_;
  Matrix6d H_;
  std::vector<uint8_t> visible_fts_;
};

// ------------------------------------------------------------------------------------------------
// svo::pose_optimizer (svo/include/svo/pose_optimizer.h:37-45)
// ------------------------------------------------------------------------------------------------
namespace pose_optimizer {
inline void optimizeGaussNewton(const double reproj_thresh, const size_t n_iter, const bool /*verbose*/, FramePtr& frame,
                                double& estimated_scale, double& error_init, double& error_final, size_t& num_obs) {
  const size_t n = frame->fts_.size();
  std::vector<double> f(3 * n), pos(3 * n, 0.0);
  std::vector<int> level(n);
  std::vector<uint8_t> has_point(n);
  size_t i = 0;
  for (Feature* ft : frame->fts_) {
    for (int k = 0; k < 3; ++k) f[3 * i + k] = ft->f[k];
    level[i] = ft->level;
    has_point[i] = ft->point != nullptr;
    if (ft->point) for (int k = 0; k < 3; ++k) pos[3 * i + k] = ft->point->pos_[k];
    ++i;
  }
  size_t n_with_point = 0;
  for (uint8_t h : has_point) n_with_point += h;
  if (n_with_point == 0) return;  // errors.empty(): outputs untouched (pose_optimizer.cpp:57-58)
  svo_b200_pose_opt_result out;
  Context& c = frame->context();
  c.check(svo_b200_pose_optimize(c.get(), reproj_thresh, (int)n_iter, frame->cam_->errorMultiplier2(), frame->T_f_w_.m,
                                 f.data(), pos.data(), level.data(), has_point.data(), (int)n, &out));
  i = 0;
  for (Feature* ft : frame->fts_) {  // culled observations lose their point (:139-143)
    if (ft->point && !has_point[i]) ft->point = nullptr;
    ++i;
  }
  for (int k = 0; k < 36; ++k) frame->Cov_[k] = out.cov[k];
  estimated_scale = out.estimated_scale;
  error_init = out.error_init;
  error_final = out.error_final;
  num_obs = (size_t)out.num_obs;
}
}  // namespace pose_optimizer

// ------------------------------------------------------------------------------------------------
// svo::feature_alignment (svo/include/svo/feature_alignment.h:29-44).  `cur_img` of the reference is
// (frame, level) here because the pyramid lives on the device.
// ------------------------------------------------------------------------------------------------
namespace feature_alignment {
inline bool align2D(const Frame& cur_frame, int level, uint8_t* ref_patch_with_border, uint8_t* ref_patch, const int n_iter,
                    Vector2d& cur_px_estimate, bool /*no_simd*/ = false) {
  uint8_t conv = 0;
  Context& c = cur_frame.context();
  c.check(svo_b200_align2d_batch(c.get(), cur_frame.device(), 1, &level, ref_patch_with_border, ref_patch, n_iter,
                                 cur_px_estimate.data(), &conv));
  return conv != 0;
}
inline bool align1D(const Frame& cur_frame, int level, const std::array<float, 2>& dir, uint8_t* ref_patch_with_border,
                    uint8_t* ref_patch, const int n_iter, Vector2d& cur_px_estimate, double& h_inv) {
  uint8_t conv = 0;
  Context& c = cur_frame.context();
  c.check(svo_b200_align1d_batch(c.get(), cur_frame.device(), 1, &level, dir.data(), ref_patch_with_border, ref_patch,
                                 n_iter, cur_px_estimate.data(), &conv, &h_inv));
  return conv != 0;
}
}  // namespace feature_alignment

// ------------------------------------------------------------------------------------------------
// svo::DepthFilter (svo/include/svo/depth_filter.h:35-51,53-158)
// ------------------------------------------------------------------------------------------------
// ------------------------------------------------------------------------------------------------
// svo::feature_detection (svo/include/svo/feature_detection.h:28-125): grid bookkeeping on the host, FAST-10 + score +
// non-maximum suppression + Shi-Tomasi + per-cell selection in one device launch (svo_b200_fast_detect).
// ------------------------------------------------------------------------------------------------
namespace feature_detection {
class AbstractDetector {
 public:
  AbstractDetector(int img_width, int img_height, int cell_size, int n_pyr_levels)
      : cell_size_(cell_size), n_pyr_levels_(n_pyr_levels), grid_n_cols_((int)std::ceil((double)img_width / cell_size)),
        grid_n_rows_((int)std::ceil((double)img_height / cell_size)), grid_occupancy_((size_t)grid_n_cols_ * grid_n_rows_, 0) {}
  virtual ~AbstractDetector() {}
  virtual void detect(Frame* frame, const double detection_threshold, Features& fts) = 0;  // img_pyr = the frame's device pyramid
  void setGridOccpuancy(const Vector2d& px) {  // feature_detection.cpp:51-56 (spelling as in the reference)
    grid_occupancy_.at((size_t)((int)(px[1] / cell_size_) * grid_n_cols_ + (int)(px[0] / cell_size_))) = 1;
  }
  void setExistingFeatures(const Features& fts) { for (Feature* f : fts) setGridOccpuancy(f->px); }  // :42-49

 protected:
  const int cell_size_, n_pyr_levels_, grid_n_cols_, grid_n_rows_;
  std::vector<uint8_t> grid_occupancy_;
  void resetGrid() { std::fill(grid_occupancy_.begin(), grid_occupancy_.end(), 0); }
};
typedef std::shared_ptr<AbstractDetector> DetectorPtr;

class FastDetector : public AbstractDetector {
 public:
  FastDetector(int img_width, int img_height, int cell_size, int n_pyr_levels)
      : AbstractDetector(img_width, img_height, cell_size, n_pyr_levels) {}
  void detect(Frame* frame, const double detection_threshold, Features& fts) override {  // feature_detection.cpp:66-115
    const svo_b200_detect_options opt = {cell_size_, n_pyr_levels_, 20, 0, detection_threshold};
    const int cap = (int)grid_occupancy_.size();
    std::vector<int> x(cap), y(cap), level(cap);
    int n = 0;
    Context& c = frame->context();
    c.check(svo_b200_fast_detect(c.get(), frame->device(), &opt, grid_occupancy_.data(), cap, x.data(), y.data(), level.data(),
                                 nullptr, &n));
    for (int i = 0; i < n; ++i) fts.push_back(new Feature(frame, Vector2d{(double)x[i], (double)y[i]}, level[i]));
    resetGrid();
  }
};
}  // namespace feature_detection

struct Seed {
  static int& batch_counter() { static int c = 0; return c; }
  static int& seed_counter() { static int c = 0; return c; }
  int batch_id, id;
  Feature* ftr;
  float a, b, mu, z_range, sigma2;
  Seed(Feature* _ftr, float depth_mean, float depth_min)  // depth_filter.cpp:37-46
      : batch_id(batch_counter()), id(seed_counter()++), ftr(_ftr), a(10), b(10), mu(1.0 / depth_mean),
        z_range(1.0 / depth_min), sigma2(z_range * z_range / 36) {}
};

class DepthFilter {
 public:
  typedef std::function<void(Point*, double)> callback_t;
  struct Options {
    int max_n_kfs = 3;
    double seed_convergence_sigma2_thresh = 200.0;
    int max_search_level = 2;  // Config::nPyrLevels()-1 with the non-ROS default n_pyr_levels = 3
  } options_;
  explicit DepthFilter(callback_t seed_converged_cb) : seed_converged_cb_(seed_converged_cb) {}
  DepthFilter(feature_detection::DetectorPtr feature_detector, callback_t seed_converged_cb)  // depth_filter.h:88-90
      : seed_converged_cb_(seed_converged_cb), feature_detector_(feature_detector) {}
  // addKeyframe + initializeSeeds (depth_filter.cpp:101-132): detect new corners away from the frame's features
  void addKeyframe(FramePtr frame, double depth_mean, double depth_min, double triang_min_corner_score = 20.0) {
    if (!feature_detector_) throw std::runtime_error("DepthFilter: no feature detector (use the overload that takes the features)");
    Features new_features;
    feature_detector_->setExistingFeatures(frame->fts_);
    feature_detector_->detect(frame.get(), triang_min_corner_score, new_features);
    addKeyframe(frame, std::vector<Feature*>(new_features.begin(), new_features.end()), depth_mean, depth_min);
  }
  // the same with the detector's output passed in
  void addKeyframe(FramePtr frame, const std::vector<Feature*>& new_features, double depth_mean, double depth_min) {
    keyframes_.push_back(frame);
    ++Seed::batch_counter();
    for (Feature* ftr : new_features) seeds_.push_back(Seed(ftr, (float)depth_mean, (float)depth_min));
  }
  void addFrame(FramePtr frame) { updateSeeds(frame); }  // synchronous branch (:95-96)
  void removeKeyframe(FramePtr frame) {                  // :134-151
    seeds_.remove_if([&](const Seed& s) { return s.ftr->frame == frame.get(); });
    keyframes_.remove(frame);
  }
  void reset() { seeds_.clear(); keyframes_.clear(); }
  std::list<Seed>& getSeeds() { return seeds_; }
  size_t n_failed_matches_ = 0, n_updates_ = 0;

  // depth_filter.cpp:197-291: one launch for all seeds, then the list side effects in list order
  virtual void updateSeeds(FramePtr frame) {
    const size_t M = seeds_.size();
    if (M == 0) return;
    std::vector<FramePtr> refs(keyframes_.begin(), keyframes_.end());
    std::vector<const svo_b200_frame*> ref_dev(refs.size());
    std::vector<double> ref_T(12 * refs.size());
    for (size_t r = 0; r < refs.size(); ++r) { ref_dev[r] = refs[r]->device(); std::memcpy(&ref_T[12 * r], refs[r]->T_f_w_.m, sizeof(double) * 12); }
    std::vector<int> ref_index(M), level(M), type(M), batch(M);
    std::vector<double> px(2 * M), f(3 * M), grad(2 * M), px_cur(2 * M), z(M);
    std::vector<float> a(M), b(M), mu(M), zr(M), s2(M);
    std::vector<uint8_t> status(M);
    size_t i = 0;
    for (const Seed& s : seeds_) {
      size_t r = 0;
      while (r < refs.size() && refs[r].get() != s.ftr->frame) ++r;
      if (r == refs.size()) throw std::runtime_error("DepthFilter: seed references a frame that is not a keyframe");
      ref_index[i] = (int)r; level[i] = s.ftr->level; type[i] = s.ftr->type; batch[i] = s.batch_id;
      px[2 * i] = s.ftr->px[0]; px[2 * i + 1] = s.ftr->px[1];
      grad[2 * i] = s.ftr->grad[0]; grad[2 * i + 1] = s.ftr->grad[1];
      for (int k = 0; k < 3; ++k) f[3 * i + k] = s.ftr->f[k];
      a[i] = s.a; b[i] = s.b; mu[i] = s.mu; zr[i] = s.z_range; s2[i] = s.sigma2;
      ++i;
    }
    const svo_b200_camera cam = frame->cam_->c_abi();
    const svo_b200_depth_options opt = {options_.max_n_kfs, options_.seed_convergence_sigma2_thresh,
                                        options_.max_search_level, 10, 1000};
    Context& c = frame->context();
    c.check(svo_b200_depth_filter_update(c.get(), ref_dev.data(), ref_T.data(), (int)refs.size(), frame->device(),
                                         frame->T_f_w_.m, &cam, &opt, (int)M, ref_index.data(), px.data(), f.data(),
                                         level.data(), type.data(), grad.data(), batch.data(), Seed::batch_counter(),
                                         a.data(), b.data(), mu.data(), zr.data(), s2.data(), status.data(),
                                         px_cur.data(), z.data(), nullptr));
    i = 0;
    for (auto it = seeds_.begin(); it != seeds_.end(); ++i) {
      it->a = a[i]; it->b = b[i]; it->mu = mu[i]; it->sigma2 = s2[i];
      switch (status[i]) {
        case SVO_B200_SEED_TOO_OLD: it = seeds_.erase(it); continue;                       // :216-219
        case SVO_B200_SEED_NO_MATCH: ++n_failed_matches_; break;                           // :240-244
        case SVO_B200_SEED_CONVERGED: {                                                    // :261-282
          ++n_updates_;
          const SE3 T_w_f = it->ftr->frame->T_f_w_.inverse();
          const double d = 1.0 / it->mu;
          const Vector3d p{it->ftr->f[0] * d, it->ftr->f[1] * d, it->ftr->f[2] * d};
          Vector3d xyz_world;
          for (int r = 0; r < 3; ++r) xyz_world[r] = T_w_f.m[r * 4] * p[0] + T_w_f.m[r * 4 + 1] * p[1] + T_w_f.m[r * 4 + 2] * p[2] + T_w_f.m[r * 4 + 3];
          Point* point = new Point(xyz_world);
          it->ftr->point = point;
          seed_converged_cb_(point, it->sigma2);
          it = seeds_.erase(it);
          continue;
        }
        case SVO_B200_SEED_NAN: ++n_updates_; it = seeds_.erase(it); continue;             // :283-287
        case SVO_B200_SEED_UPDATED: ++n_updates_; break;
        default: break;  // behind the camera / not in frame: untouched
      }
      ++it;
    }
  }
  virtual ~DepthFilter() = default;

 protected:
  callback_t seed_converged_cb_;
  feature_detection::DetectorPtr feature_detector_;
  std::list<Seed> seeds_;
  std::list<FramePtr> keyframes_;
};

// ------------------------------------------------------------------------------------------------
// svo::Map / MapPointCandidates (svo/include/svo/map.h:32-129): the parts Reprojector::reprojectMap touches.
// ------------------------------------------------------------------------------------------------
class MapPointCandidates {
 public:
  typedef std::pair<Point*, Feature*> PointCandidate;
  std::list<PointCandidate> candidates_;
  std::list<Point*> trash_points_;
  ~MapPointCandidates() { reset(); }
  void newCandidatePoint(Point* point, double /*depth_sigma2*/) {  // map.cpp:213-218
    point->type_ = Point::TYPE_CANDIDATE;
    candidates_.push_back(PointCandidate(point, point->obs_.front()));
  }
  void deleteCandidate(PointCandidate& c) {  // map.cpp:280-287
    delete c.second; c.second = nullptr;
    c.first->type_ = Point::TYPE_DELETED;
    trash_points_.push_back(c.first);
  }
  bool deleteCandidatePoint(Point* point) {  // map.cpp:239-252
    for (auto it = candidates_.begin(); it != candidates_.end(); ++it)
      if (it->first == point) { deleteCandidate(*it); candidates_.erase(it); return true; }
    return false;
  }
  void emptyTrash() { for (Point* p : trash_points_) delete p; trash_points_.clear(); }
  void reset() { for (auto& c : candidates_) { delete c.first; delete c.second; } candidates_.clear(); }
};

class Map {
 public:
  std::list<FramePtr> keyframes_;
  std::list<Point*> trash_points_;
  MapPointCandidates point_candidates_;
  void addKeyframe(FramePtr kf) { keyframes_.push_back(kf); }
  void safeDeletePoint(Point* pt) {  // map.cpp:82-99
    for (Feature* ftr : pt->obs_) {
      ftr->point = nullptr;
      for (Feature*& k : ftr->frame->key_pts_) if (k == ftr) k = nullptr;  // Frame::removeKeyPoint without re-selection
    }
    pt->obs_.clear();
    pt->type_ = Point::TYPE_DELETED;
    trash_points_.push_back(pt);
  }
  void emptyTrash() { for (Point* p : trash_points_) delete p; trash_points_.clear(); point_candidates_.emptyTrash(); }
};

// ------------------------------------------------------------------------------------------------
// svo::Reprojector (svo/include/svo/reprojector.h:37-99): reprojectMap gathers the pointer graph into a flat
// svo_b200_map_view, makes ONE device call (projection + speculative alignment of every in-frame point), and applies
// the results the C ABI replayed in the reference's cell order: new Features on the frame, point counters / types,
// safeDeletePoint / deleteCandidatePoint.
// ------------------------------------------------------------------------------------------------
struct ReprojectorOptions {
  size_t max_n_kfs = 10;          // reprojector.h:44
  bool find_match_direct = true;  // reprojector.h:45
  int grid_size = 30, max_fts = 120, n_pyr_levels = 3;  // Config::gridSize(), maxFts(), nPyrLevels()
};
class Reprojector {
 public:
  typedef ReprojectorOptions Options;
  Options options_;
  size_t n_matches_ = 0, n_trials_ = 0;

  Reprojector(PinholeCamera* cam, Map& map, Options opt = Options(), unsigned shuffle_seed = 1) : options_(opt), map_(map) {
    // initializeGrid (reprojector.cpp:47-58); the reference shuffles with rand(), here a seeded LCG Fisher-Yates
    const int cols = (cam->width_ + options_.grid_size - 1) / options_.grid_size, rows = (cam->height_ + options_.grid_size - 1) / options_.grid_size;
    cell_order_.resize((size_t)cols * rows);
    for (size_t i = 0; i < cell_order_.size(); ++i) cell_order_[i] = (int)i;
    uint64_t s = shuffle_seed;
    for (size_t i = cell_order_.size(); i > 1; --i) {
      s = s * 6364136223846793005ULL + 1442695040888963407ULL;
      std::swap(cell_order_[i - 1], cell_order_[(size_t)((s >> 33) % i)]);
    }
  }
  std::vector<int>& cellOrder() { return cell_order_; }

  void reprojectMap(FramePtr frame, std::vector<std::pair<FramePtr, size_t>>& overlap_kfs) {
    n_matches_ = n_trials_ = 0;
    std::vector<FramePtr> kfs(map_.keyframes_.begin(), map_.keyframes_.end());
    std::map<const Frame*, int> kf_index;
    for (size_t k = 0; k < kfs.size(); ++k) kf_index[kfs[k].get()] = (int)k;
    std::map<Feature*, int> ftr_index;
    std::map<Point*, int> pt_index;
    std::vector<Feature*> ftrs;
    std::vector<Point*> pts;
    auto ftr_id = [&](Feature* f) { auto it = ftr_index.find(f); if (it != ftr_index.end()) return it->second; ftr_index[f] = (int)ftrs.size(); ftrs.push_back(f); return (int)ftrs.size() - 1; };
    auto pt_id = [&](Point* p) { auto it = pt_index.find(p); if (it != pt_index.end()) return it->second; pt_index[p] = (int)pts.size(); pts.push_back(p); return (int)pts.size() - 1; };
    std::vector<int> kf_fts_offset(kfs.size() + 1, 0), kf_fts;
    std::vector<double> kf_T(12 * kfs.size()), keypt_pos(15 * kfs.size(), 0.0);
    std::vector<uint8_t> keypt_valid(5 * kfs.size(), 0);
    std::vector<const svo_b200_frame*> kf_dev(kfs.size());
    for (size_t k = 0; k < kfs.size(); ++k) {
      kf_dev[k] = kfs[k]->device();
      std::memcpy(&kf_T[12 * k], kfs[k]->T_f_w_.m, sizeof(double) * 12);
      for (Feature* f : kfs[k]->fts_) { kf_fts.push_back(ftr_id(f)); if (f->point) pt_id(f->point); }
      kf_fts_offset[k + 1] = (int)kf_fts.size();
      for (int i = 0; i < 5; ++i) {
        Feature* kp = kfs[k]->key_pts_[i];
        if (!kp || !kp->point) continue;
        keypt_valid[5 * k + i] = 1;
        for (int c = 0; c < 3; ++c) keypt_pos[3 * (5 * k + i) + c] = kp->point->pos_[c];
      }
    }
    std::vector<int> cand_point;
    for (auto& c : map_.point_candidates_.candidates_) cand_point.push_back(pt_id(c.first));
    std::vector<int> pt_obs_offset(1, 0), pt_obs;
    for (size_t p = 0; p < pts.size(); ++p) {  // obs_ may reach features outside any fts_ list (candidates) -> extends ftrs
      for (Feature* f : pts[p]->obs_) pt_obs.push_back(ftr_id(f));
      pt_obs_offset.push_back((int)pt_obs.size());
    }
    const size_t F = ftrs.size(), P = pts.size();
    std::vector<int> ftr_kf(F), ftr_level(F), ftr_type(F), ftr_point(F), pt_type(P), pt_failed(P), pt_succ(P);
    std::vector<double> ftr_px(2 * F), ftr_f(3 * F), ftr_grad(2 * F), pt_pos(3 * P);
    for (size_t i = 0; i < F; ++i) {
      const Feature* f = ftrs[i];
      auto it = kf_index.find(f->frame);
      if (it == kf_index.end()) throw std::runtime_error("Reprojector: a point is observed from a frame that is not a map keyframe");
      ftr_kf[i] = it->second; ftr_level[i] = f->level; ftr_type[i] = f->type;
      ftr_point[i] = f->point ? pt_index.at(f->point) : -1;
      ftr_px[2 * i] = f->px[0]; ftr_px[2 * i + 1] = f->px[1];
      ftr_grad[2 * i] = f->grad[0]; ftr_grad[2 * i + 1] = f->grad[1];
      for (int c = 0; c < 3; ++c) ftr_f[3 * i + c] = f->f[c];
    }
    for (size_t p = 0; p < P; ++p) {
      pt_type[p] = pts[p]->type_; pt_failed[p] = pts[p]->n_failed_reproj_; pt_succ[p] = pts[p]->n_succeeded_reproj_;
      for (int c = 0; c < 3; ++c) pt_pos[3 * p + c] = pts[p]->pos_[c];
    }
    const svo_b200_map_view view = {(int)kfs.size(), kf_T.data(), keypt_pos.data(), keypt_valid.data(), kf_fts_offset.data(),
                                    kf_fts.data(), (int)F, ftr_kf.data(), ftr_px.data(), ftr_f.data(), ftr_level.data(),
                                    ftr_type.data(), ftr_grad.data(), ftr_point.data(), (int)P, pt_pos.data(),
                                    pt_obs_offset.data(), pt_obs.data(), (int)cand_point.size(), cand_point.data()};
    const svo_b200_reproject_options opt = {options_.grid_size, options_.max_fts, (int)options_.max_n_kfs,
                                            options_.find_match_direct ? 1 : 0, options_.n_pyr_levels - 1, 10};
    const size_t cap = (size_t)options_.max_fts + 1;
    std::vector<uint8_t> action(P);
    std::vector<int> ov_kf(options_.max_n_kfs), new_point(cap), new_level(cap), new_type(cap);
    std::vector<int64_t> ov_count(options_.max_n_kfs);
    std::vector<double> new_px(2 * cap), new_grad(2 * cap);
    svo_b200_reproject_stats st;
    const svo_b200_camera cam = frame->cam_->c_abi();
    Context& c = frame->context();
    c.check(svo_b200_reproject_map(c.get(), &view, kf_dev.data(), frame->device(), frame->T_f_w_.m, &cam, &opt, cell_order_.data(),
                                   pt_type.data(), pt_failed.data(), pt_succ.data(), action.data(), ov_kf.data(), ov_count.data(),
                                   new_point.data(), new_px.data(), new_level.data(), new_type.data(), new_grad.data(), &st));
    n_matches_ = (size_t)st.n_matches;
    n_trials_ = (size_t)st.n_trials;
    overlap_kfs.reserve(options_.max_n_kfs);
    for (int i = 0; i < st.n_overlap; ++i) overlap_kfs.push_back(std::make_pair(kfs[ov_kf[i]], (size_t)ov_count[i]));
    for (int q = 0; q < st.n_new; ++q) {  // reprojector.cpp:183-196
      Feature* nf = new Feature(frame.get(), Vector2d{new_px[2 * q], new_px[2 * q + 1]}, new_level[q]);
      nf->point = pts[new_point[q]];
      if (new_type[q]) { nf->type = Feature::EDGELET; nf->grad = Vector2d{new_grad[2 * q], new_grad[2 * q + 1]}; }
      frame->addFeature(nf);
    }
    for (size_t p = 0; p < P; ++p) {
      pts[p]->n_failed_reproj_ = pt_failed[p];
      pts[p]->n_succeeded_reproj_ = pt_succ[p];
      switch (action[p]) {
        case SVO_B200_PT_SAFE_DELETE: map_.safeDeletePoint(pts[p]); break;
        case SVO_B200_PT_DELETE_CANDIDATE:
        case SVO_B200_PT_CANDIDATE_ERASED: map_.point_candidates_.deleteCandidatePoint(pts[p]); break;
        default: pts[p]->type_ = (Point::PointType)pt_type[p];
      }
    }
  }

 private:
  Map& map_;
  std::vector<int> cell_order_;
};

}  // namespace svo
