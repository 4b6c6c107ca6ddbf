// oracle/shim/vikit/abstract_camera.h -- TEST INFRASTRUCTURE ONLY: [EXT] vk::AbstractCamera, vk::PinholeCamera (+ radial-tangential distortion), vk::ATANCamera.
#pragma once
#include <Eigen/Core>
#include <cmath>
namespace vk {
using namespace Eigen;
class AbstractCamera {
 protected:
  int width_, height_;
 public:
  AbstractCamera(int w, int h) : width_(w), height_(h) {}
  virtual ~AbstractCamera() {}
  virtual Vector3d cam2world(const double& x, const double& y) const = 0;
  virtual Vector3d cam2world(const Vector2d& px) const = 0;
  virtual Vector2d world2cam(const Vector3d& xyz_c) const = 0;
  virtual Vector2d world2cam(const Vector2d& uv) const = 0;
  virtual double errorMultiplier2() const = 0;
  virtual double errorMultiplier() const = 0;
  inline int width() const { return width_; }
  inline int height() const { return height_; }
  inline bool isInFrame(const Vector2i& obs, int boundary = 0) const {
    return obs[0] >= boundary && obs[0] < width() - boundary && obs[1] >= boundary && obs[1] < height() - boundary;
  }
  inline bool isInFrame(const Vector2i& obs, int boundary, int level) const {
    return obs[0] >= boundary && obs[0] < width() / (1 << level) - boundary && obs[1] >= boundary &&
           obs[1] < height() / (1 << level) - boundary;
  }
};
// [EXT] vk::PinholeCamera (rpg_vikit pinhole_camera.h/.cpp): optional radial-tangential distortion d0..d4 (k1 k2 p1 p2 k3);
// cam2world with distortion = cv::undistortPoints on one CV_32FC2 point [EXT OpenCV 2.4 cvUndistortPoints]: float in,
// five fixed-point iterations in double, float out.
class PinholeCamera : public AbstractCamera {
  double fx_, fy_, cx_, cy_;
  bool distortion_;
  double d_[5];
 public:
  PinholeCamera(int w, int h, double fx, double fy, double cx, double cy, double d0 = 0.0, double d1 = 0.0, double d2 = 0.0,
                double d3 = 0.0, double d4 = 0.0)
      : AbstractCamera(w, h), fx_(fx), fy_(fy), cx_(cx), cy_(cy), distortion_(std::fabs(d0) > 0.0000001) {
    d_[0] = d0; d_[1] = d1; d_[2] = d2; d_[3] = d3; d_[4] = d4;
  }
  Vector3d cam2world(const double& u, const double& v) const override {
    Vector3d xyz;
    if (!distortion_) {
      xyz[0] = (u - cx_) / fx_; xyz[1] = (v - cy_) / fy_; xyz[2] = 1.0;
    } else {
      const float uf = (float)u, vf = (float)v;  // cv::Point2f uv(u, v)
      const double ifx = 1. / fx_, ify = 1. / fy_;
      double x, y, x0, y0;
      x = uf; y = vf;
      x0 = x = (x - cx_) * ifx;
      y0 = y = (y - cy_) * ify;
      for (int j = 0; j < 5; j++) {
        double r2 = x * x + y * y;
        double icdist = 1. / (1 + ((d_[4] * r2 + d_[1]) * r2 + d_[0]) * r2);
        double deltaX = 2 * d_[2] * x * y + d_[3] * (r2 + 2 * x * x);
        double deltaY = d_[2] * (r2 + 2 * y * y) + 2 * d_[3] * x * y;
        x = (x0 - deltaX) * icdist;
        y = (y0 - deltaY) * icdist;
      }
      const float pxx = (float)x, pxy = (float)y;  // dst is CV_32FC2
      xyz[0] = pxx; xyz[1] = pxy; xyz[2] = 1.0;
    }
    return xyz.normalized();
  }
  Vector3d cam2world(const Vector2d& px) const override { return cam2world(px[0], px[1]); }
  Vector2d world2cam(const Vector3d& xyz_c) const override { return world2cam(Vector2d(xyz_c[0] / xyz_c[2], xyz_c[1] / xyz_c[2])); }
  Vector2d world2cam(const Vector2d& uv) const override {
    Vector2d px;
    if (!distortion_) {
      px[0] = fx_ * uv[0] + cx_;
      px[1] = fy_ * uv[1] + cy_;
    } else {
      double x, y, r2, r4, r6, a1, a2, a3, cdist, xd, yd;
      x = uv[0];
      y = uv[1];
      r2 = x * x + y * y;
      r4 = r2 * r2;
      r6 = r4 * r2;
      a1 = 2 * x * y;
      a2 = r2 + 2 * x * x;
      a3 = r2 + 2 * y * y;
      cdist = 1 + d_[0] * r2 + d_[1] * r4 + d_[4] * r6;
      xd = x * cdist + d_[2] * a1 + d_[3] * a2;
      yd = y * cdist + d_[2] * a3 + d_[3] * a1;
      px[0] = xd * fx_ + cx_;
      px[1] = yd * fy_ + cy_;
    }
    return px;
  }
  double errorMultiplier2() const override { return std::fabs(fx_); }
  double errorMultiplier() const override { return std::fabs(4.0 * fx_ * fy_); }
};
// [EXT] vk::ATANCamera (rpg_vikit atan_camera.h/.cpp; PTAM's FOV model).  The constructor takes the NORMALISED
// parameters of camera_atan.yaml and derives the pixel ones.
class ATANCamera : public AbstractCamera {
  double fx_, fy_, fx_inv_, fy_inv_, cx_, cy_, s_, s_inv_, tans_, tans_inv_;
  bool distortion_;
  double rtrans_factor(double r) const {
    if (r < 0.001 || s_ == 0.0) return 1.0;
    return (s_inv_ * std::atan(r * tans_) / r);
  }
  double invrtrans(double r) const {
    if (s_ == 0.0) return r;
    return (std::tan(r * s_) * tans_inv_);
  }
 public:
  ATANCamera(double width, double height, double fx, double fy, double cx, double cy, double s)
      : AbstractCamera((int)width, (int)height), fx_(width * fx), fy_(height * fy), fx_inv_(1.0 / fx_), fy_inv_(1.0 / fy_),
        cx_(cx * width - 0.5), cy_(cy * height - 0.5), s_(s), s_inv_(1.0 / s_) {
    if (s_ != 0.0) {
      tans_ = 2.0 * std::tan(s_ / 2.0);
      tans_inv_ = 1.0 / tans_;
      s_inv_ = 1.0 / s_;
      distortion_ = true;
    } else {
      s_inv_ = 0.0;
      tans_ = 0.0;
      tans_inv_ = 0.0;
      distortion_ = false;
    }
  }
  Vector3d cam2world(const double& x, const double& y) const override {
    Vector2d dist_cam((x - cx_) * fx_inv_, (y - cy_) * fy_inv_);
    double dist_r = dist_cam.norm();
    double r = invrtrans(dist_r);
    double d_factor;
    if (dist_r > 0.01) d_factor = r / dist_r;
    else d_factor = 1.0;
    Vector2d uv = d_factor * dist_cam;
    return Vector3d(uv[0], uv[1], 1.0).normalized();
  }
  Vector3d cam2world(const Vector2d& px) const override { return cam2world(px[0], px[1]); }
  Vector2d world2cam(const Vector3d& xyz_c) const override { return world2cam(Vector2d(xyz_c[0] / xyz_c[2], xyz_c[1] / xyz_c[2])); }
  Vector2d world2cam(const Vector2d& uv) const override {
    double r = uv.norm();
    double factor = rtrans_factor(r);
    Vector2d dist_cam = factor * uv;
    return Vector2d(cx_ + fx_ * dist_cam[0], cy_ + fy_ * dist_cam[1]);
  }
  double errorMultiplier2() const override { return std::fabs(fx_); }
  double errorMultiplier() const override { return std::fabs(4.0 * fx_ * fy_); }
};
}  // namespace vk
