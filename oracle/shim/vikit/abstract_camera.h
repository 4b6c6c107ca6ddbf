// oracle/shim/vikit/abstract_camera.h -- TEST INFRASTRUCTURE ONLY: [EXT] vk::AbstractCamera / PinholeCamera (no distortion).
#pragma once
#include <Eigen/Core>
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
class PinholeCamera : public AbstractCamera {
  double fx_, fy_, cx_, cy_;
 public:
  PinholeCamera(int w, int h, double fx, double fy, double cx, double cy) : AbstractCamera(w, h), fx_(fx), fy_(fy), cx_(cx), cy_(cy) {}
  Vector3d cam2world(const double& u, const double& v) const override {
    Vector3d xyz;
    xyz[0] = (u - cx_) / fx_; xyz[1] = (v - cy_) / fy_; xyz[2] = 1.0;
    return xyz.normalized();
  }
  Vector3d cam2world(const Vector2d& px) const override { return cam2world(px[0], px[1]); }
  Vector2d world2cam(const Vector3d& xyz_c) const override { return world2cam(Vector2d(xyz_c[0] / xyz_c[2], xyz_c[1] / xyz_c[2])); }
  Vector2d world2cam(const Vector2d& uv) const override { Vector2d px; px[0] = fx_ * uv[0] + cx_; px[1] = fy_ * uv[1] + cy_; return px; }
  double errorMultiplier2() const override { return std::fabs(fx_); }
  double errorMultiplier() const override { return std::fabs(4.0 * fx_ * fy_); }
};
}  // namespace vk
