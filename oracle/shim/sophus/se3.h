// oracle/shim/sophus/se3.h -- TEST INFRASTRUCTURE ONLY.
// Stand-in for the old non-templated Sophus::SE3 (unit quaternion + translation) with exactly the arithmetic of
// oracle_math.h (restated from the published Sophus sources, [EXT]): used only to compile the reference's own
// files for oracle/_ref.
#pragma once
#include <Eigen/Core>
#include <algorithm>
#include <iostream>
#include <list>
#include <vector>
#include "../../oracle_math.h"

namespace Sophus {
using namespace Eigen;
using namespace std;
typedef Matrix<double, 6, 1> Vector6d;
typedef Matrix<double, 6, 6> Matrix6d;

class SE3 {
 public:
  orc::SE3 T;
  SE3() {}
  explicit SE3(const orc::SE3& t) : T(t) {}
  SE3(const Matrix3d& R, const Vector3d& t) {
    orc::M3 m;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) m.m[i][j] = R(i, j);
    T.q = orc::qnormalized(orc::qfrommatrix(m));
    T.t = orc::V3{t[0], t[1], t[2]};
  }
  SE3 operator*(const SE3& o) const { return SE3(T * o.T); }
  Vector3d operator*(const Vector3d& p) const { const orc::V3 r = T * orc::V3{p[0], p[1], p[2]}; return Vector3d(r.x, r.y, r.z); }
  SE3 inverse() const { return SE3(orc::inverse(T)); }
  Matrix3d rotation_matrix() const {
    const orc::M3 m = orc::qmatrix(T.q);
    Matrix3d R;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) R(i, j) = m.m[i][j];
    return R;
  }
  Vector3d translation() const { return Vector3d(T.t.x, T.t.y, T.t.z); }
  static SE3 exp(const Vector6d& u) { double x[6]; for (int i = 0; i < 6; ++i) x[i] = u[i]; return SE3(orc::se3_exp(x)); }
};
}  // namespace Sophus
