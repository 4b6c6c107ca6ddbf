// oracle/oracle_math.h -- TEST INFRASTRUCTURE ONLY (never linked into the product path).
//
// Tiny fixed-size linear algebra that restates the semantics of the un-vendored
// third-party types the reference's hot path leans on ([EXT] in SURVEY.md):
//   * Eigen  : small dense vectors/matrices, LDLT with symmetric pivoting, 2x2/3x3 inverse
//   * Sophus : the old non-templated SE3/SO3 (unit quaternion + translation)
// None of those libraries is available in this container, so every routine here is a
// restatement from the published algorithms.  PARITY UNPINNED at this boundary: no
// reference test pins the numbers produced by these routines.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>

namespace orc {

// ---------------------------------------------------------------- 3-vectors / 3x3
struct V3 {
  double x, y, z;
};
inline V3 v3(double x, double y, double z) { return V3{x, y, z}; }
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(V3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }
inline V3 operator-(V3 a) { return {-a.x, -a.y, -a.z}; }
inline double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 cross(V3 a, V3 b) {
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
inline double norm(V3 a) { return std::sqrt(dot(a, a)); }
inline V3 normalized(V3 a) {
  double n = norm(a);
  return {a.x / n, a.y / n, a.z / n};
}

struct V2 {
  double x, y;
};
inline V2 operator-(V2 a, V2 b) { return {a.x - b.x, a.y - b.y}; }
inline V2 operator+(V2 a, V2 b) { return {a.x + b.x, a.y + b.y}; }
inline V2 operator*(V2 a, double s) { return {a.x * s, a.y * s}; }
inline double norm(V2 a) { return std::sqrt(a.x * a.x + a.y * a.y); }

struct M3 {
  double m[3][3];
};
inline V3 operator*(const M3& A, V3 v) {
  return {A.m[0][0] * v.x + A.m[0][1] * v.y + A.m[0][2] * v.z,
          A.m[1][0] * v.x + A.m[1][1] * v.y + A.m[1][2] * v.z,
          A.m[2][0] * v.x + A.m[2][1] * v.y + A.m[2][2] * v.z};
}
inline M3 operator*(const M3& A, const M3& B) {
  M3 C;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      C.m[i][j] = A.m[i][0] * B.m[0][j] + A.m[i][1] * B.m[1][j] + A.m[i][2] * B.m[2][j];
  return C;
}

// ---------------------------------------------------------------- quaternion / SO3 / SE3
// [EXT] Eigen::Quaterniond (w, x, y, z) as stored by the old Sophus::SO3.
struct Quat {
  double w, x, y, z;
};
inline Quat qmul(Quat a, Quat b) {  // Hamilton product, Eigen operator*
  return {a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z,
          a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
          a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z,
          a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x};
}
inline Quat qnormalized(Quat q) {
  double n = std::sqrt(q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z);
  return {q.w / n, q.x / n, q.y / n, q.z / n};
}
inline Quat qconj(Quat q) { return {q.w, -q.x, -q.y, -q.z}; }
// [EXT] Eigen QuaternionBase::_transformVector: v + 2w (q x v) + 2 q x (q x v)
inline V3 qrotate(Quat q, V3 v) {
  V3 qv{q.x, q.y, q.z};
  V3 uv = cross(qv, v);
  uv = uv + uv;
  return v + uv * q.w + cross(qv, uv);
}
// [EXT] Eigen QuaternionBase::toRotationMatrix
inline M3 qmatrix(Quat q) {
  const double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
  const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
  const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
  const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  M3 R;
  R.m[0][0] = 1 - (tyy + tzz); R.m[0][1] = txy - twz;       R.m[0][2] = txz + twy;
  R.m[1][0] = txy + twz;       R.m[1][1] = 1 - (txx + tzz); R.m[1][2] = tyz - twx;
  R.m[2][0] = txz - twy;       R.m[2][1] = tyz + twx;       R.m[2][2] = 1 - (txx + tyy);
  return R;
}
// [EXT] Eigen quaternion-from-rotation-matrix (Shoemake), as used by Sophus::SO3(Matrix3d)
inline Quat qfrommatrix(const M3& R) {
  Quat q;
  double t = R.m[0][0] + R.m[1][1] + R.m[2][2];
  if (t > 0) {
    t = std::sqrt(t + 1.0);
    q.w = 0.5 * t;
    t = 0.5 / t;
    q.x = (R.m[2][1] - R.m[1][2]) * t;
    q.y = (R.m[0][2] - R.m[2][0]) * t;
    q.z = (R.m[1][0] - R.m[0][1]) * t;
  } else {
    int i = 0;
    if (R.m[1][1] > R.m[0][0]) i = 1;
    if (R.m[2][2] > R.m[i][i]) i = 2;
    int j = (i + 1) % 3, k = (j + 1) % 3;
    t = std::sqrt(R.m[i][i] - R.m[j][j] - R.m[k][k] + 1.0);
    double qv[3];
    qv[i] = 0.5 * t;
    t = 0.5 / t;
    q.w = (R.m[k][j] - R.m[j][k]) * t;
    qv[j] = (R.m[j][i] + R.m[i][j]) * t;
    qv[k] = (R.m[k][i] + R.m[i][k]) * t;
    q.x = qv[0]; q.y = qv[1]; q.z = qv[2];
  }
  return q;
}

// [EXT] old Sophus::SE3 = { SO3 (unit quaternion), translation }.
struct SE3 {
  Quat q{1, 0, 0, 0};
  V3 t{0, 0, 0};
};
inline V3 operator*(const SE3& T, V3 p) { return qrotate(T.q, p) + T.t; }  // so3_*p + t
// SE3::operator*: result.translation_ += so3_*other.translation_; so3_ *= other.so3_ (+normalize)
inline SE3 operator*(const SE3& A, const SE3& B) {
  SE3 C;
  C.t = A.t + qrotate(A.q, B.t);
  C.q = qnormalized(qmul(A.q, B.q));
  return C;
}
inline SE3 inverse(const SE3& T) {
  SE3 I;
  I.q = qconj(T.q);
  I.t = qrotate(I.q, T.t * -1.0);
  return I;
}
inline M3 rotation_matrix(const SE3& T) { return qmatrix(T.q); }

inline M3 hat(V3 w) {
  M3 O;
  O.m[0][0] = 0;    O.m[0][1] = -w.z; O.m[0][2] = w.y;
  O.m[1][0] = w.z;  O.m[1][1] = 0;    O.m[1][2] = -w.x;
  O.m[2][0] = -w.y; O.m[2][1] = w.x;  O.m[2][2] = 0;
  return O;
}

// [EXT] Sophus::SO3::expAndTheta + SE3::exp (tangent = [upsilon, omega], translation first).
inline SE3 se3_exp(const double u[6]) {
  const double SMALL_EPS = 1e-10;
  V3 upsilon{u[0], u[1], u[2]}, omega{u[3], u[4], u[5]};
  const double theta = norm(omega);
  const double half_theta = 0.5 * theta;
  double imag_factor;
  const double real_factor = std::cos(half_theta);
  if (theta < SMALL_EPS) {
    const double theta_sq = theta * theta;
    const double theta_po4 = theta_sq * theta_sq;
    imag_factor = 0.5 - 0.0208333 * theta_sq + 0.000260417 * theta_po4;
  } else {
    imag_factor = std::sin(half_theta) / theta;
  }
  SE3 T;
  T.q = Quat{real_factor, imag_factor * omega.x, imag_factor * omega.y, imag_factor * omega.z};
  const M3 Omega = hat(omega);
  const M3 Omega_sq = Omega * Omega;
  M3 V;
  if (theta < SMALL_EPS) {
    V = qmatrix(T.q);
  } else {
    const double theta_sq = theta * theta;
    const double a = (1 - std::cos(theta)) / theta_sq;
    const double b = (theta - std::sin(theta)) / (theta_sq * theta);
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j)
        V.m[i][j] = (i == j ? 1.0 : 0.0) + a * Omega.m[i][j] + b * Omega_sq.m[i][j];
  }
  T.t = V * upsilon;
  return T;
}

// Row-major 3x4 [R|t] <-> SE3 (the layout used at the C boundary of both oracle and product).
inline SE3 se3_from_rt12(const double* T12) {
  M3 R;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) R.m[i][j] = T12[i * 4 + j];
  SE3 T;
  T.q = qnormalized(qfrommatrix(R));
  T.t = V3{T12[3], T12[7], T12[11]};
  return T;
}
inline void se3_to_rt12(const SE3& T, double* T12) {
  M3 R = qmatrix(T.q);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) T12[i * 4 + j] = R.m[i][j];
  T12[3] = T.t.x; T12[7] = T.t.y; T12[11] = T.t.z;
}

// ---------------------------------------------------------------- LDLT (Eigen semantics) [EXT]
// Restates Eigen::LDLT<Matrix<double,N,N>, Lower>: symmetric pivoting on the largest remaining
// |diagonal|, unblocked in-place factorisation, and a solve that uses the pseudo-inverse of D
// (|d_i| <= 1/highest() -> 0).  Only the lower triangle of A is referenced.
template <int N>
struct LDLT {
  double m[N][N];
  int tr[N];
  void compute(const double A[N][N]) {
    for (int i = 0; i < N; ++i)
      for (int j = 0; j < N; ++j) m[i][j] = (j <= i) ? A[i][j] : 0.0;
    for (int k = 0; k < N; ++k) {
      int big = k;
      double bigv = std::fabs(m[k][k]);
      for (int i = k + 1; i < N; ++i)
        if (std::fabs(m[i][i]) > bigv) { bigv = std::fabs(m[i][i]); big = i; }
      tr[k] = big;
      if (big != k) {
        const int s = N - big - 1;
        for (int j = 0; j < k; ++j) std::swap(m[k][j], m[big][j]);
        for (int i = 0; i < s; ++i) std::swap(m[big + 1 + i][k], m[big + 1 + i][big]);
        std::swap(m[k][k], m[big][big]);
        for (int i = k + 1; i < big; ++i) std::swap(m[i][k], m[big][i]);
      }
      const int rs = N - k - 1;
      if (k > 0) {
        double temp[N];
        for (int j = 0; j < k; ++j) temp[j] = m[j][j] * m[k][j];
        double acc = 0;
        for (int j = 0; j < k; ++j) acc += m[k][j] * temp[j];
        m[k][k] -= acc;
        for (int i = 0; i < rs; ++i) {
          double a2 = 0;
          for (int j = 0; j < k; ++j) a2 += m[k + 1 + i][j] * temp[j];
          m[k + 1 + i][k] -= a2;
        }
      }
      const double akk = m[k][k];
      const bool pivot_ok = std::fabs(akk) > 0.0;
      if (k == 0 && !pivot_ok) {
        for (int j = 0; j < N; ++j) tr[j] = j;
        return;
      }
      if (rs > 0 && pivot_ok)
        for (int i = 0; i < rs; ++i) m[k + 1 + i][k] /= akk;
    }
  }
  void solve(const double b[N], double x[N]) const {
    for (int i = 0; i < N; ++i) x[i] = b[i];
    for (int i = 0; i < N; ++i) std::swap(x[i], x[tr[i]]);            // P b
    for (int i = 0; i < N; ++i)                                         // L^-1
      for (int j = 0; j < i; ++j) x[i] -= m[i][j] * x[j];
    const double tol = 1.0 / std::numeric_limits<double>::max();
    for (int i = 0; i < N; ++i) {                                       // D^+
      if (std::fabs(m[i][i]) > tol) x[i] /= m[i][i];
      else x[i] = 0.0;
    }
    for (int i = N - 1; i >= 0; --i)                                    // L^-T
      for (int j = i + 1; j < N; ++j) x[i] -= m[j][i] * x[j];
    for (int i = N - 1; i >= 0; --i) std::swap(x[i], x[tr[i]]);       // P^-1
  }
};

// [EXT] Eigen fixed-size inverse for 2x2 / 3x3 float (cofactor formulas, Eigen compute_inverse).
inline void inv2f(const float A[2][2], float I[2][2]) {
  const float invdet = 1.0f / (A[0][0] * A[1][1] - A[1][0] * A[0][1]);
  I[0][0] = A[1][1] * invdet;  I[0][1] = -A[0][1] * invdet;
  I[1][0] = -A[1][0] * invdet; I[1][1] = A[0][0] * invdet;
}
inline void inv3f(const float A[3][3], float I[3][3]) {
  // cofactors of the first column -> determinant, then the full adjugate * (1/det)
  const float c00 = A[1][1] * A[2][2] - A[1][2] * A[2][1];
  const float c10 = A[1][2] * A[2][0] - A[1][0] * A[2][2];  // cofactor(0,1) laid at row 1
  const float c20 = A[1][0] * A[2][1] - A[1][1] * A[2][0];
  const float det = A[0][0] * c00 + A[0][1] * c10 + A[0][2] * c20;
  const float invdet = 1.0f / det;
  I[0][0] = c00 * invdet;
  I[1][0] = c10 * invdet;
  I[2][0] = c20 * invdet;
  I[0][1] = (A[0][2] * A[2][1] - A[0][1] * A[2][2]) * invdet;
  I[1][1] = (A[0][0] * A[2][2] - A[0][2] * A[2][0]) * invdet;
  I[2][1] = (A[0][1] * A[2][0] - A[0][0] * A[2][1]) * invdet;
  I[0][2] = (A[0][1] * A[1][2] - A[0][2] * A[1][1]) * invdet;
  I[1][2] = (A[0][2] * A[1][0] - A[0][0] * A[1][2]) * invdet;
  I[2][2] = (A[0][0] * A[1][1] - A[0][1] * A[1][0]) * invdet;
}
inline void inv2d(const double A[2][2], double I[2][2]) {
  const double invdet = 1.0 / (A[0][0] * A[1][1] - A[1][0] * A[0][1]);
  I[0][0] = A[1][1] * invdet;  I[0][1] = -A[0][1] * invdet;
  I[1][0] = -A[1][0] * invdet; I[1][1] = A[0][0] * invdet;
}

}  // namespace orc
