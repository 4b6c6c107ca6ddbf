// rpg_svo_b200/csrc/svo_math.cuh -- device-side small math for the sm_100a kernels.
//
// SE3 as unit quaternion + translation (the storage the reference's Sophus::SE3 uses, so the two
// sides drift the same way), Rodrigues/expmap, a 6x6 LDL^T with symmetric pivoting (the behaviour
// of Eigen::LDLT that svo::SparseImgAlign::solve and pose_optimizer rely on:
// svo/src/sparse_img_align.cpp:245-251, svo/src/pose_optimizer.cpp:97), and block reductions.
// All files in csrc/ are compiled with -fmad=false: every fused multiply-add below is explicit.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/svo_b200.h"

namespace svo {

struct Quat {
  double w, x, y, z;
};
struct Pose {  // T = [R(q) | t]
  Quat q;
  double t[3];
};

__device__ __forceinline__ Quat qmul(const Quat& a, const Quat& b) {
  Quat r;
  r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
  r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
  r.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
  r.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
  return r;
}
__device__ __forceinline__ Quat qnormalized(const Quat& q) {
  const double n = sqrt(q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z);
  Quat r = {q.w / n, q.x / n, q.y / n, q.z / n};
  return r;
}
__device__ __forceinline__ void qmatrix(const Quat& q, double* R /*9 row-major*/) {
  const double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
  const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
  const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
  const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  R[0] = 1 - (tyy + tzz); R[1] = txy - twz;       R[2] = txz + twy;
  R[3] = txy + twz;       R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy;       R[7] = tyz + twx;       R[8] = 1 - (txx + tyy);
}
__device__ inline Quat qfrommatrix(const double* R /*9 row-major*/) {
  Quat q;
  double t = R[0] + R[4] + R[8];
  if (t > 0) {
    t = sqrt(t + 1.0);
    q.w = 0.5 * t;
    t = 0.5 / t;
    q.x = (R[7] - R[5]) * t;
    q.y = (R[2] - R[6]) * t;
    q.z = (R[3] - R[1]) * t;
  } else if (R[0] >= R[4] && R[0] >= R[8]) {
    t = sqrt(R[0] - R[4] - R[8] + 1.0);
    q.x = 0.5 * t;
    t = 0.5 / t;
    q.w = (R[7] - R[5]) * t;
    q.y = (R[3] + R[1]) * t;
    q.z = (R[6] + R[2]) * t;
  } else if (R[4] >= R[8]) {
    t = sqrt(R[4] - R[8] - R[0] + 1.0);
    q.y = 0.5 * t;
    t = 0.5 / t;
    q.w = (R[2] - R[6]) * t;
    q.z = (R[7] + R[5]) * t;
    q.x = (R[1] + R[3]) * t;
  } else {
    t = sqrt(R[8] - R[0] - R[4] + 1.0);
    q.z = 0.5 * t;
    t = 0.5 / t;
    q.w = (R[3] - R[1]) * t;
    q.x = (R[6] + R[2]) * t;
    q.y = (R[7] + R[5]) * t;
  }
  return qnormalized(q);
}
__device__ __forceinline__ void qrotate(const Quat& q, const double* v, double* out) {
  // v + 2w (qv x v) + 2 qv x (qv x v)
  double ux = q.y * v[2] - q.z * v[1], uy = q.z * v[0] - q.x * v[2], uz = q.x * v[1] - q.y * v[0];
  ux += ux; uy += uy; uz += uz;
  out[0] = v[0] + q.w * ux + (q.y * uz - q.z * uy);
  out[1] = v[1] + q.w * uy + (q.z * ux - q.x * uz);
  out[2] = v[2] + q.w * uz + (q.x * uy - q.y * ux);
}

// SE3 exponential, tangent = [upsilon, omega] (translation first).
__device__ inline Pose se3_exp(const double* u) {
  const double wx = u[3], wy = u[4], wz = u[5];
  const double th2 = wx * wx + wy * wy + wz * wz;
  const double th = sqrt(th2);
  double imag, real, a, b;  // a = (1-cos)/th^2, b = (th - sin)/th^3
  if (th < 1e-10) {
    real = 1.0 - 0.125 * th2;
    imag = 0.5 - th2 * (1.0 / 48.0);
    a = 0.5;
    b = 1.0 / 6.0;
  } else {
    double sh, ch;
    sincos(0.5 * th, &sh, &ch);
    real = ch;
    imag = sh / th;
    if (th < 1e-2) {  // series keep full precision where the closed forms cancel
      a = 0.5 - th2 * (1.0 / 24.0) * (1.0 - th2 * (1.0 / 30.0) * (1.0 - th2 * (1.0 / 56.0)));
      b = (1.0 / 6.0) - th2 * (1.0 / 120.0) * (1.0 - th2 * (1.0 / 42.0) * (1.0 - th2 * (1.0 / 72.0)));
    } else {
      double s, c;
      sincos(th, &s, &c);
      a = (1.0 - c) / th2;
      b = (th - s) / (th2 * th);
    }
  }
  Pose P;
  P.q.w = real; P.q.x = imag * wx; P.q.y = imag * wy; P.q.z = imag * wz;
  // V = I + a*W + b*W^2,  t = V * upsilon ;  W v = w x v
  const double cx = wy * u[2] - wz * u[1], cy = wz * u[0] - wx * u[2], cz = wx * u[1] - wy * u[0];
  const double ccx = wy * cz - wz * cy, ccy = wz * cx - wx * cz, ccz = wx * cy - wy * cx;
  P.t[0] = u[0] + a * cx + b * ccx;
  P.t[1] = u[1] + a * cy + b * ccy;
  P.t[2] = u[2] + a * cz + b * ccz;
  return P;
}

__device__ inline Pose pose_mul(const Pose& A, const Pose& B) {
  Pose C;
  double r[3];
  qrotate(A.q, B.t, r);
  C.t[0] = A.t[0] + r[0]; C.t[1] = A.t[1] + r[1]; C.t[2] = A.t[2] + r[2];
  C.q = qnormalized(qmul(A.q, B.q));
  return C;
}
__device__ inline Pose pose_inv(const Pose& A) {
  Pose I;
  I.q.w = A.q.w; I.q.x = -A.q.x; I.q.y = -A.q.y; I.q.z = -A.q.z;
  double nt[3] = {-A.t[0], -A.t[1], -A.t[2]};
  qrotate(I.q, nt, I.t);
  return I;
}
__device__ inline Pose pose_from_rt12(const double* T) {
  double R[9] = {T[0], T[1], T[2], T[4], T[5], T[6], T[8], T[9], T[10]};
  Pose P;
  P.q = qfrommatrix(R);
  P.t[0] = T[3]; P.t[1] = T[7]; P.t[2] = T[11];
  return P;
}
__device__ inline void pose_to_rt12(const Pose& P, double* T) {
  double R[9];
  qmatrix(P.q, R);
  T[0] = R[0]; T[1] = R[1]; T[2] = R[2];  T[3] = P.t[0];
  T[4] = R[3]; T[5] = R[4]; T[6] = R[5];  T[7] = P.t[1];
  T[8] = R[6]; T[9] = R[7]; T[10] = R[8]; T[11] = P.t[2];
}

// ------------------------------------------------------------------------------------------
// 6x6 LDL^T with symmetric pivoting (largest remaining |diagonal|), in place on a row-major
// 6x6 buffer `m` (lower triangle used) that may live in shared memory, plus a solve with the
// pseudo-inverse of D.  One thread executes these; they run once per pyramid level.
// ------------------------------------------------------------------------------------------
static __device__ __noinline__ void ldlt6_factor(double* m, int* tr) {
  for (int k = 0; k < 6; ++k) {
    int big = k;
    double bigv = fabs(m[k * 6 + k]);
    for (int i = k + 1; i < 6; ++i)
      if (fabs(m[i * 6 + i]) > bigv) { bigv = fabs(m[i * 6 + i]); big = i; }
    tr[k] = big;
    if (big != k) {
      for (int j = 0; j < k; ++j) { double s = m[k * 6 + j]; m[k * 6 + j] = m[big * 6 + j]; m[big * 6 + j] = s; }
      for (int i = big + 1; i < 6; ++i) { double s = m[i * 6 + k]; m[i * 6 + k] = m[i * 6 + big]; m[i * 6 + big] = s; }
      { double s = m[k * 6 + k]; m[k * 6 + k] = m[big * 6 + big]; m[big * 6 + big] = s; }
      for (int i = k + 1; i < big; ++i) { double s = m[i * 6 + k]; m[i * 6 + k] = m[big * 6 + i]; m[big * 6 + i] = s; }
    }
    if (k > 0) {
      double temp[6];
      double acc = 0;
      for (int j = 0; j < k; ++j) {
        temp[j] = m[j * 6 + j] * m[k * 6 + j];
        acc += m[k * 6 + j] * temp[j];
      }
      m[k * 6 + k] -= acc;
      for (int i = k + 1; i < 6; ++i) {
        double a2 = 0;
        for (int j = 0; j < k; ++j) a2 += m[i * 6 + j] * temp[j];
        m[i * 6 + k] -= a2;
      }
    }
    const double akk = m[k * 6 + k];
    const bool ok = fabs(akk) > 0.0;
    if (k == 0 && !ok) {
      for (int j = 0; j < 6; ++j) tr[j] = j;
      return;
    }
    if (ok)
      for (int i = k + 1; i < 6; ++i) m[i * 6 + k] /= akk;
  }
}
// x (in/out) holds b on entry.  x may be in shared memory (dynamic indexing).
static __device__ __noinline__ void ldlt6_solve(const double* m, const int* tr, double* x) {
  for (int i = 0; i < 6; ++i) { const int j = tr[i]; double s = x[i]; x[i] = x[j]; x[j] = s; }
  for (int i = 1; i < 6; ++i) {
    double s = x[i];
    for (int j = 0; j < i; ++j) s = fma(-m[i * 6 + j], x[j], s);
    x[i] = s;
  }
  for (int i = 0; i < 6; ++i) {
    const double d = m[i * 6 + i];
    x[i] = (fabs(d) > 5.562684646268003e-309) ? x[i] / d : 0.0;  // 1/DBL_MAX
  }
  for (int i = 4; i >= 0; --i) {
    double s = x[i];
    for (int j = i + 1; j < 6; ++j) s = fma(-m[j * 6 + i], x[j], s);
    x[i] = s;
  }
  for (int i = 5; i >= 0; --i) { const int j = tr[i]; double s = x[i]; x[i] = x[j]; x[j] = s; }
}


// ------------------------------------------------------------------------------------------
// Latency-tuned variants used on the serial critical path of the Gauss-Newton loops (one thread
// works while the CTA waits): explicit FMAs, no divisions, no sincos for small rotations.
// Results agree with the plain versions to a few ulp.
// ------------------------------------------------------------------------------------------
static __device__ __noinline__ void se3_exp_coeffs_large(double th2, double& imag, double& real, double& a, double& b) {
  const double th = sqrt(th2);
  double sh, ch, s, c;
  sincos(0.5 * th, &sh, &ch);
  sincos(th, &s, &c);
  real = ch;
  imag = sh / th;
  a = (1.0 - c) / th2;
  b = (th - s) / (th2 * th);
}
__device__ __forceinline__ Pose se3_exp_fast(const double* u) {
  const double wx = u[3], wy = u[4], wz = u[5];
  const double th2 = fma(wx, wx, fma(wy, wy, wz * wz));
  double imag, real, a, b;
  if (th2 < 0.25) {
    // Taylor series in th2 (|th| < 0.5: the 8th term is < 1e-17 of the leading one)
    // imag = sin(th/2)/th, real = cos(th/2), a = (1-cos th)/th^2, b = (th - sin th)/th^3
    const double q2 = 0.25 * th2;  // (th/2)^2
    imag = 0.5 * fma(q2, fma(q2, fma(q2, fma(q2, fma(q2, fma(q2, fma(q2, -1.0 / 1307674368000.0, 1.0 / 6227020800.0),
                     -1.0 / 39916800.0), 1.0 / 362880.0), -1.0 / 5040.0), 1.0 / 120.0), -1.0 / 6.0), 1.0);
    real = fma(q2, fma(q2, fma(q2, fma(q2, fma(q2, fma(q2, fma(q2, -1.0 / 87178291200.0, 1.0 / 479001600.0),
               -1.0 / 3628800.0), 1.0 / 40320.0), -1.0 / 720.0), 1.0 / 24.0), -0.5), 1.0);
    a = fma(th2, fma(th2, fma(th2, fma(th2, fma(th2, fma(th2, fma(th2, -1.0 / 20922789888000.0, 1.0 / 87178291200.0),
            -1.0 / 479001600.0), 1.0 / 3628800.0), -1.0 / 40320.0), 1.0 / 720.0), -1.0 / 24.0), 0.5);
    b = fma(th2, fma(th2, fma(th2, fma(th2, fma(th2, fma(th2, fma(th2, -1.0 / 355687428096000.0, 1.0 / 1307674368000.0),
            -1.0 / 6227020800.0), 1.0 / 39916800.0), -1.0 / 362880.0), 1.0 / 5040.0), -1.0 / 120.0), 1.0 / 6.0);
  } else {
    se3_exp_coeffs_large(th2, imag, real, a, b);  // |omega| >= 0.5 rad in one update: cold, kept out of line
  }
  Pose P;
  P.q.w = real; P.q.x = imag * wx; P.q.y = imag * wy; P.q.z = imag * wz;
  const double cx = fma(wy, u[2], -(wz * u[1])), cy = fma(wz, u[0], -(wx * u[2])), cz = fma(wx, u[1], -(wy * u[0]));
  const double ccx = fma(wy, cz, -(wz * cy)), ccy = fma(wz, cx, -(wx * cz)), ccz = fma(wx, cy, -(wy * cx));
  P.t[0] = fma(b, ccx, fma(a, cx, u[0]));
  P.t[1] = fma(b, ccy, fma(a, cy, u[1]));
  P.t[2] = fma(b, ccz, fma(a, cz, u[2]));
  return P;
}

// C = A * B with the product quaternion re-normalised by one Newton step of 1/sqrt at 1
// (both inputs are unit to ~1e-16, so the step is exact to double precision).
__device__ __forceinline__ Pose pose_mul_fast(const Pose& A, const Pose& B) {
  Pose C;
  const Quat &a = A.q, &b = B.q;
  Quat r;
  r.w = fma(a.w, b.w, -fma(a.x, b.x, fma(a.y, b.y, a.z * b.z)));
  r.x = fma(a.w, b.x, fma(a.x, b.w, fma(a.y, b.z, -(a.z * b.y))));
  r.y = fma(a.w, b.y, fma(a.y, b.w, fma(a.z, b.x, -(a.x * b.z))));
  r.z = fma(a.w, b.z, fma(a.z, b.w, fma(a.x, b.y, -(a.y * b.x))));
  const double n2 = fma(r.w, r.w, fma(r.x, r.x, fma(r.y, r.y, r.z * r.z)));
  const double sc = fma(-0.5, n2, 1.5);
  C.q.w = r.w * sc; C.q.x = r.x * sc; C.q.y = r.y * sc; C.q.z = r.z * sc;
  // t = A.t + rotate(A.q, B.t)
  const double* v = B.t;
  double ux = fma(a.y, v[2], -(a.z * v[1])), uy = fma(a.z, v[0], -(a.x * v[2])), uz = fma(a.x, v[1], -(a.y * v[0]));
  ux += ux; uy += uy; uz += uz;
  C.t[0] = A.t[0] + fma(a.w, ux, v[0]) + fma(a.y, uz, -(a.z * uy));
  C.t[1] = A.t[1] + fma(a.w, uy, v[1]) + fma(a.z, ux, -(a.x * uz));
  C.t[2] = A.t[2] + fma(a.w, uz, v[2]) + fma(a.x, uy, -(a.y * ux));
  return C;
}

// 1/z, correctly rounded for finite normal z: f32 reciprocal seed, two f64 Newton steps, residual correction.
__device__ __forceinline__ double rcp_rn(double z) {
  double r = (double)__frcp_rn((float)z);
  double e = fma(-z, r, 1.0);
  r = fma(r, e, r);
  e = fma(-z, r, 1.0);
  r = fma(r, e, r);
  e = fma(-z, r, 1.0);
  return fma(r, e, r);
}

// Unpivoted LDL^T of a symmetric positive definite 6x6, fully unrolled in registers.
// L: strict lower triangle row-major (15), dinv: 1/d.  Returns false when a pivot is not safely
// positive (caller falls back to the pivoted Eigen-like routine above).
struct Fact6 {
  double L[15];
  double dinv[6];
};
__device__ __forceinline__ constexpr int tri(int i, int j) { return i * (i - 1) / 2 + j; }  // i > j
__device__ __forceinline__ bool fact6_compute(const double* H /*36 row-major*/, Fact6& F) {
  double d[6];
  double maxdiag = 0.0;
#pragma unroll
  for (int j = 0; j < 6; ++j) maxdiag = fmax(maxdiag, fabs(H[j * 6 + j]));
  bool ok = maxdiag > 0.0 && maxdiag < 1e300;
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    double dj = H[j * 6 + j];
#pragma unroll
    for (int k = 0; k < j; ++k) dj = fma(-(F.L[tri(j, k)] * F.L[tri(j, k)]), d[k], dj);
    d[j] = dj;
    ok = ok && (dj > 1e-13 * maxdiag);
    const double inv = 1.0 / dj;
    F.dinv[j] = inv;
#pragma unroll
    for (int i = j + 1; i < 6; ++i) {
      double v = H[i * 6 + j];
#pragma unroll
      for (int k = 0; k < j; ++k) v = fma(-(F.L[tri(i, k)] * F.L[tri(j, k)]), d[k], v);
      F.L[tri(i, j)] = v * inv;
    }
  }
  return ok;
}
// same, from the 21 upper-triangle entries (row-major: (0,0..5), (1,1..5), ...) held in registers
__device__ __forceinline__ constexpr int upper_idx(int r, int c) {  // r <= c
  return r * 6 - r * (r - 1) / 2 + (c - r);
}
__device__ __forceinline__ bool fact6_compute_upper(const double (&h)[21], Fact6& F) {
  double d[6];
  double maxdiag = 0.0;
#pragma unroll
  for (int j = 0; j < 6; ++j) maxdiag = fmax(maxdiag, fabs(h[upper_idx(j, j)]));
  bool ok = maxdiag > 0.0 && maxdiag < 1e300;
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    double dj = h[upper_idx(j, j)];
#pragma unroll
    for (int k = 0; k < j; ++k) dj = fma(-(F.L[tri(j, k)] * F.L[tri(j, k)]), d[k], dj);
    d[j] = dj;
    ok = ok && (dj > 1e-13 * maxdiag);
    const double adj = fabs(dj);
    const double inv = (adj > 1e-30 && adj < 1e30) ? rcp_rn(dj) : 1.0 / dj;  // correctly rounded reciprocal without the IEEE-division slow path
    F.dinv[j] = inv;
#pragma unroll
    for (int i = j + 1; i < 6; ++i) {
      double v = h[upper_idx(j, i)];
#pragma unroll
      for (int k = 0; k < j; ++k) v = fma(-(F.L[tri(i, k)] * F.L[tri(j, k)]), d[k], v);
      F.L[tri(i, j)] = v * inv;
    }
  }
  return ok;
}
__device__ __forceinline__ void fact6_solve(const Fact6& F, const double* b, double* x) {
  double y[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    double s = b[i];
#pragma unroll
    for (int j = 0; j < i; ++j) s = fma(-F.L[tri(i, j)], y[j], s);
    y[i] = s;
  }
#pragma unroll
  for (int i = 0; i < 6; ++i) y[i] *= F.dinv[i];
#pragma unroll
  for (int i = 5; i >= 0; --i) {
    double s = y[i];
#pragma unroll
    for (int j = i + 1; j < 6; ++j) s = fma(-F.L[tri(j, i)], x[j], s);
    x[i] = s;
  }
}

// LDL^T of the 6x6 normal matrix: the register-resident unpivoted factorisation when H is safely
// positive definite (always, outside degenerate inputs), else the pivoted Eigen-like routine.
struct Solver6 {
  Fact6 F;
  double ldl[36];
  int tr[8];
  int pivoted;
};
static __device__ __noinline__ void solver_factor(Solver6& S, const double* H) {
  Fact6 F;
  if (fact6_compute(H, F)) {
    S.F = F;
    S.pivoted = 0;
  } else {
    for (int k = 0; k < 36; ++k) S.ldl[k] = H[k];
    ldlt6_factor(S.ldl, S.tr);
    S.pivoted = 1;
  }
}


// ------------------------------------------------------------------------------------------
// Reductions: each warp folds its K doubles with __shfl_down, lane 0 parks them in shared memory.
// ------------------------------------------------------------------------------------------
template <int K>
__device__ __forceinline__ void warp_sum(double (&v)[K]) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
    for (int k = 0; k < K; ++k) v[k] += __shfl_down_sync(0xffffffffu, v[k], off);
  }
}


// Transposed ("reduce-scatter") warp sum of K = 8 or 16 doubles: each butterfly step halves the number
// of values a lane carries, so the whole reduction costs K-1+... f64 exchanges instead of 5*K.
// On return v[0] of lane (32/K)*e holds the warp total of element e.
template <int K>
__device__ __forceinline__ void warp_reduce_t(double (&v)[K]) {
  const unsigned lane = threadIdx.x & 31u;
  int bit = 16;
#pragma unroll
  for (int n = K; n > 1; n >>= 1, bit >>= 1) {
    const bool hi = (lane & (unsigned)bit) != 0;
#pragma unroll
    for (int i = 0; i < n / 2; ++i) {
      const double keep = hi ? v[n / 2 + i] : v[i];
      const double send = hi ? v[i] : v[n / 2 + i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
    }
  }
#pragma unroll
  for (; bit > 0; bit >>= 1) v[0] += __shfl_xor_sync(0xffffffffu, v[0], bit);
}

// 1/z to ~1 ulp without the IEEE-division slow path: f32 reciprocal seed + two f64 Newton steps.
__device__ __forceinline__ double fast_rcp(double z) {
  double r = (double)__frcp_rn((float)z);
  double e = fma(-z, r, 1.0);
  r = fma(r, e, r);
  e = fma(-z, r, 1.0);
  return fma(r, e, r);
}

// x / z given rz ~ 1/z (relative error << 2^-53, e.g. from fast_rcp): one residual correction turns the product
// x*rz into the correctly rounded quotient (Markstein's division step; the same sequence the compiler's own f64
// division ends with).  The reference divides (Eigen's project2d = head<2>() / z), so this keeps the f32-cast pixel
// coordinate on the reference's side of every rounding boundary.
__device__ __forceinline__ double div_rn(double x, double z, double rz) {
  const double q = x * rz;
  const double r = fma(-z, q, x);
  return fma(r, rz, q);
}

// ------------------------------------------------------------------------------------------
// [EXT] vk::AbstractCamera models, restated from the published rpg_vikit sources (pinhole_camera.cpp,
// atan_camera.cpp); `c` is the CamDev of ctx.h (kernel parameter, read through the constant bank).
//   world2cam(uv): unit-plane point -> pixel.      cam2world(px): pixel -> UNIT bearing vector.
// ------------------------------------------------------------------------------------------
template <class Cam>
__host__ __device__ __forceinline__ void cam_world2cam(const Cam& c, double x, double y, double& u, double& v) {
  if (!c.distorted) {  // px = fx*uv + cx  (both models without distortion)
    u = fma(c.fx, x, c.cx);
    v = fma(c.fy, y, c.cy);
  } else if (c.model == SVO_B200_CAM_PINHOLE) {  // vk::PinholeCamera::world2cam(const Vector2d&), radial-tangential
    const double r2 = fma(x, x, y * y), r4 = r2 * r2, r6 = r4 * r2;
    const double a1 = 2.0 * x * y, a2 = fma(2.0 * x, x, r2), a3 = fma(2.0 * y, y, r2);
    const double cdist = fma(c.d[4], r6, fma(c.d[1], r4, fma(c.d[0], r2, 1.0)));
    const double xd = fma(c.d[3], a2, fma(c.d[2], a1, x * cdist));
    const double yd = fma(c.d[3], a1, fma(c.d[2], a3, y * cdist));
    u = fma(xd, c.fx, c.cx);
    v = fma(yd, c.fy, c.cy);
  } else {  // vk::ATANCamera::world2cam: factor = rtrans_factor(|uv|) = atan(r * tans) / (s * r)
    const double r = sqrt(fma(x, x, y * y));
    const double factor = r < 0.001 ? 1.0 : c.s_inv * atan(r * c.tans) / r;
    u = fma(c.fx * factor, x, c.cx);
    v = fma(c.fy * factor, y, c.cy);
  }
}
template <class Cam>
__host__ __device__ __forceinline__ void cam_cam2world(const Cam& c, double u, double v, double (&f)[3]) {
  double x, y;
  if (c.model == SVO_B200_CAM_PINHOLE) {
    if (!c.distorted) {
      x = (u - c.cx) / c.fx;
      y = (v - c.cy) / c.fy;
    } else {
      // cv::undistortPoints on ONE CV_32FC2 point [EXT OpenCV]: float in, 5 fixed-point iterations in double, float out
      const double uf = (double)(float)u, vf = (double)(float)v;
      const double x0 = (uf - c.cx) * c.fx_inv, y0 = (vf - c.cy) * c.fy_inv;
      x = x0; y = y0;
      for (int j = 0; j < 5; ++j) {
        const double r2 = x * x + y * y;
        const double icdist = 1.0 / (1.0 + ((c.d[4] * r2 + c.d[1]) * r2 + c.d[0]) * r2);
        const double dX = 2.0 * c.d[2] * x * y + c.d[3] * (r2 + 2.0 * x * x);
        const double dY = c.d[2] * (r2 + 2.0 * y * y) + 2.0 * c.d[3] * x * y;
        x = (x0 - dX) * icdist;
        y = (y0 - dY) * icdist;
      }
      x = (double)(float)x;
      y = (double)(float)y;
    }
  } else {  // vk::ATANCamera::cam2world
    const double dx = (u - c.cx) * c.fx_inv, dy = (v - c.cy) * c.fy_inv;
    const double dist_r = sqrt(dx * dx + dy * dy);
    const double r = c.distorted ? tan(dist_r * c.d[0]) * c.tans_inv : dist_r;  // invrtrans
    const double d_factor = dist_r > 0.01 ? r / dist_r : 1.0;
    x = d_factor * dx;
    y = d_factor * dy;
  }
  const double n = sqrt(x * x + y * y + 1.0);
  f[0] = x / n; f[1] = y / n; f[2] = 1.0 / n;
}

// byte k of w -> float, exactly: PRMT builds the float 2^23 + byte, one FADD removes the 2^23.
template <int KB>
__device__ __forceinline__ float byte_to_float(uint32_t w) {
  return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7540u | (unsigned)KB)) - 8388608.0f;
}

// Canonical bilinear blend shared with the oracle (oracle/svo_oracle.cpp `bilin`):
// fma(wbr,d, fma(wbl,c, fma(wtl,a, wtr*b))).
__device__ __forceinline__ float bilin(float wtl, float wtr, float wbl, float wbr, float a, float b,
                                       float c, float d) {
  return fmaf(wbr, d, fmaf(wbl, c, fmaf(wtl, a, __fmul_rn(wtr, b))));
}

// Bilinear weights exactly as the reference forms them: (1.0 - su) promotes to double, the
// product is rounded once to float (svo/src/sparse_img_align.cpp:115-120,194-199).  Pure-f32
// evaluation is bit-identical whenever the coordinate is >= 2 (always true inside the +-3 / +-4 pixel
// borders every caller enforces): su = u - floor(u) is then a multiple of 2^-22, so 1 - su is exact
// in f32, and the product of two exact f32 operands rounded once to f32 equals the double product
// rounded to f32.  This keeps the conversion (XU) pipe out of the inner loops.
__device__ __forceinline__ void bilin_weights(float su, float sv, float& wtl, float& wtr, float& wbl,
                                              float& wbr) {
  const float omu = __fsub_rn(1.0f, su), omv = __fsub_rn(1.0f, sv);
  wtl = __fmul_rn(omu, omv);
  wtr = __fmul_rn(su, omv);
  wbl = __fmul_rn(omu, sv);
  wbr = __fmul_rn(su, sv);
}

// floor() of a float known to lie in [0, 2^22) without the conversion pipe: adding 2^23 rounds to the
// nearest integer (result integer in the low mantissa bits), one compare fixes round-up cases.
// Returns the integer; `fl` receives floor(u) as a float.
__device__ __forceinline__ int floor_pos(float u, float& fl) {
  const float t = __fadd_rn(u, 8388608.0f);
  float r = __fsub_rn(t, 8388608.0f);
  int i = __float_as_int(t) - 0x4B000000;
  if (r > u) { r = __fsub_rn(r, 1.0f); i -= 1; }
  fl = r;
  return i;
}

// ------------------------------------------------------------------------------------------
// TMA (bulk async copy) + mbarrier primitives, sm_90+/sm_100a PTX.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes,
                                             uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

}  // namespace svo
