// rpg_svo_b200/csrc/pose_opt.cu -- pose_optimizer::optimizeGaussNewton on sm_100a.
//
// Replaces svo/src/pose_optimizer.cpp:28-161: robust (Tukey weights, MAD scale) Gauss-Newton on the
// unit-plane reprojection error of one frame, outlier culling, covariance and the median error
// statistics.  One CTA runs the whole optimisation on the device (<= n_iter iterations without a host
// round trip): threads stride over the features, every iteration ends in a 28-value block reduction
// (21 unique A entries, 6 b entries, chi2), thread 0 factorises the 6x6 system, applies
// T <- exp(dT) * T and the accept / rollback rule, and broadcasts the pose.  The three medians
// ([EXT] vk::getMedian = nth_element at floor(n/2)) are exact order statistics obtained by an
// 8-pass radix select on order-preserving keys in shared memory.  All arithmetic is f64 except the f32 error vector / Tukey weight, as
// in the reference.
#include <cstring>

#include "ctx.h"
#include "svo_math.cuh"

namespace svo {

constexpr int kPoThreads = 512;
constexpr int kPoWarps = kPoThreads / 32;
constexpr int kPoK = 28;

struct PoseOptParams {
  const double* f;
  const double* pos;
  const int* level;
  uint8_t* has_point;  // in/out
  int N;
  int n_iter;
  double fx, reproj_thresh;
  double* T_io;  // 12
  svo_b200_pose_opt_result* out;
};

struct PoseOptShared {
  double part[kPoWarps * kPoK];
  double sums[kPoK];
  double R[9], t[3];
  double A[36];
  double x[8];
  double med;
  Pose T, T_old;
  Solver6 sol;
  double chi2, scale;
  int done, num_obs, iters, n_deleted;
  unsigned hist[256];
  int sel_bin, sel_k;
};

// [EXT] vk::robust_cost::TukeyWeightFunction::value, b = 4.6851f
__device__ __forceinline__ float tukey_weight(float x) {
  const float b_square = __fmul_rn(4.6851f, 4.6851f);
  const float x_square = __fmul_rn(x, x);
  if (x_square <= b_square) {
    const float tmp = __fsub_rn(1.0f, __fdiv_rn(x_square, b_square));
    return __fmul_rn(tmp, tmp);
  }
  return 0.0f;
}

template <int K>
__device__ __forceinline__ void po_block_sum(double (&v)[K], PoseOptShared& s) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (K == kPoK) {  // 28 = 16 + 8 (transposed reductions) + 4 (tree)
    double a16[16], a8[8], a4[4];
#pragma unroll
    for (int k = 0; k < 16; ++k) a16[k] = v[k];
#pragma unroll
    for (int k = 0; k < 8; ++k) a8[k] = v[16 + k];
#pragma unroll
    for (int k = 0; k < 4; ++k) a4[k] = v[24 + k];
    warp_reduce_t<16>(a16);
    warp_reduce_t<8>(a8);
    warp_sum<4>(a4);
    if ((lane & 1) == 0) s.part[warp * kPoK + (lane >> 1)] = a16[0];
    if ((lane & 3) == 0) s.part[warp * kPoK + 16 + (lane >> 2)] = a8[0];
    if (lane == 0) {
#pragma unroll
      for (int k = 0; k < 4; ++k) s.part[warp * kPoK + 24 + k] = a4[k];
    }
  } else {
    warp_sum<K>(v);
    if (lane == 0) {
#pragma unroll
      for (int k = 0; k < K; ++k) s.part[warp * kPoK + k] = v[k];
    }
  }
  __syncthreads();
  if (warp == 0) {
    if (lane < K) {
      double acc = 0.0;
      for (int w = 0; w < kPoWarps; ++w) acc += s.part[w * kPoK + lane];
      s.sums[lane] = acc;
    }
    __syncwarp();
  }
  __syncthreads();
}

// k-th smallest (0-based) of the entries of v[0..N) whose valid flag is set: exact order statistic by
// MSB-first radix select on order-preserving 64-bit keys (8 passes of 8 bits; shared-memory histogram,
// warp-parallel bin scan).  Result in s.med (NaN if there are no more than k valid entries).
__device__ __forceinline__ unsigned long long order_key(double d) {
  const unsigned long long u = (unsigned long long)__double_as_longlong(d);
  return (u >> 63) ? ~u : (u | 0x8000000000000000ULL);
}
__device__ void block_kth(const double* v, const uint8_t* valid, int N, int k, PoseOptShared& s) {
  unsigned long long prefix = 0, mask = 0;
  int kk = k;
  for (int pass = 7; pass >= 0; --pass) {
    const int shift = pass * 8;
    for (int b = threadIdx.x; b < 256; b += blockDim.x) s.hist[b] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
      if (!valid[i]) continue;
      const unsigned long long key = order_key(v[i]);
      if ((key & mask) == prefix) atomicAdd(&s.hist[(unsigned)(key >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (threadIdx.x < 32) {  // warp 0: lane l owns bins [8l, 8l+8)
      const int lane = threadIdx.x;
      unsigned loc[8], sum = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) { loc[j] = s.hist[8 * lane + j]; sum += loc[j]; }
      unsigned inc = sum;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
      }
      const unsigned exc = inc - sum;
      const unsigned hit = __ballot_sync(0xffffffffu, inc > (unsigned)kk);
      if (hit == 0) {
        if (lane == 0) { s.sel_bin = -1; }
      } else if (lane == __ffs(hit) - 1) {
        unsigned cum = exc;
        int bin = 8 * lane;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (cum + loc[j] > (unsigned)kk) { bin = 8 * lane + j; break; }
          cum += loc[j];
        }
        s.sel_bin = bin;
        s.sel_k = kk - (int)cum;
      }
    }
    __syncthreads();
    if (s.sel_bin < 0) {  // fewer than k+1 valid entries
      if (threadIdx.x == 0) s.med = __longlong_as_double(0x7ff8000000000000LL);
      __syncthreads();
      return;
    }
    prefix |= (unsigned long long)s.sel_bin << shift;
    mask |= 0xffULL << shift;
    kk = s.sel_k;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const unsigned long long u = (prefix >> 63) ? (prefix & 0x7fffffffffffffffULL) : ~prefix;
    s.med = __longlong_as_double((long long)u);
  }
  __syncthreads();
}

__device__ __forceinline__ void reproj_error(const PoseOptParams& P, int i, const double* R, const double* t,
                                             double& ex, double& ey, double* xyz_f) {
  const double px = P.pos[3 * i], py = P.pos[3 * i + 1], pz = P.pos[3 * i + 2];
  // xyz_f = T_f_w * pos
  xyz_f[0] = R[0] * px + R[1] * py + R[2] * pz + t[0];
  xyz_f[1] = R[3] * px + R[4] * py + R[5] * pz + t[1];
  xyz_f[2] = R[6] * px + R[7] * py + R[8] * pz + t[2];
  const double fxn = P.f[3 * i] / P.f[3 * i + 2], fyn = P.f[3 * i + 1] / P.f[3 * i + 2];  // project2d(f)
  const double sic = 1.0 / (double)(1 << P.level[i]);
  ex = (fxn - xyz_f[0] / xyz_f[2]) * sic;
  ey = (fyn - xyz_f[1] / xyz_f[2]) * sic;
}

__global__ void __launch_bounds__(kPoThreads) pose_opt_kernel(PoseOptParams P) {
  extern __shared__ __align__(16) unsigned char po_smem[];
  PoseOptShared& s = *reinterpret_cast<PoseOptShared*>(po_smem);
  double* work = reinterpret_cast<double*>(po_smem + ((sizeof(PoseOptShared) + 15) & ~size_t(15)));
  double* init = work + P.N;
  const int tid = threadIdx.x, N = P.N;

  if (tid == 0) {
    s.T = pose_from_rt12(P.T_io);
    s.T_old = s.T;
    qmatrix(s.T.q, s.R);
    s.t[0] = s.T.t[0]; s.t[1] = s.T.t[1]; s.t[2] = s.T.t[2];
    s.chi2 = 0.0; s.done = 0; s.iters = 0; s.n_deleted = 0;
    for (int k = 0; k < 36; ++k) s.A[k] = 0.0;
  }
  __syncthreads();

  // ---- scale of the error for robust estimation (:47-60) ------------------------------------
  double cnt[1] = {0.0};
  {
    double R[9], t[3];
    for (int k = 0; k < 9; ++k) R[k] = s.R[k];
    for (int k = 0; k < 3; ++k) t[k] = s.t[k];
    for (int i = tid; i < N; i += kPoThreads) {
      if (!P.has_point[i]) continue;
      double ex, ey, xyz[3];
      reproj_error(P, i, R, t, ex, ey, xyz);
      work[i] = (double)(float)sqrt(ex * ex + ey * ey);  // errors.push_back(e.norm()) -> float
      cnt[0] += 1.0;
    }
  }
  po_block_sum<1>(cnt, s);
  const int num_obs = (int)s.sums[0];
  if (num_obs == 0) {  // errors.empty() -> return (:57-58)
    if (tid == 0) {
      svo_b200_pose_opt_result r;
      memset(&r, 0, sizeof(r));
      *P.out = r;
    }
    return;
  }
  block_kth(work, P.has_point, N, num_obs / 2, s);
  // [EXT] MADScaleEstimator: 1.48f * median (float arithmetic)
  const double estimated_scale = (double)__fmul_rn(1.48f, (float)s.med);
  double scale = estimated_scale;

  // ---- Gauss-Newton (:63-121) -------------------------------------------------------------------
  for (int iter = 0; iter < P.n_iter; ++iter) {
    if (iter == 5) scale = 0.85 / P.fx;  // (:69-70)
    double R[9], t[3];
    for (int k = 0; k < 9; ++k) R[k] = s.R[k];
    for (int k = 0; k < 3; ++k) t[k] = s.t[k];
    double acc[kPoK];
#pragma unroll
    for (int k = 0; k < kPoK; ++k) acc[k] = 0.0;
    for (int i = tid; i < N; i += kPoThreads) {
      if (!P.has_point[i]) continue;
      double ex, ey, p[3];
      reproj_error(P, i, R, t, ex, ey, p);
      const double sic = 1.0 / (double)(1 << P.level[i]);
      // Frame::jacobian_xyz2uv (frame.h:116-138), then J *= sqrt_inv_cov
      const double x = p[0], y = p[1], z_inv = 1. / p[2], z_inv_2 = z_inv * z_inv;
      double J0[6], J1[6];
      J0[0] = -z_inv; J0[1] = 0.0; J0[2] = x * z_inv_2; J0[3] = y * J0[2]; J0[4] = -(1.0 + x * J0[2]); J0[5] = y * z_inv;
      J1[0] = 0.0; J1[1] = -z_inv; J1[2] = y * z_inv_2; J1[3] = 1.0 + y * J1[2]; J1[4] = -J0[3]; J1[5] = -x * z_inv;
#pragma unroll
      for (int k = 0; k < 6; ++k) { J0[k] *= sic; J1[k] *= sic; }
      const double e_sq = ex * ex + ey * ey;
      if (iter == 0) init[i] = e_sq;  // chi2_vec_init (:87-88)
      const double w = (double)tukey_weight((float)(sqrt(e_sq) / scale));
      int idx = 0;
#pragma unroll
      for (int r = 0; r < 6; ++r)
#pragma unroll
        for (int c = r; c < 6; ++c, ++idx) acc[idx] += (J0[r] * J0[c] + J1[r] * J1[c]) * w;
#pragma unroll
      for (int r = 0; r < 6; ++r) acc[21 + r] -= (J0[r] * ex + J1[r] * ey) * w;
      acc[27] += e_sq * w;
    }
    po_block_sum<kPoK>(acc, s);
    if (tid == 0) {
      int idx = 0;
      for (int r = 0; r < 6; ++r)
        for (int c = r; c < 6; ++c, ++idx) { s.A[r * 6 + c] = s.sums[idx]; s.A[c * 6 + r] = s.sums[idx]; }
      const double new_chi2 = s.sums[27];
      double dT[6];
      solver_factor(s.sol, s.A);
      if (!s.sol.pivoted) {
        double b[6];
        for (int k = 0; k < 6; ++k) b[k] = s.sums[21 + k];
        fact6_solve(s.sol.F, b, dT);
      } else {
        for (int k = 0; k < 6; ++k) s.x[k] = s.sums[21 + k];
        ldlt6_solve(s.sol.ldl, s.sol.tr, s.x);
        for (int k = 0; k < 6; ++k) dT[k] = s.x[k];
      }
      s.iters++;
      if ((iter > 0 && new_chi2 > s.chi2) || isnan(dT[0])) {
        s.T = s.T_old;  // roll-back (:100-107)
        s.done = 1;
      } else {
        const Pose Tn = pose_mul_fast(se3_exp_fast(dT), s.T);  // exp(dT) * T  (:110)
        s.T_old = s.T;
        s.T = Tn;
        s.chi2 = new_chi2;
        double m = 0;
        for (int k = 0; k < 6; ++k) m = fmax(m, fabs(dT[k]));
        if (m <= 0.0000000001) s.done = 1;  // EPS (global.h:77)
      }
      qmatrix(s.T.q, s.R);
      s.t[0] = s.T.t[0]; s.t[1] = s.T.t[1]; s.t[2] = s.T.t[2];
    }
    __syncthreads();
    if (s.done) break;
  }

  // ---- remove measurements with too large reprojection error (:129-145) ------------------------
  const double thresh = P.reproj_thresh / P.fx;
  double del[1] = {0.0};
  {
    double R[9], t[3];
    for (int k = 0; k < 9; ++k) R[k] = s.R[k];
    for (int k = 0; k < 3; ++k) t[k] = s.t[k];
    for (int i = tid; i < N; i += kPoThreads) {
      if (!P.has_point[i]) continue;
      double ex, ey, xyz[3];
      reproj_error(P, i, R, t, ex, ey, xyz);
      const double e_sq = ex * ex + ey * ey;
      work[i] = e_sq;  // chi2_vec_final
      if (sqrt(e_sq) > thresh) del[0] += 1.0;
    }
  }
  // medians use the pre-culling validity flags: both vectors hold one entry per original observation
  block_kth(init, P.has_point, N, num_obs / 2, s);
  const double med_init = (P.n_iter > 0) ? s.med : 0.0;
  __syncthreads();
  block_kth(work, P.has_point, N, num_obs / 2, s);
  const double med_final = s.med;
  __syncthreads();
  for (int i = tid; i < N; i += kPoThreads)
    if (P.has_point[i] && sqrt(work[i]) > thresh) P.has_point[i] = 0;  // point = NULL
  po_block_sum<1>(del, s);

  if (tid == 0) {
    svo_b200_pose_opt_result r;
    memset(&r, 0, sizeof(r));
    // Cov_ = (A * fx^2)^-1  (:125-126), Gauss-Jordan with partial pivoting on [A|I]
    double M[6][12];
    const double f2 = P.fx * P.fx;
    for (int a = 0; a < 6; ++a)
      for (int b = 0; b < 6; ++b) { M[a][b] = s.A[a * 6 + b] * f2; M[a][6 + b] = (a == b) ? 1.0 : 0.0; }
    for (int c = 0; c < 6; ++c) {
      int p = c;
      for (int rr = c + 1; rr < 6; ++rr)
        if (fabs(M[rr][c]) > fabs(M[p][c])) p = rr;
      if (p != c)
        for (int j = 0; j < 12; ++j) { const double tmp = M[c][j]; M[c][j] = M[p][j]; M[p][j] = tmp; }
      const double d = 1.0 / M[c][c];
      for (int j = 0; j < 12; ++j) M[c][j] *= d;
      for (int rr = 0; rr < 6; ++rr)
        if (rr != c) {
          const double fct = M[rr][c];
          for (int j = 0; j < 12; ++j) M[rr][j] -= fct * M[c][j];
        }
    }
    for (int a = 0; a < 6; ++a)
      for (int b = 0; b < 6; ++b) r.cov[a * 6 + b] = M[a][6 + b];
    r.estimated_scale = estimated_scale * P.fx;
    r.error_init = sqrt(med_init) * P.fx;
    r.error_final = sqrt(med_final) * P.fx;
    r.num_obs = (long long)num_obs - (long long)s.sums[0];
    r.n_iter_done = s.iters;
    *P.out = r;
    pose_to_rt12(s.T, P.T_io);
  }
}

}  // namespace svo

using namespace svo;

extern "C" int svo_b200_pose_optimize(svo_b200_ctx* ctx, double reproj_thresh, int n_iter, double fx,
                                      double* T_f_w_io, const double* f, const double* point_pos, const int* level,
                                      uint8_t* has_point_io, int N, svo_b200_pose_opt_result* out) {
  if (!ctx || !T_f_w_io || !out || N < 0 || n_iter < 0 || (N > 0 && (!f || !point_pos || !level || !has_point_io)))
    return set_err(ctx, SVO_B200_EINVAL, "pose_optimize: bad arguments");
  memset(out, 0, sizeof(*out));
  if (N == 0) return 0;  // errors.empty() -> return
  cudaSetDevice(ctx->device);
  const size_t smem = ((sizeof(PoseOptShared) + 15) & ~size_t(15)) + sizeof(double) * 2 * (size_t)N;
  if (smem > (size_t)ctx->max_smem_optin)
    return set_err(ctx, SVO_B200_ELIMIT, "pose_optimize: %d features need %zu B of shared memory", N, smem);
  Carver c;
  const size_t o_T = c.take(sizeof(double) * 12), o_hp = c.take(N), o_out = c.take(sizeof(svo_b200_pose_opt_result));
  const size_t io_end = c.off;
  const size_t o_f = c.take(sizeof(double) * 3 * N), o_pos = c.take(sizeof(double) * 3 * N), o_lv = c.take(sizeof(int) * N);
  int rc;
  if ((rc = ensure_host(ctx, ctx->h_in, c.off))) return rc;
  if ((rc = ensure_dev(ctx, ctx->d_in, c.off))) return rc;
  SVO_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  uint8_t* h = static_cast<uint8_t*>(ctx->h_in.p);
  uint8_t* d = static_cast<uint8_t*>(ctx->d_in.p);
  memcpy(h + o_T, T_f_w_io, sizeof(double) * 12);
  memcpy(h + o_hp, has_point_io, N);
  memset(h + o_out, 0, sizeof(svo_b200_pose_opt_result));
  memcpy(h + o_f, f, sizeof(double) * 3 * N);
  memcpy(h + o_pos, point_pos, sizeof(double) * 3 * N);
  memcpy(h + o_lv, level, sizeof(int) * N);
  SVO_CUDA_CHECK(ctx, cudaMemcpyAsync(d, h, c.off, cudaMemcpyHostToDevice, ctx->stream));
  PoseOptParams P;
  P.f = reinterpret_cast<const double*>(d + o_f);
  P.pos = reinterpret_cast<const double*>(d + o_pos);
  P.level = reinterpret_cast<const int*>(d + o_lv);
  P.has_point = d + o_hp;
  P.N = N;
  P.n_iter = n_iter;
  P.fx = fx;
  P.reproj_thresh = reproj_thresh;
  P.T_io = reinterpret_cast<double*>(d + o_T);
  P.out = reinterpret_cast<svo_b200_pose_opt_result*>(d + o_out);
  SVO_CUDA_CHECK(ctx, cudaFuncSetAttribute(pose_opt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  pose_opt_kernel<<<1, kPoThreads, smem, ctx->stream>>>(P);
  ctx->launches++;
  SVO_CUDA_CHECK(ctx, cudaGetLastError());
  SVO_CUDA_CHECK(ctx, cudaMemcpyAsync(h, d, io_end, cudaMemcpyDeviceToHost, ctx->stream));
  SVO_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  memcpy(T_f_w_io, h + o_T, sizeof(double) * 12);
  memcpy(has_point_io, h + o_hp, N);
  memcpy(out, h + o_out, sizeof(*out));
  return 0;
}

// ================================================================================================
// Point::optimize (svo/src/point.cpp:119-177): 3-DoF Gauss-Newton on a point's world position over
// the frames observing it.  Points are independent: one thread per point, observations streamed from
// global memory; the 3x3 system is solved with the same pivoted LDL^T as Eigen's (tiny, in registers).
// ================================================================================================
namespace svo {

__device__ inline void ldlt3_solve(double (&A)[3][3], const double (&b)[3], double (&x)[3]) {
  int tr[3];
  for (int k = 0; k < 3; ++k) {
    int big = k;
    double bigv = fabs(A[k][k]);
    for (int i = k + 1; i < 3; ++i)
      if (fabs(A[i][i]) > bigv) { bigv = fabs(A[i][i]); big = i; }
    tr[k] = big;
    if (big != k) {
      for (int j = 0; j < k; ++j) { const double s = A[k][j]; A[k][j] = A[big][j]; A[big][j] = s; }
      for (int i = big + 1; i < 3; ++i) { const double s = A[i][k]; A[i][k] = A[i][big]; A[i][big] = s; }
      { const double s = A[k][k]; A[k][k] = A[big][big]; A[big][big] = s; }
      for (int i = k + 1; i < big; ++i) { const double s = A[i][k]; A[i][k] = A[big][i]; A[big][i] = s; }
    }
    if (k > 0) {
      double temp[3], acc = 0;
      for (int j = 0; j < k; ++j) { temp[j] = A[j][j] * A[k][j]; acc += A[k][j] * temp[j]; }
      A[k][k] -= acc;
      for (int i = k + 1; i < 3; ++i) {
        double a2 = 0;
        for (int j = 0; j < k; ++j) a2 += A[i][j] * temp[j];
        A[i][k] -= a2;
      }
    }
    const double akk = A[k][k];
    const bool ok = fabs(akk) > 0.0;
    if (k == 0 && !ok) { tr[0] = 0; tr[1] = 1; tr[2] = 2; break; }
    if (ok)
      for (int i = k + 1; i < 3; ++i) A[i][k] /= akk;
  }
  for (int i = 0; i < 3; ++i) x[i] = b[i];
  for (int i = 0; i < 3; ++i) { const int j = tr[i]; const double s = x[i]; x[i] = x[j]; x[j] = s; }
  for (int i = 1; i < 3; ++i)
    for (int j = 0; j < i; ++j) x[i] -= A[i][j] * x[j];
  for (int i = 0; i < 3; ++i) x[i] = (fabs(A[i][i]) > 5.562684646268003e-309) ? x[i] / A[i][i] : 0.0;
  for (int i = 1; i >= 0; --i)
    for (int j = i + 1; j < 3; ++j) x[i] -= A[j][i] * x[j];
  for (int i = 2; i >= 0; --i) { const int j = tr[i]; const double s = x[i]; x[i] = x[j]; x[j] = s; }
}

__global__ void point_optimize_kernel(int P, int n_iter, const int* __restrict__ obs_offset,
                                      const int* __restrict__ obs_frame, const double* __restrict__ obs_f,
                                      const double* __restrict__ frame_T, double* __restrict__ pos_io) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  double pos[3] = {pos_io[3 * p], pos_io[3 * p + 1], pos_io[3 * p + 2]};
  double old_point[3] = {pos[0], pos[1], pos[2]};
  double chi2 = 0.0;
  const int o0 = obs_offset[p], o1 = obs_offset[p + 1];
  for (int it = 0; it < n_iter; ++it) {
    double A[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, b[3] = {0, 0, 0}, new_chi2 = 0.0;
    for (int o = o0; o < o1; ++o) {
      const double* T = frame_T + 12 * (size_t)obs_frame[o];
      const double px = T[0] * pos[0] + T[1] * pos[1] + T[2] * pos[2] + T[3];
      const double py = T[4] * pos[0] + T[5] * pos[1] + T[6] * pos[2] + T[7];
      const double pz = T[8] * pos[0] + T[9] * pos[1] + T[10] * pos[2] + T[11];
      const double z_inv = 1.0 / pz, z_inv_sq = z_inv * z_inv;
      const double pj[2][3] = {{z_inv, 0.0, -px * z_inv_sq}, {0.0, z_inv, -py * z_inv_sq}};
      double J[2][3];
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) J[r][c] = -(pj[r][0] * T[c] + pj[r][1] * T[4 + c] + pj[r][2] * T[8 + c]);
      const double ex = obs_f[3 * o] / obs_f[3 * o + 2] - px / pz, ey = obs_f[3 * o + 1] / obs_f[3 * o + 2] - py / pz;
      new_chi2 += ex * ex + ey * ey;
#pragma unroll
      for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int c = 0; c < 3; ++c) A[r][c] += J[0][r] * J[0][c] + J[1][r] * J[1][c];
        b[r] -= J[0][r] * ex + J[1][r] * ey;
      }
    }
    double dp[3];
    ldlt3_solve(A, b, dp);
    if ((it > 0 && new_chi2 > chi2) || isnan(dp[0])) {
      pos[0] = old_point[0]; pos[1] = old_point[1]; pos[2] = old_point[2];  // roll-back
      break;
    }
    for (int k = 0; k < 3; ++k) { old_point[k] = pos[k]; pos[k] += dp[k]; }
    chi2 = new_chi2;
    if (fmax(fabs(dp[0]), fmax(fabs(dp[1]), fabs(dp[2]))) <= 0.0000000001) break;
  }
  pos_io[3 * p] = pos[0]; pos_io[3 * p + 1] = pos[1]; pos_io[3 * p + 2] = pos[2];
}

}  // namespace svo

extern "C" int svo_b200_point_optimize_batch(svo_b200_ctx* ctx, int P, int n_iter, const int* obs_offset,
                                             const int* obs_frame, const double* obs_f, const double* frame_T_f_w,
                                             int n_frames, double* pos_io) {
  if (!ctx || P < 0 || n_iter < 0 || n_frames <= 0 || (P > 0 && (!obs_offset || !obs_frame || !obs_f || !frame_T_f_w || !pos_io)))
    return set_err(ctx, SVO_B200_EINVAL, "point_optimize_batch: bad arguments");
  if (P == 0) return 0;
  const int n_obs = obs_offset[P] - obs_offset[0];
  for (int o = 0; o < n_obs; ++o)
    if (obs_frame[obs_offset[0] + o] < 0 || obs_frame[obs_offset[0] + o] >= n_frames)
      return set_err(ctx, SVO_B200_EINVAL, "point_optimize_batch: obs_frame[%d] out of range", o);
  cudaSetDevice(ctx->device);
  Carver c;
  const size_t o_pos = c.take(sizeof(double) * 3 * P);
  const size_t io_end = c.off;
  const size_t o_off = c.take(sizeof(int) * (P + 1)), o_fr = c.take(sizeof(int) * (n_obs + 1)),
               o_f = c.take(sizeof(double) * 3 * (n_obs + 1)), o_T = c.take(sizeof(double) * 12 * n_frames);
  int rc;
  if ((rc = ensure_host(ctx, ctx->h_in, c.off))) return rc;
  if ((rc = ensure_dev(ctx, ctx->d_in, c.off))) return rc;
  SVO_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  uint8_t* h = static_cast<uint8_t*>(ctx->h_in.p);
  uint8_t* d = static_cast<uint8_t*>(ctx->d_in.p);
  memcpy(h + o_pos, pos_io, sizeof(double) * 3 * P);
  int* off = reinterpret_cast<int*>(h + o_off);
  for (int p = 0; p <= P; ++p) off[p] = obs_offset[p] - obs_offset[0];
  memcpy(h + o_fr, obs_frame + obs_offset[0], sizeof(int) * n_obs);
  memcpy(h + o_f, obs_f + 3 * (size_t)obs_offset[0], sizeof(double) * 3 * n_obs);
  memcpy(h + o_T, frame_T_f_w, sizeof(double) * 12 * n_frames);
  SVO_CUDA_CHECK(ctx, cudaMemcpyAsync(d, h, c.off, cudaMemcpyHostToDevice, ctx->stream));
  const int threads = 128, blocks = (P + threads - 1) / threads;
  point_optimize_kernel<<<blocks, threads, 0, ctx->stream>>>(P, n_iter, reinterpret_cast<const int*>(d + o_off),
                                                             reinterpret_cast<const int*>(d + o_fr),
                                                             reinterpret_cast<const double*>(d + o_f),
                                                             reinterpret_cast<const double*>(d + o_T),
                                                             reinterpret_cast<double*>(d + o_pos));
  ctx->launches++;
  SVO_CUDA_CHECK(ctx, cudaGetLastError());
  SVO_CUDA_CHECK(ctx, cudaMemcpyAsync(h + o_pos, d + o_pos, io_end - o_pos, cudaMemcpyDeviceToHost, ctx->stream));
  SVO_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  memcpy(pos_io, h + o_pos, sizeof(double) * 3 * P);
  return 0;
}
