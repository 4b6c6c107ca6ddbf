// rpg_svo_b200/csrc/sparse_align.cu -- svo::SparseImgAlign on sm_100a.
//
// Replaces svo/src/sparse_img_align.cpp:43-258 (run / precomputeReferencePatches /
// computeResiduals / solve / update) together with the Gauss-Newton driver of
// vk::NLLSSolver<6,SE3>::optimizeGaussNewton [EXT] that the class derives from.
//
// Design (B200-first, not a translation):
//   * one CTA per frame pair runs the WHOLE coarse-to-fine loop on the device: no host round trip
//     per Gauss-Newton iteration, batches of pairs fill the 148 SMs (grid = #pairs).
//   * one thread owns one feature (FPT features when N > blockDim): its bearing/depth state lives in
//     registers for the whole run; the 4x4 reference patch, and its two gradient images, live in
//     shared memory in pixel-major (SoA) order so a warp's accesses are conflict free.
//   * inverse-compositional structure is exploited: the per-pixel Jacobian is
//     J_p = dx_p * a + dy_p * b with a, b per-FEATURE 6-vectors, hence
//        sum_p J_p J_p^T = Sxx aa^T + Sxy (ab^T + ba^T) + Syy bb^T     (pose independent)
//        sum_p J_p r_p   = (sum dx_p r_p) a + (sum dy_p r_p) b
//     so the 6x6 normal matrix is reduced and factorised ONCE per level (and re-formed only in the
//     iterations where some patch leaves the current image), and an iteration costs 3 f32 FMAs per
//     pixel plus ~15 f64 FMAs per feature instead of the reference's 27 f64 MACs per pixel.
//   * the packed feature records of the pair and the coarse current-level images are staged into
//     shared memory with TMA bulk copies (cp.async.bulk + mbarrier); fine levels are gathered with
//     two aligned 32-bit read-only loads per 5-byte footprint row.
//   * each warp folds its partial sums (6 Jres + chi2 + counts) with __shfl_down; one thread does
//     the 6x6 substitution, SE3 exp and the accept / rollback decision and broadcasts the new pose.
// Precision follows the reference per quantity: f32 interpolation/residual/chi2, f64 geometry and
// normal equations.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <type_traits>
#include <vector>

#include "ctx.h"
#include "svo_math.cuh"

namespace svo {

#ifndef SVO_SIA_DEBUG
#define SVO_SIA_DEBUG 0  // 1: thread 0 accumulates clock64 section timings and a few CTAs print them
#endif
constexpr int kPatchArea = 16;
constexpr int kPartK = 24;  // widest block reduction: 21 unique H entries + 1 count, padded to 16 + 8
constexpr int kMaxWarps = 16;  // blockDim <= 512

struct SiaJob {  // one frame pair; array lives in device memory
  const uint8_t* ref_lvl[SVO_B200_MAX_LEVELS];
  const uint8_t* cur_lvl[SVO_B200_MAX_LEVELS];
  const uint8_t* blob;  // packed features: px[np*2] f[np*3] pos[np*3] (f64) then has_point[np] (u8)
  int n_feat, n_pad;
  int feat_off;  // offset of this pair in visible_out
  int pad_;
  double T[12];
  double ref_pos[3];
};

struct SiaParams {
  const SiaJob* jobs;
  int w[SVO_B200_MAX_LEVELS], h[SVO_B200_MAX_LEVELS];
  double fx, fy, cx, cy;
  int max_level, min_level, n_iter;
  double eps;
  int stage_cap;  // bytes of the TMA staging region in shared memory
  int slots;      // blockDim * FPT feature slots (patch arrays are [3][16][slots])
  double* T_out;
  double* H_out;
  uint8_t* visible_out;
  svo_b200_sia_stats* stats;
  svo_b200_sia_iter* trace;
  int trace_cap;
  int* n_trace;
  int debug;  // env SVO_B200_SIA_DEBUG=1: thread 0 of a few CTAs prints clock64 section timings
  // EVAL mode (svo_b200_sparse_residuals)
  int eval_level;
  const uint8_t* visible_in;
  float* ref_patch_out;
  float* residuals_out;
  uint8_t* in_image_out;
  double* Jres_out;
  double* chi2_out;
  long long* n_meas_out;
};

struct SiaShared {
  uint64_t mbar;
  double R[9];
  double t[3];
  double part[kMaxWarps * kPartK];
  double Hs[36];       // H_ of the current pass (scaled), full symmetric
  double Htot[36];     // sum over the level's visible set (scaled)
  Solver6 sol_cur;     // factorisation used for the current solve (slow path)
  Solver6 sol_tot;     // factorisation of Htot
  double x[8];
  double sums[kPartK];
  Pose model, old_model;
  double chi2_prev;
  int stop, done, slow, n_in_last, h_is_tot;
  int n_iters, sum_vis, sum_in, n_trace;
  unsigned mbar_phase;
  int cnt[kMaxWarps][2];
  double keep_chi2;
  int keep_n_in;
  long long tkx[2];
  long long tk[8];  // debug: cycles in [pre-parallel, pre-serial, pass, pass-reduce, serial, total]
};

// ---------------------------------------------------------------------------------------------
// Unaligned byte-row fetch: consecutive bytes starting at byte offset `off` from a 4-byte aligned
// base, as two (or three) aligned 32-bit loads + funnel shifts.
// ---------------------------------------------------------------------------------------------
template <bool SMEM>
__device__ __forceinline__ uint32_t ld_word(const uint8_t* base, int word_off) {
  if (SMEM) return *reinterpret_cast<const uint32_t*>(base + word_off);
  return __ldg(reinterpret_cast<const uint32_t*>(base + word_off));
}
template <bool SMEM>
__device__ __forceinline__ void fetch8(const uint8_t* base, int off, uint32_t& lo, uint32_t& hi) {
  const int a = off & ~3;
  const unsigned sh = (unsigned)(off & 3) * 8u;
  const uint32_t w0 = ld_word<SMEM>(base, a), w1 = ld_word<SMEM>(base, a + 4);
  lo = __funnelshift_r(w0, w1, sh);
  hi = w1 >> sh;  // byte 4 of the row in its low byte
}
template <bool SMEM>
__device__ __forceinline__ void fetch12(const uint8_t* base, int off, uint32_t& lo, uint32_t& hi) {
  const int a = off & ~3;
  const unsigned sh = (unsigned)(off & 3) * 8u;
  const uint32_t w0 = ld_word<SMEM>(base, a), w1 = ld_word<SMEM>(base, a + 4),
                 w2 = ld_word<SMEM>(base, a + 8);
  lo = __funnelshift_r(w0, w1, sh);
  hi = __funnelshift_r(w1, w2, sh);
}

// per-feature unscaled Jacobian rows: a = row0 of jacobian_xyz2uv, b = row1 (frame.h:116-138)
__device__ __forceinline__ void jac_rows(double x, double y, double zi, double (&a)[6], double (&b)[6]) {
  const double X = x * zi, Y = y * zi;
  a[0] = -zi; a[1] = 0.0; a[2] = X * zi; a[3] = X * Y; a[4] = -(1.0 + X * X); a[5] = Y;
  b[0] = 0.0; b[1] = -zi; b[2] = Y * zi; b[3] = 1.0 + Y * Y; b[4] = -(X * Y); b[5] = -X;
}
// Entry `idx` (0..20, upper triangle row-major) of Sxx aa^T + Sxy (ab^T + ba^T) + Syy bb^T.
template <int IDX>
__device__ __forceinline__ double h_entry(const double (&a)[6], const double (&b)[6], double sxx, double sxy, double syy) {
  constexpr int R = IDX < 6 ? 0 : IDX < 11 ? 1 : IDX < 15 ? 2 : IDX < 18 ? 3 : IDX < 20 ? 4 : 5;
  constexpr int BASE = R == 0 ? 0 : R == 1 ? 6 : R == 2 ? 11 : R == 3 ? 15 : R == 4 ? 18 : 20;
  constexpr int C = R + (IDX - BASE);
  return fma(sxx, a[R] * a[C], fma(sxy, fma(a[R], b[C], b[R] * a[C]), syy * (b[R] * b[C])));
}
template <int CHUNK, int J>
__device__ __forceinline__ double h_chunk_value(const double (&a)[6], const double (&b)[6], double sxx, double sxy,
                                                double syy, double cnt) {
  constexpr int IDX = CHUNK * 8 + J;
  if constexpr (IDX < 21) return h_entry<IDX>(a, b, sxx, sxy, syy);
  else if constexpr (IDX == 21) return cnt;
  else return 0.0;
}

// Block sum of the 21 unique H entries + one count over per-feature moments, once per level (and in the
// rare "slow path"): three transposed 8-value warp reductions computed chunk by chunk so that only ~8
// accumulators are live at a time (no register spills), one shared-memory hop, warp 0 adds the
// per-warp partials.  ONE __syncthreads.  `get(k, x, y, zi, sxx, sxy, syy, cnt)` yields feature k's data.
template <int FPT, class Get>
__device__ __forceinline__ void block_sum_h_to_warp0(Get get, SiaShared& s, int nwarps) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  auto do_chunk = [&](auto chunk_tag) {
    constexpr int CH = decltype(chunk_tag)::value;
    double v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = 0.0;
#pragma unroll
    for (int k = 0; k < FPT; ++k) {
      double x, y, zi, sxx, sxy, syy, cnt;
      get(k, x, y, zi, sxx, sxy, syy, cnt);
      double a[6], b[6];
      jac_rows(x, y, zi, a, b);
      v[0] += h_chunk_value<CH, 0>(a, b, sxx, sxy, syy, cnt);
      v[1] += h_chunk_value<CH, 1>(a, b, sxx, sxy, syy, cnt);
      v[2] += h_chunk_value<CH, 2>(a, b, sxx, sxy, syy, cnt);
      v[3] += h_chunk_value<CH, 3>(a, b, sxx, sxy, syy, cnt);
      v[4] += h_chunk_value<CH, 4>(a, b, sxx, sxy, syy, cnt);
      v[5] += h_chunk_value<CH, 5>(a, b, sxx, sxy, syy, cnt);
      v[6] += h_chunk_value<CH, 6>(a, b, sxx, sxy, syy, cnt);
      v[7] += h_chunk_value<CH, 7>(a, b, sxx, sxy, syy, cnt);
    }
    warp_reduce_t<8>(v);
    if ((lane & 3) == 0) s.part[warp * kPartK + CH * 8 + (lane >> 2)] = v[0];
  };
  do_chunk(std::integral_constant<int, 0>{});
  do_chunk(std::integral_constant<int, 1>{});
  do_chunk(std::integral_constant<int, 2>{});
  __syncthreads();
  if (warp == 0) {
    if (lane < kPartK) {
      double acc = 0.0;
      for (int wv = 0; wv < nwarps; ++wv) acc += s.part[wv * kPartK + lane];
      s.sums[lane] = acc;
    }
    __syncwarp();
  }
}

// Block sum of the 7 per-iteration doubles (6 Jres + chi2; slot 7 unused) and the two patch counts.
// ONE __syncthreads; on return (warp 0 only) s.sums[0..6] hold the totals, the counts are returned.
__device__ __forceinline__ void block_sum8_to_warp0(double (&v)[8], int n_in, int n_out, SiaShared& s, int nwarps,
                                                    int& tot_in, int& tot_out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  warp_reduce_t<8>(v);
  const int w_in = __reduce_add_sync(0xffffffffu, n_in), w_out = __reduce_add_sync(0xffffffffu, n_out);
  if ((lane & 3) == 0) s.part[warp * kPartK + (lane >> 2)] = v[0];
  if (lane == 0) { s.cnt[warp][0] = w_in; s.cnt[warp][1] = w_out; }
  __syncthreads();
  tot_in = tot_out = 0;
  if (warp == 0) {
    const int k = lane & 7;
    double acc = 0.0;
    for (int wv = lane >> 3; wv < nwarps; wv += 4) acc += s.part[wv * kPartK + k];
    acc += __shfl_xor_sync(0xffffffffu, acc, 8);
    acc += __shfl_xor_sync(0xffffffffu, acc, 16);
    if (lane < 8) s.sums[lane] = acc;
    const int ci = lane < nwarps ? s.cnt[lane][0] : 0, co = lane < nwarps ? s.cnt[lane][1] : 0;
    tot_in = __reduce_add_sync(0xffffffffu, ci);
    tot_out = __reduce_add_sync(0xffffffffu, co);
    __syncwarp();
  }
}

__device__ inline void publish_model(SiaShared& s) {
  qmatrix(s.model.q, s.R);
  s.t[0] = s.model.t[0]; s.t[1] = s.model.t[1]; s.t[2] = s.model.t[2];
}

// Thread 0: given Jres in s.x (already scaled/negated) and a factorisation of H_ (or S == nullptr
// when H_ is exactly zero), run the tail of one NLLSSolver::optimizeGaussNewton iteration [EXT]:
// solve, accept/rollback, update.  Everything is pulled into registers first (independent loads),
// the dependent chain is FMAs only.
// (eps / trace are passed by value: taking the kernel's parameter struct by reference here would force every thread to
// copy all of it to local memory at kernel entry -- 632 B/thread of stores that end up as DRAM write traffic)
static __device__ __noinline__ void gn_finish(SiaShared& s, double eps, svo_b200_sia_iter* trace, int trace_cap, const Solver6* S, int level,
                                          int iter, double chi2sum, int n_in) {
  const int n_meas = n_in * kPatchArea;
  const float chi2f = (float)chi2sum;
  const double new_chi2 = (double)(chi2f / (float)n_meas);  // sparse_img_align.cpp:242 (NaN if 0)
  double x[6];
  if (S == nullptr) {  // Eigen's LDLT of an all-zero matrix solves to x = 0
#pragma unroll
    for (int k = 0; k < 6; ++k) x[k] = 0.0;
  } else if (!S->pivoted) {
    const Fact6 F = S->F;
    double b[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) b[k] = s.x[k];
    fact6_solve(F, b, x);
  } else {
    ldlt6_solve(S->ldl, S->tr, s.x);
#pragma unroll
    for (int k = 0; k < 6; ++k) x[k] = s.x[k];
  }
  const Pose model = s.model;
  int stop = s.stop;
  if (isnan(x[0])) stop = 1;  // solve() == 0 (:248-250)
  int accepted, done = 0;
  Pose out;
  if ((iter > 0 && new_chi2 > s.chi2_prev) || stop) {
    out = s.old_model;  // rollback
    done = 1;
    accepted = 0;
  } else {
    double mx[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) mx[k] = -x[k];
    out = pose_mul_fast(model, se3_exp_fast(mx));  // T_new = T_old * exp(-x)  (:257)
    s.old_model = model;
    s.chi2_prev = new_chi2;
    accepted = 1;
    const double m = fmax(fmax(fmax(fabs(x[0]), fabs(x[1])), fmax(fabs(x[2]), fabs(x[3]))), fmax(fabs(x[4]), fabs(x[5])));
    if (m <= eps) done = 1;
  }
  s.model = out;
  qmatrix(out.q, s.R);
  s.t[0] = out.t[0]; s.t[1] = out.t[1]; s.t[2] = out.t[2];
  s.stop = stop;
  s.done = done;
  s.n_in_last = n_in;
  s.n_iters++;
  s.sum_in += n_in;
  if (trace) {
    if (s.n_trace < trace_cap) {
      svo_b200_sia_iter& r = trace[s.n_trace];
      r.level = level; r.iter = iter; r.accepted = accepted; r.n_meas = n_meas; r.chi2 = new_chi2;
      for (int k = 0; k < 6; ++k) r.x[k] = x[k];
      pose_to_rt12(out, r.T);
    }
    s.n_trace++;
  }
}

template <int FPT, bool EVAL, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB) sia_kernel(const SiaParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SiaShared& s = *reinterpret_cast<SiaShared*>(smem_raw);
  constexpr int S = MAXT * FPT;  // feature slots (== P.slots, checked on the host)
  float* pat_ref = reinterpret_cast<float*>(smem_raw + ((sizeof(SiaShared) + 15) & ~size_t(15)));
  float2* pat_dxy = reinterpret_cast<float2*>(pat_ref + kPatchArea * S);
  uint8_t* stage = reinterpret_cast<uint8_t*>(pat_dxy + kPatchArea * S);  // 16-byte aligned

  const SiaJob& job = P.jobs[blockIdx.x];
  const int tid = threadIdx.x, T = blockDim.x, nwarps = (T + 31) >> 5;
  const int N = job.n_feat;
  const int np = job.n_pad;
  const uint32_t blob_bytes = (uint32_t)np * 65u;  // <= 192*S: the patch arrays are free until the first level

  if (tid == 0) {
    mbar_init(&s.mbar, 1);
    fence_mbar_init();
    s.mbar_phase = 0;  // use k of the barrier completes phase parity k&1; the blob copy is use 0
    s.model = pose_from_rt12(job.T);
    s.old_model = s.model;
    s.chi2_prev = 1e10;  // NLLSSolver::reset() [EXT]
    s.stop = 0; s.done = 0; s.slow = 0; s.n_in_last = 0; s.h_is_tot = 0;
    s.n_iters = 0; s.sum_vis = 0; s.sum_in = 0; s.n_trace = 0;
    for (int k = 0; k < 8; ++k) s.tk[k] = 0;
    s.tkx[0] = s.tkx[1] = 0;
    s.tk[5] = clock64();
    for (int k = 0; k < 36; ++k) s.Hs[k] = 0.0;
    publish_model(s);
    // ---- TMA: the packed feature records of this pair, one bulk copy into the (idle) patch arrays
    fence_proxy_async();
    mbar_expect_tx(&s.mbar, blob_bytes);
    tma_bulk_g2s(pat_ref, job.blob, blob_bytes, &s.mbar);
  }
  __syncthreads();
  mbar_wait(&s.mbar, 0);
  const uint8_t* blob = reinterpret_cast<const uint8_t*>(pat_ref);
  const double* b_px = reinterpret_cast<const double*>(blob);
  const double* b_f = b_px + 2 * np;
  const double* b_pos = b_f + 3 * np;
  const uint8_t* b_hp = reinterpret_cast<const uint8_t*>(b_pos + 3 * np);

  // per-feature register state
  double fx_[FPT], fy_[FPT], fz_[FPT], fzi_[FPT];
  unsigned hp_mask = 0, vis_mask = 0, in_mask = 0;
#pragma unroll
  for (int k = 0; k < FPT; ++k) {
    const int i = tid + k * T;
    fx_[k] = fy_[k] = 0.0; fz_[k] = fzi_[k] = 1.0;
    if (i < N) {
      const double dxp = b_pos[3 * i] - job.ref_pos[0], dyp = b_pos[3 * i + 1] - job.ref_pos[1],
                   dzp = b_pos[3 * i + 2] - job.ref_pos[2];
      const double depth = sqrt(dxp * dxp + dyp * dyp + dzp * dzp);  // :107  |pos - ref_pos|
      fx_[k] = b_f[3 * i] * depth;                                    // :108  xyz_ref = f * depth
      fy_[k] = b_f[3 * i + 1] * depth;
      fz_[k] = b_f[3 * i + 2] * depth;
      fzi_[k] = 1.0 / fz_[k];
      if (b_hp[i]) hp_mask |= 1u << k;
      if (EVAL && P.visible_in[i]) vis_mask |= 1u << k;
    }
  }
  __syncthreads();  // everyone is done with the staged blob; the patch arrays may be written

  const int lvl_hi = EVAL ? P.eval_level : P.max_level;
  const int lvl_lo = EVAL ? P.eval_level : P.min_level;
  for (int level = lvl_hi; level >= lvl_lo; --level) {
    const int W = P.w[level], Hh = P.h[level];
    const float scale = 1.0f / (float)(1 << level);
    const double jscale = P.fx / (double)(1 << level);  // focal_length / (1<<level_)  (:140)
    const uint8_t* ref_img = job.ref_lvl[level];
    const uint8_t* cur_img = job.cur_lvl[level];

    // ---- TMA: stage the current level image when it fits the staging region -----------------
    const uint32_t img_bytes = ((uint32_t)(W * Hh) + 15u) & ~15u;
    const bool staged = !EVAL && img_bytes + 16u <= (uint32_t)P.stage_cap;
    if (staged && tid == 0) {
      s.mbar_phase ^= 1u;
      fence_proxy_async();
      mbar_expect_tx(&s.mbar, img_bytes);
      tma_bulk_g2s(stage, cur_img, img_bytes, &s.mbar);
    }

    long long tq0 = 0;
    if ((SVO_SIA_DEBUG && P.debug) && tid == 0) tq0 = clock64();
    // ---- precomputeReferencePatches (:84-145), one feature per thread -----------------------
    double m_sxx[FPT], m_sxy[FPT], m_syy[FPT], m_cnt[FPT];
#pragma unroll
    for (int k = 0; k < FPT; ++k) {
      m_sxx[k] = m_sxy[k] = m_syy[k] = m_cnt[k] = 0.0;
      const int slot = tid + k * T;
      // px is re-read from the pair's blob in global memory (L2) once per level instead of living in registers
      const double2 pxy = slot < N ? __ldg(reinterpret_cast<const double2*>(job.blob) + slot) : make_double2(-1e6, -1e6);
      const float u_ref = (float)(pxy.x * (double)scale);
      const float v_ref = (float)(pxy.y * (double)scale);
      const bool rng = u_ref >= 0.f && v_ref >= 0.f && u_ref < 1e6f && v_ref < 1e6f;  // else: outside, floor not needed
      float ufl = 0.f, vfl = 0.f;
      const int ui = rng ? floor_pos(u_ref, ufl) : -1, vi = rng ? floor_pos(v_ref, vfl) : -1;
      const bool ok = ((hp_mask >> k) & 1u) && ui - 3 >= 0 && vi - 3 >= 0 && ui + 3 < W && vi + 3 < Hh;
      if (ok) {
        vis_mask |= 1u << k;
        float wtl, wtr, wbl, wbr;
        bilin_weights(__fsub_rn(u_ref, ufl), __fsub_rn(v_ref, vfl), wtl, wtr, wbl, wbr);
        // 7x7 footprint rows vi-3..vi+3, cols ui-3..ui+3, streamed row by row to keep the register
        // footprint small: Bq row r (bilinear blends with top-left tap P[r][c]) needs footprint rows
        // r, r+1; patch row y needs Bq rows y, y+1, y+2.
        float pr0[7], pr1[7], b0[6], b1[6], b2[6];
        double sxx = 0, sxy = 0, syy = 0;
        // all seven footprint rows are requested before the first use: one exposed L2/HBM latency per level instead
        // of seven dependent ones
        uint32_t rlo[7], rhi[7];
#pragma unroll
        for (int r = 0; r < 7; ++r) fetch12<false>(ref_img, (vi - 3 + r) * W + (ui - 3), rlo[r], rhi[r]);
        auto load_row = [&](int r, float (&dst)[7]) {
          const uint32_t lo = rlo[r], hi = rhi[r];
          dst[0] = byte_to_float<0>(lo); dst[1] = byte_to_float<1>(lo); dst[2] = byte_to_float<2>(lo);
          dst[3] = byte_to_float<3>(lo); dst[4] = byte_to_float<0>(hi); dst[5] = byte_to_float<1>(hi);
          dst[6] = byte_to_float<2>(hi);
        };
        load_row(0, pr0);
#pragma unroll
        for (int r = 0; r < 6; ++r) {
          load_row(r + 1, pr1);
#pragma unroll
          for (int c = 0; c < 6; ++c) b2[c] = bilin(wtl, wtr, wbl, wbr, pr0[c], pr0[c + 1], pr1[c], pr1[c + 1]);
          if (r >= 2) {  // rows b0 (= Bq[y]), b1 (= Bq[y+1]), b2 (= Bq[y+2]) with y = r-2 are complete
            const int y = r - 2;
#pragma unroll
            for (int x = 0; x < 4; ++x) {
              const int p = y * 4 + x;
              const float val = b1[x + 1];
              const float dx = __fmul_rn(0.5f, __fsub_rn(b1[x + 2], b1[x]));
              const float dy = __fmul_rn(0.5f, __fsub_rn(b2[x + 1], b0[x + 1]));
              pat_ref[p * S + slot] = val;
              pat_dxy[p * S + slot] = make_float2(dx, dy);
              sxx = fma((double)dx, (double)dx, sxx);
              sxy = fma((double)dx, (double)dy, sxy);
              syy = fma((double)dy, (double)dy, syy);
            }
          }
#pragma unroll
          for (int c = 0; c < 6; ++c) { b0[c] = b1[c]; b1[c] = b2[c]; }
#pragma unroll
          for (int c = 0; c < 7; ++c) pr0[c] = pr1[c];
        }
        m_sxx[k] = sxx; m_sxy[k] = sxy; m_syy[k] = syy; m_cnt[k] = 1.0;
      } else if ((vis_mask >> k) & 1u) {
        // visible from a coarser level but failing here: the reference would keep the stale patch
        // and a zeroed Jacobian (jacobian_cache_.setZero() per level, :64).  Unreachable for
        // dyadic pyramids but kept bit-faithful.
#pragma unroll
        for (int p = 0; p < kPatchArea; ++p) pat_dxy[p * S + slot] = make_float2(0.f, 0.f);
        m_cnt[k] = 1.0;
      }
    }
    block_sum_h_to_warp0<FPT>(
        [&](int k, double& x, double& y, double& zi, double& sxx, double& sxy, double& syy, double& cnt) {
          x = fx_[k]; y = fy_[k]; zi = fzi_[k]; sxx = m_sxx[k]; sxy = m_sxy[k]; syy = m_syy[k]; cnt = m_cnt[k];
          // opaque to the optimiser: otherwise the level-invariant Jacobian rows are hoisted out of the level loop and
          // parked in local memory (17 doubles per thread, written once and re-read every level)
          asm volatile("" : "+d"(x), "+d"(y), "+d"(zi));
        },
        s, nwarps);
    if (tid == 0) {
      if (SVO_SIA_DEBUG && P.debug) s.tk[0] += clock64() - tq0;
      s.sum_vis += (int)s.sums[21];
      s.done = 0;
    }
    if (staged) mbar_wait(&s.mbar, s.mbar_phase);
    __syncthreads();
    // The scaling and LDL^T factorisation of this level's H is serial work nobody needs before the first solve: the
    // last thread of the block (a spare lane whenever the pair has fewer features than slots) does it while the other
    // warps already run the first residual pass; it rejoins its warp at the pass reduction, i.e. before the
    // iteration's first __syncthreads, after which thread 0 reads s.sol_tot.
    if (tid == T - 1) {
      long long tq1 = 0;
      if (SVO_SIA_DEBUG && P.debug) tq1 = clock64();
      const double s2 = jscale * jscale;
      int idx = 0;
      for (int r = 0; r < 6; ++r)
        for (int c = r; c < 6; ++c, ++idx) {
          const double v = s.sums[idx] * s2;
          s.Htot[r * 6 + c] = v;
          s.Htot[c * 6 + r] = v;
        }
      solver_factor(s.sol_tot, s.Htot);
      if (SVO_SIA_DEBUG && P.debug) s.tk[1] += clock64() - tq1;
    }

    // ---- Gauss-Newton iterations at this level ---------------------------------------------
    const int n_iter = EVAL ? 1 : P.n_iter;
    for (int iter = 0; iter < n_iter; ++iter) {
      long long ti0 = 0;
      if ((SVO_SIA_DEBUG && P.debug) && tid == 0) ti0 = clock64();
      double acc[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] = 0.0;
      int n_in_t = 0, n_out_t = 0;
      in_mask = 0;
#pragma unroll
      for (int k = 0; k < FPT; ++k) {
        if (!((vis_mask >> k) & 1u)) continue;
        const int slot = tid + k * T;
        long long tp0 = 0, tp1 = 0, tp2 = 0, tp3 = 0;
        if ((SVO_SIA_DEBUG && P.debug) && tid == 0) tp0 = clock64();
        const double x = fx_[k], y = fy_[k], z = fz_[k];
        const double xc = fma(s.R[0], x, fma(s.R[1], y, fma(s.R[2], z, s.t[0])));
        const double yc = fma(s.R[3], x, fma(s.R[4], y, fma(s.R[5], z, s.t[1])));
        const double zc = fma(s.R[6], x, fma(s.R[7], y, fma(s.R[8], z, s.t[2])));
        const double rz = fast_rcp(zc);
        const double ud = fma(P.fx, xc * rz, P.cx);  // [EXT] world2cam: fx * (x/z) + cx
        const double vd = fma(P.fy, yc * rz, P.cy);
        const float u_cur = __fmul_rn((float)ud, scale);  // .cast<float>() * scale (:183)
        const float v_cur = __fmul_rn((float)vd, scale);
        // negative / non-finite / huge coordinates can never pass the border test (:190)
        const bool rng = u_cur >= 0.f && v_cur >= 0.f && u_cur < 1e6f && v_cur < 1e6f;
        float ufl = 0.f, vfl = 0.f;
        const int ui = rng ? floor_pos(u_cur, ufl) : -1, vi = rng ? floor_pos(v_cur, vfl) : -1;
        const bool in = ui >= 0 && vi >= 0 && ui - 3 >= 0 && vi - 3 >= 0 && ui + 3 < W && vi + 3 < Hh;  // :190
        if (!in) {
          ++n_out_t;
          if (EVAL) {
            for (int p = 0; p < kPatchArea; ++p)
              P.residuals_out[(size_t)slot * kPatchArea + p] = __int_as_float(0x7fc00000);
          }
          continue;
        }
        in_mask |= 1u << k;
        ++n_in_t;
        if ((SVO_SIA_DEBUG && P.debug) && tid == 0) tp1 = clock64();
        float wtl, wtr, wbl, wbr;
        bilin_weights(__fsub_rn(u_cur, ufl), __fsub_rn(v_cur, vfl), wtl, wtr, wbl, wbr);
        uint32_t lo[5], hi[5];
#pragma unroll
        for (int r = 0; r < 5; ++r) {
          const int off = (vi - 2 + r) * W + (ui - 2);
          if (staged) fetch8<true>(stage, off, lo[r], hi[r]);
          else fetch8<false>(cur_img, off, lo[r], hi[r]);
        }
        float c2 = 0.f, gx = 0.f, gy = 0.f;
        float q0[5], q1[5];
        if ((SVO_SIA_DEBUG && P.debug) && tid == 0) { tp2 = clock64() + (long long)(lo[4] & 0u); }
        q0[0] = byte_to_float<0>(lo[0]); q0[1] = byte_to_float<1>(lo[0]); q0[2] = byte_to_float<2>(lo[0]);
        q0[3] = byte_to_float<3>(lo[0]); q0[4] = byte_to_float<0>(hi[0]);
#pragma unroll
        for (int yy = 0; yy < 4; ++yy) {
          q1[0] = byte_to_float<0>(lo[yy + 1]); q1[1] = byte_to_float<1>(lo[yy + 1]); q1[2] = byte_to_float<2>(lo[yy + 1]);
          q1[3] = byte_to_float<3>(lo[yy + 1]); q1[4] = byte_to_float<0>(hi[yy + 1]);
#pragma unroll
          for (int xx = 0; xx < 4; ++xx) {
            const int p = yy * 4 + xx;
            const float I = bilin(wtl, wtr, wbl, wbr, q0[xx], q0[xx + 1], q1[xx], q1[xx + 1]);
            const float res = __fsub_rn(I, pat_ref[p * S + slot]);
            const float2 g = pat_dxy[p * S + slot];
            c2 = fmaf(res, res, c2);  // chi2 += res*res*weight, weight == 1 (:222); order differs from the serial sum anyway
            gx = fmaf(g.x, res, gx);
            gy = fmaf(g.y, res, gy);
            if (EVAL) P.residuals_out[(size_t)slot * kPatchArea + p] = res;
          }
#pragma unroll
          for (int c = 0; c < 5; ++c) q0[c] = q1[c];
        }
        if ((SVO_SIA_DEBUG && P.debug) && tid == 0) { tp3 = clock64() + (long long)(__float_as_int(gx) & 0); s.tk[6] += tp1 - tp0; s.tk[7] += tp2 - tp1; s.tk[0] += 0; }
        const double zi = fzi_[k], X = x * zi, Y = y * zi, dgx = (double)gx, dgy = (double)gy;
        acc[0] = fma(-zi, dgx, acc[0]);
        acc[1] = fma(-zi, dgy, acc[1]);
        acc[2] = fma(zi, fma(X, dgx, Y * dgy), acc[2]);
        acc[3] = fma(X * Y, dgx, fma(fma(Y, Y, 1.0), dgy, acc[3]));
        acc[4] = fma(-fma(X, X, 1.0), dgx, fma(-(X * Y), dgy, acc[4]));
        acc[5] = fma(Y, dgx, fma(-X, dgy, acc[5]));
        acc[6] += (double)c2;
        if ((SVO_SIA_DEBUG && P.debug) && tid == 0) { s.tkx[0] += tp3 - tp2; s.tkx[1] += clock64() - tp3 + (long long)(acc[5] != acc[5]); }
      }
      int n_in, n_out;
      long long ti1 = 0;
      if ((SVO_SIA_DEBUG && P.debug) && tid == 0) ti1 = clock64();
      block_sum8_to_warp0(acc, n_in_t, n_out_t, s, nwarps, n_in, n_out);  // first __syncthreads of the iteration
      if (tid == 0) {
        long long ti2 = 0;
        if (SVO_SIA_DEBUG && P.debug) { ti2 = clock64(); s.tk[2] += ti1 - ti0; s.tk[3] += ti2 - ti1; }
        for (int k = 0; k < 6; ++k) s.x[k] = -(s.sums[k] * jscale);  // Jres_ = -sum J r
        s.keep_chi2 = s.sums[6];
        s.keep_n_in = n_in;
        s.slow = (EVAL || (n_out > 0 && n_in > 0)) ? 1 : 0;
        if (!s.slow) {
          if (n_in == 0) {  // H_ == 0 exactly: Eigen's LDLT yields x = 0
            for (int k = 0; k < 36; ++k) s.Hs[k] = 0.0;
            s.h_is_tot = 0;
            gn_finish(s, P.eps, P.trace, P.trace_cap, nullptr, level, iter, s.sums[6], n_in);
          } else {
            s.h_is_tot = 1;
            gn_finish(s, P.eps, P.trace, P.trace_cap, &s.sol_tot, level, iter, s.sums[6], n_in);
          }
        }
        if (SVO_SIA_DEBUG && P.debug) s.tk[4] += clock64() - ti2;
      }
      __syncthreads();
      if (s.slow) {
        // some visible patches fell outside the current image (or EVAL wants H): H_ = sum over the
        // patches that contributed in this pass.
        double q_sxx[FPT], q_sxy[FPT], q_syy[FPT];
#pragma unroll
        for (int k = 0; k < FPT; ++k) {
          q_sxx[k] = q_sxy[k] = q_syy[k] = 0.0;
          if (!((in_mask >> k) & 1u)) continue;
          const int slot = tid + k * T;
#pragma unroll
          for (int p = 0; p < kPatchArea; ++p) {
            const float2 g = pat_dxy[p * S + slot];
            const double dx = (double)g.x, dy = (double)g.y;
            q_sxx[k] = fma(dx, dx, q_sxx[k]);
            q_sxy[k] = fma(dx, dy, q_sxy[k]);
            q_syy[k] = fma(dy, dy, q_syy[k]);
          }
        }
        block_sum_h_to_warp0<FPT>(
            [&](int k, double& x, double& y, double& zi, double& sxx, double& sxy, double& syy, double& cnt) {
              x = fx_[k]; y = fy_[k]; zi = fzi_[k]; sxx = q_sxx[k]; sxy = q_sxy[k]; syy = q_syy[k]; cnt = 0.0;
              asm volatile("" : "+d"(x), "+d"(y), "+d"(zi));  // see the per-level call: no hoisting into local memory
            },
            s, nwarps);
        if (tid == 0) {
          const double s2 = jscale * jscale;
          int idx = 0;
          s.h_is_tot = 0;
          for (int r = 0; r < 6; ++r)
            for (int c = r; c < 6; ++c, ++idx) {
              const double v = s.sums[idx] * s2;
              s.Hs[r * 6 + c] = v;
              s.Hs[c * 6 + r] = v;
            }
          const double chi2_keep = s.keep_chi2;
          const int n_in_keep = s.keep_n_in;
          if (EVAL) {
            for (int k = 0; k < 6; ++k) P.Jres_out[k] = s.x[k];
            const float chi2f = (float)chi2_keep;
            *P.chi2_out = (double)(chi2f / (float)(n_in_keep * kPatchArea));
            *P.n_meas_out = (long long)n_in_keep * kPatchArea;
            s.n_in_last = n_in_keep;
            s.done = 1;
          } else {
            solver_factor(s.sol_cur, s.Hs);
            gn_finish(s, P.eps, P.trace, P.trace_cap, &s.sol_cur, level, iter, chi2_keep, n_in_keep);
          }
        }
        __syncthreads();
      }
      if (EVAL) {
#pragma unroll
        for (int k = 0; k < FPT; ++k) {
          const int i = tid + k * T;
          if (i < N) {
            P.in_image_out[i] = (in_mask >> k) & 1u;
            if (!((vis_mask >> k) & 1u))
              for (int p = 0; p < kPatchArea; ++p)
                P.residuals_out[(size_t)i * kPatchArea + p] = __int_as_float(0x7fc00000);
            for (int p = 0; p < kPatchArea; ++p)
              P.ref_patch_out[(size_t)i * kPatchArea + p] = pat_ref[p * S + i];
          }
        }
      }
      if (s.done) break;
    }
    __syncthreads();  // stage region / patches are rewritten by the next level
  }

  // ---- outputs ---------------------------------------------------------------------------------
#pragma unroll
  for (int k = 0; k < FPT; ++k) {
    const int i = tid + k * T;
    if (i < N && P.visible_out) P.visible_out[job.feat_off + i] = (vis_mask >> k) & 1u;
  }
  if (tid == 0) {
    if (P.T_out) pose_to_rt12(s.model, P.T_out + 12 * (size_t)blockIdx.x);
    if (P.H_out)
      for (int k = 0; k < 36; ++k) P.H_out[36 * (size_t)blockIdx.x + k] = s.h_is_tot ? s.Htot[k] : s.Hs[k];
    if (P.stats) {
      svo_b200_sia_stats st;
      st.n_iters = s.n_iters; st.sum_visible = s.sum_vis; st.sum_in_image = s.sum_in;
      st.n_tracked = s.n_in_last;  // n_meas_/patch_area_ of the last pass (:74)
      P.stats[blockIdx.x] = st;
    }
    if (P.n_trace) *P.n_trace = s.n_trace;
    if (SVO_SIA_DEBUG && P.debug && (blockIdx.x == 0 || blockIdx.x == gridDim.x / 2 || blockIdx.x == gridDim.x - 1))
      printf("[sia dbg] cta %d iters %d cycles: pre_par %lld pre_ser %lld pass %lld reduce %lld serial %lld total %lld | t0: geom %lld wts+ld %lld pix %lld acc %lld\n", blockIdx.x,
             s.n_iters, s.tk[0], s.tk[1], s.tk[2], s.tk[3], s.tk[4], (long long)clock64() - s.tk[5], s.tk[6], s.tk[7], s.tkx[0], s.tkx[1]);
  }
}

// =============================================================================================
// Host side
// =============================================================================================
struct SiaBatchState {
  int B = 0;
  int total_feat = 0;
  int max_feat = 0;
  SiaParams P;
  size_t in_bytes = 0;
  // device output offsets inside ctx->d_out
  size_t o_T = 0, o_H = 0, o_vis = 0, o_stats = 0, out_bytes = 0;
  int threads = 0, fpt = 1;
  size_t smem = 0;
  bool staged = false;
};

void sia_batch_free(svo_b200_ctx* ctx) {
  delete ctx->sia;
  ctx->sia = nullptr;
}

static int g_sia_spare = 0;     // env SVO_B200_SIA_SPARE=1 (experimental): 352-thread CTAs for <= 320 features, so that the last warp owns
                                // no feature and the per-level LDL^T truly overlaps the first residual pass
static int g_sia_minb = 2;      // tuning knobs (env SVO_B200_SIA_MINB / SVO_B200_SIA_STAGE_KB), read once
static int g_sia_stage_kb = 20;

static int pick_launch(svo_b200_ctx* ctx, int max_feat, int& threads, int& fpt, int& stage_cap, size_t& smem) {
  static bool env_read = false;
  if (!env_read) {
    env_read = true;
    if (const char* e = getenv("SVO_B200_SIA_MINB")) g_sia_minb = atoi(e) == 3 ? 3 : 2;
    if (const char* e = getenv("SVO_B200_SIA_SPARE")) g_sia_spare = atoi(e) != 0;
    if (const char* e = getenv("SVO_B200_SIA_STAGE_KB")) g_sia_stage_kb = atoi(e) > 0 ? atoi(e) : 20;
  }
  if (max_feat <= 512) fpt = 1;
  else if (max_feat <= 1024) fpt = 2;
  else if (max_feat <= 2048) fpt = 4;
  else return set_err(ctx, SVO_B200_ELIMIT, "sparse_img_align: %d features per pair > 2048", max_feat);
  threads = ((max_feat + fpt - 1) / fpt + 31) / 32 * 32;
  // the kernels are instantiated for MAXT in {320, 384, 512}; launching exactly MAXT threads makes the
  // slot count S = MAXT*FPT a compile-time constant (immediate shared-memory offsets)
  threads = fpt == 1 ? (threads <= 320 ? (g_sia_spare ? 352 : 320) : threads <= 384 ? 384 : 512) : 512;
  const size_t base = ((sizeof(SiaShared) + 15) & ~size_t(15)) + (size_t)3 * kPatchArea * threads * fpt * sizeof(float);
  const size_t budget = (size_t)ctx->max_smem_optin;
  if (base + 1024 > budget)
    return set_err(ctx, SVO_B200_ELIMIT, "sparse_img_align: %d features need %zu B of shared memory", max_feat, base);
  // staging region for the coarse current-level images (TMA); default 20 KB holds level >= 2 of 640x480
  size_t cap = (size_t)g_sia_stage_kb * 1024;
  if (base + cap > budget) cap = (budget - base) & ~size_t(15);
  stage_cap = (int)cap;
  smem = base + cap;
  return 0;
}

template <bool EVAL>
static int launch_sia(svo_b200_ctx* ctx, const SiaParams& P, int B, int threads, int fpt, size_t smem) {
  auto go = [&](auto kern) -> int {
    SVO_CUDA_CHECK(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // ask for the full shared-memory carveout so that two CTAs of ~90 KB fit one SM
    SVO_CUDA_CHECK(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout,
                                             (int)cudaSharedmemCarveoutMaxShared));
    kern<<<B, threads, smem, ctx->stream>>>(P);
    ctx->launches++;
    SVO_CUDA_CHECK(ctx, cudaGetLastError());
    return 0;
  };
  // <= 384 threads: cap registers so that two CTAs are resident per SM
  if (fpt == 1) {
    if (threads <= 320 && g_sia_minb == 3) return go(sia_kernel<1, EVAL, 320, 3>);
    if (threads <= 320) return go(sia_kernel<1, EVAL, 320, 2>);
    if (threads == 352) return go(sia_kernel<1, EVAL, 352, 2>);
    if (threads <= 384) return go(sia_kernel<1, EVAL, 384, 2>);
    return go(sia_kernel<1, EVAL, 512, 1>);
  }
  if (fpt == 2) return go(sia_kernel<2, EVAL, 512, 1>);
  return go(sia_kernel<4, EVAL, 512, 1>);
}

static inline int pad16(int n) { return (n + 15) / 16 * 16; }

// Pack one pair's features into the blob layout the kernel stages with TMA.
static void pack_blob(uint8_t* dst, int n, int np, const double* px, const double* f, const double* pos,
                      const uint8_t* hp) {
  double* d = reinterpret_cast<double*>(dst);
  memset(dst, 0, (size_t)np * 65);
  memcpy(d, px, sizeof(double) * 2 * n);
  memcpy(d + 2 * np, f, sizeof(double) * 3 * n);
  memcpy(d + 5 * np, pos, sizeof(double) * 3 * n);
  memcpy(dst + (size_t)np * 64, hp, n);
}

static int fill_common(svo_b200_ctx* ctx, SiaParams& P, const svo_b200_frame* fr, const svo_b200_camera* cam,
                       const svo_b200_sia_options* opt) {
  memset(&P, 0, sizeof(P));
  if (opt->max_level < opt->min_level || opt->min_level < 0 || opt->max_level >= fr->n_levels)
    return set_err(ctx, SVO_B200_EINVAL, "sparse_img_align: levels [%d,%d] outside the pyramid (%d levels)",
                   opt->min_level, opt->max_level, fr->n_levels);
  for (int l = 0; l < fr->n_levels; ++l) { P.w[l] = fr->w[l]; P.h[l] = fr->h[l]; }
  P.fx = cam->fx; P.fy = cam->fy; P.cx = cam->cx; P.cy = cam->cy;
  P.max_level = opt->max_level; P.min_level = opt->min_level; P.n_iter = opt->n_iter; P.eps = opt->eps;
  P.debug = getenv("SVO_B200_SIA_DEBUG") ? 1 : 0;
  return 0;
}

}  // namespace svo

using namespace svo;

extern "C" {

int svo_b200_sia_batch_stage(svo_b200_ctx* ctx, int B, const svo_b200_frame* const* ref,
                             const svo_b200_frame* const* cur, const svo_b200_camera* cam,
                             const svo_b200_sia_options* opt, const double* T, const int* feat_offset,
                             const double* px, const double* f, const double* point_pos,
                             const uint8_t* has_point, const double* ref_pos) {
  if (!ctx || B <= 0 || !ref || !cur || !cam || !opt || !T || !feat_offset || !ref_pos)
    return set_err(ctx, SVO_B200_EINVAL, "sia_batch_stage: bad arguments");
  cudaSetDevice(ctx->device);
  if (!ctx->sia) ctx->sia = new SiaBatchState();
  SiaBatchState& st = *ctx->sia;
  st.staged = false;
  st.B = B;
  st.total_feat = feat_offset[B] - feat_offset[0];
  st.max_feat = 0;
  for (int b = 0; b < B; ++b) {
    const int n = feat_offset[b + 1] - feat_offset[b];
    if (n < 0) return set_err(ctx, SVO_B200_EINVAL, "sia_batch_stage: feat_offset not monotone");
    if (n > st.max_feat) st.max_feat = n;
    if (!ref[b] || !cur[b] || ref[b]->n_levels != ref[0]->n_levels || ref[b]->width != ref[0]->width ||
        ref[b]->height != ref[0]->height || cur[b]->width != ref[0]->width ||
        cur[b]->height != ref[0]->height || cur[b]->n_levels != ref[0]->n_levels)
      return set_err(ctx, SVO_B200_EINVAL, "sia_batch_stage: all frames of a batch must share one geometry");
  }
  if (st.total_feat > 0 && (!px || !f || !point_pos || !has_point))
    return set_err(ctx, SVO_B200_EINVAL, "sia_batch_stage: NULL feature arrays");
  int rc = fill_common(ctx, st.P, ref[0], cam, opt);
  if (rc) return rc;
  int stage_cap = 0;
  rc = pick_launch(ctx, st.max_feat, st.threads, st.fpt, stage_cap, st.smem);
  if (rc) return rc;
  st.P.stage_cap = stage_cap;
  st.P.slots = st.threads * st.fpt;

  // input staging: [jobs B][blobs]
  Carver cin;
  const size_t o_jobs = cin.take(sizeof(SiaJob) * (size_t)B);
  std::vector<size_t> o_blob(B);
  for (int b = 0; b < B; ++b) {
    const int n = feat_offset[b + 1] - feat_offset[b];
    o_blob[b] = cin.take((size_t)pad16(n) * 65 + 16, 128);
  }
  st.in_bytes = cin.off;
  if ((rc = ensure_host(ctx, ctx->h_in, st.in_bytes))) return rc;
  if ((rc = ensure_dev(ctx, ctx->d_in, st.in_bytes))) return rc;
  // a previous async copy out of the pinned buffer must be finished before it is rewritten
  SVO_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  uint8_t* hin = static_cast<uint8_t*>(ctx->h_in.p);
  uint8_t* din = static_cast<uint8_t*>(ctx->d_in.p);
  SiaJob* jobs = reinterpret_cast<SiaJob*>(hin + o_jobs);
  for (int b = 0; b < B; ++b) {
    SiaJob& j = jobs[b];
    memset(&j, 0, sizeof(j));
    for (int l = 0; l < ref[b]->n_levels; ++l) { j.ref_lvl[l] = ref[b]->lvl(l); j.cur_lvl[l] = cur[b]->lvl(l); }
    const int o = feat_offset[b], n = feat_offset[b + 1] - o;
    j.n_feat = n;
    j.n_pad = pad16(n);
    j.feat_off = o - feat_offset[0];
    j.blob = din + o_blob[b];
    memcpy(j.T, T + 12 * (size_t)b, sizeof(double) * 12);
    memcpy(j.ref_pos, ref_pos + 3 * (size_t)b, sizeof(double) * 3);
    if (n > 0) pack_blob(hin + o_blob[b], n, j.n_pad, px + 2 * (size_t)o, f + 3 * (size_t)o,
                         point_pos + 3 * (size_t)o, has_point + o);
  }
  SVO_CUDA_CHECK(ctx, cudaMemcpyAsync(din, hin, st.in_bytes, cudaMemcpyHostToDevice, ctx->stream));

  Carver co;
  st.o_T = co.take(sizeof(double) * 12 * (size_t)B);
  st.o_H = co.take(sizeof(double) * 36 * (size_t)B);
  st.o_stats = co.take(sizeof(svo_b200_sia_stats) * (size_t)B);
  st.o_vis = co.take((size_t)st.total_feat + 16);
  st.out_bytes = co.off;
  if ((rc = ensure_dev(ctx, ctx->d_out, st.out_bytes))) return rc;
  if ((rc = ensure_host(ctx, ctx->h_out, st.out_bytes))) return rc;
  uint8_t* dout = static_cast<uint8_t*>(ctx->d_out.p);
  st.P.jobs = reinterpret_cast<const SiaJob*>(din + o_jobs);
  st.P.T_out = reinterpret_cast<double*>(dout + st.o_T);
  st.P.H_out = reinterpret_cast<double*>(dout + st.o_H);
  st.P.stats = reinterpret_cast<svo_b200_sia_stats*>(dout + st.o_stats);
  st.P.visible_out = dout + st.o_vis;
  st.staged = true;
  return 0;
}

int svo_b200_sia_batch_run(svo_b200_ctx* ctx) {
  if (!ctx || !ctx->sia || !ctx->sia->staged) return set_err(ctx, SVO_B200_EINVAL, "sia_batch_run: nothing staged");
  cudaSetDevice(ctx->device);
  SiaBatchState& st = *ctx->sia;
  return launch_sia<false>(ctx, st.P, st.B, st.threads, st.fpt, st.smem);
}

int svo_b200_sia_batch_fetch(svo_b200_ctx* ctx, double* T_out, uint8_t* visible_out, double* H_out,
                             svo_b200_sia_stats* stats_out) {
  if (!ctx || !ctx->sia || !ctx->sia->staged) return set_err(ctx, SVO_B200_EINVAL, "sia_batch_fetch: nothing staged");
  cudaSetDevice(ctx->device);
  SiaBatchState& st = *ctx->sia;
  SVO_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->h_out.p, ctx->d_out.p, st.out_bytes, cudaMemcpyDeviceToHost, ctx->stream));
  SVO_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  const uint8_t* h = static_cast<const uint8_t*>(ctx->h_out.p);
  if (T_out) memcpy(T_out, h + st.o_T, sizeof(double) * 12 * (size_t)st.B);
  if (H_out) memcpy(H_out, h + st.o_H, sizeof(double) * 36 * (size_t)st.B);
  if (stats_out) memcpy(stats_out, h + st.o_stats, sizeof(svo_b200_sia_stats) * (size_t)st.B);
  if (visible_out) memcpy(visible_out, h + st.o_vis, (size_t)st.total_feat);
  return 0;
}

int svo_b200_sparse_img_align(svo_b200_ctx* ctx, const svo_b200_frame* ref, const svo_b200_frame* cur,
                              const svo_b200_camera* cam, const svo_b200_sia_options* opt,
                              double* T_io, const double* px, const double* f, const double* point_pos,
                              const uint8_t* has_point, const double* ref_pos, int N,
                              uint8_t* visible_out, double* H_out, svo_b200_sia_stats* stats_out,
                              svo_b200_sia_iter* trace_out, int trace_cap, int* n_trace_out) {
  if (!ctx || !ref || !cur || !cam || !opt || !T_io || !ref_pos || N < 0)
    return set_err(ctx, SVO_B200_EINVAL, "sparse_img_align: bad arguments");
  if (N == 0) {  // "SparseImgAlign: no features to track!" -> return 0 (sparse_img_align.cpp:47-51)
    if (stats_out) memset(stats_out, 0, sizeof(*stats_out));
    if (H_out) memset(H_out, 0, sizeof(double) * 36);
    if (n_trace_out) *n_trace_out = 0;
    return 0;
  }
  const int off[2] = {0, N};
  int rc = svo_b200_sia_batch_stage(ctx, 1, &ref, &cur, cam, opt, T_io, off, px, f, point_pos, has_point, ref_pos);
  if (rc) return rc;
  SiaBatchState& st = *ctx->sia;
  size_t o_tr = 0, o_ntr = 0;
  if (trace_out && trace_cap > 0) {
    Carver c;
    o_tr = c.take(sizeof(svo_b200_sia_iter) * (size_t)trace_cap);
    o_ntr = c.take(sizeof(int));
    if ((rc = ensure_dev(ctx, ctx->d_scratch, c.off))) return rc;
    uint8_t* ds = static_cast<uint8_t*>(ctx->d_scratch.p);
    st.P.trace = reinterpret_cast<svo_b200_sia_iter*>(ds + o_tr);
    st.P.trace_cap = trace_cap;
    st.P.n_trace = reinterpret_cast<int*>(ds + o_ntr);
  }
  if ((rc = svo_b200_sia_batch_run(ctx))) return rc;
  if ((rc = svo_b200_sia_batch_fetch(ctx, T_io, visible_out, H_out, stats_out))) return rc;
  if (trace_out && trace_cap > 0) {
    int ntr = 0;
    uint8_t* ds = static_cast<uint8_t*>(ctx->d_scratch.p);
    SVO_CUDA_CHECK(ctx, cudaMemcpy(&ntr, ds + o_ntr, sizeof(int), cudaMemcpyDeviceToHost));
    const int ncopy = ntr < trace_cap ? ntr : trace_cap;
    if (ncopy > 0)
      SVO_CUDA_CHECK(ctx, cudaMemcpy(trace_out, ds + o_tr, sizeof(svo_b200_sia_iter) * (size_t)ncopy, cudaMemcpyDeviceToHost));
    if (n_trace_out) *n_trace_out = ntr;
  } else if (n_trace_out) {
    *n_trace_out = 0;
  }
  st.P.trace = nullptr;
  st.P.n_trace = nullptr;
  return 0;
}

int svo_b200_sparse_residuals(svo_b200_ctx* ctx, const svo_b200_frame* ref, const svo_b200_frame* cur,
                              const svo_b200_camera* cam, int level, const double* T, const double* px,
                              const double* f, const double* point_pos, const uint8_t* has_point,
                              const double* ref_pos, int N, uint8_t* visible_io, float* ref_patch_out,
                              float* residuals_out, uint8_t* in_image_out, double* H_out, double* Jres_out,
                              double* chi2_out, int64_t* n_meas_out) {
  if (!ctx || !ref || !cur || !cam || !T || !ref_pos || N <= 0 || !visible_io)
    return set_err(ctx, SVO_B200_EINVAL, "sparse_residuals: bad arguments");
  svo_b200_sia_options opt = {level, level, 1, 1e-6};
  const int off[2] = {0, N};
  int rc = svo_b200_sia_batch_stage(ctx, 1, &ref, &cur, cam, &opt, T, off, px, f, point_pos, has_point, ref_pos);
  if (rc) return rc;
  SiaBatchState& st = *ctx->sia;
  Carver c;
  const size_t o_vin = c.take(N), o_rp = c.take(sizeof(float) * 16 * (size_t)st.P.slots),
               o_res = c.take(sizeof(float) * 16 * (size_t)st.P.slots), o_in = c.take(N),
               o_j = c.take(sizeof(double) * 6), o_c = c.take(sizeof(double)), o_n = c.take(sizeof(long long));
  if ((rc = ensure_dev(ctx, ctx->d_scratch, c.off))) return rc;
  uint8_t* ds = static_cast<uint8_t*>(ctx->d_scratch.p);
  SVO_CUDA_CHECK(ctx, cudaMemcpyAsync(ds + o_vin, visible_io, N, cudaMemcpyHostToDevice, ctx->stream));
  SVO_CUDA_CHECK(ctx, cudaMemsetAsync(ds + o_res, 0xff, sizeof(float) * 16 * (size_t)st.P.slots, ctx->stream));
  st.P.eval_level = level;
  st.P.visible_in = ds + o_vin;
  st.P.ref_patch_out = reinterpret_cast<float*>(ds + o_rp);
  st.P.residuals_out = reinterpret_cast<float*>(ds + o_res);
  st.P.in_image_out = ds + o_in;
  st.P.Jres_out = reinterpret_cast<double*>(ds + o_j);
  st.P.chi2_out = reinterpret_cast<double*>(ds + o_c);
  st.P.n_meas_out = reinterpret_cast<long long*>(ds + o_n);
  if ((rc = launch_sia<true>(ctx, st.P, 1, st.threads, st.fpt, st.smem))) return rc;
  double Tdummy[12];
  if ((rc = svo_b200_sia_batch_fetch(ctx, Tdummy, visible_io, H_out, nullptr))) return rc;
  if (ref_patch_out) SVO_CUDA_CHECK(ctx, cudaMemcpy(ref_patch_out, ds + o_rp, sizeof(float) * 16 * (size_t)N, cudaMemcpyDeviceToHost));
  if (residuals_out) SVO_CUDA_CHECK(ctx, cudaMemcpy(residuals_out, ds + o_res, sizeof(float) * 16 * (size_t)N, cudaMemcpyDeviceToHost));
  if (in_image_out) SVO_CUDA_CHECK(ctx, cudaMemcpy(in_image_out, ds + o_in, N, cudaMemcpyDeviceToHost));
  if (Jres_out) SVO_CUDA_CHECK(ctx, cudaMemcpy(Jres_out, ds + o_j, sizeof(double) * 6, cudaMemcpyDeviceToHost));
  if (chi2_out) SVO_CUDA_CHECK(ctx, cudaMemcpy(chi2_out, ds + o_c, sizeof(double), cudaMemcpyDeviceToHost));
  if (n_meas_out) {
    long long nm = 0;
    SVO_CUDA_CHECK(ctx, cudaMemcpy(&nm, ds + o_n, sizeof(long long), cudaMemcpyDeviceToHost));
    *n_meas_out = nm;
  }
  st.staged = false;
  return 0;
}

}  // extern "C"
