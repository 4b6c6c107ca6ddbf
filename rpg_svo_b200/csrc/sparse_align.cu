// rpg_svo_b200/csrc/sparse_align.cu -- svo::SparseImgAlign on sm_100a.
//
// Replaces svo/src/sparse_img_align.cpp:43-258 (run / precomputeReferencePatches /
// computeResiduals / solve / update) together with the Gauss-Newton driver of
// vk::NLLSSolver<6,SE3>::optimizeGaussNewton [EXT] that the class derives from.
//
// Design (B200-first, not a translation):
//   * one CTA per frame pair -- or, for small batches, one thread-block CLUSTER per pair with the
//     features split over its CTAs -- runs the WHOLE coarse-to-fine loop on the device: no host
//     round trip per Gauss-Newton iteration, batches of pairs fill the 148 SMs.
//   * one thread owns one feature (FPT features when N > blockDim): the 4x4 reference patch, and its two
//     gradient images, live in shared memory in pixel-major (SoA) order so a warp's accesses are conflict
//     free; the feature's bearing/depth state lives in registers, or -- throughput geometry -- in shared memory.
//   * inverse-compositional structure is exploited: the per-pixel Jacobian is
//     J_p = dx_p * a + dy_p * b with a, b per-FEATURE 6-vectors, hence
//        sum_p J_p J_p^T = Sxx aa^T + Sxy (ab^T + ba^T) + Syy bb^T     (pose independent)
//        sum_p J_p r_p   = (sum dx_p r_p) a + (sum dy_p r_p) b
//     so the 6x6 normal matrix is reduced and factorised ONCE per level (and re-formed only in the
//     iterations where some patch leaves the current image), and an iteration costs 3 f32 FMAs per
//     pixel plus ~15 f64 FMAs per feature instead of the reference's 27 f64 MACs per pixel.
//   * the current image is staged in shared memory where the instantiation has room: whole coarse levels with
//     one TMA bulk copy (cp.async.bulk + mbarrier), at the fine levels a 16x8-byte window around each
//     feature's projection with cp.async (once per level; a footprint that drifts out of its window
//     falls back to global loads).  The packed feature records of the pair arrive by TMA as well.
//   * ONE block barrier pair per iteration: each warp folds its partial sums (6 Jres + chi2 + counts)
//     with a transposed shuffle reduction and parks them in a double-buffered shared array; after
//     the barrier warp 0 adds the per-warp partials and runs the 6x6 substitution, SE3 exp and the
//     accept / rollback decision in registers (all lanes redundantly), and publishes the pose in shared memory.
//     Cluster geometry: every warp of every CTA receives all partials and runs that tail redundantly
//     (bit-identical everywhere); its "upfront" variant prepares the reference patches, H and LDL^T of ALL
//     levels before the first iteration and exchanges the per-iteration sums with st.async + complete_tx on
//     the receivers' mbarriers instead of DSMEM stores + barrier.cluster.
//   * the throughput instantiation is kept SMALL: staging modes and features it cannot use are compiled out
//     and its loops over a thread's features are loops, not unrolled copies -- three CTAs in different phases
//     share one instruction cache (ncu r02h -> r02j: no-instruction stalls 16 % -> 3 %, +20 % throughput).
// Precision follows the reference per quantity: f32 interpolation/residual/chi2, f64 geometry and
// normal equations.
#include <cstddef>
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
#if SVO_SIA_DEBUG
#define SIA_DBG(...) __VA_ARGS__
#else
#define SIA_DBG(...)
#endif
constexpr int kPatchArea = 16;
constexpr int kPartK = 24;     // widest block reduction: 21 unique H entries + 1 count, padded to 16 + 8
constexpr int kWinRows = 8;    // per-feature window of the current image: 8 rows x 16 bytes
constexpr int kWinBytes = kWinRows * 16;  // 128 B per feature slot

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

// ---- single-stream feature split over GPUs: the per-iteration sums of one pair are exchanged between the
// ranks' kernels through peer memory (NVLink P2P stores into every rank's exchange buffer, system-scope flags), so the
// whole coarse-to-fine loop still runs inside ONE kernel per GPU -- no host round trip, no collective library call.
constexpr int kMaxSplit = 8;
struct XgSlot {  // what one rank publishes for one exchange
  double v[kPartK];
  int cnt[2];
  unsigned seq;  // sequence number of the exchange this slot holds (written last, release)
  unsigned pad_;
};
struct XgPair {  // per frame pair, in every rank's exchange buffer
  XgSlot slot[2][kMaxSplit];  // [parity of the exchange][publishing rank]
  unsigned xseq;              // exchanges completed so far (persists across launches; identical on all ranks)
  unsigned err;               // sticky: a peer did not arrive within the timeout
  unsigned pad_[2];
};
struct XgParams {
  int rank, world;
  XgPair* peer[kMaxSplit];  // rank r's exchange buffer as mapped in this process (peer[rank] = our own)
};

struct SiaParams {
  const SiaJob* jobs;
  XgParams xg;
  int w[SVO_B200_MAX_LEVELS], h[SVO_B200_MAX_LEVELS];
  CamDev cam;
  int max_level, min_level, n_iter;
  double eps;
  int stage_cap;  // bytes of the staging region in shared memory (TMA image / cp.async windows)
  int slots;      // blockDim * FPT feature slots per CTA (patch arrays are [3][16][slots])
  int use_windows, use_prefetch;
  int async_xchg;  // upfront cluster variant: per-iteration sums by st.async + mbarrier instead of DSMEM stores + barrier.cluster
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

struct SiaState {  // model_ / old_model of NLLSSolver::optimizeGaussNewton [EXT]
  Pose model, old_model;
};

// Shared-memory control block of one CTA.  NWC = warps of this CTA, CS = CTAs of the pair (cluster size).
template <int NWC, int CS>
struct SiaSharedT {
  static constexpr int kPairWarps = NWC * CS;
  uint64_t mbar;
  uint64_t xbar[2];  // cluster geometry, asynchronous exchange: one mbarrier per parity of the running iteration counter
  // per-warp partial sums of one residual pass (6 Jres + chi2, slot 7 unused) for every warp of the
  // pair (all CTAs of the cluster), double-buffered by the parity of the running iteration counter
  double part[2][kPairWarps][8];
  int cnt[2][kPairWarps][2];
  double hpart[NWC * kPartK];                   // per-warp partials of the 24-value H reduction
  double hsum_cta[CS > 1 ? CS : 1][kPartK];     // per-CTA H sums (cluster variant: written remotely)
  double sums[kPartK];                          // H totals of the pair
  double Hs[36];                                // H_ of the current pass (scaled), full symmetric
  double Htot[36];                              // sum over the level's visible set (scaled)
  Solver6 sol_cur;                              // factorisation used for the current solve (slow path)
  Solver6 sol_tot;                              // factorisation of Htot
  SiaState st[2];                               // double-buffered like `part`
  int h_is_tot, n_iters, sum_vis, sum_in, n_in_last, n_trace;  // thread 0 of CTA rank 0 only
  unsigned xg_seq;     // feature split over GPUs: exchanges completed (warp 0) ...
  unsigned xg_failed;  // ... and "an exchange timed out" (must directly follow xg_seq)
  alignas(16) double pub[12];  // CS == 1: the pose (R row-major, t) warp 0 publishes after its Gauss-Newton tail
  int pub_done, pub_slow;
#if SVO_SIA_DEBUG
  long long tkx[4];
  long long tk[8];  // debug: cycles in [level setup, pass, reduce, tail, total]
#endif
};

// "Upfront" variant (small batches: a thread-block cluster per pair, every CTA with an SM of its own): the reference patches,
// H sums and factorisations of ALL pyramid levels -- none of which depends on the pose -- are computed before the first
// Gauss-Newton iteration, with one cluster exchange for all levels, so a level starts with nothing but the staging of its
// current image.  Per-level results live here (behind the control block) and in per-level patch arrays.
template <int NWC, int CS>
struct SiaUpT {
  double hpart[SVO_B200_MAX_LEVELS][NWC * kPartK];      // per-warp partials of every level
  double hsum_cta[SVO_B200_MAX_LEVELS][CS][kPartK];     // per-CTA sums of every level (written remotely)
  double sums[SVO_B200_MAX_LEVELS][kPartK];
  double Htot[SVO_B200_MAX_LEVELS][36];
  Solver6 sol[SVO_B200_MAX_LEVELS];
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
// 7 consecutive bytes at byte offset `off` from an 8-byte aligned global base: one aligned 64-bit load, plus the next
// one only when the span crosses it -- 1.75 memory requests per row on average instead of 3 (the residual loops are
// bound by the number of uncoalesced requests the L1 can take, not by bytes)
__device__ __forceinline__ void fetch7_g64(const uint8_t* base, int off, uint32_t& lo, uint32_t& hi) {
  const int a = off & ~7, p = off & 7;
  const uint2 A = __ldg(reinterpret_cast<const uint2*>(base + a));
  uint2 B = make_uint2(0u, 0u);
  if (p > 1) B = __ldg(reinterpret_cast<const uint2*>(base + a + 8));
  const bool k = p >= 4;
  const uint32_t x0 = k ? A.y : A.x, x1 = k ? B.x : A.y, x2 = k ? B.y : B.x;
  const unsigned sh = (unsigned)(p & 3) * 8u;
  lo = __funnelshift_r(x0, x1, sh);
  hi = __funnelshift_r(x1, x2, sh);
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

// ---- cluster helpers (CS == 1: plain CTA, everything below folds to local shared memory) ------------
__device__ __forceinline__ unsigned cluster_rank() {
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
template <int CS>
__device__ __forceinline__ void pair_sync() {  // all threads of the pair: the CTA, or every CTA of its cluster
  if constexpr (CS == 1) {
    __syncthreads();
  } else {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
}
// store a double into the same shared-memory variable of CTA `rank` of the cluster (DSMEM)
__device__ __forceinline__ void st_cluster_f64(double* local_ptr, unsigned rank, double v) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(local_ptr)), "r"(rank));
  asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(remote), "d"(v) : "memory");
}
__device__ __forceinline__ void st_cluster_v2s32(int* local_ptr, unsigned rank, int a, int b) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(local_ptr)), "r"(rank));
  asm volatile("st.shared::cluster.v2.s32 [%0], {%1, %2};" ::"r"(remote), "r"(a), "r"(b) : "memory");
}

// asynchronous remote store that also counts its bytes on an mbarrier of the SAME remote CTA (st.async + complete_tx): data
// and "it has arrived" travel together, one DSMEM hop, and the receiver sleeps on its own mbarrier instead of a cluster barrier
__device__ __forceinline__ uint32_t cluster_addr(const void* local_ptr, unsigned rank) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(local_ptr)), "r"(rank));
  return remote;
}
__device__ __forceinline__ void st_async_f64(uint32_t remote_addr, double v, uint32_t remote_mbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.f64 [%0], %1, [%2];" ::"r"(remote_addr), "d"(v), "r"(remote_mbar)
               : "memory");
}
__device__ __forceinline__ void st_async_v2s32(uint32_t remote_addr, int a, int b, uint32_t remote_mbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.s32 [%0], {%1, %2}, [%3];" ::"r"(remote_addr), "r"(a), "r"(b),
               "r"(remote_mbar)
               : "memory");
}

// ---- system-scope accesses to (peer) global memory
__device__ __forceinline__ void st_sys_f64(double* p, double v) { asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory"); }
__device__ __forceinline__ void st_sys_s32(int* p, int v) { asm volatile("st.relaxed.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void st_release_sys_u32(unsigned* p, unsigned v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned ld_acquire_sys_u32(const unsigned* p) { unsigned v; asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ double ld_sys_f64(const double* p) { double v; asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ int ld_sys_s32(const int* p) { int v; asm volatile("ld.relaxed.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }

// All-reduce (sum) of K <= 24 doubles + two counts of one pair over the ranks of a feature split, executed by ONE warp of
// the pair's CTA on every rank: lane k holds value k.  Every rank stores its values into slot [parity][rank] of EVERY
// rank's exchange buffer (peer memory), then its sequence number (release); it then waits until all slots of its own
// buffer carry that number (acquire) and adds them in rank order, so all ranks obtain bit-identical sums.  A slot is
// rewritten two exchanges later, by which time every peer has consumed it (a rank cannot run more than one exchange
// ahead).  A peer that never arrives trips a ~2 s timeout: the sticky error flag is set and the kernel finishes with
// meaningless numbers instead of hanging the GPU.
template <int K>
__device__ __forceinline__ void xg_allreduce(const XgParams& X, int pair, unsigned* seq_smem, double& val, int& c0, int& c1) {
  const int lane = threadIdx.x & 31;
  const unsigned failed = seq_smem[1];  // a previous exchange of this launch already timed out: do not wait again
  const unsigned xseq = *seq_smem;
  const unsigned par = xseq & 1u, want = xseq + 1u;
  for (int r = 0; r < X.world; ++r) {
    XgSlot* dst = &X.peer[r][pair].slot[par][X.rank];
    if (lane < K) st_sys_f64(&dst->v[lane], val);
    if (lane == 0) { st_sys_s32(&dst->cnt[0], c0); st_sys_s32(&dst->cnt[1], c1); }
  }
  __threadfence_system();
  __syncwarp();
  if (lane == 0)
    for (int r = 0; r < X.world; ++r) st_release_sys_u32(&X.peer[r][pair].slot[par][X.rank].seq, want);
  XgPair* own = &X.peer[X.rank][pair];
  bool ok = true;
  if (lane < X.world && !failed) {
    const long long t0 = clock64();
    while (ld_acquire_sys_u32(&own->slot[par][lane].seq) != want) {
      if (clock64() - t0 > 4000000000LL) { ok = false; break; }
    }
  }
  ok = __all_sync(0xffffffffu, ok);
  __threadfence_system();
  double acc = 0.0;
  int a0 = 0, a1 = 0;
  for (int r = 0; r < X.world; ++r) {
    if (lane < K) acc += ld_sys_f64(&own->slot[par][r].v[lane]);
    a0 += ld_sys_s32(&own->slot[par][r].cnt[0]);
    a1 += ld_sys_s32(&own->slot[par][r].cnt[1]);
  }
  val = acc; c0 = a0; c1 = a1;
  if (lane == 0) {
    *seq_smem = want;
    if (!ok) { own->err = 1u; seq_smem[1] = 1u; }
  }
  __syncwarp();
}

// Per-warp part of the H reduction: the 24 values (21 unique entries of sum_f Sxx aa^T + Sxy (ab^T + ba^T) + Syy bb^T, the
// feature count, two pads) of this warp's features, into dst[0..23] (shared memory): three transposed 8-value warp
// reductions computed chunk by chunk so that only ~8 accumulators are live at a time.
// `get(k, x, y, zi, sxx, sxy, syy, cnt)` yields feature k's data.
template <int FPT, bool ROLL, class Get>
__device__ __forceinline__ void warp_h_partials(Get get, double* dst) {
  const int lane = threadIdx.x & 31;
  auto do_chunk = [&](auto chunk_tag) {
    constexpr int CH = decltype(chunk_tag)::value;
    double v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = 0.0;
#pragma unroll(ROLL ? 1 : FPT)
    for (int k = 0; k < FPT; ++k) {  // ROLL: a loop, not FPT copies (instruction-cache footprint of the throughput geometry)
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
    if ((lane & 3) == 0) dst[CH * 8 + (lane >> 2)] = v[0];
  };
  do_chunk(std::integral_constant<int, 0>{});
  do_chunk(std::integral_constant<int, 1>{});
  do_chunk(std::integral_constant<int, 2>{});
}

// Sum of the 21 unique H entries + one count over the per-feature moments of the whole pair, once per level
// (and in the rare "slow path"): per-warp partials, one shared-memory hop, warp 0 adds the per-warp partials (and, in the
// cluster variant, the per-CTA sums every CTA received through DSMEM, after a cluster barrier).  On return the totals are in
// s.sums[0..23] of EVERY CTA of the pair, visible to warp 0 only (callers that need them elsewhere synchronise).
// (`sum_warp`: the warp that adds the per-warp partials and afterwards sees s.sums -- warp 0, or, one CTA per pair, another
// warp of the caller's choice.)
template <int FPT, int CS, bool XG, bool ROLL, class SH, class Get>
__device__ __forceinline__ void pair_sum_h_to_warp0(Get get, SH& s, int nwarps, const XgParams& xg, int xg_pair, int sum_warp = 0) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  warp_h_partials<FPT, ROLL>(get, &s.hpart[warp * kPartK]);
  __syncthreads();
  if constexpr (CS == 1) {
    if (warp == sum_warp) {
      double acc = 0.0;
      if (lane < kPartK)
        for (int wv = 0; wv < nwarps; ++wv) acc += s.hpart[wv * kPartK + lane];
      if (XG && xg.world > 1) {  // feature split over GPUs: the other ranks' partial sums arrive through peer memory
        int z0 = 0, z1 = 0;
        xg_allreduce<kPartK>(xg, xg_pair, &s.xg_seq, acc, z0, z1);
      }
      if (lane < kPartK) s.sums[lane] = acc;
      __syncwarp();
    }
  } else {
    const unsigned rank = cluster_rank();
    if (warp == 0 && lane < kPartK) {
      double acc = 0.0;
      for (int wv = 0; wv < nwarps; ++wv) acc += s.hpart[wv * kPartK + lane];
#pragma unroll
      for (int r = 0; r < CS; ++r) st_cluster_f64(&s.hsum_cta[rank][lane], (unsigned)r, acc);
    }
    pair_sync<CS>();
    if (warp == 0) {
      if (lane < kPartK) {
        double acc = 0.0;
#pragma unroll
        for (int r = 0; r < CS; ++r) acc += s.hsum_cta[r][lane];
        s.sums[lane] = acc;
      }
      __syncwarp();
    }
  }
}

// Warp 0: scale the 21 summed H entries into the full symmetric 6x6 `Hdst` and factorise it into `S`.
// Every lane runs the register-resident unpivoted LDL^T redundantly (same cost as one lane), lane 0 stores
// the factors; the pivoted Eigen-like fallback handles a degenerate H.
__device__ __forceinline__ void warp_scale_and_factor(const double* sums, double s2, double* Hdst, Solver6& S) {
  const int lane = threadIdx.x & 31;
  double h[21];
#pragma unroll
  for (int k = 0; k < 21; ++k) h[k] = sums[k] * s2;
  if (lane == 0) {
#pragma unroll
    for (int r = 0; r < 6; ++r)
#pragma unroll
      for (int c = r; c < 6; ++c) { Hdst[r * 6 + c] = h[upper_idx(r, c)]; Hdst[c * 6 + r] = h[upper_idx(r, c)]; }
  }
  Fact6 F;
  const bool ok = fact6_compute_upper(h, F);
  if (lane == 0) {
    if (ok) {
      S.F = F;
      S.pivoted = 0;
    } else {
      for (int k = 0; k < 36; ++k) S.ldl[k] = Hdst[k];
      ldlt6_factor(S.ldl, S.tr);
      S.pivoted = 1;
    }
  }
  __syncwarp();
}

// Tail of one NLLSSolver::optimizeGaussNewton iteration [EXT] -- solve, accept/rollback, update -- executed
// redundantly by EVERY thread of the pair on bit-identical inputs (tot[], n_in, the shared factorisation and the
// double-buffered state), so that all threads leave with the same new pose in registers and the same `done`.
// Thread 0 of every CTA (`cta_leader`) writes that CTA's state buffer of the next iteration; thread 0 of CTA rank 0
// (`leader`) also keeps the counters and the trace.
template <class SH>
__device__ __forceinline__ void gn_tail(SH& s, unsigned g, const Solver6* S, const double (&tot)[7], double jscale,
                                        int n_in, int iter, int level, double eps, bool cta_leader, bool leader, svo_b200_sia_iter* trace,
                                        int trace_cap, double& chi2_prev, int& stop, int& done, double (&R)[9],
                                        double (&t)[3]) {
#if SVO_SIA_DEBUG
  long long tg0 = clock64();
#endif
  const int n_meas = n_in * kPatchArea;
  const float chi2f = (float)tot[6];
  const double new_chi2 = (double)(chi2f / (float)n_meas);  // sparse_img_align.cpp:242 (NaN if 0)
  double x[6];
  if (S == nullptr) {  // Eigen's LDLT of an all-zero matrix solves to x = 0
#pragma unroll
    for (int k = 0; k < 6; ++k) x[k] = 0.0;
  } else if (!S->pivoted) {
    const Fact6 F = S->F;
    double b[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) b[k] = -(tot[k] * jscale);  // Jres_ = -sum J r
    fact6_solve(F, b, x);
  } else {
    double b[6];
    for (int k = 0; k < 6; ++k) b[k] = -(tot[k] * jscale);
    ldlt6_solve(S->ldl, S->tr, b);
    for (int k = 0; k < 6; ++k) x[k] = b[k];
  }
#if SVO_SIA_DEBUG
  long long tg1 = clock64() + (long long)(x[0] != x[0]) + (long long)(x[3] != x[3]);
#endif
  const SiaState& cur = s.st[g & 1u];
  const Pose model = cur.model;
  if (isnan(x[0])) stop = 1;  // solve() == 0 (:248-250); stop_ latches
  int accepted;
  Pose out;
  done = 0;
  if ((iter > 0 && new_chi2 > chi2_prev) || stop) {
    out = cur.old_model;  // rollback
    done = 1;
    accepted = 0;
    if (cta_leader) { s.st[(g + 1u) & 1u].model = out; s.st[(g + 1u) & 1u].old_model = out; }
  } else {
    double mx[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) mx[k] = -x[k];
    out = pose_mul_fast(model, se3_exp_fast(mx));  // T_new = T_old * exp(-x)  (:257)
    chi2_prev = new_chi2;
    accepted = 1;
    const double m = fmax(fmax(fmax(fabs(x[0]), fabs(x[1])), fmax(fabs(x[2]), fabs(x[3]))), fmax(fabs(x[4]), fabs(x[5])));
    if (m <= eps) done = 1;
    if (cta_leader) { s.st[(g + 1u) & 1u].model = out; s.st[(g + 1u) & 1u].old_model = model; }
  }
  qmatrix(out.q, R);
  t[0] = out.t[0]; t[1] = out.t[1]; t[2] = out.t[2];
#if SVO_SIA_DEBUG
  long long tg2 = clock64() + (long long)(R[0] != R[0]) + (long long)(t[0] != t[0]);
  if (leader) { s.tk[5] += tg1 - tg0; s.tk[6] += tg2 - tg1; }
#endif
  if (leader) {
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
}

// 4-byte asynchronous global->shared copy (LDGSTS): no register staging, completion per thread
__device__ __forceinline__ void cp_async4(void* dst_smem, const void* src_gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
}
__device__ __forceinline__ void cp_async8(void* dst_smem, const void* src_gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
}
__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src_gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void prefetch_l2_bulk(const void* p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

enum { kModeGlobal = 0, kModeImage = 1, kModeWindow = 2 };

// slot of Bq[r][c] in the 32-value patch cache (BQ): row 0 cols 1..4 -> 0..3, rows 1..4 cols 0..5 -> 4..27, row 5 cols 1..4 -> 28..31
__host__ __device__ constexpr int bq_idx(int r, int c) { return r == 0 ? c - 1 : r == 5 ? 28 + c - 1 : 4 + (r - 1) * 6 + c; }
static_assert(bq_idx(0, 1) == 0 && bq_idx(0, 4) == 3 && bq_idx(1, 0) == 4 && bq_idx(1, 5) == 9 && bq_idx(4, 5) == 27 &&
                  bq_idx(5, 1) == 28 && bq_idx(5, 4) == 31,
              "the 32 used entries of the 6x6 bilinear reference array map onto 0..31 without gaps");

// xyz_cur = T_cur_from_ref * xyz_ref.  One CTA per pair (CS == 1): the pose is read from shared memory (s.pub, published by
// warp 0's Gauss-Newton tail) at the point of use -- six 128-bit shared loads per feature instead of 24 registers that stay
// live across the residual loops (at 128 registers per thread those were spilled to local memory, which misses the small L1
// left beside 3 x 75 KB of shared memory: ncu r02c, 15 M local loads per launch, 17 % L1 hits).  Cluster per pair: every
// warp runs the tail itself and keeps the pose in registers.
template <int CS>
__device__ __forceinline__ void sia_transform(const double* pub, const double (&R)[9], const double (&t)[3], double x, double y,
                                              double z, double& xc, double& yc, double& zc) {
  if constexpr (CS == 1) {
    double r[12];
    const uint32_t a = smem_u32(pub);
#pragma unroll
    for (int k = 0; k < 6; ++k)
      asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(r[2 * k]), "=d"(r[2 * k + 1]) : "r"(a + 16u * k));
    xc = fma(r[0], x, fma(r[1], y, fma(r[2], z, r[9])));
    yc = fma(r[3], x, fma(r[4], y, fma(r[5], z, r[10])));
    zc = fma(r[6], x, fma(r[7], y, fma(r[8], z, r[11])));
  } else {
    xc = fma(R[0], x, fma(R[1], y, fma(R[2], z, t[0])));
    yc = fma(R[3], x, fma(R[4], y, fma(R[5], z, t[1])));
    zc = fma(R[6], x, fma(R[7], y, fma(R[8], z, t[2])));
  }
}

template <bool CG>
__device__ __forceinline__ void sia_world2cam(const CamDev& c, double x, double y, double& u, double& v) {
  if constexpr (CG) {
    cam_world2cam(c, x, y, u, v);
  } else {  // undistorted pinhole (the host dispatches on svo_b200_camera: model == PINHOLE, no distortion)
    u = fma(c.fx, x, c.cx);
    v = fma(c.fy, y, c.cy);
  }
}

// One CTA (CS == 1) or one cluster of CS CTAs per frame pair; the pair's features are dealt to the CTAs in
// contiguous blocks of S = MAXT*FPT slots.
// (__launch_bounds__(160, 3) yields 128 registers although 3 x 160 x 136 <= 64 K: the register file is split over the
// four SM sub-partitions, 15 warps put 4 on one of them, and 4 warps x 32 lanes x 136 > 16 K.  Measured: forcing 136 with
// __maxnreg__ drops the kernel to two CTAs per SM and costs 20 % throughput.  Also measured and not kept: laying the
// two-feature residual pass out phase by phase for instruction-level parallelism (no change), and issuing both features'
// reference-footprint loads before computing either patch (-5 %: the extra live registers spill).)
//
// CG = false compiles the projection for the undistorted pinhole only (px = fx * uv + cx): the general vk::AbstractCamera
// dispatch (radial-tangential pinhole, ATAN with its atan() slow path) stays out of the instruction stream of the
// residual loop, which is what BASELINE's synthetic camera and any rectified stream run.
//
// UP = true (cluster geometry, every CTA alone on its SM): patches / H / factorisations of all levels are computed before
// the first iteration (SiaUpT); shared memory then holds one patch array set per level.
template <int FPT, bool EVAL, int MAXT, int MINB, int CS, bool CG, bool UP>
__global__ void __launch_bounds__(MAXT, MINB) sia_kernel(const SiaParams P) {
  static_assert(!UP || (CS > 1 && FPT == 1 && !EVAL), "the upfront variant exists for the cluster geometry only");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  using SH = SiaSharedT<MAXT / 32, CS>;
  using UPT = SiaUpT<MAXT / 32, CS>;
  SH& s = *reinterpret_cast<SH*>(smem_raw);
  constexpr int S = MAXT * FPT;  // feature slots of this CTA (== P.slots, checked on the host)
  // Throughput geometry (160 threads x 2 features, three CTAs per SM): the shared arrays are allocated for SA = 304 slots
  // (the host selects it for <= 304 features only), which leaves room for the per-feature state xyz_ref in shared memory
  // next to the two coarsest current images.  Kept in registers that state was spilled at 128 registers per thread, and local
  // memory misses the small L1 left beside 3 x 75 KB of shared memory (ncu r02c: 15 M local loads per launch, 17 % L1 hits,
  // each at the head of a feature's projection chain).
  constexpr bool SS = (FPT == 2 && MAXT == 160 && CS == 1);
  constexpr int SA = SS ? 304 : S;  // stride of the per-slot shared arrays
  // The throughput geometry has no room for windows (its staging region holds the two coarsest current images) and is never
  // used for the multi-GPU feature split: both code paths are compiled out of it, and its residual pass loops over the
  // thread's features instead of being unrolled -- the instruction stream of one Gauss-Newton iteration shrinks from ~27 KB to
  // ~20 KB, which matters with three CTAs in different phases sharing one instruction cache (ncu r02h: 16 % of the stall
  // samples are instruction-fetch stalls).
  // BQ (throughput geometry compiled for FOUR CTAs per SM): the patch cache holds the 32 values of the bilinear reference
  // array the patch and both gradients are made of (rows 1..4 x cols 0..5 and cols 1..4 of rows 0 and 5 of the 6x6 array Bq)
  // instead of 16 values + 16 gradient pairs -- 128 instead of 192 bytes per feature; dx = (Bq[y][x+1] - Bq[y][x-1]) / 2 and
  // dy likewise are formed in the residual pass with the same two roundings precomputeReferencePatches uses.  That brings a
  // CTA to ~55 KB of shared memory: four pairs per SM instead of three interleave their serial phases.
  constexpr bool BQ = SS && MINB == 4;
  constexpr int kPatFloats = BQ ? 32 : 3 * kPatchArea;  // floats per feature slot in one patch array set
  constexpr bool WIN = !SS;  // per-feature cp.async windows of the current image exist in this instantiation
  constexpr bool XG = (CS == 1) && !SS;  // multi-GPU feature split (svo_b200_sia_split_*) compiled in
  constexpr size_t kCtlBytes = ((sizeof(SH) + 15) & ~size_t(15)) + (UP ? ((sizeof(UPT) + 15) & ~size_t(15)) : 0);
  UPT& up = *reinterpret_cast<UPT*>(smem_raw + ((sizeof(SH) + 15) & ~size_t(15)));  // only touched when UP
  const int n_lvl_bufs = UP ? (P.max_level - P.min_level + 1) : 1;  // patch array sets (one per level when UP)
  float* const pat_base = reinterpret_cast<float*>(smem_raw + kCtlBytes);
  // set li (0 = coarsest level) : [16][S] f32 reference patch, then [16][S] float2 gradients
  auto pat_ref_of = [&](int li) -> float* { return pat_base + (size_t)li * kPatFloats * SA; };
  auto pat_dxy_of = [&](int li) -> float2* { return reinterpret_cast<float2*>(pat_ref_of(li) + kPatchArea * SA); };
  float* pat_ref = pat_ref_of(0);
  float2* pat_dxy = pat_dxy_of(0);
  double* const st_xyz = reinterpret_cast<double*>(pat_base + (size_t)n_lvl_bufs * kPatFloats * SA);  // SS: [3][SA] xyz_ref
  uint8_t* stage = reinterpret_cast<uint8_t*>(st_xyz + (SS ? 3 * SA : 0));  // 16-byte aligned
  uint4* win = reinterpret_cast<uint4*>(stage);                                                      // [kWinRows][S] 16-byte window rows

  const unsigned crank = CS == 1 ? 0u : cluster_rank();
  const int pair = CS == 1 ? (int)blockIdx.x : (int)(blockIdx.x / CS);
  const SiaJob& job = P.jobs[pair];
  const int tid = threadIdx.x, T = blockDim.x, nwarps = (T + 31) >> 5;
  const int lane = tid & 31, warp = tid >> 5;
  const int fbase = (int)crank * S;  // first feature of this CTA
  const int N = job.n_feat;          // features of the pair
  const int np = job.n_pad;
  const bool cta_leader = tid == 0, leader = tid == 0 && crank == 0;
  // features of this CTA: [fbase, min(N, fbase + S)); their records are np_loc padded entries of the pair's blob
  const int n_loc = max(0, min(N - fbase, S));
  const int np_loc = (n_loc + 15) & ~15;

  if (tid == 0) {
    mbar_init(&s.mbar, 1);
    mbar_init(&s.xbar[0], 1);
    mbar_init(&s.xbar[1], 1);
    fence_mbar_init();
    // ---- TMA: the packed feature records of this CTA's features, four bulk copies (px, f, pos, has_point
    //      sections of the pair's blob) into the (idle) patch arrays -- issued first, everything else this thread
    //      initialises runs in the shadow of that copy
    fence_proxy_async();
    if (np_loc > 0) {
      const uint32_t bytes = (uint32_t)np_loc * 65u;
      mbar_expect_tx(&s.mbar, bytes);
      uint8_t* dst = reinterpret_cast<uint8_t*>(pat_ref);
      const uint8_t* src = job.blob;
      tma_bulk_g2s(dst, src + (size_t)fbase * 16, (uint32_t)np_loc * 16u, &s.mbar);                                   // px
      tma_bulk_g2s(dst + (size_t)np_loc * 16, src + (size_t)np * 16 + (size_t)fbase * 24, (uint32_t)np_loc * 24u, &s.mbar);  // f
      tma_bulk_g2s(dst + (size_t)np_loc * 40, src + (size_t)np * 40 + (size_t)fbase * 24, (uint32_t)np_loc * 24u, &s.mbar);  // pos
      tma_bulk_g2s(dst + (size_t)np_loc * 64, src + (size_t)np * 64 + (size_t)fbase, (uint32_t)np_loc, &s.mbar);             // has_point
    } else {
      mbar_arrive(&s.mbar);  // a CTA without features still completes use 0 of the barrier: the phase parities of the image copies stay in step
    }
    s.st[0].model = pose_from_rt12(job.T);
    s.st[0].old_model = s.st[0].model;
    s.h_is_tot = 0; s.n_in_last = 0;
    s.n_iters = 0; s.sum_vis = 0; s.sum_in = 0; s.n_trace = 0;
    s.xg_seq = (CS == 1 && P.xg.world > 1) ? P.xg.peer[P.xg.rank][pair].xseq : 0u;
    s.xg_failed = 0u;
#if SVO_SIA_DEBUG
    for (int k = 0; k < 8; ++k) s.tk[k] = 0;
    for (int k = 0; k < 4; ++k) s.tkx[k] = 0;
    s.tk[4] = clock64();
#endif
    for (int k = 0; k < 36; ++k) s.Hs[k] = 0.0;
  }
  __syncthreads();
  mbar_wait(&s.mbar, 0);
  const uint8_t* blob = reinterpret_cast<const uint8_t*>(pat_ref);
  const double* b_f = reinterpret_cast<const double*>(blob) + 2 * np_loc;
  const double* b_pos = b_f + 3 * np_loc;
  const uint8_t* b_hp = reinterpret_cast<const uint8_t*>(b_pos + 3 * np_loc);

  // per-feature register state
  double fx_[FPT], fy_[FPT], fz_[FPT], fzi_[FPT];
  double fxs[FPT], fys[FPT], fzs[FPT];  // SS: copies that go to shared memory once the staged blob has been consumed
  int wx_[FPT], wy_[FPT];  // origin of the feature's current-image window (kModeWindow)
  unsigned hp_mask = 0, vis_mask = 0, in_mask = 0, stale_mask = 0;
#pragma unroll
  for (int k = 0; k < FPT; ++k) {
    const int i = tid + k * T;
    fx_[k] = fy_[k] = 0.0; fz_[k] = fzi_[k] = 1.0;
    wx_[k] = wy_[k] = -(1 << 20);
    if (i < n_loc) {
      const double dxp = b_pos[3 * i] - job.ref_pos[0], dyp = b_pos[3 * i + 1] - job.ref_pos[1],
                   dzp = b_pos[3 * i + 2] - job.ref_pos[2];
      const double depth = sqrt(dxp * dxp + dyp * dyp + dzp * dzp);  // :107  |pos - ref_pos|
      fx_[k] = b_f[3 * i] * depth;                                    // :108  xyz_ref = f * depth
      fy_[k] = b_f[3 * i + 1] * depth;
      fz_[k] = b_f[3 * i + 2] * depth;
      fzi_[k] = 1.0 / fz_[k];
      if constexpr (SS) { fxs[k] = fx_[k]; fys[k] = fy_[k]; fzs[k] = fz_[k]; }
      if (b_hp[i]) hp_mask |= 1u << k;
      if (EVAL && P.visible_in[fbase + i]) vis_mask |= 1u << k;
    }
  }
  // solver state every thread carries (uniform over the pair): chi2_ and stop_ of NLLSSolver (reset(): 1e10 / false
  // [EXT]), the running iteration counter that selects the state / partial-sum buffers, and the current pose
  double chi2_prev = 1e10;
  int stop = 0;
  unsigned g = 0;
  double R[9], t[3];
  {
    const Pose m0 = pose_from_rt12(job.T);  // same arithmetic as thread 0's s.st[0].model
    qmatrix(m0.q, R);
    t[0] = m0.t[0]; t[1] = m0.t[1]; t[2] = m0.t[2];
    if (CS == 1 && tid == 0) {
#pragma unroll
      for (int k = 0; k < 9; ++k) s.pub[k] = R[k];
      s.pub[9] = t[0]; s.pub[10] = t[1]; s.pub[11] = t[2];
      s.pub_done = 0; s.pub_slow = 0;
    }
  }
  if constexpr (UP) {
    // Latency of a live pair is a chain of first-touch DRAM reads (reference footprints and current-image windows of every
    // level, ~0.7 us each): all of them are known now -- the footprints exactly, the windows at the initial pose, which the
    // iterations move by a few pixels at most -- so every thread pulls its feature's rows of every level into L2 at once.
    // (Only here: with the CTA alone on its SM the extra L1 requests cost nothing; in the batched geometries they do.)
    const int i = tid;
    if (i < n_loc) {
      const double2 pxy = *reinterpret_cast<const double2*>(blob + (size_t)i * 16);
      double xc, yc, zc;
      sia_transform<CS>(s.pub, R, t, fx_[0], fy_[0], fz_[0], xc, yc, zc);
      const double rz = fast_rcp(zc);
      double ud, vd;
      sia_world2cam<CG>(P.cam, div_rn(xc, zc, rz), div_rn(yc, zc, rz), ud, vd);
      for (int level = P.max_level; level >= P.min_level; --level) {
        const int W = P.w[level], Hh = P.h[level];
        const double sc = 1.0 / (double)(1 << level);
        const int ur = (int)(pxy.x * sc), vr = (int)(pxy.y * sc);
        if (ur - 3 >= 0 && vr - 3 >= 0 && ur + 4 < W && vr + 4 < Hh) {
          const uint8_t* p0 = job.ref_lvl[level] + (size_t)(vr - 3) * W + (ur - 3);
#pragma unroll
          for (int r = 0; r < 7; ++r) prefetch_l2(p0 + (size_t)r * W);
        }
        const int uc = (int)(ud * sc), vc = (int)(vd * sc);
        if ((uint32_t)(W * Hh) + 32u > (uint32_t)P.stage_cap && uc - 4 >= 0 && vc - 4 >= 0 && uc + 4 < W && vc + 4 < Hh && ud >= 0.0 && vd >= 0.0) {
          const uint8_t* p0 = job.cur_lvl[level] + (size_t)(vc - 4) * W + (uc - 4);
#pragma unroll
          for (int r = 0; r < kWinRows; ++r) prefetch_l2(p0 + (size_t)r * W);
        }
      }
    }
    if (tid == 0 && crank == 0) {  // the coarse current images that are staged whole
      for (int level = P.max_level; level >= P.min_level; --level) {
        const uint32_t nb = ((uint32_t)(P.w[level] * P.h[level]) + 15u) & ~15u;
        if (nb + 16u <= (uint32_t)P.stage_cap) prefetch_l2_bulk(job.cur_lvl[level], nb);
      }
    }
  }
  pair_sync<CS>();  // everyone is done with the staged blob (the patch arrays may be written) and, in the cluster
                    // variant, every CTA's shared memory is initialised before remote stores arrive
    SIA_DBG(if ((SVO_SIA_DEBUG && P.debug) && tid == 0) s.tkx[3] = clock64() - s.tk[4];)
  if constexpr (SS) {
#pragma unroll
    for (int k = 0; k < FPT; ++k) {
      const int slot = tid + k * T;
      if (slot < n_loc) { st_xyz[slot] = fxs[k]; st_xyz[SA + slot] = fys[k]; st_xyz[2 * SA + slot] = fzs[k]; }
    }
  }
  // xyz_ref of feature k of this thread, and 1/z (SS: re-read from shared memory / recomputed, correctly rounded like the division)
  auto feat_xyz = [&](const int k, const int slot, double& x, double& y, double& z) {
    if constexpr (SS) {
      // slots without a feature (and the slots >= SA the last half-warp maps to) yield the neutral (0, 0, 1) the register
      // variant initialises: the H reduction multiplies their Jacobian rows by zero moments, which must stay finite
      x = 0.0; y = 0.0; z = 1.0;
      if (slot < n_loc) { x = st_xyz[slot]; y = st_xyz[SA + slot]; z = st_xyz[2 * SA + slot]; }
    } else { x = fx_[k]; y = fy_[k]; z = fz_[k]; }
  };
  auto feat_zi = [&](const int k, const double z) -> double {
    if constexpr (SS) return rcp_rn(z);
    else return fzi_[k];
  };

  // ---- one feature's window of the current image at this level (kModeWindow): 16 columns x 8 rows around the projection
  //      with the pose the level starts from, requested with cp.async (completion: cp_async_wait_all by the same thread)
  auto stage_window = [&](const int k, const int slot, const int W, const int Hh, const float scale, const uint8_t* cur_img) {
    double x, y, z;
    feat_xyz(k, slot, x, y, z);
    double xc, yc, zc;
    sia_transform<CS>(s.pub, R, t, x, y, z, xc, yc, zc);
    const double rz = fast_rcp(zc);
    double ud, vd;
    sia_world2cam<CG>(P.cam, div_rn(xc, zc, rz), div_rn(yc, zc, rz), ud, vd);
    const float u0 = __fmul_rn((float)ud, scale), v0 = __fmul_rn((float)vd, scale);
    if (u0 >= 0.f && v0 >= 0.f && u0 < 1e6f && v0 < 1e6f) {
      float tmp;
      const int cu = floor_pos(__fadd_rn(u0, 0.5f), tmp), cv = floor_pos(__fadd_rn(v0, 0.5f), tmp);
      {
        // 16 columns x 8 rows around (round(u), round(v)): the 5x5 footprint stays inside for at least +-1.5 px of
        // drift.  One 16-byte cp.async per row when columns round(u)-4 .. round(u)+3 fall into one 16-byte aligned
        // block (and the pitch keeps every row 16-byte aligned), else two 8-byte copies from the 8-byte aligned column.
        const int c4 = cu - 4, wy = cv - 4;
        const bool one = ((c4 & 15) <= 8) && (W & 15) == 0;
        const int wx = one ? (c4 & ~15) : (c4 & ~7);
        if (wx >= 0 && wy >= 0 && wx + 16 <= W && wy + kWinRows <= Hh) {
          wx_[k] = wx; wy_[k] = wy;
          const uint8_t* src = cur_img + (size_t)wy * W + wx;
          if (one) {
#pragma unroll
            for (int r = 0; r < kWinRows; ++r) cp_async16(win + r * SA + slot, src + (size_t)r * W);
          } else {
#pragma unroll
            for (int r = 0; r < kWinRows; ++r) {
              cp_async8(reinterpret_cast<uint8_t*>(win + r * SA + slot), src + (size_t)r * W);
              cp_async8(reinterpret_cast<uint8_t*>(win + r * SA + slot) + 8, src + (size_t)r * W + 8);
            }
          }
        }
      }
    }
  };
  // ---- precomputeReferencePatches (:84-145) of one level into the patch arrays (pr, pd): visibility bits, the f32 patch and
  //      its gradients, and the per-feature gradient moments m_* the H reduction needs.  `with_windows`: also request the
  //      current-image windows (between the footprint loads and the arithmetic, so that both latencies overlap).
  auto level_patches = [&](const int level, float* pr, float2* pd, const float* pr_stale, const int mode, const bool with_windows,
                           double (&m_sxx)[FPT], double (&m_sxy)[FPT], double (&m_syy)[FPT], double (&m_cnt)[FPT]) {
    const int W = P.w[level], Hh = P.h[level];
    const float scale = 1.0f / (float)(1 << level);
    const uint8_t* ref_img = job.ref_lvl[level];
    const uint8_t* cur_img = job.cur_lvl[level];
    // (throughput geometry: a loop over the thread's features -- the per-feature moments m_* are then indexed dynamically and
    // live in local memory, eight values per level, which is cheaper than a second copy of the patch arithmetic in the
    // instruction stream)
#pragma unroll(SS ? 1 : FPT)
    for (int k = 0; k < FPT; ++k) {
      m_sxx[k] = m_sxy[k] = m_syy[k] = m_cnt[k] = 0.0;
      stale_mask &= ~(1u << k);
      const int slot = tid + k * T;
      // px is re-read from the pair's blob in global memory (L2) once per level instead of living in registers
      const double2 pxy = slot < n_loc ? __ldg(reinterpret_cast<const double2*>(job.blob) + fbase + slot) : make_double2(-1e6, -1e6);
      const float u_ref = (float)(pxy.x * (double)scale);
      const float v_ref = (float)(pxy.y * (double)scale);
      const bool rng = u_ref >= 0.f && v_ref >= 0.f && u_ref < 1e6f && v_ref < 1e6f;  // else: outside, floor not needed
      float ufl = 0.f, vfl = 0.f;
      const int ui = rng ? floor_pos(u_ref, ufl) : -1, vi = rng ? floor_pos(v_ref, vfl) : -1;
      const bool ok = ((hp_mask >> k) & 1u) && ui - 3 >= 0 && vi - 3 >= 0 && ui + 3 < W && vi + 3 < Hh;
      // all seven footprint rows are requested before anything else: one exposed L2/HBM latency per level
      uint32_t rlo[7], rhi[7];
      if (ok) {
        vis_mask |= 1u << k;
#pragma unroll
        for (int r = 0; r < 7; ++r) fetch7_g64(ref_img, (vi - 3 + r) * W + (ui - 3), rlo[r], rhi[r]);
      }
      // ---- current-image window of this feature (fine levels): projected with the pose the level starts from
      if constexpr (WIN) {
        if (with_windows) {
          wx_[k] = wy_[k] = -(1 << 20);
          if (((vis_mask >> k) & 1u) && mode == kModeWindow) stage_window(k, slot, W, Hh, scale, cur_img);
        }
      }
      if (ok) {
        float wtl, wtr, wbl, wbr;
        bilin_weights(__fsub_rn(u_ref, ufl), __fsub_rn(v_ref, vfl), wtl, wtr, wbl, wbr);
        // 7x7 footprint rows vi-3..vi+3, cols ui-3..ui+3, streamed row by row to keep the register
        // footprint small: Bq row r (bilinear blends with top-left tap P[r][c]) needs footprint rows
        // r, r+1; patch row y needs Bq rows y, y+1, y+2.
        float pr0[7], pr1[7], b0[6], b1[6], b2[6];
        double sxx = 0, sxy = 0, syy = 0;
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
          if constexpr (BQ) {  // b2 is row r of Bq: cache it (corner columns of rows 0 and 5 are never used)
#pragma unroll
            for (int c = (r == 0 || r == 5) ? 1 : 0; c < ((r == 0 || r == 5) ? 5 : 6); ++c) pr[bq_idx(r, c) * SA + slot] = b2[c];
          }
          if (r >= 2) {  // rows b0 (= Bq[y]), b1 (= Bq[y+1]), b2 (= Bq[y+2]) with y = r-2 are complete
            const int y = r - 2;
#pragma unroll
            for (int x = 0; x < 4; ++x) {
              const int p = y * 4 + x;
              const float val = b1[x + 1];
              const float dx = __fmul_rn(0.5f, __fsub_rn(b1[x + 2], b1[x]));
              const float dy = __fmul_rn(0.5f, __fsub_rn(b2[x + 1], b0[x + 1]));
              if constexpr (!BQ) {
                pr[p * SA + slot] = val;
                pd[p * SA + slot] = make_float2(dx, dy);
              }
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
        if constexpr (BQ) {
          stale_mask |= 1u << k;  // the cache keeps the previous level's Bq: stale values, gradients scaled by zero
        } else {
#pragma unroll
          for (int p = 0; p < kPatchArea; ++p) pd[p * SA + slot] = make_float2(0.f, 0.f);
          if (pr_stale)  // per-level arrays (upfront variant): the stale patch is the previous level's
            for (int p = 0; p < kPatchArea; ++p) pr[p * SA + slot] = pr_stale[p * SA + slot];
        }
        m_cnt[k] = 1.0;
      }
    }
  };

  const int lvl_hi = EVAL ? P.eval_level : P.max_level;
  const int lvl_lo = EVAL ? P.eval_level : P.min_level;
  unsigned vis_levels = 0;  // UP: bit li = this thread's feature is visible at level lvl_hi - li (set-only across levels)
  unsigned img_phase = 0;   // parity of the most recent use of s.mbar (the blob copy of the prologue was use 0)
  if constexpr (UP) {
    SIA_DBG(long long tu0 = 0;)
      SIA_DBG(if ((SVO_SIA_DEBUG && P.debug) && tid == 0) tu0 = clock64();)
    // ---- all levels' reference patches and per-warp H partials, coarse to fine (the visibility mask accumulates in that order)
    for (int level = lvl_hi; level >= lvl_lo; --level) {
      const int li = lvl_hi - level;
      double m_sxx[FPT], m_sxy[FPT], m_syy[FPT], m_cnt[FPT];
      level_patches(level, pat_ref_of(li), pat_dxy_of(li), li > 0 ? pat_ref_of(li - 1) : (const float*)nullptr, kModeGlobal, false,
                    m_sxx, m_sxy, m_syy, m_cnt);
      vis_levels |= (vis_mask & 1u) << li;
      warp_h_partials<FPT, false>(
          [&](int k, double& x, double& y, double& zi, double& sxx, double& sxy, double& syy, double& cnt) {
            { double z_; feat_xyz(k, (int)threadIdx.x + k * (int)blockDim.x, x, y, z_); zi = feat_zi(k, z_); } sxx = m_sxx[k]; sxy = m_sxy[k]; syy = m_syy[k]; cnt = m_cnt[k];
          },
          &up.hpart[li][warp * kPartK]);
    }
    __syncthreads();
    // ---- one exchange for all levels: every CTA receives every CTA's per-level sums (DSMEM), one cluster barrier
    const int nlv = lvl_hi - lvl_lo + 1;
    for (int e = tid; e < nlv * kPartK; e += T) {
      const int li = e / kPartK, j = e - li * kPartK;
      double acc = 0.0;
      for (int wv = 0; wv < nwarps; ++wv) acc += up.hpart[li][wv * kPartK + j];
#pragma unroll
      for (int r = 0; r < CS; ++r) st_cluster_f64(&up.hsum_cta[li][crank][j], (unsigned)r, acc);
    }
    pair_sync<CS>();
    // ---- every CTA scales and factorises its own copy, the levels dealt to its warps
    for (int li = warp; li < nlv; li += nwarps) {
      if (lane < kPartK) {
        double acc = 0.0;
#pragma unroll
        for (int r = 0; r < CS; ++r) acc += up.hsum_cta[li][r][lane];
        up.sums[li][lane] = acc;
      }
      __syncwarp();
      const double js = P.cam.fx / (double)(1 << (lvl_hi - li));
      warp_scale_and_factor(up.sums[li], js * js, up.Htot[li], up.sol[li]);
    }
    __syncthreads();
      SIA_DBG(if ((SVO_SIA_DEBUG && P.debug) && tid == 0) s.tkx[2] += clock64() - tu0;)
  }
  for (int level = lvl_hi; level >= lvl_lo; --level) {
    const int W = P.w[level], Hh = P.h[level];
    const float scale = 1.0f / (float)(1 << level);
    const double jscale = P.cam.fx / (double)(1 << level);  // focal_length / (1<<level_)  (:140)
    const uint8_t* cur_img = job.cur_lvl[level];

    // ---- how the current image of this level reaches the residual loop --------------------------
    const uint32_t img_bytes = ((uint32_t)(W * Hh) + 15u) & ~15u;
    int mode = kModeGlobal;
    if (img_bytes + 16u <= (uint32_t)P.stage_cap) mode = kModeImage;
    else if (WIN && P.use_windows && (W & 7) == 0 && kWinBytes * SA <= P.stage_cap) mode = kModeWindow;
    // Phase parity of this use of s.mbar, kept by every thread in a register (use k completes parity k & 1; the blob copy was
    // use 0).  NOT read from shared memory: in the upfront variant no barrier separates thread 0's update from the other
    // warps' wait, and a stale parity lets them through before the image has landed (found by running the tests under
    // compute-sanitizer, whose timing exposed it; `mode` is the same in all threads).
    if (mode == kModeImage) img_phase ^= 1u;
    if (mode == kModeImage && tid == 0) {
      fence_proxy_async();
      mbar_expect_tx(&s.mbar, img_bytes);
      tma_bulk_g2s(stage, cur_img, img_bytes, &s.mbar);
    }
    if (cta_leader) s.st[g & 1u].old_model = s.st[g & 1u].model;  // optimizeGaussNewton: ModelType old_model(model) [EXT]
    // next level: pull its current image (coarse levels, staged whole) into L2 while this level iterates
    if (P.use_prefetch && tid == 0 && crank == 0 && level > lvl_lo) {
      const uint32_t nb = ((uint32_t)(P.w[level - 1] * P.h[level - 1]) + 15u) & ~15u;
      if (nb + 16u <= (uint32_t)P.stage_cap) prefetch_l2_bulk(job.cur_lvl[level - 1], nb);
    }

    SIA_DBG(long long tq0 = 0;)
      SIA_DBG(if ((SVO_SIA_DEBUG && P.debug) && tid == 0) tq0 = clock64();)
    SIA_DBG(long long tq1 = 0;)
    SIA_DBG(long long tq2 = 0;)
    const Solver6* sol_level = &s.sol_tot;  // factorisation of this level's H over its visible set
    const int fwarp = SS ? nwarps - 1 : 0;  // the warp that sums and factorises this level's H
    if constexpr (!UP) {
      // ---- precomputeReferencePatches (:84-145), one feature per thread; the windows of the current image are requested
      //      between the footprint loads and the patch arithmetic
      double m_sxx[FPT], m_sxy[FPT], m_syy[FPT], m_cnt[FPT];
      level_patches(level, pat_ref, pat_dxy, (const float*)nullptr, mode, true, m_sxx, m_sxy, m_syy, m_cnt);
      SIA_DBG(if ((SVO_SIA_DEBUG && P.debug) && tid == 0) { tq1 = clock64(); s.tk[7] += tq1 - tq0; })
      pair_sum_h_to_warp0<FPT, CS, XG, SS, SH>(
          [&](int k, double& x, double& y, double& zi, double& sxx, double& sxy, double& syy, double& cnt) {
            { double z_; feat_xyz(k, (int)threadIdx.x + k * (int)blockDim.x, x, y, z_); zi = feat_zi(k, z_); } sxx = m_sxx[k]; sxy = m_sxy[k]; syy = m_syy[k]; cnt = m_cnt[k];
            // opaque to the optimiser: otherwise the level-invariant Jacobian rows are hoisted out of the level loop and
            // parked in local memory (17 doubles per thread, written once and re-read every level)
            asm volatile("" : "+d"(x), "+d"(y), "+d"(zi));
          },
          s, nwarps, P.xg, pair, fwarp);
      // The scaling and LDL^T factorisation of this level's H is serial work nobody needs before the first solve: one warp
      // does it while the others already run the first residual pass; its results (s.sol_tot, s.Htot) become visible to
      // everybody through the barrier of that pass.  Throughput geometry: the LAST warp, which owns fewer features than the
      // others (44 of 300 against 64) and would reach that barrier early -- not warp 0, which runs the Gauss-Newton tail.
      SIA_DBG(if ((SVO_SIA_DEBUG && P.debug) && tid == 0) { tq2 = clock64(); s.tkx[0] += tq2 - tq1; })
      if (warp == fwarp) {
        if (lane == 0) s.sum_vis += (int)s.sums[21];
        warp_scale_and_factor(s.sums, jscale * jscale, s.Htot, s.sol_tot);
      }
    } else {
      // everything pose independent was prepared before the first level: select this level's arrays, request the windows
      const int li = lvl_hi - level;
      pat_ref = pat_ref_of(li);
      pat_dxy = pat_dxy_of(li);
      sol_level = &up.sol[li];
      vis_mask = (vis_levels >> li) & 1u;
#pragma unroll
      for (int k = 0; k < FPT; ++k) {
        wx_[k] = wy_[k] = -(1 << 20);
        if (((vis_mask >> k) & 1u) && mode == kModeWindow) stage_window(k, tid + k * T, W, Hh, scale, cur_img);
      }
      if (leader) s.sum_vis += (int)up.sums[li][21];
      SIA_DBG(if ((SVO_SIA_DEBUG && P.debug) && tid == 0) { tq1 = tq2 = clock64(); s.tk[7] += tq1 - tq0; })
    }
    if (mode == kModeImage) mbar_wait(&s.mbar, img_phase);
    if (mode == kModeWindow) cp_async_wait_all();  // each thread reads only the window it copied itself
      SIA_DBG(if ((SVO_SIA_DEBUG && P.debug) && tid == 0) { s.tk[0] += clock64() - tq0; s.tkx[1] += clock64() - tq2; })

    // ---- Gauss-Newton iterations at this level ---------------------------------------------
    const int n_iter = EVAL ? 1 : P.n_iter;
    for (int iter = 0; iter < n_iter; ++iter) {
      SIA_DBG(long long ti0 = 0;)
      SIA_DBG(if ((SVO_SIA_DEBUG && P.debug) && tid == 0) ti0 = clock64();)
      double acc[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] = 0.0;
      int n_in_t = 0, n_out_t = 0;
      in_mask = 0;
#pragma unroll(SS ? 1 : FPT)
      for (int k = 0; k < FPT; ++k) {
        if (!((vis_mask >> k) & 1u)) continue;
        const int slot = tid + k * T;
        double x, y, z;
        feat_xyz(k, slot, x, y, z);
        double xc, yc, zc;
        sia_transform<CS>(s.pub, R, t, x, y, z, xc, yc, zc);
        const double rz = fast_rcp(zc);
        double ud, vd;
        sia_world2cam<CG>(P.cam, div_rn(xc, zc, rz), div_rn(yc, zc, rz), ud, vd);  // [EXT] world2cam(project2d(xyz))
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
              P.residuals_out[(size_t)(fbase + slot) * kPatchArea + p] = __int_as_float(0x7fc00000);
          }
          continue;
        }
        in_mask |= 1u << k;
        ++n_in_t;
        float wtl, wtr, wbl, wbr;
        bilin_weights(__fsub_rn(u_cur, ufl), __fsub_rn(v_cur, vfl), wtl, wtr, wbl, wbr);
        uint32_t lo[5], hi[5];
        if (mode == kModeImage) {
#pragma unroll
          for (int r = 0; r < 5; ++r) fetch8<true>(stage, (vi - 2 + r) * W + (ui - 2), lo[r], hi[r]);
        } else {
          const int c0 = WIN ? (ui - 2) - wx_[SS ? 0 : k] : -1, r0 = WIN ? (vi - 2) - wy_[SS ? 0 : k] : -1;
          if (WIN && mode == kModeWindow && (unsigned)c0 <= 11u && (unsigned)r0 <= (unsigned)(kWinRows - 5)) {
            const uint4* wp = win + r0 * SA + slot;
            const int kw = c0 >> 2;  // 0..2: the footprint row starts in word kw of the 16-byte window row
            const unsigned sh = (unsigned)(c0 & 3) * 8u;
#pragma unroll
            for (int r = 0; r < 5; ++r) {
              const uint4 q = wp[r * SA];
              const uint32_t w0 = kw == 0 ? q.x : kw == 1 ? q.y : q.z, w1 = kw == 0 ? q.y : kw == 1 ? q.z : q.w;
              lo[r] = __funnelshift_r(w0, w1, sh);
              hi[r] = w1 >> sh;
            }
          } else {
#pragma unroll
            for (int r = 0; r < 5; ++r) fetch8<false>(cur_img, (vi - 2 + r) * W + (ui - 2), lo[r], hi[r]);
          }
        }
        float c2 = 0.f, gx = 0.f, gy = 0.f;
        float q0[5], q1[5];
        q0[0] = byte_to_float<0>(lo[0]); q0[1] = byte_to_float<1>(lo[0]); q0[2] = byte_to_float<2>(lo[0]);
        q0[3] = byte_to_float<3>(lo[0]); q0[4] = byte_to_float<0>(hi[0]);
        // BQ: three rows of the cached bilinear reference array rotate through registers (row yy cols 1..4, rows yy+1 and yy+2)
        const float ghalf = ((stale_mask >> k) & 1u) ? 0.f : 0.5f;
        float bprev[4], bmid[6], bnext[6];
        if constexpr (BQ) {
#pragma unroll
          for (int c = 0; c < 4; ++c) bprev[c] = pat_ref[bq_idx(0, c + 1) * SA + slot];
#pragma unroll
          for (int c = 0; c < 6; ++c) bmid[c] = pat_ref[bq_idx(1, c) * SA + slot];
        }
#pragma unroll
        for (int yy = 0; yy < 4; ++yy) {
          q1[0] = byte_to_float<0>(lo[yy + 1]); q1[1] = byte_to_float<1>(lo[yy + 1]); q1[2] = byte_to_float<2>(lo[yy + 1]);
          q1[3] = byte_to_float<3>(lo[yy + 1]); q1[4] = byte_to_float<0>(hi[yy + 1]);
          if constexpr (BQ) {
#pragma unroll
            for (int c = (yy == 3) ? 1 : 0; c < ((yy == 3) ? 5 : 6); ++c) bnext[c] = pat_ref[bq_idx(yy + 2, c) * SA + slot];
          }
#pragma unroll
          for (int xx = 0; xx < 4; ++xx) {
            const int p = yy * 4 + xx;
            const float I = bilin(wtl, wtr, wbl, wbr, q0[xx], q0[xx + 1], q1[xx], q1[xx + 1]);
            float val, dx, dy;
            if constexpr (BQ) {
              val = bmid[xx + 1];
              dx = __fmul_rn(ghalf, __fsub_rn(bmid[xx + 2], bmid[xx]));          // as precomputeReferencePatches forms them (:121-126)
              dy = __fmul_rn(ghalf, __fsub_rn(bnext[xx + 1], bprev[xx]));
            } else {
              val = pat_ref[p * SA + slot];
              const float2 gr = pat_dxy[p * SA + slot];
              dx = gr.x; dy = gr.y;
            }
            const float res = __fsub_rn(I, val);
            c2 = fmaf(res, res, c2);  // chi2 += res*res*weight, weight == 1 (:222); order differs from the serial sum anyway
            gx = fmaf(dx, res, gx);
            gy = fmaf(dy, res, gy);
            if (EVAL) P.residuals_out[(size_t)(fbase + slot) * kPatchArea + p] = res;
          }
#pragma unroll
          for (int c = 0; c < 5; ++c) q0[c] = q1[c];
          if constexpr (BQ) {
#pragma unroll
            for (int c = 0; c < 4; ++c) bprev[c] = bmid[c + 1];
#pragma unroll
            for (int c = 0; c < 6; ++c) bmid[c] = bnext[c];
          }
        }
        const double zi = feat_zi(k, z), X = x * zi, Y = y * zi, dgx = (double)gx, dgy = (double)gy;
        acc[0] = fma(-zi, dgx, acc[0]);
        acc[1] = fma(-zi, dgy, acc[1]);
        acc[2] = fma(zi, fma(X, dgx, Y * dgy), acc[2]);
        acc[3] = fma(X * Y, dgx, fma(fma(Y, Y, 1.0), dgy, acc[3]));
        acc[4] = fma(-fma(X, X, 1.0), dgx, fma(-(X * Y), dgy, acc[4]));
        acc[5] = fma(Y, dgx, fma(-X, dgy, acc[5]));
        acc[6] += (double)c2;
      }
      SIA_DBG(long long ti1 = 0;)
      SIA_DBG(if ((SVO_SIA_DEBUG && P.debug) && tid == 0) ti1 = clock64();)

      // ---- pair-wide sums: per-warp transposed reduction, one shared-memory hop, ONE barrier; then every warp adds
      //      the per-warp partials in the same order (bit-identical totals everywhere)
      const unsigned buf = g & 1u;
      warp_reduce_t<8>(acc);
      const int w_in = __reduce_add_sync(0xffffffffu, n_in_t), w_out = __reduce_add_sync(0xffffffffu, n_out_t);
      const int gw = (int)crank * nwarps + warp;  // warp index within the pair
      if constexpr (CS == 1) {
        if ((lane & 3) == 0) s.part[buf][gw][lane >> 2] = acc[0];
        if (lane == 0) { s.cnt[buf][gw][0] = w_in; s.cnt[buf][gw][1] = w_out; }
      } else if (UP && P.async_xchg) {
        // every CTA of the cluster receives every warp's partials by st.async: each store completes its 8 bytes on the
        // receiver's mbarrier of this parity, which one local thread arms with the byte count of all warps of the pair; the
        // receiver wakes when the last byte has landed -- one DSMEM hop, no cluster barrier.  (Two barriers: traffic of
        // iteration g+1 goes to the other one, and nobody can send g+2 before everybody has consumed g.)
        uint64_t* xb = &s.xbar[buf];
        if (tid == 0) mbar_expect_tx(xb, 72u * (uint32_t)(nwarps * CS));
        if ((lane & 3) == 0) {
#pragma unroll
          for (int r = 0; r < CS; ++r) st_async_f64(cluster_addr(&s.part[buf][gw][lane >> 2], (unsigned)r), acc[0], cluster_addr(xb, (unsigned)r));
        }
        if (lane == 0) {
#pragma unroll
          for (int r = 0; r < CS; ++r) st_async_v2s32(cluster_addr(&s.cnt[buf][gw][0], (unsigned)r), w_in, w_out, cluster_addr(xb, (unsigned)r));
        }
        mbar_wait(xb, (g >> 1) & 1u);
      } else {
        // every CTA of the cluster receives every warp's partials (distributed shared memory stores)
        if ((lane & 3) == 0) {
#pragma unroll
          for (int r = 0; r < CS; ++r) st_cluster_f64(&s.part[buf][gw][lane >> 2], (unsigned)r, acc[0]);
        }
        if (lane == 0) {
#pragma unroll
          for (int r = 0; r < CS; ++r) st_cluster_v2s32(&s.cnt[buf][gw][0], (unsigned)r, w_in, w_out);
        }
      }
      if (!(CS > 1 && UP && P.async_xchg)) pair_sync<CS>();  // barrier A: every warp's partial sums are in place
      const int nw_pair = nwarps * CS;
      double tot[7];
      int n_in = 0, n_out = 0, done = 0;
      auto compute_totals = [&]() {
        const int kk = lane & 7;
        double a = 0.0;
        for (int wv = lane >> 3; wv < nw_pair; wv += 4) a += s.part[buf][wv][kk];
        a += __shfl_xor_sync(0xffffffffu, a, 8);
        a += __shfl_xor_sync(0xffffffffu, a, 16);
#pragma unroll
        for (int e = 0; e < 7; ++e) tot[e] = __shfl_sync(0xffffffffu, a, e);
        int ci = 0, co = 0;
        for (int wv = lane; wv < nw_pair; wv += 32) { ci += s.cnt[buf][wv][0]; co += s.cnt[buf][wv][1]; }
        n_in = __reduce_add_sync(0xffffffffu, ci);
        n_out = __reduce_add_sync(0xffffffffu, co);
      };
      // some visible patches fell outside the current image (or EVAL wants H): H_ = sum over the patches that
      // contributed in this pass ("slow path"; all threads, one block barrier inside)
      auto slow_sum_h = [&]() {
        double q_sxx[FPT], q_sxy[FPT], q_syy[FPT];
#pragma unroll
        for (int k = 0; k < FPT; ++k) {
          q_sxx[k] = q_sxy[k] = q_syy[k] = 0.0;
          if (!((in_mask >> k) & 1u)) continue;
          const int slot = tid + k * T;
#pragma unroll
          for (int p = 0; p < kPatchArea; ++p) {
            float2 gr;
            if constexpr (BQ) {
              const int yy = p >> 2, xx = p & 3;
              const float gh = ((stale_mask >> k) & 1u) ? 0.f : 0.5f;
              gr.x = __fmul_rn(gh, __fsub_rn(pat_ref[bq_idx(yy + 1, xx + 2) * SA + slot], pat_ref[bq_idx(yy + 1, xx) * SA + slot]));
              gr.y = __fmul_rn(gh, __fsub_rn(pat_ref[bq_idx(yy + 2, xx + 1) * SA + slot], pat_ref[bq_idx(yy, xx + 1) * SA + slot]));
            } else {
              gr = pat_dxy[p * SA + slot];
            }
            const double dx = (double)gr.x, dy = (double)gr.y;
            q_sxx[k] = fma(dx, dx, q_sxx[k]);
            q_sxy[k] = fma(dx, dy, q_sxy[k]);
            q_syy[k] = fma(dy, dy, q_syy[k]);
          }
        }
        pair_sum_h_to_warp0<FPT, CS, XG, SS, SH>(
            [&](int k, double& x, double& y, double& zi, double& sxx, double& sxy, double& syy, double& cnt) {
              { double z_; feat_xyz(k, (int)threadIdx.x + k * (int)blockDim.x, x, y, z_); zi = feat_zi(k, z_); } sxx = q_sxx[k]; sxy = q_sxy[k]; syy = q_syy[k]; cnt = 0.0;
              asm volatile("" : "+d"(x), "+d"(y), "+d"(zi));  // see the per-level call: no hoisting into local memory
            },
            s, nwarps, P.xg, pair);
      };
      auto eval_outputs = [&]() {  // EVAL, leader: computeResiduals' scalar outputs
        for (int k = 0; k < 6; ++k) P.Jres_out[k] = -(tot[k] * jscale);
        const float chi2f = (float)tot[6];
        *P.chi2_out = (double)(chi2f / (float)(n_in * kPatchArea));
        *P.n_meas_out = (long long)n_in * kPatchArea;
        s.n_in_last = n_in;
      };
      auto fast_path_solver = [&]() -> const Solver6* {  // every visible patch contributed: H_ is the level's H
        if (n_in == 0) {  // H_ == 0 exactly: Eigen's LDLT yields x = 0
          if (leader) {
            for (int k = 0; k < 36; ++k) s.Hs[k] = 0.0;
            s.h_is_tot = 0;
          }
          return nullptr;
        }
        if (leader) s.h_is_tot = 1;
        return sol_level;
      };
      SIA_DBG(long long ti2 = 0;)
      if constexpr (CS == 1) {
        // One CTA per pair: warp 0 alone adds the per-warp partials and runs the Gauss-Newton tail, then publishes the
        // new pose through shared memory (barrier B); the other warps would only replicate that work on the same SM.
        bool slow = false;
        if (warp == 0) {
          compute_totals();
          if (XG && P.xg.world > 1) {  // feature split over GPUs: sum the 7 doubles + 2 counts of this pass over the ranks
            double v = 0.0;
#pragma unroll
            for (int e = 0; e < 7; ++e) v = lane == e ? tot[e] : v;
            xg_allreduce<8>(P.xg, pair, &s.xg_seq, v, n_in, n_out);
#pragma unroll
            for (int e = 0; e < 7; ++e) tot[e] = __shfl_sync(0xffffffffu, v, e);
          }
      SIA_DBG(if ((SVO_SIA_DEBUG && P.debug) && tid == 0) { ti2 = clock64(); s.tk[1] += ti1 - ti0; s.tk[2] += ti2 - ti1; })
          slow = EVAL || (n_out > 0 && n_in > 0);
          if (!slow) {
            const Solver6* S6 = fast_path_solver();
            gn_tail(s, g, S6, tot, jscale, n_in, iter, level, P.eps, cta_leader, leader, P.trace, P.trace_cap, chi2_prev, stop,
                    done, R, t);
            if (lane == 0) {
#pragma unroll
              for (int k = 0; k < 9; ++k) s.pub[k] = R[k];
              s.pub[9] = t[0]; s.pub[10] = t[1]; s.pub[11] = t[2];
              s.pub_done = done;
            }
          }
          if (lane == 0) s.pub_slow = slow ? 1 : 0;
        }
        __syncthreads();  // barrier B
        if (s.pub_slow) {
          slow_sum_h();
          if (warp == 0) {
            warp_scale_and_factor(s.sums, jscale * jscale, s.Hs, s.sol_cur);
            if (leader) s.h_is_tot = 0;
            if (EVAL) {
              if (leader) eval_outputs();
            } else {
              gn_tail(s, g, &s.sol_cur, tot, jscale, n_in, iter, level, P.eps, cta_leader, leader, P.trace, P.trace_cap,
                      chi2_prev, stop, done, R, t);
              if (lane == 0) {
#pragma unroll
                for (int k = 0; k < 9; ++k) s.pub[k] = R[k];
                s.pub[9] = t[0]; s.pub[10] = t[1]; s.pub[11] = t[2];
                s.pub_done = done;
              }
            }
          }
          __syncthreads();  // barrier C
        }
        if (!EVAL) done = s.pub_done;  // the new pose stays in s.pub: sia_transform reads it there
      } else {
        // Cluster per pair: every warp of every CTA adds the partials in the same order and runs the tail redundantly in
        // registers (bit-identical everywhere), which saves a second cluster barrier + DSMEM broadcast per iteration.
        compute_totals();
      SIA_DBG(if ((SVO_SIA_DEBUG && P.debug) && tid == 0) { ti2 = clock64(); s.tk[1] += ti1 - ti0; s.tk[2] += ti2 - ti1; })
        const Solver6* S6;
        const bool slow = EVAL || (n_out > 0 && n_in > 0);
        if (!slow) {
          S6 = fast_path_solver();
        } else {
          slow_sum_h();
          if (warp == 0) {
            warp_scale_and_factor(s.sums, jscale * jscale, s.Hs, s.sol_cur);
            if (leader) s.h_is_tot = 0;
          }
          if (EVAL && leader) eval_outputs();
          __syncthreads();  // s.sol_cur is visible to every warp of this CTA (each CTA factorised its own copy)
          S6 = &s.sol_cur;
        }
        if (!EVAL)
          gn_tail(s, g, S6, tot, jscale, n_in, iter, level, P.eps, cta_leader, leader, P.trace, P.trace_cap, chi2_prev, stop,
                  done, R, t);
      }
      if (EVAL) {
#pragma unroll
        for (int k = 0; k < FPT; ++k) {
          const int i = tid + k * T;
          if (i < n_loc) {
            P.in_image_out[fbase + i] = (in_mask >> k) & 1u;
            if (!((vis_mask >> k) & 1u))
              for (int p = 0; p < kPatchArea; ++p)
                P.residuals_out[(size_t)(fbase + i) * kPatchArea + p] = __int_as_float(0x7fc00000);
            for (int p = 0; p < kPatchArea; ++p)
              P.ref_patch_out[(size_t)(fbase + i) * kPatchArea + p] = BQ ? pat_ref[bq_idx((p >> 2) + 1, (p & 3) + 1) * SA + i] : pat_ref[p * SA + i];
          }
        }
        break;
      }
      ++g;
      SIA_DBG(if ((SVO_SIA_DEBUG && P.debug) && tid == 0) s.tk[3] += clock64() - ti2;)
      if (done) break;
    }
    // stage region / patches are rewritten by the next level; the leader's state writes are visible.  Upfront variant: only
    // CTA-local buffers are reused between levels (every level has its own patch arrays, the exchange buffers alternate by
    // the parity of the running iteration counter), so the CTAs of the pair need not meet here
    if constexpr (UP) __syncthreads();
    else pair_sync<CS>();
  }
  if constexpr (UP) pair_sync<CS>();  // no CTA of the cluster exits while another may still write into its shared memory

  // ---- outputs ---------------------------------------------------------------------------------
#pragma unroll
  for (int k = 0; k < FPT; ++k) {
    const int i = tid + k * T;
    if (i < n_loc && P.visible_out) P.visible_out[job.feat_off + fbase + i] = (vis_mask >> k) & 1u;
  }
  if (leader) {
    if (P.T_out) pose_to_rt12(s.st[g & 1u].model, P.T_out + 12 * (size_t)pair);
    if (P.H_out)
      for (int k = 0; k < 36; ++k)  // H_ of the last residual pass: the level's H (of the finest level run) or the slow path's
        P.H_out[36 * (size_t)pair + k] = s.h_is_tot ? (UP ? up.Htot[lvl_hi - lvl_lo][k] : s.Htot[k]) : s.Hs[k];
    if (P.stats) {
      svo_b200_sia_stats st;
      st.n_iters = s.n_iters; st.sum_visible = s.sum_vis; st.sum_in_image = s.sum_in;
      st.n_tracked = s.n_in_last;  // n_meas_/patch_area_ of the last pass (:74)
      P.stats[pair] = st;
    }
    if (P.n_trace) *P.n_trace = s.n_trace;
    if (CS == 1 && P.xg.world > 1) P.xg.peer[P.xg.rank][pair].xseq = s.xg_seq;
#if SVO_SIA_DEBUG
    if (SVO_SIA_DEBUG && P.debug && (pair == 0 || pair == (int)(gridDim.x / CS) / 2 || pair == (int)(gridDim.x / CS) - 1))
      printf("[sia dbg] pair %d iters %d cycles: setup %lld pass %lld reduce %lld tail %lld (solve %lld update %lld) total %lld | setup parts: loads+patches %lld hsum %lld factor+wait %lld upfront %lld prologue %lld\n", pair, s.n_iters,
             s.tk[0], s.tk[1], s.tk[2], s.tk[3], s.tk[5], s.tk[6], (long long)clock64() - s.tk[4], s.tk[7], s.tkx[0], s.tkx[1], s.tkx[2], s.tkx[3]);
#endif
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
  // The batch API owns its staging buffers: stage / run / fetch may be interleaved with any other entry point of the
  // context (align, pose optimizer, depth filter ... reuse the context's generic scratch) without clobbering a staged
  // batch.
  DevBuf d_in, d_out;
  HostBuf h_in, h_out;
  size_t o_T = 0, o_H = 0, o_vis = 0, o_stats = 0, out_bytes = 0;
  int threads = 0, fpt = 1, cluster = 1;
  bool upfront = false;  // cluster geometry with all levels prepared before the first iteration (SiaUpT)
  bool bq = false;       // throughput geometry with the 32-value patch cache, four CTAs per SM
  size_t smem = 0;
  bool staged = false;
};

void sia_batch_free(svo_b200_ctx* ctx) {
  if (ctx->sia) {
    if (ctx->sia->d_in.p) cudaFree(ctx->sia->d_in.p);
    if (ctx->sia->d_out.p) cudaFree(ctx->sia->d_out.p);
    if (ctx->sia->h_in.p) cudaFreeHost(ctx->sia->h_in.p);
    if (ctx->sia->h_out.p) cudaFreeHost(ctx->sia->h_out.p);
  }
  delete ctx->sia;
  ctx->sia = nullptr;
}

void sia_split_free(svo_b200_ctx* ctx) {
  for (int r = 0; r < 8; ++r) {
    if (ctx->xg_peer_ipc[r] && ctx->xg_peer[r]) cudaIpcCloseMemHandle(ctx->xg_peer[r]);
    ctx->xg_peer[r] = nullptr;
    ctx->xg_peer_ipc[r] = false;
  }
  if (ctx->xg_buf) cudaFree(ctx->xg_buf);
  ctx->xg_buf = nullptr;
  ctx->xg_world = 1; ctx->xg_rank = 0; ctx->xg_pairs = 0; ctx->xg_connected = false;
}

// tuning knobs, read once from the environment (defaults are the measured best)
static int g_sia_minb = 2;        // SVO_B200_SIA_MINB: resident CTAs per SM the <=320-feature kernel is compiled for
static int g_sia_stage_kb = 20;   // SVO_B200_SIA_STAGE_KB: minimum staging region (holds level >= 2 of 640x480)
static int g_sia_windows = 1;     // SVO_B200_SIA_WINDOWS=0: fine levels gather from global memory instead of cp.async windows
static int g_sia_prefetch = 1;    // SVO_B200_SIA_PREFETCH=0: no bulk L2 prefetch of the next (coarse, TMA-staged) current image.
                                  // (Per-feature L2 prefetches of the next level's footprints were measured and removed: the extra
                                  // uncoalesced requests cost more L1 time than the DRAM latency they hide.)
static int g_sia_cluster = -1;    // SVO_B200_SIA_CLUSTER: force the CTAs per pair (1, 2, 4, 8); -1 = by batch size
static int g_sia_upfront = 1;     // SVO_B200_SIA_UPFRONT=0: the cluster geometry prepares each level when it reaches it (round-2a behaviour)
static int g_sia_bq = 0;          // SVO_B200_SIA_BQ=1: throughput geometry with the 128-byte patch cache at four CTAs per SM
static int g_sia_async = 1;       // SVO_B200_SIA_ASYNC=0: the upfront variant exchanges its sums through barrier.cluster like the others
static int g_sia_plain = 1;       // SVO_B200_SIA_PLAIN=0: the undistorted pinhole runs the general-camera instantiation too
static int g_sia_fpt2 = 1;        // SVO_B200_SIA_FPT2: <= 320 features per CTA as 160 threads x 2 features: 1 = three CTAs per SM (default,
                                  // measured best for full batches), 2 = two CTAs per SM with windows, 0 = 320 threads x 1 feature

static void read_env_once() {
  static bool env_read = false;
  if (env_read) return;
  env_read = true;
  if (const char* e = getenv("SVO_B200_SIA_MINB")) g_sia_minb = atoi(e) == 3 ? 3 : 2;
  if (const char* e = getenv("SVO_B200_SIA_STAGE_KB")) g_sia_stage_kb = atoi(e) > 0 ? atoi(e) : 20;
  if (const char* e = getenv("SVO_B200_SIA_WINDOWS")) g_sia_windows = atoi(e) != 0;
  if (const char* e = getenv("SVO_B200_SIA_PREFETCH")) g_sia_prefetch = atoi(e) != 0;
  if (const char* e = getenv("SVO_B200_SIA_CLUSTER")) g_sia_cluster = atoi(e);
  if (const char* e = getenv("SVO_B200_SIA_FPT2")) g_sia_fpt2 = atoi(e);
  if (const char* e = getenv("SVO_B200_SIA_PLAIN")) g_sia_plain = atoi(e) != 0;
  if (const char* e = getenv("SVO_B200_SIA_UPFRONT")) g_sia_upfront = atoi(e) != 0;
  if (const char* e = getenv("SVO_B200_SIA_ASYNC")) g_sia_async = atoi(e) != 0;
  if (const char* e = getenv("SVO_B200_SIA_BQ")) g_sia_bq = atoi(e) != 0;
}

// Launch geometry for a batch of B pairs with at most max_feat features each.
//   cluster == 1: one CTA per pair, 320 / 384 / 512 threads (one feature per thread up to 512, two up to 1024);
//   cluster  > 1: (small batches) the pair's features are split over `cluster` CTAs of 96 threads.
static size_t sia_shared_bytes(int threads, int cluster) {
  if (cluster == 2) return sizeof(SiaSharedT<3, 2>);
  if (cluster == 4) return sizeof(SiaSharedT<3, 4>);
  if (cluster == 8) return sizeof(SiaSharedT<3, 8>);
  if (threads == 160) return sizeof(SiaSharedT<5, 1>);
  if (threads == 320) return sizeof(SiaSharedT<10, 1>);
  if (threads == 384) return sizeof(SiaSharedT<12, 1>);
  return sizeof(SiaSharedT<16, 1>);
}

static size_t sia_upfront_bytes(int cluster) {
  if (cluster == 2) return sizeof(SiaUpT<3, 2>);
  if (cluster == 4) return sizeof(SiaUpT<3, 4>);
  return sizeof(SiaUpT<3, 8>);
}

static int pick_launch(svo_b200_ctx* ctx, int B, int max_feat, int n_lvl, int& threads, int& fpt, int& cluster, bool& upfront,
                       bool& bq, int& stage_cap, size_t& smem) {
  read_env_once();
  if (max_feat > 1024)
    return set_err(ctx, SVO_B200_ELIMIT, "sparse_img_align: %d features per pair > 1024 (shared-memory patch cache)", max_feat);
  cluster = 1;
  int want = ctx->sia_cluster >= 0 ? ctx->sia_cluster : g_sia_cluster;
  if (ctx->xg_connected) {
    // feature split over GPUs: the ranks' CTAs of a pair wait for each other, so every CTA must be resident at once
    if (B > ctx->xg_pairs || B > 2 * ctx->sm_count)
      return set_err(ctx, SVO_B200_ELIMIT, "sia split: %d pairs per launch exceed the split's capacity (%d) or the resident CTAs (%d)",
                     B, ctx->xg_pairs, 2 * ctx->sm_count);
    want = 1;
  }
  // small batch (live streams, BASELINE configs[4]'s 32 pairs per GPU): spread each pair over 4 SMs while every CTA
  // still has an SM of its own (measured: 4*B = 296 CTAs, two per SM, is already slower than one 320-thread CTA per pair)
  if (want < 0) want = (B * 4 <= ctx->sm_count) ? 4 : 1;
  // full batches (more pairs than 2 per SM): 160 threads x 2 features, three CTAs per SM; in between, 320 x 1 with windows
  int fpt2 = ctx->sia_fpt > 0 ? (ctx->sia_fpt == 2 ? 1 : 0) : g_sia_fpt2;
  // (round 2b: with its state in shared memory and its loops rolled the 160 x 2 geometry also wins below two CTAs per SM --
  // 148 pairs 117 vs 131 us, 296 pairs 145 vs 165 us, 444 pairs 178 vs 304 us, profiles/r02l -- so the 320 x 1 geometry is
  // left for > 304 features per pair, the multi-GPU split and explicit requests)
  if (ctx->xg_connected) fpt2 = 0;
  if (want > 1 && max_feat <= 96 * want && (want == 2 || want == 4 || want == 8)) cluster = want;
  if (cluster > 1) {
    fpt = 1;
    threads = 96;
  } else {
    fpt = max_feat <= 512 ? 1 : 2;
    threads = ((max_feat + fpt - 1) / fpt + 31) / 32 * 32;
    // the kernels are instantiated for MAXT in {320, 384, 512}; launching exactly MAXT threads makes the
    // slot count S = MAXT*FPT a compile-time constant (immediate shared-memory offsets)
    threads = fpt == 1 ? (threads <= 320 ? 320 : threads <= 384 ? 384 : 512) : 512;
    if (fpt2 && max_feat <= 304) { fpt = 2; threads = 160; }  // its shared arrays are allocated for 304 slots (kernel: SA)
  }
  const bool throughput_geom = cluster == 1 && fpt == 2 && threads == 160;
  bq = throughput_geom && g_sia_bq && fpt2 != 2;
  const int slots = throughput_geom ? 304 : threads * fpt;
  size_t base = ((sia_shared_bytes(threads, cluster) + 15) & ~size_t(15)) + (size_t)(bq ? 32 : 3 * kPatchArea) * slots * sizeof(float);
  if (throughput_geom) base += (size_t)3 * slots * sizeof(double);  // xyz_ref of every feature (kernel: st_xyz)
  // cluster geometry with every CTA alone on its SM: one patch array set per level, everything pose independent prepared
  // before the first iteration (the launch asks for the 4-CTA instantiation; 2 and 8 keep the per-level flow)
  upfront = false;
  if (cluster == 4 && g_sia_upfront && ctx->sia_upfront != 0 && B * cluster <= ctx->sm_count && n_lvl >= 1 && n_lvl <= SVO_B200_MAX_LEVELS) {
    const size_t base_up = ((sia_shared_bytes(threads, cluster) + 15) & ~size_t(15)) + ((sia_upfront_bytes(cluster) + 15) & ~size_t(15)) +
                           (size_t)n_lvl * 3 * kPatchArea * slots * sizeof(float);
    if (base_up + (size_t)g_sia_stage_kb * 1024 + 1024 <= (size_t)ctx->max_smem_optin) {
      upfront = true;
      base = base_up;
    }
  }
  // shared memory one CTA may use so that the intended number of CTAs stays resident per SM (228 KB per SM, 1 KB
  // reserved per CTA)
  const int resident = upfront ? 1 : cluster > 1 ? 2 : (threads == 160 ? (bq ? 4 : fpt2 == 2 ? 2 : 3) : threads <= 384 ? 2 : 1);
  size_t budget = (size_t)ctx->max_smem_optin;
  const size_t per_cta = (size_t)(228 * 1024) / resident - 1024;
  if (per_cta < budget) budget = per_cta;
  if (base + 1024 > budget)
    return set_err(ctx, SVO_B200_ELIMIT, "sparse_img_align: %d features need %zu B of shared memory", max_feat, base);
  // staging region: the coarse current-level images (TMA; 20 KB holds level >= 2 of 640x480) or, at the fine levels,
  // one 96-byte window per feature slot
  size_t cap = (size_t)g_sia_stage_kb * 1024;
  const size_t win = (size_t)kWinBytes * slots;
  if (g_sia_windows && win > cap && base + win <= budget) cap = win;
  if (base + cap > budget) cap = (budget - base) & ~size_t(15);
  stage_cap = (int)cap;
  smem = base + cap;
  return 0;
}

template <bool EVAL>
static int launch_sia(svo_b200_ctx* ctx, const SiaParams& P, int B, int threads, int fpt, int cluster, bool upfront, bool bq, size_t smem) {
  auto go = [&](auto kern) -> int {
    SVO_CUDA_CHECK(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // ask for the full shared-memory carveout so that two CTAs of ~95 KB fit one SM
    SVO_CUDA_CHECK(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout,
                                             (int)cudaSharedmemCarveoutMaxShared));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(B * cluster));
    cfg.blockDim = dim3((unsigned)threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = ctx->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)cluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = cluster > 1 ? 1 : 0;
    kt_begin(ctx);
    SVO_CUDA_CHECK(ctx, cudaLaunchKernelEx(&cfg, kern, P));
    kt_end(ctx);
    ctx->launches++;
    SVO_CUDA_CHECK(ctx, cudaGetLastError());
    return 0;
  };
  // the undistorted pinhole gets its own instantiation of the geometries that carry the throughput / latency figures
  // (the general-camera code also handles it; EVAL and the rarely used geometries are compiled once)
  const bool plain = !EVAL && !P.cam.distorted && P.cam.model == SVO_B200_CAM_PINHOLE && g_sia_plain;
  if (cluster == 2) return go(sia_kernel<1, EVAL, 96, 2, 2, true, false>);
  if (cluster == 4) {
    if (upfront && !EVAL)
      return plain ? go(sia_kernel<1, EVAL, 96, 1, 4, EVAL, !EVAL>) : go(sia_kernel<1, EVAL, 96, 1, 4, true, !EVAL>);
    return plain ? go(sia_kernel<1, EVAL, 96, 2, 4, EVAL, false>) : go(sia_kernel<1, EVAL, 96, 2, 4, true, false>);
  }
  if (cluster == 8) return go(sia_kernel<1, EVAL, 96, 2, 8, true, false>);
  // <= 384 threads: cap registers so that two CTAs are resident per SM
  if (fpt == 1) {
    if (threads <= 320 && g_sia_minb == 3) return go(sia_kernel<1, EVAL, 320, 3, 1, true, false>);
    if (threads <= 320) return plain ? go(sia_kernel<1, EVAL, 320, 2, 1, EVAL, false>) : go(sia_kernel<1, EVAL, 320, 2, 1, true, false>);
    if (threads <= 384) return go(sia_kernel<1, EVAL, 384, 2, 1, true, false>);
    return go(sia_kernel<1, EVAL, 512, 1, 1, true, false>);
  }
  if (threads == 160 && bq) return plain ? go(sia_kernel<2, EVAL, 160, 4, 1, EVAL, false>) : go(sia_kernel<2, EVAL, 160, 4, 1, true, false>);
  if (threads == 160) return plain ? go(sia_kernel<2, EVAL, 160, 3, 1, EVAL, false>) : go(sia_kernel<2, EVAL, 160, 3, 1, true, false>);
  return go(sia_kernel<2, EVAL, 512, 1, 1, true, false>);
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
  int rc = cam_to_dev(ctx, cam, P.cam);
  if (rc) return rc;
  P.max_level = opt->max_level; P.min_level = opt->min_level; P.n_iter = opt->n_iter; P.eps = opt->eps;
  P.debug = getenv("SVO_B200_SIA_DEBUG") ? 1 : 0;
  read_env_once();
  P.use_windows = g_sia_windows;
  P.use_prefetch = g_sia_prefetch;
  P.async_xchg = g_sia_async;
  P.xg.rank = 0; P.xg.world = 1;
  if (ctx->xg_connected) {
    P.xg.rank = ctx->xg_rank; P.xg.world = ctx->xg_world;
    for (int r = 0; r < ctx->xg_world; ++r) P.xg.peer[r] = static_cast<XgPair*>(ctx->xg_peer[r]);
  }
  return 0;
}

}  // namespace svo

using namespace svo;

extern "C" {

int svo_b200_sia_split_create(svo_b200_ctx* ctx, int rank, int world, int max_pairs, void* ipc_handle_out, void** local_ptr_out) {
  if (!ctx || world < 1 || world > kMaxSplit || rank < 0 || rank >= world || max_pairs < 1)
    return set_err(ctx, SVO_B200_EINVAL, "sia_split_create: need 1 <= world <= %d, 0 <= rank < world, max_pairs >= 1", kMaxSplit);
  cudaSetDevice(ctx->device);
  sia_split_free(ctx);
  const size_t bytes = sizeof(XgPair) * (size_t)max_pairs;
  SVO_CUDA_CHECK(ctx, cudaMalloc(&ctx->xg_buf, bytes));
  SVO_CUDA_CHECK(ctx, cudaMemset(ctx->xg_buf, 0, bytes));
  ctx->xg_rank = rank; ctx->xg_world = world; ctx->xg_pairs = max_pairs;
  if (ipc_handle_out) {
    static_assert(sizeof(cudaIpcMemHandle_t) == SVO_B200_IPC_HANDLE_BYTES, "IPC handle size");
    cudaIpcMemHandle_t h;
    SVO_CUDA_CHECK(ctx, cudaIpcGetMemHandle(&h, ctx->xg_buf));
    memcpy(ipc_handle_out, &h, sizeof(h));
  }
  if (local_ptr_out) *local_ptr_out = ctx->xg_buf;
  return 0;
}

int svo_b200_sia_split_connect(svo_b200_ctx* ctx, const void* ipc_handles, void* const* in_process_ptrs) {
  if (!ctx || !ctx->xg_buf || (!ipc_handles && !in_process_ptrs))
    return set_err(ctx, SVO_B200_EINVAL, "sia_split_connect: call sia_split_create first and pass the peers' handles or pointers");
  cudaSetDevice(ctx->device);
  for (int r = 0; r < ctx->xg_world; ++r) {
    if (r == ctx->xg_rank) { ctx->xg_peer[r] = ctx->xg_buf; continue; }
    if (in_process_ptrs) {
      if (!in_process_ptrs[r]) return set_err(ctx, SVO_B200_EINVAL, "sia_split_connect: NULL pointer for rank %d", r);
      ctx->xg_peer[r] = in_process_ptrs[r];
    } else {
      cudaIpcMemHandle_t h;
      memcpy(&h, static_cast<const uint8_t*>(ipc_handles) + (size_t)r * sizeof(h), sizeof(h));
      SVO_CUDA_CHECK(ctx, cudaIpcOpenMemHandle(&ctx->xg_peer[r], h, cudaIpcMemLazyEnablePeerAccess));
      ctx->xg_peer_ipc[r] = true;
    }
  }
  // the buffers start from a known state on every rank (sequence counters 0): the caller connects all ranks before any
  // of them launches (a barrier of the launcher's own, e.g. torch.distributed.barrier)
  SVO_CUDA_CHECK(ctx, cudaMemset(ctx->xg_buf, 0, sizeof(XgPair) * (size_t)ctx->xg_pairs));
  SVO_CUDA_CHECK(ctx, cudaDeviceSynchronize());
  ctx->xg_connected = ctx->xg_world > 1;
  return 0;
}

int svo_b200_sia_split_destroy(svo_b200_ctx* ctx) {
  if (!ctx) return SVO_B200_EINVAL;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  sia_split_free(ctx);
  return 0;
}

int svo_b200_sia_config(svo_b200_ctx* ctx, int ctas_per_pair, int features_per_thread) {
  if (!ctx) return SVO_B200_EINVAL;
  if (!(ctas_per_pair == -1 || ctas_per_pair == 1 || ctas_per_pair == 2 || ctas_per_pair == 4 || ctas_per_pair == 8) ||
      features_per_thread < 0 || features_per_thread > 2)
    return set_err(ctx, SVO_B200_EINVAL, "sia_config: ctas_per_pair must be -1, 1, 2, 4 or 8 and features_per_thread 0, 1 or 2");
  ctx->sia_cluster = ctas_per_pair;
  ctx->sia_fpt = features_per_thread;
  return 0;
}

int svo_b200_sia_upfront(svo_b200_ctx* ctx, int mode) {
  if (!ctx) return SVO_B200_EINVAL;
  if (mode < -1 || mode > 1) return set_err(ctx, SVO_B200_EINVAL, "sia_upfront: mode must be -1, 0 or 1");
  ctx->sia_upfront = mode;
  return 0;
}

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
  rc = pick_launch(ctx, B, st.max_feat, opt->max_level - opt->min_level + 1, st.threads, st.fpt, st.cluster, st.upfront, st.bq,
                   stage_cap, st.smem);
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
  if ((rc = ensure_host(ctx, st.h_in, st.in_bytes))) return rc;
  if ((rc = ensure_dev(ctx, st.d_in, st.in_bytes))) return rc;
  // a previous async copy out of the pinned buffer must be finished before it is rewritten
  SVO_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  uint8_t* hin = static_cast<uint8_t*>(st.h_in.p);
  uint8_t* din = static_cast<uint8_t*>(st.d_in.p);
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
  if ((rc = ensure_dev(ctx, st.d_out, st.out_bytes))) return rc;
  if ((rc = ensure_host(ctx, st.h_out, st.out_bytes))) return rc;
  uint8_t* dout = static_cast<uint8_t*>(st.d_out.p);
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
  return launch_sia<false>(ctx, st.P, st.B, st.threads, st.fpt, st.cluster, st.upfront, st.bq, st.smem);
}

int svo_b200_sia_batch_fetch(svo_b200_ctx* ctx, double* T_out, uint8_t* visible_out, double* H_out,
                             svo_b200_sia_stats* stats_out) {
  if (!ctx || !ctx->sia || !ctx->sia->staged) return set_err(ctx, SVO_B200_EINVAL, "sia_batch_fetch: nothing staged");
  cudaSetDevice(ctx->device);
  SiaBatchState& st = *ctx->sia;
  SVO_CUDA_CHECK(ctx, cudaMemcpyAsync(st.h_out.p, st.d_out.p, st.out_bytes, cudaMemcpyDeviceToHost, ctx->stream));
  SVO_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  if (ctx->xg_connected) {  // did every peer take part in every exchange?
    std::vector<unsigned> err((size_t)st.B, 0u);
    SVO_CUDA_CHECK(ctx, cudaMemcpy2D(err.data(), sizeof(unsigned), static_cast<const uint8_t*>(ctx->xg_buf) + offsetof(XgPair, err),
                                     sizeof(XgPair), sizeof(unsigned), (size_t)st.B, cudaMemcpyDeviceToHost));
    for (int b = 0; b < st.B; ++b)
      if (err[b]) return set_err(ctx, SVO_B200_ECUDA, "sia split: a peer rank did not arrive at an exchange of pair %d (timeout); reconnect the split", b);
  }
  const uint8_t* h = static_cast<const uint8_t*>(st.h_out.p);
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
  st.staged = false;  // the single-pair call leaves nothing staged behind
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
  const size_t nslots = (size_t)st.P.slots * (size_t)st.cluster;
  Carver c;
  const size_t o_vin = c.take(N), o_rp = c.take(sizeof(float) * 16 * nslots),
               o_res = c.take(sizeof(float) * 16 * nslots), o_in = c.take(N),
               o_j = c.take(sizeof(double) * 6), o_c = c.take(sizeof(double)), o_n = c.take(sizeof(long long));
  if ((rc = ensure_dev(ctx, ctx->d_scratch, c.off))) return rc;
  uint8_t* ds = static_cast<uint8_t*>(ctx->d_scratch.p);
  SVO_CUDA_CHECK(ctx, cudaMemcpyAsync(ds + o_vin, visible_io, N, cudaMemcpyHostToDevice, ctx->stream));
  SVO_CUDA_CHECK(ctx, cudaMemsetAsync(ds + o_res, 0xff, sizeof(float) * 16 * nslots, ctx->stream));
  st.P.eval_level = level;
  st.P.visible_in = ds + o_vin;
  st.P.ref_patch_out = reinterpret_cast<float*>(ds + o_rp);
  st.P.residuals_out = reinterpret_cast<float*>(ds + o_res);
  st.P.in_image_out = ds + o_in;
  st.P.Jres_out = reinterpret_cast<double*>(ds + o_j);
  st.P.chi2_out = reinterpret_cast<double*>(ds + o_c);
  st.P.n_meas_out = reinterpret_cast<long long*>(ds + o_n);
  if ((rc = launch_sia<true>(ctx, st.P, 1, st.threads, st.fpt, st.cluster, false, st.bq, st.smem))) return rc;
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
