// rpg_svo_b200/csrc/ctx.h -- internal: context, frame and helper declarations shared by the .cu files.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/svo_b200.h"

struct svo_b200_frame {
  int width = 0, height = 0, n_levels = 0;
  int w[SVO_B200_MAX_LEVELS] = {0}, h[SVO_B200_MAX_LEVELS] = {0};
  uint8_t* base = nullptr;  // one allocation, levels at 256-byte aligned offsets
  size_t off[SVO_B200_MAX_LEVELS] = {0};
  size_t bytes = 0;
  bool pooled = false;  // memory belongs to a svo_b200_frame_pool
  uint8_t* lv[SVO_B200_MAX_LEVELS] = {nullptr};  // level pointers (base + off[l], or into a pool's per-level slabs)
  uint8_t* lvl(int l) const { return lv[l]; }
};

struct svo_b200_frame_pool {
  int count = 0;
  int n_levels = 0;
  // one slab per pyramid level: level l of frame i at slab[l] + i*stride[l].  Level 0 of consecutive
  // frames is contiguous up to the 256-byte rounding, so a window uploads as ONE strided copy whose
  // rows are whole images.
  uint8_t* slab[SVO_B200_MAX_LEVELS] = {nullptr};
  size_t stride[SVO_B200_MAX_LEVELS] = {0};
  uint8_t* mem = nullptr;  // the single allocation behind all slabs
  std::vector<svo_b200_frame> frames;
};

// Grow-only device / pinned-host buffers owned by the context.
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
};
struct HostBuf {
  void* p = nullptr;
  size_t cap = 0;
};

namespace svo {
struct SiaBatchState;  // defined in sparse_align.cu
}

struct svo_b200_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  std::string err;
  uint64_t launches = 0;
  int sm_count = 0;
  int max_smem_optin = 0;
  // generic scratch (each entry point carves what it needs)
  DevBuf d_in, d_out, d_scratch;
  HostBuf h_in, h_out;
  svo::SiaBatchState* sia = nullptr;
  cudaEvent_t ev_k0 = nullptr, ev_k1 = nullptr;  // around the kernel(s) of the last entry point (svo_b200_last_kernel_ms)
  int pyramid_rule = SVO_B200_PYR_X86;  // svo_b200_set_pyramid_rule
  // feature split of single pairs over GPUs (svo_b200_sia_split_*): exchange buffers in peer memory
  int xg_rank = 0, xg_world = 1, xg_pairs = 0;
  bool xg_connected = false;
  void* xg_buf = nullptr;      // our exchange buffer
  void* xg_peer[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // rank r's buffer as mapped here
  bool xg_peer_ipc[8] = {false, false, false, false, false, false, false, false};
  int sia_cluster = -1;  // svo_b200_sia_config: CTAs per pair (-1 = by batch size)
  int sia_fpt = 0;       //                      features per thread (0 = automatic)
  int sia_upfront = -1;  // svo_b200_sia_upfront: all levels prepared before the first iteration (-1 = automatic, 0 = off)
};

namespace svo {

int set_err(svo_b200_ctx* ctx, int code, const char* fmt, ...);
int ensure_dev(svo_b200_ctx* ctx, DevBuf& b, size_t bytes);
int ensure_host(svo_b200_ctx* ctx, HostBuf& b, size_t bytes);

#define SVO_CUDA_CHECK(ctx, call)                                                                   \
  do {                                                                                              \
    cudaError_t _e = (call);                                                                        \
    if (_e != cudaSuccess)                                                                          \
      return svo::set_err((ctx), SVO_B200_ECUDA, "%s failed: %s (%s:%d)", #call,                    \
                          cudaGetErrorString(_e), __FILE__, __LINE__);                              \
  } while (0)

// Bump allocator over a byte buffer (host staging mirrored 1:1 on the device).
struct Carver {
  size_t off = 0;
  size_t take(size_t bytes, size_t align = 256) {
    off = (off + align - 1) / align * align;
    size_t o = off;
    off += bytes;
    return o;
  }
};

void sia_batch_free(svo_b200_ctx* ctx);
void sia_split_free(svo_b200_ctx* ctx);
// CUDA events on the context's stream bracketing the kernel launch(es) of an entry point (no copies): the live
// per-kernel device time bench.py's roofline figures divide by
inline void kt_begin(svo_b200_ctx* ctx) { if (ctx->ev_k0) cudaEventRecord(ctx->ev_k0, ctx->stream); }
inline void kt_end(svo_b200_ctx* ctx) { if (ctx->ev_k1) cudaEventRecord(ctx->ev_k1, ctx->stream); }

// Device-side camera ([EXT] vk::PinholeCamera / vk::ATANCamera): the C-ABI parameters plus the derived constants
// the vikit constructors precompute.  Passed by value inside kernel parameter structs.
struct CamDev {
  double fx, fy, cx, cy;
  double fx_inv, fy_inv;
  double d[5];       // pinhole: k1 k2 p1 p2 k3;  ATAN: d[0] = s
  double s_inv, tans, tans_inv;  // ATAN: 1/s, 2 tan(s/2), 1/tans
  int model, distorted;
  int width, height;
};
int cam_to_dev(svo_b200_ctx* ctx, const svo_b200_camera* cam, CamDev& out);

}  // namespace svo
