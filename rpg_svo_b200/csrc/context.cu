// rpg_svo_b200/csrc/context.cu -- context lifetime, frame (image pyramid) residency in HBM and the
// on-device pyramid build.  Replaces the image side of svo::Frame (svo/include/svo/frame.h:40-84,
// svo/src/frame.cpp:48-59,156-165).
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "ctx.h"

namespace svo {

int set_err(svo_b200_ctx* ctx, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (ctx) ctx->err = buf;
  return code;
}

int ensure_dev(svo_b200_ctx* ctx, DevBuf& b, size_t bytes) {
  if (bytes <= b.cap) return 0;
  if (b.p) {
    SVO_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(b.p);
    b.p = nullptr;
    b.cap = 0;
  }
  size_t cap = bytes + bytes / 4 + 4096;
  SVO_CUDA_CHECK(ctx, cudaMalloc(&b.p, cap));
  b.cap = cap;
  return 0;
}

// [EXT] the constants vk::PinholeCamera / vk::ATANCamera derive in their constructors
int cam_to_dev(svo_b200_ctx* ctx, const svo_b200_camera* cam, CamDev& o) {
  memset(&o, 0, sizeof(o));
  if (!cam || cam->width <= 0 || cam->height <= 0 || cam->fx == 0.0 || cam->fy == 0.0)
    return set_err(ctx, SVO_B200_EINVAL, "camera: bad parameters");
  if (cam->model != SVO_B200_CAM_PINHOLE && cam->model != SVO_B200_CAM_ATAN)
    return set_err(ctx, SVO_B200_EINVAL, "camera: unknown model %d", cam->model);
  o.fx = cam->fx; o.fy = cam->fy; o.cx = cam->cx; o.cy = cam->cy;
  o.fx_inv = 1.0 / cam->fx; o.fy_inv = 1.0 / cam->fy;
  o.width = cam->width; o.height = cam->height;
  o.model = cam->model;
  for (int k = 0; k < 5; ++k) o.d[k] = cam->d[k];
  if (cam->model == SVO_B200_CAM_PINHOLE) {
    o.distorted = fabs(cam->d[0]) > 0.0000001;  // vk::PinholeCamera: distortion_(fabs(d0) > 0.0000001)
  } else {
    const double sv = cam->d[0];
    if (sv != 0.0) {  // vk::ATANCamera ctor
      o.tans = 2.0 * tan(sv / 2.0);
      o.tans_inv = 1.0 / o.tans;
      o.s_inv = 1.0 / sv;
      o.distorted = 1;
    }
  }
  return 0;
}

int ensure_host(svo_b200_ctx* ctx, HostBuf& b, size_t bytes) {
  if (bytes <= b.cap) return 0;
  if (b.p) {
    SVO_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFreeHost(b.p);
    b.p = nullptr;
    b.cap = 0;
  }
  size_t cap = bytes + bytes / 4 + 4096;
  SVO_CUDA_CHECK(ctx, cudaMallocHost(&b.p, cap));
  b.cap = cap;
  return 0;
}

// [EXT] vk::halfSample (svo/src/frame.cpp:156-165), both branches of rpg_vikit's vision.cpp:
//   scalar: out = (a + b + c + d) / 4, integer division;
//   SSE2 (what an x86 build runs when in_w % 16 == 0): vertical _mm_avg_epu8, then _mm_avg_epu16 of adjacent columns,
//          i.e. avg(avg(a, c), avg(b, d)) with round-half-up at both steps.
// Packed forms on 32-bit words holding four consecutive pixels of the top / bottom row: the two results come back in
// bytes 0 and 2.
__device__ __forceinline__ uint32_t half2_scalar(uint32_t top, uint32_t bot) {
  const uint32_t s = (top & 0x00ff00ffu) + ((top >> 8) & 0x00ff00ffu) + (bot & 0x00ff00ffu) + ((bot >> 8) & 0x00ff00ffu);
  return (s >> 2) & 0x00ff00ffu;
}
__device__ __forceinline__ uint32_t half2_avg(uint32_t top, uint32_t bot) {
  const uint32_t v = __vavgu4(top, bot);          // per byte (a + c + 1) >> 1
  return __vavgu4(v, v >> 8) & 0x00ff00ffu;       // bytes 0, 2: (v0 + v1 + 1) >> 1, (v2 + v3 + 1) >> 1
}
__device__ __forceinline__ uint32_t half2(uint32_t top, uint32_t bot, bool avg) {
  return avg ? half2_avg(top, bot) : half2_scalar(top, bot);
}
__device__ __forceinline__ uint32_t half1(uint32_t a, uint32_t b, uint32_t c, uint32_t d, bool avg) {  // one output pixel
  return avg ? ((((a + c + 1u) >> 1) + ((b + d + 1u) >> 1) + 1u) >> 1) : ((a + b + c + d) >> 2);
}
// HBM-bound byte kernel: each thread produces 4 output pixels from two 8-byte row segments (coalesced 64-bit loads,
// one 32-bit store).  `avg` selects the SSE2 rounding.
__global__ void half_sample_kernel(const uint8_t* __restrict__ in, int in_w, int in_h,
                                   uint8_t* __restrict__ out, int out_w, int out_h, int avg) {
  const int quads = (out_w + 3) / 4;
  const int total = quads * out_h;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int y = idx / quads, qx = idx - y * quads;
    const int x0 = qx * 4;
    const uint8_t* top = in + (size_t)(2 * y) * in_w + 2 * x0;
    const uint8_t* bot = top + in_w;
    if (x0 + 4 <= out_w && ((in_w & 7) == 0)) {
      const uint2 t = *reinterpret_cast<const uint2*>(top);
      const uint2 b = *reinterpret_cast<const uint2*>(bot);
      const uint32_t q0 = half2(t.x, b.x, avg != 0), q1 = half2(t.y, b.y, avg != 0);
      const uint32_t r = (q0 & 0xffu) | ((q0 >> 8) & 0xff00u) | ((q1 & 0xffu) << 16) | ((q1 << 8) & 0xff000000u);
      *reinterpret_cast<uint32_t*>(out + (size_t)y * out_w + x0) = r;
    } else {
      for (int k = 0; k < 4 && x0 + k < out_w; ++k)
        out[(size_t)y * out_w + x0 + k] = (uint8_t)half1(top[2 * k], top[2 * k + 1], bot[2 * k], bot[2 * k + 1], avg != 0);
    }
  }
}


// Fused pyramid build for a batch of frames laid out with a constant stride: one CTA turns a
// 128x16 tile of level 0 into the matching 64x8 / 32x4 / 16x2 / 8x1 tiles of levels 1..4 (as many as
// the frame has), reading level 0 from HBM exactly once.  Same arithmetic as half_sample_kernel
// (level l+1 = (a+b+c+d)/4 of level l), so the result is identical to the level-by-level build.
struct PyrGeom {
  int w[SVO_B200_MAX_LEVELS], h[SVO_B200_MAX_LEVELS];
  uint8_t* slab[SVO_B200_MAX_LEVELS];            // level l of frame i at slab[l] + i*stride[l]
  unsigned long long stride[SVO_B200_MAX_LEVELS];
  int n_levels;
  int tiles_x;
  unsigned avg_mask;  // bit l: level l is produced with the SSE2 rounding (avg of avg) instead of (a+b+c+d)/4
};
__global__ void __launch_bounds__(128) pyramid_fused_kernel(int first, PyrGeom g) {
  __shared__ __align__(16) uint8_t t0[16][128];
  __shared__ __align__(16) uint8_t t1[8][64];
  __shared__ __align__(16) uint8_t t2[4][32];
  __shared__ __align__(16) uint8_t t3[2][16];
  const size_t fi = (size_t)(first + blockIdx.y);
  const int tx = blockIdx.x % g.tiles_x, ty = blockIdx.x / g.tiles_x;
  const int t = threadIdx.x;
  const int W0 = g.w[0], H0 = g.h[0];
  {  // level-0 tile -> shared (16 B per thread)
    const int r = t >> 3, c = (t & 7) * 16;
    const int y = ty * 16 + r, x = tx * 128 + c;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (y < H0) {
      const uint8_t* src = g.slab[0] + fi * g.stride[0] + (size_t)y * W0 + x;
      if (x + 16 <= W0 && ((W0 & 15) == 0)) {
        v = *reinterpret_cast<const uint4*>(src);
      } else {
        uint8_t tmp[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) tmp[k] = (x + k < W0) ? src[k] : 0;
        v = *reinterpret_cast<uint4*>(tmp);
      }
    }
    *reinterpret_cast<uint4*>(&t0[r][c]) = v;
  }
  __syncthreads();
  if (g.n_levels > 1) {  // level 1: 8 x 64, four pixels per thread
    const int r = t >> 4, c = (t & 15) * 4;
    const uint2 a = *reinterpret_cast<const uint2*>(&t0[2 * r][2 * c]);
    const uint2 b = *reinterpret_cast<const uint2*>(&t0[2 * r + 1][2 * c]);
    const bool avg = (g.avg_mask >> 1) & 1u;
    const uint32_t q0 = half2(a.x, b.x, avg), q1 = half2(a.y, b.y, avg);
    const uint32_t o = (q0 & 0xffu) | ((q0 >> 8) & 0xff00u) | ((q1 & 0xffu) << 16) | ((q1 << 8) & 0xff000000u);
    *reinterpret_cast<uint32_t*>(&t1[r][c]) = o;
    const int y = ty * 8 + r, x = tx * 64 + c, W1 = g.w[1];
    if (y < g.h[1]) {
      uint8_t* dst = g.slab[1] + fi * g.stride[1] + (size_t)y * W1 + x;
      if (x + 4 <= W1 && ((W1 & 3) == 0)) *reinterpret_cast<uint32_t*>(dst) = o;
      else
        for (int k = 0; k < 4; ++k)
          if (x + k < W1) dst[k] = (uint8_t)(o >> (8 * k));
    }
  }
  __syncthreads();
  if (g.n_levels > 2) {  // level 2: 4 x 32
    const int r = t >> 5, c = t & 31;
    const uint8_t o = (uint8_t)half1(t1[2 * r][2 * c], t1[2 * r][2 * c + 1], t1[2 * r + 1][2 * c], t1[2 * r + 1][2 * c + 1],
                                     (g.avg_mask >> 2) & 1u);
    t2[r][c] = o;
    const int y = ty * 4 + r, x = tx * 32 + c;
    if (y < g.h[2] && x < g.w[2]) g.slab[2][fi * g.stride[2] + (size_t)y * g.w[2] + x] = o;
  }
  __syncthreads();
  if (g.n_levels > 3 && t < 32) {  // level 3: 2 x 16
    const int r = t >> 4, c = t & 15;
    const uint8_t o = (uint8_t)half1(t2[2 * r][2 * c], t2[2 * r][2 * c + 1], t2[2 * r + 1][2 * c], t2[2 * r + 1][2 * c + 1],
                                     (g.avg_mask >> 3) & 1u);
    t3[r][c] = o;
    const int y = ty * 2 + r, x = tx * 16 + c;
    if (y < g.h[3] && x < g.w[3]) g.slab[3][fi * g.stride[3] + (size_t)y * g.w[3] + x] = o;
  }
  __syncthreads();
  if (g.n_levels > 4 && t < 8) {  // level 4: 1 x 8
    const int y = ty, x = tx * 8 + t;
    if (y < g.h[4] && x < g.w[4])
      g.slab[4][fi * g.stride[4] + (size_t)y * g.w[4] + x] =
          (uint8_t)half1(t3[0][2 * t], t3[0][2 * t + 1], t3[1][2 * t], t3[1][2 * t + 1], (g.avg_mask >> 4) & 1u);
  }
}


// Level 0 -> level 1 for a batch of frames as a pure streaming kernel: 94 % of the pyramid's bytes move
// here, so it is written against the HBM roofline -- a persistent grid (a few CTAs per SM), each work
// item = 16 level-0 pixels of two consecutive rows (2 x 128-bit loads) -> 8 level-1 pixels (one 64-bit
// store); the 2x2 reductions are formed SIMD-in-register (scalar rule: two 16-bit lanes per word; SSE2 rule: two
// __vavgu4).  Requires W0 % 16 == 0.
__global__ void __launch_bounds__(256) pyramid_l0_l1_stream_kernel(const uint8_t* __restrict__ l0, size_t stride0,
                                                                   uint8_t* __restrict__ l1, size_t stride1, int first,
                                                                   int count, int W0, int H1, int avg) {
  const int items_x = W0 >> 4, W1 = W0 >> 1;
  const long long per_frame = (long long)items_x * H1, total = per_frame * count;
  for (long long it = (long long)blockIdx.x * blockDim.x + threadIdx.x; it < total; it += (long long)gridDim.x * blockDim.x) {
    const int fr = (int)(it / per_frame);
    const int rem = (int)(it - (long long)fr * per_frame);
    const int y = rem / items_x, ix = rem - y * items_x;
    const uint8_t* src = l0 + (size_t)(first + fr) * stride0 + (size_t)(2 * y) * W0 + 16 * ix;
    const uint4 a = __ldg(reinterpret_cast<const uint4*>(src));
    const uint4 b = __ldg(reinterpret_cast<const uint4*>(src + W0));
    const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
    uint32_t o[2] = {0u, 0u};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t q = half2(aw[k], bw[k], avg != 0);        // two results, bytes 0 and 2
      const uint32_t two = (q & 0xffu) | ((q >> 8) & 0xff00u);  // packed into the low 16 bits
      o[k >> 1] |= two << (16 * (k & 1));
    }
    *reinterpret_cast<uint2*>(l1 + (size_t)(first + fr) * stride1 + (size_t)y * W1 + 8 * ix) = make_uint2(o[0], o[1]);
  }
}

// does producing a level from a source level of width `src_w` use vikit's SSE2 rounding?  (x86 rule: width % 16 == 0;
// cv::Mat buffers are always 16-byte aligned)
static inline int pyr_avg(const svo_b200_ctx* ctx, int src_w) {
  return ctx->pyramid_rule == SVO_B200_PYR_X86 && (src_w % 16) == 0;
}

static int build_levels(svo_b200_ctx* ctx, svo_b200_frame* fr, int from_level, bool timed = true) {
  if (timed) kt_begin(ctx);
  for (int l = from_level; l < fr->n_levels; ++l) {
    const int total = ((fr->w[l] + 3) / 4) * fr->h[l];
    if (total <= 0) continue;
    const int threads = 256;
    int blocks = (total + threads - 1) / threads;
    if (blocks > ctx->sm_count * 8) blocks = ctx->sm_count * 8;
    half_sample_kernel<<<blocks, threads, 0, ctx->stream>>>(fr->lvl(l - 1), fr->w[l - 1], fr->h[l - 1],
                                                            fr->lvl(l), fr->w[l], fr->h[l], pyr_avg(ctx, fr->w[l - 1]));
    ctx->launches++;
  }
  if (timed) kt_end(ctx);
  SVO_CUDA_CHECK(ctx, cudaGetLastError());
  return 0;
}

}  // namespace svo

using namespace svo;

extern "C" {

int svo_b200_last_kernel_ms(svo_b200_ctx* ctx, float* ms_out) {
  if (!ctx || !ms_out) return SVO_B200_EINVAL;
  cudaSetDevice(ctx->device);
  SVO_CUDA_CHECK(ctx, cudaEventSynchronize(ctx->ev_k1));
  SVO_CUDA_CHECK(ctx, cudaEventElapsedTime(ms_out, ctx->ev_k0, ctx->ev_k1));
  return 0;
}

int svo_b200_set_pyramid_rule(svo_b200_ctx* ctx, int rule) {
  if (!ctx) return SVO_B200_EINVAL;
  if (rule != SVO_B200_PYR_X86 && rule != SVO_B200_PYR_SCALAR) return set_err(ctx, SVO_B200_EINVAL, "set_pyramid_rule: unknown rule %d", rule);
  ctx->pyramid_rule = rule;
  return 0;
}

const char* svo_b200_version(void) { return "svo_b200 0.1 (sm_100a)"; }

int svo_b200_create(svo_b200_ctx** ctx_out, int device) {
  if (!ctx_out) return SVO_B200_EINVAL;
  *ctx_out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0 || device < 0 || device >= n) {
    fprintf(stderr, "svo_b200_create: no usable CUDA device %d (%s); there is no CPU fallback\n", device,
            e != cudaSuccess ? cudaGetErrorString(e) : "device index out of range");
    return SVO_B200_ECUDA;
  }
  svo_b200_ctx* ctx = new svo_b200_ctx();
  ctx->device = device;
  if (cudaSetDevice(device) != cudaSuccess ||
      cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
    fprintf(stderr, "svo_b200_create: cannot initialise device %d: %s\n", device,
            cudaGetErrorString(cudaGetLastError()));
    delete ctx;
    return SVO_B200_ECUDA;
  }
  cudaEventCreate(&ctx->ev_k0);
  cudaEventCreate(&ctx->ev_k1);
  cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
  cudaDeviceGetAttribute(&ctx->max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
  *ctx_out = ctx;
  return SVO_B200_OK;
}

void svo_b200_destroy(svo_b200_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  sia_batch_free(ctx);
  sia_split_free(ctx);
  if (ctx->d_in.p) cudaFree(ctx->d_in.p);
  if (ctx->d_out.p) cudaFree(ctx->d_out.p);
  if (ctx->d_scratch.p) cudaFree(ctx->d_scratch.p);
  if (ctx->h_in.p) cudaFreeHost(ctx->h_in.p);
  if (ctx->h_out.p) cudaFreeHost(ctx->h_out.p);
  if (ctx->ev_k0) cudaEventDestroy(ctx->ev_k0);
  if (ctx->ev_k1) cudaEventDestroy(ctx->ev_k1);
  cudaStreamDestroy(ctx->stream);
  delete ctx;
}

const char* svo_b200_last_error(const svo_b200_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
void* svo_b200_stream(svo_b200_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
uint64_t svo_b200_launch_count(const svo_b200_ctx* ctx) { return ctx ? ctx->launches : 0; }

int svo_b200_synchronize(svo_b200_ctx* ctx) {
  if (!ctx) return SVO_B200_EINVAL;
  SVO_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  return 0;
}

int svo_b200_frame_create(svo_b200_ctx* ctx, int width, int height, int n_levels,
                          svo_b200_frame** frame_out) {
  if (!ctx || !frame_out || width <= 0 || height <= 0 || n_levels < 1 || n_levels > SVO_B200_MAX_LEVELS)
    return set_err(ctx, SVO_B200_EINVAL, "frame_create: bad size %dx%d levels %d", width, height, n_levels);
  cudaSetDevice(ctx->device);
  svo_b200_frame* fr = new svo_b200_frame();
  fr->width = width;
  fr->height = height;
  fr->n_levels = n_levels;
  size_t off = 0;
  for (int l = 0; l < n_levels; ++l) {
    fr->w[l] = l ? fr->w[l - 1] / 2 : width;
    fr->h[l] = l ? fr->h[l - 1] / 2 : height;
    if (fr->w[l] <= 0 || fr->h[l] <= 0) {
      delete fr;
      return set_err(ctx, SVO_B200_EINVAL, "frame_create: level %d is empty", l);
    }
    fr->off[l] = off;
    off += ((size_t)fr->w[l] * fr->h[l] + 255 + 16) / 256 * 256;  // +16: kernels may read one word past
  }
  fr->bytes = off;
  cudaError_t e = cudaMalloc((void**)&fr->base, fr->bytes);
  if (e != cudaSuccess) {
    delete fr;
    return set_err(ctx, SVO_B200_ENOMEM, "frame_create: cudaMalloc(%zu): %s", off, cudaGetErrorString(e));
  }
  for (int l = 0; l < n_levels; ++l) fr->lv[l] = fr->base + fr->off[l];
  cudaMemsetAsync(fr->base, 0, fr->bytes, ctx->stream);
  *frame_out = fr;
  return 0;
}

int svo_b200_frame_upload(svo_b200_ctx* ctx, svo_b200_frame* fr, const uint8_t* const* levels, int n_given) {
  if (!ctx || !fr || !levels || n_given < 1 || n_given > fr->n_levels)
    return set_err(ctx, SVO_B200_EINVAL, "frame_upload: bad arguments");
  cudaSetDevice(ctx->device);
  for (int l = 0; l < n_given; ++l) {
    if (!levels[l]) return set_err(ctx, SVO_B200_EINVAL, "frame_upload: level %d is NULL", l);
    SVO_CUDA_CHECK(ctx, cudaMemcpyAsync(fr->lvl(l), levels[l], (size_t)fr->w[l] * fr->h[l],
                                        cudaMemcpyHostToDevice, ctx->stream));
  }
  return build_levels(ctx, fr, n_given);
}

int svo_b200_frame_upload_device(svo_b200_ctx* ctx, svo_b200_frame* fr, const void* level0_dev) {
  if (!ctx || !fr || !level0_dev) return set_err(ctx, SVO_B200_EINVAL, "frame_upload_device: bad arguments");
  cudaSetDevice(ctx->device);
  SVO_CUDA_CHECK(ctx, cudaMemcpyAsync(fr->lvl(0), level0_dev, (size_t)fr->w[0] * fr->h[0],
                                      cudaMemcpyDeviceToDevice, ctx->stream));
  return build_levels(ctx, fr, 1);
}

int svo_b200_frame_download_level(svo_b200_ctx* ctx, const svo_b200_frame* fr, int level, uint8_t* out) {
  if (!ctx || !fr || !out || level < 0 || level >= fr->n_levels)
    return set_err(ctx, SVO_B200_EINVAL, "frame_download_level: bad arguments");
  cudaSetDevice(ctx->device);
  SVO_CUDA_CHECK(ctx, cudaMemcpyAsync(out, fr->lvl(level), (size_t)fr->w[level] * fr->h[level],
                                      cudaMemcpyDeviceToHost, ctx->stream));
  SVO_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  return 0;
}

void svo_b200_frame_destroy(svo_b200_ctx* ctx, svo_b200_frame* fr) {
  if (!fr) return;
  if (ctx) {
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
  }
  if (fr->pooled) return;  // borrowed handle of a pool
  if (fr->base) cudaFree(fr->base);
  delete fr;
}


int svo_b200_frame_pool_create(svo_b200_ctx* ctx, int width, int height, int n_levels, int count,
                               svo_b200_frame_pool** pool_out) {
  if (!ctx || !pool_out || width <= 0 || height <= 0 || n_levels < 1 || n_levels > SVO_B200_MAX_LEVELS || count <= 0)
    return set_err(ctx, SVO_B200_EINVAL, "frame_pool_create: bad arguments");
  cudaSetDevice(ctx->device);
  svo_b200_frame proto;
  proto.width = width; proto.height = height; proto.n_levels = n_levels; proto.pooled = true;
  svo_b200_frame_pool* pool = new svo_b200_frame_pool();
  pool->count = count;
  pool->n_levels = n_levels;
  size_t total = 0, slab_off[SVO_B200_MAX_LEVELS];
  for (int l = 0; l < n_levels; ++l) {
    proto.w[l] = l ? proto.w[l - 1] / 2 : width;
    proto.h[l] = l ? proto.h[l - 1] / 2 : height;
    if (proto.w[l] <= 0 || proto.h[l] <= 0) {
      delete pool;
      return set_err(ctx, SVO_B200_EINVAL, "frame_pool_create: level %d is empty", l);
    }
    pool->stride[l] = ((size_t)proto.w[l] * proto.h[l] + 255) / 256 * 256;
    slab_off[l] = total;
    total += pool->stride[l] * (size_t)count + 256;  // slack: kernels fetch aligned words around footprints
  }
  cudaError_t e = cudaMalloc((void**)&pool->mem, total);
  if (e != cudaSuccess) {
    delete pool;
    return set_err(ctx, SVO_B200_ENOMEM, "frame_pool_create: cudaMalloc(%zu): %s", total, cudaGetErrorString(e));
  }
  cudaMemsetAsync(pool->mem, 0, total, ctx->stream);
  for (int l = 0; l < n_levels; ++l) pool->slab[l] = pool->mem + slab_off[l];
  pool->frames.assign(count, proto);
  for (int i = 0; i < count; ++i)
    for (int l = 0; l < n_levels; ++l) pool->frames[i].lv[l] = pool->slab[l] + (size_t)i * pool->stride[l];
  *pool_out = pool;
  return 0;
}

svo_b200_frame* svo_b200_frame_pool_get(svo_b200_frame_pool* pool, int index) {
  if (!pool || index < 0 || index >= pool->count) return nullptr;
  return &pool->frames[index];
}

int svo_b200_frame_pool_upload(svo_b200_ctx* ctx, svo_b200_frame_pool* pool, int first, int count,
                               const uint8_t* level0_host, size_t host_stride_bytes) {
  if (!ctx || !pool || !level0_host || first < 0 || count <= 0 || first + count > pool->count)
    return set_err(ctx, SVO_B200_EINVAL, "frame_pool_upload: bad arguments");
  cudaSetDevice(ctx->device);
  const svo_b200_frame& f0 = pool->frames[0];
  const size_t img = (size_t)f0.w[0] * f0.h[0];
  if (host_stride_bytes < img) return set_err(ctx, SVO_B200_EINVAL, "frame_pool_upload: host stride < image size");
  uint8_t* dst = pool->slab[0] + (size_t)first * pool->stride[0];
  if (host_stride_bytes == img && pool->stride[0] == img) {  // fully contiguous on both sides: one flat copy
    SVO_CUDA_CHECK(ctx, cudaMemcpyAsync(dst, level0_host, img * (size_t)count, cudaMemcpyHostToDevice, ctx->stream));
  } else {  // ONE strided copy: row i = level 0 of frame first+i
    SVO_CUDA_CHECK(ctx, cudaMemcpy2DAsync(dst, pool->stride[0], level0_host, host_stride_bytes, img, (size_t)count,
                                          cudaMemcpyHostToDevice, ctx->stream));
  }
  if (f0.n_levels > 1) {
    // level 0 -> 1 with the streaming kernel when the width allows 128-bit rows, then levels 2.. from
    // level 1 with the fused tile kernel (base level shifted by one); otherwise everything from level 0.
    const bool stream01 = (f0.w[0] % 16) == 0;
    const int base = stream01 ? 1 : 0;
    kt_begin(ctx);
    if (stream01) {
      const int blocks = ctx->sm_count * 8;
      pyramid_l0_l1_stream_kernel<<<blocks, 256, 0, ctx->stream>>>(pool->slab[0], pool->stride[0], pool->slab[1],
                                                                   pool->stride[1], first, count, f0.w[0], f0.h[1],
                                                                   pyr_avg(ctx, f0.w[0]));
      ctx->launches++;
    }
    const int n_sub = f0.n_levels - base;  // levels seen by the fused kernel, its level 0 = our level `base`
    if (n_sub > 1) {
      PyrGeom g;
      memset(&g, 0, sizeof(g));
      g.n_levels = n_sub < 5 ? n_sub : 5;
      for (int l = 0; l < n_sub; ++l) {
        g.w[l] = f0.w[base + l]; g.h[l] = f0.h[base + l]; g.slab[l] = pool->slab[base + l]; g.stride[l] = pool->stride[base + l];
      }
      for (int l = 1; l < g.n_levels; ++l)
        if (pyr_avg(ctx, g.w[l - 1])) g.avg_mask |= 1u << l;
      g.tiles_x = (g.w[0] + 127) / 128;
      const int tiles_y = (g.h[0] + 15) / 16;
      for (int done = 0; done < count; done += 32768) {  // gridDim.y limit 65535
        const int n = count - done < 32768 ? count - done : 32768;
        dim3 grid(g.tiles_x * tiles_y, n);
        pyramid_fused_kernel<<<grid, 128, 0, ctx->stream>>>(first + done, g);
        ctx->launches++;
      }
    }
    SVO_CUDA_CHECK(ctx, cudaGetLastError());
    for (int i = 0; i < count && f0.n_levels > base + 5; ++i) {  // deeper levels: plain per-level kernel
      int rc = build_levels(ctx, &pool->frames[first + i], base + 5, false);
      if (rc) return rc;
    }
    kt_end(ctx);
  }
  return 0;
}

void svo_b200_frame_pool_destroy(svo_b200_ctx* ctx, svo_b200_frame_pool* pool) {
  if (!pool) return;
  if (ctx) {
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
  }
  if (pool->mem) cudaFree(pool->mem);
  delete pool;
}

}  // extern "C"
