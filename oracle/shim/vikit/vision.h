// oracle/shim/vikit/vision.h -- TEST INFRASTRUCTURE ONLY: [EXT] vk::interpolateMat_8u, vk::halfSample (SSE2 + scalar paths).
#pragma once
#include <opencv2/opencv.hpp>
#include <cmath>
namespace vk {
inline float interpolateMat_8u(const cv::Mat& mat, float u, float v) {
  int x = floor(u);
  int y = floor(v);
  float subpix_x = u - x;
  float subpix_y = v - y;
  float w00 = (1.0f - subpix_x) * (1.0f - subpix_y);
  float w01 = (1.0f - subpix_x) * subpix_y;
  float w10 = subpix_x * (1.0f - subpix_y);
  float w11 = 1.0f - w00 - w01 - w10;
  const int stride = mat.step.p[0];
  unsigned char* ptr = mat.data + y * stride + x;
  return w00 * ptr[0] + w01 * ptr[stride] + w10 * ptr[1] + w11 * ptr[stride + 1];
}
// [EXT] vk::halfSample (rpg_vikit vision.cpp).  On x86 builds (the reference's platform) the SSE2 branch runs whenever
// both buffers are 16-byte aligned and in.cols % 16 == 0 -- true for every BASELINE width at the fine levels: it averages
// VERTICALLY with _mm_avg_epu8 (round half up), then adjacent columns with _mm_avg_epu16 (round half up again); the scalar
// branch is (a+b+c+d)/4 with integer division.  The two differ (e.g. 0,0,0,1 -> 1 vs 0).
#ifdef __SSE2__
#include <emmintrin.h>
inline void halfSampleSSE2(const unsigned char* in, unsigned char* out, int w, int h) {
  const unsigned long long mask[2] = {0x00FF00FF00FF00FFull, 0x00FF00FF00FF00FFull};
  const unsigned char* nextRow = in + w;
  __m128i m = _mm_loadu_si128((const __m128i*)mask);
  int sw = w >> 4;
  int sh = h >> 1;
  for (int i = 0; i < sh; i++) {
    for (int j = 0; j < sw; j++) {
      __m128i here = _mm_load_si128((const __m128i*)in);
      __m128i next = _mm_load_si128((const __m128i*)nextRow);
      here = _mm_avg_epu8(here, next);
      next = _mm_and_si128(_mm_srli_si128(here, 1), m);
      here = _mm_and_si128(here, m);
      here = _mm_avg_epu16(here, next);
      _mm_storel_epi64((__m128i*)out, _mm_packus_epi16(here, here));
      in += 16;
      nextRow += 16;
      out += 8;
    }
    in += w;
    nextRow += w;
  }
}
#endif
inline void halfSample(const cv::Mat& in, cv::Mat& out) {
#ifdef __SSE2__
  if ((reinterpret_cast<size_t>(in.data) & 15) == 0 && (reinterpret_cast<size_t>(out.data) & 15) == 0 && (in.cols % 16) == 0) {
    halfSampleSSE2(in.data, out.data, in.cols, in.rows);
    return;
  }
#endif
  const int stride = in.step.p[0];
  for (int y = 0; y < out.rows; ++y)
    for (int x = 0; x < out.cols; ++x) {
      const unsigned char* top = in.data + 2 * y * stride + 2 * x;
      out.data[y * out.step.p[0] + x] = static_cast<uint8_t>((uint16_t(top[0]) + top[1] + top[stride] + top[stride + 1]) / 4);
    }
}
// [EXT] vk::shiTomasiScore: defined in oracle/ref_wrap.cpp from the restatement in oracle/fast_ext.h.
float shiTomasiScore(const cv::Mat& img, int u, int v);
}  // namespace vk
