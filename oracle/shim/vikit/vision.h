// oracle/shim/vikit/vision.h -- TEST INFRASTRUCTURE ONLY: [EXT] vk::interpolateMat_8u, vk::halfSample (scalar path).
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
inline void halfSample(const cv::Mat& in, cv::Mat& out) {
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
