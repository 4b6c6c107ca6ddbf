// oracle/shim/opencv2/opencv.hpp -- TEST INFRASTRUCTURE ONLY: the three cv::Mat members the hot path reads.
#pragma once
#include <cstddef>
#include <cstdint>
namespace cv {
struct MatStep { size_t p[2]; };
struct Mat {
  unsigned char* data = nullptr;
  int rows = 0, cols = 0;
  MatStep step{{0, 1}};
  Mat() {}
  Mat(int r, int c, unsigned char* d, size_t pitch) : data(d), rows(r), cols(c) { step.p[0] = pitch; step.p[1] = 1; }
};
}  // namespace cv
