// oracle/shim/opencv2/opencv.hpp -- TEST INFRASTRUCTURE ONLY: the cv::Mat members the hot path touches.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <memory>
#include <string>
#define CV_8U 0
#define CV_8UC1 0
#include <cstdlib>
#define CV_32F 5
#define CV_WINDOW_AUTOSIZE 1
namespace cv {
struct MatStep { size_t p[2]; };
struct Size { int width, height; };
struct Scalar { double v; Scalar(double x = 0) : v(x) {} };
struct Mat {
  unsigned char* data = nullptr;
  int rows = 0, cols = 0, type_ = CV_8U;
  MatStep step{{0, 1}};
  std::shared_ptr<unsigned char> own;
  Mat() {}
  Mat(int r, int c, unsigned char* d, size_t pitch) : data(d), rows(r), cols(c) { step.p[0] = pitch; step.p[1] = 1; }
  Mat(int r, int c, int type) { create(r, c, type); }
  Mat(Size s, int type, Scalar = Scalar()) { create(s.height, s.width, type); }
  void create(int r, int c, int type) {
    rows = r; cols = c; type_ = type;
    const size_t es = (type == CV_32F) ? 4 : 1;
    step.p[0] = es * c; step.p[1] = es;
    // 64-byte aligned like OpenCV's fastMalloc (vk::halfSample's SSE2 branch tests 16-byte alignment)
    const size_t bytes = (es * (size_t)r * c + 64 + 63) & ~size_t(63);
    own.reset(static_cast<unsigned char*>(std::aligned_alloc(64, bytes)), [](unsigned char* q) { std::free(q); });
    data = own.get();
    std::memset(data, 0, es * (size_t)r * c);
  }
  Size size() const { return Size{cols, rows}; }
  int type() const { return type_; }
  bool empty() const { return data == nullptr; }
  template <class T> T& at(int r, int c) { return *reinterpret_cast<T*>(data + r * step.p[0] + c * sizeof(T)); }
  Mat operator*(double) const { return *this; }
  Mat clone() const { Mat m(rows, cols, type_); std::memcpy(m.data, data, step.p[0] * rows); return m; }
};
inline void namedWindow(const std::string&, int) {}
inline void imshow(const std::string&, const Mat&) {}
inline int waitKey(int) { return 0; }
}  // namespace cv
