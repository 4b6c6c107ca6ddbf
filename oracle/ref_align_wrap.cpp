// oracle/ref_align_wrap.cpp -- TEST INFRASTRUCTURE ONLY.  C wrappers around the reference's own
// svo::feature_alignment::align1D / align2D / align2D_SSE2, compiled from /root/reference/svo/src/feature_alignment.cpp
// (see oracle/Makefile target _ref) against the stand-in headers in oracle/shim/.
#include <svo/feature_alignment.h>
extern "C" {
int ref_align2d(const uint8_t* img, int cols, int rows, int step, uint8_t* pwb, uint8_t* patch, int n_iter, double* px_io) {
  cv::Mat m(rows, cols, const_cast<uint8_t*>(img), (size_t)step);
  Eigen::Vector2d px(px_io[0], px_io[1]);
  const bool ok = svo::feature_alignment::align2D(m, pwb, patch, n_iter, px, true);
  px_io[0] = px[0]; px_io[1] = px[1];
  return ok ? 1 : 0;
}
int ref_align1d(const uint8_t* img, int cols, int rows, int step, const float* dir, uint8_t* pwb, uint8_t* patch, int n_iter,
                double* px_io, double* h_inv) {
  cv::Mat m(rows, cols, const_cast<uint8_t*>(img), (size_t)step);
  Eigen::Vector2d px(px_io[0], px_io[1]);
  Eigen::Vector2f d(dir[0], dir[1]);
  const bool ok = svo::feature_alignment::align1D(m, d, pwb, patch, n_iter, px, *h_inv);
  px_io[0] = px[0]; px_io[1] = px[1];
  return ok ? 1 : 0;
}
}
