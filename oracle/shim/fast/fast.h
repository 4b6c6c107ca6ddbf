// oracle/shim/fast/fast.h -- TEST INFRASTRUCTURE ONLY: declarations of the [EXT] `fast` corner detector library
// (uzh-rpg/fast) so that svo/src/feature_detection.cpp compiles and links; corner detection itself is outside the
// hot path proper; the definitions live in oracle/ref_wrap.cpp (restated in oracle/fast_ext.h).
#pragma once
#include <vector>
namespace fast {
typedef unsigned char fast_byte;
struct fast_xy { short x, y; fast_xy(short x_, short y_) : x(x_), y(y_) {} };
void fast_corner_detect_10(const fast_byte* img, int w, int h, int stride, short thresh, std::vector<fast_xy>& corners);
void fast_corner_detect_10_sse2(const fast_byte* img, int w, int h, int stride, short thresh, std::vector<fast_xy>& corners);
void fast_corner_detect_10_neon(const fast_byte* img, int w, int h, int stride, short thresh, std::vector<fast_xy>& corners);
void fast_corner_score_10(const fast_byte* img, int stride, const std::vector<fast_xy>& corners, int thresh, std::vector<int>& scores);
void fast_nonmax_3x3(const std::vector<fast_xy>& corners, const std::vector<int>& scores, std::vector<int>& nonmax);
}  // namespace fast
