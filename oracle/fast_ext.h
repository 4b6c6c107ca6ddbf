// oracle/fast_ext.h -- TEST INFRASTRUCTURE ONLY.
// [EXT] restatement of the un-vendored `fast` corner detector library (uzh-rpg/fast, derived from E. Rosten's FAST /
// libCVD) and of vk::shiTomasiScore (rpg_vikit vision.cpp), from their published algorithms.  PARITY UNPINNED: neither
// library is in /root/reference; used by the oracle (svo_oracle_detect.inc) and as the link-time definition behind the
// reference's own feature_detection.cpp in oracle/_ref (ref_wrap.cpp), so only FastDetector::detect itself is pinned.
//   fast_corner_detect_10 : segment test on the 16-pixel Bresenham circle of radius 3, >= 10 contiguous pixels all
//                           brighter than centre+b or all darker than centre-b; raster order, 3-pixel border.
//   fast_corner_score_10  : largest b for which the pixel is still a corner (the library bisects b in [b0, 255]).
//   fast_nonmax_3x3       : a corner survives unless one of its 8 neighbours is a detected corner with a larger score
//                           (libCVD's non-strict nonmax_suppression: "a point must be at least as large as its
//                           neighbours", equal neighbours all survive).  The strict variant (a neighbour with an equal
//                           score suppresses too) is selectable: kFastTiesSuppress / the ties_suppress argument.
#pragma once
#include <cmath>
#include <cstdint>
#include <vector>

namespace fast_ext {

constexpr bool kFastTiesSuppress = false;
constexpr int kRing[16][2] = {{0, -3}, {1, -3}, {2, -2}, {3, -1}, {3, 0}, {3, 1}, {2, 2}, {1, 3},
                              {0, 3}, {-1, 3}, {-2, 2}, {-3, 1}, {-3, 0}, {-3, -1}, {-2, -2}, {-1, -3}};

// M = max over the 16 arcs of 10 contiguous ring pixels of min over the arc of d_k; the pixel is a corner at
// threshold b iff M > b (for d_k = I_k - c: brighter arcs; for d_k = c - I_k: darker arcs).
inline int arc_max_min(const int d[16]) {
  int best = -256;
  for (int s = 0; s < 16; ++s) {
    int m = 255;
    for (int j = 0; j < 10; ++j) m = d[(s + j) & 15] < m ? d[(s + j) & 15] : m;
    best = m > best ? m : best;
  }
  return best;
}
inline int corner_strength(const uint8_t* img, int stride, int x, int y) {  // M of the stronger polarity
  const int c = img[y * stride + x];
  int db[16], dd[16];
  for (int k = 0; k < 16; ++k) {
    const int v = img[(y + kRing[k][1]) * stride + x + kRing[k][0]];
    db[k] = v - c;
    dd[k] = c - v;
  }
  const int a = arc_max_min(db), b = arc_max_min(dd);
  return a > b ? a : b;
}

struct xy { short x, y; };

inline void detect10(const uint8_t* img, int w, int h, int stride, int b, std::vector<xy>& corners) {
  corners.clear();
  for (int y = 3; y < h - 3; ++y)
    for (int x = 3; x < w - 3; ++x)
      if (corner_strength(img, stride, x, y) > b) corners.push_back(xy{(short)x, (short)y});
}
inline void score10(const uint8_t* img, int stride, const std::vector<xy>& corners, int b0, std::vector<int>& scores) {
  scores.resize(corners.size());
  for (size_t i = 0; i < corners.size(); ++i) {
    // bisection of the library: bmin = b0, bmax = 255; returns the largest b in [b0, 254] that is still a corner
    const int M = corner_strength(img, stride, corners[i].x, corners[i].y);
    int s = M - 1;
    if (s < b0) s = b0;
    if (s > 254) s = 254;
    scores[i] = s;
  }
}
inline void nonmax3x3(const std::vector<xy>& corners, const std::vector<int>& scores, std::vector<int>& nonmax,
                      bool ties_suppress = kFastTiesSuppress) {
  nonmax.clear();
  const int n = (int)corners.size();
  if (n < 1) return;
  const int last_row = corners.back().y;
  std::vector<int> row_start(last_row + 2, -1);
  for (int i = n - 1; i >= 0; --i) row_start[corners[i].y] = i;  // first index of each row (raster order)
  auto beats = [&](int j, int score) { return ties_suppress ? scores[j] >= score : scores[j] > score; };
  for (int i = 0; i < n; ++i) {
    const int score = scores[i];
    const xy p = corners[i];
    bool bad = false;
    if (i > 0 && corners[i - 1].y == p.y && corners[i - 1].x == p.x - 1 && beats(i - 1, score)) bad = true;
    if (!bad && i < n - 1 && corners[i + 1].y == p.y && corners[i + 1].x == p.x + 1 && beats(i + 1, score)) bad = true;
    for (int dy = -1; dy <= 1 && !bad; dy += 2) {
      const int r = p.y + dy;
      if (r < 0 || r > last_row || row_start[r] < 0) continue;
      for (int j = row_start[r]; j < n && corners[j].y == r && corners[j].x <= p.x + 1; ++j)
        if (corners[j].x >= p.x - 1 && beats(j, score)) { bad = true; break; }
    }
    if (!bad) nonmax.push_back(i);
  }
}

// [EXT] vk::shiTomasiScore (rpg_vikit vision.cpp): smaller eigenvalue of the 8x8 central-difference structure tensor.
// (contraction is pinned off so that the oracle build and the oracle/_ref build of this header agree bit for bit)
__attribute__((optimize("fp-contract=off"))) inline float shiTomasiScore(const uint8_t* img, int cols, int rows, int stride, int u,
                                                                         int v) {
  float dXX = 0.0, dYY = 0.0, dXY = 0.0;
  const int halfbox_size = 4;
  const int box_size = 2 * halfbox_size;
  const int box_area = box_size * box_size;
  const int x_min = u - halfbox_size, x_max = u + halfbox_size, y_min = v - halfbox_size, y_max = v + halfbox_size;
  if (x_min < 1 || x_max >= cols - 1 || y_min < 1 || y_max >= rows - 1) return 0.0;  // patch is too close to the boundary
  for (int y = y_min; y < y_max; ++y) {
    const uint8_t* ptr_left = img + stride * y + x_min - 1;
    const uint8_t* ptr_right = img + stride * y + x_min + 1;
    const uint8_t* ptr_top = img + stride * (y - 1) + x_min;
    const uint8_t* ptr_bottom = img + stride * (y + 1) + x_min;
    for (int x = 0; x < box_size; ++x, ++ptr_left, ++ptr_right, ++ptr_top, ++ptr_bottom) {
      const float dx = float(*ptr_right - *ptr_left);
      const float dy = float(*ptr_bottom - *ptr_top);
      dXX += dx * dx;  // integers < 2^24: exact in float in any order
      dYY += dy * dy;
      dXY += dx * dy;
    }
  }
  // the literals are double: the division and the eigenvalue formula run in double, the result narrows to float
  const float fXX = float(dXX / (2.0 * box_area)), fYY = float(dYY / (2.0 * box_area)), fXY = float(dXY / (2.0 * box_area));
  return float(0.5 * (double(fXX + fYY) - std::sqrt(double((fXX + fYY) * (fXX + fYY) - 4 * (fXX * fYY - fXY * fXY)))));
}

}  // namespace fast_ext
