// oracle/shim/vikit/math_utils.h -- TEST INFRASTRUCTURE ONLY: [EXT] rpg_vikit helpers restated.
#pragma once
#include <Eigen/Core>
#include <sophus/se3.h>
#include <algorithm>
#include <cmath>
#include <vector>
namespace vk {
using namespace Eigen;
using namespace std;
using namespace Sophus;
inline Vector2d project2d(const Vector3d& v) { return v.head<2>() / v[2]; }
inline Vector3d unproject2d(const Vector2d& v) { return Vector3d(v[0], v[1], 1.0); }
template <class V> inline double norm_max(const V& v) { double m = 0; for (int i = 0; i < (int)v.size(); ++i) m = std::max(m, std::fabs((double)v[i])); return m; }
template <class T> T getMedian(vector<T>& data_vec) {
  typename vector<T>::iterator it = data_vec.begin() + floor(data_vec.size() / 2);
  nth_element(data_vec.begin(), it, data_vec.end());
  return *it;
}
}  // namespace vk
