// oracle/shim/vikit/timer.h -- TEST INFRASTRUCTURE ONLY: vk::Timer stand-in (wall-clock stopwatch).
#pragma once
#include <chrono>
namespace vk {
class Timer {
 public:
  Timer() { start(); }
  void start() { t0_ = std::chrono::steady_clock::now(); }
  double stop() { dt_ = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0_).count(); return dt_; }
  double getTime() const { return dt_; }
  void reset() { dt_ = 0; }
 private:
  std::chrono::steady_clock::time_point t0_;
  double dt_ = 0;
};
}  // namespace vk
