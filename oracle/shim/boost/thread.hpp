// oracle/shim/boost/thread.hpp -- TEST INFRASTRUCTURE ONLY: std:: stand-ins for the few Boost.Thread names
// svo/include/svo/depth_filter.h uses (the wrapper never starts the filter thread).
#pragma once
#include <condition_variable>
#include <mutex>
#include <thread>
namespace boost {
using mutex = std::mutex;
template <class M> using unique_lock = std::unique_lock<M>;
using condition_variable = std::condition_variable;
class thread {
 public:
  template <class F, class... A> explicit thread(F&& f, A&&... a) : t_(std::forward<F>(f), std::forward<A>(a)...) {}
  void interrupt() {}
  void join() { if (t_.joinable()) t_.join(); }
 private:
  std::thread t_;
};
namespace this_thread { inline bool interruption_requested() { return true; } }
}  // namespace boost
