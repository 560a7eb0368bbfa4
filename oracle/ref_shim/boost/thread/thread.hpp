// Stand-in for <boost/thread/thread.hpp> on <thread> (Boost is not installed here).  Test infrastructure only.
#pragma once
#include <chrono>
#include <thread>
#include "xtime.hpp"
namespace boost {
class thread {
 public:
  thread() {}
  template <class F>
  explicit thread(F f) : m_t(f) {}
  ~thread() { if (m_t.joinable()) m_t.detach(); }  // boost::thread detaches on destruction
  void join() { if (m_t.joinable()) m_t.join(); }
  bool joinable() const { return m_t.joinable(); }
  static unsigned hardware_concurrency() { return std::thread::hardware_concurrency(); }
  static void yield() { std::this_thread::yield(); }
  static void sleep(const xtime& until) {
    const auto tp = std::chrono::system_clock::time_point(
        std::chrono::duration_cast<std::chrono::system_clock::duration>(std::chrono::seconds(until.sec) +
                                                                        std::chrono::nanoseconds(until.nsec)));
    std::this_thread::sleep_until(tp);
  }
 private:
  thread(const thread&);
  thread& operator=(const thread&);
  std::thread m_t;
};
}
