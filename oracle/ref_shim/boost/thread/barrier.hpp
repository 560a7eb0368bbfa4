// Stand-in for <boost/thread/barrier.hpp>: a cyclic barrier; wait() returns true for one thread per cycle
// (Boost is not installed here).  Test infrastructure only.
#pragma once
#include <condition_variable>
#include <mutex>
namespace boost {
class barrier {
 public:
  explicit barrier(unsigned int count) : m_threshold(count), m_count(count), m_generation(0) {}
  bool wait() {
    std::unique_lock<std::mutex> lock(m_mutex);
    const unsigned int gen = m_generation;
    if (--m_count == 0) {
      ++m_generation;
      m_count = m_threshold;
      m_cond.notify_all();
      return true;
    }
    while (gen == m_generation) m_cond.wait(lock);
    return false;
  }
 private:
  std::mutex m_mutex;
  std::condition_variable m_cond;
  unsigned int m_threshold, m_count, m_generation;
};
}
