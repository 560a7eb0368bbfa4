// Stand-in for <boost/thread/xtime.hpp> (Boost is not installed here).  Test infrastructure only.
#pragma once
#include <chrono>
namespace boost {
struct xtime { long long sec; long nsec; };
enum { TIME_UTC_ = 1 };
inline int xtime_get(xtime* t, int base) {
  const auto now = std::chrono::system_clock::now().time_since_epoch();
  const long long ns = std::chrono::duration_cast<std::chrono::nanoseconds>(now).count();
  t->sec = ns / 1000000000ll;
  t->nsec = (long)(ns % 1000000000ll);
  return base;
}
}
