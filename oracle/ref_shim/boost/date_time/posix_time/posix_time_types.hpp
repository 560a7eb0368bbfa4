// Stand-in for the sliver of Boost.DateTime the reference's SgRandom.cpp uses to derive a time-based seed
// (microseconds since 2000-01-01), on <chrono>.  Test infrastructure only.
#pragma once
#include <chrono>
#include <cstdint>
namespace boost {
namespace gregorian {
enum months_of_year { Jan = 1 };
struct date {
  date(int y, int m, int d) : year(y), month(m), day(d) {}
  int year, month, day;
};
}  // namespace gregorian
namespace posix_time {
struct time_duration {
  int64_t us;
  int64_t total_microseconds() const { return us; }
  int64_t total_nanoseconds() const { return us * 1000; }
  int64_t total_milliseconds() const { return us / 1000; }
  int64_t total_seconds() const { return us / 1000000; }
};
struct ptime {
  int64_t us_since_epoch;
  ptime() : us_since_epoch(0) {}
  explicit ptime(const gregorian::date& d) : us_since_epoch(0) {
    // days from 1970-01-01 to the civil date (proleptic Gregorian)
    int y = d.year, m = d.month;
    y -= m <= 2;
    const int era = (y >= 0 ? y : y - 399) / 400;
    const unsigned yoe = (unsigned)(y - era * 400);
    const unsigned doy = (153u * (unsigned)(m + (m > 2 ? -3 : 9)) + 2u) / 5u + (unsigned)d.day - 1u;
    const unsigned doe = yoe * 365u + yoe / 4u - yoe / 100u + doy;
    const int64_t days = (int64_t)era * 146097 + (int64_t)doe - 719468;
    us_since_epoch = days * 86400ll * 1000000ll;
  }
};
inline time_duration operator-(const ptime& a, const ptime& b) {
  return time_duration{a.us_since_epoch - b.us_since_epoch};
}
struct microsec_clock {
  static ptime universal_time() {
    ptime t;
    t.us_since_epoch = std::chrono::duration_cast<std::chrono::microseconds>(
                           std::chrono::system_clock::now().time_since_epoch()).count();
    return t;
  }
};
}  // namespace posix_time
}  // namespace boost
