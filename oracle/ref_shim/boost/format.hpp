// Stand-in for <boost/format.hpp> (Boost is not installed here): enough of boost::format for the reference's
// exception messages — "%N%"/"%s"-style directives are substituted in order.  Test infrastructure only.
#pragma once
#include <sstream>
#include "io/ios_state.hpp"  // real Boost.Format pulls it in, and search/UctSearch.cpp relies on that
#include <string>
#include <vector>
namespace boost {
class format {
 public:
  explicit format(const std::string& s) : m_fmt(s) {}
  template <class T>
  format& operator%(const T& v) {
    std::ostringstream o;
    o << v;
    m_args.push_back(o.str());
    return *this;
  }
  std::string str() const {
    std::string out;
    size_t next = 0;
    for (size_t i = 0; i < m_fmt.size(); ++i) {
      if (m_fmt[i] != '%') { out += m_fmt[i]; continue; }
      size_t j = i + 1;
      if (j < m_fmt.size() && m_fmt[j] == '%') { out += '%'; i = j; continue; }
      while (j < m_fmt.size() && (m_fmt[j] >= '0' && m_fmt[j] <= '9')) ++j;
      if (j < m_fmt.size() && m_fmt[j] == '%') {           // %N%
        const size_t n = (size_t)std::stoi(m_fmt.substr(i + 1, j - i - 1));
        if (n >= 1 && n <= m_args.size()) out += m_args[n - 1];
        i = j;
      } else {                                             // %s %d ...
        if (next < m_args.size()) out += m_args[next++];
        i = j < m_fmt.size() ? j : m_fmt.size() - 1;
      }
    }
    return out;
  }
 private:
  std::string m_fmt;
  std::vector<std::string> m_args;
};
inline std::ostream& operator<<(std::ostream& o, const format& f) { return o << f.str(); }
inline std::string str(const format& f) { return f.str(); }
}  // namespace boost
