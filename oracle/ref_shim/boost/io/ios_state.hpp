// Stand-in for <boost/io/ios_state.hpp>: savers restore a stream's formatting state on scope exit
// (Boost is not installed here).  Test infrastructure only.
#pragma once
#include <ios>
namespace boost { namespace io {
class ios_flags_saver {
 public:
  explicit ios_flags_saver(std::ios_base& s) : m_s(s), m_f(s.flags()) {}
  ~ios_flags_saver() { m_s.flags(m_f); }
 private:
  std::ios_base& m_s;
  std::ios_base::fmtflags m_f;
};
class ios_all_saver {
 public:
  explicit ios_all_saver(std::ios& s) : m_s(s), m_f(s.flags()), m_p(s.precision()), m_w(s.width()), m_fill(s.fill()) {}
  ~ios_all_saver() { m_s.flags(m_f); m_s.precision(m_p); m_s.width(m_w); m_s.fill(m_fill); }
 private:
  std::ios& m_s;
  std::ios_base::fmtflags m_f;
  std::streamsize m_p, m_w;
  char m_fill;
};
} }
