// Stand-in for <boost/scoped_array.hpp> (Boost is not installed here).  Test infrastructure only.
#pragma once
#include <cstddef>
namespace boost {
template <class T>
class scoped_array {
 public:
  explicit scoped_array(T* p = 0) : m_p(p) {}
  ~scoped_array() { delete[] m_p; }
  void reset(T* p = 0) { if (p != m_p) { delete[] m_p; m_p = p; } }
  T& operator[](std::ptrdiff_t i) const { return m_p[i]; }
  T* get() const { return m_p; }
 private:
  scoped_array(const scoped_array&);
  scoped_array& operator=(const scoped_array&);
  T* m_p;
};
}
