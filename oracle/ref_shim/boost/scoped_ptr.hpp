// Stand-in for <boost/scoped_ptr.hpp> (Boost is not installed here).  Test infrastructure only.
#pragma once
#include <memory>
namespace boost {
template <class T>
class scoped_ptr {
 public:
  explicit scoped_ptr(T* p = 0) : m_p(p) {}
  ~scoped_ptr() { delete m_p; }
  void reset(T* p = 0) { if (p != m_p) { delete m_p; m_p = p; } }
  T& operator*() const { return *m_p; }
  T* operator->() const { return m_p; }
  T* get() const { return m_p; }
  explicit operator bool() const { return m_p != 0; }
  bool operator!() const { return m_p == 0; }
 private:
  scoped_ptr(const scoped_ptr&);
  scoped_ptr& operator=(const scoped_ptr&);
  T* m_p;
};
template <class T> inline bool operator==(const scoped_ptr<T>& a, T* b) { return a.get() == b; }
template <class T> inline bool operator!=(const scoped_ptr<T>& a, T* b) { return a.get() != b; }
}
