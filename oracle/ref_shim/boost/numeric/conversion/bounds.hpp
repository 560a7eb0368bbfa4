// Stand-in for <boost/numeric/conversion/bounds.hpp> (Boost is not installed here).  Test infrastructure only.
#pragma once
#include <limits>
namespace boost { namespace numeric {
template <class T> struct bounds {
  static T lowest() { return std::numeric_limits<T>::lowest(); }
  static T highest() { return std::numeric_limits<T>::max(); }
  static T smallest() { return std::numeric_limits<T>::min(); }
};
} }
