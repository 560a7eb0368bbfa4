// Stand-in for <boost/lexical_cast.hpp> via a stringstream (Boost is not installed here).  Test infrastructure only.
#pragma once
#include <sstream>
#include <stdexcept>
#include <string>
namespace boost {
struct bad_lexical_cast : std::runtime_error { bad_lexical_cast() : std::runtime_error("bad lexical cast") {} };
template <class To, class From>
inline To lexical_cast(const From& v) {
  std::stringstream s;
  To out;
  if (!(s << v) || !(s >> out)) throw bad_lexical_cast();
  return out;
}
}
