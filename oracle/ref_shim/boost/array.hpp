// Stand-in for <boost/array.hpp> (Boost is not installed here).  Test infrastructure only.
#pragma once
#include <array>
namespace boost { using std::array; }
