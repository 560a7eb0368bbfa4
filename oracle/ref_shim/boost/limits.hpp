// Stand-in for <boost/limits.hpp> (Boost is not installed here).  Test infrastructure only.
#pragma once
#include <limits>
