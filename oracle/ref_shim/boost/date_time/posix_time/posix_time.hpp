// Stand-in for <boost/date_time/posix_time/posix_time.hpp> (Boost is not installed here).  Test infrastructure only.
#pragma once
#include "posix_time_types.hpp"
