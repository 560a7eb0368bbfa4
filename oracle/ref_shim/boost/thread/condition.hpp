// Stand-in for <boost/thread/condition.hpp> on <condition_variable> (Boost is not installed here).  Test infrastructure only.
#pragma once
#include <condition_variable>
namespace boost {
typedef std::condition_variable_any condition;
typedef std::condition_variable_any condition_variable_any;
}
