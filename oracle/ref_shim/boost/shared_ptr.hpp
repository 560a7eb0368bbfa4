// Stand-in for <boost/shared_ptr.hpp> (Boost is not installed here).  Test infrastructure only.
#pragma once
#include <memory>
namespace boost {
using std::shared_ptr;
using std::make_shared;
}
