// Stand-in for <boost/thread/recursive_mutex.hpp> on <mutex> (Boost is not installed here).  Test infrastructure only.
#pragma once
#include <mutex>
namespace boost {
class recursive_mutex : public std::recursive_mutex {
 public:
  typedef std::unique_lock<recursive_mutex> scoped_lock;
};
}
