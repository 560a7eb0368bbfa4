// Stand-in for <boost/thread/mutex.hpp> on <mutex> (Boost is not installed here).  Test infrastructure only.
#pragma once
#include <mutex>
namespace boost {
using std::defer_lock;
class mutex : public std::mutex {
 public:
  typedef std::unique_lock<mutex> scoped_lock;
};
}
