// Stand-in for <boost/thread.hpp> (Boost is not installed here).  Test infrastructure only.
#pragma once
#include "thread/barrier.hpp"
#include "thread/condition.hpp"
#include "thread/mutex.hpp"
#include "thread/recursive_mutex.hpp"
#include "thread/thread.hpp"
