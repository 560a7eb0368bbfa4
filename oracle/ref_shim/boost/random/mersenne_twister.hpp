// Stand-in for Boost.Random's mt19937 (Boost is not installed here): the same generator from <random>.
// Test infrastructure only.
#pragma once
#include <random>
namespace boost {
using std::mt19937;
}
