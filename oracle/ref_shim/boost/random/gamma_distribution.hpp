// Stand-in for boost::gamma_distribution<>.  Test infrastructure only.
#pragma once
#include <random>
namespace boost {
template <class Real = double>
using gamma_distribution = std::gamma_distribution<Real>;
}
