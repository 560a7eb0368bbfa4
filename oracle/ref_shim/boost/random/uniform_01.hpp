// Stand-in for boost::uniform_01<Engine, Real> (engine held by value, operator() draws in [0,1)).
// Test infrastructure only.
#pragma once
#include <random>
namespace boost {
template <class Engine, class Real = double>
class uniform_01 {
 public:
  explicit uniform_01(Engine e) : m_engine(e) {}
  Real operator()() {
    const Real r = (Real)(m_engine() - Engine::min()) / ((Real)(Engine::max() - Engine::min()) + (Real)1);
    return r < (Real)1 ? r : (Real)0;
  }
  Engine& base() { return m_engine; }
 private:
  Engine m_engine;
};
}
