// Stand-in for boost::variate_generator<Engine&, Distribution>.  Test infrastructure only.
#pragma once
#include <type_traits>
namespace boost {
template <class EngineRef, class Dist>
class variate_generator {
 public:
  typedef typename std::remove_reference<EngineRef>::type Engine;
  variate_generator(Engine& e, Dist d) : m_engine(e), m_dist(d) {}
  typename Dist::result_type operator()() { return m_dist(m_engine); }
  Dist& distribution() { return m_dist; }
 private:
  Engine& m_engine;
  Dist m_dist;
};
}
