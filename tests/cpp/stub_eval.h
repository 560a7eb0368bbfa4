// A deterministic stand-in "network" for CPU tests: policy and value are a hash of the 17 feature planes, produced as
// floats.  Shared by the stub evaluator the reference's search engine is run against (ugeval_stub.c, preloaded under
// oracle/_ref/ref_search) and by the stub leaf queue of the product's self-play driver (selfplay_stub.cpp), so that
// both searches see exactly the same numbers for the same position.
#ifndef UG_TEST_STUB_EVAL_H
#define UG_TEST_STUB_EVAL_H
#include <stdint.h>
#include <stdlib.h>

static inline void stub_eval(const int8_t* planes /* [17][361] */, float* policy /* [362] */, float* value) {
  uint64_t h = 1469598103934665603ull;
  for (int i = 0; i < 17 * 361; ++i) { h ^= (uint8_t)planes[i]; h *= 1099511628211ull; }
  float sum = 0.f;
  for (int i = 0; i < 362; ++i) {  // a peaked distribution, like a trained policy head
    h ^= h >> 29; h *= 0xBF58476D1CE4E5B9ull;
    const float u = (float)((h >> 40) & 0xFFFF) / 65536.f;
    policy[i] = u * u * u * u + 1e-4f;
    sum += policy[i];
  }
  for (int i = 0; i < 362; ++i) policy[i] /= sum;
  *value = 0.05f * (float)((int)(h % 21) - 10);
}

/* value_transform (funcapproximator/DlTFNetworkEvaluator.cc:300-310): set through ug_eval_set_value_transform like on
 * the product; a driver that has no evaluator object (selfplay_stub.cpp) passes it in UGSTUB_VT. */
static int g_stub_vt = -1;
static inline int stub_vt(void) {
  if (g_stub_vt < 0) { const char* e = getenv("UGSTUB_VT"); g_stub_vt = e ? atoi(e) : 0; }
  return g_stub_vt;
}
static inline float stub_transform(float v) { return stub_vt() == 1 ? -v : (stub_vt() == 2 ? 1.f - v : v); }
#endif
