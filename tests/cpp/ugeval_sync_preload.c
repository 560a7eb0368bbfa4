/* LD_PRELOAD shim in front of the REAL libugeval.so for oracle/_ref/ref_search on a GPU: ug_eval_run_planes waits until
 * the caller's buffer holds new content that stayed unchanged for UGSTUB_SETTLE_US (250) microseconds, then calls the
 * product's own ug_eval_run_planes.  Needed because the reference's NetworkEvalThread evaluates in a free-running loop
 * and delivers whatever it computed to a leaf whose planes were written later (tests/cpp/ugeval_stub.c has the long
 * version); with the wait, the reference's search on the CUDA evaluator is reproducible and can be compared with the
 * product's self-play driver on the same evaluator.  Test infrastructure only. */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <stdint.h>
#include <stdlib.h>
#include <unistd.h>

typedef struct ug_eval ug_eval;
typedef int (*run_planes_fn)(ug_eval*, const int8_t*, int, double*, double*);

static uint64_t hash_planes(const volatile int8_t* p, size_t n) {
  uint64_t h = 1469598103934665603ull;
  for (size_t i = 0; i < n; ++i) { h ^= (uint8_t)p[i]; h *= 1099511628211ull; }
  return h;
}

int ug_eval_run_planes(ug_eval* e, const int8_t* planes, int n, double* policy, double* value) {
  static run_planes_fn real = 0;
  static uint64_t last = 0;
  if (!real) real = (run_planes_fn)dlsym(RTLD_NEXT, "ug_eval_run_planes");
  if (!real) abort();
  const char* w = getenv("UGSTUB_SETTLE_US");
  const useconds_t settle = w ? (useconds_t)atoi(w) : 250;
  const size_t bytes = (size_t)n * 17 * 361;
  for (int spin = 0; spin < 20000 / (int)(2 * settle + 1) + 1; ++spin) {
    usleep(settle);
    const uint64_t h1 = hash_planes(planes, bytes);
    usleep(settle);
    const uint64_t h2 = hash_planes(planes, bytes);
    if (h1 == h2 && h1 != last) { last = h1; break; }
  }
  return real(e, planes, n, policy, value);
}
