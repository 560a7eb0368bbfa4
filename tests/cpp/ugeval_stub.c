/* CPU stand-in for the nine ug_eval_* entry points the reference's UctBoardEvaluator binds (see
 * tests/test_reference_dropin.py), preloaded in front of libugeval.so when oracle/_ref/ref_search is run without a
 * GPU: LD_PRELOAD=ugeval_stub.so oracle/_ref/ref_search ...  Test infrastructure only; the product has no CPU path. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "../../include/ugeval.h"
#include "stub_eval.h"

struct ug_eval { int unused; };
ug_eval* ug_eval_create(int device, int precision, int max_batch) { return (ug_eval*)calloc(1, sizeof(ug_eval)); }
void ug_eval_destroy(ug_eval* e) { free(e); }
int ug_eval_load_graph(ug_eval* e, const char* path) { return 0; }
int ug_eval_load_checkpoint(ug_eval* e, const char* prefix) { return 0; }
int ug_eval_ready(const ug_eval* e) { return 1; }
int ug_eval_set_io(ug_eval* e, const char* in, const char* const* outs, int n) { return 0; }
/* value_transform: state and helpers in stub_eval.h (shared with the stub leaf queue) */
int ug_eval_set_value_transform(ug_eval* e, int vt) { g_stub_vt = (vt == 1 || vt == 2) ? vt : 0; return 0; }
const char* ug_last_error(const ug_eval* e) { return "stub evaluator"; }
static uint64_t hash_planes(const volatile int8_t* p, size_t n) {
  uint64_t h = 1469598103934665603ull;
  for (size_t i = 0; i < n; ++i) { h ^= (uint8_t)p[i]; h *= 1099511628211ull; }
  return h;
}
int ug_eval_run_planes(ug_eval* e, const int8_t* planes, int n, double* policy, double* value) {
  const size_t bytes = (size_t)n * 17 * 361;
  if (getenv("UGSTUB_SYNC")) {
    /* The reference's NetworkEvalThread calls EvaluateState in a free-running loop and hands the result to a search
     * thread whose state_ready flag it finds set AFTERWARDS (search/UctSearch.cpp:186-199), so an evaluation that
     * began before the leaf's features were written is delivered to that leaf.  To observe the search the code
     * intends, this mode does not start until the buffer holds new content that stayed unchanged for UGSTUB_SETTLE_US (200)
     * microseconds (or 20 ms pass: end of search, or a transposition with identical history).  A heuristic — the
     * reference offers no ready-before-evaluate signal to hook — so a search thread descheduled in the middle of
     * writing its planes for longer than that still gets a wrong answer; the test retries. */
    static uint64_t last = 0;
    const char* w = getenv("UGSTUB_SETTLE_US");
    const useconds_t settle = w ? (useconds_t)atoi(w) : 200;
    for (int spin = 0; spin < 20000 / (int)(2 * settle + 1) + 1; ++spin) {
      usleep(settle);
      const uint64_t h1 = hash_planes(planes, bytes);
      usleep(settle);
      const uint64_t h2 = hash_planes(planes, bytes);
      if (h1 == h2 && h1 != last) { last = h1; break; }
    }
  }
  /* like a real evaluator: take the inputs first (TransformFeature / the H2D copy), then spend the evaluation time */
  int8_t* snap = (int8_t*)malloc(bytes);
  memcpy(snap, planes, bytes);
  const char* d = getenv("UGSTUB_DELAY_US");
  if (d) usleep((useconds_t)atoi(d));
  for (int b = 0; b < n; ++b) {
    float p[362], v;
    stub_eval(snap + (size_t)b * 17 * 361, p, &v);
    for (int i = 0; i < 362; ++i) policy[(size_t)b * 362 + i] = (double)p[i];  /* float -> double, as ug_eval_run_planes */
    value[b] = stub_vt() == 1 ? -(double)v : (stub_vt() == 2 ? 1 - (double)v : (double)v);  /* in double, as .cc:300-310 */
  }
  if (getenv("UGSTUB_STATS")) {  /* how often the answer was computed from something other than what the buffer holds now */
    static uint64_t calls = 0, stale = 0;
    ++calls;
    if (hash_planes(planes, bytes) != hash_planes(snap, bytes)) ++stale;
    if (calls % 1000 == 0)
      fprintf(stderr, "ugstub: %llu calls, %llu answered from inputs that were replaced during the call\n",
              (unsigned long long)calls, (unsigned long long)stale);
  }
  free(snap);
  return 0;
}
