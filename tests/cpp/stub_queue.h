// The stub leaf queue of the CPU tests: ug_queue_* answered synchronously at poll/wait time by the stand-in network of
// stub_eval.h, applied to the planes the packed position expands to, followed by the legal-move renormalisation the
// head kernel does in float.  Included by selfplay_stub.cpp (the product's self-play driver on the CPU) and by
// ugqueue_stub.cpp (preloaded under oracle/_ref/ref_search_queue, the reference's engine with the leaf-queue patch).
// Test infrastructure only.
#ifndef UG_TEST_STUB_QUEUE_H
#define UG_TEST_STUB_QUEUE_H
#include <cstring>
#include <mutex>
#include <vector>

#include "../../include/ugeval.h"
#include "stub_eval.h"

struct StubSlot { ug_packed_pos pos; uint32_t legal[UG_MASK_WORDS]; };
struct ug_queue { std::mutex mu; std::vector<StubSlot*> slots; std::vector<int64_t> free_ids; bool compact = false; };

extern "C" {
ug_queue* ug_queue_create(ug_eval*, int, int, int) { return new ug_queue(); }
void ug_queue_destroy(ug_queue* q) { for (StubSlot* s : q->slots) delete s; delete q; }
int64_t ug_queue_submit(ug_queue* q, const ug_packed_pos* pos, const uint32_t* legal) {
  std::lock_guard<std::mutex> lk(q->mu);
  int64_t id;
  if (!q->free_ids.empty()) { id = q->free_ids.back(); q->free_ids.pop_back(); }
  else { id = (int64_t)q->slots.size(); q->slots.push_back(new StubSlot()); }
  q->slots[(size_t)id]->pos = *pos;
  memcpy(q->slots[(size_t)id]->legal, legal, sizeof(q->slots[0]->legal));
  return id;
}
int ug_queue_poll(ug_queue* q, int64_t ticket, float* policy, float* value) {
  if (ticket < 0) return -1;
  StubSlot s;
  {
    std::lock_guard<std::mutex> lk(q->mu);
    s = *q->slots[(size_t)ticket];
    q->free_ids.push_back(ticket);
  }
  // the 17 planes CollectFeatures would hand the network (csrc/encode.cu does this on the GPU), then the shared stub
  // network, then the legal-move renormalisation the head kernel applies in float (UpdatePrior)
  const uint32_t* w = reinterpret_cast<const uint32_t*>(&s.pos);
  const uint32_t tp = w[2 * UG_HISTORY * UG_MASK_WORDS] & 1u;
  int8_t planes[17 * 361];
  for (int k = 0; k < UG_HISTORY; ++k)
    for (int c = 0; c < 2; ++c) {
      const uint32_t* m = w + (c * UG_HISTORY + k) * UG_MASK_WORDS;  // black[8][12] then white[8][12]
      int8_t* out = planes + (2 * k + (int)(c ^ tp)) * 361;
      for (int i = 0; i < 361; ++i) out[i] = (int8_t)((m[i >> 5] >> (i & 31)) & 1u);
    }
  memset(planes + 16 * 361, 1 - (int)tp, 361);
  stub_eval(planes, policy, value);
  *value = stub_transform(*value);  // the head kernel applies value_transform on the packed path
  float sum = 0.f;
  for (int i = 0; i < UG_NUM_MOVES; ++i) {
    if (!((s.legal[i >> 5] >> (i & 31)) & 1u)) policy[i] = 0.f;
    sum += policy[i];
  }
  for (int i = 0; i < UG_NUM_MOVES; ++i) policy[i] /= sum;
  if (q->compact) {  // only the legal priors, ascending move order (ug_queue_set_compact_priors)
    int k = 0;
    for (int i = 0; i < UG_NUM_MOVES; ++i)
      if ((s.legal[i >> 5] >> (i & 31)) & 1u) policy[k++] = policy[i];
  }
  return 1;
}
int ug_queue_wait(ug_queue* q, int64_t t, float* p, float* v) { return ug_queue_poll(q, t, p, v) == 1 ? 0 : -1; }
int ug_queue_stats(ug_queue*, uint64_t* a, uint64_t* b, uint64_t* c) { *a = *b = *c = 0; return 0; }
int ug_queue_swap_checkpoint(ug_queue*, const char*) { return 0; }
int ug_queue_set_compact_priors(ug_queue* q, int on) { q->compact = on != 0; return 0; }
int ug_queue_stats_ex(ug_queue*, uint64_t* out, int n) { for (int i = 0; i < n; ++i) out[i] = 0; return 0; }
}
#endif
