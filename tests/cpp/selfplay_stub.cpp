// The self-play driver (unrealgo_b200/csrc/selfplay.cpp) on the CPU: the leaf queue is replaced by a stub that
// answers at the next poll with a policy/value computed from the submitted position alone (stub_eval.h: a hash of
// its 17 feature planes), so the whole search — descent, expansion, priors, backup, tree reuse, move choice, virtual loss — runs
// without a GPU and is a pure function of the seed.  Prints the stats and every game's moves.
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "../../include/ugeval.h"
#include "stub_eval.h"

struct StubSlot { ug_packed_pos pos; uint32_t legal[UG_MASK_WORDS]; };
struct ug_queue { std::mutex mu; std::vector<StubSlot*> slots; std::vector<int64_t> free_ids; };
struct ug_eval { int unused; };
#ifndef UGSTUB_REAL_TFRECORD
struct ug_tfrecord { int unused; };
#endif

extern "C" {
#ifndef UGSTUB_REAL_TFRECORD  // (with it the real csrc/tfrecord.cpp is linked and UGSTUB_RECORD_DIR receives the files)
// with UGSTUB_PRINT_VISITS=1 the driver is asked for game records and each record's policy block — the root's visit
// counts of that move (search/UctSearch.cpp:782-801) — is printed instead of written
static ug_tfrecord g_rec;
static int g_rec_no = 0;
ug_tfrecord* ug_tfrecord_open(const char*) { g_rec_no = 0; return &g_rec; }
int ug_tfrecord_write_example(ug_tfrecord*, const void*, size_t, const float* policy, size_t n, float) {
  printf("visits %d", g_rec_no++);
  for (size_t i = 0; i < n; ++i) if (policy[i] > 0) printf(" %zu:%.0f", i, policy[i]);
  printf("\n");
  return 0;
}
int64_t ug_tfrecord_close(ug_tfrecord*) { return g_rec_no; }
#endif
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
  float sum = 0.f;
  for (int i = 0; i < UG_NUM_MOVES; ++i) {
    if (!((s.legal[i >> 5] >> (i & 31)) & 1u)) policy[i] = 0.f;
    sum += policy[i];
  }
  for (int i = 0; i < UG_NUM_MOVES; ++i) policy[i] /= sum;
  return 1;
}
int ug_queue_wait(ug_queue* q, int64_t t, float* p, float* v) { return ug_queue_poll(q, t, p, v) == 1 ? 0 : -1; }
int ug_queue_stats(ug_queue*, uint64_t* a, uint64_t* b, uint64_t* c) { *a = *b = *c = 0; return 0; }
}

int main(int argc, char** argv) {
  if (argc < 7) { fprintf(stderr, "usage: games threads playouts moves reuse leaves_per_game [random_opening_moves=2] [puct]\n"); return 2; }
  ug_selfplay_config cfg{};
  cfg.games = atoi(argv[1]); cfg.threads = atoi(argv[2]); cfg.playouts_per_move = atoi(argv[3]);
  cfg.max_moves = atoi(argv[4]); cfg.reuse_tree = atoi(argv[5]); cfg.leaves_per_game = atoi(argv[6]);
  cfg.max_batch = 64; cfg.timeout_us = 100; cfg.random_opening_moves = argc > 7 ? atoi(argv[7]) : 2; cfg.seed = 9;
  if (argc > 8) cfg.puct = (float)atof(argv[8]);
  if (getenv("UGSTUB_PRINT_VISITS")) cfg.record_dir = "/tmp";
  if (getenv("UGSTUB_RECORD_DIR")) cfg.record_dir = getenv("UGSTUB_RECORD_DIR");
  std::vector<int16_t> moves((size_t)cfg.games * cfg.max_moves);
  std::vector<int> n_moves((size_t)cfg.games);
  ug_selfplay_stats st{};
  ug_eval e{};
  const int rc = ug_selfplay_run(&e, &cfg, &st, moves.data(), n_moves.data());
  printf("rc %d playouts %llu moves %llu games_finished %llu terminal %llu\n", rc, (unsigned long long)st.playouts,
         (unsigned long long)st.moves, (unsigned long long)st.games_finished, (unsigned long long)st.terminal_leaves);
  for (int g = 0; g < cfg.games; ++g) {
    printf("game %d:", g);
    for (int i = 0; i < n_moves[(size_t)g]; ++i) printf(" %d", moves[(size_t)g * cfg.max_moves + i]);
    printf("\n");
  }
  return rc;
}
