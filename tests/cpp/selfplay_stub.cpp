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
}
#include "stub_queue.h"

int main(int argc, char** argv) {
  if (argc < 7) { fprintf(stderr, "usage: games threads playouts moves reuse leaves_per_game [random_opening_moves=2] [puct] [prelude_moves]\n"); return 2; }
  ug_selfplay_config cfg{};
  cfg.games = atoi(argv[1]); cfg.threads = atoi(argv[2]); cfg.playouts_per_move = atoi(argv[3]);
  cfg.max_moves = atoi(argv[4]); cfg.reuse_tree = atoi(argv[5]); cfg.leaves_per_game = atoi(argv[6]);
  cfg.max_batch = 64; cfg.timeout_us = 100; cfg.random_opening_moves = argc > 7 ? atoi(argv[7]) : 2; cfg.seed = 9;
  if (argc > 8) cfg.puct = (float)atof(argv[8]);
  if (argc > 9) cfg.prelude_moves = atoi(argv[9]);
  const int row = cfg.max_moves + cfg.prelude_moves;
  if (getenv("UGSTUB_PRINT_VISITS")) cfg.record_dir = "/tmp";
  if (getenv("UGSTUB_RECORD_DIR")) cfg.record_dir = getenv("UGSTUB_RECORD_DIR");
  std::vector<int16_t> moves((size_t)cfg.games * row);
  std::vector<int> n_moves((size_t)cfg.games);
  ug_selfplay_stats st{};
  ug_eval e{};
  const int rc = ug_selfplay_run(&e, &cfg, &st, moves.data(), n_moves.data());
  printf("rc %d playouts %llu moves %llu games_finished %llu terminal %llu\n", rc, (unsigned long long)st.playouts,
         (unsigned long long)st.moves, (unsigned long long)st.games_finished, (unsigned long long)st.terminal_leaves);
  for (int g = 0; g < cfg.games; ++g) {
    printf("game %d:", g);
    for (int i = 0; i < n_moves[(size_t)g]; ++i) printf(" %d", moves[(size_t)g * row + i]);
    printf("\n");
  }
  return rc;
}
