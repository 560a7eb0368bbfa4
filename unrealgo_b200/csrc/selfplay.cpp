// Self-play driver: many independent games per GPU, each running the reference's single-threaded PUCT search,
// multiplexed over a few host threads so that hundreds of leaves are in flight at once and the GPU sees full
// batches through the leaf queue (queue.cpp).
//
// Per game this is the loop of UctDeepPlayer::SelfPlayOneGame -> DoSearch -> UctSearch::DeepUctSearchTree
// (search/UctDeepPlayer.cpp:61-121, :288-349; search/UctSearch.cpp:803-854) as built with FORCE_SINGLE_THREAD
// (CMakeLists.txt:15): one playout = descend with Select (PUCT, c = 2.5, :1062-1098; :370), expand the leaf
// with all generated moves, evaluate it with the network, UpdatePrior (:1257-1270; done on the GPU through the
// legal mask), BackupTree with the sign flipping every ply (:1282-1295).  The move played is the child with
// the largest visit count (:782-801, :945); with reuse_searchtree the chosen subtree is kept
// (search/UctDeepPlayer.cpp:223-259).  A game ends after two passes or 722 moves (:30, :74, :82).
// What differs from the reference is only *where a search waits*: instead of one blocked thread per game
// (search/UctSearch.cpp:515-517) a worker thread parks the game and advances another one.
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <random>
#include <thread>
#include <vector>

#include "../../include/ugeval.h"
#include "goboard.h"

namespace {

struct Node {
  float prior;
  float w;            // sum of backed-up values (perspective of the player who made `move`)
  uint32_t visits;
  uint32_t first_child;
  uint32_t parent;
  uint16_t n_children;
  uint16_t move;      // policy index, 361 = pass
};

struct Game {
  ugo::Board root_board;
  ugo::Board work;              // board at the pending leaf
  std::vector<Node> tree, scratch;
  uint32_t leaf = 0;
  int64_t ticket = -1;          // >= 0: a leaf is waiting for its evaluation
  int playouts_this_move = 0;
  bool finished = false;
  std::vector<int16_t> moves;
  std::vector<float> visit_policy;  // [moves][362] root visit counts per move (only when records are written)
  int index = 0;                    // game number on this evaluator
  std::mt19937_64 rng;
  float policy[UG_NUM_MOVES];
};

struct Shared {
  ug_queue* q;
  ug_selfplay_config cfg;
  std::atomic<uint64_t> playouts{0}, moves{0}, games_done{0}, nodes{0}, terminal_leaves{0};
  std::atomic<uint64_t> record_errors{0};
  std::atomic<bool> stop{false};
};

// CollectFeatures (gouct/GoUctGlobalSearch.h:160-206) from a packed position: plane 2k / 2k+1 = stones of the
// side to move / the opponent k moves ago, plane 16 = 1 - to_play
void planes_from_packed(const ug_packed_pos& p, int8_t* out) {
  const uint32_t tp = p.to_play & 1u;
  for (int k = 0; k < UG_HISTORY; ++k) {
    const uint32_t* mine = tp == 0 ? p.black[k] : p.white[k];
    const uint32_t* theirs = tp == 0 ? p.white[k] : p.black[k];
    for (int i = 0; i < UG_NUM_POINTS; ++i) {
      out[(2 * k) * UG_NUM_POINTS + i] = (int8_t)((mine[i >> 5] >> (i & 31)) & 1u);
      out[(2 * k + 1) * UG_NUM_POINTS + i] = (int8_t)((theirs[i >> 5] >> (i & 31)) & 1u);
    }
  }
  memset(out + 16 * UG_NUM_POINTS, tp == 0 ? 1 : 0, UG_NUM_POINTS);
}

// UctDeepPlayer::WriteTFRecord, search/UctDeepPlayer.cpp:362-388 (see ug_selfplay_config.record_dir)
bool write_game_record(const Game& g, const ug_selfplay_config& cfg) {
  const int n = (int)g.moves.size();
  int last = n - 1;
  while (last >= 0 && g.moves[(size_t)last] == ugo::PASS) --last;  // trailing passes are taken back first (:371-374)
  char path[4096];
  snprintf(path, sizeof(path), "%s/selfplay_%d.tfrecord", cfg.record_dir, cfg.game_index_base + g.index);
  ug_tfrecord* w = ug_tfrecord_open(path);
  if (!w) return false;
  std::vector<ug_packed_pos> after((size_t)(last + 1));
  ugo::Board b;
  for (int i = 0; i <= last; ++i) {
    b.play(g.moves[(size_t)i]);
    b.pack(&after[(size_t)i]);
  }
  const float score = b.tromp_taylor(cfg.komi > 0 ? cfg.komi : 7.5f);
  const int colour = (last % 2 == 0) ? ugo::BLACK : ugo::WHITE;  // Black moves first
  const bool win = (score > 0 && colour == ugo::BLACK) || (score < 0 && colour == ugo::WHITE);
  float reward = win ? 1.f : -1.f;
  int8_t planes[UG_NUM_PLANES * UG_NUM_POINTS];
  bool ok = true;
  for (int i = last; i >= 0 && ok; --i) {
    planes_from_packed(after[(size_t)i], planes);
    ok = ug_tfrecord_write_example(w, planes, sizeof(planes), &g.visit_policy[(size_t)i * UG_NUM_MOVES],
                                   UG_NUM_MOVES, reward) == 0;
    reward = -reward;
  }
  return ug_tfrecord_close(w) >= 0 && ok;
}

inline float mean_value(const Node& n) { return n.visits ? n.w / (float)n.visits : 0.f; }

// UctSearch::Select, search/UctSearch.cpp:1062-1098 (num_threads == 1: no virtual loss)
uint32_t select_child(const std::vector<Node>& t, uint32_t parent, float c_puct) {
  const Node& p = t[parent];
  double total = 0;
  for (uint32_t i = 0; i < p.n_children; ++i) total += t[p.first_child + i].visits;
  const float spc = (float)std::max(std::sqrt(total), 1.0);
  const double default_mean = -(double)mean_value(p);
  double best_v = std::numeric_limits<double>::min();  // sic: the reference initialises with min(), not lowest()
  int64_t best = -1;
  for (uint32_t i = 0; i < p.n_children; ++i) {
    const Node& c = t[p.first_child + i];
    if (c.move == ugo::PASS) continue;  // PASS is only taken when it is the only child (:1082, :1093-1096)
    const double q = c.visits ? (double)mean_value(c) : default_mean;
    const double v = q + (double)c_puct * c.prior * spc / (1.0 + c.visits);
    if (best < 0 || v > best_v) { best = p.first_child + i; best_v = v; }
  }
  return best < 0 ? p.first_child : (uint32_t)best;
}

void new_root(Game& g) {
  g.tree.clear();
  Node r{};
  r.parent = 0;
  r.move = ugo::PASS;
  g.tree.push_back(r);
}

// keep the subtree below `child` as the new tree (UctSearchTree::CopySubtree, search/UctSearchTree.cpp:159-225)
void reuse_subtree(Game& g, uint32_t child) {
  std::vector<Node>& src = g.tree;
  std::vector<Node>& dst = g.scratch;
  dst.clear();
  dst.push_back(src[child]);
  dst[0].parent = 0;
  // breadth-first copy keeps every node's children contiguous
  std::vector<uint32_t> map_src;  // dst index -> src index
  map_src.push_back(child);
  for (size_t di = 0; di < dst.size(); ++di) {
    const Node s = src[map_src[di]];
    if (s.n_children == 0) continue;
    const uint32_t base = (uint32_t)dst.size();
    dst[di].first_child = base;
    for (uint32_t i = 0; i < s.n_children; ++i) {
      Node c = src[s.first_child + i];
      c.parent = (uint32_t)di;
      dst.push_back(c);
      map_src.push_back(s.first_child + i);
    }
  }
  src.swap(dst);
}

// one descent; returns true if a leaf was submitted (false: terminal leaf, counted as a playout like the
// reference's ExpandAndBackup returning false, search/UctSearch.cpp:478-490)
bool start_playout(Game& g, Shared& sh) {
  g.work = g.root_board;
  uint32_t node = 0;
  while (g.tree[node].n_children) {
    node = select_child(g.tree, node, sh.cfg.puct);
    g.work.play(g.tree[node].move);
  }
  uint16_t moves[UG_NUM_MOVES];
  uint32_t legal[UG_MASK_WORDS];
  const int n = g.work.generate_moves(moves, legal);
  if (n == 0) {
    sh.terminal_leaves.fetch_add(1, std::memory_order_relaxed);
    return false;
  }
  const uint32_t base = (uint32_t)g.tree.size();
  g.tree[node].first_child = base;
  g.tree[node].n_children = (uint16_t)n;
  for (int i = 0; i < n; ++i) {
    Node c{};
    c.parent = node;
    c.move = moves[i];
    g.tree.push_back(c);  // prior 0 until the evaluation arrives (UctSearchTree::Expand, :675-691)
  }
  ug_packed_pos pos;
  g.work.pack(&pos);
  g.leaf = node;
  g.ticket = ug_queue_submit(sh.q, &pos, legal);
  if (g.ticket < 0) sh.stop.store(true);  // queue closed or evaluator failed: every worker winds down
  return g.ticket >= 0;
}

void finish_playout(Game& g, float value) {
  Node& leaf = g.tree[g.leaf];
  for (uint32_t i = 0; i < leaf.n_children; ++i) {
    Node& c = g.tree[leaf.first_child + i];
    c.prior = g.policy[c.move];  // already renormalised over the legal children on the GPU (UpdatePrior)
  }
  uint32_t cur = g.leaf;
  float v = value;
  for (;;) {  // BackupTree, search/UctSearch.cpp:1282-1295
    Node& n = g.tree[cur];
    n.visits += 1;
    n.w += v;
    v = -v;
    if (cur == 0) break;
    cur = n.parent;
  }
}

// after playouts_per_move playouts: play the most visited child (first wins ties), keep or drop the tree
void make_move(Game& g, Shared& sh) {
  const Node& r = g.tree[0];
  int move = ugo::PASS;
  uint32_t chosen = 0;
  if (r.n_children) {
    if (g.root_board.move_number < sh.cfg.random_opening_moves) {
      // harness option (not in the reference): diversify games with uniformly random legal opening moves
      std::uniform_int_distribution<uint32_t> d(0, r.n_children > 1 ? r.n_children - 2 : 0);  // PASS is last
      chosen = r.first_child + d(g.rng);
    } else {
      uint32_t best_visits = 0;
      chosen = r.first_child;
      bool first = true;
      for (uint32_t i = 0; i < r.n_children; ++i) {
        const Node& c = g.tree[r.first_child + i];
        if (first || c.visits > best_visits) { best_visits = c.visits; chosen = r.first_child + i; first = false; }
      }
    }
    move = g.tree[chosen].move;
  }
  if (sh.cfg.record_dir) {
    // the policy block of the move: zero, then the visit count of every root child (search/UctSearch.cpp:782-801)
    const size_t at = g.visit_policy.size();
    g.visit_policy.resize(at + UG_NUM_MOVES, 0.f);
    for (uint32_t i = 0; i < r.n_children; ++i) {
      const Node& c = g.tree[r.first_child + i];
      g.visit_policy[at + c.move] = (float)c.visits;
    }
  }
  g.root_board.play(move);
  g.moves.push_back((int16_t)move);
  sh.moves.fetch_add(1, std::memory_order_relaxed);
  sh.nodes.fetch_add(g.tree.size(), std::memory_order_relaxed);
  if (sh.cfg.reuse_tree && r.n_children && g.tree[chosen].n_children) reuse_subtree(g, chosen);
  else new_root(g);
  g.playouts_this_move = 0;
  const int limit = sh.cfg.max_moves > 0 ? sh.cfg.max_moves : 2 * 19 * 19;
  if (g.root_board.passes >= 2 || g.root_board.move_number >= limit) {
    g.finished = true;
    sh.games_done.fetch_add(1, std::memory_order_relaxed);
    if (sh.cfg.record_dir && !write_game_record(g, sh.cfg)) sh.record_errors.fetch_add(1, std::memory_order_relaxed);
  }
}

// after a playout completed (or at game start): play the move when the budget is used up, then put the next
// leaf in flight.  Terminal leaves (no moves after two passes) cost a playout but no evaluation.
// Postcondition: the game is finished or has a ticket >= 0.
void advance(Game& g, Shared& sh) {
  while (!g.finished) {
    if (g.playouts_this_move >= sh.cfg.playouts_per_move) {
      make_move(g, sh);
      continue;
    }
    if (start_playout(g, sh)) return;
    if (sh.stop.load(std::memory_order_relaxed)) return;  // queue closed
    ++g.playouts_this_move;
    sh.playouts.fetch_add(1, std::memory_order_relaxed);
  }
}

void worker(std::vector<Game>& games, size_t lo, size_t hi, Shared& sh) {
  size_t live = 0;
  for (size_t i = lo; i < hi; ++i) if (!games[i].finished) ++live;
  while (live > 0 && !sh.stop.load(std::memory_order_relaxed)) {
    bool progressed = false;
    for (size_t i = lo; i < hi; ++i) {
      Game& g = games[i];
      if (g.finished) continue;
      float value = 0.f;
      const int rc = ug_queue_poll(sh.q, g.ticket, g.policy, &value);
      if (rc == 0) continue;  // not ready yet: look at the next game
      if (rc < 0) { sh.stop.store(true); return; }
      g.ticket = -1;
      finish_playout(g, value);
      ++g.playouts_this_move;
      sh.playouts.fetch_add(1, std::memory_order_relaxed);
      advance(g, sh);
      if (g.finished) --live;
      progressed = true;
    }
    if (!progressed) std::this_thread::sleep_for(std::chrono::microseconds(20));
  }
}

}  // namespace

extern "C" int ug_selfplay_run(ug_eval* h, const ug_selfplay_config* cfg_in, ug_selfplay_stats* stats,
                               int16_t* moves_out, int* n_moves_out) {
  if (!h || !cfg_in || !stats) return -1;
  ug_selfplay_config cfg = *cfg_in;
  if (cfg.games <= 0 || cfg.playouts_per_move <= 0) return -1;
  if (cfg.threads <= 0) cfg.threads = 1;
  if (cfg.threads > cfg.games) cfg.threads = cfg.games;
  if (cfg.puct <= 0) cfg.puct = 2.5f;  // puct_const, search/UctSearch.cpp:370
  if (cfg.max_batch <= 0) cfg.max_batch = 256;
  Shared sh;
  sh.cfg = cfg;
  sh.q = ug_queue_create(h, cfg.max_batch, cfg.timeout_us > 0 ? cfg.timeout_us : 200, 4 * cfg.games + 1024);
  if (!sh.q) return -2;
  std::vector<Game> games((size_t)cfg.games);
  for (int i = 0; i < cfg.games; ++i) {
    games[i].rng.seed(cfg.seed * 1000003ull + (uint64_t)i);
    games[i].index = i;
    new_root(games[i]);
  }
  const auto t0 = std::chrono::steady_clock::now();
  // prime: every game puts its first leaf (the fresh root) in flight
  for (Game& g : games) advance(g, sh);
  std::vector<std::thread> threads;
  const size_t per = ((size_t)cfg.games + cfg.threads - 1) / cfg.threads;
  for (int t = 0; t < cfg.threads; ++t) {
    const size_t lo = (size_t)t * per, hi = std::min((size_t)cfg.games, lo + per);
    if (lo >= hi) break;
    threads.emplace_back([&games, lo, hi, &sh] { worker(games, lo, hi, sh); });
  }
  // optional progress log (UGEVAL_SP_LOG=<seconds>): playouts and queue counters while the games run
  std::atomic<bool> done{false};
  std::thread monitor;
  if (const char* lg = getenv("UGEVAL_SP_LOG")) {
    const double every = atof(lg) > 0 ? atof(lg) : 2.0;
    monitor = std::thread([&sh, &done, every, t0] {
      auto next = std::chrono::steady_clock::now();
      while (!done.load(std::memory_order_acquire)) {
        std::this_thread::sleep_for(std::chrono::milliseconds(50));
        if (std::chrono::steady_clock::now() < next) continue;
        next += std::chrono::microseconds((long long)(every * 1e6));
        uint64_t b = 0, l = 0, to = 0;
        ug_queue_stats(sh.q, &b, &l, &to);
        fprintf(stderr, "[selfplay %.1fs] playouts=%llu moves=%llu games_done=%llu batches=%llu leaves=%llu timeouts=%llu\n",
                std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(),
                (unsigned long long)sh.playouts.load(), (unsigned long long)sh.moves.load(),
                (unsigned long long)sh.games_done.load(), (unsigned long long)b, (unsigned long long)l,
                (unsigned long long)to);
      }
    });
  }
  for (std::thread& t : threads) t.join();
  done.store(true, std::memory_order_release);
  if (monitor.joinable()) monitor.join();
  const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  uint64_t batches = 0, leaves = 0, timeouts = 0;
  ug_queue_stats(sh.q, &batches, &leaves, &timeouts);
  ug_queue_destroy(sh.q);
  stats->playouts = sh.playouts.load();
  stats->evaluations = leaves;
  stats->batches = batches;
  stats->flush_timeouts = timeouts;
  stats->moves = sh.moves.load();
  stats->games_finished = sh.games_done.load();
  stats->terminal_leaves = sh.terminal_leaves.load();
  stats->seconds = secs;
  stats->playouts_per_s = secs > 0 ? (double)stats->playouts / secs : 0;
  stats->mean_batch = batches ? (double)leaves / (double)batches : 0;
  stats->mean_tree_nodes = stats->moves ? (double)sh.nodes.load() / (double)stats->moves : 0;
  if (moves_out && n_moves_out) {
    const int cap = cfg.max_moves > 0 ? cfg.max_moves : 2 * 19 * 19;
    for (int i = 0; i < cfg.games; ++i) {
      const int n = (int)std::min<size_t>(games[i].moves.size(), (size_t)cap);
      n_moves_out[i] = n;
      memcpy(moves_out + (size_t)i * cap, games[i].moves.data(), (size_t)n * sizeof(int16_t));
    }
  }
  if (sh.record_errors.load()) return -4;  // a record file could not be written
  return sh.stop.load() ? -3 : 0;
}

// ---- test hooks: drive the product board from Python to compare with the oracle's restated GoBoard
extern "C" {
void* ug_board_new(void) { return new ugo::Board(); }
void ug_board_free(void* b) { delete static_cast<ugo::Board*>(b); }
int ug_board_legal(void* b, int idx) {
  ugo::Board* bd = static_cast<ugo::Board*>(b);
  return idx == ugo::PASS ? 1 : (bd->legal(ugo::to_b(idx), bd->to_play) ? 1 : 0);
}
void ug_board_play(void* b, int idx) { static_cast<ugo::Board*>(b)->play(idx); }
void ug_board_pack(void* b, ug_packed_pos* out) { static_cast<ugo::Board*>(b)->pack(out); }
int ug_board_generate(void* b, uint16_t* moves, uint32_t* legal) {
  return static_cast<ugo::Board*>(b)->generate_moves(moves, legal);
}
int ug_board_to_play(void* b) { return static_cast<ugo::Board*>(b)->to_play; }
float ug_board_score(void* b, float komi) { return static_cast<ugo::Board*>(b)->tromp_taylor(komi); }
void ug_board_planes(void* b, int8_t* chw_out) {
  ug_packed_pos p;
  static_cast<ugo::Board*>(b)->pack(&p);
  planes_from_packed(p, chw_out);
}
}
