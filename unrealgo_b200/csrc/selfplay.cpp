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

#if defined(__linux__)
#include <sys/mman.h>
#endif

#include "../../include/ugeval.h"
#include "goboard.h"

namespace {

// Search tree, structure-of-arrays: one descent reads prior/w/visits of every child of every node on the path
// (up to 362 children per level), so the three hot fields live in their own arrays (12 bytes per child instead of
// a 24-byte record) and the scan in select_child() is a plain loop over contiguous floats.
struct Tree {
  std::vector<float> prior;          // network prior of the move leading here
  std::vector<float> w;              // sum of backed-up values (perspective of the player who made `move`)
  std::vector<uint32_t> visits;
  std::vector<uint32_t> first_child; // children of a node are contiguous
  std::vector<uint32_t> parent;
  std::vector<uint16_t> n_children;
  std::vector<uint16_t> move;        // policy index, 361 = pass
  std::vector<uint16_t> vloss;       // UctNode::v_loss_cnt (search/UctSearchTree.h:242-252); empty unless several
                                     // leaves of one game are in flight (use_vloss)
  bool use_vloss = false;
  size_t size() const { return visits.size(); }
  void clear() {
    prior.clear(); w.clear(); visits.clear(); first_child.clear(); parent.clear(); n_children.clear(); move.clear();
    vloss.clear();
  }
  void reserve(size_t n, bool huge = false);
  void resize(size_t n) {
    prior.resize(n, 0.f); w.resize(n, 0.f); visits.resize(n, 0u); first_child.resize(n, 0u); parent.resize(n, 0u);
    n_children.resize(n, 0); move.resize(n, 0);
    if (use_vloss) vloss.resize(n, 0);
  }
  void swap(Tree& o) {
    prior.swap(o.prior); w.swap(o.w); visits.swap(o.visits); first_child.swap(o.first_child); parent.swap(o.parent);
    n_children.swap(o.n_children); move.swap(o.move);
    vloss.swap(o.vloss);
  }
  void assign_from(const Tree& o) {
    prior.assign(o.prior.begin(), o.prior.end()); w.assign(o.w.begin(), o.w.end());
    visits.assign(o.visits.begin(), o.visits.end()); first_child.assign(o.first_child.begin(), o.first_child.end());
    parent.assign(o.parent.begin(), o.parent.end()); n_children.assign(o.n_children.begin(), o.n_children.end());
    move.assign(o.move.begin(), o.move.end());
    if (use_vloss) vloss.assign(o.vloss.begin(), o.vloss.end());
  }
  void copy_node(size_t dst, const Tree& src, size_t s) {
    prior[dst] = src.prior[s]; w[dst] = src.w[s]; visits[dst] = src.visits[s]; first_child[dst] = src.first_child[s];
    parent[dst] = src.parent[s]; n_children[dst] = src.n_children[s]; move[dst] = src.move[s];
    if (use_vloss) vloss[dst] = 0;  // no leaf is in flight when a subtree is carried over
  }
};

// The trees of the games on one GPU add up to gigabytes walked at random (768 games x ~400 k nodes x 24 bytes): ask for
// transparent huge pages so that a descent does not pay a TLB miss per level (no-op where THP is off or `always`).
template <typename T>
void advise_huge(std::vector<T>& v) {
#if defined(__linux__) && defined(MADV_HUGEPAGE)
  const uintptr_t lo = (reinterpret_cast<uintptr_t>(v.data()) + 4095) & ~uintptr_t(4095);
  const uintptr_t hi = (reinterpret_cast<uintptr_t>(v.data()) + v.capacity() * sizeof(T)) & ~uintptr_t(4095);
  if (hi > lo + (2u << 20)) madvise(reinterpret_cast<void*>(lo), hi - lo, MADV_HUGEPAGE);
#else
  (void)v;
#endif
}

void Tree::reserve(size_t n, bool huge) {
  prior.reserve(n); w.reserve(n); visits.reserve(n); first_child.reserve(n); parent.reserve(n); n_children.reserve(n);
  move.reserve(n);
  if (use_vloss) vloss.reserve(n);
  if (huge) {
    advise_huge(prior); advise_huge(w); advise_huge(visits); advise_huge(first_child); advise_huge(parent);
    advise_huge(n_children); advise_huge(move);
  }
}

struct Game {
  ugo::Board root_board;
  ugo::Board work;              // board at the pending leaf
  Tree tree, scratch;
  std::vector<uint32_t> map_src;  // scratch of reuse_subtree
  // leaves waiting for their evaluation: one in the shipped single-thread configuration, up to leaves_per_game
  // with the reference's multi-thread virtual-loss search
  static constexpr int kMaxLeaves = 64;
  struct Pending { uint32_t leaf; int64_t ticket; };
  Pending pend[kMaxLeaves];
  int npend = 0;
  int playouts_this_move = 0;
  bool finished = false;
  std::vector<int16_t> moves;
  std::vector<float> visit_policy;  // [moves][362] root visit counts per move (only when records are written)
  int index = 0;                    // game number on this evaluator
  int first_search_move = 0;        // moves already on the board when the first search starts (prelude_moves)
  std::mt19937_64 rng;
  float policy[UG_NUM_MOVES];
};

struct Shared {
  ug_queue* q;
  ug_selfplay_config cfg;
  std::atomic<uint64_t> playouts{0}, moves{0}, games_done{0}, nodes{0}, terminal_leaves{0};
  std::atomic<uint64_t> record_errors{0};
  std::atomic<uint64_t> idle_ns{0};   // search threads: time asleep because no evaluation had come back
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

// value_i = Q_i + c_puct * prior_i * spc / (1 + visits_i), first maximum wins; Q of an unvisited child is
// default_mean.  Two passes so that the arithmetic pass is branch-free and vectorises (the children's prior / w /
// visits are contiguous arrays); the operation order per element is the reference's (search/UctSearch.cpp:1084).
#if defined(__GNUC__) && defined(__x86_64__)
__attribute__((optimize("O3"), target("avx2,fma")))
static uint32_t argmax_puct_avx2(const float* __restrict__ pr, const float* __restrict__ ww,
                                 const uint32_t* __restrict__ vv, uint32_t n, double default_mean, double cp, float spc) {
  double val[UG_NUM_MOVES];
  const double dspc = (double)spc;
  for (uint32_t i = 0; i < n; ++i) {
    const uint32_t vis = vv[i];
    const float qf = ww[i] / (float)(vis ? vis : 1u);
    const double q = vis ? (double)qf : default_mean;
    val[i] = q + cp * (double)pr[i] * dspc / (1.0 + (double)vis);
  }
  uint32_t best = 0;
  double best_v = val[0];
  for (uint32_t i = 1; i < n; ++i)
    if (val[i] > best_v) { best_v = val[i]; best = i; }
  return best;
}
#endif
static uint32_t argmax_puct_scalar(const float* __restrict__ pr, const float* __restrict__ ww,
                                   const uint32_t* __restrict__ vv, uint32_t n, double default_mean, double cp, float spc) {
  double best_v = 0;
  uint32_t best = 0;
  for (uint32_t i = 0; i < n; ++i) {
    const uint32_t vis = vv[i];
    const double q = vis ? (double)(ww[i] / (float)vis) : default_mean;
    const double v = q + cp * (double)pr[i] * (double)spc / (1.0 + (double)vis);
    if (i == 0 || v > best_v) { best = i; best_v = v; }
  }
  return best;
}
static uint32_t argmax_puct(const float* pr, const float* ww, const uint32_t* vv, uint32_t n, double default_mean,
                            double cp, float spc) {
#if defined(__GNUC__) && defined(__x86_64__)
  static const bool have_avx2 = __builtin_cpu_supports("avx2") && __builtin_cpu_supports("fma");
  if (have_avx2) return argmax_puct_avx2(pr, ww, vv, n, default_mean, cp, spc);
#endif
  return argmax_puct_scalar(pr, ww, vv, n, default_mean, cp, spc);
}

// UctSearch::Select, search/UctSearch.cpp:1062-1098 (num_threads == 1: no virtual loss).
// The reference sums the children's visit counts; every backup through a node passes through exactly one child
// except the one that expanded it, so that sum is visits(parent) - 1 and the first pass over the children is saved.
uint32_t select_child(const Tree& t, uint32_t parent, float c_puct) {
  const uint32_t base = t.first_child[parent];
  uint32_t n = t.n_children[parent];
  const uint32_t pv = t.visits[parent];
  const double total = pv > 0 ? (double)(pv - 1) : 0.0;
  const float spc = (float)std::max(std::sqrt(total), 1.0);  // float, as in the reference (:1075)
  const double default_mean = pv ? -(double)(t.w[parent] / (float)pv) : 0.0;
  // PASS is generated last and only taken when it is the only child (:1082, :1093-1096)
  if (n > 1 && t.move[base + n - 1] == ugo::PASS) --n;
  else if (n == 1) return base;
  const uint32_t best = argmax_puct(t.prior.data() + base, t.w.data() + base, t.visits.data() + base, n, default_mean,
                                    (double)c_puct, spc);
  return base + best;
}

// The same selection with several descents of one tree in flight (the reference with num_threads > 1 and
// use_virtual_loss): the parent's total gains its virtual-loss count minus one (search/UctSearch.cpp:1066-1070) and a
// child's mean is diluted by its virtual losses counted as zero-valued visits — the statistics branch of GetMeanValue
// (:665-679).  The shipped build defines USE_DIRECT_VISITCOUNT (CMakeLists.txt:16), whose branch (:661-664) returns the
// plain mean and so does not steer concurrent descents apart at all; with several leaves of one tree in flight that
// would send every descent to the same leaf, hence the other branch here.
uint32_t select_child_vloss(const Tree& t, uint32_t parent, float c_puct) {
  const uint32_t base = t.first_child[parent];
  uint32_t n = t.n_children[parent];
  double total = 0;
  for (uint32_t i = 0; i < n; ++i) total += t.visits[base + i];
  if (t.vloss[parent] > 1) total += (double)(t.vloss[parent] - 1);
  const float spc = (float)std::max(std::sqrt(total), 1.0);
  const uint32_t pv = t.visits[parent];
  const double default_mean = pv ? -(double)(t.w[parent] / (float)pv) : 0.0;
  if (n > 1 && t.move[base + n - 1] == ugo::PASS) --n;
  else if (n == 1) return base;
  double best_v = 0;
  uint32_t best = 0;
  for (uint32_t i = 0; i < n; ++i) {
    const uint32_t vis = t.visits[base + i], vl = t.vloss[base + i];
    double q = default_mean;
    if (vis + vl > 0) q = (double)t.w[base + i] / (double)(vis + vl);  // mean * count / (count + virtual losses)
    const double v = q + (double)c_puct * (double)t.prior[base + i] * (double)spc / (1.0 + (double)vis);
    if (i == 0 || v > best_v) { best = i; best_v = v; }
  }
  return base + best;
}

void new_root(Game& g) {
  g.tree.clear();
  g.tree.resize(1);
  g.tree.move[0] = ugo::PASS;
}

// keep the subtree below `child` as the new tree (UctSearchTree::CopySubtree, search/UctSearchTree.cpp:159-225)
void reuse_subtree(Game& g, uint32_t child) {
  Tree& src = g.tree;
  Tree& dst = g.scratch;
  dst.clear();
  dst.resize(1);
  dst.copy_node(0, src, child);
  dst.parent[0] = 0;
  // breadth-first copy keeps every node's children contiguous
  std::vector<uint32_t>& map_src = g.map_src;  // dst index -> src index
  map_src.clear();
  map_src.push_back(child);
  for (size_t di = 0; di < dst.size(); ++di) {
    const uint32_t si = map_src[di];
    const uint32_t nc = src.n_children[si];
    if (nc == 0) continue;
    const uint32_t base = (uint32_t)dst.size();
    const uint32_t sbase = src.first_child[si];
    dst.first_child[di] = base;
    dst.resize(base + nc);
    for (uint32_t i = 0; i < nc; ++i) {
      dst.copy_node(base + i, src, sbase + i);
      dst.parent[base + i] = (uint32_t)di;
      map_src.push_back(sbase + i);
    }
  }
  // Ping-pong: both trees of a game are sized and prefaulted for a full move, so the swap costs nothing and the next
  // search never regrows an array (one copy of the kept subtree per move, not two: every game reaches its move
  // boundary in the same evaluation wave — one leaf per game per wave — so this work arrives in bursts of all games
  // at once, and on an 8-GPU box that burst is memory-bandwidth-bound).
  src.swap(dst);
}

// RemoveVirtualLoss on every node of a finished descent (:847-850): from the leaf up to the root
void remove_virtual_loss(Tree& t, uint32_t leaf) {
  uint32_t cur = leaf;
  for (;;) {
    --t.vloss[cur];
    if (cur == 0) break;
    cur = t.parent[cur];
  }
}

// one descent; returns true if a leaf was submitted (false: terminal leaf, counted as a playout like the
// reference's ExpandAndBackup returning false, search/UctSearch.cpp:478-490)
bool start_playout(Game& g, Shared& sh) {
  g.work = g.root_board;
  uint32_t node = 0;
  const bool vl = g.tree.use_vloss;
  if (vl) ++g.tree.vloss[0];  // DeepUctSearchTree adds a virtual loss to every node of the path (:814-829)
  while (g.tree.n_children[node]) {
    node = vl ? select_child_vloss(g.tree, node, sh.cfg.puct) : select_child(g.tree, node, sh.cfg.puct);
    g.work.play(g.tree.move[node]);
    if (vl) ++g.tree.vloss[node];
  }
  uint16_t moves[UG_NUM_MOVES];
  uint32_t legal[UG_MASK_WORDS];
  const int n = g.work.generate_moves(moves, legal);
  if (n == 0) {
    sh.terminal_leaves.fetch_add(1, std::memory_order_relaxed);
    if (vl) remove_virtual_loss(g.tree, node);
    return false;
  }
  const uint32_t base = (uint32_t)g.tree.size();
  g.tree.first_child[node] = base;
  g.tree.n_children[node] = (uint16_t)n;
  g.tree.resize(base + (size_t)n);  // prior 0 until the evaluation arrives (UctSearchTree::Expand, :675-691)
  for (int i = 0; i < n; ++i) {
    g.tree.parent[base + i] = node;
    g.tree.move[base + i] = moves[i];
  }
  ug_packed_pos pos;
  g.work.pack(&pos);
  const int64_t ticket = ug_queue_submit(sh.q, &pos, legal);
  if (ticket < 0) {
    sh.stop.store(true);  // queue closed or evaluator failed: every worker winds down
    return false;
  }
  g.pend[g.npend].leaf = node;
  g.pend[g.npend].ticket = ticket;
  ++g.npend;
  return true;
}

void finish_playout(Game& g, uint32_t leaf, float value) {
  Tree& t = g.tree;
  const uint32_t base = t.first_child[leaf], nc = t.n_children[leaf];
  // compact priors: the queue returns exactly the children's priors, renormalised over the legal set on the GPU, in
  // ascending move order with pass last — the order generate_moves created the children in
  memcpy(&t.prior[base], g.policy, (size_t)nc * sizeof(float));
  uint32_t cur = leaf;
  float v = value;
  for (;;) {  // BackupTree, search/UctSearch.cpp:1282-1295
    t.visits[cur] += 1;
    t.w[cur] += v;
    v = -v;
    if (cur == 0) break;
    cur = t.parent[cur];
  }
  if (t.use_vloss) remove_virtual_loss(t, leaf);
}

// after playouts_per_move playouts: play the most visited child (first wins ties), keep or drop the tree
void make_move(Game& g, Shared& sh) {
  const Tree& t = g.tree;
  const uint32_t rbase = t.first_child[0], rn = t.n_children[0];
  int move = ugo::PASS;
  uint32_t chosen = 0;
  if (rn) {
    if (g.root_board.move_number - g.first_search_move < sh.cfg.random_opening_moves) {
      // harness option (not in the reference): diversify games with uniformly random legal opening moves
      std::uniform_int_distribution<uint32_t> d(0, rn > 1 ? rn - 2 : 0);  // PASS is last
      chosen = rbase + d(g.rng);
    } else {
      uint32_t best_visits = 0;
      chosen = rbase;
      bool first = true;
      for (uint32_t i = 0; i < rn; ++i) {
        const uint32_t vis = t.visits[rbase + i];
        if (first || vis > best_visits) { best_visits = vis; chosen = rbase + i; first = false; }
      }
    }
    move = t.move[chosen];
  }
  if (sh.cfg.record_dir) {
    // the policy block of the move: zero, then the visit count of every root child (search/UctSearch.cpp:782-801)
    const size_t at = g.visit_policy.size();
    g.visit_policy.resize(at + UG_NUM_MOVES, 0.f);
    for (uint32_t i = 0; i < rn; ++i) g.visit_policy[at + t.move[rbase + i]] = (float)t.visits[rbase + i];
  }
  g.root_board.play(move);
  g.moves.push_back((int16_t)move);
  sh.moves.fetch_add(1, std::memory_order_relaxed);
  sh.nodes.fetch_add(g.tree.size(), std::memory_order_relaxed);
  g.playouts_this_move = 0;
  const int limit = sh.cfg.max_moves > 0 ? sh.cfg.max_moves : 2 * 19 * 19;
  const bool over = g.root_board.passes >= 2 || g.root_board.move_number - g.first_search_move >= limit ||
                    g.root_board.move_number >= 2 * 19 * 19;
  if (!over) {  // a finished game needs no tree for a next search
    if (sh.cfg.reuse_tree && rn && t.n_children[chosen]) reuse_subtree(g, chosen);
    else new_root(g);
  }
  if (over) {
    g.finished = true;
    sh.games_done.fetch_add(1, std::memory_order_relaxed);
    if (sh.cfg.record_dir && !write_game_record(g, sh.cfg)) sh.record_errors.fetch_add(1, std::memory_order_relaxed);
  }
}

// after playouts completed (or at game start): play the move when the budget is used up and nothing is in flight,
// otherwise put leaves in flight up to leaves_per_game.  Terminal leaves (no moves after two passes) cost a playout
// but no evaluation.  Postcondition: the game is finished or has at least one leaf in flight.
void advance(Game& g, Shared& sh) {
  const int max_leaves = sh.cfg.leaves_per_game > 1 ? sh.cfg.leaves_per_game : 1;
  while (!g.finished) {
    if (g.playouts_this_move + g.npend >= sh.cfg.playouts_per_move) {
      if (g.npend > 0) return;  // the move is decided once the outstanding evaluations are back
      make_move(g, sh);
      continue;
    }
    if (g.npend >= max_leaves) return;
    if (start_playout(g, sh)) continue;
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
      bool any = false;
      for (int k = g.npend - 1; k >= 0; --k) {
        float value = 0.f;
        const int rc = ug_queue_poll(sh.q, g.pend[k].ticket, g.policy, &value);
        if (rc == 0) continue;  // not ready yet
        if (rc < 0) { sh.stop.store(true); return; }
        const uint32_t leaf = g.pend[k].leaf;
        g.pend[k] = g.pend[--g.npend];
        finish_playout(g, leaf, value);
        ++g.playouts_this_move;
        sh.playouts.fetch_add(1, std::memory_order_relaxed);
        any = true;
      }
      if (!any) continue;  // nothing came back for this game: look at the next one
      advance(g, sh);
      if (g.finished) --live;
      progressed = true;
    }
    if (!progressed) {
      const auto t0 = std::chrono::steady_clock::now();
      if (hi - lo <= 2) {
        // one or two games on this thread: every playout waits for its own evaluation, so a sleep (50 us and more of
        // timer slack) would be paid per playout — spin briefly instead
        for (int i = 0; i < 2000; ++i) {
#if defined(__x86_64__)
          __builtin_ia32_pause();
#endif
        }
      } else {
        std::this_thread::sleep_for(std::chrono::microseconds(20));
      }
      sh.idle_ns.fetch_add((uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(
                               std::chrono::steady_clock::now() - t0).count(), std::memory_order_relaxed);
    }
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
  if (cfg.leaves_per_game > Game::kMaxLeaves) cfg.leaves_per_game = Game::kMaxLeaves;
  if (cfg.prelude_moves < 0 || (cfg.prelude_moves > 0 && cfg.record_dir)) return -1;
  Shared sh;
  sh.cfg = cfg;
  sh.q = ug_queue_create(h, cfg.max_batch, cfg.timeout_us < 0 ? 0 : (cfg.timeout_us > 0 ? cfg.timeout_us : 200), 4 * cfg.games * (cfg.leaves_per_game > 1 ? cfg.leaves_per_game : 1) + 1024);
  if (!sh.q) return -2;
  ug_queue_set_compact_priors(sh.q, 1);
  std::vector<Game> games((size_t)cfg.games);
  {
    // Tree storage: one expansion per playout, up to 362 children each, plus the subtree carried over by tree
    // reuse.  A game recycles this memory move after move, so in steady state the search never touches a fresh
    // page; the pages are faulted in here, before the clock starts, by the search threads in parallel (first-touch
    // page faults otherwise cost the short benchmark runs about a core per GPU).  UGEVAL_SP_PREFAULT=0 skips it.
    const size_t cap = (size_t)cfg.playouts_per_move * UG_NUM_MOVES * (cfg.reuse_tree ? 5 : 4) / 4;
    const size_t bytes_per_node = 2 * sizeof(float) + 3 * sizeof(uint32_t) + 2 * sizeof(uint16_t);
    const char* pf = getenv("UGEVAL_SP_PREFAULT");
    const bool prefault = !(pf && pf[0] == '0') && (double)cap * 2.0 * bytes_per_node * cfg.games < 64e9;
    const char* th = getenv("UGEVAL_SP_THP");
    const bool huge = !(th && th[0] == '0');
    std::vector<std::thread> init;
    const size_t per_init = ((size_t)cfg.games + cfg.threads - 1) / cfg.threads;
    for (int t = 0; t < cfg.threads; ++t) {
      const size_t lo = (size_t)t * per_init, hi = std::min((size_t)cfg.games, lo + per_init);
      if (lo >= hi) break;
      init.emplace_back([&games, &cfg, lo, hi, cap, prefault, huge] {
        for (size_t i = lo; i < hi; ++i) {
          Game& g = games[i];
          g.rng.seed(cfg.seed * 1000003ull + (uint64_t)i);
          g.index = (int)i;
          g.tree.use_vloss = g.scratch.use_vloss = cfg.leaves_per_game > 1;
          g.tree.reserve(cap, huge);
          if (cfg.reuse_tree) g.scratch.reserve(cap, huge);
          if (prefault) {
            g.tree.resize(cap);
            g.tree.clear();
            if (cfg.reuse_tree) { g.scratch.resize(cap); g.scratch.clear(); }
          }
          new_root(g);
          // mid-game sample: random legal moves before the first search (see ug_selfplay_config.prelude_moves)
          for (int k = 0; k < cfg.prelude_moves && g.root_board.passes < 2; ++k) {
            uint16_t mv[UG_NUM_MOVES];
            uint32_t legal[UG_MASK_WORDS];
            const int n = g.root_board.generate_moves(mv, legal);
            if (n == 0) break;
            // PASS is generated last: only when nothing else is left
            std::uniform_int_distribution<int> d(0, n > 1 ? n - 2 : 0);
            const uint16_t m = mv[d(g.rng)];
            g.root_board.play(m);
            g.moves.push_back((int16_t)m);
          }
          g.first_search_move = g.root_board.move_number;
        }
      });
    }
    for (std::thread& t : init) t.join();
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
  uint64_t qx[11] = {0};
  ug_queue_stats_ex(sh.q, qx, 11);
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
  {
    const double thread_secs = secs * (double)threads.size();
    const double idle = (double)sh.idle_ns.load() * 1e-9;
    stats->search_busy_fraction = thread_secs > 0 ? 1.0 - idle / thread_secs : 0;
    stats->host_us_per_playout = stats->playouts ? (thread_secs - idle) * 1e6 / (double)stats->playouts : 0;
    stats->full_batches = qx[3];
    stats->wave_cut_batches = qx[4];
    stats->max_queue_depth = qx[6];
    stats->gather_sleeps = qx[7];
    stats->gpu_busy_fraction = secs > 0 ? (double)qx[8] * 1e-6 / secs : 0;
    stats->idle_launches = qx[9];
    stats->idle_gap_seconds = (double)qx[10] * 1e-6;
  }
  if (moves_out && n_moves_out) {
    const int cap = (cfg.max_moves > 0 ? cfg.max_moves : 2 * 19 * 19) + cfg.prelude_moves;  // row length of moves_out
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
void ug_pack_planes(const int8_t* planes_chw, ug_packed_pos* out) {
  memset(out, 0, sizeof(*out));
  const uint32_t tp = planes_chw[16 * UG_NUM_POINTS] ? 0u : 1u;  // plane 16 = 1 - to_play
  out->to_play = tp;
  for (int k = 0; k < UG_HISTORY; ++k) {
    uint32_t* mine = tp == 0 ? out->black[k] : out->white[k];      // plane 2k: the side to move, k moves ago
    uint32_t* theirs = tp == 0 ? out->white[k] : out->black[k];    // plane 2k+1: the opponent
    const int8_t* pm = planes_chw + (2 * k) * UG_NUM_POINTS;
    const int8_t* pt = planes_chw + (2 * k + 1) * UG_NUM_POINTS;
    for (int i = 0; i < UG_NUM_POINTS; ++i) {
      if (pm[i]) mine[i >> 5] |= 1u << (i & 31);
      if (pt[i]) theirs[i >> 5] |= 1u << (i & 31);
    }
  }
}
void ug_board_planes(void* b, int8_t* chw_out) {
  ug_packed_pos p;
  static_cast<ugo::Board*>(b)->pack(&p);
  planes_from_packed(p, chw_out);
}
}
