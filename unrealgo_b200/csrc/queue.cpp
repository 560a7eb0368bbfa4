// Q1 — leaf-batching queue.
//
// Replaces the reference's hand-off between search threads and the evaluator:
//   producer: UctSearch::ExpandAndBackup writes EvalBuffer::feature_buf[threadId], raises state_ready and
//             blocks on a per-thread condvar (search/UctSearch.cpp:511-520, :88-91);
//   consumer: NetworkEvalThread spins, always evaluates all num_threads slots, and marks whichever slots
//             are ready *after* the evaluation as evaluated (search/UctSearch.cpp:164-205) — a slot that
//             became ready mid-evaluation gets a stale result, and the batch is capped at MAX_BATCHES=16.
// Here: leaves live in a pool of slots.  A producer takes a free slot (lock-free bounded MPMC free list), fills
// it and pushes its id on a lock-free bounded multi-producer submission queue; the gather thread pops ids in
// arrival order (only *published* leaves, so every ticket gets exactly its own result), flushes on max_batch or
// timeout, and runs double-buffered pinned staging: the H2D copy of batch k+1 and the D2H copy of batch k-1
// run on their own streams while batch k is in the tower.  Results are written back into the slots; the
// consumer frees the slot when it takes the result.  Because a slot is owned by one leaf from submit to
// poll/wait — not by a position in a ring — a producer that is slow to collect a result can never block
// another producer's submit (an earlier ring-of-slots version deadlocked when the ring lapped a parked result).
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstring>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>

#if defined(__linux__)
#include <sys/resource.h>
#include <sys/syscall.h>
#include <unistd.h>
#endif

#include "idqueue.h"
#include "net.h"

namespace {

enum : uint32_t { SLOT_FREE = 0, SLOT_QUEUED = 1, SLOT_READY = 2 };

struct Slot {
  std::atomic<uint32_t> state{SLOT_FREE};
  std::atomic<uint64_t> ticket{0};  // (allocation count << 20) | slot id: stale tickets never match
  uint64_t allocs = 0;              // touched by the owner only
  ug_packed_pos pos;
  uint32_t legal[UG_MASK_WORDS];
  uint32_t has_legal;
  uint32_t n_out;                   // floats of `policy` that carry the result: #legal in compact mode, else 362
  float policy[UG_NUM_MOVES];
  float value;
  int failed;                       // the batch this leaf rode in did not run: policy/value are not results
};

constexpr int kIdBits = 20;

struct Staging {
  ug_packed_pos* h_pos = nullptr;   // pinned
  uint32_t* h_legal = nullptr;      // pinned
  float* h_out = nullptr;           // pinned: policy[n][362] then value[n]
  ug_packed_pos* d_pos = nullptr;
  uint32_t* d_legal = nullptr;
  float* d_policy = nullptr;
  float* d_value = nullptr;
  cudaEvent_t ev_in = nullptr, ev_done = nullptr, ev_out = nullptr;
  cudaEvent_t ev_t0 = nullptr;       // timing pair with ev_done: how long the evaluator stream worked on this batch
  std::vector<uint32_t> ids;         // slot id of every leaf in the batch
  int n = 0;
  bool any_legal = false;
  bool busy = false;
  int n_run = 0;                     // positions the batch was launched with (n rounded up to 8)
  int max_out = UG_NUM_MOVES;        // widest result row of the batch (compact priors: max #legal)
  int rc = 0;                        // evaluator return code of this batch (failures belong to a batch, not the queue)
};

// The gather and completion threads do little work but every millisecond they wait for a CPU is a millisecond the GPU
// may idle; the search threads are throughput work.  Best effort (needs CAP_SYS_NICE; silently skipped otherwise).
inline void prefer_this_thread() {
#if defined(__linux__)
  const pid_t tid = (pid_t)syscall(SYS_gettid);
  (void)setpriority(PRIO_PROCESS, (id_t)tid, -10);
#endif
}

inline void cpu_relax() {
#if defined(__x86_64__)
  __builtin_ia32_pause();
#endif
}

}  // namespace

struct ug_queue {
  ug_eval* eval = nullptr;
  int max_batch = 0;
  int timeout_us = 0;
  uint64_t capacity = 0;
  std::vector<Slot> slots;
  ug::IdQueue free_ids, submitted;
  std::vector<uint32_t> carry;     // gathered but cut from the last batch (gather thread only)
  Staging st[2];
  cudaStream_t s_in = nullptr, s_out = nullptr;
  std::thread gather_thread, complete_thread;
  std::mutex mu;                    // protects inflight, st[].busy, paused
  std::condition_variable cv_inflight, cv_free, cv_results, cv_pause;
  std::deque<int> inflight;         // staging indices in submission order
  std::atomic<bool> quit{false};
  bool pause_req = false, paused = false;
  // compact priors: a leaf submitted with a legal mask gets back its #legal priors in ascending
  // move order instead of a 362-float row with zeros — the head kernel packs them, the D2H copy and the two host copies
  // shrink to the widest row of the batch
  std::atomic<bool> compact{false};
  // the gather thread sleeps instead of spinning while it waits for leaves (4 CPUs per GPU on the 8-GPU boxes): a
  // producer wakes it when the queue holds what it is waiting for
  std::mutex mu_g;
  std::condition_variable cv_g;
  std::atomic<bool> sleeping{false};
  std::atomic<uint64_t> wake_at{0};
  std::atomic<uint64_t> n_full{0}, n_wave_cut{0}, n_sleeps{0}, max_depth{0}, gpu_busy_us{0};
  // Just-in-time flush.  A partial batch only has to leave when the GPU would otherwise run dry: while a batch is still
  // running there is nothing to gain from flushing 300 us after the first leaf — the partial batch would just sit in
  // the stream behind it, and partial batches cost more per position (fixed per-launch work, unfilled waves).  The
  // gather thread therefore waits for a full batch until shortly before the evaluator stream is expected to run out of
  // work: gpu_free_at_ns is that estimate (steady clock), kept from the measured device time per kernel wave.
  // how often a batch was launched while the evaluator stream was already idle, and for how long it had been (host clock)
  std::atomic<uint64_t> n_launch_idle{0}, idle_gap_us{0};
  std::atomic<int64_t> last_done_ns{0};
  std::atomic<int64_t> last_leaf_ns{0};   // when the gather thread last saw a leaf (latency mode spins only while hot)
  std::atomic<int64_t> gpu_free_at_ns{0};
  std::atomic<uint64_t> ema_wave_ns{0};
  static int64_t now_ns() {
    return std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
  }
  int waves_of(int n) const {
    const long per = (long)(eval->sms / 2) * 256;
    return (int)(((long)n * UG_NUM_POINTS + per - 1) / per);
  }
  int64_t estimate_ns(int n) const { return (int64_t)waves_of(n) * (int64_t)ema_wave_ns.load(std::memory_order_relaxed) + 60000; }
  uint64_t depth() const {
    const uint64_t e = submitted.enq.load(std::memory_order_acquire), d = submitted.deq.load(std::memory_order_acquire);
    return e > d ? e - d : 0;
  }
  void wait_for_leaves(uint64_t want, std::chrono::steady_clock::time_point deadline);
  std::atomic<uint64_t> n_batches{0}, n_leaves{0}, n_timeouts{0}, n_failed{0};
  std::atomic<int> error{0};        // set only by a CUDA (context-wide) error: closes the queue

  void gather_loop();
  void complete_loop();
};

void ug_queue::wait_for_leaves(uint64_t want, std::chrono::steady_clock::time_point deadline) {
  // leaves arrive in bursts: a short spin first.  With timeout_us == 0 the caller asked for latency, not batching (a
  // single search thread waiting for every evaluation): spin for ~100 us before giving the CPU away, a sleeping
  // thread costs a wake-up of 50 us and more per leaf
  // (only while leaves keep coming: an engine that is idle between searches must not keep a core spinning)
  const bool hot = timeout_us == 0 && now_ns() - last_leaf_ns.load(std::memory_order_relaxed) < 5000000;
  const int spins = hot ? 40000 : 128;
  for (int i = 0; i < spins; ++i) {
    if (depth() >= want) return;
    cpu_relax();
  }
  std::unique_lock<std::mutex> lk(mu_g);
  wake_at.store(want, std::memory_order_relaxed);
  sleeping.store(true, std::memory_order_seq_cst);
  if (depth() < want && !quit.load(std::memory_order_acquire)) {
    n_sleeps.fetch_add(1, std::memory_order_relaxed);
    cv_g.wait_until(lk, deadline);  // a missed wake-up costs at most the time to the deadline
  }
  sleeping.store(false, std::memory_order_relaxed);
}

void ug_queue::gather_loop() {
  cudaSetDevice(eval->device);
  prefer_this_thread();
  int which = 0;
  using clock = std::chrono::steady_clock;
  while (!quit.load(std::memory_order_acquire)) {
    {
      std::unique_lock<std::mutex> lk(mu);
      if (pause_req) {
        cv_free.wait(lk, [&] { return inflight.empty() || quit.load(); });
        paused = true;
        cv_pause.notify_all();
        cv_pause.wait(lk, [&] { return !pause_req || quit.load(); });
        paused = false;
        continue;
      }
    }
    // wait for the first published leaf (re-checking pause/quit every millisecond)
    if (carry.empty()) {
      uint32_t id;
      if (!submitted.pop(&id)) {
        wait_for_leaves(1, clock::now() + std::chrono::milliseconds(1));
        continue;
      }
      carry.push_back(id);
    }
    // a staging buffer must be free before leaves are copied out of their slots
    Staging& S = st[which];
    {
      std::unique_lock<std::mutex> lk(mu);
      cv_free.wait(lk, [&] { return !S.busy || quit.load(); });
      if (quit.load()) break;
      S.busy = true;
    }
    auto deadline = clock::now() + std::chrono::microseconds(timeout_us);
    if (timeout_us > 0 && ema_wave_ns.load(std::memory_order_relaxed) > 0) {
      // just-in-time: keep collecting until shortly before the stream runs out of work.  The margin covers the staging
      // copy, H2D and launch (~0.2 ms) and, above all, this thread's wake-up on a host whose CPUs are all busy with
      // search threads (measured on 8 GPUs x 4 CPUs: with 0.3 ms the GPUs idled 11 % of the time): a quarter of a
      // batch time, at least 0.5 ms.  A partial batch costs one extra launch; an idle GPU costs the whole gap.
      const int64_t now = now_ns();
      const int64_t margin = std::max<int64_t>(500000, estimate_ns(max_batch) / 4);
      int64_t jit = gpu_free_at_ns.load(std::memory_order_relaxed) - margin;
      if (jit > now + 20000000) jit = now + 20000000;  // an estimate gone wrong must not park the queue
      const auto jit_tp = clock::time_point(std::chrono::duration_cast<clock::duration>(std::chrono::nanoseconds(jit)));
      if (jit_tp > deadline) deadline = jit_tp;
    }
    {
      const uint64_t d = depth() + carry.size();
      uint64_t m = max_depth.load(std::memory_order_relaxed);
      while (d > m && !max_depth.compare_exchange_weak(m, d, std::memory_order_relaxed)) {}
    }
    int n = 0;
    bool timed_out = false;
    S.any_legal = false;
    S.ids.clear();
    size_t from_carry = 0;
    while (n < max_batch) {
      uint32_t id;
      if (from_carry < carry.size()) id = carry[from_carry++];
      else if (!submitted.pop(&id)) {
        if (clock::now() >= deadline) { timed_out = true; break; }
        wait_for_leaves((uint64_t)(max_batch - n), deadline);  // sleeps until the rest of the batch is there
        continue;
      }
      S.ids.push_back(id);
      ++n;
    }
    last_leaf_ns.store(now_ns(), std::memory_order_relaxed);
    carry.erase(carry.begin(), carry.begin() + (long)from_carry);
    if (timed_out) {
      n_timeouts.fetch_add(1, std::memory_order_relaxed);
      // a partial batch is cut back to whole waves of the tower kernel; the leaves beyond open the next batch
      const int keep = ug_eval_efficient_batch(eval, n);
      if (keep < n) n_wave_cut.fetch_add(1, std::memory_order_relaxed);
      carry.insert(carry.begin(), S.ids.begin() + keep, S.ids.end());
      S.ids.resize((size_t)keep);
      n = keep;
    } else {
      n_full.fetch_add(1, std::memory_order_relaxed);
    }
    const bool cmp = compact.load(std::memory_order_relaxed);
    int max_out = cmp ? 1 : UG_NUM_MOVES;
    for (int i = 0; i < n; ++i) {
      const Slot& sl = slots[S.ids[(size_t)i]];
      memcpy(&S.h_pos[i], &sl.pos, sizeof(ug_packed_pos));
      if (sl.has_legal) {
        memcpy(&S.h_legal[(size_t)i * UG_MASK_WORDS], sl.legal, sizeof(sl.legal));
        S.any_legal = true;
      } else {
        memset(&S.h_legal[(size_t)i * UG_MASK_WORDS], 0xff, sizeof(sl.legal));
      }
      if ((int)sl.n_out > max_out) max_out = (int)sl.n_out;
    }
    S.n = n;
    S.max_out = max_out;
    // H2D on the copy-in stream, tower on the evaluator stream, D2H on the copy-out stream
    cudaMemcpyAsync(S.d_pos, S.h_pos, (size_t)n * sizeof(ug_packed_pos), cudaMemcpyHostToDevice, s_in);
    if (S.any_legal)
      cudaMemcpyAsync(S.d_legal, S.h_legal, (size_t)n * UG_MASK_WORDS * 4, cudaMemcpyHostToDevice, s_in);
    cudaEventRecord(S.ev_in, s_in);
    cudaStreamWaitEvent(eval->stream, S.ev_in, 0);
    // Launch on a batch size rounded up to a multiple of 8 (capped): the extra positions are stale staging rows
    // whose outputs are never copied back, and the evaluator's CUDA-graph cache then sees 32 batch sizes per
    // staging buffer instead of 256.
    int n_run = (n + 7) & ~7;
    if (n_run > max_batch) n_run = max_batch;
    int rc;
    cudaEventRecord(S.ev_t0, eval->stream);
    if (!eval->weights) rc = eval->fail(-1, "no network loaded");
    else rc = eval->enqueue(ug_eval::SRC_PACKED, S.d_pos, n_run, S.any_legal ? S.d_legal : nullptr, eval->value_transform,
                            S.d_policy, S.d_value, eval->stream, cmp && S.any_legal);
    S.rc = rc;
    S.n_run = n_run;
    {
      const int64_t now = now_ns();
      const int64_t base = std::max(gpu_free_at_ns.load(std::memory_order_relaxed), now);
      gpu_free_at_ns.store(base + estimate_ns(n_run), std::memory_order_relaxed);
    }
    if (rc != 0 && n_failed.fetch_add(1, std::memory_order_relaxed) == 0)
      fprintf(stderr, "ugeval queue: evaluator failure %d on a batch of %d: %s\n", rc, n_run, ug_last_error(eval));
    cudaEventRecord(S.ev_done, eval->stream);
    cudaStreamWaitEvent(s_out, S.ev_done, 0);
    // policy rows: only the first max_out floats of each (all 362 unless the priors are compact)
    cudaMemcpy2DAsync(S.h_out, (size_t)UG_NUM_MOVES * 4, S.d_policy, (size_t)UG_NUM_MOVES * 4, (size_t)max_out * 4, (size_t)n,
                      cudaMemcpyDeviceToHost, s_out);
    cudaMemcpyAsync(S.h_out + (size_t)max_batch * UG_NUM_MOVES, S.d_value, (size_t)n * 4, cudaMemcpyDeviceToHost, s_out);
    cudaEventRecord(S.ev_out, s_out);
    {
      std::lock_guard<std::mutex> lk(mu);
      if (inflight.empty()) {  // nothing was running or queued: the GPU has been idle since the last completion
        n_launch_idle.fetch_add(1, std::memory_order_relaxed);
        const int64_t since = last_done_ns.load(std::memory_order_relaxed);
        if (since > 0) idle_gap_us.fetch_add((uint64_t)std::max<int64_t>(0, (now_ns() - since) / 1000), std::memory_order_relaxed);
      }
      inflight.push_back(which);
    }
    cv_inflight.notify_one();
    n_batches.fetch_add(1, std::memory_order_relaxed);
    n_leaves.fetch_add((uint64_t)n, std::memory_order_relaxed);
    which ^= 1;
  }
}

void ug_queue::complete_loop() {
  cudaSetDevice(eval->device);
  prefer_this_thread();
  for (;;) {
    int which;
    {
      std::unique_lock<std::mutex> lk(mu);
      cv_inflight.wait(lk, [&] { return !inflight.empty() || quit.load(); });
      if (inflight.empty()) return;  // quit and drained
      which = inflight.front();
    }
    Staging& S = st[which];
    if (timeout_us == 0) {
      // latency mode (one leaf in flight): poll the event for up to ~1 ms before blocking — the blocking wait's wake-up
      // costs 50 us and more per evaluation
      for (int i = 0; i < 200000 && cudaEventQuery(S.ev_out) == cudaErrorNotReady; ++i) cpu_relax();
    }
    const cudaError_t ce = cudaEventSynchronize(S.ev_out);
    // a CUDA error is sticky for the whole context: every later batch fails too, so it closes the queue
    if (ce != cudaSuccess && error.exchange(-1) == 0)
      fprintf(stderr, "ugeval queue: CUDA error while waiting for a batch: %s\n", cudaGetErrorString(ce));
    const int failed = (S.rc != 0 || ce != cudaSuccess) ? 1 : 0;
    if (!failed) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, S.ev_t0, S.ev_done) == cudaSuccess) {
        gpu_busy_us.fetch_add((uint64_t)(ms * 1e3f));
        const uint64_t per_wave = (uint64_t)(ms * 1e6f) / (uint64_t)std::max(waves_of(S.n_run), 1);
        const uint64_t old = ema_wave_ns.load(std::memory_order_relaxed);
        ema_wave_ns.store(old ? (3 * old + per_wave) / 4 : per_wave, std::memory_order_relaxed);
      }
    }
    for (int i = 0; i < S.n; ++i) {
      Slot& sl = slots[S.ids[(size_t)i]];
      if (!failed) {
        memcpy(sl.policy, S.h_out + (size_t)i * UG_NUM_MOVES, (size_t)sl.n_out * 4);
        sl.value = S.h_out[(size_t)max_batch * UG_NUM_MOVES + i];
      }
      sl.failed = failed;
      sl.state.store(SLOT_READY, std::memory_order_release);
    }
    {
      std::lock_guard<std::mutex> lk(mu);
      inflight.pop_front();
      S.busy = false;
      // re-anchor the estimate: the stream is idle now, or has just started on the other staging buffer's batch
      const int64_t now = now_ns();
      last_done_ns.store(now, std::memory_order_relaxed);
      gpu_free_at_ns.store(inflight.empty() ? now : now + estimate_ns(st[inflight.front()].n_run), std::memory_order_relaxed);
    }
    cv_free.notify_all();
    cv_results.notify_all();
  }
}

extern "C" {

ug_queue* ug_queue_create(ug_eval* h, int max_batch, int timeout_us, int capacity) {
  if (!h) return nullptr;
  if (max_batch <= 0 || max_batch > h->max_batch) { h->fail(-1, "queue max_batch exceeds the evaluator's"); return nullptr; }
  uint64_t cap = 1;
  while (cap < (uint64_t)(capacity > 2 * max_batch ? capacity : 2 * max_batch) || cap < 4) cap <<= 1;
  cudaSetDevice(h->device);
  ug_queue* q = new ug_queue();
  q->eval = h;
  q->max_batch = max_batch;
  q->timeout_us = timeout_us < 0 ? 0 : timeout_us;
  if (cap > (1ull << kIdBits)) cap = 1ull << kIdBits;
  q->capacity = cap;
  q->slots = std::vector<Slot>(cap);
  q->free_ids.init(cap);
  q->submitted.init(cap);
  for (uint64_t i = 0; i < cap; ++i) q->free_ids.push((uint32_t)i);
  bool ok = cudaStreamCreateWithFlags(&q->s_in, cudaStreamNonBlocking) == cudaSuccess &&
            cudaStreamCreateWithFlags(&q->s_out, cudaStreamNonBlocking) == cudaSuccess;
  for (int i = 0; i < 2 && ok; ++i) {
    Staging& S = q->st[i];
    const size_t nb = (size_t)max_batch;
    S.ids.reserve(nb);
    ok = cudaMallocHost((void**)&S.h_pos, nb * sizeof(ug_packed_pos)) == cudaSuccess &&
         cudaMallocHost((void**)&S.h_legal, nb * UG_MASK_WORDS * 4) == cudaSuccess &&
         cudaMallocHost((void**)&S.h_out, nb * (UG_NUM_MOVES + 1) * 4) == cudaSuccess &&
         cudaMalloc((void**)&S.d_pos, nb * sizeof(ug_packed_pos)) == cudaSuccess &&
         cudaMalloc((void**)&S.d_legal, nb * UG_MASK_WORDS * 4) == cudaSuccess &&
         cudaMalloc((void**)&S.d_policy, nb * UG_NUM_MOVES * 4) == cudaSuccess &&
         cudaMalloc((void**)&S.d_value, nb * 4) == cudaSuccess &&
         cudaEventCreateWithFlags(&S.ev_in, cudaEventDisableTiming) == cudaSuccess &&
         cudaEventCreate(&S.ev_done) == cudaSuccess && cudaEventCreate(&S.ev_t0) == cudaSuccess &&
         // the completion thread sleeps on this one instead of spinning a core away (4 CPUs per GPU on 8-GPU boxes)
         cudaEventCreateWithFlags(&S.ev_out, cudaEventDisableTiming | cudaEventBlockingSync) == cudaSuccess;
  }
  for (int i = 0; i < 2 && ok; ++i) {
    // batches are launched on sizes rounded up to 8: the rows past n must be defined data (an all-zero position, every
    // move legal), not whatever cudaMalloc returned — a zero legal sum would put NaNs in rows nobody reads
    ok = cudaMemset(q->st[i].d_pos, 0, (size_t)max_batch * sizeof(ug_packed_pos)) == cudaSuccess &&
         cudaMemset(q->st[i].d_legal, 0xff, (size_t)max_batch * UG_MASK_WORDS * 4) == cudaSuccess;
  }
  if (!ok) {
    h->fail(-1, "queue: cannot allocate staging buffers");
    ug_queue_destroy(q);
    return nullptr;
  }
  q->gather_thread = std::thread([q] { q->gather_loop(); });
  q->complete_thread = std::thread([q] { q->complete_loop(); });
  return q;
}

void ug_queue_destroy(ug_queue* q) {
  if (!q) return;
  q->quit.store(true, std::memory_order_release);
  {
    std::lock_guard<std::mutex> lk(q->mu);
    q->pause_req = false;
  }
  q->cv_free.notify_all();
  q->cv_pause.notify_all();
  { std::lock_guard<std::mutex> lk(q->mu_g); q->cv_g.notify_all(); }
  if (q->gather_thread.joinable()) q->gather_thread.join();
  q->cv_inflight.notify_all();
  if (q->complete_thread.joinable()) q->complete_thread.join();
  q->cv_results.notify_all();
  cudaSetDevice(q->eval->device);
  for (int i = 0; i < 2; ++i) {
    Staging& S = q->st[i];
    if (S.h_pos) cudaFreeHost(S.h_pos);
    if (S.h_legal) cudaFreeHost(S.h_legal);
    if (S.h_out) cudaFreeHost(S.h_out);
    if (S.d_pos) cudaFree(S.d_pos);
    if (S.d_legal) cudaFree(S.d_legal);
    if (S.d_policy) cudaFree(S.d_policy);
    if (S.d_value) cudaFree(S.d_value);
    if (S.ev_in) cudaEventDestroy(S.ev_in);
    if (S.ev_done) cudaEventDestroy(S.ev_done);
    if (S.ev_t0) cudaEventDestroy(S.ev_t0);
    if (S.ev_out) cudaEventDestroy(S.ev_out);
  }
  if (q->s_in) cudaStreamDestroy(q->s_in);
  if (q->s_out) cudaStreamDestroy(q->s_out);
  delete q;
}

int64_t ug_queue_submit(ug_queue* q, const ug_packed_pos* pos, const uint32_t* legal) {
  if (!q || !pos) return -1;
  uint32_t id;
  int spins = 0;
  // a free slot: there is always one unless more leaves are outstanding than the pool holds
  while (!q->free_ids.pop(&id)) {
    if (q->quit.load(std::memory_order_acquire) || q->error.load(std::memory_order_relaxed)) return -1;
    if (++spins < 256) cpu_relax(); else std::this_thread::yield();
  }
  Slot& sl = q->slots[id];
  memcpy(&sl.pos, pos, sizeof(ug_packed_pos));
  sl.has_legal = legal ? 1u : 0u;
  sl.n_out = UG_NUM_MOVES;
  if (legal) {
    memcpy(sl.legal, legal, sizeof(sl.legal));
    if (q->compact.load(std::memory_order_relaxed)) {
      uint32_t c = 0;
      for (int w = 0; w < UG_MASK_WORDS; ++w) c += (uint32_t)__builtin_popcount(w == UG_MASK_WORDS - 1 ? legal[w] & 0x3ffu : legal[w]);
      sl.n_out = c;   // bits 0..361 only: word 11 holds moves 352..361
    }
  }
  const uint64_t ticket = (++sl.allocs << kIdBits) | id;
  sl.ticket.store(ticket, std::memory_order_relaxed);
  sl.state.store(SLOT_QUEUED, std::memory_order_release);
  q->submitted.push(id);  // cannot be full: at most `capacity` ids exist
  if (q->sleeping.load(std::memory_order_seq_cst) && q->depth() >= q->wake_at.load(std::memory_order_relaxed)) {
    std::lock_guard<std::mutex> lk(q->mu_g);
    q->cv_g.notify_one();
  }
  return (int64_t)ticket;
}

static inline Slot* slot_of(ug_queue* q, int64_t ticket) {
  const uint64_t id = (uint64_t)ticket & ((1ull << kIdBits) - 1);
  return id < q->capacity ? &q->slots[id] : nullptr;
}

// result ready for exactly this ticket?  (a consumed or foreign ticket never matches)
static inline bool ready(const Slot& sl, int64_t ticket) {
  return sl.state.load(std::memory_order_acquire) == SLOT_READY &&
         sl.ticket.load(std::memory_order_relaxed) == (uint64_t)ticket;
}

// copies the result out (nothing on a failed batch: outputs stay untouched) and frees the slot; returns the slot's
// failure flag.  The ticket is cleared so that waiting on it again reports "not outstanding" instead of spinning.
static inline int take(ug_queue* q, Slot& sl, int64_t ticket, float* policy, float* value) {
  const int failed = sl.failed;
  if (!failed) {
    if (policy) memcpy(policy, sl.policy, (size_t)sl.n_out * 4);
    if (value) *value = sl.value;
  }
  sl.ticket.store(0, std::memory_order_relaxed);
  sl.state.store(SLOT_FREE, std::memory_order_release);
  q->free_ids.push((uint32_t)((uint64_t)ticket & ((1ull << kIdBits) - 1)));
  return failed;
}

int ug_queue_wait(ug_queue* q, int64_t ticket, float* policy, float* value) {
  if (!q || ticket < 0) return -1;
  Slot* sl = slot_of(q, ticket);
  if (!sl) return -1;
  int spins = 0;
  while (!ready(*sl, ticket)) {
    if (q->quit.load(std::memory_order_acquire)) return -1;
    if (sl->ticket.load(std::memory_order_relaxed) != (uint64_t)ticket) return -1;  // not an outstanding ticket
    if (++spins < 2000) { cpu_relax(); continue; }
    std::unique_lock<std::mutex> lk(q->mu);
    q->cv_results.wait_for(lk, std::chrono::microseconds(200));
  }
  return (take(q, *sl, ticket, policy, value) || q->error.load()) ? -1 : 0;
}

int ug_queue_poll(ug_queue* q, int64_t ticket, float* policy, float* value) {
  if (!q || ticket < 0) return -1;
  Slot* sl = slot_of(q, ticket);
  if (!sl) return -1;
  if (!ready(*sl, ticket)) return q->quit.load(std::memory_order_acquire) ? -1 : 0;
  return (take(q, *sl, ticket, policy, value) || q->error.load()) ? -1 : 1;
}

int ug_queue_stats(ug_queue* q, uint64_t* batches, uint64_t* leaves, uint64_t* flush_timeouts) {
  if (!q) return -1;
  if (batches) *batches = q->n_batches.load();
  if (leaves) *leaves = q->n_leaves.load();
  if (flush_timeouts) *flush_timeouts = q->n_timeouts.load();
  return 0;
}

int ug_queue_set_compact_priors(ug_queue* q, int on) {
  if (!q) return -1;
  q->compact.store(on != 0, std::memory_order_relaxed);
  return 0;
}

int ug_queue_stats_ex(ug_queue* q, uint64_t* out, int count) {
  if (!q || !out) return -1;
  const uint64_t v[11] = {q->n_batches.load(), q->n_leaves.load(), q->n_timeouts.load(), q->n_full.load(),
                          q->n_wave_cut.load(), q->n_failed.load(), q->max_depth.load(), q->n_sleeps.load(),
                          q->gpu_busy_us.load(), q->n_launch_idle.load(), q->idle_gap_us.load()};
  for (int i = 0; i < count && i < 11; ++i) out[i] = v[i];
  return 0;
}

int ug_queue_swap_checkpoint(ug_queue* q, const char* prefix) {
  if (!q) return -1;
  {
    std::unique_lock<std::mutex> lk(q->mu);
    q->pause_req = true;
    q->cv_free.notify_all();
    q->cv_pause.wait(lk, [&] { return q->paused; });
  }
  cudaSetDevice(q->eval->device);
  const int rc = ug_eval_load_checkpoint(q->eval, prefix);
  {
    std::lock_guard<std::mutex> lk(q->mu);
    q->pause_req = false;
  }
  q->cv_pause.notify_all();
  return rc;
}

}  // extern "C"
