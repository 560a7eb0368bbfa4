// Bounded lock-free MPMC queue of 32-bit ids, used twice by the leaf queue (queue.cpp): as the free list of leaf
// slots and as the submission queue that carries slot ids from the search threads to the gather thread.
#pragma once
#include <atomic>
#include <cstdint>
#include <vector>

namespace ug {

// Bounded MPMC queue of slot ids (D. Vyukov's array queue): per-cell sequence numbers, one CAS per operation.
struct IdQueue {
  struct Cell { std::atomic<uint64_t> seq; uint32_t id; };
  std::vector<Cell> cells;
  uint64_t mask = 0;
  alignas(64) std::atomic<uint64_t> enq{0};
  alignas(64) std::atomic<uint64_t> deq{0};
  void init(uint64_t cap) {
    cells = std::vector<Cell>(cap);
    mask = cap - 1;
    for (uint64_t i = 0; i < cap; ++i) cells[i].seq.store(i, std::memory_order_relaxed);
  }
  bool push(uint32_t id) {
    uint64_t pos = enq.load(std::memory_order_relaxed);
    for (;;) {
      Cell& c = cells[pos & mask];
      const int64_t dif = (int64_t)c.seq.load(std::memory_order_acquire) - (int64_t)pos;
      if (dif == 0) {
        if (enq.compare_exchange_weak(pos, pos + 1, std::memory_order_relaxed)) {
          c.id = id;
          c.seq.store(pos + 1, std::memory_order_release);
          return true;
        }
      } else if (dif < 0) {
        return false;  // full
      } else {
        pos = enq.load(std::memory_order_relaxed);
      }
    }
  }
  bool pop(uint32_t* id) {
    uint64_t pos = deq.load(std::memory_order_relaxed);
    for (;;) {
      Cell& c = cells[pos & mask];
      const int64_t dif = (int64_t)c.seq.load(std::memory_order_acquire) - (int64_t)(pos + 1);
      if (dif == 0) {
        if (deq.compare_exchange_weak(pos, pos + 1, std::memory_order_relaxed)) {
          *id = c.id;
          c.seq.store(pos + mask + 1, std::memory_order_release);
          return true;
        }
      } else if (dif < 0) {
        return false;  // empty (or the next cell is claimed but not yet published)
      } else {
        pos = deq.load(std::memory_order_relaxed);
      }
    }
  }
};

}  // namespace ug
