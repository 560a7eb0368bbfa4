// CPU stress test of the lock-free id queue behind the leaf queue (unrealgo_b200/csrc/idqueue.h), in the two roles it
// plays there: (1) free list + submission queue cycling a pool of slot ids between many producers and one consumer,
// every id delivered exactly once per cycle and in per-producer FIFO order; (2) full / empty behaviour.
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>

#include "../../unrealgo_b200/csrc/idqueue.h"

int main(int argc, char** argv) {
  const int producers = argc > 1 ? atoi(argv[1]) : 16;
  const int per_producer = argc > 2 ? atoi(argv[2]) : 200000;
  const uint64_t cap = 1024;
  ug::IdQueue free_ids, submitted;
  free_ids.init(cap);
  submitted.init(cap);
  uint32_t tmp;
  if (submitted.pop(&tmp)) { printf("FAIL pop from empty\n"); return 1; }
  for (uint64_t i = 0; i < cap; ++i)
    if (!free_ids.push((uint32_t)i)) { printf("FAIL push into non-full\n"); return 1; }
  if (free_ids.push(7)) { printf("FAIL push into full\n"); return 1; }

  // payload written by the producer that owns a slot, checked by the consumer: (producer << 32) | sequence
  std::vector<uint64_t> payload(cap, 0);
  std::atomic<bool> failed{false};
  std::atomic<uint64_t> consumed{0};
  const uint64_t total = (uint64_t)producers * per_producer;
  std::thread consumer([&] {
    std::vector<uint64_t> next((size_t)producers, 0);
    uint32_t id;
    while (consumed.load(std::memory_order_relaxed) < total && !failed.load()) {
      if (!submitted.pop(&id)) continue;
      if (id >= cap) { failed.store(true); break; }
      const uint64_t v = payload[id];
      const uint64_t p = v >> 32, seq = v & 0xffffffffu;
      if (p >= (uint64_t)producers || seq != next[p]) { failed.store(true); break; }  // exactly once, FIFO per producer
      ++next[p];
      consumed.fetch_add(1, std::memory_order_relaxed);
      while (!free_ids.push(id)) {}  // hand the slot back
    }
  });
  std::vector<std::thread> th;
  for (int p = 0; p < producers; ++p)
    th.emplace_back([&, p] {
      uint32_t id;
      for (int s = 0; s < per_producer && !failed.load(); ++s) {
        while (!free_ids.pop(&id)) { if (failed.load()) return; }
        payload[id] = ((uint64_t)p << 32) | (uint64_t)s;
        while (!submitted.push(id)) {}
      }
    });
  for (auto& t : th) t.join();
  consumer.join();
  if (failed.load() || consumed.load() != total) { printf("FAIL consumed %llu of %llu\n", (unsigned long long)consumed.load(), (unsigned long long)total); return 1; }
  // every slot id is back in the free list exactly once
  std::vector<int> seen(cap, 0);
  uint64_t n = 0;
  while (free_ids.pop(&tmp)) { ++seen[tmp]; ++n; }
  for (uint64_t i = 0; i < cap; ++i) if (seen[i] != 1) { printf("FAIL slot %llu seen %d times\n", (unsigned long long)i, seen[i]); return 1; }
  printf("OK %llu ids through %d producers, %llu slots intact\n", (unsigned long long)total, producers, (unsigned long long)n);
  return 0;
}
