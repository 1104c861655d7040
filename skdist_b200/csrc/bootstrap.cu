// bootstrap.cu -- host-only: the bootstrap multiplicities and splitter seeds of a forest's trees.
//
// Per tree seed s the reference's task (ref ensemble.py:51-55, 68-109) draws
//   indices = RandomState(s).randint(0, n, n);  sample_weight = bincount(indices, minlength=n)
// and scikit-learn's splitter takes  rand_r_state = RandomState(s).randint(0, 2^31 - 1)
// (SK/tree/_splitter.pyx:155).  numpy runs this one tree at a time under the GIL (~29 ms per tree at
// n = 2M: the host side was the bottleneck of the forest fit); here the trees are spread over host
// threads.  Restated bit for bit:
//   * RandomState(int seed) -> MT19937 init_genrand(seed)            (numpy/random/_mt19937.pyx _legacy_seeding)
//   * randint(0, n) for n - 1 < 2^32 - 1 -> masked rejection on 32-bit outputs:
//     do v = genrand_uint32() & mask; while (v > n - 1)               (numpy distributions.c
//     random_bounded_uint64_fill -> buffered_bounded_masked_uint32, legacy use_masked = True)
// No CUDA in this file: the function runs (and is tested) on machines without a GPU.
#include <stdint.h>

#include <algorithm>
#include <atomic>
#include <thread>
#include <vector>

namespace {

struct Mt19937 {
  uint32_t mt[624];
  int idx;
  explicit Mt19937(uint32_t seed) {
    mt[0] = seed;
    for (int i = 1; i < 624; ++i) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i;
    idx = 624;
  }
  static inline uint32_t twist(uint32_t u, uint32_t v) {
    const uint32_t y = (u & 0x80000000u) | (v & 0x7fffffffu);
    return (y >> 1) ^ ((v & 1u) ? 0x9908b0dfu : 0u);
  }
  void refill() {
    int k = 0;
    for (; k < 624 - 397; ++k) mt[k] = mt[k + 397] ^ twist(mt[k], mt[k + 1]);
    for (; k < 623; ++k) mt[k] = mt[k + 397 - 624] ^ twist(mt[k], mt[k + 1]);
    mt[623] = mt[396] ^ twist(mt[623], mt[0]);
    idx = 0;
  }
  inline uint32_t next() {
    if (idx >= 624) refill();
    uint32_t y = mt[idx++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
  }
};

inline uint32_t mask_of(uint32_t rng) {   // smallest 2^k - 1 >= rng
  uint32_t m = rng;
  m |= m >> 1; m |= m >> 2; m |= m >> 4; m |= m >> 8; m |= m >> 16;
  return m;
}

}  // namespace

extern "C" int skd_bootstrap_counts(int32_t n_trees, const uint32_t* seeds, int64_t n, int32_t bootstrap,
                                    uint8_t* counts_out, uint32_t* rand_r_out, int32_t n_threads) {
  if (n_trees <= 0 || !seeds || !rand_r_out || n <= 0 || n > 0xFFFFFFFELL || (bootstrap && !counts_out)) return 2;
  unsigned hw = std::thread::hardware_concurrency();
  int nt = n_threads > 0 ? n_threads : (int)(hw ? hw : 8);
  if (nt > n_trees) nt = n_trees;
  if (nt > 64) nt = 64;
  std::atomic<int> next_tree{0};
  std::atomic<int> overflow{0};
  auto work = [&]() {
    for (;;) {
      const int t = next_tree.fetch_add(1);
      if (t >= n_trees) break;
      {   // splitter seed: a fresh generator, one draw from [0, 2^31 - 2]
        Mt19937 g(seeds[t]);
        const uint32_t rng = 2147483646u, mask = 0x7fffffffu;
        uint32_t v;
        do { v = g.next() & mask; } while (v > rng);
        rand_r_out[t] = v;
      }
      if (!bootstrap) continue;
      Mt19937 g(seeds[t]);
      const uint32_t rng = (uint32_t)(n - 1), mask = mask_of(rng);
      // counted in place in the caller's uint8 row (half the cache footprint of wider counters); the
      // scatter is latency-bound, so the draws are made in blocks whose targets are prefetched first
      uint8_t* cnt = counts_out + (size_t)t * (size_t)n;
      std::fill(cnt, cnt + n, (uint8_t)0);
      bool over = false;
      if (rng == 0) {   // numpy consumes no random number for a single value
        over = n > 255;
        cnt[0] = (uint8_t)(n > 255 ? 255 : n);
      } else {
        constexpr int BLK = 32;
        uint32_t v[BLK];
        int64_t i = 0;
        for (; i + BLK <= n; i += BLK) {
          for (int k = 0; k < BLK; ++k) {
            uint32_t x;
            do { x = g.next() & mask; } while (x > rng);
            v[k] = x;
            __builtin_prefetch(cnt + x, 1, 1);
          }
          for (int k = 0; k < BLK; ++k) {
            uint8_t& c = cnt[v[k]];
            if (c == 255) over = true; else c += 1;
          }
        }
        for (; i < n; ++i) {
          uint32_t x;
          do { x = g.next() & mask; } while (x > rng);
          uint8_t& c = cnt[x];
          if (c == 255) over = true; else c += 1;
        }
      }
      if (over) overflow.store(1);
    }
  };
  std::vector<std::thread> pool;
  for (int i = 1; i < nt; ++i) pool.emplace_back(work);
  work();
  for (auto& th : pool) th.join();
  return overflow.load() ? 1 : 0;
}
