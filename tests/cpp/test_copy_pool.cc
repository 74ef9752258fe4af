// tests/cpp/test_copy_pool.cc -- CPU stress test of eigencuda_b200/csrc/copy_pool.hpp: every item
// of every job runs exactly once whatever the timing (back-to-back jobs that catch the workers
// polling, jobs after a pause that finds them asleep, jobs of one item, jobs with more items than
// threads), and staging_copy is a byte-exact memcpy for every alignment and length class.
// Built and run by tests/test_host.py (no GPU needed).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>

#include "../../eigencuda_b200/csrc/copy_pool.hpp"

static int g_failed = 0;
#define EXPECT(cond)                                              \
  do {                                                            \
    if (!(cond)) {                                                \
      std::printf("FAIL line %d: %s\n", __LINE__, #cond);         \
      ++g_failed;                                                 \
    }                                                             \
  } while (0)

static void each_item_once(int workers, int jobs, bool pauses) {
  ec_host::CopyPool pool(workers);
  std::vector<std::atomic<int>> hits(4096);
  unsigned seed = 12345u + workers;
  for (int j = 0; j < jobs; ++j) {
    seed = seed * 1664525u + 1013904223u;
    const int n = 1 + (int)((seed >> 8) % (j % 7 == 0 ? 4096 : 9));
    for (int i = 0; i < n; ++i) hits[i].store(0);
    pool.parallel_for(n, [&](int64_t i) { hits[i].fetch_add(1); });
    bool ok = true;
    for (int i = 0; i < n; ++i) ok = ok && hits[i].load() == 1;
    EXPECT(ok);
    if (!ok) return;
    if (pauses && j % 50 == 0) std::this_thread::sleep_for(std::chrono::microseconds(400));
  }
}

static void copies_are_exact() {
  const size_t cap = (size_t)3 << 20;
  std::vector<unsigned char> src(cap + 64), dst(cap + 128);
  for (size_t i = 0; i < src.size(); ++i) src[i] = (unsigned char)(i * 131u + (i >> 9));
  const size_t lengths[] = {0, 1, 31, 4096, (64 << 10) - 1, 64 << 10, (64 << 10) + 33, (1 << 20) + 127, cap};
  for (size_t len : lengths)
    for (size_t so : {0, 1, 24})
      for (size_t d_o : {0, 8, 31}) {
        std::fill(dst.begin(), dst.end(), 0xEE);
        ec_host::staging_copy(dst.data() + d_o, src.data() + so, len);
        bool ok = std::memcmp(dst.data() + d_o, src.data() + so, len) == 0;
        ok = ok && dst[d_o + len] == 0xEE && (d_o == 0 || dst[d_o - 1] == 0xEE);
        EXPECT(ok);
      }
  // one copy cut over the pool, as compat_copy does
  ec_host::CopyPool pool(3);
  std::fill(dst.begin(), dst.end(), 0);
  const size_t piece = 600 << 10, pieces = (cap + piece - 1) / piece;
  pool.parallel_for((int64_t)pieces, [&](int64_t i) {
    const size_t lo = (size_t)i * piece, hi = std::min(cap, lo + piece);
    ec_host::staging_copy(dst.data() + lo, src.data() + lo, hi - lo);
  });
  EXPECT(std::memcmp(dst.data(), src.data(), cap) == 0);
}

int main() {
  each_item_once(0, 200, false);   // no workers: the caller does everything
  each_item_once(1, 4000, true);
  each_item_once(3, 20000, true);
  each_item_once(7, 20000, false);
  copies_are_exact();
  if (g_failed == 0) std::printf("ALL PASSED\n");
  return g_failed == 0 ? 0 : 1;
}
