// eigencuda_b200/csrc/copy_pool.hpp -- host-side staging copies between pageable and pinned memory:
// a streaming-store memcpy and the small worker pool that spreads one copy over a few cores.
// Plain C++ with no CUDA in it, so tests/cpp/test_copy_pool.cc exercises it on the CPU.
// Used by the blocking transfers of CudaMatrix (reference src/cudamatrix.cc:8-12, :47-51) and by
// the tensor-stream engine.
#pragma once
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cerrno>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>
#if defined(__x86_64__)
#include <immintrin.h>
#endif
#if defined(__linux__)
#include <sys/mman.h>
#endif

namespace ec_host {

// Staging copy with streaming (non-temporal) stores.  Neither direction of the pageable <-> pinned
// staging re-reads what it has just written, so ordinary stores only add a read-for-ownership of
// every destination line: on this memcpy-bound path that is a third of the host memory traffic.
// AVX2 when the CPU has it (checked once), std::memcpy otherwise and for small pieces.
#if defined(__x86_64__)
__attribute__((target("avx2"))) inline void stream_copy_avx2(char *dst, const char *src, size_t n) {
  // head: bring dst to 32-byte alignment
  const size_t head = (32 - ((uintptr_t)dst & 31)) & 31;
  if (head) {
    std::memcpy(dst, src, head);
    dst += head; src += head; n -= head;
  }
  size_t i = 0;
  for (; i + 128 <= n; i += 128) {
    const __m256i a = _mm256_loadu_si256((const __m256i *)(src + i));
    const __m256i b = _mm256_loadu_si256((const __m256i *)(src + i + 32));
    const __m256i c = _mm256_loadu_si256((const __m256i *)(src + i + 64));
    const __m256i d = _mm256_loadu_si256((const __m256i *)(src + i + 96));
    _mm256_stream_si256((__m256i *)(dst + i), a);
    _mm256_stream_si256((__m256i *)(dst + i + 32), b);
    _mm256_stream_si256((__m256i *)(dst + i + 64), c);
    _mm256_stream_si256((__m256i *)(dst + i + 96), d);
  }
  _mm_sfence();
  if (i < n) std::memcpy(dst + i, src + i, n - i);
}
#endif

inline void staging_copy(void *dst, const void *src, size_t n) {
#if defined(__x86_64__)
  static const bool have_avx2 = __builtin_cpu_supports("avx2");
  if (have_avx2 && n >= (size_t)64 << 10) {
    stream_copy_avx2((char *)dst, (const char *)src, n);
    return;
  }
#endif
  std::memcpy(dst, src, n);
}

// Ask the kernel to map a freshly allocated destination in one call (MADV_POPULATE_WRITE, Linux
// 5.14+) instead of taking a page fault every 4 KiB inside the copy: the results of the
// reference's loop land in new matrices every iteration, and in a VM each of those faults costs
// about as much as copying the page.  Done while the DMA is still in flight.  Also keeps the
// parallel copy from serialising on the address-space lock.  A kernel without the call is noticed
// once and never asked again; EC_PREFAULT=0 turns it off.
inline void prefault_for_write(void *p, size_t n) {
#if defined(__linux__)
  static std::atomic<int> usable{[] {
    const char *e = std::getenv("EC_PREFAULT");
    return (e && e[0] == '0') ? 0 : 1;
  }()};
  if (!usable.load(std::memory_order_relaxed)) return;
  const uintptr_t lo = ((uintptr_t)p + 4095) & ~(uintptr_t)4095, hi = ((uintptr_t)p + n) & ~(uintptr_t)4095;
  if (hi <= lo) return;
#ifndef MADV_POPULATE_WRITE
#define MADV_POPULATE_WRITE 23
#endif
  // Recycled heap memory is already mapped and walking its page tables would cost a third of the
  // copy: look at the first and the last page and leave if both are resident.
  unsigned char first = 0, last = 0;
  if (mincore((void *)lo, 4096, &first) == 0 && mincore((void *)(hi - 4096), 4096, &last) == 0 &&
      (first & 1) && (last & 1))
    return;
  if (madvise((void *)lo, hi - lo, MADV_POPULATE_WRITE) != 0 && errno == EINVAL)
    usable.store(0, std::memory_order_relaxed);
#else
  (void)p; (void)n;
#endif
}

// Small persistent worker pool for pageable <-> pinned staging copies.  One caller at a time.
// Work is handed out through ONE packed atomic word, generation | item count | next item, so a
// worker that draws an index also learns -- from the same atomic read -- which job it belongs to
// and whether it is in range; a late worker can therefore never take an item of a job it did not
// see start.  Workers poll that word for a short while after each job before they go to sleep: the
// reference's loop hands over a 2 MiB copy every hundred microseconds or so, and a futex wake-up
// costs about as much as the copy.
class CopyPool {
 public:
  // spin_us: how long an idle worker keeps polling before it sleeps.  Should cover the gap between
  // two jobs of a steady caller: once workers start sleeping every job pays a wake-up, the caller's
  // loop gets slower, and the gaps stay long (measured as two stable speeds a factor 1.8 apart).
  explicit CopyPool(int n, int spin_us = 150) : spin_us_(spin_us) {
    for (int i = 0; i < n; ++i) workers_.emplace_back([this] { run(); });
  }
  ~CopyPool() {
    stop_.store(true);
    {
      std::lock_guard<std::mutex> lk(mu_);
    }
    cv_.notify_all();
    for (auto &t : workers_) t.join();
  }
  int size() const { return (int)workers_.size(); }
  // Run fn(i) for i in [0, n) on the workers AND the calling thread; returns when all are done.
  void parallel_for(int64_t n, const std::function<void(int64_t)> &fn) {
    if (n <= 0) return;
    if (workers_.empty() || n == 1) {
      for (int64_t i = 0; i < n; ++i) fn(i);
      return;
    }
    for (int64_t base = 0; base < n; base += MAX_ITEMS) {   // far more items than any real call has
      const int64_t count = std::min<int64_t>(MAX_ITEMS, n - base);
      fn_ = &fn;
      base_ = base;
      pending_.store(count, std::memory_order_relaxed);
      const uint64_t gen = ((word_.load(std::memory_order_relaxed) >> GEN_SHIFT) + 1) & GEN_MASK;
      word_.store((gen << GEN_SHIFT) | ((uint64_t)count << COUNT_SHIFT), std::memory_order_seq_cst);
      if (sleepers_.load(std::memory_order_seq_cst) > 0) {
        {
          std::lock_guard<std::mutex> lk(mu_);
        }
        cv_.notify_all();
      }
      drain();
      while (pending_.load(std::memory_order_acquire) != 0) cpu_relax();
    }
  }

 private:
  static constexpr int GEN_SHIFT = 48, COUNT_SHIFT = 24;
  static constexpr uint64_t GEN_MASK = 0xFFFF, FIELD_MASK = 0xFFFFFF;
  static constexpr int64_t MAX_ITEMS = (int64_t)FIELD_MASK - 4096;   // head room for overshoot
  static void cpu_relax() {
#if defined(__x86_64__)
    __builtin_ia32_pause();
#endif
  }
  // take items of the current job until it has none left
  void drain() {
    for (;;) {
      const uint64_t v = word_.fetch_add(1, std::memory_order_acq_rel);
      const uint64_t idx = v & FIELD_MASK, count = (v >> COUNT_SHIFT) & FIELD_MASK;
      if (idx >= count) return;
      (*fn_)(base_ + (int64_t)idx);
      pending_.fetch_sub(1, std::memory_order_release);
    }
  }
  void run() {
    uint64_t seen = 0;   // generation of the last job this worker looked at
    for (;;) {
      // poll for spin_us_, then sleep until the generation moves
      bool fresh = false;
      const auto t0 = std::chrono::steady_clock::now();
      for (int spin = 0;; ++spin) {
        if (stop_.load(std::memory_order_relaxed)) return;
        if ((word_.load(std::memory_order_acquire) >> GEN_SHIFT) != seen) {
          fresh = true;
          break;
        }
        cpu_relax();
        if ((spin & 255) == 255 &&
            std::chrono::steady_clock::now() - t0 > std::chrono::microseconds(spin_us_))
          break;
      }
      if (!fresh) {
        std::unique_lock<std::mutex> lk(mu_);
        sleepers_.fetch_add(1, std::memory_order_seq_cst);
        cv_.wait(lk, [&] {
          return stop_.load() || (word_.load(std::memory_order_seq_cst) >> GEN_SHIFT) != seen;
        });
        sleepers_.fetch_sub(1, std::memory_order_seq_cst);
        if (stop_.load()) return;
      }
      seen = word_.load(std::memory_order_acquire) >> GEN_SHIFT;
      drain();
    }
  }
  const int spin_us_;
  std::vector<std::thread> workers_;
  std::mutex mu_;
  std::condition_variable cv_;
  std::atomic<bool> stop_{false};
  std::atomic<int> sleepers_{0};
  std::atomic<uint64_t> word_{0};
  std::atomic<int64_t> pending_{0};
  const std::function<void(int64_t)> *fn_ = nullptr;
  int64_t base_ = 0;
};

}  // namespace ec_host
