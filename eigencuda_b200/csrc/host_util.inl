// eigencuda_b200/csrc/host_util.inl -- error reporting, driver entry point, pinned staging ring, pointer queries, copy worker pool
// Part of the single translation unit ec_api.cu (included there, in order); not a standalone header.
#pragma once

// ------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------
namespace {

thread_local std::string g_last_error;

int fail(int code, const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}

#define EC_CUDA(call)                                                                        \
  do {                                                                                       \
    cudaError_t _e = (call);                                                                 \
    if (_e != cudaSuccess)                                                                   \
      return fail(EC_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e),       \
                  __FILE__, __LINE__);                                                       \
  } while (0)

// ------------------------------------------------------------------------------------------
// driver entry point for tensor-map encoding (no link-time dependency on libcuda, so the
// library loads -- and its symbols can be checked -- on a machine without a GPU driver)
// ------------------------------------------------------------------------------------------
typedef CUresult (*encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
                                    const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                    const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

encode_tiled_fn get_encode_tiled() {
  static encode_tiled_fn fn = []() -> encode_tiled_fn {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) !=
            cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    return reinterpret_cast<encode_tiled_fn>(p);
  }();
  return fn;
}

// ------------------------------------------------------------------------------------------
// pinned staging ring: lets a pageable source be released to the caller as soon as
// copy_to_gpu returns (the reference's observable contract, src/cudamatrix.cc:47-51) while the
// DMA itself runs asynchronously from pinned memory.
// ------------------------------------------------------------------------------------------
struct StagingRing {
  static constexpr int SLOTS = 4;
  void *buf[SLOTS] = {};
  size_t cap[SLOTS] = {};
  cudaEvent_t ev[SLOTS] = {};
  bool busy[SLOTS] = {};
  int next = 0;

  // Returns a pinned buffer of at least `bytes`, free for the host to write.
  int acquire(size_t bytes, int *slot) {
    int s = next;
    next = (next + 1) % SLOTS;
    if (!ev[s]) EC_CUDA(cudaEventCreateWithFlags(&ev[s], cudaEventDisableTiming));
    if (busy[s]) {
      EC_CUDA(cudaEventSynchronize(ev[s]));
      busy[s] = false;
    }
    if (cap[s] < bytes) {
      if (buf[s]) EC_CUDA(cudaFreeHost(buf[s]));
      buf[s] = nullptr;
      cap[s] = 0;
      size_t want = std::max(bytes, (size_t)1 << 20);
      EC_CUDA(cudaHostAlloc(&buf[s], want, cudaHostAllocDefault));
      cap[s] = want;
    }
    *slot = s;
    return EC_OK;
  }
  int release_after(int s, cudaStream_t st) {
    EC_CUDA(cudaEventRecord(ev[s], st));
    busy[s] = true;
    return EC_OK;
  }
  void destroy() {
    for (int s = 0; s < SLOTS; ++s) {
      if (busy[s]) cudaEventSynchronize(ev[s]);
      if (ev[s]) cudaEventDestroy(ev[s]);
      if (buf[s]) cudaFreeHost(buf[s]);
      ev[s] = nullptr;
      buf[s] = nullptr;
      cap[s] = 0;
      busy[s] = false;
    }
  }
};

bool is_pinned_host(const void *p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeHost;
}
bool is_device_ptr(const void *p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// Staging copy with streaming (non-temporal) stores.  Neither direction of the pageable <-> pinned
// staging re-reads what it has just written, so ordinary stores only add a read-for-ownership of
// every destination line: on this memcpy-bound path that is a third of the host memory traffic.
// AVX2 when the CPU has it (checked once), std::memcpy otherwise and for small pieces.
#if defined(__x86_64__)
#include <immintrin.h>
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

// Small persistent worker pool for pageable <-> pinned staging copies.
class CopyPool {
 public:
  explicit CopyPool(int n) : stop_(false), pending_(0) {
    for (int i = 0; i < n; ++i) workers_.emplace_back([this] { run(); });
  }
  ~CopyPool() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      stop_ = true;
    }
    cv_.notify_all();
    for (auto &t : workers_) t.join();
  }
  int size() const { return (int)workers_.size(); }
  // Run fn(i) for i in [0, n) on the pool and wait.
  void parallel_for(int64_t n, const std::function<void(int64_t)> &fn) {
    if (n <= 0) return;
    if (workers_.empty()) {
      for (int64_t i = 0; i < n; ++i) fn(i);
      return;
    }
    {
      std::lock_guard<std::mutex> lk(mu_);
      fn_ = &fn;
      next_ = 0;
      end_ = n;
      pending_ = n;
    }
    cv_.notify_all();
    std::unique_lock<std::mutex> lk(mu_);
    done_cv_.wait(lk, [this] { return pending_ == 0; });
    fn_ = nullptr;
  }

 private:
  void run() {
    std::unique_lock<std::mutex> lk(mu_);
    for (;;) {
      cv_.wait(lk, [this] { return stop_ || (fn_ && next_ < end_); });
      if (stop_) return;
      while (fn_ && next_ < end_) {
        int64_t i = next_++;
        const std::function<void(int64_t)> *f = fn_;
        lk.unlock();
        (*f)(i);
        lk.lock();
        if (--pending_ == 0) done_cv_.notify_all();
      }
    }
  }
  std::vector<std::thread> workers_;
  std::mutex mu_;
  std::condition_variable cv_, done_cv_;
  bool stop_;
  const std::function<void(int64_t)> *fn_ = nullptr;
  int64_t next_ = 0, end_ = 0, pending_;
};

}  // namespace
