// eigencuda_b200/csrc/host_util.inl -- error reporting, driver entry point, pinned staging ring, pointer queries
// Part of the single translation unit ec_api.cu (included there, in order); not a standalone header.
#pragma once
#include "copy_pool.hpp"   // staging_copy, CopyPool (plain C++, tested on the CPU)
using ec_host::CopyPool;
using ec_host::staging_copy;

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
// NVTX ranges: the library's phases show up by name on a profiler
// timeline.  nvtx3 is header-only and costs one pointer test per call when no tool is attached.
// ------------------------------------------------------------------------------------------
struct TraceRange {
  explicit TraceRange(const char *name) { nvtxRangePushA(name); }
  ~TraceRange() { nvtxRangePop(); }
  TraceRange(const TraceRange &) = delete;
  TraceRange &operator=(const TraceRange &) = delete;
};

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
    // slots only grow on demand and shrink again once a much smaller request follows an
    // oversized one (a one-off large upload must not keep hundreds of MiB pinned per slot)
    if (cap[s] < bytes || (cap[s] > ((size_t)64 << 20) && bytes <= cap[s] / 8)) {
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

}  // namespace
