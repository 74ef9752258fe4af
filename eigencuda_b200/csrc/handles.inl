// eigencuda_b200/csrc/handles.inl -- the opaque handles behind the C ABI (ec_pipeline, ec_matrix), stream registry, device guard
// Part of the single translation unit ec_api.cu (included there, in order); not a standalone header.
#pragma once

// ------------------------------------------------------------------------------------------
// handles
// ------------------------------------------------------------------------------------------
struct StreamEngine;
struct MultiGpu;

struct ec_pipeline {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool owns_stream = true;
  cudaEvent_t ev_order = nullptr;   // cross-stream ordering of ec_gemm operands (created on first use)
  int backend = EC_BACKEND_FP64_DMMA;
  int sm_count = 0;
  int64_t launches = 0;
  const char *last_kernel = "none";
  StagingRing staging;
  StreamEngine *engine = nullptr;
  MultiGpu *multi = nullptr;   // child pipelines + NCCL communicators of the multi-GPU scheduler (multi_gpu.inl)
  // workers for the matrix-at-a-time staging copies of copy_to_gpu / the conversion back
  // (created on the first pageable transfer of 1 MiB or more; EC_COMPAT_THREADS=0 disables)
  bool mem_pool = false;   // device buffers of CudaMatrix come from the stream-ordered pool
  CopyPool *compat_pool = nullptr;
  bool compat_pool_tried = false;
  int cfg_chunk = 0, cfg_ring = 0, cfg_threads = 0;
  // stream-K fix-up workspace (partial accumulator tiles + flags), allocated on first use
  double *sk_workspace = nullptr;
  unsigned *sk_flags = nullptr;
  unsigned sk_epoch = 0;
  int stream_k = 1;   // 0 disables the stream-K schedule (EC_STREAM_K=0)
  int sm_margin = 0;  // SMs left free for concurrent kernels (NCCL during SUMMA)
  int one_cta_per_tile = 0;  // non-persistent launch for long tiles (shared-GPU friendly)
  // dynamic tile schedule + contention detection (dgemm_launch.inl: ensure_sched_words)
  int dynamic_tiles = 1;                       // EC_DYNAMIC_TILES=0: static round-robin, no statistics
  unsigned long long *sched_words = nullptr;   // device: [0] claim counter, [1..3] start-skew statistics
  unsigned long long dp_claims = 0;            // claims made so far (the counter is never reset)
  unsigned *skew_host = nullptr, *skew_dev = nullptr;   // host-mapped: start skew (ns) of the last persistent grid
  unsigned long long stats_launches = 0;       // launches that carried the start-skew check (slot parity)
  int contended_left = 0;                      // launches left in "the GPU is shared" mode
  int64_t contended_events = 0;                // how often that mode was entered
  int fuse_min_kblocks = 2;  // ... only for tiles of at least this many k-iterations, or L2-resident operands (EC_FUSE_MIN_KBLOCKS)
  int fuse_epilogue = 1;     // 128x128 tile: C stores ride under the next tile's first k-iteration (EC_FUSE_EPILOGUE=0: plain epilogue)
  int pdl = 1;               // programmatic dependent launch of the GEMM kernels (EC_PDL=0 turns it off)
  // small direct-mapped cache of encoded tensor maps: the compat loop and every tensor-stream
  // chunk reuse the same few device buffers, and cuTensorMapEncodeTiled costs about a microsecond
  struct TmapEntry {
    const void *ptr = nullptr;
    int64_t inner = 0, outer = 0, batch = 0, ld = 0, stride = 0;
    int box_outer = 0;
    CUtensorMap map;
  } tmap_cache[16];
  // EC_BACKEND_OZAKI_I8 workspace: digit planes and row/column scales, grown on demand
  int8_t *oz_planes_a = nullptr, *oz_planes_b = nullptr;
  double *oz_scale_a = nullptr, *oz_scale_b = nullptr;
  unsigned long long *oz_amax = nullptr;
  size_t oz_planes_a_bytes = 0, oz_planes_b_bytes = 0, oz_scale_a_bytes = 0, oz_scale_b_bytes = 0, oz_amax_bytes = 0;
  cudaStream_t oz_split_stream = nullptr;          // splits of chunk c+1 overlap the GEMM of chunk c
  cudaEvent_t oz_ev_split[2] = {nullptr, nullptr}, oz_ev_gemm[2] = {nullptr, nullptr};
};

struct ec_matrix {
  double *data = nullptr;
  int64_t rows = 0, cols = 0;
  cudaStream_t stream = nullptr;
  int device = 0;
  bool pooled = false;   // allocated with cudaMallocAsync on `stream` (freed in stream order)
};

namespace {

// stream -> pipeline registry, so that CudaMatrix (which only knows its stream, reference
// include/cudamatrix.hpp:57) can find its pipeline's staging ring.
std::mutex g_registry_mu;
std::unordered_map<void *, ec_pipeline *> g_registry;

ec_pipeline *pipeline_of_stream(cudaStream_t s) {
  std::lock_guard<std::mutex> lk(g_registry_mu);
  auto it = g_registry.find((void *)s);
  return it == g_registry.end() ? nullptr : it->second;
}

struct DeviceGuard {
  int prev = -1;
  bool active = false;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) {
      cudaSetDevice(dev);
      active = true;
    }
  }
  ~DeviceGuard() {
    if (active) cudaSetDevice(prev);
  }
};


// Run a few launches of this pipeline on another of its internal streams (the engine's compute
// stream, the Ozaki split stream); restores the stream on every exit path.
struct StreamScope {
  ec_pipeline *p;
  cudaStream_t saved;
  StreamScope(ec_pipeline *p_, cudaStream_t s) : p(p_), saved(p_->stream) { p->stream = s; }
  ~StreamScope() { p->stream = saved; }
  StreamScope(const StreamScope &) = delete;
  StreamScope &operator=(const StreamScope &) = delete;
};

}  // namespace
