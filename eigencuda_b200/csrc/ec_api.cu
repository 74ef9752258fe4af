// eigencuda_b200/csrc/ec_api.cu -- implementation of the C ABI declared in include/eigencuda_b200.h.
//
// Host side of the hot path: the memory subsystem that replaces CudaMatrix
// (reference src/cudamatrix.cc:20-79), the op dispatch that replaces CudaPipeline::gemm
// (reference src/cudapipeline.cc:17-32) and the overlapped tensor-stream engine that replaces the
// caller-side loop of README.md:62-66.  No cuBLAS, no CPU compute fallback: every product is a
// launch of a kernel from dgemm_kernels.cuh.
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../include/eigencuda_b200.h"
#include "dgemm_kernels.cuh"
#include "ozaki_kernels.cuh"

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

// ------------------------------------------------------------------------------------------
// handles
// ------------------------------------------------------------------------------------------
struct StreamEngine;

struct ec_pipeline {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool owns_stream = true;
  int backend = EC_BACKEND_FP64_DMMA;
  int sm_count = 0;
  int64_t launches = 0;
  const char *last_kernel = "none";
  StagingRing staging;
  StreamEngine *engine = nullptr;
  int cfg_chunk = 0, cfg_ring = 0, cfg_threads = 0;
  // stream-K fix-up workspace (partial accumulator tiles + flags), allocated on first use
  double *sk_workspace = nullptr;
  unsigned *sk_flags = nullptr;
  unsigned sk_epoch = 0;
  int stream_k = 1;   // 0 disables the stream-K schedule (EC_STREAM_K=0)
  int sm_margin = 0;  // SMs left free for concurrent kernels (NCCL during SUMMA)
  int one_cta_per_tile = 0;  // non-persistent launch for long tiles (shared-GPU friendly)
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

// ------------------------------------------------------------------------------------------
// kernel dispatch
// ------------------------------------------------------------------------------------------
struct Operand {
  const double *ptr;
  int64_t ld, stride;
  bool kmajor;     // contiguous along k
  int64_t mn, k;   // extents of op(X): mn x k
};

bool tma_ok(const Operand &o, int64_t batch) {
  if (((uintptr_t)o.ptr & 15u) != 0) return false;
  if ((o.ld & 1) != 0) return false;                       // row pitch must be a multiple of 16 B
  if (batch > 1 && o.stride != 0 && (o.stride & 1) != 0) return false;
  if (o.mn >= ((int64_t)1 << 31) || o.k >= ((int64_t)1 << 31)) return false;
  if (o.ld * 8 >= ((int64_t)1 << 40) || o.stride * 8 >= ((int64_t)1 << 40)) return false;
  return true;
}

int make_tensor_map(ec_pipeline *p, CUtensorMap *tm, const Operand &o, int64_t batch, int tile_extent) {
  const bool batched = batch > 1 && o.stride != 0;
  const int64_t inner = o.kmajor ? o.k : o.mn, outer = o.kmajor ? o.mn : o.k;
  const int box_outer = o.kmajor ? tile_extent : ec::BK;
  const int64_t nb = batched ? batch : 1, st = batched ? o.stride : 0;
  // cache lookup
  uint64_t hsh = (uint64_t)(uintptr_t)o.ptr * 0x9E3779B97F4A7C15ull ^ (uint64_t)inner * 0xBF58476D1CE4E5B9ull ^
                 (uint64_t)outer * 0x94D049BB133111EBull ^ (uint64_t)o.ld * 31 ^ (uint64_t)box_outer;
  ec_pipeline::TmapEntry &e = p->tmap_cache[(hsh >> 40) & 15];
  if (e.ptr == o.ptr && e.inner == inner && e.outer == outer && e.batch == nb && e.ld == o.ld && e.stride == st &&
      e.box_outer == box_outer) {
    *tm = e.map;
    return EC_OK;
  }
  encode_tiled_fn enc = get_encode_tiled();
  if (!enc) return fail(EC_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t dims[3], strides[2];
  cuuint32_t box[3], estr[3] = {1, 1, 1};
  dims[0] = (cuuint64_t)inner;
  dims[1] = (cuuint64_t)outer;
  dims[2] = (cuuint64_t)nb;
  strides[0] = (cuuint64_t)o.ld * 8;
  // stride of the (possibly degenerate) batch dimension: any legal multiple of 16 when unused
  strides[1] = batched ? (cuuint64_t)o.stride * 8 : (cuuint64_t)o.ld * 8 * (cuuint64_t)std::max<int64_t>(outer, 1);
  box[0] = 16;
  box[1] = (cuuint32_t)box_outer;
  box[2] = 1;
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, (void *)o.ptr, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(EC_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  e.ptr = o.ptr; e.inner = inner; e.outer = outer; e.batch = nb; e.ld = o.ld; e.stride = st; e.box_outer = box_outer;
  e.map = *tm;
  return EC_OK;
}

constexpr size_t SK_WORKSPACE_BYTES = (size_t)640 * 128 * 128 * 8 / 4;  // >= grid * BM*BN*8 for every config
constexpr int SK_MAX_GRID = 1024;

int ensure_sk_workspace(ec_pipeline *p) {
  if (p->sk_workspace) return EC_OK;
  EC_CUDA(cudaMalloc((void **)&p->sk_workspace, SK_WORKSPACE_BYTES));
  EC_CUDA(cudaMalloc((void **)&p->sk_flags, SK_MAX_GRID * sizeof(unsigned)));
  EC_CUDA(cudaMemsetAsync(p->sk_flags, 0, SK_MAX_GRID * sizeof(unsigned), p->stream));
  return EC_OK;
}

// Resident CTAs per SM of one kernel instantiation on one device (cached).
template <class Cfg, bool AK, bool BK_>
int ctas_per_sm(ec_pipeline *p, int *out) {
  auto kern = ec::dgemm_tma_dmma_kernel<Cfg, AK, BK_>;
  static std::atomic<int> cached[64];
  int v = cached[p->device & 63].load();
  if (v == 0) {
    EC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    int occ = 0;
    EC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, Cfg::THREADS, Cfg::SMEM_BYTES));
    if (occ < 1) return fail(EC_ERR_CUDA, "kernel does not fit on an SM (occupancy 0)");
    v = std::min(occ, Cfg::MIN_CTAS);
    cached[p->device & 63].store(v);
  }
  *out = v;
  return EC_OK;
}

template <class Cfg, bool AK, bool BK_>
int launch_tma_dmma(ec_pipeline *p, const CUtensorMap &tmA, const CUtensorMap &tmB,
                    ec::GemmParams &gp, bool want_sk) {
  auto kern = ec::dgemm_tma_dmma_kernel<Cfg, AK, BK_>;
  int ctas = 1;
  int rc = ctas_per_sm<Cfg, AK, BK_>(p, &ctas);
  if (rc) return rc;
  const int64_t slots = (int64_t)(p->sm_count - p->sm_margin) * ctas;
  const int64_t kblocks = (gp.K + ec::BK - 1) / ec::BK;
  int grid = (int)std::min<int64_t>(gp.total_tiles, slots);
  gp.sk_tiles = 0;
  gp.sk_sms = p->sm_count - p->sm_margin;
  gp.sk_workspace = nullptr;
  gp.sk_flags = nullptr;
  gp.sk_epoch = 0;
  // Stream-K region: the tiles of the last, partial wave plus (when there is one) one full wave,
  // so that every CTA gets between one and two tiles' worth of k-iterations in the region.
  const int64_t rem = gp.total_tiles % slots;
  if (want_sk && p->stream_k && (rem != 0 || getenv("EC_SK_ALL")) && slots <= SK_MAX_GRID &&
      slots * (int64_t)(Cfg::BM * Cfg::BN * 8) <= (int64_t)SK_WORKSPACE_BYTES) {
    int64_t sk_tiles = rem + (gp.total_tiles > slots ? slots : 0);
    if (const char *e = getenv("EC_SK_ALL")) {   // experiment: stream-K over every tile
      if (atoi(e)) sk_tiles = gp.total_tiles;
    }
    if (sk_tiles * kblocks >= 4 * slots) {   // at least 4 k-iterations per CTA in the region
      rc = ensure_sk_workspace(p);
      if (rc) return rc;
      gp.sk_tiles = sk_tiles;
      gp.sk_workspace = p->sk_workspace;
      gp.sk_flags = p->sk_flags;
      gp.sk_epoch = ++p->sk_epoch;
      if (gp.sk_epoch == 0) gp.sk_epoch = ++p->sk_epoch;   // 0 is the cleared state
      grid = (int)slots;
    }
  }
  if (gp.sk_tiles == 0 && p->one_cta_per_tile && kblocks >= 64 && gp.total_tiles > slots)
    grid = (int)std::min<int64_t>(gp.total_tiles, (int64_t)1 << 30);   // one tile per CTA
  if (gp.sk_tiles > 0) {
    // CTAs wait on one another's partials: a cooperative launch guarantees co-residency.
    void *args[3] = {(void *)&tmA, (void *)&tmB, (void *)&gp};
    EC_CUDA(cudaLaunchCooperativeKernel((const void *)kern, dim3(grid), dim3(Cfg::THREADS), args,
                                        (size_t)Cfg::SMEM_BYTES, p->stream));
  } else {
    kern<<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, p->stream>>>(tmA, tmB, gp);
  }
  EC_CUDA(cudaGetLastError());
  p->launches++;
  return EC_OK;
}

template <class Cfg>
int run_tma_dmma(ec_pipeline *p, const Operand &oa, const Operand &ob, int64_t m, int64_t n,
                 int64_t k, double alpha, double beta, double *C, int64_t ldc, int64_t stride_c,
                 int64_t stride_a, int64_t stride_b, int64_t batch, bool want_sk,
                 const char *const names[8]) {
  CUtensorMap tmA, tmB;
  int rc = make_tensor_map(p, &tmA, oa, batch, Cfg::BM);
  if (rc) return rc;
  rc = make_tensor_map(p, &tmB, ob, batch, Cfg::BN);
  if (rc) return rc;
  ec::GemmParams gp;
  gp.M = m; gp.N = n; gp.K = k; gp.batch = batch;
  gp.alpha = alpha; gp.beta = beta;
  gp.C = C; gp.ldc = ldc; gp.stride_c = stride_c;
  gp.tiles_m = (int)((m + Cfg::BM - 1) / Cfg::BM);
  gp.tiles_n = (int)((n + Cfg::BN - 1) / Cfg::BN);
  gp.total_tiles = (int64_t)gp.tiles_m * gp.tiles_n * batch;
  gp.a_batched = (batch > 1 && stride_a != 0);
  gp.b_batched = (batch > 1 && stride_b != 0);
  gp.c_vec2 = (((uintptr_t)C & 15u) == 0 && (ldc & 1) == 0 && (batch <= 1 || (stride_c & 1) == 0)) ? 1 : 0;
  int idx;
  if (oa.kmajor && ob.kmajor) { idx = 0; rc = launch_tma_dmma<Cfg, true, true>(p, tmA, tmB, gp, want_sk); }
  else if (oa.kmajor && !ob.kmajor) { idx = 1; rc = launch_tma_dmma<Cfg, true, false>(p, tmA, tmB, gp, want_sk); }
  else if (!oa.kmajor && ob.kmajor) { idx = 2; rc = launch_tma_dmma<Cfg, false, true>(p, tmA, tmB, gp, want_sk); }
  else { idx = 3; rc = launch_tma_dmma<Cfg, false, false>(p, tmA, tmB, gp, want_sk); }
  p->last_kernel = names[idx + (gp.sk_tiles > 0 ? 4 : 0)];
  return rc;
}

// Tile and schedule choice.  All configurations run the DMMA pipe at nearly the same rate once
// it is fed; what differs is (a) load balance -- under the data-parallel schedule the busiest SM
// does ceil(tiles / SMs) tiles, under stream-K every SM does tiles / SMs of them plus one
// workspace round trip -- and (b) per-tile overhead and main-loop efficiency, which get worse
// as tiles shrink.  Cost model in SM clocks, fitted to profiles/r01_tile_sweep*.json.
enum { TILE_AUTO = 0, TILE_L = 1, TILE_S = 2, TILE_XS = 3 };

struct TileChoice { int tile; bool sk; };

TileChoice choose_tile(const ec_pipeline *p, int64_t m, int64_t n, int64_t k, int64_t batch) {
  int force = 0;
  bool allow_sk = p->stream_k != 0, force_sk = false;
  // EC_FORCE_TILE=1..4 and EC_STREAM_K=0 (never) | 1 (cost model) | 2 (whenever feasible) override
  // the choice for tools/tile_sweep.py and the tests; read per call
  if (const char *e = getenv("EC_FORCE_TILE")) force = atoi(e);
  if (const char *e = getenv("EC_STREAM_K")) {
    allow_sk = atoi(e) != 0;
    force_sk = atoi(e) == 2;
  }
  // eff[r-1]: fraction of the DMMA issue rate the main loop sustains with r CTAs resident per SM
  // (the 4-warp tiles put one warp on each SM sub-partition per CTA and hide their LDS latency
  // badly until a second or third CTA shares the SM); overhead: exposed clocks per tile; fixup:
  // exposed clocks of one stream-K tail (publish partial, wait, read, epilogue).  Fitted to
  // profiles/r01_tile_sweep*.json and profiles/r01_small_probe.txt (back-to-back launches).
  struct Cand { int id, bm, bn, ctas; double eff[3], overhead, fixup; };
  const Cand cands[3] = {{TILE_L, 128, 128, 1, {0.987, 0.987, 0.987}, 2300.0, 30000.0},
                         {TILE_S, 64, 64, 2, {0.59, 0.955, 0.955}, 1200.0, 17000.0},
                         {TILE_XS, 64, 32, 3, {0.50, 0.76, 0.80}, 800.0, 20000.0}};
  const double kpad = (double)((k + 15) / 16 * 16);
  const int64_t kblocks = (k + 15) / 16;
  double best = 1e300;
  TileChoice pick{TILE_L, false};
  for (const Cand &c : cands) {
    if (force >= TILE_L && force <= TILE_XS && c.id != force) continue;
    const int64_t tiles = ((m + c.bm - 1) / c.bm) * ((n + c.bn - 1) / c.bn) * batch;
    const int sms = p->sm_count - p->sm_margin;
    const int64_t slots = (int64_t)sms * c.ctas;
    const int64_t resident = std::max<int64_t>(1, std::min<int64_t>(c.ctas, (tiles + sms - 1) / sms));
    const double tile_clk = (double)c.bm * c.bn * kpad / 64.0 / c.eff[resident - 1] + c.overhead;
    const double dp = (double)((tiles + sms - 1) / sms) * tile_clk;
    if (dp < best) { best = dp; pick = {c.id, false}; }
    const int64_t rem = tiles % slots;
    const int64_t sk_tiles = rem + (tiles > slots ? slots : 0);
    if (allow_sk && rem != 0 && sk_tiles * kblocks >= 4 * slots) {
      // every SM does its exact share; one partial tile written and one read per CTA
      // the stream-K grid fills every slot, so the SM runs at its full-occupancy efficiency
      const double tile_clk_sk = (double)c.bm * c.bn * kpad / 64.0 / c.eff[c.ctas - 1] + c.overhead;
      // with far fewer tiles than CTAs a tile is cut into several partials and its owner adds
      // them one after the other: the exposed tail grows with the partials per tile
      const double splits = std::max(1.0, (double)slots / (double)tiles);
      const double sk = force_sk ? dp * 1e-3
                                 : (double)tiles / sms * tile_clk_sk + c.fixup * (1.0 + 0.35 * (splits - 1.0));
      if (sk < best) { best = sk; pick = {c.id, true}; }
    }
  }
  return pick;
}


// ------------------------------------------------------------------------------------------
// EC_BACKEND_OZAKI_I8: FP64 product emulated on the int8 tensor cores (ozaki_kernels.cuh)
// ------------------------------------------------------------------------------------------
constexpr int OZ_S = 7;   // digit planes: 49 fractional bits below each row/column maximum

// Shapes the emulation takes; everything else runs the native FP64 kernels (which are at least
// as accurate).  K <= 65536 keeps the int32 accumulators exact for 7 planes.
constexpr int64_t OZ_MAX_K = 65536;   // 7 * 65536 * 65^2 < 2^31: deeper products are cut along k
bool ozaki_eligible(int64_t m, int64_t n, int64_t k) {
  return k >= 64 && m >= 64 && n >= 32 && m * n * k >= (int64_t)128 * 128 * 128 &&
         m < ((int64_t)1 << 31) && n < ((int64_t)1 << 31);
}

template <class T>
int oz_grow(T **buf, size_t *have, size_t want) {
  if (want <= *have) return EC_OK;
  if (*buf) EC_CUDA(cudaFree(*buf));
  *buf = nullptr;
  *have = 0;
  EC_CUDA(cudaMalloc((void **)buf, want));
  *have = want;
  return EC_OK;
}

int oz_tensor_map(CUtensorMap *tm, const int8_t *planes, int64_t K, int64_t Kp, int64_t R, int64_t batch,
                  int box_rows) {
  encode_tiled_fn enc = get_encode_tiled();
  if (!enc) return fail(EC_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t dims[4] = {(cuuint64_t)K, (cuuint64_t)R, (cuuint64_t)OZ_S, (cuuint64_t)std::max<int64_t>(batch, 1)};
  cuuint64_t strides[3] = {(cuuint64_t)Kp, (cuuint64_t)(R * Kp), (cuuint64_t)(OZ_S * R * Kp)};
  cuuint32_t box[4] = {(cuuint32_t)ec::oz::OBK, (cuuint32_t)box_rows, (cuuint32_t)OZ_S, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, (void *)planes, dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(EC_ERR_CUDA, "cuTensorMapEncodeTiled (int8 planes) failed with CUresult %d", (int)r);
  return EC_OK;
}

// Scale + split one operand: op(X) viewed as R rows x K, element (r, q) at X[r*rs + q*ks].
// `amax` is scratch for R*batch abs-max bit patterns.
int oz_split(ec_pipeline *p, const double *X, int64_t rs, int64_t ks, int64_t bstride, int64_t R, int64_t K,
             int64_t Kp, int64_t batch, int8_t *planes, double *scale, unsigned long long *amax) {
  ec::oz::SplitParams sp{X, rs, ks, bstride, R, K, Kp, batch, planes, scale};
  const bool row_contig = (rs == 1 && ks != 1);
  EC_CUDA(cudaMemsetAsync(amax, 0, (size_t)(R * batch) * 8, p->stream));
  const int64_t rt = (R + ec::oz::SPLIT_ROWS - 1) / ec::oz::SPLIT_ROWS;
  const int64_t tiles_k = ((K + ec::oz::SPLIT_K - 1) / ec::oz::SPLIT_K) * rt * batch;
  const int64_t tiles_kp = ((Kp + ec::oz::SPLIT_K - 1) / ec::oz::SPLIT_K) * rt * batch;
  const int g1 = (int)std::min<int64_t>(tiles_k, (int64_t)p->sm_count * 8);
  const int g2 = (int)std::min<int64_t>(tiles_kp, (int64_t)p->sm_count * 8);
  if (row_contig) {
    ec::oz::ozaki_absmax_kernel<true><<<g1, ec::oz::SPLIT_THREADS, 0, p->stream>>>(sp, amax);
    ec::oz::ozaki_split_kernel<OZ_S, true><<<g2, ec::oz::SPLIT_THREADS, 0, p->stream>>>(sp, amax);
  } else {
    ec::oz::ozaki_absmax_kernel<false><<<g1, ec::oz::SPLIT_THREADS, 0, p->stream>>>(sp, amax);
    ec::oz::ozaki_split_kernel<OZ_S, false><<<g2, ec::oz::SPLIT_THREADS, 0, p->stream>>>(sp, amax);
  }
  EC_CUDA(cudaGetLastError());
  p->launches += 2;
  return EC_OK;
}

int ozaki_dispatch(ec_pipeline *p, int transa, int transb, int64_t m, int64_t n, int64_t k, double alpha,
                   const double *A, int64_t lda, int64_t stride_a, const double *B, int64_t ldb, int64_t stride_b,
                   double beta, double *C, int64_t ldc, int64_t stride_c, int64_t batch) {
  using Cfg = ec::oz::OzCfg<OZ_S>;
  const int64_t Kp = (k + 15) / 16 * 16;
  const bool a_batched = batch > 1 && stride_a != 0, b_batched = batch > 1 && stride_b != 0;
  // chunk the batch so that the digit planes of a chunk stay around L2 size; the batched
  // operand(s) are double-buffered so that the split of chunk c+1 (LSU / FP64 pipe / HBM) runs on
  // a side stream underneath the GEMM of chunk c (tensor cores / L2)
  const int64_t per_problem = OZ_S * Kp * ((a_batched ? m : 0) + (b_batched ? n : 0));
  int64_t chunk = batch;
  if (per_problem > 0) chunk = std::max<int64_t>(1, std::min<int64_t>(batch, ((int64_t)192 << 20) / per_problem));
  if (const char *e = getenv("EC_OZAKI_CHUNK")) chunk = std::max<int64_t>(1, std::min<int64_t>(batch, atoll(e)));
  const int64_t nchunks = (batch + chunk - 1) / chunk;
  const bool pipelined = nchunks > 1 && (a_batched || b_batched);
  const int nbuf = pipelined ? 2 : 1;
  const int64_t ca = a_batched ? chunk : 1, cb = b_batched ? chunk : 1;
  const size_t pa = (size_t)(OZ_S * m * Kp * ca), pb = (size_t)(OZ_S * n * Kp * cb);
  const size_t sa_b = (size_t)(m * ca * 8), sb_b = (size_t)(n * cb * 8), am_b = (size_t)(std::max(m * ca, n * cb) * 8);
  int rc;
  if ((rc = oz_grow(&p->oz_planes_a, &p->oz_planes_a_bytes, pa * (a_batched ? nbuf : 1)))) return rc;
  if ((rc = oz_grow(&p->oz_planes_b, &p->oz_planes_b_bytes, pb * (b_batched ? nbuf : 1)))) return rc;
  if ((rc = oz_grow(&p->oz_scale_a, &p->oz_scale_a_bytes, sa_b * (a_batched ? nbuf : 1)))) return rc;
  if ((rc = oz_grow(&p->oz_scale_b, &p->oz_scale_b_bytes, sb_b * (b_batched ? nbuf : 1)))) return rc;
  if ((rc = oz_grow(&p->oz_amax, &p->oz_amax_bytes, am_b * 3))) return rc;   // [shared operand, buffer 0, buffer 1]
  if (pipelined && !p->oz_split_stream) {
    EC_CUDA(cudaStreamCreateWithFlags(&p->oz_split_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
      EC_CUDA(cudaEventCreateWithFlags(&p->oz_ev_split[i], cudaEventDisableTiming));
      EC_CUDA(cudaEventCreateWithFlags(&p->oz_ev_gemm[i], cudaEventDisableTiming));
    }
  }
  auto kern = ec::oz::ozaki_gemm_kernel<OZ_S>;
  static std::atomic<uint64_t> configured{0};
  const uint64_t bit = 1ull << (p->device & 63);
  if (!(configured.load() & bit)) {
    EC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    configured.fetch_or(bit);
  }
  // op(A)(r, q): N -> A[r + q*lda]; T -> A[r*lda + q].  op(B)^T(j, q) = op(B)(q, j): N -> B[j*ldb + q]; T -> B[q*ldb + j]
  const int64_t a_rs = transa ? lda : 1, a_ks = transa ? 1 : lda;
  const int64_t b_rs = transb ? 1 : ldb, b_ks = transb ? ldb : 1;
  unsigned long long *amax_shared = p->oz_amax, *amax_buf[2] = {p->oz_amax + am_b / 8, p->oz_amax + 2 * (am_b / 8)};
  if (!a_batched && (rc = oz_split(p, A, a_rs, a_ks, 0, m, k, Kp, 1, p->oz_planes_a, p->oz_scale_a, amax_shared))) return rc;
  if (!b_batched && (rc = oz_split(p, B, b_rs, b_ks, 0, n, k, Kp, 1, p->oz_planes_b, p->oz_scale_b, amax_shared))) return rc;

  cudaStream_t main_stream = p->stream;
  // split of chunk c into buffer c & 1, on the side stream when pipelined
  auto split_chunk = [&](int64_t c) -> int {
    const int64_t b0 = c * chunk, nb = std::min(chunk, batch - b0);
    const int buf = pipelined ? (int)(c & 1) : 0;
    if (pipelined) {
      p->stream = p->oz_split_stream;
      if (c >= 2) EC_CUDA(cudaStreamWaitEvent(p->oz_split_stream, p->oz_ev_gemm[buf], 0));   // buffer consumed
    }
    int r = EC_OK;
    if (a_batched)
      r = oz_split(p, A + b0 * stride_a, a_rs, a_ks, stride_a, m, k, Kp, nb, p->oz_planes_a + buf * pa,
                   p->oz_scale_a + buf * (sa_b / 8), amax_buf[buf]);
    if (!r && b_batched)
      r = oz_split(p, B + b0 * stride_b, b_rs, b_ks, stride_b, n, k, Kp, nb, p->oz_planes_b + buf * pb,
                   p->oz_scale_b + buf * (sb_b / 8), amax_buf[buf]);
    if (pipelined) {
      if (!r) {
        cudaError_t e = cudaEventRecord(p->oz_ev_split[buf], p->oz_split_stream);
        if (e != cudaSuccess) r = fail(EC_ERR_CUDA, "cudaEventRecord failed: %s", cudaGetErrorString(e));
      }
      p->stream = main_stream;
    }
    return r;
  };
  if (pipelined) {
    // the side stream starts after whatever produced the operands on the main stream
    EC_CUDA(cudaEventRecord(p->oz_ev_gemm[0], main_stream));
    EC_CUDA(cudaStreamWaitEvent(p->oz_split_stream, p->oz_ev_gemm[0], 0));
  }
  if (a_batched || b_batched) {
    if ((rc = split_chunk(0))) { p->stream = main_stream; return rc; }
  }
  for (int64_t c = 0; c < nchunks; ++c) {
    const int64_t b0 = c * chunk, nb = std::min(chunk, batch - b0);
    const int buf = pipelined ? (int)(c & 1) : 0;
    if (pipelined && c + 1 < nchunks && (rc = split_chunk(c + 1))) { p->stream = main_stream; return rc; }
    if (!pipelined && c > 0 && (a_batched || b_batched) && (rc = split_chunk(c))) return rc;
    if (pipelined) EC_CUDA(cudaStreamWaitEvent(main_stream, p->oz_ev_split[buf], 0));
    CUtensorMap tmA, tmB;
    if ((rc = oz_tensor_map(&tmA, p->oz_planes_a + (a_batched ? buf * pa : 0), k, Kp, m, a_batched ? nb : 1, ec::oz::OM))) return rc;
    if ((rc = oz_tensor_map(&tmB, p->oz_planes_b + (b_batched ? buf * pb : 0), k, Kp, n, b_batched ? nb : 1, ec::oz::ON))) return rc;
    ec::oz::GemmOzParams gp;
    gp.M = m; gp.N = n; gp.K = k; gp.batch = nb;
    gp.alpha = alpha; gp.beta = beta;
    gp.C = C + b0 * stride_c; gp.ldc = ldc; gp.stride_c = stride_c;
    gp.sa = p->oz_scale_a + (a_batched ? buf * (sa_b / 8) : 0);
    gp.sb = p->oz_scale_b + (b_batched ? buf * (sb_b / 8) : 0);
    gp.a_batched = a_batched; gp.b_batched = b_batched;
    gp.tiles_m = (int)((m + ec::oz::OM - 1) / ec::oz::OM);
    gp.tiles_n = (int)((n + ec::oz::ON - 1) / ec::oz::ON);
    gp.total_tiles = (int64_t)gp.tiles_m * gp.tiles_n * nb;
    const int grid = (int)std::min<int64_t>(gp.total_tiles, p->sm_count - p->sm_margin);
    kern<<<grid, ec::oz::OTHREADS, Cfg::SMEM_BYTES, main_stream>>>(tmA, tmB, gp);
    EC_CUDA(cudaGetLastError());
    p->launches++;
    if (pipelined) EC_CUDA(cudaEventRecord(p->oz_ev_gemm[buf], main_stream));
  }
  p->last_kernel = "ozaki_i8_s7";
  return EC_OK;
}

int dgemm_dispatch(ec_pipeline *p, int transa, int transb, int64_t m, int64_t n, int64_t k,
                   double alpha, const double *A, int64_t lda, int64_t stride_a, const double *B,
                   int64_t ldb, int64_t stride_b, double beta, double *C, int64_t ldc,
                   int64_t stride_c, int64_t batch) {
  if (!p) return fail(EC_ERR_ARG, "null pipeline");
  if (m < 0 || n < 0 || k < 0 || batch < 0) return fail(EC_ERR_ARG, "negative dimension");
  if ((transa != EC_OP_N && transa != EC_OP_T) || (transb != EC_OP_N && transb != EC_OP_T))
    return fail(EC_ERR_ARG, "bad transpose flag");
  if (m == 0 || n == 0 || batch == 0) return EC_OK;
  const int64_t a_rows = transa ? k : m, b_rows = transb ? n : k;
  if (lda < std::max<int64_t>(1, a_rows) || ldb < std::max<int64_t>(1, b_rows) ||
      ldc < std::max<int64_t>(1, m))
    return fail(EC_ERR_ARG, "leading dimension too small");
  if (!C || (k > 0 && alpha != 0.0 && (!A || !B))) return fail(EC_ERR_ARG, "null operand");
  DeviceGuard guard(p->device);

  if (k == 0 || alpha == 0.0) {
    const int64_t total = m * n * batch;
    const int grid = (int)std::min<int64_t>((total + 255) / 256, (int64_t)p->sm_count * 8);
    ec::scale_kernel<<<grid, 256, 0, p->stream>>>(C, m, n, ldc, stride_c, batch, beta);
    EC_CUDA(cudaGetLastError());
    p->launches++;
    p->last_kernel = "scale";
    return EC_OK;
  }

  if (p->backend == EC_BACKEND_OZAKI_I8 && ozaki_eligible(m, n, k)) {
    // k beyond what the int32 accumulators hold exactly: consecutive k-chunks, accumulated with beta = 1
    for (int64_t k0 = 0; k0 < k; k0 += OZ_MAX_K) {
      const int64_t kc = std::min(OZ_MAX_K, k - k0);
      const double *Ak = A + (transa ? k0 : k0 * lda);
      const double *Bk = B + (transb ? k0 * ldb : k0);
      const int rc = ozaki_dispatch(p, transa, transb, m, n, kc, alpha, Ak, lda, stride_a, Bk, ldb, stride_b,
                                    k0 == 0 ? beta : 1.0, C, ldc, stride_c, batch);
      if (rc) return rc;
    }
    return EC_OK;
  }

  Operand oa{A, lda, stride_a, transa == EC_OP_T, m, k};
  Operand ob{B, ldb, stride_b, transb == EC_OP_N, n, k};
  // Tiny problems are latency-bound whichever kernel runs; the TMA pipeline only pays off once a
  // problem has at least a small tile's worth of rows, columns and depth.
  const bool big_enough = (m >= 32 && n >= 32 && k >= 16) && (m * n * k >= (int64_t)64 * 64 * 64);
  const bool c_ok = m < ((int64_t)1 << 31) && n < ((int64_t)1 << 31);
  if (big_enough && c_ok && tma_ok(oa, batch) && tma_ok(ob, batch)) {
    static const char *const nL[8] = {"tma_dmma_TN", "tma_dmma_TT", "tma_dmma_NN", "tma_dmma_NT",
                                      "tma_dmma_TN_sk", "tma_dmma_TT_sk", "tma_dmma_NN_sk", "tma_dmma_NT_sk"};
    static const char *const nS[8] = {"tma_dmma_TN_s", "tma_dmma_TT_s", "tma_dmma_NN_s", "tma_dmma_NT_s",
                                      "tma_dmma_TN_s_sk", "tma_dmma_TT_s_sk", "tma_dmma_NN_s_sk", "tma_dmma_NT_s_sk"};
    static const char *const nX[8] = {"tma_dmma_TN_xs", "tma_dmma_TT_xs", "tma_dmma_NN_xs", "tma_dmma_NT_xs",
                                      "tma_dmma_TN_xs_sk", "tma_dmma_TT_xs_sk", "tma_dmma_NN_xs_sk", "tma_dmma_NT_xs_sk"};
    const TileChoice ch = choose_tile(p, m, n, k, batch);
    switch (ch.tile) {
      case TILE_S: return run_tma_dmma<ec::CfgS>(p, oa, ob, m, n, k, alpha, beta, C, ldc, stride_c, stride_a, stride_b, batch, ch.sk, nS);
      case TILE_XS: return run_tma_dmma<ec::CfgXS>(p, oa, ob, m, n, k, alpha, beta, C, ldc, stride_c, stride_a, stride_b, batch, ch.sk, nX);
      default: return run_tma_dmma<ec::CfgL>(p, oa, ob, m, n, k, alpha, beta, C, ldc, stride_c, stride_a, stride_b, batch, ch.sk, nL);
    }
  }

  ec::SimtParams sp;
  sp.M = m; sp.N = n; sp.K = k; sp.batch = batch;
  sp.alpha = alpha; sp.beta = beta;
  sp.A = A; sp.lda = lda; sp.stride_a = stride_a;
  sp.B = B; sp.ldb = ldb; sp.stride_b = stride_b;
  sp.C = C; sp.ldc = ldc; sp.stride_c = stride_c;
  sp.transa = transa; sp.transb = transb;
  sp.tiles_m = (int)((m + ec::SB - 1) / ec::SB);
  sp.tiles_n = (int)((n + ec::SB - 1) / ec::SB);
  sp.total_tiles = (int64_t)sp.tiles_m * sp.tiles_n * batch;
  const int grid = (int)std::min<int64_t>(sp.total_tiles, (int64_t)p->sm_count * 4);
  ec::dgemm_simt_kernel<<<grid, ec::SIMT_THREADS, 0, p->stream>>>(sp);
  EC_CUDA(cudaGetLastError());
  p->launches++;
  p->last_kernel = "simt";
  return EC_OK;
}

}  // namespace

// ------------------------------------------------------------------------------------------
// tensor-stream engine
// ------------------------------------------------------------------------------------------
// One job = the reference's whole streaming loop (README.md:62-66).  The tensor is cut into
// chunks of `chunk` matrices; chunk c lives in ring slot c % ring.  Per slot: a device input
// buffer, a device output buffer, (BASIS only) a device intermediate, and -- when the caller's
// memory is pageable -- a pinned input and a pinned output staging buffer.  Three streams:
//     h2d:     [wait compute(c-ring)]  copy chunk c in            -> ev_h2d[s]
//     compute: [wait ev_h2d[s], d2h(c-ring)]  batched DGEMM(s)    -> ev_cmp[s]
//     d2h:     [wait ev_cmp[s]]  copy chunk c out                 -> ev_d2h[s]
// With pinned/registered caller memory the host thread only enqueues (DMA straight from/to the
// caller's buffers).  With pageable memory a feeder thread stages chunks in (worker pool memcpy
// into pinned) and a drainer thread stages results out, so both host copies overlap the GPU.
struct StreamEngine {
  ec_pipeline *p = nullptr;
  cudaStream_t s_h2d = nullptr, s_cmp = nullptr, s_d2h = nullptr;
  int ring = 0;
  size_t in_bytes = 0, out_bytes = 0, tmp_bytes = 0, pin_in_bytes = 0, pin_out_bytes = 0;
  std::vector<double *> d_in, d_out, d_tmp;
  std::vector<void *> pin_in, pin_out;
  std::vector<cudaEvent_t> ev_h2d, ev_cmp, ev_d2h;
  double *d_fixed = nullptr;
  size_t fixed_bytes = 0;
  cudaEvent_t ev_fixed = nullptr;
  CopyPool *pool_in = nullptr, *pool_out = nullptr;
  int pool_threads = 0;

  int ensure(int ring_, size_t in_b, size_t out_b, size_t tmp_b, size_t fixed_b, bool staged_in,
             bool staged_out, int threads) {
    if (!s_h2d) {
      EC_CUDA(cudaStreamCreateWithFlags(&s_h2d, cudaStreamNonBlocking));
      EC_CUDA(cudaStreamCreateWithFlags(&s_cmp, cudaStreamNonBlocking));
      EC_CUDA(cudaStreamCreateWithFlags(&s_d2h, cudaStreamNonBlocking));
      EC_CUDA(cudaEventCreateWithFlags(&ev_fixed, cudaEventDisableTiming));
    }
    if (ring_ != ring) {
      int rc = release_buffers();
      if (rc) return rc;
      ring = ring_;
      d_in.assign(ring, nullptr);
      d_out.assign(ring, nullptr);
      d_tmp.assign(ring, nullptr);
      pin_in.assign(ring, nullptr);
      pin_out.assign(ring, nullptr);
      ev_h2d.assign(ring, nullptr);
      ev_cmp.assign(ring, nullptr);
      ev_d2h.assign(ring, nullptr);
      for (int s = 0; s < ring; ++s) {
        EC_CUDA(cudaEventCreateWithFlags(&ev_h2d[s], cudaEventDisableTiming));
        EC_CUDA(cudaEventCreateWithFlags(&ev_cmp[s], cudaEventDisableTiming));
        EC_CUDA(cudaEventCreateWithFlags(&ev_d2h[s], cudaEventDisableTiming));
      }
    }
    auto grow_dev = [&](std::vector<double *> &v, size_t &have, size_t want) -> int {
      if (want <= have) return EC_OK;
      for (int s = 0; s < ring; ++s) {
        if (v[s]) EC_CUDA(cudaFree(v[s]));
        v[s] = nullptr;
      }
      have = 0;
      for (int s = 0; s < ring; ++s) EC_CUDA(cudaMalloc((void **)&v[s], want));
      have = want;
      return EC_OK;
    };
    auto grow_pin = [&](std::vector<void *> &v, size_t &have, size_t want) -> int {
      if (want <= have) return EC_OK;
      for (int s = 0; s < ring; ++s) {
        if (v[s]) EC_CUDA(cudaFreeHost(v[s]));
        v[s] = nullptr;
      }
      have = 0;
      for (int s = 0; s < ring; ++s) EC_CUDA(cudaHostAlloc(&v[s], want, cudaHostAllocDefault));
      have = want;
      return EC_OK;
    };
    int rc;
    if ((rc = grow_dev(d_in, in_bytes, in_b))) return rc;
    if ((rc = grow_dev(d_out, out_bytes, out_b))) return rc;
    if (tmp_b && (rc = grow_dev(d_tmp, tmp_bytes, tmp_b))) return rc;
    if (staged_in && (rc = grow_pin(pin_in, pin_in_bytes, in_b))) return rc;
    if (staged_out && (rc = grow_pin(pin_out, pin_out_bytes, out_b))) return rc;
    if (fixed_b > fixed_bytes) {
      if (d_fixed) EC_CUDA(cudaFree(d_fixed));
      d_fixed = nullptr;
      fixed_bytes = 0;
      EC_CUDA(cudaMalloc((void **)&d_fixed, fixed_b));
      fixed_bytes = fixed_b;
    }
    if ((staged_in || staged_out) && (!pool_in || pool_threads != threads)) {
      delete pool_in;
      delete pool_out;
      pool_in = new CopyPool(std::max(1, threads / 2));
      pool_out = new CopyPool(std::max(1, threads - threads / 2));
      pool_threads = threads;
    }
    return EC_OK;
  }

  int release_buffers() {
    for (auto &x : d_in) if (x) { EC_CUDA(cudaFree(x)); x = nullptr; }
    for (auto &x : d_out) if (x) { EC_CUDA(cudaFree(x)); x = nullptr; }
    for (auto &x : d_tmp) if (x) { EC_CUDA(cudaFree(x)); x = nullptr; }
    for (auto &x : pin_in) if (x) { EC_CUDA(cudaFreeHost(x)); x = nullptr; }
    for (auto &x : pin_out) if (x) { EC_CUDA(cudaFreeHost(x)); x = nullptr; }
    for (auto &e : ev_h2d) if (e) { cudaEventDestroy(e); e = nullptr; }
    for (auto &e : ev_cmp) if (e) { cudaEventDestroy(e); e = nullptr; }
    for (auto &e : ev_d2h) if (e) { cudaEventDestroy(e); e = nullptr; }
    in_bytes = out_bytes = tmp_bytes = pin_in_bytes = pin_out_bytes = 0;
    return EC_OK;
  }

  void destroy() {
    release_buffers();
    if (d_fixed) cudaFree(d_fixed);
    d_fixed = nullptr;
    if (ev_fixed) cudaEventDestroy(ev_fixed);
    if (s_h2d) cudaStreamDestroy(s_h2d);
    if (s_cmp) cudaStreamDestroy(s_cmp);
    if (s_d2h) cudaStreamDestroy(s_d2h);
    delete pool_in;
    delete pool_out;
    pool_in = pool_out = nullptr;
  }
};

namespace {

struct StreamShape {
  int64_t r_rows, r_cols;  // result shape
};

int stream_shape(int kind, int64_t f_rows, int64_t f_cols, int64_t t_rows, int64_t t_cols,
                 StreamShape *out) {
  switch (kind) {
    case EC_STREAM_RIGHT:  // T(t_rows x t_cols) * F(f_rows x f_cols)
      if (t_cols != f_rows) return fail(EC_ERR_SHAPE, "Shape mismatch in gemm: T.cols != F.rows");
      *out = {t_rows, f_cols};
      return EC_OK;
    case EC_STREAM_LEFT:
      if (f_cols != t_rows) return fail(EC_ERR_SHAPE, "Shape mismatch in gemm: F.cols != T.rows");
      *out = {f_rows, t_cols};
      return EC_OK;
    case EC_STREAM_LEFT_T:
      if (f_rows != t_rows) return fail(EC_ERR_SHAPE, "Shape mismatch in gemm: F^T.cols != T.rows");
      *out = {f_cols, t_cols};
      return EC_OK;
    case EC_STREAM_BASIS:
      if (f_rows != t_rows || t_cols != f_rows)
        return fail(EC_ERR_SHAPE, "Shape mismatch in basis transform F^T * T * F");
      *out = {f_cols, f_cols};
      return EC_OK;
  }
  return fail(EC_ERR_ARG, "unknown tensor-stream kind %d", kind);
}

// Launch the product(s) for `nb` matrices at d_in -> d_out on the pipeline's *current* stream.
int stream_compute(ec_pipeline *p, int kind, const double *dF, int64_t f_rows, int64_t f_cols,
                   const double *d_in, double *d_out, double *d_tmp, int64_t nb, int64_t t_rows,
                   int64_t t_cols, const StreamShape &sh) {
  const int64_t t_sz = t_rows * t_cols, r_sz = sh.r_rows * sh.r_cols;
  switch (kind) {
    case EC_STREAM_RIGHT:
      return dgemm_dispatch(p, EC_OP_N, EC_OP_N, t_rows, f_cols, t_cols, 1.0, d_in, t_rows, t_sz,
                            dF, f_rows, 0, 0.0, d_out, sh.r_rows, r_sz, nb);
    case EC_STREAM_LEFT:
      return dgemm_dispatch(p, EC_OP_N, EC_OP_N, f_rows, t_cols, f_cols, 1.0, dF, f_rows, 0, d_in,
                            t_rows, t_sz, 0.0, d_out, sh.r_rows, r_sz, nb);
    case EC_STREAM_LEFT_T:
      return dgemm_dispatch(p, EC_OP_T, EC_OP_N, f_cols, t_cols, f_rows, 1.0, dF, f_rows, 0, d_in,
                            t_rows, t_sz, 0.0, d_out, sh.r_rows, r_sz, nb);
    case EC_STREAM_BASIS: {
      const int64_t tmp_sz = f_cols * t_cols;
      int rc = dgemm_dispatch(p, EC_OP_T, EC_OP_N, f_cols, t_cols, f_rows, 1.0, dF, f_rows, 0,
                              d_in, t_rows, t_sz, 0.0, d_tmp, f_cols, tmp_sz, nb);
      if (rc) return rc;
      return dgemm_dispatch(p, EC_OP_N, EC_OP_N, f_cols, f_cols, t_cols, 1.0, d_tmp, f_cols,
                            tmp_sz, dF, f_rows, 0, 0.0, d_out, sh.r_rows, r_sz, nb);
    }
  }
  return fail(EC_ERR_ARG, "unknown tensor-stream kind %d", kind);
}

// Copy matrices [i0, i0+nb) between per-matrix host pointers and one contiguous device chunk,
// merging runs of adjacent host matrices into single DMA transfers.
int dma_chunk(bool to_device, double *dev, const double *const *host, int64_t i0, int64_t nb,
              int64_t elems, cudaStream_t st) {
  int64_t i = 0;
  while (i < nb) {
    int64_t j = i + 1;
    while (j < nb && host[i0 + j] == host[i0 + j - 1] + elems) ++j;
    const size_t bytes = (size_t)(j - i) * elems * 8;
    if (to_device)
      EC_CUDA(cudaMemcpyAsync(dev + i * elems, host[i0 + i], bytes, cudaMemcpyHostToDevice, st));
    else
      EC_CUDA(cudaMemcpyAsync((void *)host[i0 + i], dev + i * elems, bytes, cudaMemcpyDeviceToHost, st));
    i = j;
  }
  return EC_OK;
}

}  // namespace

// ------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------
extern "C" {

const char *ec_version(void) { return "eigencuda_b200 0.1.0 (sm_100a, FP64 DMMA)"; }
const char *ec_last_error_string(void) { return g_last_error.c_str(); }

int64_t ec_device_count(void) {
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return count;
}

static int pipeline_create_impl(int device, cudaStream_t external, bool use_external,
                                ec_pipeline **out) {
  if (!out) return fail(EC_ERR_ARG, "null out pointer");
  *out = nullptr;
  int dev = device;
  if (dev < 0) EC_CUDA(cudaGetDevice(&dev));
  int count = 0;
  EC_CUDA(cudaGetDeviceCount(&count));
  if (dev >= count) return fail(EC_ERR_ARG, "device %d out of range (%d devices)", dev, count);
  DeviceGuard guard(dev);
  cudaDeviceProp prop;
  EC_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10)
    return fail(EC_ERR_UNSUPPORTED,
                "device %d is sm_%d%d; this library contains sm_100a code only (B200)", dev,
                prop.major, prop.minor);
  ec_pipeline *p = new ec_pipeline();
  p->device = dev;
  p->sm_count = prop.multiProcessorCount;
  if (use_external) {
    p->stream = external;
    p->owns_stream = false;
  } else {
    // A blocking stream, like the reference's cudaStreamCreate (include/cudapipeline.hpp:24),
    // so code that also uses the legacy NULL stream keeps its implicit ordering.
    cudaError_t e = cudaStreamCreate(&p->stream);
    if (e != cudaSuccess) {
      delete p;
      return fail(EC_ERR_CUDA, "cudaStreamCreate failed: %s", cudaGetErrorString(e));
    }
  }
  {
    std::lock_guard<std::mutex> lk(g_registry_mu);
    g_registry[(void *)p->stream] = p;
  }
  *out = p;
  return EC_OK;
}

int ec_pipeline_create(int device, ec_pipeline **out) {
  return pipeline_create_impl(device, nullptr, false, out);
}
int ec_pipeline_create_on_stream(int device, void *cuda_stream, ec_pipeline **out) {
  return pipeline_create_impl(device, (cudaStream_t)cuda_stream, true, out);
}

int ec_pipeline_destroy(ec_pipeline *p) {
  if (!p) return EC_OK;
  DeviceGuard guard(p->device);
  cudaStreamSynchronize(p->stream);
  {
    std::lock_guard<std::mutex> lk(g_registry_mu);
    auto it = g_registry.find((void *)p->stream);
    if (it != g_registry.end() && it->second == p) g_registry.erase(it);
  }
  if (p->engine) {
    p->engine->destroy();
    delete p->engine;
  }
  p->staging.destroy();
  if (p->sk_workspace) cudaFree(p->sk_workspace);
  if (p->sk_flags) cudaFree(p->sk_flags);
  if (p->oz_planes_a) cudaFree(p->oz_planes_a);
  if (p->oz_planes_b) cudaFree(p->oz_planes_b);
  if (p->oz_scale_a) cudaFree(p->oz_scale_a);
  if (p->oz_scale_b) cudaFree(p->oz_scale_b);
  if (p->oz_amax) cudaFree(p->oz_amax);
  if (p->oz_split_stream) cudaStreamDestroy(p->oz_split_stream);
  for (int i = 0; i < 2; ++i) {
    if (p->oz_ev_split[i]) cudaEventDestroy(p->oz_ev_split[i]);
    if (p->oz_ev_gemm[i]) cudaEventDestroy(p->oz_ev_gemm[i]);
  }
  if (p->owns_stream) cudaStreamDestroy(p->stream);
  delete p;
  return EC_OK;
}

void *ec_pipeline_stream(const ec_pipeline *p) { return p ? (void *)p->stream : nullptr; }
const void *ec_pipeline_stream_ref(const ec_pipeline *p) { return p ? (const void *)&p->stream : nullptr; }
int ec_pipeline_device(const ec_pipeline *p) { return p ? p->device : -1; }

int ec_pipeline_synchronize(ec_pipeline *p) {
  if (!p) return fail(EC_ERR_ARG, "null pipeline");
  EC_CUDA(cudaStreamSynchronize(p->stream));
  return EC_OK;
}

int ec_pipeline_set_backend(ec_pipeline *p, int backend) {
  if (!p) return fail(EC_ERR_ARG, "null pipeline");
  if (backend != EC_BACKEND_FP64_DMMA && backend != EC_BACKEND_OZAKI_I8)
    return fail(EC_ERR_UNSUPPORTED, "unknown backend %d", backend);
  p->backend = backend;
  return EC_OK;
}
int ec_pipeline_get_backend(const ec_pipeline *p) { return p ? p->backend : -1; }
int ec_pipeline_set_schedule(ec_pipeline *p, int stream_k, int one_cta_per_tile) {
  if (!p) return fail(EC_ERR_ARG, "null pipeline");
  p->stream_k = stream_k ? 1 : 0;
  p->one_cta_per_tile = one_cta_per_tile ? 1 : 0;
  return EC_OK;
}
int ec_pipeline_set_sm_margin(ec_pipeline *p, int sms) {
  if (!p) return fail(EC_ERR_ARG, "null pipeline");
  if (sms < 0 || sms >= p->sm_count) return fail(EC_ERR_ARG, "sm margin out of range");
  p->sm_margin = sms;
  return EC_OK;
}
int64_t ec_pipeline_launch_count(const ec_pipeline *p) { return p ? p->launches : 0; }
const char *ec_pipeline_last_kernel(const ec_pipeline *p) { return p ? p->last_kernel : "none"; }

int ec_pipeline_set_stream_config(ec_pipeline *p, int chunk_matrices, int ring_depth,
                                  int staging_threads) {
  if (!p) return fail(EC_ERR_ARG, "null pipeline");
  if (chunk_matrices < 0 || ring_depth < 0 || staging_threads < 0 || ring_depth == 1)
    return fail(EC_ERR_ARG, "bad stream config");
  p->cfg_chunk = chunk_matrices;
  p->cfg_ring = ring_depth;
  p->cfg_threads = staging_threads;
  return EC_OK;
}

// ---- matrices ------------------------------------------------------------------------------

int ec_matrix_alloc(void *cuda_stream, int64_t rows, int64_t cols, ec_matrix **out) {
  if (!out) return fail(EC_ERR_ARG, "null out pointer");
  *out = nullptr;
  if (rows < 0 || cols < 0) return fail(EC_ERR_ARG, "negative matrix dimension");
  cudaStream_t st = (cudaStream_t)cuda_stream;
  ec_pipeline *p = pipeline_of_stream(st);
  int dev = 0;
  if (p)
    dev = p->device;
  else
    EC_CUDA(cudaGetDevice(&dev));
  DeviceGuard guard(dev);
  const size_t bytes = (size_t)rows * (size_t)cols * sizeof(double);
  // The reference pre-checks free memory on every allocation and throws when the request does
  // not fit (src/cudamatrix.cc:63-79).  Same observable behaviour, but the query (a device
  // synchronising call) is only made when cudaMalloc has actually failed.
  double *d = nullptr;
  if (bytes > 0) {
    cudaError_t e = cudaMalloc((void **)&d, bytes);
    if (e != cudaSuccess) {
      cudaGetLastError();
      size_t free_b = 0, total_b = 0;
      cudaMemGetInfo(&free_b, &total_b);
      return fail(EC_ERR_NOMEM,
                  "There were requested : %zu bytes in the device\nDevice Free memory (bytes): %zu\n"
                  "Device total Memory (bytes): %zu\nThere is not enough memory in the Device!",
                  bytes, free_b, total_b);
    }
  }
  ec_matrix *m = new ec_matrix();
  m->data = d;
  m->rows = rows;
  m->cols = cols;
  m->stream = st;
  m->device = dev;
  *out = m;
  return EC_OK;
}

int ec_matrix_free(ec_matrix *m) {
  if (!m) return EC_OK;
  if (m->data) {
    DeviceGuard guard(m->device);
    cudaFree(m->data);  // synchronises with outstanding work on the buffer
  }
  delete m;
  return EC_OK;
}

int ec_matrix_upload(ec_matrix *m, const double *host, int64_t count) {
  if (!m) return fail(EC_ERR_ARG, "null matrix");
  if (count < 0 || count > m->rows * m->cols)
    return fail(EC_ERR_ARG, "copy_to_gpu: source has %lld elements, device matrix holds %lld",
                (long long)count, (long long)(m->rows * m->cols));
  if (count == 0) return EC_OK;
  if (!host) return fail(EC_ERR_ARG, "null host pointer");
  DeviceGuard guard(m->device);
  const size_t bytes = (size_t)count * sizeof(double);
  ec_pipeline *p = pipeline_of_stream(m->stream);
  if (p && !is_pinned_host(host)) {
    int slot = -1;
    int rc = p->staging.acquire(bytes, &slot);
    if (rc) return rc;
    std::memcpy(p->staging.buf[slot], host, bytes);
    cudaError_t e =
        cudaMemcpyAsync(m->data, p->staging.buf[slot], bytes, cudaMemcpyHostToDevice, m->stream);
    if (e != cudaSuccess) return fail(EC_ERR_COPY, "Error copy arrays to device: %s", cudaGetErrorString(e));
    return p->staging.release_after(slot, m->stream);
  }
  cudaError_t e = cudaMemcpyAsync(m->data, host, bytes, cudaMemcpyHostToDevice, m->stream);
  if (e != cudaSuccess) return fail(EC_ERR_COPY, "Error copy arrays to device: %s", cudaGetErrorString(e));
  // A pinned source is read asynchronously, exactly as cudaMemcpyAsync does in the reference;
  // only pageable sources carry the "reusable on return" guarantee.
  return EC_OK;
}

int ec_matrix_create_from_host(void *cuda_stream, const double *host, int64_t rows, int64_t cols,
                               ec_matrix **out) {
  int rc = ec_matrix_alloc(cuda_stream, rows, cols, out);
  if (rc) return rc;
  rc = ec_matrix_upload(*out, host, rows * cols);
  if (rc) {
    ec_matrix_free(*out);
    *out = nullptr;
    if (rc == EC_ERR_CUDA) rc = EC_ERR_COPY;
  }
  return rc;
}

int ec_matrix_download(const ec_matrix *m, double *host) {
  if (!m) return fail(EC_ERR_ARG, "null matrix");
  const int64_t count = m->rows * m->cols;
  if (count == 0) return EC_OK;
  if (!host) return fail(EC_ERR_ARG, "null host pointer");
  DeviceGuard guard(m->device);
  const size_t bytes = (size_t)count * sizeof(double);
  ec_pipeline *p = pipeline_of_stream(m->stream);
  constexpr size_t CHUNK = (size_t)4 << 20;
  if (p && bytes > CHUNK && !is_pinned_host(host)) {
    // Pageable destination: DMA into two pinned slots alternately and copy out behind the DMA.
    size_t off = 0;
    int prev_slot = -1;
    size_t prev_off = 0, prev_len = 0;
    while (off < bytes || prev_slot >= 0) {
      int slot = -1;
      size_t len = 0;
      if (off < bytes) {
        len = std::min(CHUNK, bytes - off);
        int rc = p->staging.acquire(len, &slot);
        if (rc) return rc;
        EC_CUDA(cudaMemcpyAsync(p->staging.buf[slot], (const char *)m->data + off, len,
                                cudaMemcpyDeviceToHost, m->stream));
        rc = p->staging.release_after(slot, m->stream);
        if (rc) return rc;
      }
      if (prev_slot >= 0) {
        EC_CUDA(cudaEventSynchronize(p->staging.ev[prev_slot]));
        p->staging.busy[prev_slot] = false;
        std::memcpy((char *)host + prev_off, p->staging.buf[prev_slot], prev_len);
      }
      prev_slot = slot;
      prev_off = off;
      prev_len = len;
      off += len;
    }
    return EC_OK;
  }
  EC_CUDA(cudaMemcpyAsync(host, m->data, bytes, cudaMemcpyDeviceToHost, m->stream));
  EC_CUDA(cudaStreamSynchronize(m->stream));
  return EC_OK;
}

int64_t ec_matrix_rows(const ec_matrix *m) { return m ? m->rows : 0; }
int64_t ec_matrix_cols(const ec_matrix *m) { return m ? m->cols : 0; }
int64_t ec_matrix_size(const ec_matrix *m) { return m ? m->rows * m->cols : 0; }
double *ec_matrix_data(const ec_matrix *m) { return m ? m->data : nullptr; }

// ---- gemm ----------------------------------------------------------------------------------

int ec_gemm(ec_pipeline *p, const ec_matrix *A, const ec_matrix *B, ec_matrix *C) {
  if (!p || !A || !B || !C) return fail(EC_ERR_ARG, "null handle");
  if (A->cols != B->rows) return fail(EC_ERR_SHAPE, "Shape mismatch in Cublas gemm");
  // Operands created on another stream than the pipeline's: the reference is ordered by the
  // legacy NULL stream in that case; here the dependency is made explicit.
  const ec_matrix *ops[3] = {A, B, C};
  for (const ec_matrix *m : ops) {
    if (m->stream != p->stream) {
      cudaEvent_t ev;
      EC_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
      EC_CUDA(cudaEventRecord(ev, m->stream));
      EC_CUDA(cudaStreamWaitEvent(p->stream, ev, 0));
      EC_CUDA(cudaEventDestroy(ev));
    }
  }
  int rc = dgemm_dispatch(p, EC_OP_N, EC_OP_N, A->rows, B->cols, A->cols, 1.0, A->data, A->rows, 0,
                          B->data, B->rows, 0, 0.0, C->data, std::max<int64_t>(C->rows, 1), 0, 1);
  if (rc) return rc;
  if (C->stream != p->stream) {
    cudaEvent_t ev;
    EC_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    EC_CUDA(cudaEventRecord(ev, p->stream));
    EC_CUDA(cudaStreamWaitEvent(C->stream, ev, 0));
    EC_CUDA(cudaEventDestroy(ev));
  }
  return EC_OK;
}

int ec_dgemm(ec_pipeline *p, int transa, int transb, int64_t m, int64_t n, int64_t k, double alpha,
             const double *A, int64_t lda, const double *B, int64_t ldb, double beta, double *C,
             int64_t ldc) {
  return dgemm_dispatch(p, transa, transb, m, n, k, alpha, A, lda, 0, B, ldb, 0, beta, C, ldc, 0, 1);
}

int ec_dgemm_strided_batched(ec_pipeline *p, int transa, int transb, int64_t m, int64_t n,
                             int64_t k, double alpha, const double *A, int64_t lda,
                             int64_t stride_a, const double *B, int64_t ldb, int64_t stride_b,
                             double beta, double *C, int64_t ldc, int64_t stride_c, int64_t batch) {
  if (stride_a < 0 || stride_b < 0 || stride_c < 0) return fail(EC_ERR_ARG, "negative stride");
  return dgemm_dispatch(p, transa, transb, m, n, k, alpha, A, lda, stride_a, B, ldb, stride_b, beta,
                        C, ldc, stride_c, batch);
}

// ---- tensor stream -------------------------------------------------------------------------

int ec_tensor_stream_run_device(ec_pipeline *p, int kind, const double *F_dev, int64_t f_rows,
                                int64_t f_cols, const double *tensor_dev, double *results_dev,
                                int64_t count, int64_t t_rows, int64_t t_cols) {
  if (!p) return fail(EC_ERR_ARG, "null pipeline");
  if (count < 0 || f_rows < 0 || f_cols < 0 || t_rows < 0 || t_cols < 0)
    return fail(EC_ERR_ARG, "negative size");
  StreamShape sh;
  int rc = stream_shape(kind, f_rows, f_cols, t_rows, t_cols, &sh);
  if (rc) return rc;
  if (count == 0) return EC_OK;
  DeviceGuard guard(p->device);
  double *tmp = nullptr;
  if (kind == EC_STREAM_BASIS) {
    // intermediate F^T * T_i for a bounded chunk at a time
    const int64_t tmp_sz = f_cols * t_cols;
    const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(count, ((int64_t)256 << 20) / std::max<int64_t>(8 * tmp_sz, 1)));
    EC_CUDA(cudaMallocAsync((void **)&tmp, (size_t)(chunk * tmp_sz * 8), p->stream));
    for (int64_t i0 = 0; i0 < count && !rc; i0 += chunk) {
      const int64_t nb = std::min(chunk, count - i0);
      rc = stream_compute(p, kind, F_dev, f_rows, f_cols, tensor_dev + i0 * t_rows * t_cols,
                          results_dev + i0 * sh.r_rows * sh.r_cols, tmp, nb, t_rows, t_cols, sh);
    }
    cudaFreeAsync(tmp, p->stream);
    return rc;
  }
  return stream_compute(p, kind, F_dev, f_rows, f_cols, tensor_dev, results_dev, nullptr, count,
                        t_rows, t_cols, sh);
}

int ec_tensor_stream_run(ec_pipeline *p, int kind, const double *F, int64_t f_rows, int64_t f_cols,
                         const double *const *tensor, double *const *results, int64_t count,
                         int64_t t_rows, int64_t t_cols) {
  if (!p) return fail(EC_ERR_ARG, "null pipeline");
  if (count < 0 || f_rows < 0 || f_cols < 0 || t_rows < 0 || t_cols < 0)
    return fail(EC_ERR_ARG, "negative size");
  StreamShape sh;
  int rc = stream_shape(kind, f_rows, f_cols, t_rows, t_cols, &sh);  // before any device work
  if (rc) return rc;
  if (count == 0) return EC_OK;
  if (!F || !tensor || !results) return fail(EC_ERR_ARG, "null pointer");
  const int64_t t_sz = t_rows * t_cols, r_sz = sh.r_rows * sh.r_cols, f_sz = f_rows * f_cols;
  if (t_sz == 0 || r_sz == 0 || f_sz == 0) {
    // degenerate: products with an empty inner dimension are all-zero
    for (int64_t i = 0; i < count; ++i)
      if (r_sz) std::memset(results[i], 0, (size_t)r_sz * 8);
    return EC_OK;
  }
  DeviceGuard guard(p->device);

  // ---- plan ----
  bool in_pinned = true, out_pinned = true;
  // One attribute query per run of adjacent matrices (first and last element of the run), not
  // per matrix: a contiguous tensor costs two queries.
  auto all_pinned = [&](const double *const *ptrs, int64_t elems) {
    int64_t i = 0;
    while (i < count) {
      int64_t j = i + 1;
      while (j < count && ptrs[j] == ptrs[j - 1] + elems) ++j;
      if (!is_pinned_host(ptrs[i]) || !is_pinned_host(ptrs[j - 1] + elems - 1)) return false;
      i = j;
    }
    return true;
  };
  in_pinned = all_pinned(tensor, t_sz);
  out_pinned = all_pinned((const double *const *)results, r_sz);
  const bool staged_in = !in_pinned, staged_out = !out_pinned;
  int64_t chunk = p->cfg_chunk;
  if (chunk <= 0) chunk = std::max<int64_t>(1, ((int64_t)32 << 20) / (8 * std::max(t_sz, r_sz)));
  chunk = std::min(chunk, count);
  const int ring = p->cfg_ring > 0 ? p->cfg_ring : 3;
  int threads = p->cfg_threads;
  if (threads <= 0) threads = std::max(2u, std::min(16u, std::thread::hardware_concurrency()));
  if (!p->engine) {
    p->engine = new StreamEngine();
    p->engine->p = p;
  }
  StreamEngine &E = *p->engine;
  const bool f_on_device = is_device_ptr(F);
  const size_t tmp_b = kind == EC_STREAM_BASIS ? (size_t)(chunk * f_cols * t_cols * 8) : 0;
  rc = E.ensure(ring, (size_t)(chunk * t_sz * 8), (size_t)(chunk * r_sz * 8), tmp_b,
                f_on_device ? 0 : (size_t)(f_sz * 8), staged_in, staged_out, threads);
  if (rc) return rc;

  // Order the engine's streams after whatever the caller queued on the pipeline stream.
  EC_CUDA(cudaEventRecord(E.ev_fixed, p->stream));
  EC_CUDA(cudaStreamWaitEvent(E.s_h2d, E.ev_fixed, 0));
  EC_CUDA(cudaStreamWaitEvent(E.s_cmp, E.ev_fixed, 0));

  // ---- fixed operand: uploaded once (or used in place when already on the device, e.g. after
  // the multi-GPU broadcast) ----
  const double *dF = F;
  if (!f_on_device) {
    int slot = -1;
    if (is_pinned_host(F)) {
      EC_CUDA(cudaMemcpyAsync(E.d_fixed, F, (size_t)f_sz * 8, cudaMemcpyHostToDevice, E.s_h2d));
    } else {
      rc = p->staging.acquire((size_t)f_sz * 8, &slot);
      if (rc) return rc;
      std::memcpy(p->staging.buf[slot], F, (size_t)f_sz * 8);
      EC_CUDA(cudaMemcpyAsync(E.d_fixed, p->staging.buf[slot], (size_t)f_sz * 8,
                              cudaMemcpyHostToDevice, E.s_h2d));
      rc = p->staging.release_after(slot, E.s_h2d);
      if (rc) return rc;
    }
    dF = E.d_fixed;
  }
  EC_CUDA(cudaEventRecord(E.ev_fixed, E.s_h2d));
  EC_CUDA(cudaStreamWaitEvent(E.s_cmp, E.ev_fixed, 0));

  const int64_t nchunks = (count + chunk - 1) / chunk;

  // ---- drainer (pageable results only) ----
  std::mutex mu;
  std::condition_variable cv;
  int64_t enqueued = 0, drained = 0;  // chunks whose D2H is enqueued / whose results are delivered
  bool abort_flag = false;
  std::atomic<int> drain_rc{EC_OK};
  std::thread drainer;
  if (staged_out) {
    drainer = std::thread([&] {
      cudaSetDevice(p->device);
      for (int64_t c = 0; c < nchunks; ++c) {
        {
          std::unique_lock<std::mutex> lk(mu);
          cv.wait(lk, [&] { return enqueued > c || abort_flag; });
          if (abort_flag) return;
        }
        const int s = (int)(c % ring);
        if (cudaEventSynchronize(E.ev_d2h[s]) != cudaSuccess) {
          drain_rc = EC_ERR_CUDA;
        } else {
          const int64_t i0 = c * chunk, nb = std::min(chunk, count - i0);
          const char *src = (const char *)E.pin_out[s];
          // split every matrix into a few pieces so the pool stays busy even for small chunks
          const int pieces = std::max<int>(1, (int)std::min<int64_t>(8, (r_sz * 8) >> 18));
          E.pool_out->parallel_for(nb * pieces, [&](int64_t t) {
            const int64_t i = t / pieces, q = t % pieces;
            const size_t lo = (size_t)(r_sz * 8) * q / pieces, hi = (size_t)(r_sz * 8) * (q + 1) / pieces;
            std::memcpy((char *)results[i0 + i] + lo, src + (size_t)i * r_sz * 8 + lo, hi - lo);
          });
        }
        {
          std::lock_guard<std::mutex> lk(mu);
          drained = c + 1;
        }
        cv.notify_all();
      }
    });
  }

  // ---- feeder (this thread) ----
  auto feeder = [&]() -> int {
    for (int64_t c = 0; c < nchunks; ++c) {
      const int s = (int)(c % ring);
      const int64_t i0 = c * chunk, nb = std::min(chunk, count - i0);
      // H2D
      if (staged_in) {
        if (c >= ring) EC_CUDA(cudaEventSynchronize(E.ev_h2d[s]));  // pinned_in[s] consumed
        char *dst = (char *)E.pin_in[s];
        const int pieces = std::max<int>(1, (int)std::min<int64_t>(8, (t_sz * 8) >> 18));
        E.pool_in->parallel_for(nb * pieces, [&](int64_t t) {
          const int64_t i = t / pieces, q = t % pieces;
          const size_t lo = (size_t)(t_sz * 8) * q / pieces, hi = (size_t)(t_sz * 8) * (q + 1) / pieces;
          std::memcpy(dst + (size_t)i * t_sz * 8 + lo, (const char *)tensor[i0 + i] + lo, hi - lo);
        });
      }
      if (c >= ring) EC_CUDA(cudaStreamWaitEvent(E.s_h2d, E.ev_cmp[s], 0));  // d_in[s] consumed
      if (staged_in)
        EC_CUDA(cudaMemcpyAsync(E.d_in[s], E.pin_in[s], (size_t)(nb * t_sz * 8),
                                cudaMemcpyHostToDevice, E.s_h2d));
      else if ((rc = dma_chunk(true, E.d_in[s], tensor, i0, nb, t_sz, E.s_h2d)))
        return rc;
      EC_CUDA(cudaEventRecord(E.ev_h2d[s], E.s_h2d));
      // compute (on the engine's compute stream: temporarily retarget the pipeline)
      EC_CUDA(cudaStreamWaitEvent(E.s_cmp, E.ev_h2d[s], 0));
      if (c >= ring) EC_CUDA(cudaStreamWaitEvent(E.s_cmp, E.ev_d2h[s], 0));  // d_out[s] drained
      {
        cudaStream_t saved = p->stream;
        p->stream = E.s_cmp;
        rc = stream_compute(p, kind, dF, f_rows, f_cols, E.d_in[s], E.d_out[s],
                            kind == EC_STREAM_BASIS ? E.d_tmp[s] : nullptr, nb, t_rows, t_cols, sh);
        p->stream = saved;
        if (rc) return rc;
      }
      EC_CUDA(cudaEventRecord(E.ev_cmp[s], E.s_cmp));
      // D2H
      EC_CUDA(cudaStreamWaitEvent(E.s_d2h, E.ev_cmp[s], 0));
      if (staged_out) {
        if (c >= ring) {  // pin_out[s] must have been delivered to the caller
          std::unique_lock<std::mutex> lk(mu);
          cv.wait(lk, [&] { return drained >= c - ring + 1 || drain_rc != EC_OK; });
        }
        EC_CUDA(cudaMemcpyAsync(E.pin_out[s], E.d_out[s], (size_t)(nb * r_sz * 8),
                                cudaMemcpyDeviceToHost, E.s_d2h));
      } else if ((rc = dma_chunk(false, E.d_out[s], (const double *const *)results, i0, nb, r_sz,
                                 E.s_d2h))) {
        return rc;
      }
      EC_CUDA(cudaEventRecord(E.ev_d2h[s], E.s_d2h));
      if (staged_out) {
        {
          std::lock_guard<std::mutex> lk(mu);
          enqueued = c + 1;
        }
        cv.notify_all();
      }
    }
    return EC_OK;
  };
  rc = feeder();
  if (rc) {
    std::string keep = g_last_error;
    {
      std::lock_guard<std::mutex> lk(mu);
      abort_flag = true;
    }
    cv.notify_all();
    if (drainer.joinable()) drainer.join();
    cudaStreamSynchronize(E.s_h2d);
    cudaStreamSynchronize(E.s_cmp);
    cudaStreamSynchronize(E.s_d2h);
    g_last_error = keep;
    return rc;
  }
  if (drainer.joinable()) drainer.join();
  EC_CUDA(cudaStreamSynchronize(E.s_d2h));
  EC_CUDA(cudaStreamSynchronize(E.s_cmp));
  EC_CUDA(cudaStreamSynchronize(E.s_h2d));
  if (drain_rc != EC_OK) return fail(EC_ERR_CUDA, "device-to-host drain failed");
  return EC_OK;
}

// ---- host memory + synthetic data ----------------------------------------------------------

int ec_host_alloc(size_t bytes, void **out) {
  if (!out) return fail(EC_ERR_ARG, "null out pointer");
  *out = nullptr;
  if (bytes == 0) return EC_OK;
  EC_CUDA(cudaHostAlloc(out, bytes, cudaHostAllocPortable));
  return EC_OK;
}
int ec_host_free(void *ptr) {
  if (ptr) EC_CUDA(cudaFreeHost(ptr));
  return EC_OK;
}

int ec_fill_uniform_batched(ec_pipeline *p, uint64_t seed, uint64_t first_id, double *dst,
                            int64_t count, int64_t batch) {
  if (!p) return fail(EC_ERR_ARG, "null pipeline");
  if (count < 0 || batch < 0) return fail(EC_ERR_ARG, "negative size");
  if (count == 0 || batch == 0) return EC_OK;
  if (!dst) return fail(EC_ERR_ARG, "null destination");
  DeviceGuard guard(p->device);
  const int64_t total = count * batch;
  const int grid = (int)std::min<int64_t>((total + 255) / 256, (int64_t)p->sm_count * 16);
  ec::fill_uniform_kernel<<<grid, 256, 0, p->stream>>>(seed, first_id, dst, count, batch);
  EC_CUDA(cudaGetLastError());
  p->launches++;
  return EC_OK;
}
int ec_fill_uniform_block(ec_pipeline *p, uint64_t seed, uint64_t matrix_id, double *dst,
                          int64_t rows, int64_t cols, int64_t ld_dst, int64_t row0, int64_t col0,
                          int64_t global_ld) {
  if (!p) return fail(EC_ERR_ARG, "null pipeline");
  if (rows < 0 || cols < 0 || row0 < 0 || col0 < 0 || ld_dst < rows || global_ld < row0 + rows)
    return fail(EC_ERR_ARG, "bad block geometry");
  if (rows == 0 || cols == 0) return EC_OK;
  if (!dst) return fail(EC_ERR_ARG, "null destination");
  DeviceGuard guard(p->device);
  const int64_t total = rows * cols;
  const int grid = (int)std::min<int64_t>((total + 255) / 256, (int64_t)p->sm_count * 16);
  ec::fill_uniform_block_kernel<<<grid, 256, 0, p->stream>>>(seed, matrix_id, dst, rows, cols, ld_dst,
                                                            row0, col0, global_ld);
  EC_CUDA(cudaGetLastError());
  p->launches++;
  return EC_OK;
}
int ec_fill_uniform(ec_pipeline *p, uint64_t seed, uint64_t matrix_id, double *dst, int64_t count) {
  return ec_fill_uniform_batched(p, seed, matrix_id, dst, count, 1);
}

}  // extern "C"
