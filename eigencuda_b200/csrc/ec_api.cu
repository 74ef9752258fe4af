// eigencuda_b200/csrc/ec_api.cu -- implementation of the C ABI declared in include/eigencuda_b200.h.
//
// Host side of the hot path: the memory subsystem that replaces CudaMatrix
// (reference src/cudamatrix.cc:20-79), the op dispatch that replaces CudaPipeline::gemm
// (reference src/cudapipeline.cc:17-32) and the overlapped tensor-stream engine that replaces the
// caller-side loop of README.md:62-66.  No cuBLAS, no CPU compute fallback: every product is a
// launch of a kernel from dgemm_kernels.cuh.
#include <cuda.h>
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>

#include <algorithm>
#include <atomic>
#include <chrono>
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


// The host side is one translation unit, split by concern:
#include "host_util.inl"        // errors, driver entry point, staging ring, copy pool
#include "handles.inl"          // ec_pipeline / ec_matrix, stream registry, device guard
#include "dgemm_launch.inl"     // native path: tensor maps, stream-K workspace, launch, cost model
#include "ozaki_dispatch.inl"   // int8 emulation backend: split + GEMM launches
#include "dgemm_dispatch.inl"   // validation, corner cases, backend / kernel choice
#include "stream_engine.inl"    // overlapped tensor-stream engine
#include "stream_run.inl"       // one tensor-stream job on one device (feeder + drainer)
#include "multi_gpu.inl"        // the same job sharded over several devices of one box (NCCL broadcast)
#include "dgemm_2d.inl"         // one huge DGEMM, 2D block-partitioned over the devices (SUMMA, NCCL panel broadcasts)

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
  if (const char *e = std::getenv("EC_DYNAMIC_TILES")) p->dynamic_tiles = std::atoi(e) != 0;
  if (const char *e = std::getenv("EC_FUSE_EPILOGUE")) p->fuse_epilogue = std::atoi(e) != 0;
  if (const char *e = std::getenv("EC_FUSE_MIN_KBLOCKS")) p->fuse_min_kblocks = std::max(1, std::atoi(e));
  if (const char *e = std::getenv("EC_PDL")) p->pdl = std::max(0, std::min(3, std::atoi(e)));
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
  // Device buffers of CudaMatrix come from the device's stream-ordered memory pool: constructing
  // and destroying matrices inside a loop (the reference pays cudaMemGetInfo + cudaMalloc and a
  // device-synchronising cudaFree each time, src/cudamatrix.cc:53-79) then costs microseconds and
  // does not stall the stream.  Freed blocks stay cached up to 1 GiB.  EC_MEM_POOL=0 turns it off.
  {
    int supported = 0;
    const char *env = std::getenv("EC_MEM_POOL");
    if (!(env && env[0] == '0') &&
        cudaDeviceGetAttribute(&supported, cudaDevAttrMemoryPoolsSupported, dev) == cudaSuccess && supported) {
      cudaMemPool_t pool;
      if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        uint64_t have = 0;
        const uint64_t want = (uint64_t)1 << 30;
        if (cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &have) == cudaSuccess && have < want) {
          uint64_t v = want;
          cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &v);
        }
        p->mem_pool = true;
      }
    }
    cudaGetLastError();
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
  if (p->multi) {
    p->multi->destroy();
    delete p->multi;
    p->multi = nullptr;
  }
  if (p->engine) {
    p->engine->destroy();
    delete p->engine;
  }
  p->staging.destroy();
  delete p->compat_pool;
  p->compat_pool = nullptr;
  if (p->ev_order) cudaEventDestroy(p->ev_order);
  if (p->sched_words) cudaFree(p->sched_words);
  if (p->skew_host) cudaFreeHost(p->skew_host);
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
int ec_plan_dgemm(int sm_count, int64_t m, int64_t n, int64_t k, int64_t batch, int *tile,
                  int *stream_k) {
  if (!tile || !stream_k) return fail(EC_ERR_ARG, "null out pointer");
  if (sm_count <= 0 || m <= 0 || n <= 0 || k <= 0 || batch <= 0) return fail(EC_ERR_ARG, "bad planning query");
  ec_pipeline model;   // host-side description only: no stream, nothing allocated
  model.sm_count = sm_count;
  const TileChoice ch = choose_tile(&model, m, n, k, batch);
  *tile = ch.tile;
  *stream_k = ch.sk ? 1 : 0;
  return EC_OK;
}
int64_t ec_pipeline_launch_count(const ec_pipeline *p) { return p ? p->launches : 0; }
int64_t ec_pipeline_contention_events(const ec_pipeline *p) { return p ? p->contended_events : 0; }
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
  const bool pooled = p && p->mem_pool;
  if (bytes > 0) {
    cudaError_t e = pooled ? cudaMallocAsync((void **)&d, bytes, st) : cudaMalloc((void **)&d, bytes);
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
  m->pooled = pooled && d != nullptr;
  *out = m;
  return EC_OK;
}

int ec_matrix_free(ec_matrix *m) {
  if (!m) return EC_OK;
  if (m->data) {
    DeviceGuard guard(m->device);
    // Pooled buffers are released in stream order, behind the work that still uses them.  If the
    // matrix has outlived its pipeline (the reference tolerates that: its deleter is a plain
    // cudaFree, include/cudamatrix.hpp:46) the stream may be gone, and cudaFree -- legal for pool
    // memory, and synchronising -- is used instead.
    ec_pipeline *p = m->pooled ? pipeline_of_stream(m->stream) : nullptr;
    if (!(p && p->device == m->device && cudaFreeAsync(m->data, m->stream) == cudaSuccess)) {
      cudaGetLastError();
      cudaFree(m->data);  // synchronises with outstanding work on the buffer
    }
  }
  delete m;
  return EC_OK;
}

namespace {
// Pageable <-> pinned copy of one whole matrix for the reference's blocking calls
// (src/cudamatrix.cc:8-12 and :47-51).  A single core moves about 12 GB/s, a quarter of what the
// host link then takes it at, so anything of 1 MiB or more is cut into pieces for a few workers
// plus the calling thread.
void compat_copy(ec_pipeline *p, void *dst, const void *src, size_t bytes, bool prefault_only = false) {
  constexpr size_t MIN_PARALLEL = (size_t)1 << 20, MIN_PIECE = (size_t)64 << 10;
  auto piece_job = [&](void *d, const void *s, size_t n) {
    if (prefault_only)
      ec_host::prefault_for_write(d, n);
    else
      staging_copy(d, s, n);
  };
  if (bytes < MIN_PARALLEL) {
    piece_job(dst, src, bytes);
    return;
  }
  if (!p->compat_pool_tried) {
    p->compat_pool_tried = true;
    int workers = std::max(1, std::min(5, (int)std::thread::hardware_concurrency() / 3));
    if (const char *e = std::getenv("EC_COMPAT_THREADS")) workers = std::max(0, std::atoi(e) - 1);
    // 1 ms of polling covers a whole iteration of the reference's loop at 512^2 .. 1000^2
    if (workers > 0) p->compat_pool = new CopyPool(workers, 1000);
  }
  if (!p->compat_pool) {
    piece_job(dst, src, bytes);
    return;
  }
  // four pieces per thread, handed out dynamically: a worker that shares its core with somebody
  // else (SMT sibling, a busy neighbour VM) then simply takes fewer of them
  const size_t lanes = (size_t)p->compat_pool->size() + 1;
  size_t piece = std::max(MIN_PIECE, (bytes + 4 * lanes - 1) / (4 * lanes));
  piece = (piece + 4095) & ~(size_t)4095;
  const int64_t pieces = (int64_t)((bytes + piece - 1) / piece);
  p->compat_pool->parallel_for(pieces, [&](int64_t i) {
    const size_t lo = (size_t)i * piece, hi = std::min(bytes, lo + piece);
    piece_job((char *)dst + lo, (const char *)src + lo, hi - lo);
  });
}
}  // namespace

int ec_matrix_upload(ec_matrix *m, const double *host, int64_t count) {
  if (!m) return fail(EC_ERR_ARG, "null matrix");
  if (count < 0 || count > m->rows * m->cols)
    return fail(EC_ERR_ARG, "copy_to_gpu: source has %lld elements, device matrix holds %lld",
                (long long)count, (long long)(m->rows * m->cols));
  if (count == 0) return EC_OK;
  if (!host) return fail(EC_ERR_ARG, "null host pointer");
  DeviceGuard guard(m->device);
  TraceRange trace("ec_matrix_upload");
  const size_t bytes = (size_t)count * sizeof(double);
  ec_pipeline *p = pipeline_of_stream(m->stream);
  if (p && !is_pinned_host(host)) {
    // Whole-matrix staging up to 8 MiB (one DMA: the bench shapes); anything larger goes in 4 MiB
    // pieces over the rotating slots, so that the host copy of piece i+1 runs under the DMA of
    // piece i and a 2 GiB matrix pins 16 MiB of staging, not 2 GiB per slot.
    const size_t PIECE = bytes <= ((size_t)8 << 20) ? bytes : ((size_t)4 << 20);
    for (size_t off = 0; off < bytes; off += PIECE) {
      const size_t len = std::min(PIECE, bytes - off);
      int slot = -1;
      int rc = p->staging.acquire(len, &slot);
      if (rc) return rc;
      compat_copy(p, p->staging.buf[slot], (const char *)host + off, len);
      cudaError_t e = cudaMemcpyAsync((char *)m->data + off, p->staging.buf[slot], len,
                                      cudaMemcpyHostToDevice, m->stream);
      if (e != cudaSuccess) return fail(EC_ERR_COPY, "Error copy arrays to device: %s", cudaGetErrorString(e));
      rc = p->staging.release_after(slot, m->stream);
      if (rc) return rc;
    }
    return EC_OK;
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
  TraceRange trace("ec_matrix_download");
  const size_t bytes = (size_t)count * sizeof(double);
  ec_pipeline *p = pipeline_of_stream(m->stream);
  // Pageable destination of 1 MiB or more: DMA into two pinned slots alternately and copy out
  // behind the DMA, in 4 MiB pieces (two halves for anything smaller than 8 MiB).
  const size_t CHUNK = std::min((size_t)4 << 20, (bytes / 2 + 4095) & ~(size_t)4095);
  if (p && bytes >= ((size_t)1 << 20) && !is_pinned_host(host)) {
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
        if (off == 0) compat_copy(p, host, host, bytes, /*prefault_only=*/true);   // behind the first DMA
      }
      if (prev_slot >= 0) {
        EC_CUDA(cudaEventSynchronize(p->staging.ev[prev_slot]));
        p->staging.busy[prev_slot] = false;
        compat_copy(p, (char *)host + prev_off, p->staging.buf[prev_slot], prev_len);
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
  DeviceGuard guard(p->device);
  // Operands created on another stream than the pipeline's: the reference is ordered by the
  // legacy NULL stream in that case; here the dependency is made explicit.  One event per
  // pipeline, created on its device and kept: nothing to leak on an error path.
  const ec_matrix *ops[3] = {A, B, C};
  bool foreign = false;
  for (const ec_matrix *m : ops) foreign = foreign || m->stream != p->stream;
  if (foreign) {
    for (const ec_matrix *m : ops)
      if (m->device != p->device)
        return fail(EC_ERR_ARG, "gemm: operand lives on device %d, the pipeline on device %d", m->device, p->device);
    if (!p->ev_order) EC_CUDA(cudaEventCreateWithFlags(&p->ev_order, cudaEventDisableTiming));
    for (const ec_matrix *m : ops) {
      if (m->stream == p->stream) continue;
      EC_CUDA(cudaEventRecord(p->ev_order, m->stream));
      EC_CUDA(cudaStreamWaitEvent(p->stream, p->ev_order, 0));
    }
  }
  int rc = dgemm_dispatch(p, EC_OP_N, EC_OP_N, A->rows, B->cols, A->cols, 1.0, A->data, A->rows, 0,
                          B->data, B->rows, 0, 0.0, C->data, std::max<int64_t>(C->rows, 1), 0, 1);
  if (rc) return rc;
  // ... and in the other direction: whatever the operands' own streams do next -- read the result,
  // overwrite an input, release a pooled buffer in stream order -- comes after this product.
  if (foreign) {
    EC_CUDA(cudaEventRecord(p->ev_order, p->stream));
    for (const ec_matrix *m : ops)
      if (m->stream != p->stream) EC_CUDA(cudaStreamWaitEvent(m->stream, p->ev_order, 0));
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
  return tensor_stream_run_impl(p, kind, F, f_rows, f_cols, tensor, results, count, t_rows, t_cols);
}

int ec_tensor_stream_run_multi(ec_pipeline *p, const int *devices, int ndev, int kind, const double *F,
                               int64_t f_rows, int64_t f_cols, const double *const *tensor,
                               double *const *results, int64_t count, int64_t t_rows, int64_t t_cols) {
  return tensor_stream_run_multi_impl(p, devices, ndev, kind, F, f_rows, f_cols, tensor, results, count,
                                      t_rows, t_cols);
}

// ---- 2D block-partitioned DGEMM over a device grid -------------------------------------------

int ec_grid_create(const int *devices, int ndev, int P, int Q, ec_grid **out) {
  return grid_create_impl(devices, ndev, P, Q, out);
}
int ec_grid_destroy(ec_grid *g) { return grid_destroy_impl(g); }
int ec_grid_shape(const ec_grid *g, int *P, int *Q) {
  if (!g || !P || !Q) return fail(EC_ERR_ARG, "null pointer");
  *P = g->P;
  *Q = g->Q;
  return EC_OK;
}
int ec_grid_set_backend(ec_grid *g, int backend) {
  if (!g) return fail(EC_ERR_ARG, "null grid");
  for (ec_pipeline *c : g->pipes) {
    int rc = ec_pipeline_set_backend(c, backend);
    if (rc) return rc;
  }
  return EC_OK;
}
int ec_grid_set_schedule(ec_grid *g, int stream_k, int one_cta_per_tile) {
  if (!g) return fail(EC_ERR_ARG, "null grid");
  for (ec_pipeline *c : g->pipes) ec_pipeline_set_schedule(c, stream_k, one_cta_per_tile);
  return EC_OK;
}
int64_t ec_grid_launch_count(const ec_grid *g) {
  int64_t total = 0;
  if (g)
    for (const ec_pipeline *c : g->pipes) total += c->launches;
  return total;
}
int64_t ec_grid_broadcast_count(const ec_grid *g) { return g ? g->broadcasts : 0; }
const char *ec_grid_last_kernel(const ec_grid *g) { return g && !g->pipes.empty() ? g->pipes[0]->last_kernel : "none"; }

int ec_dist_matrix_create(ec_grid *g, int64_t rows, int64_t cols, ec_dist_matrix **out) {
  return dist_matrix_create_impl(g, rows, cols, out);
}
int ec_dist_matrix_destroy(ec_dist_matrix *m) { return dist_matrix_destroy_impl(m); }
int ec_dist_matrix_upload(ec_dist_matrix *m, const double *host, int64_t ld) {
  return dist_matrix_transfer(m, const_cast<double *>(host), ld, true);
}
int ec_dist_matrix_download(const ec_dist_matrix *m, double *host, int64_t ld) {
  return dist_matrix_transfer(m, host, ld, false);
}
int ec_dist_matrix_fill_uniform(ec_dist_matrix *m, uint64_t seed, uint64_t matrix_id) {
  if (!m) return fail(EC_ERR_ARG, "null matrix");
  const ec_grid *g = m->grid;
  for (int pr = 0; pr < g->P; ++pr)
    for (int pc = 0; pc < g->Q; ++pc) {
      const int i = g->dev(pr, pc);
      const BlockGeom b = block_of(g, m->rows, m->cols, pr, pc);
      int rc = ec_fill_uniform_block(g->pipes[i], seed, matrix_id, m->block[i], b.nrows(), b.ncols(), b.nrows(),
                                     b.r0, b.c0, m->rows);
      if (rc) return rc;
    }
  return EC_OK;
}
int ec_dist_matrix_get(const ec_dist_matrix *m, int64_t i, int64_t j, double *out) {
  if (!m || !out) return fail(EC_ERR_ARG, "null pointer");
  if (i < 0 || j < 0 || i >= m->rows || j >= m->cols) return fail(EC_ERR_ARG, "element (%lld, %lld) out of range", (long long)i, (long long)j);
  const ec_grid *g = m->grid;
  int64_t r0 = 0, c0 = 0;
  const int pr = owner_of(m->rows, g->P, i, &r0), pc = owner_of(m->cols, g->Q, j, &c0);
  const BlockGeom b = block_of(g, m->rows, m->cols, pr, pc);
  const int d = g->dev(pr, pc);
  DeviceGuard guard(g->devices[d]);
  EC_CUDA(cudaMemcpyAsync(out, m->block[d] + (j - c0) * b.nrows() + (i - r0), 8, cudaMemcpyDeviceToHost, g->pipes[d]->stream));
  EC_CUDA(cudaStreamSynchronize(g->pipes[d]->stream));
  return EC_OK;
}
int ec_dgemm_2d(ec_grid *g, const ec_dist_matrix *A, const ec_dist_matrix *B, ec_dist_matrix *C, int64_t panel,
                double *elapsed_ms) {
  return dgemm_2d_impl(g, A, B, C, panel, elapsed_ms);
}
int ec_dgemm_2d_host(const int *devices, int ndev, int64_t m, int64_t n, int64_t k, const double *A, int64_t lda,
                     const double *B, int64_t ldb, double *C, int64_t ldc) {
  if (m < 0 || n < 0 || k < 0) return fail(EC_ERR_ARG, "negative dimension");
  ec_grid *g = nullptr;
  int rc = ec_grid_create(devices, ndev, 0, 0, &g);
  if (rc) return rc;
  ec_dist_matrix *dA = nullptr, *dB = nullptr, *dC = nullptr;
  rc = ec_dist_matrix_create(g, m, k, &dA);
  if (!rc) rc = ec_dist_matrix_create(g, k, n, &dB);
  if (!rc) rc = ec_dist_matrix_create(g, m, n, &dC);
  if (!rc) rc = ec_dist_matrix_upload(dA, A, lda);
  if (!rc) rc = ec_dist_matrix_upload(dB, B, ldb);
  if (!rc) rc = ec_dgemm_2d(g, dA, dB, dC, 0, nullptr);
  if (!rc) rc = ec_dist_matrix_download(dC, C, ldc);
  std::string keep = g_last_error;
  ec_dist_matrix_destroy(dA);
  ec_dist_matrix_destroy(dB);
  ec_dist_matrix_destroy(dC);
  ec_grid_destroy(g);
  g_last_error = keep;
  return rc;
}

int64_t ec_plan_chunks(int64_t count, int64_t chunk, int ramp, int64_t *starts, int64_t cap) {
  if (count < 0 || chunk <= 0 || cap < 0 || (cap > 0 && !starts)) {
    fail(EC_ERR_ARG, "bad chunk planning query");
    return -1;
  }
  const std::vector<int64_t> plan = plan_chunks(count, chunk, ramp != 0);
  for (size_t i = 0; i < plan.size() && (int64_t)i < cap; ++i) starts[i] = plan[i];
  return (int64_t)plan.size();
}

int64_t ec_plan_panels(int64_t k, int P, int Q, int64_t panel, int64_t *starts, int64_t cap) {
  if (k < 0 || P <= 0 || Q <= 0 || panel < 0 || cap < 0 || (cap > 0 && !starts)) {
    fail(EC_ERR_ARG, "bad panel planning query");
    return -1;
  }
  const std::vector<std::pair<int64_t, int64_t>> panels = summa_panels(k, P, Q, panel > 0 ? panel : 4096);
  int64_t n = 0;
  if (n < cap) starts[n] = 0;
  ++n;
  for (const auto &pn : panels) {
    if (n < cap) starts[n] = pn.second;
    ++n;
  }
  return n;
}

int ec_plan_shard(int64_t count, int ndev, int idx, int64_t *lo, int64_t *hi) {
  if (!lo || !hi) return fail(EC_ERR_ARG, "null out pointer");
  if (count < 0 || ndev <= 0 || idx < 0 || idx >= ndev) return fail(EC_ERR_ARG, "bad shard query");
  shard_range(count, ndev, idx, lo, hi);
  return EC_OK;
}

int ec_nccl_version(int *version) {
  if (!version) return fail(EC_ERR_ARG, "null out pointer");
  *version = 0;
  NcclApi *api = nccl_api();
  if (!api->handle) return fail(EC_ERR_UNSUPPORTED, "NCCL is not available: %s", api->why.c_str());
  EC_NCCL(api, api->GetVersion(version));
  return EC_OK;
}

int ec_pick_stream_devices(const int *devices, int ndev, int *out_devices, int *out_ndev, double *out_gbs) {
  if (!out_devices || !out_ndev) return fail(EC_ERR_ARG, "null out pointer");
  std::vector<int> devs;
  int rc = device_list_or_all(devices, ndev, &devs);
  if (rc) return rc;
  DevicePick pick;
  if ((rc = pick_stream_devices_impl(devs, &pick))) return rc;
  for (size_t i = 0; i < pick.picked.size(); ++i) out_devices[i] = pick.picked[i];
  *out_ndev = (int)pick.picked.size();
  if (out_gbs) {
    out_gbs[0] = pick.gbs_picked;
    out_gbs[1] = pick.gbs_all;
  }
  return EC_OK;
}

int ec_host_link_probe(const int *devices, int ndev, size_t bytes, int reps, double *out_gbs) {
  return host_link_probe_impl(devices, ndev, bytes, reps, out_gbs);
}
int ec_host_link_probe_buffers(const int *devices, int ndev, const void *host_in, void *host_out,
                               size_t bytes_per_device, size_t piece_bytes, int reps, double *out_gbs) {
  if (!out_gbs) return fail(EC_ERR_ARG, "null out pointer");
  out_gbs[0] = out_gbs[1] = out_gbs[2] = 0.0;
  if (!host_in || !host_out || bytes_per_device == 0 || piece_bytes == 0 || reps <= 0)
    return fail(EC_ERR_ARG, "bad probe arguments");
  std::vector<int> devs;
  int rc = device_list_or_all(devices, ndev, &devs);
  if (rc) return rc;
  if (!is_pinned_host(host_in) || !is_pinned_host(host_out)) return fail(EC_ERR_ARG, "the probe needs pinned host buffers");
  return host_link_probe_user_impl(devs, (const char *)host_in, (char *)host_out, bytes_per_device, piece_bytes, reps, out_gbs);
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

int ec_host_register(void *ptr, size_t bytes) {
  if (bytes == 0) return EC_OK;
  if (!ptr) return fail(EC_ERR_ARG, "null host pointer");
  cudaError_t e = cudaHostRegister(ptr, bytes, cudaHostRegisterPortable);
  if (e == cudaErrorHostMemoryAlreadyRegistered) {
    // Either the same range again (fine), or a small matrix that shares its first or last page
    // with a pinned neighbour -- then the middle of it is NOT pinned and must not be taken for
    // DMA-able.  Look at the pages (all of them up to 256, evenly spaced samples beyond).
    cudaGetLastError();
    const uintptr_t lo = (uintptr_t)ptr & ~(uintptr_t)4095, hi = (uintptr_t)ptr + bytes - 1;
    const size_t pages = (size_t)((hi - lo) / 4096) + 1, step = std::max<size_t>(1, pages / 256);
    for (size_t pg = 0; pg < pages; pg += step)
      if (!is_pinned_host((const void *)std::max(lo + pg * 4096, (uintptr_t)ptr)))
        return fail(EC_ERR_ARG, "range overlaps a registered range without being covered by it");
    if (!is_pinned_host((const void *)hi))
      return fail(EC_ERR_ARG, "range overlaps a registered range without being covered by it");
    return EC_OK;
  }
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(EC_ERR_NOMEM, "cudaHostRegister of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
  }
  return EC_OK;
}
int ec_host_unregister(void *ptr) {
  if (!ptr) return EC_OK;
  cudaError_t e = cudaHostUnregister(ptr);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(EC_ERR_ARG, "cudaHostUnregister failed: %s", cudaGetErrorString(e));
  }
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
uint32_t ec_debug_fastdiv(uint32_t n, uint32_t d) { return d == 0 ? 0u : ec::FastDiv::make(d).div(n); }   // d = 0: outside the contract, but no SIGFPE

int ec_debug_hold_sms(int device, void *cuda_stream, int sms, int64_t microseconds) {
  if (sms <= 0 || microseconds < 0) return fail(EC_ERR_ARG, "bad hold request");
  DeviceGuard guard(device);
  constexpr int SMEM = 200 * 1024;
  EC_CUDA(cudaFuncSetAttribute(ec::hold_sms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
  ec::hold_sms_kernel<<<sms, 32, SMEM, (cudaStream_t)cuda_stream>>>((unsigned long long)microseconds * 1000ull);
  EC_CUDA(cudaGetLastError());
  return EC_OK;
}

int ec_fill_uniform(ec_pipeline *p, uint64_t seed, uint64_t matrix_id, double *dst, int64_t count) {
  return ec_fill_uniform_batched(p, seed, matrix_id, dst, count, 1);
}

}  // extern "C"
