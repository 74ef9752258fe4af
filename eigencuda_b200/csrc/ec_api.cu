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
  if (p->engine) {
    p->engine->destroy();
    delete p->engine;
  }
  p->staging.destroy();
  delete p->compat_pool;
  p->compat_pool = nullptr;
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
    int slot = -1;
    int rc = p->staging.acquire(bytes, &slot);
    if (rc) return rc;
    compat_copy(p, p->staging.buf[slot], host, bytes);
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
  // ... and in the other direction: whatever the operands' own streams do next -- read the result,
  // overwrite an input, release a pooled buffer in stream order -- comes after this product.
  cudaEvent_t done = nullptr;
  for (const ec_matrix *m : ops) {
    if (m->stream == p->stream) continue;
    if (!done) {
      EC_CUDA(cudaEventCreateWithFlags(&done, cudaEventDisableTiming));
      EC_CUDA(cudaEventRecord(done, p->stream));
    }
    EC_CUDA(cudaStreamWaitEvent(m->stream, done, 0));
  }
  if (done) EC_CUDA(cudaEventDestroy(done));
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
  TraceRange trace("ec_tensor_stream_run");

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
          TraceRange trace_chunk("deliver chunk (pinned -> caller)");
          const char *src = (const char *)E.pin_out[s];
          // split every matrix into a few pieces so the pool stays busy even for small chunks
          const int pieces = std::max<int>(1, (int)std::min<int64_t>(8, (r_sz * 8) >> 18));
          E.pool_out->parallel_for(nb * pieces, [&](int64_t t) {
            const int64_t i = t / pieces, q = t % pieces;
            const size_t lo = (size_t)(r_sz * 8) * q / pieces, hi = (size_t)(r_sz * 8) * (q + 1) / pieces;
            // result matrices are often freshly allocated: map the piece in one call, not per page
            ec_host::prefault_for_write((char *)results[i0 + i] + lo, hi - lo);
            staging_copy((char *)results[i0 + i] + lo, src + (size_t)i * r_sz * 8 + lo, hi - lo);
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
        TraceRange trace_chunk("stage chunk (caller -> pinned)");
        char *dst = (char *)E.pin_in[s];
        const int pieces = std::max<int>(1, (int)std::min<int64_t>(8, (t_sz * 8) >> 18));
        E.pool_in->parallel_for(nb * pieces, [&](int64_t t) {
          const int64_t i = t / pieces, q = t % pieces;
          const size_t lo = (size_t)(t_sz * 8) * q / pieces, hi = (size_t)(t_sz * 8) * (q + 1) / pieces;
          staging_copy(dst + (size_t)i * t_sz * 8 + lo, (const char *)tensor[i0 + i] + lo, hi - lo);
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
        StreamScope scope(p, E.s_cmp);
        rc = stream_compute(p, kind, dF, f_rows, f_cols, E.d_in[s], E.d_out[s],
                            kind == EC_STREAM_BASIS ? E.d_tmp[s] : nullptr, nb, t_rows, t_cols, sh);
      }
      if (rc) return rc;
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
int ec_fill_uniform(ec_pipeline *p, uint64_t seed, uint64_t matrix_id, double *dst, int64_t count) {
  return ec_fill_uniform_batched(p, seed, matrix_id, dst, count, 1);
}

}  // extern "C"
