/*
 * eigencuda_b200.h -- the C ABI of the B200-native replacement for EigenCuda's hot path.
 *
 * This is the drop-in boundary: plain pointers, sizes and opaque handles; no C++ types, no torch
 * types, no exceptions.  include/cudamatrix.hpp and include/cudapipeline.hpp (source-compatible
 * with the reference's headers of the same names) are thin inline wrappers over these functions,
 * and eigencuda_b200/ (Python) binds the same symbols through ctypes.
 *
 * Every entry point cites the reference interface (path:line under NLESC-JCER/EigenCuda) it
 * replaces.  All matrices are FP64, column-major, contiguous unless a leading dimension is given.
 * Status codes: 0 = ok; nonzero = error, with a message available from ec_last_error_string()
 * (thread-local).  Handles are thread-compatible (one host thread at a time per pipeline), as in
 * the reference, which has no locks either.
 *
 * There is no CPU fallback behind this ABI: every compute entry point launches hand-written
 * sm_100a kernels (eigencuda_b200/csrc/) and fails with EC_ERR_CUDA when no B200 is present.
 */
#ifndef EIGENCUDA_B200_H
#define EIGENCUDA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ec_pipeline ec_pipeline; /* replaces class CudaPipeline, include/cudapipeline.hpp:20-42 */
typedef struct ec_matrix ec_matrix;     /* replaces class CudaMatrix,   include/cudamatrix.hpp:27-60   */

enum {
  EC_OK = 0,
  EC_ERR_SHAPE = 1,   /* "Shape mismatch in Cublas gemm", src/cudapipeline.cc:26-28          */
  EC_ERR_NOMEM = 2,   /* throw_if_not_enough_memory_in_gpu, src/cudamatrix.cc:63-79          */
  EC_ERR_COPY = 3,    /* "Error copy arrays to device", src/cudamatrix.cc:28-30              */
  EC_ERR_CUDA = 4,    /* any other CUDA failure (the reference swallows these, :5-12)        */
  EC_ERR_ARG = 5,     /* null handle / negative size / bad enum                              */
  EC_ERR_UNSUPPORTED = 6
};

enum { EC_OP_N = 0, EC_OP_T = 1 };

/* Arithmetic backend for the GEMM kernels.  EC_BACKEND_FP64_DMMA is the native path (FP64
 * mma.sync fragments, parity bar 1e-12 relative Frobenius).  EC_BACKEND_OZAKI_I8 is the selectable
 * int8 tcgen05 emulation with a looser, explicitly stated bound. */
enum { EC_BACKEND_FP64_DMMA = 0, EC_BACKEND_OZAKI_I8 = 1 };

/* Tensor-stream job kinds (ec_tensor_stream_run): what each tensor element T_i is multiplied by.
 * RIGHT is the order the reference itself uses (gemm(cuma_B, cuma_A, cuma_C) with cuma_B the
 * streamed element, README.md:62-66 / src/tests/test_dot.cc:81-85). */
enum {
  EC_STREAM_RIGHT = 0,   /* R_i = T_i * F                                                 */
  EC_STREAM_LEFT = 1,    /* R_i = F * T_i                                                 */
  EC_STREAM_LEFT_T = 2,  /* R_i = F^T * T_i            (BASELINE.json config 4, A^T*B)    */
  EC_STREAM_BASIS = 3    /* R_i = F^T * T_i * F        (BASELINE.json config 4, A^T*M*A)  */
};

/* ---- library / device --------------------------------------------------------------------- */

const char *ec_version(void);
const char *ec_last_error_string(void);

/* count_available_gpus(), include/cudamatrix.hpp:25, src/cudamatrix.cc:14-18.  Returns the device
 * count, or 0 on error (the reference's ternary is inverted and returns garbage on error). */
int64_t ec_device_count(void);

/* ---- pipeline: CudaPipeline, include/cudapipeline.hpp:20-42 ------------------------------- */

/* CudaPipeline::CudaPipeline(), include/cudapipeline.hpp:22-25: one stream per pipeline (the
 * reference also creates a cuBLAS handle; there is none here).  device < 0 = current device. */
int ec_pipeline_create(int device, ec_pipeline **out);
/* Same, but runs on a caller-owned stream (e.g. a framework's current stream); not destroyed. */
int ec_pipeline_create_on_stream(int device, void *cuda_stream, ec_pipeline **out);
/* CudaPipeline::~CudaPipeline(), src/cudapipeline.cc:5-11. */
int ec_pipeline_destroy(ec_pipeline *p);
/* CudaPipeline::get_stream(), include/cudapipeline.hpp:34.  Returns the cudaStream_t. */
void *ec_pipeline_stream(const ec_pipeline *p);
/* Address of the stored cudaStream_t (get_stream() returns a const reference). */
const void *ec_pipeline_stream_ref(const ec_pipeline *p);
int ec_pipeline_device(const ec_pipeline *p);
int ec_pipeline_synchronize(ec_pipeline *p);
int ec_pipeline_set_backend(ec_pipeline *p, int backend);
int ec_pipeline_get_backend(const ec_pipeline *p);

/* ---- device matrix: CudaMatrix, include/cudamatrix.hpp:27-60 ------------------------------ */

/* CudaMatrix(Index nrows, Index ncols, const cudaStream_t&), src/cudamatrix.cc:33-37:
 * allocate rows*cols doubles on the stream's device, contents uninitialised.
 * EC_ERR_NOMEM when the request exceeds free device memory (src/cudamatrix.cc:63-79). */
int ec_matrix_alloc(void *cuda_stream, int64_t rows, int64_t cols, ec_matrix **out);
/* CudaMatrix(const Eigen::MatrixXd&, const cudaStream_t&), src/cudamatrix.cc:20-31:
 * allocate + asynchronous H2D of rows*cols doubles.  EC_ERR_COPY if the copy cannot be issued.
 * `host` may be pageable; it is reusable as soon as the call returns. */
int ec_matrix_create_from_host(void *cuda_stream, const double *host, int64_t rows, int64_t cols,
                               ec_matrix **out);
/* ~CudaMatrix (the unique_ptr deleter, include/cudamatrix.hpp:46,55-56). */
int ec_matrix_free(ec_matrix *m);
/* CudaMatrix::copy_to_gpu(const Eigen::MatrixXd&), src/cudamatrix.cc:47-51: asynchronous H2D of
 * `count` doubles into the existing buffer, on the matrix's stream; `host` reusable on return.
 * Unlike the reference (Appendix A #5) count > rows*cols is rejected with EC_ERR_ARG. */
int ec_matrix_upload(ec_matrix *m, const double *host, int64_t count);
/* CudaMatrix::operator Eigen::MatrixXd() const, src/cudamatrix.cc:39-45: D2H of the whole
 * matrix; the result is complete in `host` when the call returns. */
int ec_matrix_download(const ec_matrix *m, double *host);
/* rows()/cols()/size()/data(), include/cudamatrix.hpp:29-32. */
int64_t ec_matrix_rows(const ec_matrix *m);
int64_t ec_matrix_cols(const ec_matrix *m);
int64_t ec_matrix_size(const ec_matrix *m);
double *ec_matrix_data(const ec_matrix *m);

/* ---- the hot op: CudaPipeline::gemm, src/cudapipeline.cc:17-32 ---------------------------- */

/* gemm(A, B, C): C = A * B.  EC_ERR_SHAPE when A.cols != B.rows, before any device work
 * (src/cudapipeline.cc:26-28).  C's shape is not validated, as in the reference: its rows() is
 * the leading dimension of the output (src/cudapipeline.cc:31). */
int ec_gemm(ec_pipeline *p, const ec_matrix *A, const ec_matrix *B, ec_matrix *C);

/* The cublasDgemm call itself, src/cudapipeline.cc:29-31, with the parameters the reference
 * hard-codes opened up: C = alpha * op(A) * op(B) + beta * C on device pointers, 64-bit
 * dimensions (the reference narrows to int, Appendix A #4).  op(A) is m x k, op(B) is k x n. */
int ec_dgemm(ec_pipeline *p, int transa, int transb, int64_t m, int64_t n, int64_t k, double alpha,
             const double *A, int64_t lda, const double *B, int64_t ldb, double beta, double *C,
             int64_t ldc);

/* The same product for `batch` independent problems at fixed element strides (stride 0 =
 * operand shared by all problems): one launch for a chunk of the matrix tensor against the
 * fixed operand.  Resurrects the gemmBatched path CHANGELOG.md:24 records as removed. */
int ec_dgemm_strided_batched(ec_pipeline *p, int transa, int transb, int64_t m, int64_t n,
                             int64_t k, double alpha, const double *A, int64_t lda,
                             int64_t stride_a, const double *B, int64_t ldb, int64_t stride_b,
                             double beta, double *C, int64_t ldc, int64_t stride_c, int64_t batch);

/* ---- the streaming loop as one overlapped job --------------------------------------------- */

/* README.md:58-66 / src/tests/test_dot.cc:77-85, the whole `for` loop:
 *     for i: cuma_B.copy_to_gpu(tensor[i]); cuda_pip.gemm(...); results[i] = cuma_C;
 * F is f_rows x f_cols, every tensor[i] is t_rows x t_cols, results[i] receives the product
 * (shape by `kind`), all host pointers, column-major, ld = rows.  H2D, compute and D2H run on
 * separate streams over a ring of device chunks; pinned/registered host buffers are DMA'd
 * directly, pageable ones go through a pinned staging ring filled by worker threads.
 * All results are complete in host memory when the call returns. */
int ec_tensor_stream_run(ec_pipeline *p, int kind, const double *F, int64_t f_rows, int64_t f_cols,
                         const double *const *tensor, double *const *results, int64_t count,
                         int64_t t_rows, int64_t t_cols);
/* Same job on a tensor that is already resident in device memory as `count` consecutive
 * matrices (element stride t_rows*t_cols); results written to `results_dev` likewise. */
int ec_tensor_stream_run_device(ec_pipeline *p, int kind, const double *F_dev, int64_t f_rows,
                                int64_t f_cols, const double *tensor_dev, double *results_dev,
                                int64_t count, int64_t t_rows, int64_t t_cols);
/* The same job sharded over several GPUs of one box -- the multi-GPU scheduler.  `tensor` and
 * `results` are the caller's ONE vector (README.md:58-66); device devices[g] takes the contiguous
 * index range [g*count/ndev, (g+1)*count/ndev) with its own worker thread, pinned ring and three
 * streams; F goes to devices[0] once and is broadcast to the others with NCCL over NVLink
 * (libnccl.so.2 is bound at run time: EC_ERR_UNSUPPORTED when ndev > 1 and it cannot be loaded);
 * every device DMAs its results straight into results[i]; no result exchange.  ndev == 0 means
 * every device of the box (devices may then be NULL), ndev == EC_DEVICES_AUTO the subset of them
 * the host streams through fastest (ec_pick_stream_devices).  `p` carries the engine knobs and the
 * backend and owns the per-device state, which is kept across calls.  The reference has no
 * multi-GPU path; this replaces its unused count_available_gpus(), src/cudamatrix.cc:14-18. */
int ec_tensor_stream_run_multi(ec_pipeline *p, const int *devices, int ndev, int kind, const double *F,
                               int64_t f_rows, int64_t f_cols, const double *const *tensor,
                               double *const *results, int64_t count, int64_t t_rows, int64_t t_cols);
/* Whether the multi-GPU entry points can run here; needs no device.  Binds libnccl.so.2 (or the
 * library the environment variable EC_NCCL_LIB names -- then only that one is tried) the way
 * ec_tensor_stream_run_multi and ec_dgemm_2d do and stores ncclGetVersion()'s code
 * (major*10000 + minor*100 + patch) in *version.  EC_ERR_UNSUPPORTED, with the loader's reason in
 * ec_last_error_string(), when the library or one of the symbols used is missing. */
int ec_nccl_version(int *version);
/* ndev value for ec_tensor_stream_run_multi: stream through the subset of the box's devices that
 * the host moves data to and from fastest, chosen by measurement (ec_pick_stream_devices). */
enum { EC_DEVICES_AUTO = -1 };
/* The streamed job is bound by the host side of the GPU links, which is not uniform across the
 * GPUs of a box (on this pool's 8-GPU boxes all eight together move LESS than GPUs 4-7 alone:
 * profiles/r02_link_probe.txt).  Measures bidirectional pinned-copy bandwidth on all listed
 * devices (ndev == 0: every device), on pairs of neighbours and on unions of pairs, and returns
 * the subset with the highest aggregate (every device unless a subset is more than 5 % faster).
 * out_devices must hold ndev (or ec_device_count()) entries; out_gbs (optional): [0] = GB/s each
 * way of the chosen set, [1] = of all devices.  Costs about a second the first time, cached per
 * process and device list afterwards. */
int ec_pick_stream_devices(const int *devices, int ndev, int *out_devices, int *out_ndev, double *out_gbs);
/* Planning query; needs no device: the index range [*lo, *hi) of a `count`-element tensor that
 * device number idx of ndev takes in ec_tensor_stream_run_multi. */
int ec_plan_shard(int64_t count, int ndev, int idx, int64_t *lo, int64_t *hi);
/* Planning query; needs no device: the chunk boundaries the tensor-stream engine uses for a job of
 * `count` matrices with at most `chunk` matrices per device chunk; ramp != 0 = the default plan
 * (short first and last chunks), 0 = uniform chunks (what an explicit
 * ec_pipeline_set_stream_config gives).  Writes up to `cap` boundaries (chunk c covers
 * [starts[c], starts[c+1])) and returns how many there are (number of chunks + 1), -1 on error. */
int64_t ec_plan_chunks(int64_t count, int64_t chunk, int ramp, int64_t *starts, int64_t cap);
/* Measured host-link ceiling (the roofline of the streamed configs):
 * plain pinned cudaMemcpyAsync of `bytes` on every listed device at once, best of `reps`.
 * out_gbs[0] = H2D only, [1] = D2H only, [2] = each direction while both run -- GB/s summed over
 * the devices.  ndev == 0 = every device. */
int ec_host_link_probe(const int *devices, int ndev, size_t bytes, int reps, double *out_gbs);
/* The same on the caller's own pinned buffers: device number g streams its slice
 * [g * bytes_per_device, (g+1) * bytes_per_device) of host_in to the device and a device buffer
 * back into its slice of host_out, piece_bytes per copy.  This touches as much unique host memory
 * as the job itself, so the host's last-level cache (which DMA can read from and write into) does
 * not flatter the figure as it can with the small rotating buffers above: it is the ceiling of
 * the tensor stream proper -- its traffic without the products.  host_out is overwritten. */
int ec_host_link_probe_buffers(const int *devices, int ndev, const void *host_in, void *host_out,
                               size_t bytes_per_device, size_t piece_bytes, int reps, double *out_gbs);
/* Leave `sms` streaming multiprocessors unused by the (persistent, register-filling) GEMM kernels
 * so that concurrent kernels -- NCCL's, during the panel broadcasts of the 2D-partitioned DGEMM --
 * can run underneath them instead of queueing behind them.  0 (default) = use every SM. */
int ec_pipeline_set_sm_margin(ec_pipeline *p, int sms);
/* Schedule policy of the GEMM kernels.  stream_k: 1 (default) lets the cost model cut the last
 * wave along k (cooperative launch); 0 never does.  one_cta_per_tile: 1 launches one CTA per
 * output tile instead of a persistent grid when tiles are long (K >= 1024), so that the hardware
 * scheduler balances the work around other kernels sharing the GPU -- NCCL's during the panel
 * broadcasts of the 2D-partitioned DGEMM, where a statically partitioned persistent grid stalls
 * on the SMs NCCL holds.  0 (default) = persistent. */
int ec_pipeline_set_schedule(ec_pipeline *p, int stream_k, int one_cta_per_tile);
/* Engine knobs: matrices per device chunk, ring depth, staging worker threads (0 = default). */
int ec_pipeline_set_stream_config(ec_pipeline *p, int chunk_matrices, int ring_depth,
                                  int staging_threads);

/* ---- one very large DGEMM, 2D block-partitioned over the GPUs of one box ------------------- */

/* BASELINE.json configs[4] (n = 65536 on 8 GPUs).  The reference's gemm is one cublasDgemm on one
 * device with its dimensions narrowed to int (src/cudapipeline.cc:29-31): a product whose
 * operands do not fit one device, or whose element count passes 2^31, cannot be expressed there.
 * Here: a P x Q device grid (ec_grid; P = Q = 0 picks the squarest grid, ndev == 0 every device),
 * matrices cut into P x Q blocks with block (pr, pc) resident on device devices[pr*Q + pc]
 * (ec_dist_matrix), and ec_dgemm_2d: C = A * B by SUMMA -- per k-panel the owning grid column
 * broadcasts its panel of A along the grid rows and the owning grid row its panel of B along the
 * grid columns (NCCL over NVLink, on a side stream, double-buffered), every device accumulates
 * C(pr,pc) += Apanel * Bpanel with the library's DGEMM kernel.  One process, one host thread.
 * EC_ERR_SHAPE on an inner-dimension mismatch, before any device work, like the reference's gemm.
 * `panel` = k-panel width (0 = 4096); *elapsed_ms (optional) = device time, max over the GPUs. */
typedef struct ec_grid ec_grid;
typedef struct ec_dist_matrix ec_dist_matrix;
int ec_grid_create(const int *devices, int ndev, int P, int Q, ec_grid **out);
int ec_grid_destroy(ec_grid *g);
int ec_grid_shape(const ec_grid *g, int *P, int *Q);
int ec_grid_set_backend(ec_grid *g, int backend);
int ec_grid_set_schedule(ec_grid *g, int stream_k, int one_cta_per_tile);
int64_t ec_grid_launch_count(const ec_grid *g);
int64_t ec_grid_broadcast_count(const ec_grid *g);
const char *ec_grid_last_kernel(const ec_grid *g);
int ec_dist_matrix_create(ec_grid *g, int64_t rows, int64_t cols, ec_dist_matrix **out);
int ec_dist_matrix_destroy(ec_dist_matrix *m);
/* whole column-major host matrix (leading dimension ld) <-> blocks; complete on return */
int ec_dist_matrix_upload(ec_dist_matrix *m, const double *host, int64_t ld);
int ec_dist_matrix_download(const ec_dist_matrix *m, double *host, int64_t ld);
/* every block synthesised where it lives: element (i, j) = generator value at index j*rows + i
 * (ec_fill_uniform_block) -- the n = 65536 operands are 34 GB each, more than the host holds */
int ec_dist_matrix_fill_uniform(ec_dist_matrix *m, uint64_t seed, uint64_t matrix_id);
int ec_dist_matrix_get(const ec_dist_matrix *m, int64_t i, int64_t j, double *out);
int ec_dgemm_2d(ec_grid *g, const ec_dist_matrix *A, const ec_dist_matrix *B, ec_dist_matrix *C, int64_t panel,
                double *elapsed_ms);
/* convenience: host operands in, host result out, on a grid made and destroyed inside the call */
int ec_dgemm_2d_host(const int *devices, int ndev, int64_t m, int64_t n, int64_t k, const double *A, int64_t lda,
                     const double *B, int64_t ldb, double *C, int64_t ldc);

/* Planning query; needs no device: the k-panels ec_dgemm_2d walks for inner dimension k on a P x Q grid with
 * panel width `panel` (0 = the default, 4096).  Panels never straddle an ownership boundary of A's column
 * blocks (Q parts) or B's row blocks (P parts), so each has one owning grid column and one owning grid row.
 * Writes up to `cap` boundaries (panel t covers [starts[t], starts[t+1])) and returns how many there are
 * (number of panels + 1; 1 for k = 0), -1 on error. */
int64_t ec_plan_panels(int64_t k, int P, int Q, int64_t panel, int64_t *starts, int64_t cap);

/* ---- host memory + synthetic data helpers ------------------------------------------------- */

/* Pinned host allocation (what CHANGELOG.md:32 records the reference once defaulted to). */
int ec_host_alloc(size_t bytes, void **out);
int ec_host_free(void *ptr);
/* Pin memory the caller already owns -- the storage of an Eigen::MatrixXd, a std::vector<double>
 * -- in place, so that copy_to_gpu, the conversion back and ec_tensor_stream_run move it by DMA
 * without a staging copy.  The range must stay allocated until
 * ec_host_unregister.  Registering is expensive -- measured on the B200 box: 3 ms for 2 MiB,
 * 30 ms for 32 MiB, not faster from several threads (tools/register_probe.py) -- which is why
 * the library stages pageable memory instead of pinning it on the fly: worth it only for
 * buffers that are reused many times.  Zero bytes or an already-pinned range is EC_OK. */
int ec_host_register(void *ptr, size_t bytes);
int ec_host_unregister(void *ptr);
/* x = 2u-1, u in [0,1): splitmix64 on (seed, matrix_id, column-major index); bit-identical to
 * oracle_fill_uniform.  Fills `count` doubles at device pointer dst. */
int ec_fill_uniform(ec_pipeline *p, uint64_t seed, uint64_t matrix_id, double *dst, int64_t count);
/* Same, `batch` matrices of `count` doubles each: matrix b gets matrix_id = first_id + b. */
int ec_fill_uniform_batched(ec_pipeline *p, uint64_t seed, uint64_t first_id, double *dst,
                            int64_t count, int64_t batch);
/* Same recipe for a rows x cols sub-block of a larger column-major matrix: element (i, j) of the
 * block is the generator's value at global index (col0 + j) * global_ld + (row0 + i); written to
 * dst with leading dimension ld_dst.  Lets every rank of the 2D-partitioned DGEMM synthesise its
 * own blocks of the n = 65536 operands (34 GB each: too large for one host). */
int ec_fill_uniform_block(ec_pipeline *p, uint64_t seed, uint64_t matrix_id, double *dst,
                          int64_t rows, int64_t cols, int64_t ld_dst, int64_t row0, int64_t col0,
                          int64_t global_ld);
/* How often the dispatcher of this pipeline has found the GPU shared with somebody else's kernels
 * (CTAs of one of its persistent grids started more than 50 us apart) and switched, for the next
 * 32 launches, from statically partitioned schedules (stream-K, round-robin tiles) to dynamic ones
 * (tiles claimed from a counter; long tiles one CTA each).  See DESIGN.md section 3.1. */
int64_t ec_pipeline_contention_events(const ec_pipeline *p);
/* Test hook: occupy `sms` streaming multiprocessors of `device` for `microseconds` with a kernel
 * that does nothing (one CTA per SM, sized so that no other CTA fits beside it), on `cuda_stream`. */
int ec_debug_hold_sms(int device, void *cuda_stream, int sms, int64_t microseconds);
/* Test hook; needs no device.  The kernels turn a tile id into (problem, tile row, tile column) with a
 * multiply-and-shift division by an invariant divisor (FastDiv in csrc/dgemm_kernels.cuh); this returns
 * what that routine gives for n / d (n < 2^31, 1 <= d < 2^31) so the CPU suite can check it against
 * integer division (0 for d = 0). */
uint32_t ec_debug_fastdiv(uint32_t n, uint32_t d);
/* Launch counter: number of kernels this library has launched on this pipeline since creation. */
int64_t ec_pipeline_launch_count(const ec_pipeline *p);
/* Name of the kernel variant the last ec_dgemm* on this pipeline dispatched to. */
const char *ec_pipeline_last_kernel(const ec_pipeline *p);
/* Planning query; needs no device.  Which CTA tile configuration (1 = 128x128, 2 = 64x64,
 * 3 = 64x32, 4 = 32x32, 5 = 128x16) and schedule (*stream_k 0 = whole tiles per CTA, 1 = stream-K) the
 * dispatcher's cost model picks for `batch` products of m x n x k on a device with `sm_count`
 * SMs.  Honours the EC_FORCE_TILE / EC_STREAM_K overrides like the dispatcher itself. */
int ec_plan_dgemm(int sm_count, int64_t m, int64_t n, int64_t k, int64_t batch, int *tile,
                  int *stream_k);

#ifdef __cplusplus
}
#endif
#endif /* EIGENCUDA_B200_H */
