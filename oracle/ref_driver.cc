// oracle/ref_driver.cc -- C entry points around the UNMODIFIED reference classes.
// TEST INFRASTRUCTURE ONLY (see oracle/dgemm_oracle.c header for who may load it).
//
// This file is compiled together with /root/reference/src/cudamatrix.cc and
// /root/reference/src/cudapipeline.cc (from where they lie; nothing is copied into this repo)
// against oracle/eigen_standin/ and cuBLAS, into oracle/_ref/libeigencuda_ref.so.  Every function
// below drives the reference exclusively through its public API (eigencuda::CudaPipeline,
// eigencuda::CudaMatrix) the way its own tests do (src/tests/test_dot.cc), so that Python can
// (a) check the CPU restatement and the new kernels against the real reference on a GPU box and
// (b) time the reference's stock code path for `bench.py --impl reference`.
#include <chrono>
#include <cstring>
#include <stdexcept>
#include <vector>

#include "cudamatrix.hpp"    // /root/reference/include
#include "cudapipeline.hpp"  // /root/reference/include

using eigencuda::CudaMatrix;
using eigencuda::CudaPipeline;
using eigencuda::Index;

namespace {
Eigen::MatrixXd from_ptr(const double *p, Index r, Index c) {
  Eigen::MatrixXd m = Eigen::MatrixXd::Zero(r, c);
  std::memcpy(m.data(), p, sizeof(double) * static_cast<size_t>(r * c));
  return m;
}
}  // namespace

extern "C" {

// src/tests/test_dot.cc:11-32 (create_cudamatrix): H2D in the ctor, D2H in the conversion.
int ref_roundtrip(const double *src, long rows, long cols, double *dst) {
  try {
    CudaPipeline cp;
    Eigen::MatrixXd B = from_ptr(src, rows, cols);
    CudaMatrix cumatrix{B, cp.get_stream()};
    Eigen::MatrixXd tmp = cumatrix;
    std::memcpy(dst, tmp.data(), sizeof(double) * static_cast<size_t>(rows * cols));
    return 0;
  } catch (const std::runtime_error &) {
    return 1;
  }
}

// src/tests/test_dot.cc:34-51 (matrix_multiplication) and :92-103 (wrong_shape_cublas):
// C = A * B through CudaPipeline::gemm; returns 1 when the reference throws std::runtime_error.
int ref_gemm(const double *A, long a_rows, long a_cols, const double *B, long b_rows, long b_cols,
             double *C) {
  try {
    Eigen::MatrixXd eA = from_ptr(A, a_rows, a_cols), eB = from_ptr(B, b_rows, b_cols);
    CudaPipeline cuda_pip;
    CudaMatrix cuma_A{eA, cuda_pip.get_stream()};
    CudaMatrix cuma_B{eB, cuda_pip.get_stream()};
    CudaMatrix cuma_C{a_rows, b_cols, cuda_pip.get_stream()};
    cuda_pip.gemm(cuma_A, cuma_B, cuma_C);
    Eigen::MatrixXd eC = cuma_C;
    std::memcpy(C, eC.data(), sizeof(double) * static_cast<size_t>(a_rows * b_cols));
    return 0;
  } catch (const std::runtime_error &) {
    return 1;
  }
}

// The streaming pattern, README.md:58-66 == src/tests/test_dot.cc:77-85:
//   results[i] = tensor[i] * F  (side = 0, the reference's order)  or  F * tensor[i]  (side = 1).
// The loop is timed with steady_clock from the first copy_to_gpu to the last result being
// resident in host memory; building the std::vector<MatrixXd> inputs and copying the results out
// to the caller happen outside the timed region.  Returns seconds in *elapsed.
int ref_tensor_stream(int side, const double *F, long f_rows, long f_cols,
                      const double *const *tensor, double *const *results, long count, long t_rows,
                      long t_cols, double *elapsed) {
  try {
    CudaPipeline cuda_pip;
    Eigen::MatrixXd eF = from_ptr(F, f_rows, f_cols);
    std::vector<Eigen::MatrixXd> ts;
    ts.reserve(count);
    for (long i = 0; i < count; ++i) ts.push_back(from_ptr(tensor[i], t_rows, t_cols));
    Index r_rows = side == 0 ? t_rows : f_rows, r_cols = side == 0 ? f_cols : t_cols;
    std::vector<Eigen::MatrixXd> res(count, Eigen::MatrixXd::Zero(r_rows, r_cols));
    CudaMatrix cuma_A{eF, cuda_pip.get_stream()};
    CudaMatrix cuma_B{t_rows, t_cols, cuda_pip.get_stream()};
    CudaMatrix cuma_C{r_rows, r_cols, cuda_pip.get_stream()};
    cudaDeviceSynchronize();
    auto t0 = std::chrono::steady_clock::now();
    for (Index i = 0; i < count; i++) {
      cuma_B.copy_to_gpu(ts[i]);
      if (side == 0)
        cuda_pip.gemm(cuma_B, cuma_A, cuma_C);
      else
        cuda_pip.gemm(cuma_A, cuma_B, cuma_C);
      res[i] = cuma_C;
    }
    auto t1 = std::chrono::steady_clock::now();
    if (elapsed) *elapsed = std::chrono::duration<double>(t1 - t0).count();
    for (long i = 0; i < count; ++i)
      std::memcpy(results[i], res[i].data(), sizeof(double) * static_cast<size_t>(r_rows * r_cols));
    return 0;
  } catch (const std::runtime_error &) {
    return 1;
  }
}

// The same loop for a benchmark: inputs and result vector are built once, the loop runs
// `warmup` untimed and `steps` timed passes over the whole tensor; seconds per timed pass go to
// elapsed[0..steps).  The results of the last pass are copied out to the caller.
int ref_tensor_stream_steps(int side, const double *F, long f_rows, long f_cols,
                            const double *const *tensor, double *const *results, long count,
                            long t_rows, long t_cols, int warmup, int steps, double *elapsed) {
  try {
    CudaPipeline cuda_pip;
    Eigen::MatrixXd eF = from_ptr(F, f_rows, f_cols);
    std::vector<Eigen::MatrixXd> ts;
    ts.reserve(count);
    for (long i = 0; i < count; ++i) ts.push_back(from_ptr(tensor[i], t_rows, t_cols));
    Index r_rows = side == 0 ? t_rows : f_rows, r_cols = side == 0 ? f_cols : t_cols;
    std::vector<Eigen::MatrixXd> res(count, Eigen::MatrixXd::Zero(r_rows, r_cols));
    CudaMatrix cuma_A{eF, cuda_pip.get_stream()};
    CudaMatrix cuma_B{t_rows, t_cols, cuda_pip.get_stream()};
    CudaMatrix cuma_C{r_rows, r_cols, cuda_pip.get_stream()};
    for (int s = 0; s < warmup + steps; ++s) {
      cudaDeviceSynchronize();
      auto t0 = std::chrono::steady_clock::now();
      for (Index i = 0; i < count; i++) {
        cuma_B.copy_to_gpu(ts[i]);
        if (side == 0)
          cuda_pip.gemm(cuma_B, cuma_A, cuma_C);
        else
          cuda_pip.gemm(cuma_A, cuma_B, cuma_C);
        res[i] = cuma_C;
      }
      auto t1 = std::chrono::steady_clock::now();
      if (elapsed && s >= warmup) elapsed[s - warmup] = std::chrono::duration<double>(t1 - t0).count();
    }
    if (results)
      for (long i = 0; i < count; ++i)
        if (results[i])
          std::memcpy(results[i], res[i].data(), sizeof(double) * static_cast<size_t>(r_rows * r_cols));
    return 0;
  } catch (const std::runtime_error &) {
    return 1;
  }
}

// Device-resident timing of the reference's gemm (== cublasDgemm NN on the legacy stream,
// src/cudapipeline.cc:29-31) on n x n operands: `reps` back-to-back calls bracketed by CUDA events
// on the stream cuBLAS actually runs on (the NULL stream; the reference never calls
// cublasSetStream).  Returns average milliseconds per gemm in *ms.
int ref_gemm_time(long n, int warmup, int reps, double *ms) {
  try {
    CudaPipeline cuda_pip;
    Eigen::MatrixXd A = Eigen::MatrixXd::Random(n, n), B = Eigen::MatrixXd::Random(n, n);
    CudaMatrix cuma_A{A, cuda_pip.get_stream()};
    CudaMatrix cuma_B{B, cuda_pip.get_stream()};
    CudaMatrix cuma_C{n, n, cuda_pip.get_stream()};
    for (int i = 0; i < warmup; ++i) cuda_pip.gemm(cuma_A, cuma_B, cuma_C);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaDeviceSynchronize();
    cudaEventRecord(e0, 0);
    for (int i = 0; i < reps; ++i) cuda_pip.gemm(cuma_A, cuma_B, cuma_C);
    cudaEventRecord(e1, 0);
    cudaEventSynchronize(e1);
    float t = 0;
    cudaEventElapsedTime(&t, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (ms) *ms = static_cast<double>(t) / reps;
    return cudaGetLastError() == cudaSuccess ? 0 : 2;
  } catch (const std::runtime_error &) {
    return 1;
  }
}

long ref_count_available_gpus() { return static_cast<long>(eigencuda::count_available_gpus()); }

}  // extern "C"
