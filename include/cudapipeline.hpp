// include/cudapipeline.hpp -- source-compatible replacement for the reference header of the same
// name (NLESC-JCER/EigenCuda include/cudapipeline.hpp:1-46).  The class owns a stream, as the
// reference's does; it owns no cuBLAS handle, because gemm() launches this library's own sm_100a
// kernels through the C ABI (eigencuda_b200.h).
#ifndef CUDA_PIPELINE__H
#define CUDA_PIPELINE__H

#include <vector>

#include "cudamatrix.hpp"

namespace eigencuda {

class CudaPipeline {
 public:
  // reference include/cudapipeline.hpp:22-25 (cublasCreate + cudaStreamCreate).
  CudaPipeline() { detail::throw_on_error(ec_pipeline_create(-1, &_p)); }
  // Additive: choose the device explicitly (the reference never calls cudaSetDevice).
  explicit CudaPipeline(int device) { detail::throw_on_error(ec_pipeline_create(device, &_p)); }
  // reference src/cudapipeline.cc:5-11.
  ~CudaPipeline() { ec_pipeline_destroy(_p); }

  CudaPipeline(const CudaPipeline &) = delete;
  CudaPipeline &operator=(const CudaPipeline &) = delete;

  // reference src/cudapipeline.cc:17-32: C = A * B.  std::runtime_error on an inner-dimension
  // mismatch, raised before any device work.
  void gemm(const CudaMatrix &A, const CudaMatrix &B, CudaMatrix &C) const {
    detail::throw_on_error(ec_gemm(_p, A.handle(), B.handle(), C.handle()));
  }

  // reference include/cudapipeline.hpp:34.
  const cudaStream_t &get_stream() const {
    return *static_cast<const cudaStream_t *>(ec_pipeline_stream_ref(_p));
  };

  // ---------------------------------------------------------------------------------------
  // Additive API (none of this exists in the reference; CHANGELOG.md:16-25 records that the
  // triple product and the batched path were removed).  Parity for these is pinned against
  // cuBLAS called with the corresponding flags and against the CPU oracle, see DESIGN.md.
  // ---------------------------------------------------------------------------------------

  // C = alpha * op(A) * op(B) + beta * C
  void gemm(bool transA, bool transB, double alpha, const CudaMatrix &A, const CudaMatrix &B,
            double beta, CudaMatrix &C) const {
    const Index m = transA ? A.cols() : A.rows(), k = transA ? A.rows() : A.cols();
    const Index kb = transB ? B.cols() : B.rows(), n = transB ? B.rows() : B.cols();
    if (k != kb) throw std::runtime_error("Shape mismatch in Cublas gemm");
    detail::throw_on_error(ec_dgemm(_p, transA ? EC_OP_T : EC_OP_N, transB ? EC_OP_T : EC_OP_N, m,
                                    n, k, alpha, A.data(), A.rows(), B.data(), B.rows(), beta,
                                    C.data(), C.rows()));
  }

  // results[i] = tensor[i] * A -- the README loop (reference README.md:58-66) as one overlapped
  // job: H2D, product and D2H of different tensor elements run concurrently.
  std::vector<Eigen::MatrixXd> right_multiply(const std::vector<Eigen::MatrixXd> &tensor,
                                              const Eigen::MatrixXd &A) const {
    return stream(EC_STREAM_RIGHT, A, tensor);
  }
  // same, into result matrices the caller keeps (reused in place when they have the right shape)
  void right_multiply(const std::vector<Eigen::MatrixXd> &tensor, const Eigen::MatrixXd &A,
                      std::vector<Eigen::MatrixXd> &results) const {
    stream_into(EC_STREAM_RIGHT, A, tensor, results);
  }
  // results[i] = A * tensor[i]   (transposeA: A^T * tensor[i])
  std::vector<Eigen::MatrixXd> left_multiply(const Eigen::MatrixXd &A,
                                             const std::vector<Eigen::MatrixXd> &tensor,
                                             bool transposeA = false) const {
    return stream(transposeA ? EC_STREAM_LEFT_T : EC_STREAM_LEFT, A, tensor);
  }
  void left_multiply(const Eigen::MatrixXd &A, const std::vector<Eigen::MatrixXd> &tensor,
                     std::vector<Eigen::MatrixXd> &results, bool transposeA = false) const {
    stream_into(transposeA ? EC_STREAM_LEFT_T : EC_STREAM_LEFT, A, tensor, results);
  }
  // results[i] = A^T * tensor[i] * A
  std::vector<Eigen::MatrixXd> basis_transform(const Eigen::MatrixXd &A,
                                               const std::vector<Eigen::MatrixXd> &tensor) const {
    return stream(EC_STREAM_BASIS, A, tensor);
  }

  void basis_transform(const Eigen::MatrixXd &A, const std::vector<Eigen::MatrixXd> &tensor,
                       std::vector<Eigen::MatrixXd> &results) const {
    stream_into(EC_STREAM_BASIS, A, tensor, results);
  }

  // ---------------------------------------------------------------------------------------
  // The same calls over SEVERAL GPUs of the box -- the multi-GPU scheduler
  // (ec_tensor_stream_run_multi).  `devices` lists CUDA device indices; an empty list means every
  // device (count_available_gpus() of them).  The caller's one vector is sharded into contiguous
  // index ranges, one worker thread + pinned ring + three streams per device; A is uploaded once
  // and broadcast with NCCL over NVLink; every device writes its results straight into
  // results[i].  The pipeline's own device need not be in the list.
  // ---------------------------------------------------------------------------------------
  // The GPUs of the box the host streams through fastest, chosen by measurement once per process
  // (ec_pick_stream_devices): the host side of the links is not uniform, and more GPUs can be
  // less throughput for this transfer-bound job.  Pass the result as `devices`.
  static std::vector<int> auto_devices() {
    std::vector<int> out(static_cast<size_t>(ec_device_count() > 0 ? ec_device_count() : 1));
    int n = 0;
    detail::throw_on_error(ec_pick_stream_devices(nullptr, 0, out.data(), &n, nullptr));
    out.resize(static_cast<size_t>(n));
    return out;
  }

  std::vector<Eigen::MatrixXd> right_multiply(const std::vector<Eigen::MatrixXd> &tensor,
                                              const Eigen::MatrixXd &A,
                                              const std::vector<int> &devices) const {
    return stream(EC_STREAM_RIGHT, A, tensor, &devices);
  }
  void right_multiply(const std::vector<Eigen::MatrixXd> &tensor, const Eigen::MatrixXd &A,
                      std::vector<Eigen::MatrixXd> &results, const std::vector<int> &devices) const {
    stream_into(EC_STREAM_RIGHT, A, tensor, results, &devices);
  }
  std::vector<Eigen::MatrixXd> left_multiply(const Eigen::MatrixXd &A,
                                             const std::vector<Eigen::MatrixXd> &tensor,
                                             const std::vector<int> &devices,
                                             bool transposeA = false) const {
    return stream(transposeA ? EC_STREAM_LEFT_T : EC_STREAM_LEFT, A, tensor, &devices);
  }
  void left_multiply(const Eigen::MatrixXd &A, const std::vector<Eigen::MatrixXd> &tensor,
                     std::vector<Eigen::MatrixXd> &results, const std::vector<int> &devices,
                     bool transposeA = false) const {
    stream_into(transposeA ? EC_STREAM_LEFT_T : EC_STREAM_LEFT, A, tensor, results, &devices);
  }
  std::vector<Eigen::MatrixXd> basis_transform(const Eigen::MatrixXd &A,
                                               const std::vector<Eigen::MatrixXd> &tensor,
                                               const std::vector<int> &devices) const {
    return stream(EC_STREAM_BASIS, A, tensor, &devices);
  }
  void basis_transform(const Eigen::MatrixXd &A, const std::vector<Eigen::MatrixXd> &tensor,
                       std::vector<Eigen::MatrixXd> &results, const std::vector<int> &devices) const {
    stream_into(EC_STREAM_BASIS, A, tensor, results, &devices);
  }

  ec_pipeline *handle() const { return _p; }

 private:
  std::vector<Eigen::MatrixXd> stream(int kind, const Eigen::MatrixXd &A,
                                      const std::vector<Eigen::MatrixXd> &tensor,
                                      const std::vector<int> *devices = nullptr) const {
    std::vector<Eigen::MatrixXd> results;
    stream_into(kind, A, tensor, results, devices);
    return results;
  }

  // Elements of `results` that already have the result shape are written in place (so a caller
  // that keeps -- and perhaps pins -- its result matrices across calls pays no allocation); the
  // vector is resized and the other elements are reallocated.
  void stream_into(int kind, const Eigen::MatrixXd &A, const std::vector<Eigen::MatrixXd> &tensor,
                   std::vector<Eigen::MatrixXd> &results,
                   const std::vector<int> *devices = nullptr) const {
    results.resize(tensor.size());
    if (tensor.empty()) return;
    const Index tr = tensor[0].rows(), tc = tensor[0].cols();
    Index rr = 0, rc = 0;
    switch (kind) {
      case EC_STREAM_RIGHT: rr = tr; rc = A.cols(); break;
      case EC_STREAM_LEFT: rr = A.rows(); rc = tc; break;
      case EC_STREAM_LEFT_T: rr = A.cols(); rc = tc; break;
      default: rr = A.cols(); rc = A.cols(); break;
    }
    std::vector<const double *> in(tensor.size());
    std::vector<double *> out(tensor.size());
    for (size_t i = 0; i < tensor.size(); ++i) {
      if (tensor[i].rows() != tr || tensor[i].cols() != tc)
        throw std::runtime_error("Shape mismatch in Cublas gemm");
      if (results[i].rows() != rr || results[i].cols() != rc) results[i] = Eigen::MatrixXd(rr, rc);
      in[i] = tensor[i].data();
      out[i] = results[i].data();
    }
    if (devices)
      detail::throw_on_error(ec_tensor_stream_run_multi(
          _p, devices->data(), static_cast<int>(devices->size()), kind, A.data(), A.rows(), A.cols(),
          in.data(), out.data(), static_cast<int64_t>(tensor.size()), tr, tc));
    else
      detail::throw_on_error(ec_tensor_stream_run(_p, kind, A.data(), A.rows(), A.cols(), in.data(),
                                                  out.data(), static_cast<int64_t>(tensor.size()),
                                                  tr, tc));
  }

  ec_pipeline *_p = nullptr;
};

}  // namespace eigencuda

#endif
