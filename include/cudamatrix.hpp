// include/cudamatrix.hpp -- source-compatible replacement for the reference header of the same
// name (NLESC-JCER/EigenCuda include/cudamatrix.hpp:1-64).  Same namespace, same class, same
// member signatures, same exceptions; the bodies are thin inline calls into the C ABI
// (eigencuda_b200.h), so a caller recompiles against this header and links libeigencuda_b200.so
// instead of the reference library + cuBLAS.  No cublas_v2.h / curand.h here (the reference
// includes both, uses neither type in this class; Appendix A #10 of SURVEY.md).
#ifndef EIGENCUDA_H_
#define EIGENCUDA_H_

#include <Eigen/Core>
#include <Eigen/Dense>
#include <cuda_runtime_api.h>

#include <memory>
#include <stdexcept>
#include <utility>
#include <vector>
#include <string>

#include "eigencuda_b200.h"

namespace eigencuda {

// reference include/cudamatrix.hpp:22, src/cudamatrix.cc:5-12: returns its argument (the
// reference only logs under -DDEBUG).
inline cudaError_t checkCuda(cudaError_t result) { return result; }

using Index = Eigen::Index;  // reference include/cudamatrix.hpp:24

// reference include/cudamatrix.hpp:25, src/cudamatrix.cc:14-18.
inline Index count_available_gpus() { return static_cast<Index>(ec_device_count()); }

namespace detail {
// Status code -> the exception type the reference throws (std::runtime_error everywhere:
// src/cudapipeline.cc:27, src/cudamatrix.cc:29,77).
inline void throw_on_error(int rc) {
  if (rc != EC_OK) throw std::runtime_error(ec_last_error_string());
}
}  // namespace detail

// Additive: pins the storage of host matrices the caller already owns, in place, for as long as
// this object lives.  Transfers from / to pinned memory are plain
// DMA: copy_to_gpu, the conversion back and the tensor-level calls skip their staging copies.
// Worth it for matrices that are transferred more than once; the matrices must not be resized or
// destroyed while pinned.
class HostPin {
 public:
  HostPin() = default;
  explicit HostPin(const Eigen::MatrixXd &m) { add(m); }
  explicit HostPin(const std::vector<Eigen::MatrixXd> &tensor) {
    try {
      for (const Eigen::MatrixXd &m : tensor) add(m);
    } catch (...) {
      release();   // the destructor does not run for a constructor that throws
      throw;
    }
  }
  // Returns whether the matrix is pinned now.  A small heap matrix can share its first or last
  // page with a neighbour that is already pinned; such a range cannot be registered as a whole
  // and is simply left pageable (transfers of it are staged, as without HostPin).
  bool add(const Eigen::MatrixXd &m) {
    if (m.size() == 0) return true;
    void *ptr = const_cast<double *>(m.data());
    const int rc = ec_host_register(ptr, static_cast<size_t>(m.size()) * sizeof(double));
    if (rc == EC_ERR_ARG) return false;
    detail::throw_on_error(rc);
    _ranges.push_back(ptr);
    return true;
  }
  ~HostPin() { release(); }
  HostPin(const HostPin &) = delete;
  HostPin &operator=(const HostPin &) = delete;
  HostPin(HostPin &&other) noexcept : _ranges(std::move(other._ranges)) { other._ranges.clear(); }

 private:
  void release() {
    for (void *ptr : _ranges) ec_host_unregister(ptr);
    _ranges.clear();
  }
  std::vector<void *> _ranges;
};

class CudaPipeline;

class CudaMatrix {
 public:
  Index size() const { return _rows * _cols; };
  Index rows() const { return _rows; };
  Index cols() const { return _cols; };
  double *data() const { return ec_matrix_data(_m.get()); };

  // reference src/cudamatrix.cc:20-31: allocate on the device and copy `matrix` asynchronously on
  // `stream`; std::runtime_error when the device is out of memory or the copy cannot be issued.
  CudaMatrix(const Eigen::MatrixXd &matrix, const cudaStream_t &stream)
      : _rows{static_cast<Index>(matrix.rows())}, _cols{static_cast<Index>(matrix.cols())} {
    ec_matrix *m = nullptr;
    detail::throw_on_error(
        ec_matrix_create_from_host(stream, matrix.data(), matrix.rows(), matrix.cols(), &m));
    _m.reset(m);
  }

  // reference src/cudamatrix.cc:33-37: allocate memory in the GPU for a matrix (uninitialised).
  CudaMatrix(Index nrows, Index ncols, const cudaStream_t &stream) : _rows{nrows}, _cols{ncols} {
    ec_matrix *m = nullptr;
    detail::throw_on_error(ec_matrix_alloc(stream, nrows, ncols, &m));
    _m.reset(m);
  }

  // reference src/cudamatrix.cc:39-45: convert a CudaMatrix to an Eigen matrix (blocking).
  operator Eigen::MatrixXd() const {
    Eigen::MatrixXd result(this->rows(), this->cols());
    detail::throw_on_error(ec_matrix_download(_m.get(), result.data()));
    return result;
  }

  // reference src/cudamatrix.cc:47-51: asynchronous H2D into the existing buffer.
  void copy_to_gpu(const Eigen::MatrixXd &A) {
    detail::throw_on_error(ec_matrix_upload(_m.get(), A.data(), static_cast<int64_t>(A.size())));
  }

  const ec_matrix *handle() const { return _m.get(); }
  ec_matrix *handle() { return _m.get(); }

 private:
  struct Deleter {
    void operator()(ec_matrix *m) const { ec_matrix_free(m); }
  };
  // Move-only ownership of the device buffer, like the reference's
  // unique_ptr<double, void(*)(double*)> (include/cudamatrix.hpp:46,55-56).
  std::unique_ptr<ec_matrix, Deleter> _m;
  Index _rows;
  Index _cols;
};

}  // namespace eigencuda

#endif  // EIGENCUDA_H_
