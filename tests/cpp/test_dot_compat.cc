// tests/cpp/test_dot_compat.cc -- the reference's four test cases (src/tests/test_dot.cc:11-103:
// create_cudamatrix, matrix_multiplication, right_matrix_multiplication, wrong_shape_cublas)
// restated against include/cudamatrix.hpp + include/cudapipeline.hpp of THIS repo, to show that
// caller code written for the reference compiles and behaves the same.  Boost.Test is not in the
// image, so a ten-line harness stands in; Eigen is the stand-in from oracle/eigen_standin when
// the real one is absent.  Needs a B200: run by tests/test_gpu_parity.py (-m gpu).
#include <cstdio>
#include <stdexcept>
#include <vector>

#include "cudamatrix.hpp"
#include "cudapipeline.hpp"

using eigencuda::CudaMatrix;
using eigencuda::CudaPipeline;
using eigencuda::Index;

static int failures = 0;
#define CHECK(cond)                                                     \
  do {                                                                  \
    if (!(cond)) {                                                      \
      std::printf("FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond);       \
      ++failures;                                                       \
    }                                                                   \
  } while (0)

// Host -> device -> host round trip, then a CPU product against the known answer.
static void create_cudamatrix() {
  CudaPipeline cp;
  Eigen::MatrixXd A = Eigen::MatrixXd::Zero(2, 2);
  Eigen::MatrixXd B = Eigen::MatrixXd::Zero(3, 2);
  Eigen::MatrixXd X = Eigen::MatrixXd::Zero(3, 2);
  A << 1., 2., 3., 4.;
  B << 5., 6., 7., 8., 9., 10.;
  X << 23., 34., 31., 46., 39., 58.;
  CudaMatrix cumatrix{B, cp.get_stream()};
  Eigen::MatrixXd tmp = cumatrix;
  Eigen::MatrixXd result = tmp * A;
  CHECK(X.isApprox(result));
}

// 200 x 200 random product against the CPU product, Eigen's default isApprox (1e-12).
static void matrix_multiplication() {
  Index dim = 200;
  Eigen::MatrixXd A = Eigen::MatrixXd::Random(dim, dim);
  Eigen::MatrixXd B = Eigen::MatrixXd::Random(dim, dim);
  CudaPipeline cuda_pip;
  CudaMatrix cuma_A{A, cuda_pip.get_stream()};
  CudaMatrix cuma_B{B, cuda_pip.get_stream()};
  CudaMatrix cuma_C{dim, dim, cuda_pip.get_stream()};
  cuda_pip.gemm(cuma_A, cuma_B, cuma_C);
  Eigen::MatrixXd C = cuma_C;
  Eigen::MatrixXd result = A * B;
  CHECK(C.isApprox(result));
}

// The streaming pattern: one fixed matrix, a tensor of matrices uploaded one by one.
static void right_matrix_multiplication() {
  CudaPipeline cuda_pip;
  Eigen::MatrixXd A = Eigen::MatrixXd::Zero(2, 2);
  Eigen::MatrixXd B = Eigen::MatrixXd::Zero(3, 2), C = B, D = B, X = B, Y = B, Z = B;
  A << 1., 2., 3., 4.;
  B << 5., 6., 7., 8., 9., 10.;
  C << 9., 10., 11., 12., 13., 14.;
  D << 13., 14., 15., 16., 17., 18.;
  X << 23., 34., 31., 46., 39., 58.;
  Y << 39., 58., 47., 70., 55., 82.;
  Z << 55., 82., 63., 94., 71., 106.;
  std::vector<Eigen::MatrixXd> tensor{B, C, D};
  std::vector<Eigen::MatrixXd> results(3, Eigen::MatrixXd::Zero(3, 2));
  CudaMatrix cuma_A{A, cuda_pip.get_stream()};
  CudaMatrix cuma_B{3, 2, cuda_pip.get_stream()};
  CudaMatrix cuma_C{3, 2, cuda_pip.get_stream()};
  for (Index i = 0; i < 3; i++) {
    cuma_B.copy_to_gpu(tensor[i]);
    cuda_pip.gemm(cuma_B, cuma_A, cuma_C);
    results[i] = cuma_C;
  }
  CHECK(X.isApprox(results[0]));
  CHECK(Y.isApprox(results[1]));
  CHECK(Z.isApprox(results[2]));

  // the same job through the additive tensor-level call
  std::vector<Eigen::MatrixXd> r2 = cuda_pip.right_multiply(tensor, A);
  CHECK(r2.size() == 3);
  CHECK(X.isApprox(r2[0]));
  CHECK(Y.isApprox(r2[1]));
  CHECK(Z.isApprox(r2[2]));
}

// Inner-dimension mismatch throws std::runtime_error.
static void wrong_shape_cublas() {
  Eigen::MatrixXd A = Eigen::MatrixXd::Random(2, 2);
  Eigen::MatrixXd B = Eigen::MatrixXd::Random(5, 5);
  CudaPipeline cuda_pip;
  CudaMatrix cuma_A{A, cuda_pip.get_stream()};
  CudaMatrix cuma_B{B, cuda_pip.get_stream()};
  CudaMatrix cuma_C{2, 5, cuda_pip.get_stream()};
  bool thrown = false;
  try {
    cuda_pip.gemm(cuma_A, cuma_B, cuma_C);
  } catch (const std::runtime_error &) {
    thrown = true;
  }
  CHECK(thrown);
}

// Additive API: transposes, basis transform, larger streamed tensor (pageable host memory).
static void additive_api() {
  CudaPipeline cuda_pip;
  const Index n = 96, count = 7;
  Eigen::MatrixXd F = Eigen::MatrixXd::Random(n, n);
  std::vector<Eigen::MatrixXd> tensor;
  for (Index i = 0; i < count; ++i) tensor.push_back(Eigen::MatrixXd::Random(n, n));
  auto r_right = cuda_pip.right_multiply(tensor, F);
  auto r_leftT = cuda_pip.left_multiply(F, tensor, true);
  auto r_basis = cuda_pip.basis_transform(F, tensor);
  Eigen::MatrixXd Ft = F.transpose();
  for (Index i = 0; i < count; ++i) {
    CHECK(r_right[i].isApprox(tensor[i] * F));
    CHECK(r_leftT[i].isApprox(Ft * tensor[i]));
    CHECK(r_basis[i].isApprox((Ft * tensor[i]) * F));
  }
  CudaMatrix dA{F, cuda_pip.get_stream()}, dB{tensor[0], cuda_pip.get_stream()};
  CudaMatrix dC{n, n, cuda_pip.get_stream()};
  cuda_pip.gemm(true, true, 1.0, dA, dB, 0.0, dC);
  Eigen::MatrixXd C = dC;
  CHECK(C.isApprox(Ft * tensor[0].transpose()));
}

int main() {
  create_cudamatrix();
  matrix_multiplication();
  right_matrix_multiplication();
  wrong_shape_cublas();
  additive_api();
  if (failures == 0) std::printf("ALL PASSED\n");
  return failures == 0 ? 0 : 1;
}
