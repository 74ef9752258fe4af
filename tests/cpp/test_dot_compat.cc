// tests/cpp/test_dot_compat.cc -- caller-side C++ against include/cudamatrix.hpp and
// include/cudapipeline.hpp of THIS repo: the four behaviours the reference's Boost tests pin
// (src/tests/test_dot.cc:11-103 -- device round trip, 200x200 product vs the CPU product, the
// streamed right-multiplication with its known answers, the shape-mismatch throw) plus the
// additive API, written table-driven with a ten-line harness (Boost.Test is not in the image).
// The point is that code using eigencuda::CudaPipeline / CudaMatrix the way the reference's
// callers do -- stream from get_stream(), construction from a matrix or from a shape, copy_to_gpu,
// gemm, implicit conversion back -- compiles and behaves the same.  Eigen is the stand-in from
// oracle/eigen_standin when the real one is absent.  Needs a B200: run by tests/test_gpu_parity.py.
#include <cstdio>
#include <initializer_list>
#include <stdexcept>
#include <vector>

#include "cudamatrix.hpp"
#include "cudapipeline.hpp"

namespace ecu = eigencuda;
using Mat = Eigen::MatrixXd;

static int g_failed = 0;
static void expect(bool ok, const char *what, int line) {
  if (!ok) {
    std::printf("FAIL line %d: %s\n", line, what);
    ++g_failed;
  }
}
#define EXPECT(cond) expect((cond), #cond, __LINE__)

// rows x cols matrix from values listed row by row (what Eigen's comma initialiser does)
static Mat from_rows(ecu::Index rows, ecu::Index cols, std::initializer_list<double> values) {
  Mat m = Mat::Zero(rows, cols);
  ecu::Index n = 0;
  for (double v : values) {
    m(n / cols, n % cols) = v;
    ++n;
  }
  return m;
}

// golden vectors: tests/golden/readme_example.json (the reference's own known answers)
static const Mat kFixed = from_rows(2, 2, {1, 2, 3, 4});
static const std::vector<Mat> kTensor = {from_rows(3, 2, {5, 6, 7, 8, 9, 10}), from_rows(3, 2, {9, 10, 11, 12, 13, 14}),
                                         from_rows(3, 2, {13, 14, 15, 16, 17, 18})};
static const std::vector<Mat> kExpected = {from_rows(3, 2, {23, 34, 31, 46, 39, 58}),
                                           from_rows(3, 2, {39, 58, 47, 70, 55, 82}),
                                           from_rows(3, 2, {55, 82, 63, 94, 71, 106})};

static void device_round_trip() {
  ecu::CudaPipeline pipeline;
  ecu::CudaMatrix on_device{kTensor[0], pipeline.get_stream()};   // H2D in the constructor
  Mat back = on_device;                                          // D2H in the conversion
  EXPECT(back.rows() == 3 && back.cols() == 2 && on_device.size() == 6);
  EXPECT(kExpected[0].isApprox(back * kFixed));
}

static void square_product_against_cpu() {
  const ecu::Index n = 200;
  Mat lhs = Mat::Random(n, n), rhs = Mat::Random(n, n);
  ecu::CudaPipeline pipeline;
  ecu::CudaMatrix d_lhs{lhs, pipeline.get_stream()}, d_rhs{rhs, pipeline.get_stream()};
  ecu::CudaMatrix d_out{n, n, pipeline.get_stream()};
  pipeline.gemm(d_lhs, d_rhs, d_out);
  Mat got = d_out;
  EXPECT(got.isApprox(lhs * rhs));
}

static void streamed_right_multiplication() {
  ecu::CudaPipeline pipeline;
  ecu::CudaMatrix d_fixed{kFixed, pipeline.get_stream()};
  ecu::CudaMatrix d_elem{3, 2, pipeline.get_stream()}, d_out{3, 2, pipeline.get_stream()};
  std::vector<Mat> results(kTensor.size(), Mat::Zero(3, 2));
  for (size_t i = 0; i < kTensor.size(); ++i) {   // the loop of the reference's README
    d_elem.copy_to_gpu(kTensor[i]);
    pipeline.gemm(d_elem, d_fixed, d_out);
    results[i] = d_out;
  }
  for (size_t i = 0; i < kTensor.size(); ++i) EXPECT(kExpected[i].isApprox(results[i]));
  // the same job through the additive tensor-level call
  std::vector<Mat> overlapped = pipeline.right_multiply(kTensor, kFixed);
  EXPECT(overlapped.size() == kTensor.size());
  for (size_t i = 0; i < overlapped.size(); ++i) EXPECT(kExpected[i].isApprox(overlapped[i]));
}

static void inner_dimension_mismatch_throws() {
  ecu::CudaPipeline pipeline;
  ecu::CudaMatrix d_a{Mat::Random(2, 2), pipeline.get_stream()}, d_b{Mat::Random(5, 5), pipeline.get_stream()};
  ecu::CudaMatrix d_c{2, 5, pipeline.get_stream()};
  bool thrown = false;
  try {
    pipeline.gemm(d_a, d_b, d_c);
  } catch (const std::runtime_error &) {
    thrown = true;
  }
  EXPECT(thrown);
}

static void additive_api() {
  ecu::CudaPipeline pipeline;
  const ecu::Index n = 96, count = 7;
  Mat fixed = Mat::Random(n, n), fixed_t = fixed.transpose();
  std::vector<Mat> tensor;
  for (ecu::Index i = 0; i < count; ++i) tensor.push_back(Mat::Random(n, n));
  std::vector<Mat> right = pipeline.right_multiply(tensor, fixed);
  std::vector<Mat> left_t = pipeline.left_multiply(fixed, tensor, true);
  std::vector<Mat> basis = pipeline.basis_transform(fixed, tensor);
  for (ecu::Index i = 0; i < count; ++i) {
    EXPECT(right[i].isApprox(tensor[i] * fixed));
    EXPECT(left_t[i].isApprox(fixed_t * tensor[i]));
    EXPECT(basis[i].isApprox((fixed_t * tensor[i]) * fixed));
  }
  {
    ecu::HostPin pin_tensor(tensor), pin_fixed(fixed);   // caller memory pinned in place: same results
    std::vector<Mat> right_pinned = pipeline.right_multiply(tensor, fixed);
    for (ecu::Index i = 0; i < count; ++i) EXPECT(right_pinned[i].isApprox(right[i]));
    // results the caller keeps (and pins) across calls are written in place
    std::vector<Mat> kept(count, Mat::Zero(n, n));
    const double *storage = kept[3].data();
    ecu::HostPin pin_results(kept);
    pipeline.left_multiply(fixed, tensor, kept, true);
    EXPECT(kept[3].data() == storage);
    for (ecu::Index i = 0; i < count; ++i) EXPECT(kept[i].isApprox(left_t[i]));
    pipeline.basis_transform(fixed, tensor, kept);
    for (ecu::Index i = 0; i < count; ++i) EXPECT(kept[i].isApprox(basis[i]));
  }
  ecu::CudaMatrix d_a{fixed, pipeline.get_stream()}, d_b{tensor[0], pipeline.get_stream()};
  ecu::CudaMatrix d_c{n, n, pipeline.get_stream()};
  pipeline.gemm(true, true, 1.0, d_a, d_b, 0.0, d_c);
  Mat got = d_c;
  EXPECT(got.isApprox(fixed_t * tensor[0].transpose()));
  EXPECT(ecu::count_available_gpus() >= 1);
}

int main() {
  device_round_trip();
  square_product_against_cpu();
  streamed_right_multiplication();
  inner_dimension_mismatch_throws();
  additive_api();
  if (g_failed == 0) std::printf("ALL PASSED\n");
  return g_failed == 0 ? 0 : 1;
}
