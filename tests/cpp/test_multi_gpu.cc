// tests/cpp/test_multi_gpu.cc -- the multi-GPU tensor scheduler from caller-side C++: the loop of
// the reference's README (README.md:58-66) over ONE std::vector<Eigen::MatrixXd>, handed to
// CudaPipeline::right_multiply / left_multiply / basis_transform together with a device list.
// Checks, for every device count the box offers (1, 2, 4, ... up to count_available_gpus()):
// results equal the CPU product (Eigen isApprox, 1e-12) AND are bit-identical to the single-GPU
// call (same kernels, same tiles: sharding must not change a single bit); ragged shards (count not
// a multiple of the device count, fewer matrices than devices); pinned and pageable memory; the
// reference's known answers; the shape-mismatch throw.  Needs a B200: run by tests/test_gpu_parity.py.
#include <cstdio>
#include <cstring>
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

static bool same_bits(const Mat &a, const Mat &b) {
  return a.rows() == b.rows() && a.cols() == b.cols() &&
         std::memcmp(a.data(), b.data(), sizeof(double) * static_cast<size_t>(a.size())) == 0;
}

static Mat from_rows(ecu::Index rows, ecu::Index cols, std::initializer_list<double> values) {
  Mat m = Mat::Zero(rows, cols);
  ecu::Index n = 0;
  for (double v : values) {
    m(n / cols, n % cols) = v;
    ++n;
  }
  return m;
}

static void known_answers(const std::vector<int> &devices) {
  // tests/golden/readme_example.json <- reference src/tests/test_dot.cc:58-74,86-89
  const Mat fixed = from_rows(2, 2, {1, 2, 3, 4});
  const std::vector<Mat> tensor = {from_rows(3, 2, {5, 6, 7, 8, 9, 10}), from_rows(3, 2, {9, 10, 11, 12, 13, 14}),
                                   from_rows(3, 2, {13, 14, 15, 16, 17, 18})};
  const std::vector<Mat> expected = {from_rows(3, 2, {23, 34, 31, 46, 39, 58}), from_rows(3, 2, {39, 58, 47, 70, 55, 82}),
                                     from_rows(3, 2, {55, 82, 63, 94, 71, 106})};
  ecu::CudaPipeline pipeline;
  std::vector<Mat> got = pipeline.right_multiply(tensor, fixed, devices);
  EXPECT(got.size() == 3);
  for (size_t i = 0; i < got.size(); ++i) EXPECT(expected[i].isApprox(got[i]));
}

static void sharded_equals_single(const std::vector<int> &devices) {
  ecu::CudaPipeline pipeline;
  const ecu::Index n = 160;
  const size_t count = 37;   // ragged: not a multiple of 2, 4 or 8
  const Mat fixed = Mat::Random(n, n), fixed_t = fixed.transpose();
  std::vector<Mat> tensor;
  for (size_t i = 0; i < count; ++i) tensor.push_back(Mat::Random(n, n));
  const std::vector<Mat> single_r = pipeline.right_multiply(tensor, fixed);
  const std::vector<Mat> single_l = pipeline.left_multiply(fixed, tensor, true);
  const std::vector<Mat> single_b = pipeline.basis_transform(fixed, tensor);
  const std::vector<Mat> multi_r = pipeline.right_multiply(tensor, fixed, devices);
  const std::vector<Mat> multi_l = pipeline.left_multiply(fixed, tensor, devices, true);
  const std::vector<Mat> multi_b = pipeline.basis_transform(fixed, tensor, devices);
  EXPECT(multi_r.size() == count && multi_l.size() == count && multi_b.size() == count);
  for (size_t i = 0; i < count; ++i) {
    EXPECT(multi_r[i].isApprox(tensor[i] * fixed));
    EXPECT(multi_l[i].isApprox(fixed_t * tensor[i]));
    EXPECT(multi_b[i].isApprox((fixed_t * tensor[i]) * fixed));
    EXPECT(same_bits(multi_r[i], single_r[i]));
    EXPECT(same_bits(multi_l[i], single_l[i]));
    EXPECT(same_bits(multi_b[i], single_b[i]));
  }
  // the caller's matrices pinned in place, results kept across calls: written where they are
  {
    std::vector<Mat> kept(count, Mat::Zero(n, n));
    ecu::HostPin pin_in(tensor), pin_out(kept), pin_fixed(fixed);
    const double *storage = kept[count / 2].data();
    pipeline.right_multiply(tensor, fixed, kept, devices);
    EXPECT(kept[count / 2].data() == storage);
    for (size_t i = 0; i < count; ++i) EXPECT(same_bits(kept[i], single_r[i]));
  }
  // fewer matrices than devices: some shards are empty
  {
    std::vector<Mat> few(tensor.begin(), tensor.begin() + 1);
    const std::vector<Mat> got = pipeline.right_multiply(few, fixed, devices);
    EXPECT(got.size() == 1 && same_bits(got[0], single_r[0]));
  }
  // empty tensor
  EXPECT(pipeline.right_multiply(std::vector<Mat>(), fixed, devices).empty());
}

static void mismatch_throws(const std::vector<int> &devices) {
  ecu::CudaPipeline pipeline;
  bool thrown = false;
  try {
    pipeline.right_multiply(std::vector<Mat>(3, Mat::Random(3, 2)), Mat::Random(5, 5), devices);
  } catch (const std::runtime_error &) {
    thrown = true;
  }
  EXPECT(thrown);
  thrown = false;
  try {
    pipeline.right_multiply(std::vector<Mat>(3, Mat::Random(4, 4)), Mat::Random(4, 4), std::vector<int>{0, 0});
  } catch (const std::runtime_error &) {   // a device listed twice
    thrown = true;
  }
  EXPECT(thrown);
}

int main() {
  const int available = static_cast<int>(ecu::count_available_gpus());
  if (available < 1) {
    std::printf("no CUDA device\n");
    return 2;
  }
  int tested = 0;
  for (int g = 1; g <= available; g *= 2) {
    std::vector<int> devices;
    for (int d = 0; d < g; ++d) devices.push_back(d);
    known_answers(devices);
    sharded_equals_single(devices);
    mismatch_throws(devices);
    ++tested;
  }
  {   // an empty list means every device of the box
    known_answers(std::vector<int>());
  }
  {   // the scheduler's own pick of devices (by host-link measurement)
    const std::vector<int> picked = ecu::CudaPipeline::auto_devices();
    EXPECT(!picked.empty() && static_cast<int>(picked.size()) <= available);
    known_answers(picked);
  }
  if (available >= 2) {   // the pipeline's own device need not lead the list
    std::vector<int> devices = {1, 0};
    sharded_equals_single(devices);
  }
  std::printf("device counts tested: %d (of %d devices)\n", tested, available);
  if (g_failed == 0) std::printf("ALL PASSED\n");
  return g_failed == 0 ? 0 : 1;
}
