// examples/readme_loop.cc -- what an EigenCuda caller's program looks like against this library,
// and what each of the three ways of driving it delivers.  Usage: readme_loop [n] [count]
//
//   (1) the loop of the reference's README (README.md:58-66), caller code unchanged:
//       copy_to_gpu -> gemm -> conversion back, pageable std::vector<Eigen::MatrixXd>, one stream;
//   (2) the same job as ONE tensor-level call (additive API): H2D, product and D2H of different
//       tensor elements overlap;
//   (3) the tensor-level call on matrices the caller keeps and has pinned in place (HostPin),
//       results written into matrices kept across calls;
//   (4) (3) over every GPU of the box -- the same call with a device list: the multi-GPU
//       scheduler shards the one vector, broadcasts A with NCCL, one worker + ring per device.
//
// Prints one JSON object with the rates in matrices/s.  Needs a B200; built by
// __graft_entry__.build() into tools/bin/readme_loop and run by bench.py's extras.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "cudamatrix.hpp"
#include "cudapipeline.hpp"

using Mat = Eigen::MatrixXd;
using Clock = std::chrono::steady_clock;

static double seconds_since(Clock::time_point t0) {
  return std::chrono::duration<double>(Clock::now() - t0).count();
}

int main(int argc, char **argv) {
  const eigencuda::Index n = argc > 1 ? std::atol(argv[1]) : 512;
  const size_t count = argc > 2 ? static_cast<size_t>(std::atol(argv[2])) : 256;
  try {
    eigencuda::CudaPipeline cuda_pip;
    const Mat A = Mat::Random(n, n);
    std::vector<Mat> tensor;
    for (size_t i = 0; i < count; ++i) tensor.push_back(Mat::Random(n, n));

    // (1) unchanged caller code
    std::vector<Mat> results(count, Mat::Zero(n, n));
    eigencuda::CudaMatrix cuma_A{A, cuda_pip.get_stream()};
    eigencuda::CudaMatrix cuma_B{n, n, cuda_pip.get_stream()};
    eigencuda::CudaMatrix cuma_C{n, n, cuda_pip.get_stream()};
    double loop_rate = 0.0;
    for (int rep = 0; rep < 3; ++rep) {   // first repetition warms up; best of the rest
      const auto t0 = Clock::now();
      for (size_t i = 0; i < count; ++i) {
        cuma_B.copy_to_gpu(tensor[i]);
        cuda_pip.gemm(cuma_B, cuma_A, cuma_C);
        results[i] = cuma_C;
      }
      const double rate = count / seconds_since(t0);
      if (rep > 0 && rate > loop_rate) loop_rate = rate;
    }
    const Mat check = tensor[count - 1] * A;
    const bool loop_ok = check.isApprox(results[count - 1]);

    // (2) one tensor-level call, pageable memory as it is
    double stream_rate = 0.0;
    std::vector<Mat> streamed;
    for (int rep = 0; rep < 3; ++rep) {
      const auto t0 = Clock::now();
      cuda_pip.right_multiply(tensor, A, streamed);
      const double rate = count / seconds_since(t0);
      if (rep > 0 && rate > stream_rate) stream_rate = rate;
    }
    const bool stream_ok = check.isApprox(streamed[count - 1]);

    // (3) the same with the caller's matrices pinned in place
    double pinned_rate = 0.0;
    {
      eigencuda::HostPin pin_in(tensor), pin_out(streamed), pin_fixed(A);
      for (int rep = 0; rep < 3; ++rep) {
        const auto t0 = Clock::now();
        cuda_pip.right_multiply(tensor, A, streamed);
        const double rate = count / seconds_since(t0);
        if (rep > 0 && rate > pinned_rate) pinned_rate = rate;
      }
    }
    bool pinned_ok = check.isApprox(streamed[count - 1]);

    // (4) the same call over every GPU of the box (and from pageable memory, as in (2))
    const int gpus = static_cast<int>(eigencuda::count_available_gpus());
    double multi_pinned_rate = 0.0, multi_pageable_rate = 0.0;
    if (gpus > 1) {
      const std::vector<int> all_devices;   // empty = every device
      for (int rep = 0; rep < 3; ++rep) {
        const auto t0 = Clock::now();
        cuda_pip.right_multiply(tensor, A, streamed, all_devices);
        const double rate = count / seconds_since(t0);
        if (rep > 0 && rate > multi_pageable_rate) multi_pageable_rate = rate;
      }
      pinned_ok = pinned_ok && check.isApprox(streamed[count - 1]);
      eigencuda::HostPin pin_in(tensor), pin_out(streamed), pin_fixed(A);
      for (int rep = 0; rep < 3; ++rep) {
        const auto t0 = Clock::now();
        cuda_pip.right_multiply(tensor, A, streamed, all_devices);
        const double rate = count / seconds_since(t0);
        if (rep > 0 && rate > multi_pinned_rate) multi_pinned_rate = rate;
      }
      pinned_ok = pinned_ok && check.isApprox(streamed[count - 1]) && check.isApprox(streamed[count - 1]);
    }

    std::printf("{\"n\": %ld, \"matrices\": %zu, \"unchanged_loop_matrices_per_s\": %.1f, "
                "\"tensor_call_pageable_matrices_per_s\": %.1f, \"tensor_call_pinned_in_place_matrices_per_s\": %.1f, "
                "\"gpus\": %d, \"multi_gpu_call_pageable_matrices_per_s\": %.1f, "
                "\"multi_gpu_call_pinned_in_place_matrices_per_s\": %.1f, \"results_ok\": %s}\n",
                static_cast<long>(n), count, loop_rate, stream_rate, pinned_rate, gpus, multi_pageable_rate,
                multi_pinned_rate,
                (loop_ok && stream_ok && pinned_ok) ? "true" : "false");
    return (loop_ok && stream_ok && pinned_ok) ? 0 : 1;
  } catch (const std::exception &e) {
    std::fprintf(stderr, "readme_loop: %s\n", e.what());
    return 2;
  }
}
