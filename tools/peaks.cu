// tools/peaks.cu -- measures the roofline denominators MEASURED_PEAKS.json does not carry:
//   * FP64 DMMA (mma.sync.m8n8k4.f64) and FP64 FMA issue-rate peaks, registers only;
//   * pinned host<->device copy bandwidth (the tensor-stream job's bounding link), each
//     direction alone and both at once;
//   * cuBLAS DGEMM (what the reference's gemm runs, reference src/cudapipeline.cc:29-31) at a few
//     sizes, as the practical library ceiling on the same box.
// Measurement tool only: not linked into the product library.  Prints one JSON object.
#include <cublas_v2.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e = (x);                                                               \
    if (e != cudaSuccess) {                                                            \
      fprintf(stderr, "%s: %s (%s:%d)\n", #x, cudaGetErrorString(e), __FILE__, __LINE__); \
      exit(1);                                                                         \
    }                                                                                  \
  } while (0)

template <int ILP>
__global__ void dmma_peak(double *out, int iters, double seed) {
  double a = seed + threadIdx.x, b = seed * 0.5 + threadIdx.x;
  double c0[ILP], c1[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) c0[i] = c1[i] = 0.0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c0[i]), "+d"(c1[i])
                   : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += c0[i] + c1[i];
  if (s == 123.456) out[0] = s;
}

template <int ILP>
__global__ void dfma_peak(double *out, int iters, double seed) {
  double a = seed + threadIdx.x * 1e-9, b = 1.0 - 1e-9;
  double c[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) c[i] = i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) c[i] = fma(c[i], b, a);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += c[i];
  if (s == 123.456) out[0] = s;
}

static float time_ms(cudaEvent_t e0, cudaEvent_t e1) {
  float t;
  CK(cudaEventElapsedTime(&t, e0, e1));
  return t;
}

int main(int argc, char **argv) {
  bool quick = argc > 1 && !strcmp(argv[1], "--quick");
  int dev = 0;
  CK(cudaSetDevice(dev));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, dev));
  int sms = prop.multiProcessorCount;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  double *dout;
  CK(cudaMalloc(&dout, 64));
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d", prop.name, sms, prop.clockRate);

  // ---- DMMA / DFMA issue peaks: warps per SM sweep ----
  {
    const int iters = quick ? 4000 : 20000;
    double best_dmma = 0, best_dfma = 0;
    int best_w_dmma = 0, best_w_dfma = 0;
    printf(", \"dmma_sweep\": [");
    bool first = true;
    for (int warps : {4, 8, 16, 32}) {
      for (int rep = 0; rep < 3; ++rep) {
        dmma_peak<8><<<sms, warps * 32>>>(dout, iters, 1.0);
      }
      CK(cudaDeviceSynchronize());
      float best = 1e30f;
      for (int rep = 0; rep < 5; ++rep) {
        CK(cudaEventRecord(e0));
        dmma_peak<8><<<sms, warps * 32>>>(dout, iters, 1.0);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        best = std::min(best, time_ms(e0, e1));
      }
      double tf = 512.0 * 8 * iters * warps * sms / (best * 1e-3) / 1e12;
      printf("%s{\"warps_per_sm\": %d, \"tflops\": %.2f}", first ? "" : ", ", warps, tf);
      first = false;
      if (tf > best_dmma) { best_dmma = tf; best_w_dmma = warps; }
    }
    printf("], \"dfma_sweep\": [");
    first = true;
    for (int warps : {8, 16, 32}) {
      for (int rep = 0; rep < 3; ++rep) dfma_peak<8><<<sms, warps * 32>>>(dout, iters * 4, 1.0);
      CK(cudaDeviceSynchronize());
      float best = 1e30f;
      for (int rep = 0; rep < 5; ++rep) {
        CK(cudaEventRecord(e0));
        dfma_peak<8><<<sms, warps * 32>>>(dout, iters * 4, 1.0);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        best = std::min(best, time_ms(e0, e1));
      }
      double tf = 2.0 * 32 * 8 * (iters * 4.0) * warps * sms / (best * 1e-3) / 1e12;
      printf("%s{\"warps_per_sm\": %d, \"tflops\": %.2f}", first ? "" : ", ", warps, tf);
      first = false;
      if (tf > best_dfma) { best_dfma = tf; best_w_dfma = warps; }
    }
    printf("], \"dmma_peak_tflops\": %.2f, \"dmma_peak_warps\": %d, \"dfma_peak_tflops\": %.2f, \"dfma_peak_warps\": %d",
           best_dmma, best_w_dmma, best_dfma, best_w_dfma);
    // sustained: ~2 s of back-to-back DMMA at the best configuration (power-capped clocks)
    {
      const int launches = quick ? 20 : 100;
      CK(cudaEventRecord(e0));
      for (int i = 0; i < launches; ++i) dmma_peak<8><<<sms, best_w_dmma * 32>>>(dout, iters * 4, 1.0);
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      double tf = 512.0 * 8 * (iters * 4.0) * best_w_dmma * sms * launches / (time_ms(e0, e1) * 1e-3) / 1e12;
      printf(", \"dmma_sustained_tflops\": %.2f, \"dmma_sustained_seconds\": %.2f", tf, time_ms(e0, e1) * 1e-3);
    }
  }

  // ---- host link ----
  {
    const size_t bytes = (size_t)256 << 20;
    void *h_in, *h_out, *d_a, *d_b;
    CK(cudaHostAlloc(&h_in, bytes, cudaHostAllocDefault));
    CK(cudaHostAlloc(&h_out, bytes, cudaHostAllocDefault));
    memset(h_in, 1, bytes);
    memset(h_out, 2, bytes);
    CK(cudaMalloc(&d_a, bytes));
    CK(cudaMalloc(&d_b, bytes));
    cudaStream_t s1, s2;
    CK(cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
    cudaEvent_t a0, a1, b0, b1;
    CK(cudaEventCreate(&a0)); CK(cudaEventCreate(&a1)); CK(cudaEventCreate(&b0)); CK(cudaEventCreate(&b1));
    float h2d = 1e30f, d2h = 1e30f, both = 1e30f;
    for (int rep = 0; rep < 10; ++rep) {
      CK(cudaEventRecord(a0, s1));
      CK(cudaMemcpyAsync(d_a, h_in, bytes, cudaMemcpyHostToDevice, s1));
      CK(cudaEventRecord(a1, s1));
      CK(cudaEventSynchronize(a1));
      h2d = std::min(h2d, time_ms(a0, a1));
      CK(cudaEventRecord(b0, s2));
      CK(cudaMemcpyAsync(h_out, d_b, bytes, cudaMemcpyDeviceToHost, s2));
      CK(cudaEventRecord(b1, s2));
      CK(cudaEventSynchronize(b1));
      d2h = std::min(d2h, time_ms(b0, b1));
      // both directions at once: wall time until both are done
      CK(cudaDeviceSynchronize());
      CK(cudaEventRecord(a0, s1));
      CK(cudaStreamWaitEvent(s2, a0, 0));
      CK(cudaMemcpyAsync(d_a, h_in, bytes, cudaMemcpyHostToDevice, s1));
      CK(cudaMemcpyAsync(h_out, d_b, bytes, cudaMemcpyDeviceToHost, s2));
      CK(cudaEventRecord(b1, s2));
      CK(cudaStreamWaitEvent(s1, b1, 0));
      CK(cudaEventRecord(a1, s1));
      CK(cudaEventSynchronize(a1));
      both = std::min(both, time_ms(a0, a1));
    }
    printf(", \"h2d_pinned_gbs\": %.2f, \"d2h_pinned_gbs\": %.2f, \"bidir_each_gbs\": %.2f",
           bytes / (h2d * 1e-3) / 1e9, bytes / (d2h * 1e-3) / 1e9, bytes / (both * 1e-3) / 1e9);
    // pageable, as the reference does it (src/cudamatrix.cc:26-27,41-42,49-50): 8 MB per call
    {
      const size_t pb = (size_t)8 << 20;
      char *pg = (char *)malloc(pb);
      memset(pg, 3, pb);
      float best_in = 1e30f, best_out = 1e30f;
      for (int rep = 0; rep < 10; ++rep) {
        CK(cudaEventRecord(a0, s1));
        CK(cudaMemcpyAsync(d_a, pg, pb, cudaMemcpyHostToDevice, s1));
        CK(cudaEventRecord(a1, s1));
        CK(cudaEventSynchronize(a1));
        best_in = std::min(best_in, time_ms(a0, a1));
        CK(cudaEventRecord(a0, s1));
        CK(cudaMemcpyAsync(pg, d_a, pb, cudaMemcpyDeviceToHost, s1));
        CK(cudaEventRecord(a1, s1));
        CK(cudaEventSynchronize(a1));
        best_out = std::min(best_out, time_ms(a0, a1));
      }
      printf(", \"h2d_pageable_8mb_gbs\": %.2f, \"d2h_pageable_8mb_gbs\": %.2f",
             pb / (best_in * 1e-3) / 1e9, pb / (best_out * 1e-3) / 1e9);
      free(pg);
    }
    // device copy (read + write bytes), for reference against MEASURED_PEAKS.json hbm_gbs
    float dd = 1e30f;
    for (int rep = 0; rep < 10; ++rep) {
      CK(cudaEventRecord(e0));
      CK(cudaMemcpyAsync(d_b, d_a, bytes, cudaMemcpyDeviceToDevice, 0));
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      dd = std::min(dd, time_ms(e0, e1));
    }
    printf(", \"d2d_copy_gbs\": %.1f", 2.0 * bytes / (dd * 1e-3) / 1e9);
    cudaFreeHost(h_in); cudaFreeHost(h_out); cudaFree(d_a); cudaFree(d_b);
  }

  // ---- cuBLAS DGEMM ----
  {
    cublasHandle_t h;
    cublasCreate(&h);
    printf(", \"cublas_dgemm\": [");
    bool first = true;
    std::vector<int> sizes = quick ? std::vector<int>{512, 4096, 8192} : std::vector<int>{256, 512, 1024, 2048, 4096, 8192, 16384};
    for (int n : sizes) {
      double *A, *B, *Cm;
      size_t bytes = (size_t)n * n * 8;
      CK(cudaMalloc(&A, bytes)); CK(cudaMalloc(&B, bytes)); CK(cudaMalloc(&Cm, bytes));
      CK(cudaMemset(A, 0x3c, bytes)); CK(cudaMemset(B, 0x3d, bytes));
      const double one = 1.0, zero = 0.0;
      int reps = n >= 8192 ? 3 : 10;
      for (int i = 0; i < 3; ++i)
        cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, n, n, n, &one, A, n, B, n, &zero, Cm, n);
      CK(cudaDeviceSynchronize());
      float best = 1e30f;
      for (int rep = 0; rep < reps; ++rep) {
        CK(cudaEventRecord(e0));
        cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, n, n, n, &one, A, n, B, n, &zero, Cm, n);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        best = std::min(best, time_ms(e0, e1));
      }
      printf("%s{\"n\": %d, \"ms\": %.4f, \"tflops\": %.2f}", first ? "" : ", ", n, best,
             2.0 * n * n * (double)n / (best * 1e-3) / 1e12);
      first = false;
      cudaFree(A); cudaFree(B); cudaFree(Cm);
    }
    printf("]");
    // batched 512^3 x 256 (the tensor-stream shape) through cublasDgemmStridedBatched
    {
      const int n = 512, batch = 256;
      double *A, *B, *Cm;
      CK(cudaMalloc(&A, (size_t)n * n * 8 * batch)); CK(cudaMalloc(&B, (size_t)n * n * 8));
      CK(cudaMalloc(&Cm, (size_t)n * n * 8 * batch));
      CK(cudaMemset(A, 0x3c, (size_t)n * n * 8 * batch)); CK(cudaMemset(B, 0x3d, (size_t)n * n * 8));
      const double one = 1.0, zero = 0.0;
      for (int i = 0; i < 3; ++i)
        cublasDgemmStridedBatched(h, CUBLAS_OP_N, CUBLAS_OP_N, n, n, n, &one, A, n, (long long)n * n, B, n, 0, &zero, Cm, n, (long long)n * n, batch);
      CK(cudaDeviceSynchronize());
      float best = 1e30f;
      for (int rep = 0; rep < 5; ++rep) {
        CK(cudaEventRecord(e0));
        cublasDgemmStridedBatched(h, CUBLAS_OP_N, CUBLAS_OP_N, n, n, n, &one, A, n, (long long)n * n, B, n, 0, &zero, Cm, n, (long long)n * n, batch);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        best = std::min(best, time_ms(e0, e1));
      }
      printf(", \"cublas_batched_512x256\": {\"ms\": %.4f, \"tflops\": %.2f, \"matrices_per_s\": %.0f}", best,
             2.0 * n * n * (double)n * batch / (best * 1e-3) / 1e12, batch / (best * 1e-3));
      cudaFree(A); cudaFree(B); cudaFree(Cm);
    }
    cublasDestroy(h);
  }
  printf("}\n");
  return 0;
}
