// eigencuda_b200/csrc/ozaki_dispatch.inl -- EC_BACKEND_OZAKI_I8 host side: plane workspace, split launches, chunked and pipelined GEMM launches
// Part of the single translation unit ec_api.cu (included there, in order); not a standalone header.
#pragma once

namespace {

// ------------------------------------------------------------------------------------------
// EC_BACKEND_OZAKI_I8: FP64 product emulated on the int8 tensor cores (ozaki_kernels.cuh)
// ------------------------------------------------------------------------------------------
constexpr int OZ_S = 7;   // digit planes: 49 fractional bits below each row/column maximum

// Shapes the emulation takes; everything else runs the native FP64 kernels (which are at least
// as accurate).  K <= 65536 keeps the int32 accumulators exact for 7 planes.
constexpr int64_t OZ_MAX_K = 65536;   // 7 * 65536 * 65^2 < 2^31: deeper products are cut along k
bool ozaki_eligible(int64_t m, int64_t n, int64_t k) {
  return k >= 64 && m >= 64 && n >= 32 && m * n * k >= (int64_t)128 * 128 * 128 &&
         m < ((int64_t)1 << 31) && n < ((int64_t)1 << 31);
}

template <class T>
int oz_grow(T **buf, size_t *have, size_t want) {
  if (want <= *have) return EC_OK;
  if (*buf) EC_CUDA(cudaFree(*buf));
  *buf = nullptr;
  *have = 0;
  EC_CUDA(cudaMalloc((void **)buf, want));
  *have = want;
  return EC_OK;
}

int oz_tensor_map(CUtensorMap *tm, const int8_t *planes, int64_t K, int64_t Kp, int64_t R, int64_t batch,
                  int box_rows) {
  encode_tiled_fn enc = get_encode_tiled();
  if (!enc) return fail(EC_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t dims[4] = {(cuuint64_t)K, (cuuint64_t)R, (cuuint64_t)OZ_S, (cuuint64_t)std::max<int64_t>(batch, 1)};
  cuuint64_t strides[3] = {(cuuint64_t)Kp, (cuuint64_t)(R * Kp), (cuuint64_t)(OZ_S * R * Kp)};
  cuuint32_t box[4] = {(cuuint32_t)ec::oz::OBK, (cuuint32_t)box_rows, (cuuint32_t)OZ_S, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, (void *)planes, dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(EC_ERR_CUDA, "cuTensorMapEncodeTiled (int8 planes) failed with CUresult %d", (int)r);
  return EC_OK;
}

// Scale + split one operand: op(X) viewed as R rows x K, element (r, q) at X[r*rs + q*ks].
// `amax` is scratch for R*batch abs-max bit patterns.
int oz_split(ec_pipeline *p, const double *X, int64_t rs, int64_t ks, int64_t bstride, int64_t R, int64_t K,
             int64_t Kp, int64_t batch, int8_t *planes, double *scale, unsigned long long *amax) {
  ec::oz::SplitParams sp{X, rs, ks, bstride, R, K, Kp, batch, planes, scale};
  const bool row_contig = (rs == 1 && ks != 1);
  EC_CUDA(cudaMemsetAsync(amax, 0, (size_t)(R * batch) * 8, p->stream));
  const int64_t rt = (R + ec::oz::SPLIT_ROWS - 1) / ec::oz::SPLIT_ROWS;
  const int64_t tiles_k = ((K + ec::oz::SPLIT_K - 1) / ec::oz::SPLIT_K) * rt * batch;
  const int64_t tiles_kp = ((Kp + ec::oz::SPLIT_K - 1) / ec::oz::SPLIT_K) * rt * batch;
  const int g1 = (int)std::min<int64_t>(tiles_k, (int64_t)p->sm_count * 8);
  const int g2 = (int)std::min<int64_t>(tiles_kp, (int64_t)p->sm_count * 8);
  if (row_contig) {
    ec::oz::ozaki_absmax_kernel<true><<<g1, ec::oz::SPLIT_THREADS, 0, p->stream>>>(sp, amax);
    ec::oz::ozaki_split_kernel<OZ_S, true><<<g2, ec::oz::SPLIT_THREADS, 0, p->stream>>>(sp, amax);
  } else {
    ec::oz::ozaki_absmax_kernel<false><<<g1, ec::oz::SPLIT_THREADS, 0, p->stream>>>(sp, amax);
    ec::oz::ozaki_split_kernel<OZ_S, false><<<g2, ec::oz::SPLIT_THREADS, 0, p->stream>>>(sp, amax);
  }
  EC_CUDA(cudaGetLastError());
  p->launches += 2;
  return EC_OK;
}

int ozaki_dispatch(ec_pipeline *p, int transa, int transb, int64_t m, int64_t n, int64_t k, double alpha,
                   const double *A, int64_t lda, int64_t stride_a, const double *B, int64_t ldb, int64_t stride_b,
                   double beta, double *C, int64_t ldc, int64_t stride_c, int64_t batch) {
  using Cfg = ec::oz::OzCfg<OZ_S>;
  const int64_t Kp = (k + 15) / 16 * 16;
  const bool a_batched = batch > 1 && stride_a != 0, b_batched = batch > 1 && stride_b != 0;
  // chunk the batch so that the digit planes of a chunk stay around L2 size; the batched
  // operand(s) are double-buffered so that the split of chunk c+1 (LSU / FP64 pipe / HBM) runs on
  // a side stream underneath the GEMM of chunk c (tensor cores / L2)
  const int64_t per_problem = OZ_S * Kp * ((a_batched ? m : 0) + (b_batched ? n : 0));
  int64_t chunk = batch;
  if (per_problem > 0) chunk = std::max<int64_t>(1, std::min<int64_t>(batch, ((int64_t)192 << 20) / per_problem));
  if (const char *e = getenv("EC_OZAKI_CHUNK")) chunk = std::max<int64_t>(1, std::min<int64_t>(batch, atoll(e)));
  const int64_t nchunks = (batch + chunk - 1) / chunk;
  const bool pipelined = nchunks > 1 && (a_batched || b_batched);
  const int nbuf = pipelined ? 2 : 1;
  const int64_t ca = a_batched ? chunk : 1, cb = b_batched ? chunk : 1;
  const size_t pa = (size_t)(OZ_S * m * Kp * ca), pb = (size_t)(OZ_S * n * Kp * cb);
  const size_t sa_b = (size_t)(m * ca * 8), sb_b = (size_t)(n * cb * 8), am_b = (size_t)(std::max(m * ca, n * cb) * 8);
  int rc;
  if ((rc = oz_grow(&p->oz_planes_a, &p->oz_planes_a_bytes, pa * (a_batched ? nbuf : 1)))) return rc;
  if ((rc = oz_grow(&p->oz_planes_b, &p->oz_planes_b_bytes, pb * (b_batched ? nbuf : 1)))) return rc;
  if ((rc = oz_grow(&p->oz_scale_a, &p->oz_scale_a_bytes, sa_b * (a_batched ? nbuf : 1)))) return rc;
  if ((rc = oz_grow(&p->oz_scale_b, &p->oz_scale_b_bytes, sb_b * (b_batched ? nbuf : 1)))) return rc;
  if ((rc = oz_grow(&p->oz_amax, &p->oz_amax_bytes, am_b * 3))) return rc;   // [shared operand, buffer 0, buffer 1]
  if (pipelined && !p->oz_split_stream) {
    EC_CUDA(cudaStreamCreateWithFlags(&p->oz_split_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
      EC_CUDA(cudaEventCreateWithFlags(&p->oz_ev_split[i], cudaEventDisableTiming));
      EC_CUDA(cudaEventCreateWithFlags(&p->oz_ev_gemm[i], cudaEventDisableTiming));
    }
  }
  auto kern = ec::oz::ozaki_gemm_kernel<OZ_S>;
  static std::atomic<uint64_t> configured{0};
  const uint64_t bit = 1ull << (p->device & 63);
  if (!(configured.load() & bit)) {
    EC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    configured.fetch_or(bit);
  }
  // op(A)(r, q): N -> A[r + q*lda]; T -> A[r*lda + q].  op(B)^T(j, q) = op(B)(q, j): N -> B[j*ldb + q]; T -> B[q*ldb + j]
  const int64_t a_rs = transa ? lda : 1, a_ks = transa ? 1 : lda;
  const int64_t b_rs = transb ? 1 : ldb, b_ks = transb ? ldb : 1;
  unsigned long long *amax_shared = p->oz_amax, *amax_buf[2] = {p->oz_amax + am_b / 8, p->oz_amax + 2 * (am_b / 8)};
  if (!a_batched && (rc = oz_split(p, A, a_rs, a_ks, 0, m, k, Kp, 1, p->oz_planes_a, p->oz_scale_a, amax_shared))) return rc;
  if (!b_batched && (rc = oz_split(p, B, b_rs, b_ks, 0, n, k, Kp, 1, p->oz_planes_b, p->oz_scale_b, amax_shared))) return rc;

  cudaStream_t main_stream = p->stream;
  // split of chunk c into buffer c & 1, on the side stream when pipelined
  auto split_chunk = [&](int64_t c) -> int {
    const int64_t b0 = c * chunk, nb = std::min(chunk, batch - b0);
    const int buf = pipelined ? (int)(c & 1) : 0;
    StreamScope scope(p, pipelined ? p->oz_split_stream : main_stream);
    if (pipelined && c >= 2) EC_CUDA(cudaStreamWaitEvent(p->oz_split_stream, p->oz_ev_gemm[buf], 0));   // buffer consumed
    int r = EC_OK;
    if (a_batched)
      r = oz_split(p, A + b0 * stride_a, a_rs, a_ks, stride_a, m, k, Kp, nb, p->oz_planes_a + buf * pa,
                   p->oz_scale_a + buf * (sa_b / 8), amax_buf[buf]);
    if (!r && b_batched)
      r = oz_split(p, B + b0 * stride_b, b_rs, b_ks, stride_b, n, k, Kp, nb, p->oz_planes_b + buf * pb,
                   p->oz_scale_b + buf * (sb_b / 8), amax_buf[buf]);
    if (!r && pipelined) EC_CUDA(cudaEventRecord(p->oz_ev_split[buf], p->oz_split_stream));
    return r;
  };
  if (pipelined) {
    // the side stream starts after whatever produced the operands on the main stream
    EC_CUDA(cudaEventRecord(p->oz_ev_gemm[0], main_stream));
    EC_CUDA(cudaStreamWaitEvent(p->oz_split_stream, p->oz_ev_gemm[0], 0));
  }
  if (a_batched || b_batched) {
    if ((rc = split_chunk(0))) return rc;
  }
  for (int64_t c = 0; c < nchunks; ++c) {
    const int64_t b0 = c * chunk, nb = std::min(chunk, batch - b0);
    const int buf = pipelined ? (int)(c & 1) : 0;
    if (pipelined && c + 1 < nchunks && (rc = split_chunk(c + 1))) return rc;
    if (!pipelined && c > 0 && (a_batched || b_batched) && (rc = split_chunk(c))) return rc;
    if (pipelined) EC_CUDA(cudaStreamWaitEvent(main_stream, p->oz_ev_split[buf], 0));
    CUtensorMap tmA, tmB;
    if ((rc = oz_tensor_map(&tmA, p->oz_planes_a + (a_batched ? buf * pa : 0), k, Kp, m, a_batched ? nb : 1, ec::oz::OM))) return rc;
    if ((rc = oz_tensor_map(&tmB, p->oz_planes_b + (b_batched ? buf * pb : 0), k, Kp, n, b_batched ? nb : 1, ec::oz::ON))) return rc;
    ec::oz::GemmOzParams gp;
    gp.M = m; gp.N = n; gp.K = k; gp.batch = nb;
    gp.alpha = alpha; gp.beta = beta;
    gp.C = C + b0 * stride_c; gp.ldc = ldc; gp.stride_c = stride_c;
    gp.sa = p->oz_scale_a + (a_batched ? buf * (sa_b / 8) : 0);
    gp.sb = p->oz_scale_b + (b_batched ? buf * (sb_b / 8) : 0);
    gp.a_batched = a_batched; gp.b_batched = b_batched;
    gp.tiles_m = (int)((m + ec::oz::OM - 1) / ec::oz::OM);
    gp.tiles_n = (int)((n + ec::oz::ON - 1) / ec::oz::ON);
    gp.total_tiles = (int64_t)gp.tiles_m * gp.tiles_n * nb;
    const int grid = (int)std::min<int64_t>(gp.total_tiles, p->sm_count - p->sm_margin);
    kern<<<grid, ec::oz::OTHREADS, Cfg::SMEM_BYTES, main_stream>>>(tmA, tmB, gp);
    EC_CUDA(cudaGetLastError());
    p->launches++;
    if (pipelined) EC_CUDA(cudaEventRecord(p->oz_ev_gemm[buf], main_stream));
  }
  p->last_kernel = "ozaki_i8_s7";
  return EC_OK;
}


}  // namespace
