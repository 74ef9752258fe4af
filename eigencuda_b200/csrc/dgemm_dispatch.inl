// eigencuda_b200/csrc/dgemm_dispatch.inl -- the one entry every product goes through: validation, corner cases, backend and kernel choice
// Part of the single translation unit ec_api.cu (included there, in order); not a standalone header.
#pragma once

namespace {

int dgemm_dispatch(ec_pipeline *p, int transa, int transb, int64_t m, int64_t n, int64_t k,
                   double alpha, const double *A, int64_t lda, int64_t stride_a, const double *B,
                   int64_t ldb, int64_t stride_b, double beta, double *C, int64_t ldc,
                   int64_t stride_c, int64_t batch) {
  if (!p) return fail(EC_ERR_ARG, "null pipeline");
  if (m < 0 || n < 0 || k < 0 || batch < 0) return fail(EC_ERR_ARG, "negative dimension");
  if ((transa != EC_OP_N && transa != EC_OP_T) || (transb != EC_OP_N && transb != EC_OP_T))
    return fail(EC_ERR_ARG, "bad transpose flag");
  if (m == 0 || n == 0 || batch == 0) return EC_OK;
  const int64_t a_rows = transa ? k : m, b_rows = transb ? n : k;
  if (lda < std::max<int64_t>(1, a_rows) || ldb < std::max<int64_t>(1, b_rows) ||
      ldc < std::max<int64_t>(1, m))
    return fail(EC_ERR_ARG, "leading dimension too small");
  if (!C || (k > 0 && alpha != 0.0 && (!A || !B))) return fail(EC_ERR_ARG, "null operand");
  DeviceGuard guard(p->device);
  TraceRange trace("ec_dgemm");

  if (k == 0 || alpha == 0.0) {
    const int64_t total = m * n * batch;
    const int grid = (int)std::min<int64_t>((total + 255) / 256, (int64_t)p->sm_count * 8);
    ec::scale_kernel<<<grid, 256, 0, p->stream>>>(C, m, n, ldc, stride_c, batch, beta);
    EC_CUDA(cudaGetLastError());
    p->launches++;
    p->last_kernel = "scale";
    return EC_OK;
  }

  if (p->backend == EC_BACKEND_OZAKI_I8 && ozaki_eligible(m, n, k)) {
    // k beyond what the int32 accumulators hold exactly: consecutive k-chunks, accumulated with beta = 1
    for (int64_t k0 = 0; k0 < k; k0 += OZ_MAX_K) {
      const int64_t kc = std::min(OZ_MAX_K, k - k0);
      const double *Ak = A + (transa ? k0 : k0 * lda);
      const double *Bk = B + (transb ? k0 * ldb : k0);
      const int rc = ozaki_dispatch(p, transa, transb, m, n, kc, alpha, Ak, lda, stride_a, Bk, ldb, stride_b,
                                    k0 == 0 ? beta : 1.0, C, ldc, stride_c, batch);
      if (rc) return rc;
    }
    return EC_OK;
  }

  Operand oa{A, lda, stride_a, transa == EC_OP_T, m, k};
  Operand ob{B, ldb, stride_b, transb == EC_OP_N, n, k};
  // Tiny problems are latency-bound whichever kernel runs; the TMA pipeline pays off once there is
  // a small tile's worth of work.  Skinny ones (a million rows by 16 columns) count: they are
  // HBM-bound, TMA zero-fills the part of the box that hangs over the edge, and the padded DMMA
  // work is free (5.2 TB/s there against 0.7 TB/s from the SIMT kernel).
  // A batch counts as a whole: a tensor of 10^5 matrices of 32 x 32 is one HBM-bound job, not
  // 10^5 tiny ones.
  const bool big_enough = (m >= 8 && n >= 8 && k >= 8) &&
                          ((double)m * (double)n * (double)k * (double)batch >= 64.0 * 64.0 * 64.0);
  const bool c_ok = m < ((int64_t)1 << 31) && n < ((int64_t)1 << 31);
  if (big_enough && c_ok && tma_ok(oa, batch) && tma_ok(ob, batch)) {
    static const char *const nL[8] = {"tma_dmma_TN", "tma_dmma_TT", "tma_dmma_NN", "tma_dmma_NT",
                                      "tma_dmma_TN_sk", "tma_dmma_TT_sk", "tma_dmma_NN_sk", "tma_dmma_NT_sk"};
    static const char *const nS[8] = {"tma_dmma_TN_s", "tma_dmma_TT_s", "tma_dmma_NN_s", "tma_dmma_NT_s",
                                      "tma_dmma_TN_s_sk", "tma_dmma_TT_s_sk", "tma_dmma_NN_s_sk", "tma_dmma_NT_s_sk"};
    static const char *const nX[8] = {"tma_dmma_TN_xs", "tma_dmma_TT_xs", "tma_dmma_NN_xs", "tma_dmma_NT_xs",
                                      "tma_dmma_TN_xs_sk", "tma_dmma_TT_xs_sk", "tma_dmma_NN_xs_sk", "tma_dmma_NT_xs_sk"};
    static const char *const nT[8] = {"tma_dmma_TN_t", "tma_dmma_TT_t", "tma_dmma_NN_t", "tma_dmma_NT_t",
                                      "tma_dmma_TN_t_sk", "tma_dmma_TT_t_sk", "tma_dmma_NN_t_sk", "tma_dmma_NT_t_sk"};
    static const char *const nN[8] = {"tma_dmma_TN_n", "tma_dmma_TT_n", "tma_dmma_NN_n", "tma_dmma_NT_n",
                                      "tma_dmma_TN_n_sk", "tma_dmma_TT_n_sk", "tma_dmma_NN_n_sk", "tma_dmma_NT_n_sk"};
    const TileChoice ch = choose_tile(p, m, n, k, batch);
    switch (ch.tile) {
      case TILE_S: return run_tma_dmma<ec::CfgS>(p, oa, ob, m, n, k, alpha, beta, C, ldc, stride_c, stride_a, stride_b, batch, ch.sk, nS);
      case TILE_N: return run_tma_dmma<ec::CfgN>(p, oa, ob, m, n, k, alpha, beta, C, ldc, stride_c, stride_a, stride_b, batch, ch.sk, nN);
      case TILE_T: return run_tma_dmma<ec::CfgT>(p, oa, ob, m, n, k, alpha, beta, C, ldc, stride_c, stride_a, stride_b, batch, ch.sk, nT);
      case TILE_XS: return run_tma_dmma<ec::CfgXS>(p, oa, ob, m, n, k, alpha, beta, C, ldc, stride_c, stride_a, stride_b, batch, ch.sk, nX);
      default: return run_tma_dmma<ec::CfgL>(p, oa, ob, m, n, k, alpha, beta, C, ldc, stride_c, stride_a, stride_b, batch, ch.sk, nL);
    }
  }

  ec::SimtParams sp;
  sp.M = m; sp.N = n; sp.K = k; sp.batch = batch;
  sp.alpha = alpha; sp.beta = beta;
  sp.A = A; sp.lda = lda; sp.stride_a = stride_a;
  sp.B = B; sp.ldb = ldb; sp.stride_b = stride_b;
  sp.C = C; sp.ldc = ldc; sp.stride_c = stride_c;
  sp.transa = transa; sp.transb = transb;
  sp.tiles_m = (int)((m + ec::SB - 1) / ec::SB);
  sp.tiles_n = (int)((n + ec::SB - 1) / ec::SB);
  sp.total_tiles = (int64_t)sp.tiles_m * sp.tiles_n * batch;
  const int grid = (int)std::min<int64_t>(sp.total_tiles, (int64_t)p->sm_count * 4);
  ec::dgemm_simt_kernel<<<grid, ec::SIMT_THREADS, 0, p->stream>>>(sp);
  EC_CUDA(cudaGetLastError());
  p->launches++;
  p->last_kernel = "simt";
  return EC_OK;
}

}  // namespace
