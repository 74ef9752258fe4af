// eigencuda_b200/csrc/dgemm_2d.inl -- one very large DGEMM, 2D block-partitioned over the GPUs of one box (SUMMA with NCCL panel broadcasts)
// Part of the single translation unit ec_api.cu (included there, in order); not a standalone header.
//
// BASELINE.json configs[4]: C = A * B at n = 65536 on 8 B200.  The
// reference cannot express this at all: one device, dimensions narrowed to int
// (src/cudapipeline.cc:29-31, Appendix A #4).
//
// Layout, everything column-major.  P x Q device grid, device (pr, pc) = devices[pr * Q + pc].
// A matrix of rows x cols is cut into P row blocks x Q column blocks, block (pr, pc) lives on
// device (pr, pc) with leading dimension = its row count.  For C(m x n) = A(m x k) * B(k x n), per
// k-panel [k0, k1) (panels never straddle an ownership boundary of A's columns or B's rows):
//     the grid column that owns those columns of A broadcasts A[pr, :][:, panel] (mloc x w,
//     contiguous in the owner's block) along every grid row;
//     the grid row that owns those rows of B broadcasts B[:, pc][panel, :] (w x nloc, packed by a
//     copy-engine 2D copy on the owner: no kernel) along every grid column;
//     every device then does C[pr, pc] += Apanel * Bpanel with this library's DGEMM kernel.
// The broadcasts of panel t+1 run on a side stream per device, double-buffered against the product
// of panel t.  One host thread drives all devices (everything is asynchronous); NCCL calls of one
// panel are grouped.  The NCCL communicators are created with a cap on the CTAs their kernels may
// use (EC_2D_NCCL_CTAS, default 8): the panels need ~10 GB/s per GPU, a fraction of what one
// channel moves over NVLink, and every SM NCCL holds is an SM the persistent GEMM grid cannot use.
//
// Why not a kernel that TMA-loads the panels straight from the peers: peer addresses bypass the
// local L2, so every CTA's re-read of a panel tile would cross NVLink (about 1.2 TB/s for this
// kernel against 0.77 TB/s available), whereas a broadcast into local HBM costs each byte one
// crossing (17 GB per GPU at n = 65536, ~25 ms against ~1.9 s of DMMA).
#pragma once

struct ec_grid {
  int P = 1, Q = 1;
  std::vector<int> devices;                 // P*Q, row-major over (pr, pc)
  std::vector<ec_pipeline *> pipes;         // one child pipeline per device (its stream = the compute stream)
  std::vector<cudaStream_t> comm;           // side stream per device for the panel traffic
  std::vector<ncclComm_t> row_comm;         // per device: communicator of its grid row (Q ranks), null when Q == 1
  std::vector<ncclComm_t> col_comm;         // per device: communicator of its grid column (P ranks), null when P == 1
  std::vector<double *> bufA[2], bufB[2];   // panel double buffers per device
  std::vector<size_t> bufA_bytes, bufB_bytes;
  std::vector<cudaEvent_t> ready[2], freed[2], t0, t1;
  int nccl_ctas = 8;
  int64_t broadcasts = 0;

  int dev(int pr, int pc) const { return pr * Q + pc; }
};

struct ec_dist_matrix {
  ec_grid *grid = nullptr;
  int64_t rows = 0, cols = 0;
  std::vector<double *> block;              // per device
  std::vector<int> devices;                 // copy of the grid's list: a matrix may be released after its grid
};

namespace {

inline void block_range(int64_t total, int parts, int idx, int64_t *lo, int64_t *hi) {
  *lo = total * idx / parts;
  *hi = total * (idx + 1) / parts;
}

void grid_shape_auto(int ndev, int *P, int *Q) {
  int p = 1;
  for (int c = 1; c * c <= ndev; ++c)
    if (ndev % c == 0) p = c;
  *P = p;             // P <= Q, as square as possible: 1x1, 1x2, 2x2, 2x4
  *Q = ndev / p;
}

int grid_destroy_impl(ec_grid *g) {
  if (!g) return EC_OK;
  NcclApi *api = nccl_api();
  for (size_t i = 0; i < g->devices.size(); ++i) {
    DeviceGuard guard(g->devices[i]);
    if (i < g->pipes.size() && g->pipes[i]) cudaStreamSynchronize(g->pipes[i]->stream);
    if (i < g->comm.size() && g->comm[i]) cudaStreamSynchronize(g->comm[i]);
  }
  for (size_t i = 0; i < g->devices.size(); ++i) {
    DeviceGuard guard(g->devices[i]);
    if (api->handle) {
      if (i < g->row_comm.size() && g->row_comm[i]) api->CommDestroy(g->row_comm[i]);
      if (i < g->col_comm.size() && g->col_comm[i]) api->CommDestroy(g->col_comm[i]);
    }
    for (int b = 0; b < 2; ++b) {
      if (i < g->bufA[b].size() && g->bufA[b][i]) cudaFree(g->bufA[b][i]);
      if (i < g->bufB[b].size() && g->bufB[b][i]) cudaFree(g->bufB[b][i]);
      if (i < g->ready[b].size() && g->ready[b][i]) cudaEventDestroy(g->ready[b][i]);
      if (i < g->freed[b].size() && g->freed[b][i]) cudaEventDestroy(g->freed[b][i]);
    }
    if (i < g->t0.size() && g->t0[i]) cudaEventDestroy(g->t0[i]);
    if (i < g->t1.size() && g->t1[i]) cudaEventDestroy(g->t1[i]);
    if (i < g->comm.size() && g->comm[i]) cudaStreamDestroy(g->comm[i]);
    if (i < g->pipes.size() && g->pipes[i]) ec_pipeline_destroy(g->pipes[i]);
  }
  delete g;
  return EC_OK;
}

int grid_create_impl(const int *devices, int ndev, int P, int Q, ec_grid **out) {
  if (!out) return fail(EC_ERR_ARG, "null out pointer");
  *out = nullptr;
  if (ndev < 0 || (ndev > 0 && !devices)) return fail(EC_ERR_ARG, "bad device list");
  std::vector<int> devs;
  if (ndev == 0)
    for (int d = 0; d < (int)ec_device_count(); ++d) devs.push_back(d);
  else
    devs.assign(devices, devices + ndev);
  const int n = (int)devs.size();
  if (n == 0) return fail(EC_ERR_CUDA, "no CUDA device visible");
  if (P <= 0 || Q <= 0) grid_shape_auto(n, &P, &Q);
  if (P * Q != n) return fail(EC_ERR_ARG, "grid %d x %d does not match %d devices", P, Q, n);
  int count = 0;
  EC_CUDA(cudaGetDeviceCount(&count));
  for (int i = 0; i < n; ++i) {
    if (devs[i] < 0 || devs[i] >= count) return fail(EC_ERR_ARG, "device %d out of range (%d devices)", devs[i], count);
    for (int j = 0; j < i; ++j)
      if (devs[j] == devs[i]) return fail(EC_ERR_ARG, "device %d listed twice", devs[i]);
  }
  ec_grid *g = new ec_grid();
  g->P = P;
  g->Q = Q;
  g->devices = devs;
  if (const char *e = std::getenv("EC_2D_NCCL_CTAS")) g->nccl_ctas = std::max(0, std::atoi(e));
  g->pipes.assign(n, nullptr);
  g->comm.assign(n, nullptr);
  g->row_comm.assign(n, nullptr);
  g->col_comm.assign(n, nullptr);
  g->bufA_bytes.assign(n, 0);
  g->bufB_bytes.assign(n, 0);
  g->t0.assign(n, nullptr);
  g->t1.assign(n, nullptr);
  for (int b = 0; b < 2; ++b) {
    g->bufA[b].assign(n, nullptr);
    g->bufB[b].assign(n, nullptr);
    g->ready[b].assign(n, nullptr);
    g->freed[b].assign(n, nullptr);
  }
  int prev = 0;
  cudaGetDevice(&prev);
  int rc = EC_OK;
  for (int i = 0; i < n && !rc; ++i) {
    rc = ec_pipeline_create(devs[i], &g->pipes[i]);
    if (rc) break;
    cudaSetDevice(devs[i]);
    // The side stream outranks the compute stream: when a panel product and the broadcast of the
    // next panel become runnable at the same moment, the broadcast's few CTAs must get the first
    // SMs that free up -- at equal priority the block scheduler keeps feeding the product's
    // 32768 CTAs and the broadcast only starts once they run out, i.e. serialised behind it
    // (measured: 88 % instead of 95 % of peak per GPU on 2 GPUs).
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    cudaError_t e = cudaStreamCreateWithPriority(&g->comm[i], cudaStreamNonBlocking, prio_hi);
    for (int b = 0; b < 2 && e == cudaSuccess; ++b) {
      e = cudaEventCreateWithFlags(&g->ready[b][i], cudaEventDisableTiming);
      if (e == cudaSuccess) e = cudaEventCreateWithFlags(&g->freed[b][i], cudaEventDisableTiming);
    }
    if (e == cudaSuccess) e = cudaEventCreate(&g->t0[i]);
    if (e == cudaSuccess) e = cudaEventCreate(&g->t1[i]);
    if (e != cudaSuccess) rc = fail(EC_ERR_CUDA, "grid setup on device %d failed: %s", devs[i], cudaGetErrorString(e));
  }
  if (!rc && n > 1) {
    NcclApi *api = nccl_api();
    if (!api->handle)
      rc = fail(EC_ERR_UNSUPPORTED, "the 2D-partitioned DGEMM needs NCCL for its panel broadcasts: %s", api->why.c_str());
    for (int pr = 0; pr < P && !rc && Q > 1; ++pr) {        // one communicator per grid row
      std::vector<int> members;
      for (int pc = 0; pc < Q; ++pc) members.push_back(devs[g->dev(pr, pc)]);
      std::vector<ncclComm_t> comms;
      rc = nccl_init_all(api, members, g->nccl_ctas, &comms);
      for (int pc = 0; pc < Q && !rc; ++pc) g->row_comm[g->dev(pr, pc)] = comms[pc];
    }
    for (int pc = 0; pc < Q && !rc && P > 1; ++pc) {        // one communicator per grid column
      std::vector<int> members;
      for (int pr = 0; pr < P; ++pr) members.push_back(devs[g->dev(pr, pc)]);
      std::vector<ncclComm_t> comms;
      rc = nccl_init_all(api, members, g->nccl_ctas, &comms);
      for (int pr = 0; pr < P && !rc; ++pr) g->col_comm[g->dev(pr, pc)] = comms[pr];
    }
  }
  // Local products share the GPU with NCCL's broadcast kernels: a statically partitioned persistent
  // grid (and the cooperative stream-K grid even more so) runs at the pace of the CTA that had to
  // wait for an SM NCCL held, so the panel products use the schedule that tolerates that.
  if (!rc && n > 1)
    for (ec_pipeline *c : g->pipes) {
      c->stream_k = 0;
      c->one_cta_per_tile = 1;
    }
  cudaSetDevice(prev);
  if (rc) {
    std::string keep = g_last_error;
    grid_destroy_impl(g);
    g_last_error = keep;
    return rc;
  }
  *out = g;
  return EC_OK;
}

// rows / columns of the block of a rows x cols matrix held by device (pr, pc)
struct BlockGeom {
  int64_t r0, r1, c0, c1;
  int64_t nrows() const { return r1 - r0; }
  int64_t ncols() const { return c1 - c0; }
};
BlockGeom block_of(const ec_grid *g, int64_t rows, int64_t cols, int pr, int pc) {
  BlockGeom b;
  block_range(rows, g->P, pr, &b.r0, &b.r1);
  block_range(cols, g->Q, pc, &b.c0, &b.c1);
  return b;
}

int dist_matrix_destroy_impl(ec_dist_matrix *m) {
  if (!m) return EC_OK;
  for (size_t i = 0; i < m->block.size(); ++i)
    if (m->block[i]) {
      DeviceGuard guard(m->devices[i]);
      cudaFree(m->block[i]);   // synchronises with outstanding work on the block
    }
  delete m;
  return EC_OK;
}

int dist_matrix_create_impl(ec_grid *g, int64_t rows, int64_t cols, ec_dist_matrix **out) {
  if (!out) return fail(EC_ERR_ARG, "null out pointer");
  *out = nullptr;
  if (!g) return fail(EC_ERR_ARG, "null grid");
  if (rows < 0 || cols < 0) return fail(EC_ERR_ARG, "negative matrix dimension");
  ec_dist_matrix *m = new ec_dist_matrix();
  m->grid = g;
  m->rows = rows;
  m->cols = cols;
  m->block.assign(g->devices.size(), nullptr);
  m->devices = g->devices;
  for (int pr = 0; pr < g->P; ++pr)
    for (int pc = 0; pc < g->Q; ++pc) {
      const BlockGeom b = block_of(g, rows, cols, pr, pc);
      const size_t bytes = (size_t)b.nrows() * (size_t)b.ncols() * 8;
      if (bytes == 0) continue;
      DeviceGuard guard(g->devices[g->dev(pr, pc)]);
      if (cudaMalloc((void **)&m->block[g->dev(pr, pc)], bytes) != cudaSuccess) {
        cudaGetLastError();
        size_t free_b = 0, total_b = 0;
        cudaMemGetInfo(&free_b, &total_b);
        dist_matrix_destroy_impl(m);
        return fail(EC_ERR_NOMEM,
                    "There were requested : %zu bytes in the device\nDevice Free memory (bytes): %zu\n"
                    "Device total Memory (bytes): %zu\nThere is not enough memory in the Device!",
                    bytes, free_b, total_b);
      }
    }
  *out = m;
  return EC_OK;
}

// k-panels [k0, k1) of width <= kb that never straddle an ownership boundary of A's columns
// (Q parts) or B's rows (P parts)
std::vector<std::pair<int64_t, int64_t>> summa_panels(int64_t k, int P, int Q, int64_t kb) {
  std::vector<int64_t> cuts{0, k};
  for (int q = 0; q < Q; ++q) cuts.push_back(k * q / Q);
  for (int p = 0; p < P; ++p) cuts.push_back(k * p / P);
  std::sort(cuts.begin(), cuts.end());
  cuts.erase(std::unique(cuts.begin(), cuts.end()), cuts.end());
  std::vector<std::pair<int64_t, int64_t>> out;
  for (size_t i = 0; i + 1 < cuts.size(); ++i)
    for (int64_t s = cuts[i]; s < cuts[i + 1]; s += kb) out.emplace_back(s, std::min(s + kb, cuts[i + 1]));
  return out;
}

int owner_of(int64_t total, int parts, int64_t x, int64_t *lo) {
  for (int i = 0; i < parts; ++i) {
    int64_t a, b;
    block_range(total, parts, i, &a, &b);
    if (a <= x && x < b) {
      *lo = a;
      return i;
    }
  }
  *lo = 0;
  return 0;
}

int dgemm_2d_impl(ec_grid *g, const ec_dist_matrix *A, const ec_dist_matrix *B, ec_dist_matrix *C,
                  int64_t panel, double *elapsed_ms) {
  if (!g || !A || !B || !C) return fail(EC_ERR_ARG, "null handle");
  if (A->grid != g || B->grid != g || C->grid != g) return fail(EC_ERR_ARG, "matrices belong to another grid");
  if (A->cols != B->rows) return fail(EC_ERR_SHAPE, "Shape mismatch in Cublas gemm");
  if (C->rows != A->rows || C->cols != B->cols) return fail(EC_ERR_SHAPE, "Shape mismatch in gemm: C is not A.rows x B.cols");
  const int64_t m = A->rows, n = B->cols, k = A->cols;
  if (elapsed_ms) *elapsed_ms = 0.0;
  if (m == 0 || n == 0) return EC_OK;
  const int P = g->P, Q = g->Q, G = P * Q;
  const int64_t kb = panel > 0 ? panel : 4096;
  const std::vector<std::pair<int64_t, int64_t>> panels = summa_panels(k, P, Q, kb);
  int64_t wmax = 0;
  for (const auto &pn : panels) wmax = std::max(wmax, pn.second - pn.first);
  NcclApi *api = nccl_api();
  int prev = 0;
  cudaGetDevice(&prev);
  struct Restore {
    int d;
    ~Restore() { cudaSetDevice(d); }
  } restore{prev};
  TraceRange trace("ec_dgemm_2d");

  // ---- per-device panel buffers (only where a broadcast actually happens) ----
  std::vector<BlockGeom> cg(G);
  for (int pr = 0; pr < P; ++pr)
    for (int pc = 0; pc < Q; ++pc) {
      const int i = g->dev(pr, pc);
      cg[i] = block_of(g, m, n, pr, pc);
      EC_CUDA(cudaSetDevice(g->devices[i]));
      const size_t needA = Q > 1 ? (size_t)cg[i].nrows() * (size_t)wmax * 8 : 0;
      const size_t needB = P > 1 ? (size_t)wmax * (size_t)cg[i].ncols() * 8 : 0;
      if (needA > g->bufA_bytes[i]) {
        for (int b = 0; b < 2; ++b) {
          if (g->bufA[b][i]) EC_CUDA(cudaFree(g->bufA[b][i]));
          g->bufA[b][i] = nullptr;
          EC_CUDA(cudaMalloc((void **)&g->bufA[b][i], needA));
        }
        g->bufA_bytes[i] = needA;
      }
      if (needB > g->bufB_bytes[i]) {
        for (int b = 0; b < 2; ++b) {
          if (g->bufB[b][i]) EC_CUDA(cudaFree(g->bufB[b][i]));
          g->bufB[b][i] = nullptr;
          EC_CUDA(cudaMalloc((void **)&g->bufB[b][i], needB));
        }
        g->bufB_bytes[i] = needB;
      }
      // the side stream starts behind whatever is queued on the compute stream (fills, uploads)
      EC_CUDA(cudaEventRecord(g->t0[i], g->pipes[i]->stream));
      EC_CUDA(cudaStreamWaitEvent(g->comm[i], g->t0[i], 0));
    }

  // ---- k == 0: C = 0 ----
  if (k == 0) {
    for (int i = 0; i < G; ++i) {
      if (!C->block[i]) continue;
      EC_CUDA(cudaSetDevice(g->devices[i]));
      EC_CUDA(cudaMemsetAsync(C->block[i], 0, (size_t)cg[i].nrows() * (size_t)cg[i].ncols() * 8, g->pipes[i]->stream));
    }
    for (int i = 0; i < G; ++i) {
      EC_CUDA(cudaSetDevice(g->devices[i]));
      EC_CUDA(cudaStreamSynchronize(g->pipes[i]->stream));
    }
    return EC_OK;
  }

  // Issue the traffic of panel t into buffer b on every device's side stream.
  auto fetch = [&](size_t t, int b) -> int {
    const int64_t k0 = panels[t].first, w = panels[t].second - k0;
    int64_t a_lo = 0, b_lo = 0;
    const int qa = owner_of(k, Q, k0, &a_lo);   // grid column that owns these columns of A
    const int pb = owner_of(k, P, k0, &b_lo);   // grid row that owns these rows of B
    for (int i = 0; i < G; ++i) {               // buffer b must have been consumed by panel t-2
      if (t < 2) break;
      EC_CUDA(cudaSetDevice(g->devices[i]));
      EC_CUDA(cudaStreamWaitEvent(g->comm[i], g->freed[b][i], 0));
    }
    if (P > 1) {
      // pack the owner's w rows of B (strided: ld = its row count) into the panel buffer: a 2D
      // device-to-device copy on the copy engine, no kernel, no SM
      for (int pc = 0; pc < Q; ++pc) {
        const int i = g->dev(pb, pc);
        const BlockGeom bg = block_of(g, k, n, pb, pc);
        if (bg.ncols() == 0) continue;
        EC_CUDA(cudaSetDevice(g->devices[i]));
        EC_CUDA(cudaMemcpy2DAsync(g->bufB[b][i], (size_t)w * 8, B->block[i] + (k0 - b_lo), (size_t)bg.nrows() * 8,
                                  (size_t)w * 8, (size_t)bg.ncols(), cudaMemcpyDeviceToDevice, g->comm[i]));
      }
    }
    if (G > 1) {
      ncclResult_t r = api->GroupStart();
      for (int pr = 0; pr < P && r == ncclSuccess && Q > 1; ++pr) {
        const int root = g->dev(pr, qa);
        const BlockGeom ag = block_of(g, m, k, pr, qa);
        const double *src = A->block[root] + (k0 - a_lo) * ag.nrows();   // w whole columns: contiguous
        for (int pc = 0; pc < Q && r == ncclSuccess; ++pc) {
          const int i = g->dev(pr, pc);
          if (cg[i].nrows() == 0) continue;
          cudaSetDevice(g->devices[i]);
          r = api->Broadcast(pc == qa ? src : g->bufA[b][i], g->bufA[b][i], (size_t)(cg[i].nrows() * w), ncclDouble, qa,
                             g->row_comm[i], g->comm[i]);
        }
      }
      for (int pc = 0; pc < Q && r == ncclSuccess && P > 1; ++pc) {
        for (int pr = 0; pr < P && r == ncclSuccess; ++pr) {
          const int i = g->dev(pr, pc);
          if (cg[i].ncols() == 0) continue;
          cudaSetDevice(g->devices[i]);
          r = api->Broadcast(g->bufB[b][i], g->bufB[b][i], (size_t)(w * cg[i].ncols()), ncclDouble, pb, g->col_comm[i], g->comm[i]);
        }
      }
      const ncclResult_t r2 = api->GroupEnd();
      if (r != ncclSuccess || r2 != ncclSuccess)
        return fail(EC_ERR_CUDA, "NCCL panel broadcast failed: %s", api->GetErrorString(r != ncclSuccess ? r : r2));
      g->broadcasts++;
    }
    for (int i = 0; i < G; ++i) {
      EC_CUDA(cudaSetDevice(g->devices[i]));
      EC_CUDA(cudaEventRecord(g->ready[b][i], g->comm[i]));
    }
    return EC_OK;
  };

  for (int i = 0; i < G; ++i) {
    EC_CUDA(cudaSetDevice(g->devices[i]));
    EC_CUDA(cudaEventRecord(g->t0[i], g->pipes[i]->stream));
  }
  int rc = fetch(0, 0);
  if (rc) return rc;
  for (size_t t = 0; t < panels.size(); ++t) {
    const int b = (int)(t & 1);
    if (t + 1 < panels.size() && (rc = fetch(t + 1, b ^ 1))) return rc;
    const int64_t k0 = panels[t].first, w = panels[t].second - k0;
    int64_t a_lo = 0, b_lo = 0;
    owner_of(k, Q, k0, &a_lo);
    owner_of(k, P, k0, &b_lo);
    for (int pr = 0; pr < P; ++pr)
      for (int pc = 0; pc < Q; ++pc) {
        const int i = g->dev(pr, pc);
        if (cg[i].nrows() == 0 || cg[i].ncols() == 0) continue;
        EC_CUDA(cudaSetDevice(g->devices[i]));
        ec_pipeline *p = g->pipes[i];
        EC_CUDA(cudaStreamWaitEvent(p->stream, g->ready[b][i], 0));
        // operands: the received panel, or -- along a grid dimension of extent 1 -- the local block itself
        const double *Ap = Q > 1 ? g->bufA[b][i] : A->block[i] + (k0 - a_lo) * cg[i].nrows();
        const int64_t lda = cg[i].nrows();
        const double *Bp = P > 1 ? g->bufB[b][i] : B->block[i] + (k0 - b_lo);
        const int64_t ldb = P > 1 ? w : block_of(g, k, n, pr, pc).nrows();
        rc = dgemm_dispatch(p, EC_OP_N, EC_OP_N, cg[i].nrows(), cg[i].ncols(), w, 1.0, Ap, lda, 0, Bp, ldb, 0,
                            t == 0 ? 0.0 : 1.0, C->block[i], cg[i].nrows(), 0, 1);
        if (rc) return rc;
        EC_CUDA(cudaEventRecord(g->freed[b][i], p->stream));
      }
  }
  double worst = 0.0;
  for (int i = 0; i < G; ++i) {
    EC_CUDA(cudaSetDevice(g->devices[i]));
    EC_CUDA(cudaEventRecord(g->t1[i], g->pipes[i]->stream));
  }
  for (int i = 0; i < G; ++i) {
    EC_CUDA(cudaSetDevice(g->devices[i]));
    EC_CUDA(cudaStreamSynchronize(g->pipes[i]->stream));
    EC_CUDA(cudaStreamSynchronize(g->comm[i]));
    float ms = 0.f;
    EC_CUDA(cudaEventElapsedTime(&ms, g->t0[i], g->t1[i]));
    worst = std::max(worst, (double)ms);
  }
  if (elapsed_ms) *elapsed_ms = worst;   // on the device, max over the GPUs
  return EC_OK;
}

// Host <-> blocks: one strided 2D copy per block (pitch = the host matrix's leading dimension).
int dist_matrix_transfer(const ec_dist_matrix *mat, double *host, int64_t ld, bool to_device) {
  if (!mat) return fail(EC_ERR_ARG, "null matrix");
  if (mat->rows == 0 || mat->cols == 0) return EC_OK;
  if (!host) return fail(EC_ERR_ARG, "null host pointer");
  if (ld < mat->rows) return fail(EC_ERR_ARG, "leading dimension too small");
  const ec_grid *g = mat->grid;
  int prev = 0;
  cudaGetDevice(&prev);
  int rc = EC_OK;
  for (int pr = 0; pr < g->P && !rc; ++pr)
    for (int pc = 0; pc < g->Q && !rc; ++pc) {
      const int i = g->dev(pr, pc);
      const BlockGeom b = block_of(g, mat->rows, mat->cols, pr, pc);
      if (b.nrows() == 0 || b.ncols() == 0) continue;
      cudaSetDevice(g->devices[i]);
      double *h = host + b.c0 * ld + b.r0;
      cudaError_t e = to_device
          ? cudaMemcpy2DAsync(mat->block[i], (size_t)b.nrows() * 8, h, (size_t)ld * 8, (size_t)b.nrows() * 8,
                              (size_t)b.ncols(), cudaMemcpyHostToDevice, g->pipes[i]->stream)
          : cudaMemcpy2DAsync(h, (size_t)ld * 8, mat->block[i], (size_t)b.nrows() * 8, (size_t)b.nrows() * 8,
                              (size_t)b.ncols(), cudaMemcpyDeviceToHost, g->pipes[i]->stream);
      if (e != cudaSuccess) rc = fail(EC_ERR_COPY, "Error copy arrays %s device: %s", to_device ? "to" : "from", cudaGetErrorString(e));
    }
  for (int i = 0; i < g->P * g->Q; ++i) {   // complete on return, like the reference's blocking conversion
    cudaSetDevice(g->devices[i]);
    if (cudaStreamSynchronize(g->pipes[i]->stream) != cudaSuccess && !rc) rc = fail(EC_ERR_CUDA, "stream synchronize failed");
  }
  cudaSetDevice(prev);
  return rc;
}

}  // namespace
