// eigencuda_b200/csrc/stream_engine.inl -- the overlapped tensor-stream engine (ring of device chunks, three streams, staging threads)
// Part of the single translation unit ec_api.cu (included there, in order); not a standalone header.
#pragma once

// ------------------------------------------------------------------------------------------
// tensor-stream engine
// ------------------------------------------------------------------------------------------
// One job = the reference's whole streaming loop (README.md:62-66).  The tensor is cut into
// chunks of `chunk` matrices; chunk c lives in ring slot c % ring.  Per slot: a device input
// buffer, a device output buffer, (BASIS only) a device intermediate, and -- when the caller's
// memory is pageable -- a pinned input and a pinned output staging buffer.  Three streams:
//     h2d:     [wait compute(c-ring)]  copy chunk c in            -> ev_h2d[s]
//     compute: [wait ev_h2d[s], d2h(c-ring)]  batched DGEMM(s)    -> ev_cmp[s]
//     d2h:     [wait ev_cmp[s]]  copy chunk c out                 -> ev_d2h[s]
// With pinned/registered caller memory the host thread only enqueues (DMA straight from/to the
// caller's buffers).  With pageable memory a feeder thread stages chunks in (worker pool memcpy
// into pinned) and a drainer thread stages results out, so both host copies overlap the GPU.
struct StreamEngine {
  ec_pipeline *p = nullptr;
  cudaStream_t s_h2d = nullptr, s_cmp = nullptr, s_d2h = nullptr;
  int ring = 0;
  size_t in_bytes = 0, out_bytes = 0, tmp_bytes = 0, pin_in_bytes = 0, pin_out_bytes = 0;
  std::vector<double *> d_in, d_out, d_tmp;
  std::vector<void *> pin_in, pin_out;
  std::vector<cudaEvent_t> ev_h2d, ev_cmp, ev_d2h;
  double *d_fixed = nullptr;
  size_t fixed_bytes = 0;
  cudaEvent_t ev_fixed = nullptr;
  CopyPool *pool_in = nullptr, *pool_out = nullptr;
  int pool_threads = 0;

  int ensure(int ring_, size_t in_b, size_t out_b, size_t tmp_b, size_t fixed_b, bool staged_in,
             bool staged_out, int threads) {
    if (!s_h2d) {
      EC_CUDA(cudaStreamCreateWithFlags(&s_h2d, cudaStreamNonBlocking));
      EC_CUDA(cudaStreamCreateWithFlags(&s_cmp, cudaStreamNonBlocking));
      EC_CUDA(cudaStreamCreateWithFlags(&s_d2h, cudaStreamNonBlocking));
      EC_CUDA(cudaEventCreateWithFlags(&ev_fixed, cudaEventDisableTiming));
    }
    if (ring_ != ring) {
      int rc = release_buffers();
      if (rc) return rc;
      ring = ring_;
      d_in.assign(ring, nullptr);
      d_out.assign(ring, nullptr);
      d_tmp.assign(ring, nullptr);
      pin_in.assign(ring, nullptr);
      pin_out.assign(ring, nullptr);
      ev_h2d.assign(ring, nullptr);
      ev_cmp.assign(ring, nullptr);
      ev_d2h.assign(ring, nullptr);
      for (int s = 0; s < ring; ++s) {
        EC_CUDA(cudaEventCreateWithFlags(&ev_h2d[s], cudaEventDisableTiming));
        EC_CUDA(cudaEventCreateWithFlags(&ev_cmp[s], cudaEventDisableTiming));
        EC_CUDA(cudaEventCreateWithFlags(&ev_d2h[s], cudaEventDisableTiming));
      }
    }
    auto grow_dev = [&](std::vector<double *> &v, size_t &have, size_t want) -> int {
      if (want <= have) return EC_OK;
      for (int s = 0; s < ring; ++s) {
        if (v[s]) EC_CUDA(cudaFree(v[s]));
        v[s] = nullptr;
      }
      have = 0;
      for (int s = 0; s < ring; ++s) EC_CUDA(cudaMalloc((void **)&v[s], want));
      have = want;
      return EC_OK;
    };
    auto grow_pin = [&](std::vector<void *> &v, size_t &have, size_t want) -> int {
      if (want <= have) return EC_OK;
      for (int s = 0; s < ring; ++s) {
        if (v[s]) EC_CUDA(cudaFreeHost(v[s]));
        v[s] = nullptr;
      }
      have = 0;
      for (int s = 0; s < ring; ++s) EC_CUDA(cudaHostAlloc(&v[s], want, cudaHostAllocDefault));
      have = want;
      return EC_OK;
    };
    int rc;
    if ((rc = grow_dev(d_in, in_bytes, in_b))) return rc;
    if ((rc = grow_dev(d_out, out_bytes, out_b))) return rc;
    if (tmp_b && (rc = grow_dev(d_tmp, tmp_bytes, tmp_b))) return rc;
    if (staged_in && (rc = grow_pin(pin_in, pin_in_bytes, in_b))) return rc;
    if (staged_out && (rc = grow_pin(pin_out, pin_out_bytes, out_b))) return rc;
    if (fixed_b > fixed_bytes) {
      if (d_fixed) EC_CUDA(cudaFree(d_fixed));
      d_fixed = nullptr;
      fixed_bytes = 0;
      EC_CUDA(cudaMalloc((void **)&d_fixed, fixed_b));
      fixed_bytes = fixed_b;
    }
    if ((staged_in || staged_out) && (!pool_in || pool_threads != threads)) {
      delete pool_in;
      delete pool_out;
      pool_in = new CopyPool(std::max(1, threads / 2));
      pool_out = new CopyPool(std::max(1, threads - threads / 2));
      pool_threads = threads;
    }
    return EC_OK;
  }

  int release_buffers() {
    for (auto &x : d_in) if (x) { EC_CUDA(cudaFree(x)); x = nullptr; }
    for (auto &x : d_out) if (x) { EC_CUDA(cudaFree(x)); x = nullptr; }
    for (auto &x : d_tmp) if (x) { EC_CUDA(cudaFree(x)); x = nullptr; }
    for (auto &x : pin_in) if (x) { EC_CUDA(cudaFreeHost(x)); x = nullptr; }
    for (auto &x : pin_out) if (x) { EC_CUDA(cudaFreeHost(x)); x = nullptr; }
    for (auto &e : ev_h2d) if (e) { cudaEventDestroy(e); e = nullptr; }
    for (auto &e : ev_cmp) if (e) { cudaEventDestroy(e); e = nullptr; }
    for (auto &e : ev_d2h) if (e) { cudaEventDestroy(e); e = nullptr; }
    in_bytes = out_bytes = tmp_bytes = pin_in_bytes = pin_out_bytes = 0;
    return EC_OK;
  }

  void destroy() {
    release_buffers();
    if (d_fixed) cudaFree(d_fixed);
    d_fixed = nullptr;
    if (ev_fixed) cudaEventDestroy(ev_fixed);
    if (s_h2d) cudaStreamDestroy(s_h2d);
    if (s_cmp) cudaStreamDestroy(s_cmp);
    if (s_d2h) cudaStreamDestroy(s_d2h);
    delete pool_in;
    delete pool_out;
    pool_in = pool_out = nullptr;
  }
};

namespace {

struct StreamShape {
  int64_t r_rows, r_cols;  // result shape
};

int stream_shape(int kind, int64_t f_rows, int64_t f_cols, int64_t t_rows, int64_t t_cols,
                 StreamShape *out) {
  switch (kind) {
    case EC_STREAM_RIGHT:  // T(t_rows x t_cols) * F(f_rows x f_cols)
      if (t_cols != f_rows) return fail(EC_ERR_SHAPE, "Shape mismatch in gemm: T.cols != F.rows");
      *out = {t_rows, f_cols};
      return EC_OK;
    case EC_STREAM_LEFT:
      if (f_cols != t_rows) return fail(EC_ERR_SHAPE, "Shape mismatch in gemm: F.cols != T.rows");
      *out = {f_rows, t_cols};
      return EC_OK;
    case EC_STREAM_LEFT_T:
      if (f_rows != t_rows) return fail(EC_ERR_SHAPE, "Shape mismatch in gemm: F^T.cols != T.rows");
      *out = {f_cols, t_cols};
      return EC_OK;
    case EC_STREAM_BASIS:
      if (f_rows != t_rows || t_cols != f_rows)
        return fail(EC_ERR_SHAPE, "Shape mismatch in basis transform F^T * T * F");
      *out = {f_cols, f_cols};
      return EC_OK;
  }
  return fail(EC_ERR_ARG, "unknown tensor-stream kind %d", kind);
}

// Chunk boundaries of one job: chunk c covers matrices [starts[c], starts[c+1]).  Uniform chunks
// of `chunk` matrices; with `ramp` the first chunks grow chunk/8, chunk/4, chunk/2 and the last
// ones shrink the same way.  The H2D of the first chunk and the D2H of the last one cannot overlap
// anything, so they should be short, while the chunks in between should be long: every DMA pays a
// fixed cost, and the link only reaches its ceiling from about 64 MiB per transfer
// (profiles/r02_engine_sweep.txt: 8 MiB chunks 39 GB/s, 32 MiB 44.6, 64-128 MiB 46 of 48 GB/s).
std::vector<int64_t> plan_chunks(int64_t count, int64_t chunk, bool ramp) {
  std::vector<int64_t> starts{0};
  chunk = std::max<int64_t>(1, std::min(chunk, std::max<int64_t>(count, 1)));
  if (!ramp || chunk < 8 || count < 6 * chunk) {
    for (int64_t i = chunk; i < count; i += chunk) starts.push_back(i);
    if (count > 0) starts.push_back(count);
    return starts;
  }
  const int64_t steps[3] = {chunk / 8, chunk / 4, chunk / 2};
  int64_t head = 0, tail_total = steps[0] + steps[1] + steps[2];
  for (int64_t sz : steps) starts.push_back(head += sz);
  while (head + chunk <= count - tail_total) starts.push_back(head += chunk);
  // what is left -- tail_total plus a remainder below one chunk -- goes out in shrinking pieces; the
  // remainder joins the first of them, or goes ahead of it when that would overflow a ring slot
  const int64_t extra = count - head - tail_total;
  int64_t first = steps[2] + extra;
  if (first > chunk) {
    starts.push_back(head += extra);
    first = steps[2];
  }
  starts.push_back(head += first);
  starts.push_back(head += steps[1]);
  starts.push_back(head += steps[0]);
  return starts;
}

// Launch the product(s) for `nb` matrices at d_in -> d_out on the pipeline's *current* stream.
int stream_compute(ec_pipeline *p, int kind, const double *dF, int64_t f_rows, int64_t f_cols,
                   const double *d_in, double *d_out, double *d_tmp, int64_t nb, int64_t t_rows,
                   int64_t t_cols, const StreamShape &sh) {
  const int64_t t_sz = t_rows * t_cols, r_sz = sh.r_rows * sh.r_cols;
  switch (kind) {
    case EC_STREAM_RIGHT:
      return dgemm_dispatch(p, EC_OP_N, EC_OP_N, t_rows, f_cols, t_cols, 1.0, d_in, t_rows, t_sz,
                            dF, f_rows, 0, 0.0, d_out, sh.r_rows, r_sz, nb);
    case EC_STREAM_LEFT:
      return dgemm_dispatch(p, EC_OP_N, EC_OP_N, f_rows, t_cols, f_cols, 1.0, dF, f_rows, 0, d_in,
                            t_rows, t_sz, 0.0, d_out, sh.r_rows, r_sz, nb);
    case EC_STREAM_LEFT_T:
      return dgemm_dispatch(p, EC_OP_T, EC_OP_N, f_cols, t_cols, f_rows, 1.0, dF, f_rows, 0, d_in,
                            t_rows, t_sz, 0.0, d_out, sh.r_rows, r_sz, nb);
    case EC_STREAM_BASIS: {
      const int64_t tmp_sz = f_cols * t_cols;
      int rc = dgemm_dispatch(p, EC_OP_T, EC_OP_N, f_cols, t_cols, f_rows, 1.0, dF, f_rows, 0,
                              d_in, t_rows, t_sz, 0.0, d_tmp, f_cols, tmp_sz, nb);
      if (rc) return rc;
      return dgemm_dispatch(p, EC_OP_N, EC_OP_N, f_cols, f_cols, t_cols, 1.0, d_tmp, f_cols,
                            tmp_sz, dF, f_rows, 0, 0.0, d_out, sh.r_rows, r_sz, nb);
    }
  }
  return fail(EC_ERR_ARG, "unknown tensor-stream kind %d", kind);
}

// Copy matrices [i0, i0+nb) between per-matrix host pointers and one contiguous device chunk,
// merging runs of adjacent host matrices into single DMA transfers.
int dma_chunk(bool to_device, double *dev, const double *const *host, int64_t i0, int64_t nb,
              int64_t elems, cudaStream_t st) {
  int64_t i = 0;
  while (i < nb) {
    int64_t j = i + 1;
    while (j < nb && host[i0 + j] == host[i0 + j - 1] + elems) ++j;
    const size_t bytes = (size_t)(j - i) * elems * 8;
    if (to_device)
      EC_CUDA(cudaMemcpyAsync(dev + i * elems, host[i0 + i], bytes, cudaMemcpyHostToDevice, st));
    else
      EC_CUDA(cudaMemcpyAsync((void *)host[i0 + i], dev + i * elems, bytes, cudaMemcpyDeviceToHost, st));
    i = j;
  }
  return EC_OK;
}

}  // namespace
