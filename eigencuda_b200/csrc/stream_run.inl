// eigencuda_b200/csrc/stream_run.inl -- one tensor-stream job on one device: plan, feeder, drainer
// Part of the single translation unit ec_api.cu (included there, in order); not a standalone header.
#pragma once

namespace {

// The reference's whole streaming loop (README.md:62-66) on ONE device: see ec_tensor_stream_run.
int tensor_stream_run_impl(ec_pipeline *p, int kind, const double *F, int64_t f_rows, int64_t f_cols,
                           const double *const *tensor, double *const *results, int64_t count,
                           int64_t t_rows, int64_t t_cols) {
  if (!p) return fail(EC_ERR_ARG, "null pipeline");
  if (count < 0 || f_rows < 0 || f_cols < 0 || t_rows < 0 || t_cols < 0)
    return fail(EC_ERR_ARG, "negative size");
  StreamShape sh;
  int rc = stream_shape(kind, f_rows, f_cols, t_rows, t_cols, &sh);  // before any device work
  if (rc) return rc;
  if (count == 0) return EC_OK;
  if (!F || !tensor || !results) return fail(EC_ERR_ARG, "null pointer");
  const int64_t t_sz = t_rows * t_cols, r_sz = sh.r_rows * sh.r_cols, f_sz = f_rows * f_cols;
  if (t_sz == 0 || r_sz == 0 || f_sz == 0) {
    // degenerate: products with an empty inner dimension are all-zero
    for (int64_t i = 0; i < count; ++i)
      if (r_sz) std::memset(results[i], 0, (size_t)r_sz * 8);
    return EC_OK;
  }
  DeviceGuard guard(p->device);
  TraceRange trace("ec_tensor_stream_run");

  // ---- plan ----
  bool in_pinned = true, out_pinned = true;
  // One attribute query per run of adjacent matrices (first and last element of the run), not
  // per matrix: a contiguous tensor costs two queries.
  auto all_pinned = [&](const double *const *ptrs, int64_t elems) {
    int64_t i = 0;
    while (i < count) {
      int64_t j = i + 1;
      while (j < count && ptrs[j] == ptrs[j - 1] + elems) ++j;
      if (!is_pinned_host(ptrs[i]) || !is_pinned_host(ptrs[j - 1] + elems - 1)) return false;
      i = j;
    }
    return true;
  };
  in_pinned = all_pinned(tensor, t_sz);
  out_pinned = all_pinned((const double *const *)results, r_sz);
  const bool staged_in = !in_pinned, staged_out = !out_pinned;
  // chunk = matrices per ring slot (the largest chunk); default 64 MiB worth, with short first and
  // last chunks (plan_chunks); an explicit setting gives uniform chunks of exactly that size
  int64_t chunk = p->cfg_chunk;
  const bool ramp = chunk <= 0;
  if (chunk <= 0) chunk = std::max<int64_t>(1, ((int64_t)64 << 20) / (8 * std::max(t_sz, r_sz)));
  chunk = std::min(chunk, count);
  const std::vector<int64_t> starts = plan_chunks(count, chunk, ramp);
  const int ring = p->cfg_ring > 0 ? p->cfg_ring : 4;
  int threads = p->cfg_threads;
  if (threads <= 0) threads = std::max(2u, std::min(16u, std::thread::hardware_concurrency()));
  if (!p->engine) {
    p->engine = new StreamEngine();
    p->engine->p = p;
  }
  StreamEngine &E = *p->engine;
  const bool f_on_device = is_device_ptr(F);
  const size_t tmp_b = kind == EC_STREAM_BASIS ? (size_t)(chunk * f_cols * t_cols * 8) : 0;
  rc = E.ensure(ring, (size_t)(chunk * t_sz * 8), (size_t)(chunk * r_sz * 8), tmp_b,
                f_on_device ? 0 : (size_t)(f_sz * 8), staged_in, staged_out, threads);
  if (rc) return rc;

  // Order the engine's streams after whatever the caller queued on the pipeline stream.
  EC_CUDA(cudaEventRecord(E.ev_fixed, p->stream));
  EC_CUDA(cudaStreamWaitEvent(E.s_h2d, E.ev_fixed, 0));
  EC_CUDA(cudaStreamWaitEvent(E.s_cmp, E.ev_fixed, 0));

  // ---- fixed operand: uploaded once (or used in place when already on the device, e.g. after
  // the multi-GPU broadcast) ----
  const double *dF = F;
  if (!f_on_device) {
    int slot = -1;
    if (is_pinned_host(F)) {
      EC_CUDA(cudaMemcpyAsync(E.d_fixed, F, (size_t)f_sz * 8, cudaMemcpyHostToDevice, E.s_h2d));
    } else {
      rc = p->staging.acquire((size_t)f_sz * 8, &slot);
      if (rc) return rc;
      std::memcpy(p->staging.buf[slot], F, (size_t)f_sz * 8);
      EC_CUDA(cudaMemcpyAsync(E.d_fixed, p->staging.buf[slot], (size_t)f_sz * 8,
                              cudaMemcpyHostToDevice, E.s_h2d));
      rc = p->staging.release_after(slot, E.s_h2d);
      if (rc) return rc;
    }
    dF = E.d_fixed;
  }
  EC_CUDA(cudaEventRecord(E.ev_fixed, E.s_h2d));
  EC_CUDA(cudaStreamWaitEvent(E.s_cmp, E.ev_fixed, 0));

  const int64_t nchunks = (int64_t)starts.size() - 1;

  // ---- drainer (pageable results only) ----
  std::mutex mu;
  std::condition_variable cv;
  int64_t enqueued = 0, drained = 0;  // chunks whose D2H is enqueued / whose results are delivered
  bool abort_flag = false;
  std::atomic<int> drain_rc{EC_OK};
  std::thread drainer;
  if (staged_out) {
    drainer = std::thread([&] {
      cudaSetDevice(p->device);
      for (int64_t c = 0; c < nchunks; ++c) {
        {
          std::unique_lock<std::mutex> lk(mu);
          cv.wait(lk, [&] { return enqueued > c || abort_flag; });
          if (abort_flag) return;
        }
        const int s = (int)(c % ring);
        if (cudaEventSynchronize(E.ev_d2h[s]) != cudaSuccess) {
          drain_rc = EC_ERR_CUDA;
        } else {
          const int64_t i0 = starts[c], nb = starts[c + 1] - i0;
          TraceRange trace_chunk("deliver chunk (pinned -> caller)");
          const char *src = (const char *)E.pin_out[s];
          // split every matrix into a few pieces so the pool stays busy even for small chunks
          const int pieces = std::max<int>(1, (int)std::min<int64_t>(8, (r_sz * 8) >> 18));
          E.pool_out->parallel_for(nb * pieces, [&](int64_t t) {
            const int64_t i = t / pieces, q = t % pieces;
            const size_t lo = (size_t)(r_sz * 8) * q / pieces, hi = (size_t)(r_sz * 8) * (q + 1) / pieces;
            // result matrices are often freshly allocated: map the piece in one call, not per page
            ec_host::prefault_for_write((char *)results[i0 + i] + lo, hi - lo);
            staging_copy((char *)results[i0 + i] + lo, src + (size_t)i * r_sz * 8 + lo, hi - lo);
          });
        }
        {
          std::lock_guard<std::mutex> lk(mu);
          drained = c + 1;
        }
        cv.notify_all();
      }
    });
  }

  // ---- feeder (this thread) ----
  auto feeder = [&]() -> int {
    for (int64_t c = 0; c < nchunks; ++c) {
      const int s = (int)(c % ring);
      const int64_t i0 = starts[c], nb = starts[c + 1] - i0;
      // H2D
      if (staged_in) {
        if (c >= ring) EC_CUDA(cudaEventSynchronize(E.ev_h2d[s]));  // pinned_in[s] consumed
        TraceRange trace_chunk("stage chunk (caller -> pinned)");
        char *dst = (char *)E.pin_in[s];
        const int pieces = std::max<int>(1, (int)std::min<int64_t>(8, (t_sz * 8) >> 18));
        E.pool_in->parallel_for(nb * pieces, [&](int64_t t) {
          const int64_t i = t / pieces, q = t % pieces;
          const size_t lo = (size_t)(t_sz * 8) * q / pieces, hi = (size_t)(t_sz * 8) * (q + 1) / pieces;
          staging_copy(dst + (size_t)i * t_sz * 8 + lo, (const char *)tensor[i0 + i] + lo, hi - lo);
        });
      }
      if (c >= ring) EC_CUDA(cudaStreamWaitEvent(E.s_h2d, E.ev_cmp[s], 0));  // d_in[s] consumed
      if (staged_in)
        EC_CUDA(cudaMemcpyAsync(E.d_in[s], E.pin_in[s], (size_t)(nb * t_sz * 8),
                                cudaMemcpyHostToDevice, E.s_h2d));
      else if ((rc = dma_chunk(true, E.d_in[s], tensor, i0, nb, t_sz, E.s_h2d)))
        return rc;
      EC_CUDA(cudaEventRecord(E.ev_h2d[s], E.s_h2d));
      // compute (on the engine's compute stream: temporarily retarget the pipeline)
      EC_CUDA(cudaStreamWaitEvent(E.s_cmp, E.ev_h2d[s], 0));
      if (c >= ring) EC_CUDA(cudaStreamWaitEvent(E.s_cmp, E.ev_d2h[s], 0));  // d_out[s] drained
      {
        StreamScope scope(p, E.s_cmp);
        rc = stream_compute(p, kind, dF, f_rows, f_cols, E.d_in[s], E.d_out[s],
                            kind == EC_STREAM_BASIS ? E.d_tmp[s] : nullptr, nb, t_rows, t_cols, sh);
      }
      if (rc) return rc;
      EC_CUDA(cudaEventRecord(E.ev_cmp[s], E.s_cmp));
      // D2H
      EC_CUDA(cudaStreamWaitEvent(E.s_d2h, E.ev_cmp[s], 0));
      if (staged_out) {
        if (c >= ring) {  // pin_out[s] must have been delivered to the caller
          std::unique_lock<std::mutex> lk(mu);
          cv.wait(lk, [&] { return drained >= c - ring + 1 || drain_rc != EC_OK; });
        }
        EC_CUDA(cudaMemcpyAsync(E.pin_out[s], E.d_out[s], (size_t)(nb * r_sz * 8),
                                cudaMemcpyDeviceToHost, E.s_d2h));
      } else if ((rc = dma_chunk(false, E.d_out[s], (const double *const *)results, i0, nb, r_sz,
                                 E.s_d2h))) {
        return rc;
      }
      EC_CUDA(cudaEventRecord(E.ev_d2h[s], E.s_d2h));
      if (staged_out) {
        {
          std::lock_guard<std::mutex> lk(mu);
          enqueued = c + 1;
        }
        cv.notify_all();
      }
    }
    return EC_OK;
  };
  rc = feeder();
  if (rc) {
    std::string keep = g_last_error;
    {
      std::lock_guard<std::mutex> lk(mu);
      abort_flag = true;
    }
    cv.notify_all();
    if (drainer.joinable()) drainer.join();
    cudaStreamSynchronize(E.s_h2d);
    cudaStreamSynchronize(E.s_cmp);
    cudaStreamSynchronize(E.s_d2h);
    g_last_error = keep;
    return rc;
  }
  if (drainer.joinable()) drainer.join();
  EC_CUDA(cudaStreamSynchronize(E.s_d2h));
  EC_CUDA(cudaStreamSynchronize(E.s_cmp));
  EC_CUDA(cudaStreamSynchronize(E.s_h2d));
  if (drain_rc != EC_OK) return fail(EC_ERR_CUDA, "device-to-host drain failed");
  return EC_OK;
}

}  // namespace
