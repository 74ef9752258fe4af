// eigencuda_b200/csrc/multi_gpu.inl -- the multi-GPU tensor scheduler: one tensor-stream job sharded over the devices of one box
// Part of the single translation unit ec_api.cu (included there, in order); not a standalone header.
//
// The caller hands over ONE std::vector of matrices (as arrays of host pointers) and a device
// list.  The scheduler keeps, per device, a child pipeline with its own tensor-stream engine
// (pinned ring + H2D / compute / D2H streams) and one host worker thread; device g takes the
// contiguous index range [g*N/G, (g+1)*N/G) (contiguous keeps each worker's host traffic
// sequential and the results in order), the fixed operand is uploaded to the first device once and
// broadcast to the others with NCCL (NVLink / NVSwitch), and every device DMAs its results straight
// into the caller's results[i].  No result exchange, no collective on the data path
// (SURVEY.md section 8(e)).  The reference has no multi-GPU code at all: its only affordance is
// count_available_gpus(), src/cudamatrix.cc:14-18, which nothing calls.
#pragma once
#include <dlfcn.h>
#include <nccl.h>   // types and prototypes only: the library is opened at run time (see NcclApi)

namespace {

// ------------------------------------------------------------------------------------------
// NCCL, bound at run time.  No link-time dependency: the product library keeps loading on a box
// without NCCL (and its exported symbols can be checked without a GPU), and a host process that
// has already loaded an NCCL of its own (PyTorch bundles one under the same soname) shares that
// copy instead of getting a second one.
// ------------------------------------------------------------------------------------------
struct NcclApi {
  void *handle = nullptr;
  decltype(&ncclGetVersion) GetVersion = nullptr;
  decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
  decltype(&ncclCommInitRankConfig) CommInitRankConfig = nullptr;
  decltype(&ncclCommInitRank) CommInitRank = nullptr;
  decltype(&ncclCommDestroy) CommDestroy = nullptr;
  decltype(&ncclBroadcast) Broadcast = nullptr;
  decltype(&ncclGroupStart) GroupStart = nullptr;
  decltype(&ncclGroupEnd) GroupEnd = nullptr;
  decltype(&ncclGetErrorString) GetErrorString = nullptr;
  std::string why;   // why it is unavailable
};

NcclApi *nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    const char *forced = std::getenv("EC_NCCL_LIB");   // when set, the only candidate
    if (forced && *forced) {
      api.handle = dlopen(forced, RTLD_NOW | RTLD_GLOBAL);
    } else {
      for (const char *n : names)
        if (!api.handle) api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    }
    if (!api.handle) {
      const char *why = dlerror();   // a second dlerror() call returns NULL: read it once
      api.why = std::string("cannot load ") + (forced && *forced ? forced : "libnccl.so.2") + ": " +
                (why ? why : "not found");
      return;
    }
    bool ok = true;
    auto sym = [&](const char *name) {
      void *s = dlsym(api.handle, name);
      if (!s) {
        ok = false;
        api.why = std::string("libnccl lacks ") + name;
      }
      return s;
    };
    api.GetVersion = (decltype(api.GetVersion))sym("ncclGetVersion");
    api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
    api.CommInitRankConfig = (decltype(api.CommInitRankConfig))sym("ncclCommInitRankConfig");
    api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
    api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
    api.Broadcast = (decltype(api.Broadcast))sym("ncclBroadcast");
    api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
    api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
    api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
    if (!ok) {
      dlclose(api.handle);
      api.handle = nullptr;
    }
  });
  return &api;   // check ->handle; ->why says what went wrong
}

#define EC_NCCL(api, call)                                                                      \
  do {                                                                                          \
    ncclResult_t _r = (call);                                                                   \
    if (_r != ncclSuccess)                                                                      \
      return fail(EC_ERR_CUDA, "%s failed: %s (%s:%d)", #call, (api)->GetErrorString(_r),       \
                  __FILE__, __LINE__);                                                          \
  } while (0)

// One communicator per device over `devices`, created from this one process (the
// ncclCommInitAll pattern, spelled out so that a config can be passed: max_ctas caps the SMs
// NCCL's kernels may hold, which matters when they run underneath a GEMM -- see dgemm_2d.inl).
int nccl_init_all(NcclApi *api, const std::vector<int> &devices, int max_ctas,
                  std::vector<ncclComm_t> *comms) {
  const int n = (int)devices.size();
  comms->assign(n, nullptr);
  ncclUniqueId id;
  EC_NCCL(api, api->GetUniqueId(&id));
  EC_NCCL(api, api->GroupStart());
  for (int i = 0; i < n; ++i) {
    cudaSetDevice(devices[i]);
    ncclResult_t r;
    if (max_ctas > 0) {
      ncclConfig_t cfg = NCCL_CONFIG_INITIALIZER;
      cfg.maxCTAs = max_ctas;
      cfg.minCTAs = 1;
      r = api->CommInitRankConfig(&(*comms)[i], n, id, i, &cfg);
    } else {
      r = api->CommInitRank(&(*comms)[i], n, id, i);
    }
    if (r != ncclSuccess) {
      api->GroupEnd();
      return fail(EC_ERR_CUDA, "ncclCommInitRank(%d of %d, device %d) failed: %s", i, n, devices[i],
                  api->GetErrorString(r));
    }
  }
  EC_NCCL(api, api->GroupEnd());
  return EC_OK;
}

}  // namespace

// ------------------------------------------------------------------------------------------
// per-parent-pipeline state of the scheduler
// ------------------------------------------------------------------------------------------
struct MultiGpu {
  std::vector<int> devices;
  std::vector<ec_pipeline *> pipes;     // one child pipeline (stream + engine + staging) per device
  std::vector<ncclComm_t> comms;        // empty for a single device
  std::vector<double *> d_fixed;        // the fixed operand on every device
  size_t fixed_bytes = 0;
  int64_t broadcasts = 0;               // NCCL broadcasts issued (one per job)

  int ensure(const std::vector<int> &devs) {
    if (devs == devices && !pipes.empty()) return EC_OK;
    destroy();
    int count = 0;
    EC_CUDA(cudaGetDeviceCount(&count));
    for (size_t i = 0; i < devs.size(); ++i) {
      if (devs[i] < 0 || devs[i] >= count)
        return fail(EC_ERR_ARG, "device %d out of range (%d devices)", devs[i], count);
      for (size_t j = 0; j < i; ++j)
        if (devs[j] == devs[i]) return fail(EC_ERR_ARG, "device %d listed twice", devs[i]);
    }
    devices = devs;
    for (int d : devices) {
      ec_pipeline *child = nullptr;
      int rc = ec_pipeline_create(d, &child);
      if (rc) {
        destroy();
        return rc;
      }
      pipes.push_back(child);
    }
    d_fixed.assign(devices.size(), nullptr);
    if (devices.size() > 1) {
      NcclApi *api = nccl_api();
      if (!api->handle) {
        destroy();
        return fail(EC_ERR_UNSUPPORTED, "multi-GPU tensor stream needs NCCL for the broadcast of the fixed operand: %s",
                    api->why.c_str());
      }
      int prev = 0;
      cudaGetDevice(&prev);
      int rc = nccl_init_all(api, devices, /*max_ctas=*/0, &comms);
      cudaSetDevice(prev);
      if (rc) {
        std::string keep = g_last_error;
        destroy();
        g_last_error = keep;
        return rc;
      }
    }
    return EC_OK;
  }

  int ensure_fixed(size_t bytes) {
    if (bytes <= fixed_bytes) return EC_OK;
    for (size_t i = 0; i < devices.size(); ++i) {
      DeviceGuard guard(devices[i]);
      if (d_fixed[i]) EC_CUDA(cudaFree(d_fixed[i]));
      d_fixed[i] = nullptr;
    }
    fixed_bytes = 0;
    for (size_t i = 0; i < devices.size(); ++i) {
      DeviceGuard guard(devices[i]);
      EC_CUDA(cudaMalloc((void **)&d_fixed[i], bytes));
    }
    fixed_bytes = bytes;
    return EC_OK;
  }

  void destroy() {
    if (!comms.empty()) {
      NcclApi *api = nccl_api();
      for (ncclComm_t c : comms)
        if (c && api->handle) api->CommDestroy(c);
      comms.clear();
    }
    for (size_t i = 0; i < d_fixed.size(); ++i)
      if (d_fixed[i]) {
        DeviceGuard guard(devices[i]);
        cudaFree(d_fixed[i]);
      }
    d_fixed.clear();
    fixed_bytes = 0;
    for (ec_pipeline *c : pipes) ec_pipeline_destroy(c);
    pipes.clear();
    devices.clear();
  }
};

namespace {

struct DevicePick {
  std::vector<int> from, picked;
  double gbs_picked = 0.0, gbs_all = 0.0;
};
int pick_stream_devices_impl(const std::vector<int> &devs, DevicePick *out);
int device_list_or_all(const int *devices, int ndev, std::vector<int> *devs);

// Balanced, contiguous, order-preserving partition of range(count) over `parts` workers
// (the same rule as eigencuda_b200/shard.py: shard_range).
inline void shard_range(int64_t count, int parts, int idx, int64_t *lo, int64_t *hi) {
  *lo = count * idx / parts;
  *hi = count * (idx + 1) / parts;
}

int tensor_stream_run_multi_impl(ec_pipeline *p, const int *devices, int ndev, int kind, const double *F,
                                 int64_t f_rows, int64_t f_cols, const double *const *tensor,
                                 double *const *results, int64_t count, int64_t t_rows, int64_t t_cols) {
  if (!p) return fail(EC_ERR_ARG, "null pipeline");
  if (ndev < EC_DEVICES_AUTO || (ndev > 0 && !devices)) return fail(EC_ERR_ARG, "bad device list");
  if (count < 0 || f_rows < 0 || f_cols < 0 || t_rows < 0 || t_cols < 0)
    return fail(EC_ERR_ARG, "negative size");
  StreamShape sh;
  int rc = stream_shape(kind, f_rows, f_cols, t_rows, t_cols, &sh);   // before any device work
  if (rc) return rc;
  if (count == 0) return EC_OK;
  if (!F || !tensor || !results) return fail(EC_ERR_ARG, "null pointer");
  std::vector<int> devs;
  if ((rc = device_list_or_all(devices, std::max(ndev, 0), &devs))) return rc;
  if (ndev == EC_DEVICES_AUTO) {   // the subset of the box's devices that streams fastest, by measurement
    DevicePick pick;
    if ((rc = pick_stream_devices_impl(devs, &pick))) return rc;
    devs = pick.picked;
  }
  const int64_t f_sz = f_rows * f_cols;
  if (t_rows * t_cols == 0 || sh.r_rows * sh.r_cols == 0 || f_sz == 0)   // degenerate: handled on the host
    return tensor_stream_run_impl(p, kind, F, f_rows, f_cols, tensor, results, count, t_rows, t_cols);
  TraceRange trace("ec_tensor_stream_run_multi");
  if (!p->multi) p->multi = new MultiGpu();
  MultiGpu &M = *p->multi;
  if ((rc = M.ensure(devs))) return rc;
  const int G = (int)devs.size();
  if ((rc = M.ensure_fixed((size_t)f_sz * 8))) return rc;

  // ---- fixed operand: onto the first device once, then one NCCL broadcast ----
  {
    ec_pipeline *p0 = M.pipes[0];
    DeviceGuard guard(p0->device);
    if (is_device_ptr(F) || is_pinned_host(F)) {
      EC_CUDA(cudaMemcpyAsync(M.d_fixed[0], F, (size_t)f_sz * 8, cudaMemcpyDefault, p0->stream));
    } else {
      int slot = -1;
      if ((rc = p0->staging.acquire((size_t)f_sz * 8, &slot))) return rc;
      std::memcpy(p0->staging.buf[slot], F, (size_t)f_sz * 8);
      EC_CUDA(cudaMemcpyAsync(M.d_fixed[0], p0->staging.buf[slot], (size_t)f_sz * 8,
                              cudaMemcpyHostToDevice, p0->stream));
      if ((rc = p0->staging.release_after(slot, p0->stream))) return rc;
    }
  }
  if (G > 1) {
    NcclApi *api = nccl_api();
    int prev = 0;
    cudaGetDevice(&prev);
    ncclResult_t r = api->GroupStart();
    for (int g = 0; g < G && r == ncclSuccess; ++g) {
      cudaSetDevice(devs[g]);
      r = api->Broadcast(M.d_fixed[0], M.d_fixed[g], (size_t)f_sz, ncclDouble, /*root=*/0, M.comms[g],
                         M.pipes[g]->stream);
    }
    const ncclResult_t r2 = api->GroupEnd();
    cudaSetDevice(prev);
    if (r != ncclSuccess || r2 != ncclSuccess)
      return fail(EC_ERR_CUDA, "ncclBroadcast of the fixed operand failed: %s",
                  api->GetErrorString(r != ncclSuccess ? r : r2));
    M.broadcasts++;
  }

  // ---- one worker per device over its shard; engines share the host's cores ----
  int threads = p->cfg_threads;
  if (threads <= 0) {
    const unsigned hw = std::max(2u, std::thread::hardware_concurrency());
    threads = (int)std::max(2u, std::min(16u, hw / (unsigned)G));
  }
  std::vector<int> rcs(G, EC_OK);
  std::vector<std::string> errs(G);
  int64_t launches_before = 0;
  for (ec_pipeline *c : M.pipes) launches_before += c->launches;
  auto work = [&](int g) {
    int64_t lo, hi;
    shard_range(count, G, g, &lo, &hi);
    ec_pipeline *c = M.pipes[g];
    c->cfg_chunk = p->cfg_chunk;
    c->cfg_ring = p->cfg_ring;
    c->cfg_threads = threads;
    c->backend = p->backend;
    cudaSetDevice(c->device);
    if (hi > lo)
      rcs[g] = tensor_stream_run_impl(c, kind, M.d_fixed[g], f_rows, f_cols, tensor + lo, results + lo,
                                      hi - lo, t_rows, t_cols);
    else
      rcs[g] = cudaStreamSynchronize(c->stream) == cudaSuccess ? EC_OK : EC_ERR_CUDA;   // drains the broadcast
    if (rcs[g]) errs[g] = g_last_error;
  };
  std::vector<std::thread> workers;
  for (int g = 1; g < G; ++g) workers.emplace_back(work, g);
  {
    int prev = 0;
    cudaGetDevice(&prev);
    work(0);   // the calling thread drives the first device
    cudaSetDevice(prev);
  }
  for (std::thread &t : workers) t.join();
  for (ec_pipeline *c : M.pipes) p->launches += c->launches;
  p->launches -= launches_before;
  p->last_kernel = M.pipes[0]->last_kernel;
  for (int g = 0; g < G; ++g)
    if (rcs[g]) {
      g_last_error = "device " + std::to_string(devs[g]) + ": " + errs[g];
      return rcs[g];
    }
  return EC_OK;
}

}  // namespace

// ------------------------------------------------------------------------------------------
// Host-link ceiling, measured (the roofline of the streamed configs is the
// host link, "L = measured pinned cudaMemcpyAsync bandwidth").  Plain pinned copies on every
// listed device AT ONCE -- what the box gives however the engine is written.
// ------------------------------------------------------------------------------------------
namespace {

// Buffers and streams for plain pinned copies on a set of devices; run() times one mode on a subset.
struct LinkProbe {
  struct Dev {
    int device = -1;
    void *d_in = nullptr, *d_out = nullptr, *h_in = nullptr, *h_out = nullptr;
    cudaStream_t s_in = nullptr, s_out = nullptr;
  };
  std::vector<Dev> D;
  size_t bytes = 0;
  int prev = 0;

  int setup(const std::vector<int> &devs, size_t bytes_) {
    bytes = bytes_;
    cudaGetDevice(&prev);
    D.resize(devs.size());
    for (size_t g = 0; g < devs.size(); ++g) {
      D[g].device = devs[g];
      EC_CUDA(cudaSetDevice(devs[g]));
      EC_CUDA(cudaMalloc(&D[g].d_in, bytes));
      EC_CUDA(cudaMalloc(&D[g].d_out, bytes));
      EC_CUDA(cudaHostAlloc(&D[g].h_in, bytes, cudaHostAllocPortable));
      EC_CUDA(cudaHostAlloc(&D[g].h_out, bytes, cudaHostAllocPortable));
      std::memset(D[g].h_in, 1, bytes);
      std::memset(D[g].h_out, 0, bytes);
      EC_CUDA(cudaMemsetAsync(D[g].d_out, 0, bytes, 0));
      EC_CUDA(cudaStreamCreateWithFlags(&D[g].s_in, cudaStreamNonBlocking));
      EC_CUDA(cudaStreamCreateWithFlags(&D[g].s_out, cudaStreamNonBlocking));
      EC_CUDA(cudaDeviceSynchronize());
    }
    return EC_OK;
  }
  // mode 0: H2D only, 1: D2H only, 2: both at once.  `use[g]` selects the devices that take part.
  // Result: GB/s in each direction, summed over the participating devices, best of `reps`
  // (after one warm-up repetition).
  int run(int mode, const std::vector<char> &use, int reps, double *gbs) {
    constexpr int INNER = 2;   // copies per direction per device per timed repetition
    double best = 1e300;
    int active = 0;
    for (size_t g = 0; g < D.size(); ++g) active += use[g] ? 1 : 0;
    for (int rep = 0; rep < reps + 1; ++rep) {
      const auto t0 = std::chrono::steady_clock::now();
      for (int it = 0; it < INNER; ++it)
        for (size_t g = 0; g < D.size(); ++g) {
          if (!use[g]) continue;
          if (mode != 1) EC_CUDA(cudaMemcpyAsync(D[g].d_in, D[g].h_in, bytes, cudaMemcpyHostToDevice, D[g].s_in));
          if (mode != 0) EC_CUDA(cudaMemcpyAsync(D[g].h_out, D[g].d_out, bytes, cudaMemcpyDeviceToHost, D[g].s_out));
        }
      for (size_t g = 0; g < D.size(); ++g) {
        if (!use[g]) continue;
        EC_CUDA(cudaStreamSynchronize(D[g].s_in));
        EC_CUDA(cudaStreamSynchronize(D[g].s_out));
      }
      const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      if (rep > 0) best = std::min(best, dt);
    }
    *gbs = (double)active * INNER * (double)bytes / best / 1e9;
    return EC_OK;
  }
  ~LinkProbe() {
    for (Dev &d : D) {
      if (d.device < 0) continue;
      cudaSetDevice(d.device);
      if (d.s_in) cudaStreamDestroy(d.s_in);
      if (d.s_out) cudaStreamDestroy(d.s_out);
      if (d.d_in) cudaFree(d.d_in);
      if (d.d_out) cudaFree(d.d_out);
      if (d.h_in) cudaFreeHost(d.h_in);
      if (d.h_out) cudaFreeHost(d.h_out);
    }
    cudaSetDevice(prev);
  }
};

// The same measurement on the CALLER's pinned buffers: device g streams its slice
// [g * bytes_per_device, (g+1) * bytes_per_device) of `host_in` to the device and a device buffer
// back into its slice of `host_out`, `piece` bytes per copy.  Unlike the small rotating buffers of
// LinkProbe this touches as much unique host memory as the real job does, so the last-level cache of
// the host (which DMA reads can hit and DMA writes can land in) does not flatter the figure: this is
// the ceiling of the tensor stream itself -- its traffic without the products.
int host_link_probe_user_impl(const std::vector<int> &devs, const char *host_in, char *host_out,
                              size_t bytes_per_device, size_t piece, int reps, double *out_gbs) {
  const int G = (int)devs.size();
  struct Dev {
    void *d_in = nullptr, *d_out = nullptr;
    cudaStream_t s_in = nullptr, s_out = nullptr;
  };
  std::vector<Dev> D(G);
  int prev = 0;
  cudaGetDevice(&prev);
  auto cleanup = [&] {
    for (int g = 0; g < G; ++g) {
      cudaSetDevice(devs[g]);
      if (D[g].s_in) cudaStreamDestroy(D[g].s_in);
      if (D[g].s_out) cudaStreamDestroy(D[g].s_out);
      if (D[g].d_in) cudaFree(D[g].d_in);
      if (D[g].d_out) cudaFree(D[g].d_out);
    }
    cudaSetDevice(prev);
  };
  int rc = EC_OK;
#define PROBE_CUDA(call)                                                                          \
  do {                                                                                            \
    cudaError_t _e = (call);                                                                      \
    if (_e != cudaSuccess) {                                                                      \
      rc = fail(EC_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
      cleanup();                                                                                  \
      return rc;                                                                                  \
    }                                                                                             \
  } while (0)
  for (int g = 0; g < G; ++g) {
    PROBE_CUDA(cudaSetDevice(devs[g]));
    PROBE_CUDA(cudaMalloc(&D[g].d_in, piece));
    PROBE_CUDA(cudaMalloc(&D[g].d_out, piece));
    PROBE_CUDA(cudaMemsetAsync(D[g].d_out, 0, piece, 0));
    PROBE_CUDA(cudaStreamCreateWithFlags(&D[g].s_in, cudaStreamNonBlocking));
    PROBE_CUDA(cudaStreamCreateWithFlags(&D[g].s_out, cudaStreamNonBlocking));
    PROBE_CUDA(cudaDeviceSynchronize());
  }
  for (int mode = 0; mode < 3; ++mode) {
    double best = 1e300;
    for (int rep = 0; rep < reps; ++rep) {
      const auto t0 = std::chrono::steady_clock::now();
      for (size_t off = 0; off < bytes_per_device; off += piece) {
        const size_t len = std::min(piece, bytes_per_device - off);
        for (int g = 0; g < G; ++g) {
          const size_t at = (size_t)g * bytes_per_device + off;
          PROBE_CUDA(cudaSetDevice(devs[g]));
          if (mode != 1) PROBE_CUDA(cudaMemcpyAsync(D[g].d_in, host_in + at, len, cudaMemcpyHostToDevice, D[g].s_in));
          if (mode != 0) PROBE_CUDA(cudaMemcpyAsync(host_out + at, D[g].d_out, len, cudaMemcpyDeviceToHost, D[g].s_out));
        }
      }
      for (int g = 0; g < G; ++g) {
        PROBE_CUDA(cudaStreamSynchronize(D[g].s_in));
        PROBE_CUDA(cudaStreamSynchronize(D[g].s_out));
      }
      const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      best = std::min(best, dt);
    }
    out_gbs[mode] = (double)G * (double)bytes_per_device / best / 1e9;
  }
#undef PROBE_CUDA
  cleanup();
  return EC_OK;
}

int device_list_or_all(const int *devices, int ndev, std::vector<int> *devs) {
  if (ndev < 0 || (ndev > 0 && !devices)) return fail(EC_ERR_ARG, "bad device list");
  if (ndev == 0)
    for (int d = 0; d < (int)ec_device_count(); ++d) devs->push_back(d);
  else
    devs->assign(devices, devices + ndev);
  if (devs->empty()) return fail(EC_ERR_CUDA, "no CUDA device visible");
  return EC_OK;
}

int host_link_probe_impl(const int *devices, int ndev, size_t bytes, int reps, double *out_gbs) {
  if (!out_gbs) return fail(EC_ERR_ARG, "null out pointer");
  out_gbs[0] = out_gbs[1] = out_gbs[2] = 0.0;
  if (bytes == 0 || reps <= 0) return fail(EC_ERR_ARG, "bad probe arguments");
  std::vector<int> devs;
  int rc = device_list_or_all(devices, ndev, &devs);
  if (rc) return rc;
  LinkProbe P;
  if ((rc = P.setup(devs, bytes))) return rc;
  const std::vector<char> all(devs.size(), 1);
  for (int mode = 0; mode < 3; ++mode)
    if ((rc = P.run(mode, all, reps, &out_gbs[mode]))) return rc;
  return EC_OK;
}

// ------------------------------------------------------------------------------------------
// Which of the box's GPUs to stream through.  The streamed job is bound by the HOST side of the
// links, and that side is not uniform: on this pool's 8-GPU boxes GPUs 0-3 hang off a path that
// gives 51 GB/s each way however many of them transfer, GPUs 4-7 off one that gives 95, and all
// eight together reach only 64 -- the narrow half drags the wide half down
// (profiles/r02_link_probe.txt).  More GPUs is then less throughput.  The choice is made by
// measurement, once per process and device list: bidirectional pinned copies (64 MiB) on all
// devices, on each pair of neighbours, then greedily on unions of pairs from the fastest down,
// keeping a pair only if the aggregate grows by more than 5 %.
// ------------------------------------------------------------------------------------------
std::mutex g_pick_mu;
std::vector<DevicePick> g_picks;

int pick_stream_devices_impl(const std::vector<int> &devs, DevicePick *out) {
  {
    std::lock_guard<std::mutex> lk(g_pick_mu);
    for (const DevicePick &p : g_picks)
      if (p.from == devs) {
        *out = p;
        return EC_OK;
      }
  }
  DevicePick pick;
  pick.from = devs;
  const int n = (int)devs.size();
  LinkProbe P;
  int rc = P.setup(devs, (size_t)64 << 20);
  if (rc) return rc;
  std::vector<char> use(n, 1);
  if ((rc = P.run(2, use, 2, &pick.gbs_all))) return rc;
  pick.picked = devs;
  pick.gbs_picked = pick.gbs_all;
  if (n > 2) {
    struct Group { int lo, hi; double gbs; };
    std::vector<Group> groups;
    for (int lo = 0; lo < n; lo += 2) {
      Group g{lo, std::min(lo + 2, n), 0.0};
      std::fill(use.begin(), use.end(), 0);
      for (int i = g.lo; i < g.hi; ++i) use[i] = 1;
      if ((rc = P.run(2, use, 2, &g.gbs))) return rc;
      groups.push_back(g);
    }
    std::sort(groups.begin(), groups.end(), [](const Group &a, const Group &b) { return a.gbs > b.gbs; });
    std::vector<char> best(n, 0);
    for (int i = groups[0].lo; i < groups[0].hi; ++i) best[i] = 1;
    double best_gbs = groups[0].gbs;
    for (size_t k = 1; k < groups.size(); ++k) {
      std::vector<char> trial = best;
      for (int i = groups[k].lo; i < groups[k].hi; ++i) trial[i] = 1;
      double gbs = 0.0;
      if ((rc = P.run(2, trial, 2, &gbs))) return rc;
      if (gbs > 1.05 * best_gbs) {
        best = trial;
        best_gbs = gbs;
      }
    }
    if (best_gbs > 1.05 * pick.gbs_all) {   // otherwise: every device (more GEMM capacity, same link)
      pick.picked.clear();
      for (int i = 0; i < n; ++i)
        if (best[i]) pick.picked.push_back(devs[i]);
      pick.gbs_picked = best_gbs;
    }
  }
  {
    std::lock_guard<std::mutex> lk(g_pick_mu);
    g_picks.push_back(pick);
  }
  *out = pick;
  return EC_OK;
}

}  // namespace
