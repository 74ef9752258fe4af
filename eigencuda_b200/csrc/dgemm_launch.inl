// eigencuda_b200/csrc/dgemm_launch.inl -- native FP64 path: tensor maps, stream-K workspace, kernel launch, tile / schedule cost model
// Part of the single translation unit ec_api.cu (included there, in order); not a standalone header.
#pragma once

namespace {

// ------------------------------------------------------------------------------------------
// kernel dispatch
// ------------------------------------------------------------------------------------------
struct Operand {
  const double *ptr;
  int64_t ld, stride;
  bool kmajor;     // contiguous along k
  int64_t mn, k;   // extents of op(X): mn x k
};

bool tma_ok(const Operand &o, int64_t batch) {
  if (((uintptr_t)o.ptr & 15u) != 0) return false;
  if ((o.ld & 1) != 0) return false;                       // row pitch must be a multiple of 16 B
  if (batch > 1 && o.stride != 0 && (o.stride & 1) != 0) return false;
  if (o.mn >= ((int64_t)1 << 31) || o.k >= ((int64_t)1 << 31)) return false;
  if (o.ld * 8 >= ((int64_t)1 << 40) || o.stride * 8 >= ((int64_t)1 << 40)) return false;
  return true;
}

int make_tensor_map(ec_pipeline *p, CUtensorMap *tm, const Operand &o, int64_t batch, int tile_extent) {
  const bool batched = batch > 1 && o.stride != 0;
  const int64_t inner = o.kmajor ? o.k : o.mn, outer = o.kmajor ? o.mn : o.k;
  const int box_outer = o.kmajor ? tile_extent : ec::BK;
  const int64_t nb = batched ? batch : 1, st = batched ? o.stride : 0;
  // cache lookup
  uint64_t hsh = (uint64_t)(uintptr_t)o.ptr * 0x9E3779B97F4A7C15ull ^ (uint64_t)inner * 0xBF58476D1CE4E5B9ull ^
                 (uint64_t)outer * 0x94D049BB133111EBull ^ (uint64_t)o.ld * 31 ^ (uint64_t)box_outer;
  ec_pipeline::TmapEntry &e = p->tmap_cache[(hsh >> 40) & 15];
  if (e.ptr == o.ptr && e.inner == inner && e.outer == outer && e.batch == nb && e.ld == o.ld && e.stride == st &&
      e.box_outer == box_outer) {
    *tm = e.map;
    return EC_OK;
  }
  encode_tiled_fn enc = get_encode_tiled();
  if (!enc) return fail(EC_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t dims[3], strides[2];
  cuuint32_t box[3], estr[3] = {1, 1, 1};
  dims[0] = (cuuint64_t)inner;
  dims[1] = (cuuint64_t)outer;
  dims[2] = (cuuint64_t)nb;
  strides[0] = (cuuint64_t)o.ld * 8;
  // stride of the (possibly degenerate) batch dimension: any legal multiple of 16 when unused
  strides[1] = batched ? (cuuint64_t)o.stride * 8 : (cuuint64_t)o.ld * 8 * (cuuint64_t)std::max<int64_t>(outer, 1);
  box[0] = 16;
  box[1] = (cuuint32_t)box_outer;
  box[2] = 1;
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, (void *)o.ptr, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(EC_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  e.ptr = o.ptr; e.inner = inner; e.outer = outer; e.batch = nb; e.ld = o.ld; e.stride = st; e.box_outer = box_outer;
  e.map = *tm;
  return EC_OK;
}

// >= 2 * grid * BM*BN*8 for every config: one contributor slot and one owner spill slot per CTA
constexpr size_t SK_WORKSPACE_BYTES = (size_t)2 * 640 * 128 * 128 * 8 / 4;
constexpr int SK_MAX_GRID = 1024;
constexpr unsigned CONTENDED_SKEW_NS = 50000;   // CTAs of one persistent grid starting > 50 us apart: somebody else holds SMs
constexpr int CONTENDED_LAUNCHES = 32;          // launches that avoid static partitions after that, before probing again

int ensure_sk_workspace(ec_pipeline *p) {
  if (p->sk_workspace) return EC_OK;
  EC_CUDA(cudaMalloc((void **)&p->sk_workspace, SK_WORKSPACE_BYTES));
  EC_CUDA(cudaMalloc((void **)&p->sk_flags, SK_MAX_GRID * sizeof(unsigned)));
  EC_CUDA(cudaMemsetAsync(p->sk_flags, 0, SK_MAX_GRID * sizeof(unsigned), p->stream));
  return EC_OK;
}

// Scheduler words of a pipeline: the claim counter of the dynamic tile schedule, the start-skew
// statistics of persistent grids (device memory) and the word the kernel reports the skew in
// (host memory mapped into the device, read by the dispatcher without any synchronisation).
int ensure_sched_words(ec_pipeline *p) {
  if (p->sched_words) return EC_OK;
  EC_CUDA(cudaMalloc((void **)&p->sched_words, 4 * sizeof(unsigned long long)));
  const unsigned long long init[4] = {0ull, ~0ull, ~0ull, 0ull};   // claim counter, earliest CTA start (two slots, by launch parity)
  EC_CUDA(cudaMemcpyAsync(p->sched_words, init, sizeof(init), cudaMemcpyHostToDevice, p->stream));
  EC_CUDA(cudaStreamSynchronize(p->stream));   // `init` is on this stack frame
  EC_CUDA(cudaHostAlloc((void **)&p->skew_host, sizeof(unsigned), cudaHostAllocMapped));
  *p->skew_host = 0;
  EC_CUDA(cudaHostGetDevicePointer((void **)&p->skew_dev, (void *)p->skew_host, 0));
  return EC_OK;
}

// Resident CTAs per SM of one kernel instantiation on one device (cached).
template <class Cfg, bool AK, bool BK_, bool EDGE>
auto pick_tma_dmma_kernel() {
  if constexpr (EDGE) return ec::dgemm_tma_dmma_edge_kernel<Cfg, AK, BK_>;
  else return ec::dgemm_tma_dmma_kernel<Cfg, AK, BK_>;
}

template <class Cfg, bool AK, bool BK_, bool EDGE>
int ctas_per_sm(ec_pipeline *p, int *out) {
  auto kern = pick_tma_dmma_kernel<Cfg, AK, BK_, EDGE>();
  static std::atomic<int> cached[64];
  int v = cached[p->device & 63].load();
  if (v == 0) {
    EC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    int occ = 0;
    EC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, Cfg::THREADS, Cfg::SMEM_BYTES));
    if (occ < 1) return fail(EC_ERR_CUDA, "kernel does not fit on an SM (occupancy 0)");
    v = std::min(occ, Cfg::MIN_CTAS);
    cached[p->device & 63].store(v);
  }
  *out = v;
  return EC_OK;
}

template <class Cfg, bool AK, bool BK_, bool EDGE = false>
int launch_tma_dmma(ec_pipeline *p, const CUtensorMap &tmA, const CUtensorMap &tmB,
                    ec::GemmParams &gp, bool want_sk) {
  // 128x128 tile and a C with partial tiles: the instantiation whose fused turn-around takes edge tiles too
  if constexpr (Cfg::FUSE_EPI && !EDGE) {
    if (gp.fuse_epilogue_wanted && gp.beta == 0.0 && (gp.M % Cfg::BM != 0 || gp.N % Cfg::BN != 0))
      return launch_tma_dmma<Cfg, AK, BK_, true>(p, tmA, tmB, gp, want_sk);
  }
  auto kern = pick_tma_dmma_kernel<Cfg, AK, BK_, EDGE>();
  int ctas = 1;
  int rc = ctas_per_sm<Cfg, AK, BK_, EDGE>(p, &ctas);
  if (rc) return rc;
  const int64_t slots = (int64_t)(p->sm_count - p->sm_margin) * ctas;
  const int64_t kblocks = (gp.K + ec::BK - 1) / ec::BK;
  int grid = (int)std::min<int64_t>(gp.total_tiles, slots);
  gp.sk_tiles = 0;
  gp.sk_sms = p->sm_count - p->sm_margin;
  gp.sk_workspace = nullptr;
  gp.sk_flags = nullptr;
  gp.sk_epoch = 0;
  gp.pdl_trigger = p->pdl;
  // One-k-iteration tiles whose operands stream from HBM keep the plain epilogue: deferred stores go out only once the NEXT
  // tile's first (and only) stage has landed, and that wait is then the whole tile (2^20 x 128 x 16: 5.07 TB/s plain, 4.24 fused;
  // with L2-resident operands -- 8192 x 8192 x 16 -- fused wins like everywhere else: tools/fuse_k_probe.py,
  // profiles/r02_fuse_k_probe.txt).
  const bool operands_in_l2 = (double)(gp.M + gp.N) * (double)gp.K * 8.0 * (double)std::max<int64_t>(1, gp.batch) <= 48e6;
  gp.fuse_epilogue = (p->fuse_epilogue && (kblocks >= p->fuse_min_kblocks || operands_in_l2)) ? 1 : 0;
  // Is the GPU being shared?  The last persistent grid reported how far apart its CTAs started:
  // microseconds on an idle GPU, as long as the other kernel runs when some of them had to wait
  // for SMs it held.  A statically cut stream-K region then runs at the pace of the CTA that
  // started last, so for the next launches the region is dropped and the work is handed out
  // dynamically (long tiles: one CTA per tile, balanced by the hardware).  Re-probed afterwards.
  if (p->skew_host && *(volatile unsigned *)p->skew_host > CONTENDED_SKEW_NS) {
    p->contended_left = CONTENDED_LAUNCHES;
    p->contended_events++;
    *(volatile unsigned *)p->skew_host = 0;
  }
  const bool contended = p->contended_left > 0;
  if (contended) p->contended_left--;
  // Stream-K region: the tiles of the last, partial wave plus (when there is one) one full wave,
  // so that every CTA gets between one and two tiles' worth of k-iterations in the region.
  const int64_t rem = gp.total_tiles % slots;
  if (want_sk && !contended && p->stream_k && (rem != 0 || getenv("EC_SK_ALL")) && slots <= SK_MAX_GRID &&
      2 * slots * (int64_t)(Cfg::BM * Cfg::BN * 8) <= (int64_t)SK_WORKSPACE_BYTES) {
    int64_t sk_tiles = rem + (gp.total_tiles > slots ? slots : 0);
    if (const char *e = getenv("EC_SK_ALL")) {   // experiment: stream-K over every tile
      if (atoi(e)) sk_tiles = gp.total_tiles;
    }
    if (sk_tiles * kblocks >= 4 * slots) {   // at least 4 k-iterations per CTA in the region
      rc = ensure_sk_workspace(p);
      if (rc) return rc;
      gp.sk_tiles = sk_tiles;
      gp.sk_workspace = p->sk_workspace;
      gp.sk_flags = p->sk_flags;
      gp.sk_epoch = ++p->sk_epoch;
      if (gp.sk_epoch == 0) gp.sk_epoch = ++p->sk_epoch;   // 0 is the cleared state
      grid = (int)slots;
    }
  }
  if (gp.sk_tiles == 0 && (p->one_cta_per_tile || contended) && kblocks >= 64 && gp.total_tiles > slots)
    grid = (int)std::min<int64_t>(gp.total_tiles, (int64_t)1 << 30);   // one tile per CTA
  // Whole tiles beyond each CTA's first are claimed dynamically when tiles are long enough for
  // the claim's round trip to hide behind the loads of the tile before (8 k-iterations).
  const int64_t dp_tiles = gp.total_tiles - gp.sk_tiles;
  gp.dp_counter = nullptr;
  gp.dp_base = 0;
  gp.dp_all_dynamic = 0;
  gp.stats = nullptr;
  gp.skew_out = nullptr;
  gp.skew_limit_ns = 0;
  gp.launch_parity = 0;
  const bool persistent = (int64_t)grid <= slots;
  if (persistent && p->dynamic_tiles) {
    rc = ensure_sched_words(p);
    if (rc) return rc;
    if (contended && dp_tiles > 0) {   // every tile claimed, the first one of each CTA included
      gp.dp_counter = p->sched_words;
      gp.dp_base = p->dp_claims;
      gp.dp_all_dynamic = 1;
      p->dp_claims += (unsigned long long)dp_tiles + (unsigned long long)grid;   // successes + one miss per CTA
    } else if (dp_tiles > grid && kblocks >= 8) {
      gp.dp_counter = p->sched_words;
      gp.dp_base = p->dp_claims;
      p->dp_claims += (unsigned long long)dp_tiles;   // (dp_tiles - grid) successes + one miss per CTA
    }
    if (grid >= p->sm_count / 2) {   // a grid wide enough for its start skew to say something about the GPU
      gp.stats = p->sched_words + 1;
      gp.skew_out = p->skew_dev;
      gp.skew_limit_ns = CONTENDED_SKEW_NS;
      gp.launch_parity = (int)(p->stats_launches++ & 1);
    }
  }
  {
    // Programmatic dependent launch: this grid may become resident while its predecessor in the
    // stream is still running; the kernel waits (griddepcontrol.wait) before it touches memory.
    // No cooperative launch any more, stream-K included: contributors never wait, owners wait only
    // at the very end of their work, and the grid never exceeds what the device can hold at once,
    // so every CTA gets an SM as soon as other kernels release theirs.
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(Cfg::THREADS);
    cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
    cfg.stream = p->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    // Only for grids that fill at most half of the resident slots: the dependent grid is placed
    // into whatever slots its predecessor leaves free, and once the two do not both fit that
    // placement is lopsided -- some SMs end up with five CTAs of the new grid, others with none
    // (measured, back to back: n = 512 14.5 -> 13.1 us with PDL, but n = 640 22.7 -> 33.6 us).
    cfg.numAttrs = (p->pdl && (int64_t)grid * 2 <= slots) ? 1 : 0;
    EC_CUDA(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, gp));
  }
  EC_CUDA(cudaGetLastError());
  p->launches++;
  return EC_OK;
}

template <class Cfg>
int run_tma_dmma(ec_pipeline *p, const Operand &oa, const Operand &ob, int64_t m, int64_t n,
                 int64_t k, double alpha, double beta, double *C, int64_t ldc, int64_t stride_c,
                 int64_t stride_a, int64_t stride_b, int64_t batch, bool want_sk,
                 const char *const names[8]) {
  CUtensorMap tmA, tmB;
  int rc = make_tensor_map(p, &tmA, oa, batch, Cfg::BM);
  if (rc) return rc;
  rc = make_tensor_map(p, &tmB, ob, batch, Cfg::BN);
  if (rc) return rc;
  ec::GemmParams gp;
  gp.M = m; gp.N = n; gp.K = k; gp.batch = batch;
  gp.alpha = alpha; gp.beta = beta;
  gp.C = C; gp.ldc = ldc; gp.stride_c = stride_c;
  gp.tiles_m = (int)((m + Cfg::BM - 1) / Cfg::BM);
  gp.tiles_n = (int)((n + Cfg::BN - 1) / Cfg::BN);
  gp.total_tiles = (int64_t)gp.tiles_m * gp.tiles_n * batch;
  gp.a_batched = (batch > 1 && stride_a != 0);
  gp.b_batched = (batch > 1 && stride_b != 0);
  gp.fast_locate = gp.total_tiles < ((int64_t)1 << 31) ? 1 : 0;
  gp.fd_tiles_per_problem = ec::FastDiv::make((uint32_t)std::max(1, gp.tiles_m * gp.tiles_n));
  gp.fd_tiles_m = ec::FastDiv::make((uint32_t)std::max(1, gp.tiles_m));
  gp.fuse_epilogue_wanted = p->fuse_epilogue;
  gp.c_vec2 = (((uintptr_t)C & 15u) == 0 && (ldc & 1) == 0 && (batch <= 1 || (stride_c & 1) == 0)) ? 1 : 0;
  int idx;
  if (oa.kmajor && ob.kmajor) { idx = 0; rc = launch_tma_dmma<Cfg, true, true>(p, tmA, tmB, gp, want_sk); }
  else if (oa.kmajor && !ob.kmajor) { idx = 1; rc = launch_tma_dmma<Cfg, true, false>(p, tmA, tmB, gp, want_sk); }
  else if (!oa.kmajor && ob.kmajor) { idx = 2; rc = launch_tma_dmma<Cfg, false, true>(p, tmA, tmB, gp, want_sk); }
  else { idx = 3; rc = launch_tma_dmma<Cfg, false, false>(p, tmA, tmB, gp, want_sk); }
  p->last_kernel = names[idx + (gp.sk_tiles > 0 ? 4 : 0)];
  return rc;
}

// Tile and schedule choice.  All configurations run the DMMA pipe at nearly the same rate once
// it is fed; what differs is (a) load balance -- under the data-parallel schedule the busiest SM
// does ceil(tiles / SMs) tiles, under stream-K every SM does tiles / SMs of them plus one
// workspace round trip -- and (b) per-tile overhead and main-loop efficiency, which get worse
// as tiles shrink.  Cost model in SM clocks, fitted to profiles/r01_tile_sweep*.json.
enum { TILE_AUTO = 0, TILE_L = 1, TILE_S = 2, TILE_XS = 3, TILE_T = 4, TILE_N = 5 };

struct TileChoice { int tile; bool sk; };

TileChoice choose_tile(const ec_pipeline *p, int64_t m, int64_t n, int64_t k, int64_t batch) {
  int force = 0;
  bool allow_sk = p->stream_k != 0, force_sk = false;
  // EC_FORCE_TILE=1..4 and EC_STREAM_K=0 (never) | 1 (cost model) | 2 (whenever feasible) override
  // the choice for tools/tile_sweep.py and the tests; read per call
  if (const char *e = getenv("EC_FORCE_TILE")) force = atoi(e);
  if (const char *e = getenv("EC_STREAM_K")) {
    allow_sk = atoi(e) != 0;
    force_sk = atoi(e) == 2;
  }
  // eff[r-1]: fraction of the DMMA issue rate the main loop sustains with r CTAs resident per SM
  // (the 4-warp tiles put one warp on each SM sub-partition per CTA and hide their LDS latency
  // badly until a second or third CTA shares the SM); overhead: exposed clocks per tile; fixup:
  // exposed clocks of one stream-K tail (publish partial, wait, read, epilogue).  Fitted to
  // profiles/r01_tile_sweep*.json and profiles/r01_small_probe.txt (back-to-back launches).
  // The 32x32 tile needs 4 flop per operand byte from L2, more than L2 delivers at the full DMMA
  // rate: it tops out near 0.72 however many CTAs share the SM (26-30 TFLOP/s), but it is the only
  // shape that fills the machine for n <= 640 and wastes nothing on matrices of 32 x 32 and less.
  struct Cand { int id, bm, bn, ctas; double eff[6], overhead, fixup; };
  const Cand cands[5] = {{TILE_L, 128, 128, 1, {0.987, 0.987, 0.987, 0.987, 0.987, 0.987}, 2300.0, 30000.0},
                         {TILE_S, 64, 64, 2, {0.59, 0.955, 0.955, 0.955, 0.955, 0.955}, 1200.0, 17000.0},
                         {TILE_XS, 64, 32, 3, {0.50, 0.76, 0.80, 0.80, 0.80, 0.80}, 450.0, 20000.0},
                         {TILE_T, 32, 32, 5, {0.30, 0.58, 0.68, 0.72, 0.72, 0.72}, 250.0, 20000.0},
                         // 128x16: the same MACs per tile as 64x32; only ever cheaper when N <= 16 halves the tile count
                         {TILE_N, 128, 16, 3, {0.50, 0.76, 0.80, 0.80, 0.80, 0.80}, 450.0, 20000.0}};
  const double kpad = (double)((k + 15) / 16 * 16);
  const int64_t kblocks = (k + 15) / 16;
  double best = 1e300;
  TileChoice pick{TILE_L, false};
  for (const Cand &c : cands) {
    if (force >= TILE_L && force <= TILE_N && c.id != force) continue;
    const int64_t tiles = ((m + c.bm - 1) / c.bm) * ((n + c.bn - 1) / c.bn) * batch;
    const int sms = p->sm_count - p->sm_margin;
    const int64_t slots = (int64_t)sms * c.ctas;
    const int64_t resident = std::max<int64_t>(1, std::min<int64_t>(c.ctas, (tiles + sms - 1) / sms));
    const double tile_clk = (double)c.bm * c.bn * kpad / 64.0 / c.eff[resident - 1] + c.overhead;
    const double dp = (double)((tiles + sms - 1) / sms) * tile_clk;
    if (dp < best) { best = dp; pick = {c.id, false}; }
    const int64_t rem = tiles % slots;
    const int64_t sk_tiles = rem + (tiles > slots ? slots : 0);
    if (allow_sk && rem != 0 && sk_tiles * kblocks >= 4 * slots) {
      // every SM does its exact share; one partial tile written and one read per CTA
      // the stream-K grid fills every slot, so the SM runs at its full-occupancy efficiency
      const double tile_clk_sk = (double)c.bm * c.bn * kpad / 64.0 / c.eff[c.ctas - 1] + c.overhead;
      // with far fewer tiles than CTAs a tile is cut into several partials and its owner adds
      // them one after the other: the exposed tail grows with the partials per tile
      // (beyond SK_TREE_MIN partials the kernel sums in two levels: a group of SK_TREE_GROUP,
      // then the group sums)
      const double splits = std::max(1.0, (double)slots / (double)tiles);
      const double chain = splits <= (double)ec::SK_TREE_MIN
                               ? splits
                               : (double)ec::SK_TREE_GROUP + splits / (double)ec::SK_TREE_GROUP;
      const double sk = force_sk ? dp * 1e-3
                                 : (double)tiles / sms * tile_clk_sk + c.fixup * (1.0 + 0.35 * (chain - 1.0));
      if (sk < best) { best = sk; pick = {c.id, true}; }
    }
  }
  return pick;
}



}  // namespace
