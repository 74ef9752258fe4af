// eigencuda_b200/csrc/dgemm_kernels.cuh -- hand-written sm_100a FP64 GEMM kernels.
//
// These kernels replace the one device computation of the reference, the cublasDgemm call in
// CudaPipeline::gemm (reference src/cudapipeline.cc:29-31): C = alpha * op(A) * op(B) + beta * C,
// FP64, column-major.  Nothing here is derived from cuBLAS or from the reference's sources.
//
//   dgemm_tma_dmma_kernel   the hot kernel.  Persistent (one CTA per SM, static tile schedule),
//                           warp-specialised: one producer warp feeds a 5-stage smem ring with
//                           TMA (cp.async.bulk.tensor, 128B swizzle) signalled through
//                           mbarriers; eight consumer warps hold a 128x128 C tile in registers
//                           (64x32 per warp) and issue FP64 DMMA (mma.sync.m8n8k4.f64 -- the only
//                           FP64 tensor shape sm_100a has: every larger PTX shape lowers to
//                           DMMA.8x8x4 in SASS).  All four op(A)/op(B) combinations are served by
//                           two smem layouts per operand ("K-major" and "MN-major"), both read
//                           with conflict-free LDS.128.  Batched over a third tensor-map
//                           dimension for the tensor-stream job (one launch per chunk).
//   dgemm_simt_kernel       generic fallback for what TMA cannot address (odd leading
//                           dimensions, unaligned bases) and for tiny problems such as the
//                           README's 3x2 * 2x2: register-tiled FP64 FMA, bounds-checked.
//   scale_kernel            C = beta * C for the k == 0 / alpha == 0 corner.
//   fill_uniform_kernel     counter-based synthetic inputs, bit-identical to the CPU oracle.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#ifndef EC_FUSE_ROW_BLOCKS
#define EC_FUSE_ROW_BLOCKS 2   // row blocks (of 8 rows) stored and restarted per step of the fused tile turn-around
#endif

namespace ec {

// ------------------------------------------------------------------------------------------
// Tile configurations of the hot kernel
// ------------------------------------------------------------------------------------------
constexpr int BK = 16;  // k extent per stage: 16 doubles = 128 B = one swizzle span
constexpr int SK_TREE_MIN = 8;     // stream-K: more contributors to one tile than this -> two-level sum
constexpr int SK_TREE_GROUP = 16;  // ... in groups of this many consecutive ranks

// One CTA computes a BM x BN tile of C with (BM/WM) x (BN/WN) consumer warps, each holding a
// WM x WN sub-tile in registers, plus one producer warpgroup (its first warp's elected lane issues
// the TMA loads).  The register file is allocated per 4 warps, so a lone producer warp would cost
// as much as four: the producer group hands its registers to the consumers with setmaxnreg.
template <int BM_, int BN_, int WM_, int WN_, int STAGES_, int MIN_CTAS_, bool REGSPLIT_>
struct TileCfg {
  static constexpr int BM = BM_, BN = BN_, WM = WM_, WN = WN_, STAGES = STAGES_;
  static constexpr int MIN_CTAS = MIN_CTAS_;
  static constexpr bool REGSPLIT = REGSPLIT_;
  // The register-rich one-CTA-per-SM tile overlaps a tile's C stores with the first k-iteration of
  // the CTA's next tile (see "fused tile turn-around" in the kernel).
  static constexpr bool FUSE_EPI = REGSPLIT_ && MIN_CTAS_ == 1;
  static constexpr int WARPS_M = BM / WM, WARPS_N = BN / WN;
  static constexpr int CONSUMER_WARPS = WARPS_M * WARPS_N;
  static constexpr int THREADS = (CONSUMER_WARPS + (REGSPLIT ? 4 : 1)) * 32;
  static constexpr int A_BYTES = BM * BK * 8, B_BYTES = BN * BK * 8;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/;
  // REGSPLIT: setmaxnreg moves registers inside the CTA's own launch allocation (THREADS x
  // LAUNCH_REGS, LAUNCH_REGS = what __launch_bounds__(THREADS, MIN_CTAS) makes ptxas allocate);
  // asking for more than that pool holds blocks forever, so the consumer budget is derived from it:
  //   CONSUMER_WARPS*32*CONSUMER_REGS + 128*PRODUCER_REGS <= THREADS*LAUNCH_REGS
  static constexpr int LAUNCH_REGS_RAW = (65536 / (MIN_CTAS * THREADS)) / 8 * 8;
  static constexpr int LAUNCH_REGS = LAUNCH_REGS_RAW > 248 ? 248 : LAUNCH_REGS_RAW;
  static constexpr int PRODUCER_REGS = MIN_CTAS == 1 ? 40 : 24;
  static constexpr int CONSUMER_REGS_RAW =
      ((THREADS * LAUNCH_REGS - 128 * PRODUCER_REGS) / (CONSUMER_WARPS * 32)) / 8 * 8;
  static constexpr int CONSUMER_REGS = CONSUMER_REGS_RAW > 232 ? 232 : CONSUMER_REGS_RAW;
  static_assert(!REGSPLIT ||
                    CONSUMER_WARPS * 32 * CONSUMER_REGS + 128 * PRODUCER_REGS <= THREADS * LAUNCH_REGS,
                "register split exceeds the CTA's launch allocation (setmaxnreg.inc would hang)");
  static_assert(!REGSPLIT || (CONSUMER_REGS >= LAUNCH_REGS && PRODUCER_REGS <= LAUNCH_REGS), "inc/dec direction");
  static_assert(WM % 16 == 0 && WN % 16 == 0 && BM % WM == 0 && BN % WN == 0, "tile shape");
  static_assert(!REGSPLIT || CONSUMER_WARPS % 4 == 0, "setmaxnreg works on whole warpgroups");
};
//                    BM   BN   WM  WN  ST CTAs regsplit
using CfgL  = TileCfg<128, 128, 64, 32, 5, 1, true>;    // 8 consumer warps, 160 KiB, 232 regs (6 stages measured: no gain)
using CfgS  = TileCfg< 64,  64, 32, 32, 4, 2, false>;   // 4 consumer warps,  64 KiB, 2 CTAs/SM
using CfgXS = TileCfg< 64,  32, 32, 16, 4, 3, false>;   // 4 consumer warps,  48 KiB, 3 CTAs/SM
using CfgT  = TileCfg< 32,  32, 16, 16, 4, 5, false>;   // 4 consumer warps,  33 KiB, 5 CTAs/SM (80 regs; 6 would spill): tensors of matrices of 32 x 32 and less
// (An 8-deep ring for the 32x32 tile was measured in round 2 for the small single-wave products -- n = 128..384
//  unchanged, 7.1 vs 7.2 us back to back, n >= 512 slower for the lost occupancy: they are bound by launch and
//  prologue, not by the TMA round trip -- and is not built.)
using CfgN  = TileCfg<128,  16, 32, 16, 4, 3, false>;   // 4 consumer warps,  73 KiB, 3 CTAs/SM: skinny products with at most 16 columns --
                                                        // a 32-wide tile would spend half its DMMAs on padding, which is what bounds these HBM-bound shapes
// (A 128x64 / 2 CTAs-per-SM variant and register-split variants of the small tiles were measured
//  and lost everywhere -- profiles/r01_tile_sweep*.json -- so they are not built.)

// Division of a 31-bit dividend by an invariant divisor as one widening multiply and a shift
// (Granlund-Montgomery: m = ceil(2^(31+l) / d), l = ceil(log2 d), fits 32 bits; q = n*m >> (31+l)
// is exact for every n < 2^31).  A tile id is turned into (problem, tile row, tile column) on the
// critical path twice per tile -- by the producer before its first load, by the consumers before
// their stores -- and a 64-bit hardware-less division is a ~100-instruction dependent chain.
struct FastDiv {
  uint32_t m, sh, d;
  __host__ __device__ __forceinline__ uint32_t div(uint32_t n) const {
    return (uint32_t)(((uint64_t)n * m) >> sh);
  }
  static __host__ __device__ inline FastDiv make(uint32_t d) {
    FastDiv f;
    uint32_t l = 0;
    while (l < 31 && (1u << l) < d) ++l;
    f.sh = 31 + l;
    f.m = (uint32_t)((((uint64_t)1 << f.sh) + d - 1) / d);
    f.d = d;
    return f;
  }
};

struct GemmParams {
  int64_t M, N, K;
  int64_t batch;
  double alpha, beta;
  double *C;
  int64_t ldc, stride_c;
  int tiles_m, tiles_n;
  int64_t total_tiles;
  int a_batched, b_batched;  // 0: operand shared by every problem of the batch (coordinate 0)
  // Stream-K region (see "work schedule" below): tiles [0, sk_tiles) are cut along k into
  // gridDim.x equal contiguous runs of k-iterations; tiles [sk_tiles, total_tiles) are whole
  // tiles dealt round-robin.  sk_tiles == 0 -> plain data-parallel schedule.
  int64_t sk_tiles;
  double *sk_workspace;    // gridDim.x partial accumulator tiles, thread-linear layout
  unsigned *sk_flags;      // gridDim.x flags; flag[c] == sk_epoch <=> CTA c's partial is published
  unsigned sk_epoch;       // unique per launch, so flags never need resetting
  int c_vec2;              // C base, ldc and batch stride are 16-byte aligned: row pairs may use 16 B stores
  // Data-parallel tiles are handed out dynamically when dp_counter is set: the CTA's first tile is
  // sk_tiles + blockIdx.x, every further one is sk_tiles + gridDim.x + (atomicAdd(dp_counter, 1) -
  // dp_base).  The counter is never reset: the host knows how many claims a launch makes.
  unsigned long long *dp_counter;
  unsigned long long dp_base;
  int dp_all_dynamic;      // 1: the first tile of a CTA is claimed too (the GPU is known to be shared:
                           //    a CTA that has to wait for its SM must not sit on a tile meanwhile)
  // Start-skew check of a persistent grid: stats[launch_parity] = earliest CTA start of this launch
  // (globaltimer ns; the other slot is re-armed for the next launch).  A CTA that starts more than
  // skew_limit_ns after the earliest one reports its lateness in *skew_out (host-mapped): it had
  // to wait for an SM that some other kernel held.  Null = off.
  unsigned long long *stats;
  unsigned *skew_out;
  unsigned skew_limit_ns;
  int launch_parity;
  int pdl_trigger;         // where griddepcontrol.launch_dependents is issued: 0 never, 1 after the wait, 2 at kernel start, 3 before the first epilogue
  FastDiv fd_tiles_per_problem, fd_tiles_m;   // valid when total_tiles < 2^31 (fast_locate)
  int fast_locate;
  int fuse_epilogue_wanted; // host side only: the pipeline's setting, read before the launch fills fuse_epilogue
  int fuse_epilogue;       // 1: tiles with FUSE_EPI defer their C stores into the next tile's first k-iteration
  int sk_sms;              // SM count: CTAs b and b + sk_sms share an SM and get ADJACENT stream-K
                           // ranges, so their tile boundaries (epilogues) fall at different times
};

// ------------------------------------------------------------------------------------------
// PTX helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(m) : "memory");
}
// 3D tiled TMA load, completion reported as transaction bytes on an mbarrier.
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *map, uint32_t bar,
                                            int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// FP64 tensor-core MMA: D(8x8) += A(8x4, row) * B(4x8, col).  lane = 4*g + c:
// A fragment = A[g][c], B fragment = B[c][g], D fragment = D[g][2c], D[g][2c+1].
__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
      : "+d"(d0), "+d"(d1)
      : "d"(a), "d"(b));
}
template <int REGS>
__device__ __forceinline__ void setmaxnreg_inc() {
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS));
}
template <int REGS>
__device__ __forceinline__ void setmaxnreg_dec() {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS));
}
__device__ __forceinline__ void st_release_gpu(unsigned *p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned *p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// Programmatic dependent launch (PDL).  launch_dependents: the next kernel of the stream may be
// scheduled from now on (it becomes resident where SM resources allow and runs its prologue
// underneath this grid); grid_dependency_wait: blocks until every prerequisite grid has completed
// and its memory is visible.  Both are no-ops for a launch without the PDL attribute.
__device__ __forceinline__ void pdl_launch_dependents() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
__device__ __forceinline__ void pdl_grid_dependency_wait() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void st_shared_s64(uint32_t addr, long long v) {
  asm volatile("st.shared.s64 [%0], %1;" ::"r"(addr), "l"(v) : "memory");
}
__device__ __forceinline__ long long ld_shared_s64(uint32_t addr) {
  long long v;
  asm volatile("ld.shared.s64 %0, [%1];" : "=l"(v) : "r"(addr) : "memory");
  return v;
}
// named barrier over the consumer warps only (barrier 0 is __syncthreads)
__device__ __forceinline__ void consumer_bar_sync(int nthreads) {
  asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory");
}
__device__ __forceinline__ double2 lds128(uint32_t addr) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
  return v;
}

// ------------------------------------------------------------------------------------------
// Shared-memory operand layouts (per stage, per operand: 128 "mn" x 16 "k" doubles = 16 KiB).
//
// K-major (memory-contiguous along k: op(A)=A^T, op(B)=B): ONE TMA box {16 k, 128 mn}; row mn is
//   128 B, the 16-byte chunk index is XORed with (mn & 7) by the 128B swizzle:
//       off(mn,k) = mn*128 + (((k>>1) ^ (mn&7)) << 4) + (k&1)*8
// MN-major (memory-contiguous along m or n: op(A)=A, op(B)=B^T): EIGHT TMA boxes {16 mn, 16 k},
//   sub-tile j = mn>>4 at j*2048, row k is 128 B, chunk XORed with (k & 7):
//       off(mn,k) = (mn>>4)*2048 + k*128 + ((((mn&15)>>1) ^ (k&7)) << 4) + (mn&1)*8
//
// One k8 step issues, per 8x8 output block, two DMMAs: "even" takes k = {0,2,4,6} on fragment
// lanes c = 0..3 and "odd" takes k = {1,3,5,7}.  Any permutation of k inside a DMMA is legal as
// long as both operands agree, and any permutation of the 8 rows (columns) a fragment lane g
// stands for is legal as long as the epilogue writes C accordingly.  The permutations below are
// what makes every LDS.128 bank-conflict-free (each quarter-warp covers 8 distinct 16 B chunks):
//   K-major : lane (g,c) loads 16 B = k {2c, 2c+1} of row  8*blk + pi(g),  pi(g) = (g>>1) + 4*(g&1)
//             -> .x feeds the even DMMA, .y the odd DMMA of block blk.
//   MN-major: lane (g,c) loads 16 B = mn {2g, 2g+1} of 16-block jb at k = 2c (then at k = 2c+1)
//             -> .x is a fragment of block 2*jb (rows 16*jb + 2g), .y of block 2*jb+1 (rows +1).
// ------------------------------------------------------------------------------------------
template <bool KMAJOR>
__device__ __forceinline__ int frag_mn(int blk, int g) {
  if (KMAJOR) return 8 * blk + (g >> 1) + 4 * (g & 1);
  return 16 * (blk >> 1) + 2 * g + (blk & 1);
}

// Load the fragments of NBLK 8-wide blocks starting at warp offset `w0` (multiple of 16) for the
// k8 step at kk (0 or 8).  fe[b] / fo[b]: even-k / odd-k fragment of block b.
template <bool KMAJOR, int NBLK>
__device__ __forceinline__ void load_frags(uint32_t base, int w0, int kk, int g, int c,
                                           double (&fe)[NBLK], double (&fo)[NBLK]) {
  if (KMAJOR) {
    const int pg = (g >> 1) + 4 * (g & 1);
    const uint32_t lane_off = (uint32_t)((w0 + pg) * 128 + ((((kk >> 1) + c) ^ pg) << 4));
#pragma unroll
    for (int b = 0; b < NBLK; ++b) {
      double2 v = lds128(base + lane_off + b * 1024);
      fe[b] = v.x;
      fo[b] = v.y;
    }
  } else {
    const int ke = kk + 2 * c, ko = ke + 1;
    const uint32_t off_e = (uint32_t)((w0 >> 4) * 2048 + ke * 128 + ((g ^ (ke & 7)) << 4));
    const uint32_t off_o = (uint32_t)((w0 >> 4) * 2048 + ko * 128 + ((g ^ (ko & 7)) << 4));
#pragma unroll
    for (int jb = 0; jb < NBLK / 2; ++jb) {
      double2 ve = lds128(base + off_e + jb * 2048);
      double2 vo = lds128(base + off_o + jb * 2048);
      fe[2 * jb] = ve.x;
      fe[2 * jb + 1] = ve.y;
      fo[2 * jb] = vo.x;
      fo[2 * jb + 1] = vo.y;
    }
  }
}

// Producer side: fill one operand's stage (EXTENT = BM or BN rows of the operand tile).
// K-major: one box {16 k, EXTENT mn}; MN-major: EXTENT/16 boxes {16 mn, 16 k}.
template <bool KMAJOR, int EXTENT>
__device__ __forceinline__ void tma_fill_operand(uint32_t dst, const CUtensorMap *map, uint32_t bar,
                                                 int mn0, int k0, int b) {
  if (KMAJOR) {
    tma_load_3d(dst, map, bar, k0, mn0, b);
  } else {
#pragma unroll
    for (int j = 0; j < EXTENT / 16; ++j) tma_load_3d(dst + j * 2048, map, bar, mn0 + 16 * j, k0, b);
  }
}

// ------------------------------------------------------------------------------------------
// Work schedule, shared by the producer and the consumers of a CTA
// ------------------------------------------------------------------------------------------
// A "segment" is a run of k-iterations [kb0, kb1) of one output tile.  CTA c of G first walks
// its share of the stream-K region -- the contiguous run [S*c/G, S*(c+1)/G) of the region's
// S = sk_tiles*kblocks iterations, which cuts into at most: the tail of one tile, some whole
// tiles, the head of one tile -- and then whole tiles sk_tiles + c, sk_tiles + c + G, ...
// With sk_tiles == 0 this is the plain persistent data-parallel schedule.
struct Segment {
  int64_t tile;
  int kb0, kb1;
};
constexpr int TQ = 4;   // (power of two) depth of the tile queue between the producer and the consumers of a CTA
struct Schedule {
  int64_t sk_it, sk_end;   // position / end inside the stream-K iteration space
  int kblocks;
  // Stream-K rank of this CTA.  The hardware places CTA b and b + #SMs on the same SM; giving
  // them consecutive ranks staggers their phases inside a tile by one CTA share.
  static __device__ __forceinline__ unsigned sk_rank(const GemmParams &p) {
    const unsigned per_sm = gridDim.x / p.sk_sms;
    if (per_sm * p.sk_sms != gridDim.x) return blockIdx.x;   // grid not a multiple of the SM count
    return (blockIdx.x % p.sk_sms) * per_sm + blockIdx.x / p.sk_sms;
  }
  __device__ __forceinline__ Schedule(const GemmParams &p, int kblocks_) {
    kblocks = kblocks_;
    sk_it = sk_end = 0;
    if (p.sk_tiles == 0) return;   // no stream-K region: keep three integer divisions off the start of every kernel
    const int64_t S = p.sk_tiles * kblocks;
    const unsigned rank = sk_rank(p);
    sk_it = S * rank / gridDim.x;
    sk_end = S * (rank + 1) / gridDim.x;
  }
  // next segment of this CTA's share of the stream-K region (static: both sides compute it)
  __device__ __forceinline__ bool next_sk(Segment &sg) {
    if (sk_it >= sk_end) return false;
    sg.tile = sk_it / kblocks;
    sg.kb0 = (int)(sk_it - sg.tile * kblocks);
    const int64_t left = sk_end - sk_it;
    sg.kb1 = (int)(left < (int64_t)(kblocks - sg.kb0) ? sg.kb0 + left : kblocks);
    sk_it += sg.kb1 - sg.kb0;
    return true;
  }
};
// After the stream-K region come whole ("data-parallel") tiles.  Which ones is decided by the
// producer thread alone -- statically (tile += gridDim.x) or by claiming them from a global counter
// -- and passed to the consumer warps through a TQ-deep queue in shared memory (tile id + a
// full/empty mbarrier pair per slot; -1 ends the stream).  With dynamic claims a CTA that starts
// late or shares its SM with somebody else's kernel simply takes fewer tiles: the grid no longer
// runs at the pace of its most delayed CTA.

// tile id -> problem of the batch, tile row, tile column (m fastest)
__device__ __forceinline__ void locate_tile(const GemmParams &p, int64_t tile, int tiles_per_problem,
                                            int &b, int &tm, int &tn) {
  if (p.fast_locate) {
    const uint32_t t = (uint32_t)tile;
    const uint32_t ub = p.fd_tiles_per_problem.div(t);
    const uint32_t r = t - ub * (uint32_t)tiles_per_problem;
    const uint32_t un = p.fd_tiles_m.div(r);
    b = (int)ub;
    tn = (int)un;
    tm = (int)(r - un * (uint32_t)p.tiles_m);
  } else {
    b = (int)(tile / tiles_per_problem);
    const int r = (int)(tile - (int64_t)b * tiles_per_problem);
    tn = r / p.tiles_m;
    tm = r - tn * p.tiles_m;
  }
}

// ------------------------------------------------------------------------------------------
// The hot kernel
// ------------------------------------------------------------------------------------------
// (FUSE_EDGE: the fused tile turn-around also takes tiles that hang over the edge of C, with predicated
//  element stores.  A second instantiation rather than a run-time branch: the branch alone costs the
//  all-interior headline launch 0.2 % -- one more place where ptxas' schedule of the fused iteration moves.)
template <class Cfg, bool A_KMAJOR, bool B_KMAJOR, bool FUSE_EDGE>
__device__ __forceinline__ void dgemm_tma_dmma_body(const CUtensorMap &tmA, const CUtensorMap &tmB,
                                                    const GemmParams &p) {
  constexpr int BM = Cfg::BM, BN = Cfg::BN, WM = Cfg::WM, WN = Cfg::WN, STAGES = Cfg::STAGES;
  constexpr int CONSUMER_WARPS = Cfg::CONSUMER_WARPS, STAGE_BYTES = Cfg::STAGE_BYTES;
  constexpr int CONSUMER_THREADS = CONSUMER_WARPS * 32;
  extern __shared__ uint8_t smem_raw[];
  // 128B-swizzled TMA destinations need 1024-byte alignment.
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + STAGES * STAGE_BYTES;  // full[STAGES], empty[STAGES], tile queue
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  static_assert(2 * STAGES * 8 <= 128, "barrier area layout");
  auto tq_full = [&](int q) { return bar_base + 128u + 8u * q; };
  auto tq_empty = [&](int q) { return bar_base + 128u + 8u * (TQ + q); };
  auto tq_tile = [&](int q) { return bar_base + 128u + 8u * (2 * TQ + q); };

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // PDL: the next product of the stream (the compat loop issues them back to back) may set itself
  // up while this one runs -- its launch latency, barrier init and descriptor fetch disappear
  // behind this grid.  Everything up to the wait below touches no global memory.
  if (p.pdl_trigger == 2) pdl_launch_dependents();

  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);                // one arrive.expect_tx by the producer
      mbar_init(empty_bar(s), CONSUMER_WARPS);  // one arrive per consumer warp
    }
#pragma unroll
    for (int q = 0; q < TQ; ++q) {
      mbar_init(tq_full(q), 1);                 // the producer, after writing the tile id
      mbar_init(tq_empty(q), CONSUMER_WARPS);   // one arrive per consumer warp, after reading it
    }
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == CONSUMER_WARPS && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  __syncthreads();
  // from here on global memory is read (operands, beta * C, stream-K flags) and written: the
  // kernels queued before this one must be complete
  pdl_grid_dependency_wait();
  // Trigger only now, with the predecessor gone: at most ONE dependent grid is ever resident
  // underneath a running one (triggering before the wait lets a whole queue of launches pile
  // into the SMs' free slots).
  if (p.pdl_trigger == 1) pdl_launch_dependents();
  const int kblocks = (int)((p.K + BK - 1) / BK);
  const int tiles_per_problem = p.tiles_m * p.tiles_n;

  if (warp >= CONSUMER_WARPS) {
    // ===================== TMA producer (one elected lane) =====================
    if (Cfg::REGSPLIT) setmaxnreg_dec<Cfg::PRODUCER_REGS>();
    if (warp == CONSUMER_WARPS && lane == 0) {
      // when did this CTA get its SM?  (compared with the rest of the grid once the loads are out)
      const unsigned long long t_start = p.stats ? globaltimer_ns() : 0ull;
      if (p.stats && blockIdx.x == 0) p.stats[p.launch_parity ^ 1] = ~0ull;   // re-arm the next launch's slot
      int stage = 0;
      uint32_t phase = 0;
      // (claim_kb, nxt: whole tiles only -- the CTA's next tile is claimed when the loads of this one
      //  reach k-iteration claim_kb)
      int64_t nxt = p.total_tiles;
      auto load_segment = [&](const Segment &sg, int claim_kb, auto &&claim_next) {
        int b, tm, tn;
        locate_tile(p, sg.tile, tiles_per_problem, b, tm, tn);
        const int m0 = tm * BM;
        const int n0 = tn * BN;
        const int ba = p.a_batched ? b : 0, bb = p.b_batched ? b : 0;
        for (int kb = sg.kb0; kb < sg.kb1; ++kb) {
          if (kb == claim_kb) nxt = claim_next();
          mbar_wait(empty_bar(stage), phase ^ 1u);
          mbar_expect_tx(full_bar(stage), STAGE_BYTES);
          const uint32_t sA = smem_base + stage * STAGE_BYTES;
          const uint32_t sB = sA + Cfg::A_BYTES;
          tma_fill_operand<A_KMAJOR, BM>(sA, &tmA, full_bar(stage), m0, kb * BK, ba);
          tma_fill_operand<B_KMAJOR, BN>(sB, &tmB, full_bar(stage), n0, kb * BK, bb);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
      };
      Schedule sched(p, kblocks);
      Segment sg;
      while (sched.next_sk(sg)) load_segment(sg, -1, [&]() -> int64_t { return 0; });
      // Whole tiles: the first one is fixed, further ones are claimed while the last (up to) 8
      // k-iterations of the tile before are being loaded: the round trip of the atomic hides behind
      // those loads, and a CTA never sits on more than the tile it is loading.  (Round 2 first
      // claimed a whole tile earlier -- two tiles at kernel start -- which let half the CTAs take
      // every dynamic tile when there were exactly two whole tiles per CTA: n = 2816 and 3072 ran
      // at 28 instead of 36 TFLOP/s, profiles/r02_sk_probe.txt.)
      const int64_t dyn_first = p.sk_tiles + (p.dp_all_dynamic ? 0 : (int64_t)gridDim.x);
      auto claim = [&]() -> int64_t {
        return dyn_first + (int64_t)(atomicAdd(p.dp_counter, 1ull) - p.dp_base);
      };
      // A CTA's first whole tile is fixed unless every tile is claimed (dp_all_dynamic): both sides
      // know it without talking.  The queue carries what comes after it -- and nothing at all when
      // no CTA can have a second tile (single-wave grids: the latency-bound small products).
      const bool first_static = !p.dp_all_dynamic;
      const bool queued = p.dp_all_dynamic || p.total_tiles - p.sk_tiles > (int64_t)gridDim.x;
      int64_t cur = first_static ? p.sk_tiles + blockIdx.x : claim();
      int q = 0;
      uint32_t qphase = 0;
      for (bool first = first_static;; first = false) {
        const bool valid = cur < p.total_tiles;
        if (queued && !(first && valid)) {   // the fixed first tile is not queued
          mbar_wait(tq_empty(q), qphase ^ 1u);
          st_shared_s64(tq_tile(q), valid ? (long long)cur : -1ll);
          mbar_arrive(tq_full(q));
          if (++q == TQ) {
            q = 0;
            qphase ^= 1u;
          }
        }
        if (!valid) break;
        sg.tile = cur;
        sg.kb0 = 0;
        sg.kb1 = kblocks;
        nxt = p.total_tiles;   // invalid
        load_segment(sg, queued ? (kblocks > 8 ? kblocks - 8 : 0) : -1,
                     [&]() -> int64_t { return p.dp_counter ? claim() : cur + (int64_t)gridDim.x; });
        if (!queued) break;
        cur = nxt;
      }
      // Start-skew check, here where it delays nobody: this thread has nothing left to do.  A CTA
      // that started much later than the first one of the grid had to wait for an SM that some
      // other kernel held; it says so in host-mapped memory, where the dispatcher finds it.
      if (p.stats) {
        const unsigned long long first = atomicMin(p.stats + p.launch_parity, t_start);
        if (first < t_start && t_start - first > p.skew_limit_ns) {
          const unsigned long long late = t_start - first;
          atomicMax(p.skew_out, late > 0xffffffffull ? 0xffffffffu : (unsigned)late);
        }
      }
    }
    return;
  }

  // ===================== DMMA consumers =====================
  if (Cfg::REGSPLIT) setmaxnreg_inc<Cfg::CONSUMER_REGS>();
  const int g = lane >> 2, c = lane & 3;
  const int wm0 = (warp % Cfg::WARPS_M) * WM;  // warp's row offset inside the CTA tile
  const int wn0 = (warp / Cfg::WARPS_M) * WN;  // warp's column offset
  constexpr int MB = WM / 8;                   // row blocks per warp
  constexpr int NB = WN / 8;                   // column blocks per warp

  int stage = 0;
  uint32_t phase = 0;
  uint32_t pending_bar = 0;  // empty barrier of the slot consumed last, not yet released
  Schedule sched(p, kblocks);
  Segment sg;
  // whole tiles: state of the consumer side of the tile queue.  qn >= 0: pops so far modulo 2*TQ
  // (slot = qn % TQ, parity = (qn / TQ) & 1); -1: the stream has ended; -2: the CTA's fixed first
  // tile is still to be done (it is not queued, see the producer)
  const bool queued = p.dp_all_dynamic || p.total_tiles - p.sk_tiles > (int64_t)gridDim.x;
  int qn = p.dp_all_dynamic ? 0 : -2;
  // A stream-K tile this CTA owns (it computed the tile's first k-iterations) is finished LAST,
  // after the CTA's whole data-parallel share: its accumulators wait in the workspace meanwhile.
  // By then the other contributors' partials have long been published, so the owner hardly ever
  // spins -- and no CTA spins in the middle of its work, which is what lets the grid be launched
  // without a co-residency guarantee (a CTA that cannot become resident yet only delays the
  // finishing touches of the tiles it contributes to, not the data-parallel work of the others).
  int own_tile = -1, own_kb1 = 0;   // (stream-K tiles come first in the numbering: they fit an int)
  // spill slot of the owner: behind the gridDim.x contributor slots
  auto own_ws = [&]() -> double * {
    return p.sk_workspace + ((size_t)gridDim.x + Schedule::sk_rank(p)) * (BM * BN) + threadIdx.x;
  };
  double acc[MB][NB][2];
  // Fused tile turn-around (FUSE_EPI): a finished interior tile with beta == 0 is not stored at
  // once.  Its 128 KB would drain through the SM's store path at 64 B/clk -- 2 k clocks during
  // which all eight consumer warps sit in the epilogue together and the DMMA pipe idles.  Instead
  // the tile stays in the accumulators ("pending") and the FIRST k-iteration of the CTA's next
  // segment is issued row-block pair by row-block pair: store the pair's 16 x 32 results, clear
  // those accumulators, issue the pair's 32 DMMAs of the new tile.  The stores of a tile thus
  // spread over one k-iteration's worth of DMMA time (4 k clocks per SM), every accumulator still
  // sees its k-steps in the same order (results are bit-identical), and no shared memory is spent.
  bool pend = false;
  double *pend_base = nullptr;
  // stores of row blocks 2*ip, 2*ip+1 of the pending tile (beta == 0, interior: no predicates)
  bool pend_edge = false;             // the pending tile hangs over the edge of C (FUSE_EDGE only)
  int pend_rows = 0, pend_cols = 0;   // rows / columns of C left from this warp's corner of the pending tile
  auto store_pair = [&](double *base, int ip) {
    if constexpr (FUSE_EDGE) {
      if (pend_edge) {
#pragma unroll
        for (int j = 0; j < NB; ++j) {
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int col = frag_mn<B_KMAJOR>(j, 2 * c + e);
            if (col < pend_cols) {
              double *pc = base + (int64_t)col * p.ldc;
#pragma unroll
              for (int ii = 0; ii < 2; ++ii) {
                const int row = frag_mn<A_KMAJOR>(2 * ip + ii, g);
                if (row < pend_rows) pc[row] = p.alpha * acc[2 * ip + ii][j][e];
              }
            }
          }
        }
        return;
      }
    }
#pragma unroll
    for (int j = 0; j < NB; ++j) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        double *pc = base + (int64_t)frag_mn<B_KMAJOR>(j, 2 * c + e) * p.ldc;
        if (A_KMAJOR) {
          pc += (g >> 1) + 4 * (g & 1);
          pc[16 * ip] = p.alpha * acc[2 * ip][j][e];
          pc[16 * ip + 8] = p.alpha * acc[2 * ip + 1][j][e];
        } else {
          pc += 2 * g;
          double2 v;
          v.x = p.alpha * acc[2 * ip][j][e];
          v.y = p.alpha * acc[2 * ip + 1][j][e];
          if (p.c_vec2) {
            *reinterpret_cast<double2 *>(pc + 16 * ip) = v;
          } else {
            pc[16 * ip] = v.x;
            pc[16 * ip + 1] = v.y;
          }
        }
      }
    }
  };
  for (;;) {
    bool resume = false;
    if (sched.next_sk(sg)) {
    } else if (qn == -2) {
      qn = queued ? 0 : -1;
      if (p.sk_tiles + blockIdx.x >= p.total_tiles) continue;   // no first tile: nothing follows either
      sg.tile = p.sk_tiles + blockIdx.x;
      sg.kb0 = 0;
      sg.kb1 = kblocks;
    } else if (qn >= 0) {
      const int q = qn & (TQ - 1);
      mbar_wait(tq_full(q), (uint32_t)(qn / TQ) & 1u);
      const long long t = ld_shared_s64(tq_tile(q));
      __syncwarp();
      if (lane == 0) mbar_arrive(tq_empty(q));
      if (t < 0) {
        qn = -1;
        continue;
      }
      qn = (qn + 1) & (2 * TQ - 1);
      sg.tile = t;
      sg.kb0 = 0;
      sg.kb1 = kblocks;
    } else if (own_tile >= 0) {
      sg.tile = own_tile;
      sg.kb0 = 0;
      sg.kb1 = own_kb1;
      own_tile = -1;
      resume = true;
    } else {
      break;
    }
    bool fused = false;
    if (Cfg::FUSE_EPI && pend) {
      pend = false;
      if (!resume) {
        fused = true;
      } else {
#pragma unroll
        for (int ip = 0; ip < MB / 2; ++ip) store_pair(pend_base, ip);
      }
    }
    if (Cfg::FUSE_EPI && fused) {
      mbar_wait(full_bar(stage), phase);
      if (pending_bar != 0) {   // deferred slot release, as in the main loop below
        __syncwarp();
        if (lane == 0) mbar_arrive(pending_bar);
      }
      const uint32_t sA = smem_base + stage * STAGE_BYTES;
      const uint32_t sB = sA + Cfg::A_BYTES;
      double be[2][NB], bo[2][NB];
      load_frags<B_KMAJOR, NB>(sB, wn0, 0, g, c, be[0], bo[0]);
      load_frags<B_KMAJOR, NB>(sB, wn0, 8, g, c, be[1], bo[1]);
      constexpr int R = (EC_FUSE_ROW_BLOCKS <= MB ? EC_FUSE_ROW_BLOCKS : MB);   // row blocks per step (even)
#pragma unroll
      for (int ig = 0; ig < MB / R; ++ig) {
        double ae[2][R], ao[2][R];
        load_frags<A_KMAJOR, R>(sA, wm0 + 8 * R * ig, 0, g, c, ae[0], ao[0]);
        load_frags<A_KMAJOR, R>(sA, wm0 + 8 * R * ig, 8, g, c, ae[1], ao[1]);
#pragma unroll
        for (int q = 0; q < R / 2; ++q) store_pair(pend_base, (R / 2) * ig + q);
#pragma unroll
        for (int ii = 0; ii < R; ++ii)
#pragma unroll
          for (int j = 0; j < NB; ++j) acc[R * ig + ii][j][0] = acc[R * ig + ii][j][1] = 0.0;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
#pragma unroll
          for (int ii = 0; ii < R; ++ii)
#pragma unroll
            for (int j = 0; j < NB; ++j)
              dmma884(acc[R * ig + ii][j][0], acc[R * ig + ii][j][1], ae[h][ii], be[h][j]);
#pragma unroll
          for (int ii = 0; ii < R; ++ii)
#pragma unroll
            for (int j = 0; j < NB; ++j)
              dmma884(acc[R * ig + ii][j][0], acc[R * ig + ii][j][1], ao[h][ii], bo[h][j]);
        }
      }
      pending_bar = empty_bar(stage);
      if (++stage == STAGES) {
        stage = 0;
        phase ^= 1u;
      }
    } else if (resume) {
      const double *ws = own_ws();
#pragma unroll
      for (int i = 0; i < MB; ++i)
#pragma unroll
        for (int j = 0; j < NB; ++j)
#pragma unroll
          for (int e = 0; e < 2; ++e) acc[i][j][e] = __ldcg(ws + ((i * NB + j) * 2 + e) * CONSUMER_THREADS);
    } else {
#pragma unroll
      for (int i = 0; i < MB; ++i)
#pragma unroll
        for (int j = 0; j < NB; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    }

    for (int kb = resume ? sg.kb1 : sg.kb0 + (fused ? 1 : 0); kb < sg.kb1; ++kb) {
      mbar_wait(full_bar(stage), phase);
      // Deferred slot release: the slot of the PREVIOUS k-iteration is handed back only now.
      // An arrive issued right behind the loads of its own stage carries no scoreboard wait,
      // and loads can sit in the memory pipeline for a long time behind another resident CTA's
      // epilogue stores -- long enough for the producer to see the slot free and for the refill
      // to land underneath them (observed on B200 as rare corrupted 16-row fragments whenever
      // more than one CTA shared an SM).  At this point every DMMA of the previous iteration has
      // been issued (in-order issue), so every LDS feeding them has returned its data.  Costs one
      // slot of ring depth, no instructions.
      if (pending_bar != 0) {
        __syncwarp();
        if (lane == 0) mbar_arrive(pending_bar);
      }
      const uint32_t sA = smem_base + stage * STAGE_BYTES;
      const uint32_t sB = sA + Cfg::A_BYTES;
#pragma unroll
      for (int kk = 0; kk < BK; kk += 8) {
        double ae[MB], ao[MB], be[NB], bo[NB];
        load_frags<A_KMAJOR, MB>(sA, wm0, kk, g, c, ae, ao);
        load_frags<B_KMAJOR, NB>(sB, wn0, kk, g, c, be, bo);
#pragma unroll
        for (int i = 0; i < MB; ++i)
#pragma unroll
          for (int j = 0; j < NB; ++j) dmma884(acc[i][j][0], acc[i][j][1], ae[i], be[j]);
#pragma unroll
        for (int i = 0; i < MB; ++i)
#pragma unroll
          for (int j = 0; j < NB; ++j) dmma884(acc[i][j][0], acc[i][j][1], ao[i], bo[j]);
      }
      pending_bar = empty_bar(stage);
      if (++stage == STAGES) {
        stage = 0;
        phase ^= 1u;
      }
    }

    // ---- stream-K fix-up -------------------------------------------------------------------
    // A segment that does not start the tile publishes its partial sums (it is always the first
    // thing this CTA computes).  The segment that starts a tile but does not finish it owns the
    // tile: it parks its accumulators, and after the CTA's data-parallel tiles it adds, in CTA
    // order, the partials of the CTAs that follow it -- published long before -- and writes C.
    //
    // Deep splits (a 64 x 64 product over K = 10^6 is cut over every CTA of the grid) would make
    // that a serial chain of hundreds of dependent workspace reads in ONE CTA.  Beyond
    // SK_TREE_MIN contributors the sum is therefore taken in two levels: the contributors form
    // groups of SK_TREE_GROUP consecutive ranks, the first of each group adds its followers'
    // partials to its own before publishing, and the owner adds the group sums.  Followers never
    // wait, leaders wait only for followers, the owner only for leaders: no cycle; the order of
    // additions is fixed by the ranks, so results stay bit-reproducible.
    if (sg.kb0 != 0 || sg.kb1 != kblocks) {
      const int64_t S = p.sk_tiles * (int64_t)kblocks;
      const int64_t G = gridDim.x;
      const unsigned rank = Schedule::sk_rank(p);
      const int64_t t0 = sg.tile * (int64_t)kblocks, t1 = t0 + kblocks;
      // first and last rank whose run [S*d/G, S*(d+1)/G) meets this tile (runs are non-empty when S >= G)
      const unsigned owner = (unsigned)(((t0 + 1) * G + S - 1) / S - 1);
      const unsigned last = (unsigned)((t1 * G + S - 1) / S - 1);
      const bool tree = S >= G && last - owner > (unsigned)SK_TREE_MIN;
      auto add_partial = [&](unsigned d) {
        if (threadIdx.x == 0) {
          while (ld_acquire_gpu(p.sk_flags + d) != p.sk_epoch) {
          }
        }
        consumer_bar_sync(CONSUMER_THREADS);
        const double *ws = p.sk_workspace + (size_t)d * (BM * BN) + threadIdx.x;
#pragma unroll
        for (int i = 0; i < MB; ++i)
#pragma unroll
          for (int j = 0; j < NB; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) acc[i][j][e] += __ldcg(ws + ((i * NB + j) * 2 + e) * CONSUMER_THREADS);
      };
      if (sg.kb0 != 0) {
        if (tree && (rank - owner - 1) % SK_TREE_GROUP == 0) {   // group leader
          const unsigned group_last = min(rank + SK_TREE_GROUP - 1, last);
          for (unsigned d = rank + 1; d <= group_last; ++d) add_partial(d);
        }
        double *ws = p.sk_workspace + (size_t)rank * (BM * BN) + threadIdx.x;
#pragma unroll
        for (int i = 0; i < MB; ++i)
#pragma unroll
          for (int j = 0; j < NB; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) __stcg(ws + ((i * NB + j) * 2 + e) * CONSUMER_THREADS, acc[i][j][e]);
        __threadfence();
        consumer_bar_sync(CONSUMER_THREADS);
        if (threadIdx.x == 0) st_release_gpu(p.sk_flags + rank, p.sk_epoch);
        continue;
      }
      if (!resume && p.total_tiles > p.sk_tiles) {
        // owner, first visit, and there are data-parallel tiles to do: park the accumulators
        // and come back after them
        double *ws = own_ws();
#pragma unroll
        for (int i = 0; i < MB; ++i)
#pragma unroll
          for (int j = 0; j < NB; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) __stcg(ws + ((i * NB + j) * 2 + e) * CONSUMER_THREADS, acc[i][j][e]);
        own_tile = (int)sg.tile;
        own_kb1 = sg.kb1;
        continue;
      }
      if (tree) {
        for (unsigned d = rank + 1; d <= last; d += SK_TREE_GROUP) add_partial(d);
      } else {
        int covered = sg.kb1;
        for (unsigned d = rank + 1; covered < kblocks; ++d) {
          const int64_t d0 = S * d / G, d1 = S * (d + 1) / G;
          if (d1 == d0) continue;  // CTA d has no stream-K work (S < gridDim.x): nothing to wait for
          add_partial(d);
          covered += (int)((d1 < t1 ? d1 : t1) - d0);
        }
      }
    }

    // ---- epilogue: registers -> global, C = alpha*acc + beta*C ----
    if (p.pdl_trigger == 3) pdl_launch_dependents();
    int b, tm, tn;
    locate_tile(p, sg.tile, tiles_per_problem, b, tm, tn);
    const int m0 = tm * BM;
    const int n0 = tn * BN;
    double *Cb = p.C + (int64_t)b * p.stride_c;
    const bool use_beta = (p.beta != 0.0);
    const bool interior = m0 + BM <= p.M && n0 + BN <= p.N;
    if (Cfg::FUSE_EPI && FUSE_EDGE && !interior && p.fuse_epilogue && !use_beta) {
      // edge tile, deferred like an interior one (e.g. 15 of the 64 tiles of a 1000 x 1000 matrix)
      pend = true;
      pend_base = Cb + (int64_t)(n0 + wn0) * p.ldc + (m0 + wm0);
      pend_edge = true;
      pend_rows = (int)(p.M - (m0 + wm0));
      pend_cols = (int)(p.N - (n0 + wn0));
      continue;
    }
    if (interior) {
      // Interior tile (CTA-uniform branch): no predicates, one base pointer per fragment column,
      // rows at compile-time offsets.  With an MN-major A the row blocks 2jb / 2jb+1 hold
      // adjacent rows 2g / 2g+1, so each pair goes out as one 16-byte store.
      double *base = Cb + (int64_t)(n0 + wn0) * p.ldc + (m0 + wm0);
      if (Cfg::FUSE_EPI && p.fuse_epilogue && !use_beta) {
        pend = true;   // stored underneath the first k-iteration of the next segment, or after the loop
        pend_base = base;
        if (FUSE_EDGE) pend_edge = false;
        continue;
      }
#pragma unroll
      for (int j = 0; j < NB; ++j) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          double *pc = base + (int64_t)frag_mn<B_KMAJOR>(j, 2 * c + e) * p.ldc;
          if (A_KMAJOR) {
            pc += (g >> 1) + 4 * (g & 1);
#pragma unroll
            for (int i = 0; i < MB; ++i) {
              double v = p.alpha * acc[i][j][e];
              if (use_beta) v += p.beta * pc[8 * i];
              pc[8 * i] = v;
            }
          } else {
            pc += 2 * g;
#pragma unroll
            for (int ib = 0; ib < MB / 2; ++ib) {
              double2 v;
              v.x = p.alpha * acc[2 * ib][j][e];
              v.y = p.alpha * acc[2 * ib + 1][j][e];
              double2 *dst = reinterpret_cast<double2 *>(pc + 16 * ib);
              if (use_beta) {
                if (p.c_vec2) {
                  const double2 o = *dst;
                  v.x += p.beta * o.x;
                  v.y += p.beta * o.y;
                } else {
                  v.x += p.beta * pc[16 * ib];
                  v.y += p.beta * pc[16 * ib + 1];
                }
              }
              if (p.c_vec2) {
                *dst = v;
              } else {
                pc[16 * ib] = v.x;
                pc[16 * ib + 1] = v.y;
              }
            }
          }
        }
      }
    } else {
      // Edge tile: bounds-predicated element stores.
#pragma unroll
      for (int j = 0; j < NB; ++j) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int64_t col = n0 + wn0 + frag_mn<B_KMAJOR>(j, 2 * c + e);
          if (col < p.N) {
            double *Ccol = Cb + col * p.ldc;
#pragma unroll
            for (int i = 0; i < MB; ++i) {
              const int64_t row = m0 + wm0 + frag_mn<A_KMAJOR>(i, g);
              if (row < p.M) {
                double v = p.alpha * acc[i][j][e];
                if (use_beta) v += p.beta * Ccol[row];
                Ccol[row] = v;
              }
            }
          }
        }
      }
    }
  }
  if (Cfg::FUSE_EPI && pend) {   // the CTA's last tile
#pragma unroll
    for (int ip = 0; ip < MB / 2; ++ip) store_pair(pend_base, ip);
  }
}

template <class Cfg, bool A_KMAJOR, bool B_KMAJOR>
__global__ void __launch_bounds__(Cfg::THREADS, Cfg::MIN_CTAS)
    dgemm_tma_dmma_kernel(const __grid_constant__ CUtensorMap tmA,
                          const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
  dgemm_tma_dmma_body<Cfg, A_KMAJOR, B_KMAJOR, false>(tmA, tmB, p);
}
// the same for products whose C has partial tiles (128x128 tile only)
template <class Cfg, bool A_KMAJOR, bool B_KMAJOR>
__global__ void __launch_bounds__(Cfg::THREADS, Cfg::MIN_CTAS)
    dgemm_tma_dmma_edge_kernel(const __grid_constant__ CUtensorMap tmA,
                               const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
  dgemm_tma_dmma_body<Cfg, A_KMAJOR, B_KMAJOR, Cfg::FUSE_EPI>(tmA, tmB, p);
}

// ------------------------------------------------------------------------------------------
// Generic fallback: register-tiled FP64 FMA, any shape / leading dimension / alignment.
// 64x64 C tile, 256 threads, 4x4 per thread, k step 16.
// ------------------------------------------------------------------------------------------
constexpr int SB = 64, SK = 16, SIMT_THREADS = 256;

struct SimtParams {
  int64_t M, N, K, batch;
  double alpha, beta;
  const double *A;
  int64_t lda, stride_a;
  const double *B;
  int64_t ldb, stride_b;
  double *C;
  int64_t ldc, stride_c;
  int transa, transb;
  int tiles_m, tiles_n;
  int64_t total_tiles;
};

__global__ void __launch_bounds__(SIMT_THREADS) dgemm_simt_kernel(const SimtParams p) {
  __shared__ double As[SK][SB + 1];  // [k][m]
  __shared__ double Bs[SK][SB + 1];  // [k][n]
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int tiles_per_problem = p.tiles_m * p.tiles_n;
  for (int64_t tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
    const int64_t b = tile / tiles_per_problem;
    const int r = (int)(tile - b * tiles_per_problem);
    const int64_t m0 = (int64_t)(r % p.tiles_m) * SB, n0 = (int64_t)(r / p.tiles_m) * SB;
    const double *A = p.A + b * p.stride_a;
    const double *B = p.B + b * p.stride_b;
    double *C = p.C + b * p.stride_c;
    double acc[4][4] = {};
    for (int64_t k0 = 0; k0 < p.K; k0 += SK) {
      // cooperative, bounds-checked loads; index order chosen so the contiguous memory
      // direction is the fastest-varying thread index for each transpose case
      for (int idx = threadIdx.x; idx < SB * SK; idx += SIMT_THREADS) {
        int mi, ki;
        if (!p.transa) { mi = idx % SB; ki = idx / SB; } else { ki = idx % SK; mi = idx / SK; }
        const int64_t gm = m0 + mi, gk = k0 + ki;
        double v = 0.0;
        if (gm < p.M && gk < p.K) v = p.transa ? A[gm * p.lda + gk] : A[gk * p.lda + gm];
        As[ki][mi] = v;
      }
      for (int idx = threadIdx.x; idx < SB * SK; idx += SIMT_THREADS) {
        int ni, ki;
        if (!p.transb) { ki = idx % SK; ni = idx / SK; } else { ni = idx % SB; ki = idx / SB; }
        const int64_t gn = n0 + ni, gk = k0 + ki;
        double v = 0.0;
        if (gn < p.N && gk < p.K) v = p.transb ? B[gk * p.ldb + gn] : B[gn * p.ldb + gk];
        Bs[ki][ni] = v;
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < SK; ++kk) {
        double a[4], bv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = As[kk][tx + 16 * i];
#pragma unroll
        for (int j = 0; j < 4; ++j) bv[j] = Bs[kk][ty + 16 * j];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], bv[j], acc[i][j]);
      }
      __syncthreads();
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t col = n0 + ty + 16 * j;
      if (col >= p.N) continue;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int64_t row = m0 + tx + 16 * i;
        if (row >= p.M) continue;
        double v = p.alpha * acc[i][j];
        if (p.beta != 0.0) v += p.beta * C[col * p.ldc + row];
        C[col * p.ldc + row] = v;
      }
    }
  }
}

// Test hook: a kernel that does nothing but hold SMs -- `gridDim.x` CTAs, each filling its SM's
// shared memory so that nothing else fits beside it, spinning for `ns` nanoseconds.
__global__ void hold_sms_kernel(unsigned long long ns) {
  extern __shared__ uint8_t hold_smem[];
  if (threadIdx.x == 0) hold_smem[0] = 1;
  const unsigned long long t0 = globaltimer_ns();
  while (globaltimer_ns() - t0 < ns) {
  }
}

// C = beta * C (beta == 0 writes zeros without reading), m x n, batched.
__global__ void scale_kernel(double *C, int64_t m, int64_t n, int64_t ldc, int64_t stride_c,
                             int64_t batch, double beta) {
  const int64_t per = m * n, total = per * batch;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = idx / per, r = idx - b * per;
    double *ptr = C + b * stride_c + (r / m) * ldc + (r % m);
    *ptr = (beta == 0.0) ? 0.0 : beta * *ptr;
  }
}

// ------------------------------------------------------------------------------------------
// Synthetic inputs: same integer recipe as oracle/dgemm_oracle.c (oracle_uniform_at).
// ------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ULL;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
  return x ^ (x >> 31);
}

__global__ void fill_uniform_kernel(uint64_t seed, uint64_t first_id, double *dst, int64_t count,
                                    int64_t batch) {
  const int64_t total = count * batch;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = idx / count, i = idx - b * count;
    const uint64_t key = splitmix64(seed ^ splitmix64(first_id + (uint64_t)b + 0x632BE59BD9B4E019ULL));
    const uint64_t rnd = splitmix64(key + (uint64_t)i * 0xD1342543DE82EF95ULL);
    const double u = (double)(rnd >> 11) * (1.0 / 9007199254740992.0);
    dst[idx] = 2.0 * u - 1.0;
  }
}

__global__ void fill_uniform_block_kernel(uint64_t seed, uint64_t matrix_id, double *dst, int64_t rows,
                                          int64_t cols, int64_t ld_dst, int64_t row0, int64_t col0,
                                          int64_t global_ld) {
  const int64_t total = rows * cols;
  const uint64_t key = splitmix64(seed ^ splitmix64(matrix_id + 0x632BE59BD9B4E019ULL));
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t j = idx / rows, i = idx - j * rows;
    const uint64_t gi = (uint64_t)((col0 + j) * global_ld + (row0 + i));
    const uint64_t rnd = splitmix64(key + gi * 0xD1342543DE82EF95ULL);
    const double u = (double)(rnd >> 11) * (1.0 / 9007199254740992.0);
    dst[j * ld_dst + i] = 2.0 * u - 1.0;
  }
}

}  // namespace ec
