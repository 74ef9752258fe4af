// eigencuda_b200/csrc/ozaki_kernels.cuh -- the selectable EC_BACKEND_OZAKI_I8 path: FP64 GEMM
// emulated on the int8 tensor cores of sm_100a (tcgen05.mma.kind::i8, accumulators in TMEM).
//
// Scheme (Ozaki splitting, fixed-point variant).  Every row i of op(A) gets a power-of-two scale
// sa_i = 2^Ei with |a_ip| / sa_i < 1/2, every column j of op(B) a scale sb_j likewise.  The scaled
// entries are expanded in S balanced base-128 digits, d in [-64, 64] (they fit int8):
//       a_ip ~= sa_i * sum_{s<S} dA_s[i,p] * 128^-(s+1)        (truncation <= 1/2 * 128^-S)
// The digit planes are multiplied pairwise EXACTLY in int32 (|sum_p d*d| <= 4225*K), pairs of
// equal weight s+t = d share one TMEM accumulator, and pairs with s+t >= S -- below the
// truncation already made -- are dropped:
//       C_ij ~= alpha * sa_i * sb_j * sum_{d<S} 128^-(d+2) * acc_d[i,j]      (+ beta * C_ij)
// Error bound (stated in DESIGN.md): |E_ij| <= (S+2) * K * 128^-S * sa_i * sb_j, i.e. relative to
// (row max of A) * (column max of B) * K, about 5e-14 * K for S = 7 in the worst case and ~1e-15
// relative Frobenius on well-scaled data.  It is a norm-wise bound per row/column, looser than
// the native FP64 path's for rows or columns whose entries span many orders of magnitude.
// Non-finite inputs poison the row/column scale with NaN.  K <= 65536 per launch keeps int32
// exact for S <= 7 (7 * 65536 * 4225 < 2^31).
//
//   ozaki_absmax_kernel     per-row abs-max bits (atomicMax), coalesced for either operand layout
//   ozaki_split_kernel      digits: x -> 128x -> rint via the 1.5*2^52 trick, all on the FP64
//                           pipe, transposed through smem so that every plane is K-major
//                           (what the UMMA shared-memory descriptor wants), 16-byte stores
//   ozaki_gemm_kernel<S>    persistent, warp-specialised: warp 0 = TMA producer (4D tensor maps,
//                           one box per operand per stage covering all S planes, 64B swizzle),
//                           warp 1 = MMA issuer (S(S+1)/2 * 2 tcgen05.mma per 64-byte k-chunk into
//                           S accumulators of 128x64 s32 = 448 TMEM columns), warps 2-9 = epilogue
//                           (tcgen05.ld, exact power-of-two weighting in FP64, coalesced stores).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "dgemm_kernels.cuh"

namespace ec {
namespace oz {

constexpr int OM = 128;     // CTA tile rows (UMMA M)
constexpr int ON = 64;      // CTA tile columns (UMMA N); S accumulators x 64 columns <= 512
constexpr int OBK = 64;     // bytes of K per stage = two K=32 MMAs; 64-byte swizzle rows
constexpr int OSTAGES = 2;
constexpr int OTHREADS = 320;   // producer warp, MMA warp, 8 epilogue warps

// ---------------------------------------------------------------------------------------------
// splitting
// ---------------------------------------------------------------------------------------------
struct SplitParams {
  const double *X;
  int64_t rs, ks, bstride;   // op(X)(r, p) of problem b lives at X[b*bstride + r*rs + p*ks]
  int64_t R, K, Kp, batch;   // rows, depth, padded plane pitch (multiple of 16 bytes), batch
  int8_t *planes;            // [batch][S][R][Kp]
  double *scale;             // [batch][R], = 2^E
};

// Row abs-max as the bit pattern of |x| (non-negative doubles order like unsigned integers, and
// Inf / NaN sort above every finite value, which is exactly the poisoning wanted).  Same 32 x 128
// tiling as the split kernel: coalesced whichever way the operand is laid out, one atomicMax per
// row per tile.  `amax` must be zeroed before the launch.
constexpr int SPLIT_ROWS = 32, SPLIT_K = 128, SPLIT_THREADS = 256;
constexpr int SPLIT_RS = 137, SPLIT_SS = 17;

template <bool ROW_CONTIG>
__global__ void __launch_bounds__(SPLIT_THREADS) ozaki_absmax_kernel(const SplitParams p, unsigned long long *amax) {
  __shared__ unsigned long long red[SPLIT_THREADS];
  const int64_t ktiles = (p.K + SPLIT_K - 1) / SPLIT_K, rtiles = (p.R + SPLIT_ROWS - 1) / SPLIT_ROWS;
  const int64_t total = ktiles * rtiles * p.batch;
  for (int64_t t = blockIdx.x; t < total; t += gridDim.x) {
    const int64_t b = t / (ktiles * rtiles);
    const int64_t rem = t - b * ktiles * rtiles;
    const int64_t r0 = (rem / ktiles) * SPLIT_ROWS, k0 = (rem % ktiles) * SPLIT_K;
    const double *X = p.X + b * p.bstride;
    // thread -> (row, k-lane): lanes run along the memory-contiguous direction
    const int r = ROW_CONTIG ? (threadIdx.x & 31) : (threadIdx.x >> 3);
    const int kl = ROW_CONTIG ? (threadIdx.x >> 5) : (threadIdx.x & 7);     // 8 k-lanes per row
    unsigned long long m = 0ull;
    if (r0 + r < p.R) {
      const double *xr = X + (r0 + r) * p.rs;
#pragma unroll 4
      for (int i = 0; i < SPLIT_K / 8; ++i) {
        // ROW_CONTIG: a warp reads 32 consecutive rows of one k; else 8 lanes read 8 consecutive k
        const int64_t k = k0 + (ROW_CONTIG ? kl + 8 * i : kl + 8 * i);
        if (k < p.K) {
          const unsigned long long bits = (unsigned long long)__double_as_longlong(fabs(xr[k * p.ks]));
          m = bits > m ? bits : m;
        }
      }
    }
    __syncthreads();
    red[ROW_CONTIG ? (kl * 32 + r) : (r * 8 + kl)] = m;
    __syncthreads();
    if (threadIdx.x < SPLIT_ROWS) {
      const int rr = threadIdx.x;
      unsigned long long mm = 0ull;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const unsigned long long v = red[ROW_CONTIG ? (j * 32 + rr) : (rr * 8 + j)];
        mm = v > mm ? v : mm;
      }
      if (r0 + rr < p.R && mm) atomicMax(amax + b * p.R + r0 + rr, mm);
    }
  }
}

// 2^E from the abs-max bits: |x| * 2^-E < 1/2 for every entry of the row; NaN when the row holds a
// non-finite value; 1 for an all-zero row.  A row whose abs-max is 2^999 or more would need an
// exponent beyond the clamp -- the split would then see |x * 2^-E| >= 1/2 and drop the entry --
// so such a row is poisoned with NaN like a non-finite one (its products are within a factor
// 2^25 of overflowing FP64 anyway) instead of coming out finite and wrong.
__device__ __forceinline__ double scale_from_amax(unsigned long long bits) {
  if (bits >= 0x7ff0000000000000ull) return __longlong_as_double(0x7ff8000000000000LL);
  if (bits == 0ull) return 1.0;
  int e = (int)((bits >> 52) & 0x7ff) - 1022 + 1;
  if (e > 1000) return __longlong_as_double(0x7ff8000000000000LL);
  e = max(-1000, e);
  return __hiloint2double((1023 + e) << 20, 0);
}

// Tile of 32 rows x 128 k per block of 256 threads; staged in smem (row stride 137, 16-element
// segments at stride 17: conflict-free for both the coalesced fill and the per-thread reads).
template <int S, bool ROW_CONTIG>
__global__ void __launch_bounds__(SPLIT_THREADS) ozaki_split_kernel(const SplitParams p, const unsigned long long *amax) {
  __shared__ double tile[SPLIT_ROWS * SPLIT_RS];
  const int64_t ktiles = (p.Kp + SPLIT_K - 1) / SPLIT_K, rtiles = (p.R + SPLIT_ROWS - 1) / SPLIT_ROWS;
  const int64_t total = ktiles * rtiles * p.batch;
  for (int64_t t = blockIdx.x; t < total; t += gridDim.x) {
    const int64_t b = t / (ktiles * rtiles);
    const int64_t rem = t - b * ktiles * rtiles;
    const int64_t r0 = (rem / ktiles) * SPLIT_ROWS, k0 = (rem % ktiles) * SPLIT_K;
    const double *X = p.X + b * p.bstride;
    __syncthreads();
    for (int idx = threadIdx.x; idx < SPLIT_ROWS * SPLIT_K; idx += SPLIT_THREADS) {
      int r, k;
      if (ROW_CONTIG) { r = idx % SPLIT_ROWS; k = idx / SPLIT_ROWS; } else { k = idx % SPLIT_K; r = idx / SPLIT_K; }
      double v = 0.0;
      if (r0 + r < p.R && k0 + k < p.K) v = X[(r0 + r) * p.rs + (k0 + k) * p.ks];
      tile[r * SPLIT_RS + (k >> 4) * SPLIT_SS + (k & 15)] = v;
    }
    __syncthreads();
    const int row = threadIdx.x >> 3, seg = threadIdx.x & 7;
    if (r0 + row < p.R && k0 + seg * 16 < p.Kp) {
      const double sc = scale_from_amax(amax[b * p.R + r0 + row]);
      if (k0 == 0 && seg == 0) p.scale[b * p.R + r0 + row] = sc;
      // 1 / 2^E, exact (a NaN scale makes everything NaN -> digits 0, the epilogue re-poisons)
      const double inv = __hiloint2double((2046 - ((__double2hiint(sc) >> 20) & 0x7ff)) << 20, 0);
      uint32_t w[S][4];
#pragma unroll
      for (int s = 0; s < S; ++s) w[s][0] = w[s][1] = w[s][2] = w[s][3] = 0u;
      const double *src = tile + row * SPLIT_RS + seg * SPLIT_SS;
      const double MAGIC = 6755399441055744.0;  // 1.5 * 2^52: low word of (y + MAGIC) is rint(y)
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        double x = src[i] * inv;                 // |x| < 1/2, exact
        if (!(fabs(x) < 0.5)) x = 0.0;           // NaN / inf rows: keep the digits harmless
#pragma unroll
        for (int s = 0; s < S; ++s) {
          const double y = x * 128.0;            // exact
          const double tq = y + MAGIC;
          const int d = __double2loint(tq);      // rint(y), |d| <= 64
          x = y - (tq - MAGIC);                  // exact remainder, |x| <= 1/2
          w[s][i >> 2] |= ((uint32_t)d & 0xffu) << (8 * (i & 3));
        }
      }
      int8_t *dst = p.planes + ((b * S) * p.R + (r0 + row)) * p.Kp + k0 + seg * 16;
#pragma unroll
      for (int s = 0; s < S; ++s)
        *reinterpret_cast<uint4 *>(dst + (int64_t)s * p.R * p.Kp) = make_uint4(w[s][0], w[s][1], w[s][2], w[s][3]);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// the int8 GEMM on the digit planes
// ---------------------------------------------------------------------------------------------
// Tile order: groups of RASTER_M m-tiles, inside a group m fastest then n, so that the ~148 tiles in
// flight share few A row-panels AND few B column-panels (operand bytes per wave ~ sqrt-like instead
// of all of A): at n = 8192 that is 7 GB instead of 26 GB of HBM reads per product.
constexpr int RASTER_M = 8;
__device__ __forceinline__ void oz_tile_coords(int r, int tiles_m, int tiles_n, int &tm, int &tn) {
  const int per_group = RASTER_M * tiles_n;
  const int grp = r / per_group, in = r - grp * per_group;
  const int gm = min(RASTER_M, tiles_m - grp * RASTER_M);   // rows in this (possibly short) group
  tm = grp * RASTER_M + in % gm;
  tn = in / gm;
}

struct GemmOzParams {
  int64_t M, N, K, batch;
  double alpha, beta;
  double *C;
  int64_t ldc, stride_c;
  const double *sa, *sb;       // row scales [batch_a][M], column scales [batch_b][N]
  int a_batched, b_batched;
  int tiles_m, tiles_n;
  int64_t total_tiles;
};

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0,
                                            int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// K-major operand in 64-byte swizzled rows: 8-row groups 512 B apart (stride byte offset), the
// leading byte offset is unused, descriptor version 1 (sm_100), layout type 4 = SWIZZLE_64B.
__device__ __forceinline__ uint64_t umma_desc_sw64(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)((8 * OBK) >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)4 << 61);
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// 64 consecutive 32-bit TMEM columns of this warp's 32 lanes: thread = row, r[j] = column j.
__device__ __forceinline__ void tmem_ld_x64(uint32_t taddr, uint32_t (&r)[64]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,"
      "%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63}, [%64];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
        "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]),
        "=r"(r[30]), "=r"(r[31]), "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]),
        "=r"(r[40]), "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]), "=r"(r[48]), "=r"(r[49]),
        "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]), "=r"(r[56]), "=r"(r[57]), "=r"(r[58]), "=r"(r[59]),
        "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
      : "r"(taddr)
      : "memory");
}

// 32 consecutive 32-bit TMEM columns of this warp's 32 lanes.
__device__ __forceinline__ void tmem_ld_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
        "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]),
        "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}

template <int S>
struct OzCfg {
  static constexpr int A_BYTES = S * OM * OBK, B_BYTES = S * ON * OBK;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int TMEM_COLS = 512;          // S * 64 <= 512, power of two
  static constexpr int SMEM_BYTES = OSTAGES * STAGE_BYTES + 1024 + 256 + ON * 8;
  static_assert(S * ON <= 512, "accumulators do not fit TMEM");
};

template <int S>
__global__ void __launch_bounds__(OTHREADS, 1)
    ozaki_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                      const GemmOzParams p) {
  using Cfg = OzCfg<S>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = base + OSTAGES * Cfg::STAGE_BYTES;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (OSTAGES + s); };
  const uint32_t tmem_full = bars + 8u * (2 * OSTAGES), tmem_empty = tmem_full + 8, tmem_slot = tmem_full + 16;
  double *colscale = reinterpret_cast<double *>(smem_raw + ((base - smem_u32(smem_raw)) + OSTAGES * Cfg::STAGE_BYTES + 256));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < OSTAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);   // one tcgen05.commit per stage
    }
    mbar_init(tmem_full, 1);        // one tcgen05.commit per tile
    mbar_init(tmem_empty, 8);       // one arrive per epilogue warp
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) {  // TMEM allocation is warp-collective; this warp also frees it
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(Cfg::TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(tmem_slot));

  const int nk = (int)((p.K + OBK - 1) / OBK);
  const int tiles_per_problem = p.tiles_m * p.tiles_n;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int64_t tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const int b = (int)(tile / tiles_per_problem);
        const int r = (int)(tile - (int64_t)b * tiles_per_problem);
        int tm, tn;
        oz_tile_coords(r, p.tiles_m, p.tiles_n, tm, tn);
        const int m0 = tm * OM, n0 = tn * ON;
        const int ba = p.a_batched ? b : 0, bb = p.b_batched ? b : 0;
        for (int kc = 0; kc < nk; ++kc) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          mbar_expect_tx(full_bar(stage), Cfg::STAGE_BYTES);
          const uint32_t sA = base + stage * Cfg::STAGE_BYTES, sB = sA + Cfg::A_BYTES;
          tma_load_4d(sA, &tmA, full_bar(stage), kc * OBK, m0, 0, ba);   // all S planes, one box
          tma_load_4d(sB, &tmB, full_bar(stage), kc * OBK, n0, 0, bb);
          if (++stage == OSTAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0) {
      // s32 accumulate, signed 8-bit A and B, both K-major, M = 128, N in the descriptor per call
      auto idesc_n = [](int n) -> uint32_t {
        return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(OM >> 4) << 24);
      };
      int stage = 0;
      uint32_t phase = 0, tphase = 0;
      for (int64_t tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        mbar_wait(tmem_empty, tphase ^ 1u);   // the epilogue has drained the accumulators
        tc_fence_after();
        for (int kc = 0; kc < nk; ++kc) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sA = base + stage * Cfg::STAGE_BYTES, sB = sA + Cfg::A_BYTES;
          // Plane s of A against planes t = 0..S-1-s of B.  Those B planes are contiguous in smem
          // (one K-major matrix of 64*(S-s) rows) and their products belong to the accumulators
          // d = s..S-1, which are contiguous in TMEM (columns 64*s ..), so ONE MMA of N up to 256
          // does what would otherwise be up to four: 10 MMAs instead of 28 per 32-byte k-step, and
          // the 128-row A tile is re-read from shared memory 10 times instead of 28 (the N = 64
          // form needs 192 B/clk of operand bandwidth; shared memory delivers 128).
#pragma unroll
          for (int kk = 0; kk < OBK / 32; ++kk) {
#pragma unroll
            for (int s = 0; s < S; ++s) {
              const uint64_t da = umma_desc_sw64(sA + s * (OM * OBK) + kk * 32);
              constexpr int dummy = 0; (void)dummy;
#pragma unroll
              for (int c0 = 0; c0 < ON * (S - s); c0 += 256) {
                const int nn = (ON * (S - s) - c0) < 256 ? (ON * (S - s) - c0) : 256;
                const uint64_t db = umma_desc_sw64(sB + (c0 / ON) * (ON * OBK) + kk * 32);
                umma_i8(tmem + s * ON + c0, da, db, idesc_n(nn), (kc > 0 || kk > 0 || s > 0) ? 1u : 0u);
              }
            }
          }
          umma_commit(empty_bar(stage));       // smem slot free once these MMAs have read it
          if (++stage == OSTAGES) { stage = 0; phase ^= 1u; }
        }
        umma_commit(tmem_full);                // accumulators complete
        tphase ^= 1u;
      }
    }
  } else {
    // ===================== epilogue warps (2..9): TMEM -> FP64 -> global =====================
    // Two warps per TMEM lane quarter, each taking 32 of the tile's 64 columns, so that one
    // warp's TMEM-load latency overlaps the other's FP64 work on the same SM sub-partition.
    constexpr int EN = ON / 2;
    const int q = warp & 3;                    // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;          // which 32 columns
    const int row_in_tile = q * 32 + lane;
    const int et = threadIdx.x - 64;           // 0..255
    uint32_t tphase = 0;
    for (int64_t tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const int b = (int)(tile / tiles_per_problem);
      const int r = (int)(tile - (int64_t)b * tiles_per_problem);
      int tm, tn;
      oz_tile_coords(r, p.tiles_m, p.tiles_n, tm, tn);
      const int m0 = tm * OM, n0 = tn * ON;
      // column scales of this tile (visible to all epilogue warps after the named barrier)
      asm volatile("bar.sync 1, 256;" ::: "memory");   // previous tile's readers are done
      if (et < ON) {
        const int64_t col = n0 + et;
        colscale[et] = col < p.N ? p.sb[(p.b_batched ? (int64_t)b * p.N : 0) + col] : 0.0;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      mbar_wait(tmem_full, tphase);
      tc_fence_after();
      double acc[EN];
#pragma unroll
      for (int j = 0; j < EN; ++j) acc[j] = 0.0;
      double wgt = 1.0 / 16384.0;              // 128^-(d+2)
#pragma unroll 1
      for (int d = 0; d < S; ++d) {
        uint32_t v[EN];
        tmem_ld_x32(tmem + ((uint32_t)(q * 32) << 16) + d * ON + half * EN, v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        // int32 -> double without the (slow) conversion unit: 2^52 + 2^31 + v is exactly
        // representable with the biased value in the low mantissa word, one exact DADD removes
        // the offset; everything stays on the FP64 pipe.
#pragma unroll
        for (int j = 0; j < EN; ++j) {
          const double x = __hiloint2double(0x43300000, (int)(v[j] ^ 0x80000000u)) - 4503601774854144.0;
          acc[j] = fma(x, wgt, acc[j]);
        }
        wgt *= (1.0 / 128.0);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tmem_empty);  // the MMA warp may overwrite the accumulators
      tphase ^= 1u;
      const int64_t row = m0 + row_in_tile;
      if (row < p.M) {
        const double rs = p.alpha * p.sa[(p.a_batched ? (int64_t)b * p.M : 0) + row];
        double *Crow = p.C + (int64_t)b * p.stride_c + row;
        const bool use_beta = p.beta != 0.0;
#pragma unroll
        for (int j = 0; j < EN; ++j) {
          const int64_t col = n0 + half * EN + j;
          if (col < p.N) {
            double v = rs * colscale[half * EN + j] * acc[j];
            if (use_beta) v += p.beta * Crow[col * p.ldc];
            Crow[col * p.ldc] = v;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(Cfg::TMEM_COLS));
}

}  // namespace oz
}  // namespace ec
