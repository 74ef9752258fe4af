// tools/umma_probe.cu -- first-contact probe for the int8 tcgen05 path (not part of the product):
// one CTA computes D(128 x N, s32) = A(128 x K, s8, K-major) * B(N x K, s8, K-major)^T with TMA
// (64B swizzle) -> smem -> tcgen05.mma.kind::i8 -> TMEM -> tcgen05.ld, and the host checks it
// against a CPU product.  Validates the shared-memory descriptor, the instruction descriptor, the
// TMEM allocation protocol and the accumulator layout before the Ozaki kernel is built on them.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

constexpr int M = 128, N = 64, BK = 64;   // BK bytes of K per stage (= 2 MMAs of K=32)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc_sw64(uint32_t saddr) {
  // K-major, SWIZZLE_64B: 8-row groups of 64-byte rows are 512 B apart (SBO), LBO unused
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);            // start address, bits [0,14)
  d |= (uint64_t)0 << 16;                             // leading byte offset
  d |= (uint64_t)((8 * BK) >> 4) << 32;               // stride byte offset, bits [32,46)
  d |= (uint64_t)1 << 46;                             // version = 1 (sm_100)
  d |= (uint64_t)4 << 61;                             // layout type: SWIZZLE_64B
  return d;
}

__global__ void __launch_bounds__(128, 1)
probe(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int32_t *out, int K) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  const uint32_t sA = base, sB = base + M * BK;         // 8 KiB + 4 KiB
  const uint32_t bar_full = base + M * BK + N * BK, bar_mma = bar_full + 8, tmem_slot = bar_full + 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_full));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_mma));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {  // TMEM allocation is warp-collective; 64 columns of 32-bit accumulators
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(64));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(tmem_slot));

  const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
  uint32_t phase = 0;
  for (int k0 = 0; k0 < K; k0 += BK) {
    if (threadIdx.x == 0) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_full), "r"((M + N) * BK) : "memory");
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                   ::"r"(sA), "l"(&tmA), "r"(bar_full), "r"(k0), "r"(0) : "memory");
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                   ::"r"(sB), "l"(&tmB), "r"(bar_full), "r"(k0), "r"(0) : "memory");
      // wait for the tiles
      uint32_t ok = 0;
      while (!ok) asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0,1,0,p;}" : "=r"(ok) : "r"(bar_full), "r"(phase) : "memory");
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      for (int kk = 0; kk < BK / 32; ++kk) {
        const uint64_t da = make_desc_sw64(sA + kk * 32), db = make_desc_sw64(sB + kk * 32);
        const uint32_t accumulate = (k0 > 0 || kk > 0) ? 1u : 0u;
        asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;}"
                     ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_mma) : "memory");
      ok = 0;   // MMAs done reading smem (and writing TMEM) before the next loads / the epilogue
      while (!ok) asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0,1,0,p;}" : "=r"(ok) : "r"(bar_mma), "r"(phase) : "memory");
      phase ^= 1;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // epilogue: warp w reads TMEM lanes 32w..32w+31 (row = lane), 64 columns
  uint32_t r[64];
  const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
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
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  const int row = warp * 32 + lane;
  for (int j = 0; j < N; ++j) out[row * N + j] = (int32_t)r[j];
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64));
}

typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                              const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                              CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  const int K = 256;
  std::vector<int8_t> A(M * K), B(N * K);
  srand(1);
  for (auto &x : A) x = (int8_t)(rand() % 129 - 64);
  for (auto &x : B) x = (int8_t)(rand() % 129 - 64);
  int8_t *dA, *dB;
  int32_t *dO;
  CK(cudaMalloc(&dA, A.size())); CK(cudaMalloc(&dB, B.size())); CK(cudaMalloc(&dO, M * N * 4));
  CK(cudaMemcpy(dA, A.data(), A.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), B.size(), cudaMemcpyHostToDevice));
  void *fp = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
  encode_fn enc = (encode_fn)fp;
  CUtensorMap tmA, tmB;
  auto mk = [&](CUtensorMap *tm, void *p, int rows) {
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows}, strides[1] = {(cuuint64_t)K};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)rows}, es[2] = {1, 1};
    return enc(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, p, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  };
  if (mk(&tmA, dA, M) != CUDA_SUCCESS || mk(&tmB, dB, N) != CUDA_SUCCESS) { printf("encode failed\n"); return 1; }
  const int smem = (M + N) * BK + 1024 + 64;
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  probe<<<1, 128, smem>>>(tmA, tmB, dO, K);
  CK(cudaDeviceSynchronize());
  std::vector<int32_t> O(M * N);
  CK(cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost));
  long bad = 0;
  for (int i = 0; i < M; ++i)
    for (int j = 0; j < N; ++j) {
      int32_t want = 0;
      for (int p = 0; p < K; ++p) want += (int32_t)A[i * K + p] * (int32_t)B[j * K + p];
      if (want != O[i * N + j]) { if (bad < 5) printf("mismatch (%d,%d): got %d want %d\n", i, j, O[i * N + j], want); ++bad; }
    }
  printf("umma probe: %ld mismatches of %d\n", bad, M * N);
  return bad ? 2 : 0;
}
