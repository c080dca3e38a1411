// Development probe (not on the product path): one CTA computes D[128,N] = A[128,K] * B[N,K]^T with operands staged
// by the threads themselves into NON-swizzled canonical UMMA layouts (8x8 "core matrices" of 128 bytes), for
// K-major or MN-major A/B. Used by tests/test_umma_probe_gpu.py to pin the LBO/SBO descriptor semantics that the
// tcgen05 attention kernels rely on (their operands are produced by threads, not by TMA).
#include "../common.cuh"
#include "../ptx.cuh"

namespace {
typedef __nv_bfloat16 bf16;

__device__ __forceinline__ uint64_t make_desc_noswz(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // version 1; layout_type 0 = SWIZZLE_NONE
  return d;
}

// A_g: row-major [128][K] (a_mn = 0) or [K][128] (a_mn = 1); same for B with N.
__global__ void __launch_bounds__(128) umma_probe_kernel(const bf16* __restrict__ Ag, const bf16* __restrict__ Bg,
                                                         float* __restrict__ Dg, int N, int K, int a_mn, int b_mn,
                                                         int swap_lbo_sbo) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int M = 128;
  // core-matrix strides: K-direction groups adjacent (128 B apart), MN-direction groups K/8*128 B apart
  const uint32_t k_str = 128, mnA_str = (K / 8) * 128, mnB_str = (K / 8) * 128;
  uint8_t* sA = smem;
  uint8_t* sB = smem + M * K * 2;
  // element (mn, k) -> byte offset inside the operand tile
  auto off = [&](int mn, int k, bool mn_major, uint32_t mn_str) -> uint32_t {
    const uint32_t base = (mn / 8) * mn_str + (k / 8) * k_str;
    return mn_major ? base + (k % 8) * 16 + (mn % 8) * 2 : base + (mn % 8) * 16 + (k % 8) * 2;
  };
  for (int i = tid; i < M * K; i += 128) {
    const int mn = a_mn ? i % M : i / K, k = a_mn ? i / M : i % K;
    *reinterpret_cast<bf16*>(sA + off(mn, k, a_mn, mnA_str)) = Ag[i];
  }
  for (int i = tid; i < N * K; i += 128) {
    const int mn = b_mn ? i % N : i / K, k = b_mn ? i / N : i % K;
    *reinterpret_cast<bf16*>(sB + off(mn, k, b_mn, mnB_str)) = Bg[i];
  }
  if (tid == 0) { ptx::mbar_init(&bar, 1); ptx::fence_mbar_init(); }
  if (warp == 0) ptx::tmem_alloc<256>(&tmem_slot);
  ptx::fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (tid == 0) {
    const uint32_t idesc = ptx::make_idesc_bf16(128, N, a_mn != 0, b_mn != 0);
    for (int ks = 0; ks < K / 16; ++ks) {
      // one MMA consumes 2 core matrices along K: advance the start address by 2 * k_str
      const uint32_t a_addr = ptx::smem_u32(sA) + ks * 2 * k_str, b_addr = ptx::smem_u32(sB) + ks * 2 * k_str;
      uint32_t lboA = k_str, sboA = mnA_str, lboB = k_str, sboB = mnB_str;
      if (swap_lbo_sbo) { uint32_t t = lboA; lboA = sboA; sboA = t; t = lboB; lboB = sboB; sboB = t; }
      ptx::umma_bf16(tmem, make_desc_noswz(a_addr, lboA, sboA), make_desc_noswz(b_addr, lboB, sboB), idesc, ks > 0);
    }
    ptx::umma_commit(&bar);
  }
  ptx::mbar_wait(&bar, 0);
  ptx::tc_fence_after();
  for (int c = 0; c < N; c += 32) {
    uint32_t r[32];
    ptx::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c, r);
    ptx::tmem_ld_wait();
    for (int j = 0; j < 32 && c + j < N; ++j) Dg[(size_t)tid * N + c + j] = __uint_as_float(r[j]);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc<256>(tmem);
}
}  // namespace

DLB_EXPORT int dlb_umma_probe(const void* A, const void* B, float* D, int N, int K, int a_mn, int b_mn, int swap_lbo_sbo,
                              cudaStream_t stream) {
  DLB_REQUIRE(N % 16 == 0 && N >= 16 && N <= 256 && K % 16 == 0 && K >= 16 && K <= 256, DLB_ERR_SHAPE, "umma_probe: bad N/K");
  const size_t smem = (size_t)(128 + N) * K * 2;
  cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  umma_probe_kernel<<<1, 128, smem, stream>>>((const bf16*)A, (const bf16*)B, D, N, K, a_mn, b_mn, swap_lbo_sbo);
  dlb_count_launch();
  return dlb_check_launch("umma_probe");
}
