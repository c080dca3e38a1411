// Development probe (not on the product path): pins the swizzled head-slice tile layout of attn_sw.cuh on hardware.
// One CTA TMA-loads a 128-row tile X and a 64-row tile Y of head h from packed activations, stages a thread-written
// P[128][64] operand, and computes with the SAME helper functions the attention kernels use
//     D1[128, 64]  = X Y^T            (both K-major: main SWIZZLE_128B k-steps + the SWIZZLE_32B tail step for hd 72)
//     D2[128, HDP] = P Y              (Y read MN-major: N = 64 main + N = 16 tail products)
// plus a streaming mode that measures how fast TMA delivers such tiles (the round-1 16-byte-inner gather reached ~15 B/clk/SM).
#include "../attn_sw.cuh"
#include "../attn_sw_host.cuh"
#include "../common.cuh"

namespace {
typedef __nv_bfloat16 bf16;

template <int HDP>
__global__ void __launch_bounds__(128) sw_probe_kernel(const __grid_constant__ CUtensorMap xm, const __grid_constant__ CUtensorMap xt,
                                                       const __grid_constant__ CUtensorMap ym, const __grid_constant__ CUtensorMap yt,
                                                       const bf16* __restrict__ P, float* __restrict__ D1, float* __restrict__ D2, int h,
                                                       int xrow, int yrow, uint8_t* dump) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int TQ = 128 * HDP * 2, TK = 64 * HDP * 2;
  uint8_t* sX = smem;
  uint8_t* sY = sX + TQ;
  uint8_t* sP = sY + TK;
  __shared__ uint64_t full, done;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) { ptx::mbar_init(&full, 1); ptx::mbar_init(&done, 1); ptx::fence_mbar_init(); }
  if (warp == 0) ptx::tmem_alloc<256>(&tmem_slot);
  for (int c = 0; c < 8; ++c)  // my row of P, 8 chunks of 8 columns
    *reinterpret_cast<uint4*>(sP + attn_sw::p_chunk_off(tid, c)) = *reinterpret_cast<const uint4*>(P + tid * 64 + c * 8);
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (warp == 1) {
    if (ptx::elect_one()) {
      ptx::mbar_expect_tx(&full, TQ + TK);
      attn_sw::tma_tile<HDP, 128>(sX, &xm, &xt, &full, h, xrow);
      attn_sw::tma_tile<HDP, 64>(sY, &ym, &yt, &full, h, yrow);
    }
    __syncwarp();
    ptx::mbar_wait(&full, 0);
    ptx::tc_fence_after();
    attn_sw::mma_scores<HDP, 64>(tmem, ptx::smem_u32(sX), ptx::smem_u32(sY));
    attn_sw::mma_accum<HDP>(tmem + 64, ptx::smem_u32(sP), ptx::smem_u32(sY), false);
    ptx::umma_commit_elect(&done);
  }
  ptx::mbar_wait(&done, 0);
  ptx::tc_fence_after();
  const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
  for (int c = 0; c < 64 + HDP; c += 8) {
    uint32_t r[8];
    ptx::tmem_ld8(tmem + lane_off + c, r);
    ptx::tmem_ld_wait();
    for (int j = 0; j < 8; ++j) {
      if (c < 64) D1[tid * 64 + c + j] = __uint_as_float(r[j]);
      else D2[tid * HDP + (c - 64) + j] = __uint_as_float(r[j]);
    }
  }
  if (dump != nullptr)
    for (int b = tid * 16; b < TQ + TK; b += 128 * 16) *reinterpret_cast<uint4*>(dump + b) = *reinterpret_cast<const uint4*>(smem + b);
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc<256>(tmem);
}

// streaming: every CTA loads `tiles_per_cta` 64-row tiles of head (blockIdx.x % H) through a 4-stage ring
template <int HDP>
__global__ void __launch_bounds__(32) sw_stream_kernel(const __grid_constant__ CUtensorMap xm, const __grid_constant__ CUtensorMap xt, int rows, int H,
                                                       int tiles_per_cta) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int TK = 64 * HDP * 2, NSTG = 4;
  __shared__ uint64_t full[NSTG];
  if (threadIdx.x == 0) {
    for (int i = 0; i < NSTG; ++i) ptx::mbar_init(&full[i], 1);
    ptx::fence_mbar_init();
  }
  __syncwarp();
  const int h = blockIdx.x % H, row_tiles = rows / 64;
  const int t0 = (int)(((long long)blockIdx.x * tiles_per_cta) % row_tiles);
  for (int i = 0; i < tiles_per_cta + NSTG; ++i) {
    if (i >= NSTG) ptx::mbar_wait(&full[(i - NSTG) % NSTG], ((i - NSTG) / NSTG) & 1);
    if (i < tiles_per_cta && ptx::elect_one()) {
      ptx::mbar_expect_tx(&full[i % NSTG], TK);
      attn_sw::tma_tile<HDP, 64>(smem + (i % NSTG) * TK, &xm, &xt, &full[i % NSTG], h, ((t0 + i) % row_tiles) * 64);
    }
    __syncwarp();
  }
}
}  // namespace

// X, Y: bf16 [rows, ld] packed activations (H heads of hd columns from column 0); P: bf16 [128, 64]; D1 fp32 [128, 64];
// D2 fp32 [128, HDP] (HDP = 64 / 80 / 128 for hd <= 64 / <= 80 / <= 128); dump (nullable): raw image of the X and Y tiles.
DLB_EXPORT int dlb_attn_sw_probe(const void* X, const void* Y, const void* P, float* D1, float* D2, int64_t rows, int64_t ld, int H, int hd, int h,
                                 int xrow, int yrow, void* dump, cudaStream_t stream) {
  DLB_REQUIRE(hd % 8 == 0 && hd >= 8 && (hd <= 80 || (hd > 96 && hd <= 128)), DLB_ERR_UNSUPPORTED, "attn_sw_probe: hd %d", hd);
  CUtensorMap xm, xt, ym, yt;
  int rc = attn_sw_host::head_map3(&xm, X, rows, ld, H, hd, 0);
  if (!rc) rc = attn_sw_host::head_map3(&xt, X, rows, ld, H, hd, 1);
  if (!rc) rc = attn_sw_host::head_map3(&ym, Y, rows, ld, H, hd, 0);
  if (!rc) rc = attn_sw_host::head_map3(&yt, Y, rows, ld, H, hd, 1);
  if (rc) return rc;
  const int hdp = hd <= 64 ? 64 : (hd <= 80 ? 80 : 128);
  const size_t smem = (size_t)(128 + 64) * hdp * 2 + 16384 + 1024;
#define DLB_SW_PROBE(HDPV)                                                                                                   \
  {                                                                                                                          \
    cudaError_t e = cudaFuncSetAttribute(sw_probe_kernel<HDPV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);     \
    DLB_REQUIRE(e == cudaSuccess, (int)e, "attn_sw_probe: cudaFuncSetAttribute: %s", cudaGetErrorString(e));                \
    sw_probe_kernel<HDPV><<<1, 128, smem, stream>>>(xm, xt, ym, yt, (const bf16*)P, D1, D2, h, xrow, yrow, (uint8_t*)dump);  \
  }
  if (hdp == 64) DLB_SW_PROBE(64) else if (hdp == 80) DLB_SW_PROBE(80) else DLB_SW_PROBE(128)
#undef DLB_SW_PROBE
  dlb_count_launch();
  return dlb_check_launch("attn_sw_probe");
}

DLB_EXPORT int dlb_attn_sw_stream_probe(const void* X, int64_t rows, int64_t ld, int H, int hd, int grid, int tiles_per_cta, cudaStream_t stream) {
  DLB_REQUIRE(hd % 8 == 0 && hd >= 8 && (hd <= 80 || (hd > 96 && hd <= 128)) && rows % 64 == 0, DLB_ERR_UNSUPPORTED, "attn_sw_stream_probe: hd %d", hd);
  CUtensorMap xm, xt;
  int rc = attn_sw_host::head_map3(&xm, X, rows, ld, H, hd, 0);
  if (!rc) rc = attn_sw_host::head_map3(&xt, X, rows, ld, H, hd, 1);
  if (rc) return rc;
  const int hdp = hd <= 64 ? 64 : (hd <= 80 ? 80 : 128);
  const size_t smem = (size_t)4 * 64 * hdp * 2 + 1024;
#define DLB_SW_STREAM(HDPV)                                                                                                  \
  {                                                                                                                          \
    cudaError_t e = cudaFuncSetAttribute(sw_stream_kernel<HDPV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);    \
    DLB_REQUIRE(e == cudaSuccess, (int)e, "attn_sw_stream_probe: cudaFuncSetAttribute: %s", cudaGetErrorString(e));         \
    sw_stream_kernel<HDPV><<<grid, 32, smem, stream>>>(xm, xt, (int)rows, H, tiles_per_cta);                                 \
  }
  if (hdp == 64) DLB_SW_STREAM(64) else if (hdp == 80) DLB_SW_STREAM(80) else DLB_SW_STREAM(128)
#undef DLB_SW_STREAM
  dlb_count_launch();
  return dlb_check_launch("attn_sw_stream_probe");
}
