// Development probe: can a 4-D tensor map {8 elements, rows, 16-byte chunks of one head, heads} drop a [64 rows x hd]
// head slice of a packed [rows, ld] bf16 activation into shared memory directly in the core-matrix layout the attention
// kernels use (chunk c, row r at c*1024 + r*16), and how fast does TMA stream such 16-byte-inner boxes?
#include "../common.cuh"
#include "../ptx.cuh"
#include <cudaTypedefs.h>

namespace {
constexpr int NSTG = 4;

// each CTA streams `tiles_per_cta` 64-row tiles of head (blockIdx.x % H); tile 0 of CTA 0 is dumped to `dump`
__global__ void __launch_bounds__(32) tma_gather_probe_kernel(const __grid_constant__ CUtensorMap tm, int rows, int H, int cpr_box,
                                                              int tiles_per_cta, uint8_t* dump) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t full[NSTG];
  const int tile_bytes = 64 * cpr_box * 16;
  if (threadIdx.x == 0) {
    for (int i = 0; i < NSTG; ++i) ptx::mbar_init(&full[i], 1);
    ptx::fence_mbar_init();
  }
  __syncwarp();
  const int h = blockIdx.x % H;
  const int row_tiles = rows / 64;
  int t0 = (int)(((long long)blockIdx.x * tiles_per_cta) % row_tiles);
  for (int i = 0; i < tiles_per_cta + NSTG; ++i) {
    if (i >= NSTG) {
      ptx::mbar_wait(&full[(i - NSTG) % NSTG], ((i - NSTG) / NSTG) & 1);
      if (dump != nullptr && blockIdx.x == 0 && i == NSTG) {
        __syncwarp();
        for (int b = threadIdx.x * 16; b < tile_bytes; b += 32 * 16)
          *reinterpret_cast<uint4*>(dump + b) = *reinterpret_cast<const uint4*>(smem + b);
        __syncwarp();
      }
    }
    if (i < tiles_per_cta && ptx::elect_one()) {
      ptx::mbar_expect_tx(&full[i % NSTG], tile_bytes);
      ptx::tma_load_4d(smem + (i % NSTG) * tile_bytes, &tm, &full[i % NSTG], 0, ((t0 + i) % row_tiles) * 64, 0, h);
    }
    __syncwarp();
  }
}
}  // namespace

// base: bf16 [rows, ld] (16-byte aligned, ld % 8 == 0); head h occupies columns [h*hd, (h+1)*hd). Streams
// grid * tiles_per_cta tiles; dump (may be null) receives CTA 0's first tile image (64 * ceil(hd/16)*2 * 16 bytes).
DLB_EXPORT int dlb_tma_gather_probe(const void* base, int64_t rows, int64_t ld, int H, int hd, int grid, int tiles_per_cta,
                                    void* dump, cudaStream_t stream) {
  static PFN_cuTensorMapEncodeTiled_v12000 enc = nullptr;
  if (!enc) {
    void* fp = nullptr;
    cudaDriverEntryPointQueryResult q;
    const bool ok = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess;
    DLB_REQUIRE(ok, DLB_ERR_DRIVER, "tma_gather_probe: cuTensorMapEncodeTiled unavailable");
    enc = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fp);
  }
  const int cpr = hd / 8, cpr_box = (hd + 15) / 16 * 2;
  CUtensorMap tm;
  cuuint64_t dims[4] = {8, (cuuint64_t)rows, (cuuint64_t)cpr, (cuuint64_t)H};
  cuuint64_t strides[3] = {(cuuint64_t)ld * 2, 16, (cuuint64_t)hd * 2};
  cuuint32_t box[4] = {8, 64, (cuuint32_t)cpr_box, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DLB_REQUIRE(r == CUDA_SUCCESS, DLB_ERR_DRIVER, "tma_gather_probe: cuTensorMapEncodeTiled failed (%d)", (int)r);
  const size_t smem = (size_t)NSTG * 64 * cpr_box * 16;
  cudaError_t e = cudaFuncSetAttribute(tma_gather_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  DLB_REQUIRE(e == cudaSuccess, (int)e, "tma_gather_probe: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  tma_gather_probe_kernel<<<grid, 32, smem, stream>>>(tm, (int)rows, H, cpr_box, tiles_per_cta, (uint8_t*)dump);
  dlb_count_launch();
  return dlb_check_launch("tma_gather_probe");
}
