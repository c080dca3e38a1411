// Swizzled head-slice tiles for the tcgen05 attention kernels (shared by attention_tc.cu and the layout probe sw_probe.cu).
//
// A head slice [rows, hd] of a packed bf16 activation [tokens, ld] (head h = columns [h*hd, (h+1)*hd)) is brought into shared
// memory by TMA through a 3-D tensor map {hd, heads, tokens} with 128-byte-wide boxes, i.e. at full TMA efficiency, instead of
// the 16-byte-inner gather the first tcgen05 attention used (~15 B/clk/SM, the measured bottleneck of round 1):
//
//   main block : box {64 elements, 1 head, 64 rows}, SWIZZLE_128B -> rows of 128 B, 8-row atoms of 1 KB      (offset 0)
//   second main: head dims > 96 (HDP = 128): the same map at element offset 64                                (offset R*128)
//   tail block : HDP = 80 (DiT-XL/2: hd 72): box {16 elements, 1, 64 rows}, SWIZZLE_32B -> rows of 32 B; elements >= hd are
//                out of bounds of dimension 0 and arrive as zeros                                              (offset R*128)
//
// R = 64 (streamed tiles) or 128 (resident tiles = two boxes). Tile bytes = R * HDP * 2 as in the unswizzled layout. The SAME
// bytes serve as a K-major operand (rows = M/N index, head dim = contraction: Q K^T, dO V^T) and as an MN-major operand (rows =
// contraction index, head dim = N: P V, dS K, dS^T Q, P^T dO) — the two canonical readings of one swizzle atom, exactly what
// the GEMM kernel does with its K-major A tiles and MN-major B tiles. Tiles must start on a 1024-byte boundary.
#pragma once
#include "ptx.cuh"

namespace attn_sw {

// UMMA shared-memory descriptor, version 1; layout_type: 2 = SWIZZLE_128B, 4 = SWIZZLE_64B, 6 = SWIZZLE_32B
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(layout_type) << 61;
  return d;
}

template <int HDP>
struct Geo {
  static_assert(HDP == 64 || HDP == 80 || HDP == 128, "swizzled head-slice tiles: padded head dims 64, 80, 128");
  static constexpr int KSTEPS = HDP / 16;               // 16-element steps of a contraction over the head dim
  static constexpr int N_MAIN = HDP == 128 ? 128 : 64;  // columns of the main MN-major product
  static constexpr int N_TAIL = HDP == 80 ? 16 : 0;     // columns of the tail MN-major product
  static constexpr bool HAS_TAIL = HDP == 80;
  static constexpr int BOXES = HDP == 64 ? 1 : 2;       // TMA boxes per 64 rows
};

// K-major operand, step ks of the contraction over the head dim; tile of R rows at shared address a
template <int HDP, int R>
__device__ __forceinline__ uint64_t k_desc(uint32_t a, int ks) {
  if (Geo<HDP>::HAS_TAIL && ks == 4) return make_desc(a + R * 128, 16, 256, 6);
  return make_desc(a + (ks >> 2) * (R * 128) + (ks & 3) * 32, 16, 1024, 2);
}
// MN-major operand (N = head dim, contraction over the R rows), 16-row step ks: main block(s)
template <int HDP, int R>
__device__ __forceinline__ uint64_t mn_desc_main(uint32_t a, int ks) {
  return make_desc(a + ks * 2048, R * 128, 1024, 2);
}
// ... and the 16-column tail block (HDP = 80)
template <int HDP, int R>
__device__ __forceinline__ uint64_t mn_desc_tail(uint32_t a, int ks) {
  return make_desc(a + R * 128 + ks * 512, 16, 256, 6);
}
// thread-written [128 rows][64 columns] bf16 operand (P, dS, P^T, dS^T), K-major, SWIZZLE_128B: 16-column step ks
__device__ __forceinline__ uint64_t p_desc(uint32_t a, int ks) { return make_desc(a + ks * 32, 16, 1024, 2); }

// byte offset of the 16-byte chunk c (8 elements) of row r inside an R-row tile
template <int HDP, int R>
__device__ __forceinline__ uint32_t chunk_off(int r, int c) {
  if (Geo<HDP>::HAS_TAIL && c >= 8) return R * 128 + r * 32 + (((c - 8) ^ ((r >> 2) & 1)) << 4);
  const int blk = c >> 3, cc = c & 7;
  return blk * (R * 128) + (r >> 3) * 1024 + (r & 7) * 128 + ((cc ^ (r & 7)) << 4);
}
// ... and inside a thread-written [128][64] operand tile
__device__ __forceinline__ uint32_t p_chunk_off(int r, int c) { return (r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4); }

// 3-D TMA load of one box (64 rows) of a head slice: coordinates (element offset inside the head, head, token row)
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n" ::"r"(
          ptx::smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(ptx::smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// all boxes of an R-row tile (R = 64 or 128) whose first token row is `row`: main (+ second main / tail)
template <int HDP, int R>
__device__ __forceinline__ void tma_tile(uint8_t* dst, const CUtensorMap* main_map, const CUtensorMap* tail_map, uint64_t* bar, int h, int row) {
#pragma unroll
  for (int i = 0; i < R / 64; ++i) {
    tma_load_3d(dst + i * 64 * 128, main_map, bar, 0, h, row + i * 64);
    if (HDP == 128) tma_load_3d(dst + R * 128 + i * 64 * 128, main_map, bar, 64, h, row + i * 64);
    if (HDP == 80) tma_load_3d(dst + R * 128 + i * 64 * 32, tail_map, bar, 64, h, row + i * 64);
  }
}

// S(128 x RB) (+)= A(128-row K-major tile) * B(RB-row K-major tile)^T over the head dim; whole warp, one elected lane issues
template <int HDP, int RB>
__device__ __forceinline__ void mma_scores(uint32_t d_tmem, uint32_t a_tile, uint32_t b_tile) {
  constexpr uint32_t idesc = ptx::make_idesc_bf16(128, RB, false, false);
#pragma unroll
  for (int ks = 0; ks < Geo<HDP>::KSTEPS; ++ks) ptx::umma_bf16_elect(d_tmem, k_desc<HDP, 128>(a_tile, ks), k_desc<HDP, RB>(b_tile, ks), idesc, ks > 0);
}
// D(128 x HDP) (+)= P(128 x 64, thread-written) * B(64-row tile read MN-major); acc0: accumulate into D from the first step on
template <int HDP>
__device__ __forceinline__ void mma_accum(uint32_t d_tmem, uint32_t p_tile, uint32_t b_tile, bool acc0) {
  constexpr uint32_t idesc_main = ptx::make_idesc_bf16(128, Geo<HDP>::N_MAIN, false, true);
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    const uint32_t acc = (acc0 || ks > 0) ? 1u : 0u;
    ptx::umma_bf16_elect(d_tmem, p_desc(p_tile, ks), mn_desc_main<HDP, 64>(b_tile, ks), idesc_main, acc);
    if constexpr (Geo<HDP>::HAS_TAIL) {
      constexpr uint32_t idesc_tail = ptx::make_idesc_bf16(128, 16, false, true);
      ptx::umma_bf16_elect(d_tmem + 64, p_desc(p_tile, ks), mn_desc_tail<HDP, 64>(b_tile, ks), idesc_tail, acc);
    }
  }
}

}  // namespace attn_sw
