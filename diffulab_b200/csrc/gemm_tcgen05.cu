// Persistent warp-specialised bf16 GEMM for sm_100a: TMA -> shared (128B swizzle) -> tcgen05.mma (TMEM fp32
// accumulators, double buffered) -> tcgen05.ld epilogue -> swizzled staging -> TMA store / TMA reduce-add.
//
//   C[M,N] (+)= A[M,K] * B[N,K]^T (+ bias[N])
//
// Replaces the cuBLASLt calls behind nn.Linear / nn.Conv2d(k=s=p) on the reference hot path
// (reference: src/diffulab/networks/denoisers/mmdit.py:70-73,260-264,539,697-699 and their autograd mirrors).
// Operand "major" flags let one kernel serve forward (K-major x K-major), dgrad (K-major x MN-major: B is the
// row-major weight read along its other dimension) and wgrad (MN-major x MN-major: dY^T * X straight from the
// row-major activations), so no transposed copies are ever materialised.
#include "common.cuh"
#include "ptx.cuh"
#include <cudaTypedefs.h>
#include <cstdlib>
#include <mutex>
#include <unordered_map>

namespace {

constexpr int BM = 128;            // UMMA M (cta_group::1)
constexpr int BK = 64;             // one 128-byte swizzle atom of bf16 along K
constexpr int UMMA_K = 16;
constexpr int A_TILE_BYTES = BM * BK * 2;            // 16 KB
constexpr int MN_BLOCK_BYTES = 64 * BK * 2;          // one 64(MN) x 64(K) TMA box of an MN-major operand: 8 KB
constexpr int EPI_WARPS = 4;
constexpr int EPI_BUF_BYTES = 32 * 128;              // 32 rows x 128 B per warp per buffer
constexpr int EPI_BYTES = EPI_WARPS * 2 * EPI_BUF_BYTES;  // 32 KB
constexpr int SMEM_LIMIT = 232448;                   // 227 KB opt-in maximum per CTA

template <int BN, int EPIB = EPI_BYTES>
struct Cfg {
  static constexpr int B_TILE_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_TILE_BYTES + B_TILE_BYTES;
  static constexpr int STAGES_RAW = (SMEM_LIMIT - 1024 - 256 - EPIB) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr int TMEM_COLS = (2 * BN <= 128) ? 128 : (2 * BN <= 256 ? 256 : 512);
  static constexpr int SMEM_BYTES = 1024 + STAGES * STAGE_BYTES + EPIB + 256;
  static_assert(STAGES >= 3, "pipeline too shallow");
  static_assert(SMEM_BYTES <= SMEM_LIMIT, "shared memory budget exceeded");
};

enum { OUT_BF16 = 0, OUT_F32 = 1, OUT_F32_ADD = 2 };
enum { EPI_NONE = 0, EPI_SWIGLU = 1, EPI_SWIGLU_BWD = 2 };
constexpr int EPI_BWD_BYTES = EPI_WARPS * 2 * 2 * EPI_BUF_BYTES;  // per warp: 2 chunk slots x {a, g} tiles = 64 KB

struct TileCoord {
  int m_blk, n_blk, kb0, kb1, ks;
};
__device__ __forceinline__ TileCoord decode_tile(int t, int mb, int nb, int kb_total, int kb_per) {
  TileCoord c;
  const int per_split = mb * nb;
  c.ks = t / per_split;
  const int r = t - c.ks * per_split;
  c.m_blk = r / nb;
  c.n_blk = r - c.m_blk * nb;
  c.kb0 = c.ks * kb_per;
  c.kb1 = min(c.kb0 + kb_per, kb_total);
  return c;
}

// SWIGLU_BWD: TMA-prefetch the (a, g) tiles of H that this warp's rows of the given output tile will need
template <int BN>
__device__ __forceinline__ void issue_h_tile_loads(const CUtensorMap& tmC2, uint8_t* sE, uint64_t* lbar, int ew, int lane, int M, int N,
                                                   int r0, int n_blk, bool wait_stores, int slot_mask = 3) {
  if (lane == 0 && r0 >= 0 && r0 < M) {
    if (wait_stores) ptx::tma_wait_group_read<0>();  // my committed staged stores have been read: their slots are free
    uint8_t* wb_ = sE + ew * (4 * EPI_BUF_BYTES);
#pragma unroll
    for (int slot = 0; slot < 2; ++slot) {
      if (!((slot_mask >> slot) & 1)) continue;
      const int col = n_blk * BN + slot * 64;
      uint64_t* b = &lbar[ew * 2 + slot];
      ptx::mbar_expect_tx(b, 2 * EPI_BUF_BYTES);
      ptx::tma_load_2d(wb_ + slot * 2 * EPI_BUF_BYTES, &tmC2, b, col, r0);
      ptx::tma_load_2d(wb_ + slot * 2 * EPI_BUF_BYTES + EPI_BUF_BYTES, &tmC2, b, N + col, r0);
    }
  }
  __syncwarp();
}

// ---------------------------------------------------------------------------------------------------------
// Epilogue of one accumulator tile for one epilogue warp (shared by the single-CTA and the CTA-pair kernels):
// this warp's 32 TMEM lanes x the tile's columns -> registers -> 128B-swizzled staging -> TMA store / reduce-add,
// with the optional fused SwiGLU forward / backward math (see the kernel comments).
// ---------------------------------------------------------------------------------------------------------
template <int BN, int OUT, int EPI>
__device__ __forceinline__ void epilogue_tile(const CUtensorMap& tmC, const CUtensorMap& tmC2, const float* __restrict__ bias,
                                              bool add_bias, uint8_t* sE, uint64_t* lbar, int ew, int lane, int M, int N, int row0,
                                              int n_blk, uint32_t taddr, int& ebuf, uint32_t& hphase, int next_row0 = -1,
                                              int next_n_blk = 0, int my_slot = -1) {
  constexpr bool SWIGLU = EPI == EPI_SWIGLU, SWIGLU_BWD = EPI == EPI_SWIGLU_BWD;
  constexpr int BNT = SWIGLU ? BN / 2 : BN;
  constexpr int CH = (OUT == OUT_BF16) ? 64 : 32;  // columns per 128-byte staging row
  uint8_t* ebase = sE + ew * (2 * EPI_BUF_BYTES);
  (void)ebase; (void)lbar; (void)hphase; (void)bias; (void)add_bias;
  if constexpr (SWIGLU_BWD) {
    uint8_t* wbase = sE + ew * (4 * EPI_BUF_BYTES);  // [slot][a | g] tiles of 32 rows x 128 B (128B-swizzled, as TMA lays them)
    uint64_t* lb = lbar + ew * 2;
    const bool rows_ok = row0 < M;
#pragma unroll 1
    for (int slot = 0; slot < 2; ++slot) {
      if (my_slot >= 0 && slot != my_slot) continue;  // 8-warp epilogue: this warp owns one 64-column slot of its 32 rows
      const int col0 = n_blk * BN + slot * 64;
      uint8_t* abuf = wbase + slot * 2 * EPI_BUF_BYTES;
      uint8_t* gbuf = abuf + EPI_BUF_BYTES;
      uint32_t r[64];
      ptx::tmem_ld32(taddr + slot * 64, r);
      ptx::tmem_ld32(taddr + slot * 64 + 32, r + 32);
      ptx::tmem_ld_wait();
      if (rows_ok) ptx::mbar_wait(&lb[slot], hphase);
      uint8_t* rowa = abuf + lane * 128;
      uint8_t* rowg = gbuf + lane * 128;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int off = (j ^ (lane & 7)) << 4;
        const bf16x8 av = *reinterpret_cast<const bf16x8*>(rowa + off), gv = *reinterpret_cast<const bf16x8*>(rowg + off);
        // packed fp32 pairs: ~10 issue slots per element instead of ~20. Neutral in the train step (same-box A/B 11.8 vs 12.0 ms):
        // the epilogue is bound by the latency of the H-tile loads and of the stores they wait behind, not by issue. A variant that
        // prefetched H into registers with plain loads (no load behind a store) measured 0.7 ms SLOWER (uncoalesced row loads).
        bf16x8 dav, dgv;
        const f32x2 half2 = make_f32x2(0.5f, 0.5f), one2 = make_f32x2(1.f, 1.f), mone2 = make_f32x2(-1.f, -1.f);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const f32x2 a2 = unpack2(av.u[i]), g2 = unpack2(gv.u[i]);
          // d(act) as the bf16 tensor autocast would hold
          const f32x2 go2 = unpack2(pack2(make_f32x2(__uint_as_float(r[8 * j + 2 * i]), __uint_as_float(r[8 * j + 2 * i + 1]))));
          float h0, h1, t0, t1;
          split_f32x2(mul2(a2, half2), h0, h1);
          asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(h0));
          asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(h1));
          const f32x2 sg2 = fma2(make_f32x2(t0, t1), half2, half2);      // sigmoid(a)
          const f32x2 t2 = fma2(a2, fma2(sg2, mone2, one2), one2);       // 1 + a (1 - sigmoid(a))
          dav.u[i] = pack2(mul2(mul2(mul2(go2, g2), sg2), t2));
          dgv.u[i] = pack2(mul2(mul2(go2, a2), sg2));
        }
        *reinterpret_cast<bf16x8*>(rowa + off) = dav;
        *reinterpret_cast<bf16x8*>(rowg + off) = dgv;
      }
      ptx::fence_proxy_async_smem();
      __syncwarp();
      // slot 0 of the NEXT tile is requested as soon as this tile's slot-0 stores have been read (they were committed a
      // whole chunk ago), not after the tile: the main loop of a K = 1152 tile is shorter than the TMA load latency
      if (slot == 1 && my_slot < 0) issue_h_tile_loads<BN>(tmC2, sE, lbar, ew, lane, M, N, next_row0, next_n_blk, true, 1);
      if (lane == 0 && rows_ok) {
        ptx::tma_store_2d(&tmC, abuf, col0, row0);
        ptx::tma_store_2d(&tmC, gbuf, N + col0, row0);
        ptx::tma_commit_group();
      }
    }
    if (rows_ok) hphase ^= 1;  // the barriers are only armed for tiles whose rows exist
  } else if constexpr (SWIGLU) {
    // 64-column chunks of the two halves: h_a -> tmC(col), h_g -> tmC(N + col), silu(a) * g -> tmC2(col)
    auto stage_store = [&](const uint32_t* packed, const CUtensorMap* tm, int col) {
      uint8_t* buf = ebase + ebuf * EPI_BUF_BYTES;
      if (lane == 0) ptx::tma_wait_group_read<1>();
      __syncwarp();
      uint8_t* rowp = buf + lane * 128;
#pragma unroll
      for (int j = 0; j < 8; ++j)
        *reinterpret_cast<uint4*>(rowp + ((j ^ (lane & 7)) << 4)) = make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
      ptx::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0 && row0 < M) {
        ptx::tma_store_2d(tm, buf, col, row0);
        ptx::tma_commit_group();
      }
      ebuf ^= 1;
    };
    auto load_pack = [&](uint32_t tcol, int bias_col, uint32_t* packed) {  // 64 fp32 columns (+ bias) -> 32 packed bf16 pairs
#pragma unroll
      for (int hlf = 0; hlf < 2; ++hlf) {
        uint32_t r[32];
        ptx::tmem_ld32(taddr + tcol + hlf * 32, r);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float v0 = __uint_as_float(r[i]), v1 = __uint_as_float(r[i + 1]);
          if (add_bias) { v0 += __ldg(bias + bias_col + hlf * 32 + i); v1 += __ldg(bias + bias_col + hlf * 32 + i + 1); }
          packed[hlf * 16 + i / 2] = pack_bf16x2(v0, v1);
        }
      }
    };
#pragma unroll 1
    for (int c = 0; c < BNT; c += 64) {
      const int col0 = n_blk * BNT + c;
      if (col0 >= N) break;
      uint32_t pa[32], pg[32];
      load_pack(c, col0, pa);
      stage_store(pa, &tmC, col0);
      load_pack(BNT + c, N + col0, pg);
      stage_store(pg, &tmC, N + col0);
#pragma unroll
      for (int i = 0; i < 32; ++i) {  // out = bf16(bf16(silu(a)) * g) on the bf16-rounded pre-activations (autocast numerics)
        const float2 a = unpack_bf16x2(pa[i]), g = unpack_bf16x2(pg[i]);
        pa[i] = pack_bf16x2(bf16_round(silu_fast(a.x)) * g.x, bf16_round(silu_fast(a.y)) * g.y);
      }
      stage_store(pa, &tmC2, col0);
    }
  } else {
#pragma unroll 1
  for (int c = 0; c < BN; c += CH) {
    const int col0 = n_blk * BN + c;
    if (col0 >= N) break;
    uint32_t r[CH];
    ptx::tmem_ld32(taddr + c, r);
    if constexpr (CH == 64) ptx::tmem_ld32(taddr + c + 32, r + 32);
    ptx::tmem_ld_wait();
    if (add_bias) {
#pragma unroll
      for (int i = 0; i < CH; ++i) {
        const float bv = (col0 + i < N) ? __ldg(bias + col0 + i) : 0.f;
        r[i] = __float_as_uint(__uint_as_float(r[i]) + bv);
      }
    }
    uint8_t* buf = ebase + ebuf * EPI_BUF_BYTES;
    if (lane == 0) ptx::tma_wait_group_read<1>();  // the store that last read this buffer is done
    __syncwarp();
    uint8_t* rowp = buf + lane * 128;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      uint4 v;
      if constexpr (OUT == OUT_BF16) {
        v.x = pack_bf16x2(__uint_as_float(r[8 * j + 0]), __uint_as_float(r[8 * j + 1]));
        v.y = pack_bf16x2(__uint_as_float(r[8 * j + 2]), __uint_as_float(r[8 * j + 3]));
        v.z = pack_bf16x2(__uint_as_float(r[8 * j + 4]), __uint_as_float(r[8 * j + 5]));
        v.w = pack_bf16x2(__uint_as_float(r[8 * j + 6]), __uint_as_float(r[8 * j + 7]));
      } else {
        v.x = r[4 * j + 0]; v.y = r[4 * j + 1]; v.z = r[4 * j + 2]; v.w = r[4 * j + 3];
      }
      *reinterpret_cast<uint4*>(rowp + ((j ^ (lane & 7)) << 4)) = v;  // 128B swizzle: chunk ^= row % 8
    }
    ptx::fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0 && row0 < M) {
      if constexpr (OUT == OUT_F32_ADD) ptx::tma_reduce_add_2d(&tmC, buf, col0, row0);
      else ptx::tma_store_2d(&tmC, buf, col0, row0);
      ptx::tma_commit_group();
    }
    ebuf ^= 1;
  }
  }
}

// SWIGLU variant (fc1 of the packed-SwiGLU MLP, reference nn.py:478-486 + mmdit.py:260-264): the weight is [2F, K]
// with the silu-ed half in rows [0, F) and the multiplied half in rows [F, 2F). A tile then covers BN/2 columns of EACH
// half (two TMA boxes of BN/2 weight rows), so the accumulator holds matching (a, g) column pairs and the epilogue
// writes the pre-activation h = [a | g] (saved for backward) AND silu(a) * g in one pass. N = F here.
//
// SWIGLU_BWD variant (dgrad of fc2 fused with the SwiGLU backward): the accumulator tile is d(act)[128 x 128]; the
// epilogue warps TMA-load the matching (a, g) tiles of the saved pre-activation H (prefetched during the tile's
// main loop), compute da = dact * g * silu'(a), dg = dact * silu(a) in place and TMA-store them to dH[:, n] / dH[:, F+n].
// d(act) is never written to memory. N = F here.
// (The SWIGLU_BWD variant runs EIGHT epilogue warps, 384 threads: its epilogue does ~22 instructions per element and with
//  four warps took 2.4x the main loop of a K = 1152 tile; warps 8-11 share the TMEM lane quarters of warps 4-7 and take the
//  second 64-column slot.)
template <int BN, bool A_MN, bool B_MN, int OUT, int EPI = EPI_NONE>
__global__ void __launch_bounds__(EPI == EPI_SWIGLU_BWD ? 384 : 256, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmC2,
                    const float* __restrict__ bias, int M, int N, int K, int split_k) {
  constexpr bool SWIGLU = EPI == EPI_SWIGLU, SWIGLU_BWD = EPI == EPI_SWIGLU_BWD;
  static_assert(!SWIGLU || (!A_MN && !B_MN && OUT == OUT_BF16 && BN == 256), "SWIGLU epilogue: K-major bf16 128x256 tiles only");
  static_assert(!SWIGLU_BWD || (!A_MN && B_MN && OUT == OUT_BF16 && BN == 128), "SWIGLU_BWD epilogue: dgrad bf16 128x128 tiles only");
  constexpr int BNT = SWIGLU ? BN / 2 : BN;  // output columns (of each half) per tile
  constexpr int EPIB = SWIGLU_BWD ? EPI_BWD_BYTES : EPI_BYTES;
  using C = Cfg<BN, EPIB>;
  constexpr int NST = C::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = sA + NST * A_TILE_BYTES;
  uint8_t* sE = sB + NST * C::B_TILE_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(sE + EPIB);
  uint64_t* empty = full + NST;
  uint64_t* tfull = empty + NST;
  uint64_t* tempty = tfull + 2;
  uint64_t* lbar = tempty + 2;  // SWIGLU_BWD: [EPI_WARPS][2] "h tiles landed" barriers
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(lbar + 2 * EPI_WARPS);

  const int lane = threadIdx.x & 31;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // warp-uniform for the compiler (role dispatch + uniform MMA issue)
  const int mb = (M + BM - 1) / BM, nb = (N + BNT - 1) / BNT;
  const int kb_total = (K + BK - 1) / BK;
  const int kb_per = (kb_total + split_k - 1) / split_k;
  const int tiles = mb * nb * split_k;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmB);
    ptx::prefetch_tmap(&tmC);
    if constexpr (SWIGLU || SWIGLU_BWD) ptx::prefetch_tmap(&tmC2);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < NST; ++i) {
      ptx::mbar_init(&full[i], 1);
      ptx::mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&tfull[i], 1);
      ptx::mbar_init(&tempty[i], SWIGLU_BWD ? 2 * EPI_WARPS : EPI_WARPS);
    }
    for (int i = 0; i < 2 * EPI_WARPS; ++i) ptx::mbar_init(&lbar[i], 1);
    ptx::fence_mbar_init();
  }
  if (warp == 2) ptx::tmem_alloc<C::TMEM_COLS>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer (whole warp loops on uniform values, one elected lane issues) =====================
    {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
        const TileCoord tc = decode_tile(t, mb, nb, kb_total, kb_per);
        for (int kb = tc.kb0; kb < tc.kb1; ++kb) {
          ptx::mbar_wait(&empty[stage], phase ^ 1);
          if (ptx::elect_one()) {
            ptx::mbar_expect_tx(&full[stage], C::STAGE_BYTES);
            uint8_t* a = sA + stage * A_TILE_BYTES;
            uint8_t* b = sB + stage * C::B_TILE_BYTES;
            if constexpr (!A_MN) {
              ptx::tma_load_2d(a, &tmA, &full[stage], kb * BK, tc.m_blk * BM);
            } else {
#pragma unroll
              for (int i = 0; i < BM / 64; ++i)
                ptx::tma_load_2d(a + i * MN_BLOCK_BYTES, &tmA, &full[stage], tc.m_blk * BM + i * 64, kb * BK);
            }
            if constexpr (SWIGLU) {
              ptx::tma_load_2d(b, &tmB, &full[stage], kb * BK, tc.n_blk * BNT);                       // rows of the silu-ed half
              ptx::tma_load_2d(b + BNT * BK * 2, &tmB, &full[stage], kb * BK, N + tc.n_blk * BNT);    // matching rows of the other half
            } else if constexpr (!B_MN) {
              ptx::tma_load_2d(b, &tmB, &full[stage], kb * BK, tc.n_blk * BN);
            } else {
#pragma unroll
              for (int i = 0; i < BN / 64; ++i)
                ptx::tma_load_2d(b + i * MN_BLOCK_BYTES, &tmB, &full[stage], tc.n_blk * BN + i * 64, kb * BK);
            }
          }
          __syncwarp();
          if (++stage == NST) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (whole warp, warp-uniform code; one elected lane issues) =====================
    {
      constexpr uint32_t idesc = ptx::make_idesc_bf16(BM, BN, A_MN, B_MN);
      // K-major: a k-step is 16 elements (32 B) inside the swizzle atom. MN-major: 16 k-rows (2 KB). In 16-byte units:
      constexpr uint64_t A_KSTEP = (A_MN ? UMMA_K * 128 : UMMA_K * 2) >> 4, B_KSTEP = (B_MN ? UMMA_K * 128 : UMMA_K * 2) >> 4;
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t a_base = ptx::smem_u32(sA), b_base = ptx::smem_u32(sB);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
        const TileCoord tc = decode_tile(t, mb, nb, kb_total, kb_per);
        ptx::mbar_wait(&tempty[as], aphase ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_u + as * BN;
        for (int kb = tc.kb0; kb < tc.kb1; ++kb) {
          ptx::mbar_wait(&full[stage], phase);
          ptx::tc_fence_after();
          const uint32_t a_addr = a_base + stage * A_TILE_BYTES;
          const uint32_t b_addr = b_base + stage * C::B_TILE_BYTES;
          const uint64_t adesc0 = A_MN ? ptx::make_smem_desc_sw128(a_addr, MN_BLOCK_BYTES, 1024) : ptx::make_smem_desc_sw128(a_addr, 16, 1024);
          const uint64_t bdesc0 = B_MN ? ptx::make_smem_desc_sw128(b_addr, MN_BLOCK_BYTES, 1024) : ptx::make_smem_desc_sw128(b_addr, 16, 1024);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k)
            ptx::umma_bf16_elect(d_tmem, adesc0 + k * A_KSTEP, bdesc0 + k * B_KSTEP, idesc, (kb > tc.kb0 || k > 0) ? 1u : 0u);
          ptx::umma_commit_elect(&empty[stage]);  // frees the smem slot once these MMAs retire
          if (++stage == NST) { stage = 0; phase ^= 1; }
        }
        ptx::umma_commit_elect(&tfull[as]);  // accumulator complete -> epilogue
        as ^= 1;
        if (as == 0) aphase ^= 1;
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue: TMEM -> registers -> swizzled smem -> TMA store =====================
    const int ew = (warp - 4) & 3;  // == warp % 4: the TMEM lane quarter this warp may access
    const int my_slot = SWIGLU_BWD ? (warp - 4) >> 2 : -1;
    const int slot_mask = SWIGLU_BWD ? 1 << ((warp - 4) >> 2) : 3;
    int ebuf = 0;
    int as = 0;
    uint32_t aphase = 0;
    [[maybe_unused]] uint32_t hphase = 0;
    [[maybe_unused]] auto issue_h_loads = [&](int tn, bool wait_stores, int slot_mask) {
      if constexpr (SWIGLU_BWD) {
        const TileCoord tcn = decode_tile(tn, mb, nb, kb_total, kb_per);
        issue_h_tile_loads<BN>(tmC2, sE, lbar, ew, lane, M, N, tcn.m_blk * BM + ew * 32, tcn.n_blk, wait_stores, slot_mask);
      }
    };
    if constexpr (SWIGLU_BWD) {
      if ((int)blockIdx.x < tiles) issue_h_loads(blockIdx.x, false, slot_mask);
    }
    for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
      const TileCoord tc = decode_tile(t, mb, nb, kb_total, kb_per);
      ptx::mbar_wait(&tfull[as], aphase);
      ptx::tc_fence_after();
      const int row0 = tc.m_blk * BM + ew * 32;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + as * BN;
      const bool add_bias = (bias != nullptr) && (tc.ks == 0);
      int next_row0 = -1, next_n_blk = 0;
      if constexpr (SWIGLU_BWD) {
        if (t + (int)gridDim.x < tiles) {
          const TileCoord tcn = decode_tile(t + gridDim.x, mb, nb, kb_total, kb_per);
          next_row0 = tcn.m_blk * BM + ew * 32;
          next_n_blk = tcn.n_blk;
        }
      }
      epilogue_tile<BN, OUT, EPI>(tmC, tmC2, bias, add_bias, sE, lbar, ew, lane, M, N, row0, tc.n_blk, taddr, ebuf, hphase, next_row0, next_n_blk, my_slot);
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tempty[as]);
      as ^= 1;
      if (as == 0) aphase ^= 1;
      if constexpr (SWIGLU_BWD) {  // h tiles of my next output tile: they land while its main loop runs
        if (t + (int)gridDim.x < tiles) issue_h_loads(t + gridDim.x, true, slot_mask);
      }
    }
    if (lane == 0) ptx::tma_wait_group<0>();
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) ptx::tmem_dealloc<C::TMEM_COLS>(tmem_base);
}

// ---------------------------------------------------------------------------------------------------------
// CTA-pair variant (cta_group::2): a cluster of two CTAs (one TPC) computes a 256 x BN tile with ONE stream of
// tcgen05.mma.cta_group::2 issued by the even CTA. Each CTA stages its own 128 rows of A and only HALF of the B tile
// (BN/2 rows), so the L2 -> shared-memory traffic per output element drops by a third against the single-CTA kernel —
// the single-CTA kernel is bound by exactly that traffic. Protocol:
//   full[s]   (even CTA)  1 arrival (its producer's expect_tx) + the bytes of BOTH CTAs' TMA loads (.cta_group::2 loads
//                         credit the even CTA's barrier)
//   empty[s]  (each CTA)  tcgen05.commit multicast to both CTAs: stage s may be refilled
//   tfull[a]  (each CTA)  tcgen05.commit multicast: accumulator stage a is complete; every CTA drains its own 128 lanes
//   tempty[a] (even CTA)  2 x EPI_WARPS arrivals (the odd CTA's epilogue warps arrive remotely)
// ---------------------------------------------------------------------------------------------------------
template <int BN, int EPIB = EPI_BYTES>
struct Cfg2 {
  static constexpr int B_HALF_BYTES = (BN / 2) * BK * 2;
  static constexpr int STAGE_BYTES = A_TILE_BYTES + B_HALF_BYTES;  // per CTA
  static constexpr int STAGES_RAW = (SMEM_LIMIT - 1024 - 256 - EPIB) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr int TMEM_COLS = (2 * BN <= 128) ? 128 : (2 * BN <= 256 ? 256 : 512);
  static constexpr int SMEM_BYTES = 1024 + STAGES * STAGE_BYTES + EPIB + 256;
};

// The fused SwiGLU epilogues carry over unchanged (every CTA drains its own 128 rows). SWIGLU: the even CTA loads the
// weight rows of the silu-ed half, the odd CTA the matching rows of the other half — exactly the pair's B split.
template <int BN, bool A_MN, bool B_MN, int OUT, int EPI = EPI_NONE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(256, 1)
gemm2_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                     const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmC2,
                     const float* __restrict__ bias, int M, int N, int K, int split_k) {
  constexpr bool SWIGLU = EPI == EPI_SWIGLU, SWIGLU_BWD = EPI == EPI_SWIGLU_BWD;
  static_assert(!B_MN || (BN / 2) % 64 == 0, "MN-major B: each CTA's half must be whole 64-column blocks");
  static_assert(!SWIGLU || (!A_MN && !B_MN && OUT == OUT_BF16 && BN == 256), "SWIGLU epilogue: K-major bf16 256-wide tiles only");
  static_assert(!SWIGLU_BWD || (!A_MN && B_MN && OUT == OUT_BF16 && BN == 128), "SWIGLU_BWD epilogue: dgrad bf16 128-wide tiles only");
  constexpr int BNT = SWIGLU ? BN / 2 : BN;
  constexpr int EPIB = SWIGLU_BWD ? EPI_BWD_BYTES : EPI_BYTES;
  using C = Cfg2<BN, EPIB>;
  constexpr int NST = C::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = sA + NST * A_TILE_BYTES;
  uint8_t* sE = sB + NST * C::B_HALF_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(sE + EPIB);
  uint64_t* empty = full + NST;
  uint64_t* tfull = empty + NST;
  uint64_t* tempty = tfull + 2;
  uint64_t* lbar = tempty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(lbar + 2 * EPI_WARPS);

  const int lane = threadIdx.x & 31;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int rank = (int)ptx::cluster_ctarank();  // 0 = the even ("leader") CTA, which issues the MMAs
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
  const int mb = (M + 2 * BM - 1) / (2 * BM), nb = (N + BNT - 1) / BNT;
  const int kb_total = (K + BK - 1) / BK;
  const int kb_per = (kb_total + split_k - 1) / split_k;
  const int tiles = mb * nb * split_k;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmB);
    ptx::prefetch_tmap(&tmC);
    if constexpr (SWIGLU || SWIGLU_BWD) ptx::prefetch_tmap(&tmC2);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < NST; ++i) {
      ptx::mbar_init(&full[i], 1);
      ptx::mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&tfull[i], 1);
      ptx::mbar_init(&tempty[i], 2 * EPI_WARPS);
    }
    for (int i = 0; i < 2 * EPI_WARPS; ++i) ptx::mbar_init(&lbar[i], 1);
    ptx::fence_mbar_init();
  }
  if (warp == 2) ptx::tmem_alloc_2sm<C::TMEM_COLS>(tmem_slot);
  ptx::tc_fence_before();
  ptx::cluster_sync_all();  // barriers of both CTAs are initialised before anyone signals them
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs; each loads its A rows and its half of B) =====================
    int stage = 0;
    uint32_t phase = 0;
    for (int t = cluster_id; t < tiles; t += n_clusters) {
      const TileCoord tc = decode_tile(t, mb, nb, kb_total, kb_per);
      const int m0 = (tc.m_blk * 2 + rank) * BM;
      const int n0 = SWIGLU ? (rank ? N : 0) + tc.n_blk * BNT : tc.n_blk * BN + rank * (BN / 2);
      for (int kb = tc.kb0; kb < tc.kb1; ++kb) {
        ptx::mbar_wait(&empty[stage], phase ^ 1);
        if (ptx::elect_one()) {
          if (rank == 0) ptx::mbar_expect_tx(&full[stage], 2 * C::STAGE_BYTES);
          uint8_t* a = sA + stage * A_TILE_BYTES;
          uint8_t* b = sB + stage * C::B_HALF_BYTES;
          if constexpr (!A_MN) {
            ptx::tma_load_2d_2sm(a, &tmA, &full[stage], kb * BK, m0);
          } else {
#pragma unroll
            for (int i = 0; i < BM / 64; ++i) ptx::tma_load_2d_2sm(a + i * MN_BLOCK_BYTES, &tmA, &full[stage], m0 + i * 64, kb * BK);
          }
          if constexpr (!B_MN) {
            ptx::tma_load_2d_2sm(b, &tmB, &full[stage], kb * BK, n0);
          } else {
#pragma unroll
            for (int i = 0; i < BN / 128; ++i) ptx::tma_load_2d_2sm(b + i * MN_BLOCK_BYTES, &tmB, &full[stage], n0 + i * 64, kb * BK);
          }
        }
        __syncwarp();
        if (++stage == NST) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (even CTA only; warp-uniform code, one elected lane issues) =====================
    if (rank == 0) {
      constexpr uint32_t idesc = ptx::make_idesc_bf16(2 * BM, BN, A_MN, B_MN);
      constexpr uint64_t A_KSTEP = (A_MN ? UMMA_K * 128 : UMMA_K * 2) >> 4, B_KSTEP = (B_MN ? UMMA_K * 128 : UMMA_K * 2) >> 4;
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t a_base = ptx::smem_u32(sA), b_base = ptx::smem_u32(sB);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int t = cluster_id; t < tiles; t += n_clusters) {
        const TileCoord tc = decode_tile(t, mb, nb, kb_total, kb_per);
        ptx::mbar_wait(&tempty[as], aphase ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_u + as * BN;
        for (int kb = tc.kb0; kb < tc.kb1; ++kb) {
          ptx::mbar_wait(&full[stage], phase);
          ptx::tc_fence_after();
          const uint32_t a_addr = a_base + stage * A_TILE_BYTES;
          const uint32_t b_addr = b_base + stage * C::B_HALF_BYTES;
          const uint64_t adesc0 = A_MN ? ptx::make_smem_desc_sw128(a_addr, MN_BLOCK_BYTES, 1024) : ptx::make_smem_desc_sw128(a_addr, 16, 1024);
          const uint64_t bdesc0 = B_MN ? ptx::make_smem_desc_sw128(b_addr, MN_BLOCK_BYTES, 1024) : ptx::make_smem_desc_sw128(b_addr, 16, 1024);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k)
            ptx::umma_bf16_elect_2sm(d_tmem, adesc0 + k * A_KSTEP, bdesc0 + k * B_KSTEP, idesc, (kb > tc.kb0 || k > 0) ? 1u : 0u);
          ptx::umma_commit_elect_2sm(&empty[stage]);
          if (++stage == NST) { stage = 0; phase ^= 1; }
        }
        ptx::umma_commit_elect_2sm(&tfull[as]);
        as ^= 1;
        if (as == 0) aphase ^= 1;
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (both CTAs, own 128 rows): TMEM -> registers -> swizzled smem -> TMA store =====================
    const int ew = warp - 4;
    int ebuf = 0;
    int as = 0;
    uint32_t aphase = 0;
    [[maybe_unused]] uint32_t hphase = 0;
    [[maybe_unused]] auto issue_h_loads = [&](int tn, bool wait_stores, int slot_mask) {
      if constexpr (SWIGLU_BWD) {
        const TileCoord tcn = decode_tile(tn, mb, nb, kb_total, kb_per);
        issue_h_tile_loads<BN>(tmC2, sE, lbar, ew, lane, M, N, (tcn.m_blk * 2 + rank) * BM + ew * 32, tcn.n_blk, wait_stores, slot_mask);
      }
    };
    if constexpr (SWIGLU_BWD) {
      if (cluster_id < tiles) issue_h_loads(cluster_id, false, 3);
    }
    for (int t = cluster_id; t < tiles; t += n_clusters) {
      const TileCoord tc = decode_tile(t, mb, nb, kb_total, kb_per);
      ptx::mbar_wait(&tfull[as], aphase);
      ptx::tc_fence_after();
      const int row0 = (tc.m_blk * 2 + rank) * BM + ew * 32;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + as * BN;
      const bool add_bias = (bias != nullptr) && (tc.ks == 0);
      int next_row0 = -1, next_n_blk = 0;
      if constexpr (SWIGLU_BWD) {
        if (t + n_clusters < tiles) {
          const TileCoord tcn = decode_tile(t + n_clusters, mb, nb, kb_total, kb_per);
          next_row0 = (tcn.m_blk * 2 + rank) * BM + ew * 32;
          next_n_blk = tcn.n_blk;
        }
      }
      epilogue_tile<BN, OUT, EPI>(tmC, tmC2, bias, add_bias, sE, lbar, ew, lane, M, N, row0, tc.n_blk, taddr, ebuf, hphase, next_row0, next_n_blk);
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive_cluster(&tempty[as], 0);  // the even CTA's MMA warp owns the accumulator hand-off
      as ^= 1;
      if (as == 0) aphase ^= 1;
      if constexpr (SWIGLU_BWD) {
        if (t + n_clusters < tiles) issue_h_loads(t + n_clusters, true, 2);
      }
    }
    if (lane == 0) ptx::tma_wait_group<0>();
  }

  ptx::tc_fence_before();
  ptx::cluster_sync_all();  // nobody leaves (or frees TMEM) while the peer may still signal or read
  if (warp == 2) ptx::tmem_dealloc_2sm<C::TMEM_COLS>(tmem_base);
}

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  }
  return fn;
}

// 2-D row-major tensor [outer][inner] with `ld` elements between rows. Descriptors are cached per
// (dtype, pointer, extents, pitch, box): the training loop and the sampling loop reuse their buffers (caching allocator),
// so after the first step a GEMM call costs three hash look-ups instead of three driver encodes (~1500 per train step).
struct TmapKey {
  const void* ptr; uint64_t inner, outer, ld; uint32_t box_inner, box_outer; int dt;
  bool operator==(const TmapKey& o) const {
    return ptr == o.ptr && inner == o.inner && outer == o.outer && ld == o.ld && box_inner == o.box_inner && box_outer == o.box_outer && dt == o.dt;
  }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    uint64_t h = reinterpret_cast<uintptr_t>(k.ptr) * 0x9E3779B97F4A7C15ull;
    for (uint64_t v : {k.inner, k.outer, k.ld, (uint64_t)k.box_inner << 32 | k.box_outer, (uint64_t)k.dt}) h = (h ^ v) * 0x100000001B3ull + (h >> 29);
    return (size_t)h;
  }
};
int encode2d(CUtensorMap* m, CUtensorMapDataType dt, int elem_bytes, const void* ptr, uint64_t inner, uint64_t outer,
             uint64_t ld, uint32_t box_inner, uint32_t box_outer) {
  static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> cache;
  static std::mutex mu;
  const TmapKey key{ptr, inner, outer, ld, box_inner, box_outer, (int)dt};
  {
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find(key);
    if (it != cache.end()) { *m = it->second; return DLB_OK; }
  }
  auto enc = get_encode();
  DLB_REQUIRE(enc != nullptr, DLB_ERR_DRIVER, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld * elem_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, dt, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DLB_REQUIRE(r == CUDA_SUCCESS, DLB_ERR_DRIVER,
              "cuTensorMapEncodeTiled failed (%d): inner=%llu outer=%llu ld=%llu box=%ux%u ptr=%p", (int)r,
              (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)ld, box_inner, box_outer, ptr);
  std::lock_guard<std::mutex> lock(mu);
  if (cache.size() > 16384) cache.clear();
  cache.emplace(key, *m);
  return DLB_OK;
}

template <int BN, bool A_MN, bool B_MN, int OUT>
int launch(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const float* bias, int M, int N, int K,
           int split_k, cudaStream_t stream) {
  using C = Cfg<BN>;
  auto kern = gemm_tcgen05_kernel<BN, A_MN, B_MN, OUT>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    DLB_REQUIRE(e == cudaSuccess, (int)e, "cudaFuncSetAttribute(smem=%d): %s", C::SMEM_BYTES, cudaGetErrorString(e));
    attr_set = true;
  }
  const int mb = (M + BM - 1) / BM, nb = (N + BN - 1) / BN;
  const int tiles = mb * nb * split_k;
  const int grid = tiles < dlb_num_sms() ? tiles : dlb_num_sms();
  kern<<<grid, 256, C::SMEM_BYTES, stream>>>(ta, tb, tc, tc, bias, M, N, K, split_k);
  dlb_count_launch();
  return dlb_check_launch("gemm_tcgen05");
}

template <int BN, bool A_MN, bool B_MN>
int dispatch_out(int out_mode, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const float* bias,
                 int M, int N, int K, int split_k, cudaStream_t s) {
  switch (out_mode) {
    case OUT_BF16: return launch<BN, A_MN, B_MN, OUT_BF16>(ta, tb, tc, bias, M, N, K, split_k, s);
    case OUT_F32: return launch<BN, A_MN, B_MN, OUT_F32>(ta, tb, tc, bias, M, N, K, split_k, s);
    default: return launch<BN, A_MN, B_MN, OUT_F32_ADD>(ta, tb, tc, bias, M, N, K, split_k, s);
  }
}
template <int BN>
int dispatch_major(int a_mn, int b_mn, int out_mode, const CUtensorMap& ta, const CUtensorMap& tb,
                   const CUtensorMap& tc, const float* bias, int M, int N, int K, int split_k, cudaStream_t s) {
  if (!a_mn && !b_mn) return dispatch_out<BN, false, false>(out_mode, ta, tb, tc, bias, M, N, K, split_k, s);
  if (!a_mn && b_mn) return dispatch_out<BN, false, true>(out_mode, ta, tb, tc, bias, M, N, K, split_k, s);
  if (a_mn && !b_mn) return dispatch_out<BN, true, false>(out_mode, ta, tb, tc, bias, M, N, K, split_k, s);
  return dispatch_out<BN, true, true>(out_mode, ta, tb, tc, bias, M, N, K, split_k, s);
}

// Cost model fitted to B200 measurements (profiles/gemm_microbench_r1.jsonl): one k-block of a 128 x BN tile costs
// ~(BN + 100) units, the epilogue ~8*BN; a launch costs waves * per-tile cost. Used to pick BN and split-K.
// (The constant was 364 while the MMA warp issued from an `if (lane == 0)` region: ~95 cycles per tcgen05.mma.)
double tile_cost(int M, int N, int kb_total, int bn, int split_k) {
  const int sms = dlb_num_sms();
  const int kb_per = (kb_total + split_k - 1) / split_k;
  const int splits = (kb_total + kb_per - 1) / kb_per;
  const long tiles = (long)((M + BM - 1) / BM) * ((N + bn - 1) / bn) * splits;
  const long waves = (tiles + sms - 1) / sms;
  return (double)waves * ((double)kb_per * (bn + 100.0) + 8.0 * bn);
}

// CTA-pair kernel: 256 x BN tiles over sms/2 clusters; a k-block costs ~(BN + 60) (half of B per CTA -> less L2 traffic)
double pair_cost(int M, int N, int kb_total, int bn, int split_k) {
  const int clusters = dlb_num_sms() / 2;
  const int kb_per = (kb_total + split_k - 1) / split_k;
  const int splits = (kb_total + kb_per - 1) / kb_per;
  const long tiles = (long)((M + 2 * BM - 1) / (2 * BM)) * ((N + bn - 1) / bn) * splits;
  const long waves = (tiles + clusters - 1) / clusters;
  return (double)waves * ((double)kb_per * (bn + 60.0) + 8.0 * bn);
}

// tile_n / split_k: > 0 = forced by the caller, else chosen here. pair: -1 = choose, 0 = single-CTA kernel, 1 = CTA pair.
void pick_config(int M, int N, int K, bool allow_split, bool b_mn, int& tile_n, int& split_k, int& pair) {
  const int kb_total = (K + BK - 1) / BK;
  const int cands[4] = {256, 192, 128, 64};
  const int splits[6] = {1, 2, 4, 8, 16, 32};
  double best = 1e300;
  int best_bn = 128, best_sk = 1, best_pair = 0;
  for (int pr = 0; pr < 2; ++pr) {
    if (pair >= 0 && pr != pair) continue;
    if (pr == 1 && (M <= 128 || N <= 64)) continue;  // a pair needs two row tiles and whole 64-column halves
    for (int i = 0; i < 4; ++i) {
      const int bn = cands[i];
      if (tile_n > 0 && bn != tile_n) continue;
      if (pr == 1 && bn != 256 && bn != 128 && !(bn == 192 && !b_mn)) continue;  // 192: each CTA stages 96 K-major rows of B
      if (tile_n <= 0 && bn == 64 && N > 64) continue;
      if (tile_n <= 0 && bn > 64 && N <= 64) continue;
      for (int j = 0; j < 6; ++j) {
        const int sk = splits[j];
        if (split_k > 0 && sk != split_k) continue;
        if (sk > 1 && (!allow_split || kb_total / sk < 4)) continue;
        const double c = pr ? pair_cost(M, N, kb_total, bn, sk) : tile_cost(M, N, kb_total, bn, sk);
        if (c < best * (1.0 - 1e-9)) { best = c; best_bn = bn; best_sk = sk; best_pair = pr; }
      }
    }
  }
  if (tile_n <= 0) tile_n = best_bn;
  if (split_k <= 0) split_k = best_sk;
  pair = best_pair;
}

int dispatch_pair(int bn, int a_mn, int b_mn, int out_mode, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc,
                  const float* bias, int M, int N, int K, int split_k, cudaStream_t s);
}  // namespace

// See include/diffulab_b200.h for the contract.
DLB_EXPORT int dlb_gemm_bf16(const void* A, const void* B, void* Cout, const float* bias, int64_t M, int64_t N,
                             int64_t K, int64_t lda, int64_t ldb, int64_t ldc, int a_mn_major, int b_mn_major,
                             int out_mode, int split_k, int tile_n, cudaStream_t stream) {
  DLB_REQUIRE(M > 0 && N > 0 && K > 0, DLB_ERR_SHAPE, "gemm: empty problem M=%lld N=%lld K=%lld", (long long)M,
              (long long)N, (long long)K);
  DLB_REQUIRE(M < (1ll << 31) && N < (1ll << 31) && K < (1ll << 31), DLB_ERR_SHAPE, "gemm: dimension overflow");
  DLB_REQUIRE(out_mode >= 0 && out_mode <= 2, DLB_ERR_UNSUPPORTED, "gemm: bad out_mode %d", out_mode);
  DLB_REQUIRE(lda % 8 == 0 && ldb % 8 == 0, DLB_ERR_ALIGN, "gemm: lda/ldb must be multiples of 8 elements (got %lld, %lld)",
              (long long)lda, (long long)ldb);
  const int c_elem = out_mode == OUT_BF16 ? 2 : 4;
  DLB_REQUIRE((ldc * c_elem) % 16 == 0, DLB_ERR_ALIGN, "gemm: ldc*elem must be a multiple of 16 bytes (ldc=%lld)",
              (long long)ldc);
  DLB_REQUIRE(((uintptr_t)A % 16) == 0 && ((uintptr_t)B % 16) == 0 && ((uintptr_t)Cout % 16) == 0, DLB_ERR_ALIGN,
              "gemm: operand pointers must be 16-byte aligned");
  const int kb_total = (int)((K + BK - 1) / BK);
  // split_k: 0 = choose automatically (only ever > 1 in accumulate mode); tile_n: 0 = choose automatically
  if (split_k > 1) DLB_REQUIRE(out_mode == OUT_F32_ADD, DLB_ERR_UNSUPPORTED, "gemm: split_k>1 needs out_mode=2 (fp32 accumulate)");
  DLB_REQUIRE(tile_n == 0 || tile_n == 64 || tile_n == 128 || tile_n == 192 || tile_n == 256, DLB_ERR_UNSUPPORTED,
              "gemm: tile_n %d unsupported", tile_n);
  if (split_k < 0) split_k = 1;
  int bn = tile_n;
  int pair = tile_n == 64 ? 0 : -1;
  pick_config((int)M, (int)N, (int)K, out_mode == OUT_F32_ADD, b_mn_major != 0, bn, split_k, pair);
  if (split_k > kb_total) split_k = kb_total;
  if (split_k > 1) {
    const int kb_per = (kb_total + split_k - 1) / split_k;
    split_k = (kb_total + kb_per - 1) / kb_per;  // no empty splits
  }

  CUtensorMap ta, tb, tc;
  int rc;
  // A: K-major = row-major [M][K] (ld=lda); MN-major = row-major [K][M] (ld=lda).
  if (!a_mn_major) rc = encode2d(&ta, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, A, K, M, lda, BK, BM);
  else rc = encode2d(&ta, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, A, M, K, lda, 64, BK);
  if (rc) return rc;
  if (!b_mn_major) rc = encode2d(&tb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, B, K, N, ldb, BK, pair ? bn / 2 : bn);
  else rc = encode2d(&tb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, B, N, K, ldb, 64, BK);
  if (rc) return rc;
  if (out_mode == OUT_BF16) rc = encode2d(&tc, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, Cout, N, M, ldc, 64, 32);
  else rc = encode2d(&tc, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, Cout, N, M, ldc, 32, 32);
  if (rc) return rc;

  if (pair) return dispatch_pair(bn, a_mn_major, b_mn_major, out_mode, ta, tb, tc, bias, (int)M, (int)N, (int)K, split_k, stream);
  switch (bn) {
    case 64: return dispatch_major<64>(a_mn_major, b_mn_major, out_mode, ta, tb, tc, bias, M, N, K, split_k, stream);
    case 128: return dispatch_major<128>(a_mn_major, b_mn_major, out_mode, ta, tb, tc, bias, M, N, K, split_k, stream);
    case 192: return dispatch_major<192>(a_mn_major, b_mn_major, out_mode, ta, tb, tc, bias, M, N, K, split_k, stream);
    default: return dispatch_major<256>(a_mn_major, b_mn_major, out_mode, ta, tb, tc, bias, M, N, K, split_k, stream);
  }
}


// H[M, 2F] = A[M,K] * W[2F,K]^T + bias (bf16, the saved pre-activation) and ACT[M, F] = silu(H[:, :F]) * H[:, F:] in one
// launch. Replaces Linear + PackedSwiGLU.forward (reference mmdit.py:260-264, nn.py:478-486). F % 128 == 0.
DLB_EXPORT int dlb_gemm_swiglu_bf16(const void* A, const void* W, const float* bias, void* H, void* ACT, int64_t M, int64_t F,
                                    int64_t K, int64_t lda, int64_t ldw, int64_t ldh, int64_t ldact, cudaStream_t stream) {
  DLB_REQUIRE(M > 0 && F > 0 && K > 0 && M < (1ll << 31) && F < (1ll << 30) && K < (1ll << 31), DLB_ERR_SHAPE,
              "gemm_swiglu: bad problem M=%lld F=%lld K=%lld", (long long)M, (long long)F, (long long)K);
  DLB_REQUIRE(F % 128 == 0, DLB_ERR_UNSUPPORTED, "gemm_swiglu: F must be a multiple of 128 (got %lld)", (long long)F);
  DLB_REQUIRE(lda % 8 == 0 && ldw % 8 == 0 && ldh % 8 == 0 && ldact % 8 == 0, DLB_ERR_ALIGN, "gemm_swiglu: strides must be multiples of 8 elements");
  DLB_REQUIRE(((uintptr_t)A % 16) == 0 && ((uintptr_t)W % 16) == 0 && ((uintptr_t)H % 16) == 0 && ((uintptr_t)ACT % 16) == 0, DLB_ERR_ALIGN,
              "gemm_swiglu: operand pointers must be 16-byte aligned");
  CUtensorMap ta, tb, tc, tc2;
  int rc = encode2d(&ta, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, A, K, M, lda, BK, BM);
  if (rc) return rc;
  rc = encode2d(&tb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, W, K, 2 * F, ldw, BK, 128);
  if (rc) return rc;
  rc = encode2d(&tc, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, H, 2 * F, M, ldh, 64, 32);
  if (rc) return rc;
  rc = encode2d(&tc2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, ACT, F, M, ldact, 64, 32);
  if (rc) return rc;
  if (M > BM) {  // CTA-pair kernel: the even CTA stages the silu-ed half of the weight rows, the odd CTA the other half
    using C2 = Cfg2<256>;
    auto kern2 = gemm2_tcgen05_kernel<256, false, false, OUT_BF16, EPI_SWIGLU>;
    static bool attr2_set = false;
    if (!attr2_set) {
      cudaError_t e = cudaFuncSetAttribute(kern2, cudaFuncAttributeMaxDynamicSharedMemorySize, C2::SMEM_BYTES);
      DLB_REQUIRE(e == cudaSuccess, (int)e, "cudaFuncSetAttribute(smem=%d): %s", C2::SMEM_BYTES, cudaGetErrorString(e));
      attr2_set = true;
    }
    const int tiles2 = (int)(((M + 2 * BM - 1) / (2 * BM)) * (F / 128));
    const int maxc = dlb_num_sms() / 2;
    kern2<<<2 * (tiles2 < maxc ? tiles2 : maxc), 256, C2::SMEM_BYTES, stream>>>(ta, tb, tc, tc2, bias, (int)M, (int)F, (int)K, 1);
    dlb_count_launch();
    return dlb_check_launch("gemm2_swiglu");
  }
  using C = Cfg<256>;
  auto kern = gemm_tcgen05_kernel<256, false, false, OUT_BF16, EPI_SWIGLU>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    DLB_REQUIRE(e == cudaSuccess, (int)e, "cudaFuncSetAttribute(smem=%d): %s", C::SMEM_BYTES, cudaGetErrorString(e));
    attr_set = true;
  }
  const int tiles = (int)(((M + BM - 1) / BM) * (F / 128));
  const int grid = tiles < dlb_num_sms() ? tiles : dlb_num_sms();
  kern<<<grid, 256, C::SMEM_BYTES, stream>>>(ta, tb, tc, tc2, bias, (int)M, (int)F, (int)K, 1);
  dlb_count_launch();
  return dlb_check_launch("gemm_swiglu");
}


// dH[M, 2F] = SwiGLU'(H) applied to d(act) = dY[M,D] * W2[D,F]: the dgrad of the MLP down-projection with the SwiGLU
// backward fused into its epilogue (d(act) never reaches memory). Replaces the autograd mirror of
// PackedSwiGLU.forward + nn.Linear (reference nn.py:478-486, mmdit.py:260-264). W2 is the row-major [D, F] weight of the
// down projection (nn.Linear(F, D).weight); F % 128 == 0.
DLB_EXPORT int dlb_gemm_swiglu_bwd_bf16(const void* dY, const void* W2, const void* H, void* dH, int64_t M, int64_t F, int64_t D,
                                        int64_t lddy, int64_t ldw2, int64_t ldh, int64_t lddh, cudaStream_t stream) {
  DLB_REQUIRE(M > 0 && F > 0 && D > 0 && M < (1ll << 31) && F < (1ll << 30) && D < (1ll << 31), DLB_ERR_SHAPE,
              "gemm_swiglu_bwd: bad problem M=%lld F=%lld D=%lld", (long long)M, (long long)F, (long long)D);
  DLB_REQUIRE(F % 128 == 0, DLB_ERR_UNSUPPORTED, "gemm_swiglu_bwd: F must be a multiple of 128 (got %lld)", (long long)F);
  DLB_REQUIRE(lddy % 8 == 0 && ldw2 % 8 == 0 && ldh % 8 == 0 && lddh % 8 == 0, DLB_ERR_ALIGN, "gemm_swiglu_bwd: strides must be multiples of 8 elements");
  DLB_REQUIRE(((uintptr_t)dY % 16) == 0 && ((uintptr_t)W2 % 16) == 0 && ((uintptr_t)H % 16) == 0 && ((uintptr_t)dH % 16) == 0, DLB_ERR_ALIGN,
              "gemm_swiglu_bwd: operand pointers must be 16-byte aligned");
  CUtensorMap ta, tb, tc, tc2;
  int rc = encode2d(&ta, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dY, D, M, lddy, BK, BM);   // A = dY, K-major (K = D)
  if (rc) return rc;
  rc = encode2d(&tb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, W2, F, D, ldw2, 64, BK);        // B = W2^T read MN-major
  if (rc) return rc;
  rc = encode2d(&tc, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dH, 2 * F, M, lddh, 64, 32);
  if (rc) return rc;
  rc = encode2d(&tc2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, H, 2 * F, M, ldh, 64, 32);
  if (rc) return rc;
  // The CTA-pair instantiation exists (and is parity-tested through DLB_SWIGLU_BWD_PAIR=1) but measured slower in the
  // train step (12.4 vs 10.7 ms / step): this epilogue is the bottleneck and a pair does not shorten it.
  static const bool use_pair = getenv("DLB_SWIGLU_BWD_PAIR") != nullptr;
  if (use_pair && M > BM) {
    using C2 = Cfg2<128, EPI_BWD_BYTES>;
    auto kern2 = gemm2_tcgen05_kernel<128, false, true, OUT_BF16, EPI_SWIGLU_BWD>;
    static bool attr2_set = false;
    if (!attr2_set) {
      cudaError_t e = cudaFuncSetAttribute(kern2, cudaFuncAttributeMaxDynamicSharedMemorySize, C2::SMEM_BYTES);
      DLB_REQUIRE(e == cudaSuccess, (int)e, "cudaFuncSetAttribute(smem=%d): %s", C2::SMEM_BYTES, cudaGetErrorString(e));
      attr2_set = true;
    }
    const int tiles2 = (int)(((M + 2 * BM - 1) / (2 * BM)) * (F / 128));
    const int maxc = dlb_num_sms() / 2;
    kern2<<<2 * (tiles2 < maxc ? tiles2 : maxc), 256, C2::SMEM_BYTES, stream>>>(ta, tb, tc, tc2, nullptr, (int)M, (int)F, (int)D, 1);
    dlb_count_launch();
    return dlb_check_launch("gemm2_swiglu_bwd");
  }
  using C = Cfg<128, EPI_BWD_BYTES>;
  auto kern = gemm_tcgen05_kernel<128, false, true, OUT_BF16, EPI_SWIGLU_BWD>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    DLB_REQUIRE(e == cudaSuccess, (int)e, "cudaFuncSetAttribute(smem=%d): %s", C::SMEM_BYTES, cudaGetErrorString(e));
    attr_set = true;
  }
  const int tiles = (int)(((M + BM - 1) / BM) * (F / 128));
  const int grid = tiles < dlb_num_sms() ? tiles : dlb_num_sms();
  kern<<<grid, 384, C::SMEM_BYTES, stream>>>(ta, tb, tc, tc2, nullptr, (int)M, (int)F, (int)D, 1);
  dlb_count_launch();
  return dlb_check_launch("gemm_swiglu_bwd");
}


namespace {
template <int BN, bool A_MN, bool B_MN, int OUT>
int launch2(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const float* bias, int M, int N, int K, int split_k,
            cudaStream_t stream) {
  using C = Cfg2<BN>;
  auto kern = gemm2_tcgen05_kernel<BN, A_MN, B_MN, OUT>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    DLB_REQUIRE(e == cudaSuccess, (int)e, "cudaFuncSetAttribute(smem=%d): %s", C::SMEM_BYTES, cudaGetErrorString(e));
    attr_set = true;
  }
  const int mb = (M + 2 * BM - 1) / (2 * BM), nb = (N + BN - 1) / BN;
  const int tiles = mb * nb * split_k;
  const int max_clusters = dlb_num_sms() / 2;
  const int clusters = tiles < max_clusters ? tiles : max_clusters;
  kern<<<2 * clusters, 256, C::SMEM_BYTES, stream>>>(ta, tb, tc, tc, bias, M, N, K, split_k);
  dlb_count_launch();
  return dlb_check_launch("gemm2_tcgen05");
}
template <int BN, bool A_MN, bool B_MN>
int dispatch_out2(int out_mode, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const float* bias, int M, int N,
                  int K, int split_k, cudaStream_t s) {
  switch (out_mode) {
    case OUT_BF16: return launch2<BN, A_MN, B_MN, OUT_BF16>(ta, tb, tc, bias, M, N, K, split_k, s);
    case OUT_F32: return launch2<BN, A_MN, B_MN, OUT_F32>(ta, tb, tc, bias, M, N, K, split_k, s);
    default: return launch2<BN, A_MN, B_MN, OUT_F32_ADD>(ta, tb, tc, bias, M, N, K, split_k, s);
  }
}
template <int BN>
int dispatch_major2(int a_mn, int b_mn, int out_mode, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc,
                    const float* bias, int M, int N, int K, int split_k, cudaStream_t s) {
  if (!a_mn && !b_mn) return dispatch_out2<BN, false, false>(out_mode, ta, tb, tc, bias, M, N, K, split_k, s);
  if (!a_mn && b_mn) return dispatch_out2<BN, false, true>(out_mode, ta, tb, tc, bias, M, N, K, split_k, s);
  if (a_mn && !b_mn) return dispatch_out2<BN, true, false>(out_mode, ta, tb, tc, bias, M, N, K, split_k, s);
  return dispatch_out2<BN, true, true>(out_mode, ta, tb, tc, bias, M, N, K, split_k, s);
}
int dispatch_pair(int bn, int a_mn, int b_mn, int out_mode, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc,
                  const float* bias, int M, int N, int K, int split_k, cudaStream_t s) {
  if (bn == 128) return dispatch_major2<128>(a_mn, b_mn, out_mode, ta, tb, tc, bias, M, N, K, split_k, s);
  if (bn == 192) {  // K-major B only (checked by the caller): 96-row halves
    if (!a_mn) return dispatch_out2<192, false, false>(out_mode, ta, tb, tc, bias, M, N, K, split_k, s);
    return dispatch_out2<192, true, false>(out_mode, ta, tb, tc, bias, M, N, K, split_k, s);
  }
  return dispatch_major2<256>(a_mn, b_mn, out_mode, ta, tb, tc, bias, M, N, K, split_k, s);
}
}  // namespace

// CTA-pair (cta_group::2, 256 x tile_n tiles) version of dlb_gemm_bf16: same contract; tile_n in {128, 256}; split_k >= 1.
DLB_EXPORT int dlb_gemm2_bf16(const void* A, const void* B, void* Cout, const float* bias, int64_t M, int64_t N, int64_t K,
                              int64_t lda, int64_t ldb, int64_t ldc, int a_mn_major, int b_mn_major, int out_mode, int split_k,
                              int tile_n, cudaStream_t stream) {
  DLB_REQUIRE(M > 0 && N > 0 && K > 0 && M < (1ll << 31) && N < (1ll << 31) && K < (1ll << 31), DLB_ERR_SHAPE, "gemm2: bad problem");
  DLB_REQUIRE(out_mode >= 0 && out_mode <= 2, DLB_ERR_UNSUPPORTED, "gemm2: bad out_mode %d", out_mode);
  DLB_REQUIRE(tile_n == 128 || tile_n == 256 || (tile_n == 192 && !b_mn_major), DLB_ERR_UNSUPPORTED,
              "gemm2: tile_n must be 128 or 256 (or 192 with a K-major B operand), got %d", tile_n);
  DLB_REQUIRE(lda % 8 == 0 && ldb % 8 == 0, DLB_ERR_ALIGN, "gemm2: lda/ldb must be multiples of 8 elements");
  const int c_elem = out_mode == OUT_BF16 ? 2 : 4;
  DLB_REQUIRE((ldc * c_elem) % 16 == 0, DLB_ERR_ALIGN, "gemm2: ldc*elem must be a multiple of 16 bytes");
  DLB_REQUIRE(((uintptr_t)A % 16) == 0 && ((uintptr_t)B % 16) == 0 && ((uintptr_t)Cout % 16) == 0, DLB_ERR_ALIGN, "gemm2: pointers must be 16-byte aligned");
  if (split_k < 1) split_k = 1;
  if (split_k > 1) DLB_REQUIRE(out_mode == OUT_F32_ADD, DLB_ERR_UNSUPPORTED, "gemm2: split_k>1 needs out_mode=2");
  const int kb_total = (int)((K + BK - 1) / BK);
  if (split_k > kb_total) split_k = kb_total;
  if (split_k > 1) {
    const int kb_per = (kb_total + split_k - 1) / split_k;
    split_k = (kb_total + kb_per - 1) / kb_per;
  }
  const int bn = tile_n;
  CUtensorMap ta, tb, tc;
  int rc;
  if (!a_mn_major) rc = encode2d(&ta, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, A, K, M, lda, BK, BM);
  else rc = encode2d(&ta, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, A, M, K, lda, 64, BK);
  if (rc) return rc;
  if (!b_mn_major) rc = encode2d(&tb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, B, K, N, ldb, BK, bn / 2);
  else rc = encode2d(&tb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, B, N, K, ldb, 64, BK);
  if (rc) return rc;
  if (out_mode == OUT_BF16) rc = encode2d(&tc, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, Cout, N, M, ldc, 64, 32);
  else rc = encode2d(&tc, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, Cout, N, M, ldc, 32, 32);
  if (rc) return rc;
  return dispatch_pair(bn, a_mn_major, b_mn_major, out_mode, ta, tb, tc, bias, (int)M, (int)N, (int)K, split_k, stream);
}
