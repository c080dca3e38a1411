// tcgen05 joint attention, forward and backward, sm_100a.
//
// Replaces F.scaled_dot_product_attention as called by DiTAttention / MMDiTAttention (reference mmdit.py:92-98,
// 184-204): softmax(q k^T * hd^-1/2 + key_padding_mask) v over the concatenation of up to two segments (text rows
// first, then image rows), bf16 operands, fp32 softmax. Q / K arrive RMS-normalised and rotated (qknorm_rope.cu), V is
// read in place from the packed qkv projection.
//
// Structure (all three kernels): one CTA owns one 128-row tile of one (sample, head) — the "resident" tile, TMEM lane =
// resident row — and streams the other sequence axis past it in 64-row tiles through a ring of shared-memory stages.
// 288 threads: warps 0-7 are compute warps (two threads per resident row, each owning 32 of the 64 columns of a score
// tile; they also stage all operands with cp.async, 8 rows x 64 bytes per warp instruction), warp 8 only issues
// tcgen05.mma (warp-uniform code, elect.sync) so that instruction issue never sits on the softmax threads' path.
// No CTA-wide barrier inside the loop; the hand-offs are mbarriers:
//   full[s]  (256 arrivals) tile in stage s has landed            compute -> MMA warp
//   bar1[b]  (tcgen05.commit) score tiles in TMEM buffer b ready   MMA warp -> compute
//   ps_full  (256 arrivals) P / dS operand tile written, TMEM buffer drained   compute -> MMA warp
//   bar2     (tcgen05.commit) accumulating products of tile j done: stage, operand tile (and O tile) free
// The MMA warp issues the score products of tile j+1 before the accumulating products of tile j, so the tensor core
// computes scores while the compute warps do the exponentials of the previous tile; loads run NST-1 tiles ahead.
// Every operand tile uses ONE shared-memory layout ("L1(R)": 16-byte chunk (row r, chunk c) of an R-row tile at
// c*R*16 + r*16), a valid non-swizzled UMMA layout both K-major (LBO R*16, SBO 128) and MN-major (LBO 128, SBO R*16)
// — pinned by tests/test_umma_probe_gpu.py — so Q / dO / K / V serve as row operands of one product and as transposed
// operands of another without data movement. P / dS tiles are written by their owning threads in the same layout.
//   forward : S = Q K_j^T -> online softmax (row max exchanged between the two column halves) -> P (smem) -> O_j = P V_j,
//             read from TMEM one iteration later and accumulated (rescaled) in registers
//   dq      : S = Q K_j^T, dP = dO V_j^T, dS = P o (dP - D), dQ += dS K_j      (also produces D = rowsum(dO o O))
//   dkv     : S^T = K Q_j^T, dP^T = V dO_j^T, dV += P^T dO_j, dK += dS^T Q_j
// Head dims that are not a multiple of 16 (DiT-XL/2: 72) are zero-padded in shared memory only.
#include "common.cuh"
#include "ptx.cuh"

namespace attn_tc {
typedef __nv_bfloat16 bf16;

constexpr int NC = 256;      // compute threads
constexpr int NT = NC + 32;  // + the MMA-issue warp
constexpr int KT = 64;       // rows of a streamed tile
constexpr float LOG2E = 1.4426950408889634f;

struct Seg {
  const bf16* q; const bf16* k; const bf16* v; const bf16* o; const bf16* dout;
  bf16* out; bf16* dq; bf16* dk; bf16* dv;
  int64_t ldq, ldk, ldv, ldo, lddo, lddq, lddk, lddv;
  int len;
};
struct Params {
  Seg seg[2];
  float* lse;            // [B,H,S] natural-log LSE of the scaled scores (written by fwd, read by bwd)
  float* dsum;           // [B,H,S] rowsum(dO o O): written by the dq kernel, read by the dkv kernel
  const uint8_t* kmask;  // [B, mask_len] 1 = attend; keys >= mask_len always attend
  int mask_len, B, H, S, hd;
  float scale, scale_log2;
  long long* trace;      // development aid (dlb_attn_set_trace): SM-clock timeline of one CTA per kernel, else null
};

// timeline slot i of kernel KIND (0 fwd, 1 dq, 2 dkv): written by thread 0 of the CTA (0, 0, B/2)
#define ATTN_TRACE(KIND, i)                                                                                   \
  do {                                                                                                        \
    if (p.trace != nullptr && threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == p.B / 2) \
      p.trace[(KIND) * 64 + (i)] = clock64();                                                                 \
  } while (0)

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// 16-byte async copy; !valid writes zeros (src-size 0), src must still be a mapped address
__device__ __forceinline__ void cp_async16_zfill(void* dst, const void* src, bool valid) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(src), "r"(sz));
}
__device__ __forceinline__ void cp_async4(void* dst, const void* src) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(d), "l"(src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }
__device__ __forceinline__ void pair_barrier(int id) { asm volatile("bar.sync %0, 64;\n" ::"r"(id) : "memory"); }
__device__ __forceinline__ void compute_barrier() { asm volatile("bar.sync 5, 256;\n" ::: "memory"); }

enum { T_Q = 0, T_K = 1, T_V = 2, T_DO = 3, T_O = 4 };
enum { O_OUT = 0, O_DQ = 1, O_DK = 2, O_DV = 3 };

__device__ __forceinline__ int64_t seg_row(const Params& p, int b, int s, int& sg) {
  sg = s < p.seg[0].len ? 0 : 1;
  return (int64_t)b * p.seg[sg].len + (sg ? s - p.seg[0].len : s);
}
__device__ __forceinline__ const bf16* row_ptr(const Params& p, int which, int b, int h, int s) {
  int sg;
  const int64_t row = seg_row(p, b, s, sg);
  const Seg& g = p.seg[sg];
  const bf16* base = which == T_Q ? g.q + row * g.ldq
                   : which == T_K ? g.k + row * g.ldk
                   : which == T_V ? g.v + row * g.ldv
                   : which == T_DO ? g.dout + row * g.lddo : g.o + row * g.ldo;
  return base + (int64_t)h * p.hd;
}

// The 8 compute warps stage a ROWS-row tile into layout L1(ROWS). One warp instruction covers 8 rows x 4 chunks: 64
// contiguous bytes of each row on the global side, and 4 conflict-free 128-byte wavefronts on the shared side.
template <int HDP, int WHICH, int ROWS>
__device__ __forceinline__ void load_tile(uint8_t* sm, const Params& p, int b, int h, int s0, int warp, int lane) {
  constexpr int CPR = HDP / 8;
  const int nvalid = p.hd >> 3;
  const int cl = lane & 3;
#pragma unroll
  for (int rg = 0; rg < ROWS / 64; ++rg) {
    const int r = (rg * 8 + warp) * 8 + (lane >> 2);
    const int s = s0 + r;
    const bool rv = s < p.S;
    const bf16* src = row_ptr(p, WHICH, b, h, rv ? s : 0);
    uint8_t* dst = sm + r * 16;
#pragma unroll
    for (int c0 = 0; c0 < CPR; c0 += 4) {
      const int c = c0 + cl;
      if (c0 + 3 < CPR || c < CPR) cp_async16_zfill(dst + c * (ROWS * 16), src + (c < nvalid ? c * 8 : 0), rv && c < nvalid);
    }
  }
}

// row-major bf16 staging tile [128][HDP] -> global (valid rows / columns only), coalesced 16-byte stores
template <int HDP, int WHICH>
__device__ __forceinline__ void store_tile(const bf16* stage, const Params& p, int b, int h, int s0, int tid) {
  const int cpr = p.hd >> 3;
  for (int idx = tid; idx < 128 * cpr; idx += NC) {
    const int r = idx / cpr, c = idx - r * cpr;
    const int s = s0 + r;
    if (s < p.S) {
      int sg;
      const int64_t row = seg_row(p, b, s, sg);
      const Seg& g = p.seg[sg];
      bf16* base = WHICH == O_OUT ? g.out + row * g.ldo
                 : WHICH == O_DQ ? g.dq + row * g.lddq
                 : WHICH == O_DK ? g.dk + row * g.lddk : g.dv + row * g.lddv;
      *reinterpret_cast<uint4*>(base + (int64_t)h * p.hd + c * 8) = *reinterpret_cast<const uint4*>(stage + r * HDP + c * 8);
    }
  }
}

// this thread's half of a finished fp32 TMEM tile (HDP columns) -> scaled bf16 in the row-major staging tile
template <int HDP>
__device__ __forceinline__ void tmem_half_to_stage(uint32_t taddr, bf16* stage, int r, int half, float mul) {
#pragma unroll
  for (int c8 = 0; c8 < HDP / 16; ++c8) {
    const int c = half * (HDP / 2) + c8 * 8;
    uint32_t v[8];
    ptx::tmem_ld8(taddr + c, v);
    ptx::tmem_ld_wait();
    float t[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) t[j] = __uint_as_float(v[j]) * mul;
    st8(stage + r * HDP + c, pack8(t));
  }
}

__device__ __forceinline__ float key_bias(const Params& p, int b, int key) {
  if (key >= p.S) return -INFINITY;
  if (p.kmask && key < p.mask_len && p.kmask[(int64_t)b * p.mask_len + key] == 0) return -INFINITY;
  return 0.f;
}
// does the 64-key tile starting at kv0 contain any key that may carry a bias (uniform over the CTA)
__device__ __forceinline__ bool tile_may_be_masked(const Params& p, int kv0) {
  return kv0 + KT > p.S || (p.kmask != nullptr && kv0 < p.mask_len);
}

// 32 columns starting at column c of row r -> 4 chunks of an L1(128) tile (rowbase = tile + r*16)
__device__ __forceinline__ void store_bf16x32(uint8_t* rowbase, int c, const float* v) {
#pragma unroll
  for (int q4 = 0; q4 < 4; ++q4) {
    uint4 u;
    u.x = pack_bf16x2(v[8 * q4 + 0], v[8 * q4 + 1]);
    u.y = pack_bf16x2(v[8 * q4 + 2], v[8 * q4 + 3]);
    u.z = pack_bf16x2(v[8 * q4 + 4], v[8 * q4 + 5]);
    u.w = pack_bf16x2(v[8 * q4 + 6], v[8 * q4 + 7]);
    *reinterpret_cast<uint4*>(rowbase + ((c >> 3) + q4) * 2048) = u;
  }
}

// Descriptors of the two tile shapes (resident 128-row tiles, streamed 64-row tiles). d0 = descriptor of the tile base;
// a k-step (16 elements of the contraction) advances the 16-byte-unit address field.
__device__ __forceinline__ uint64_t desc_k128(uint32_t a) { return ptx::make_smem_desc_noswz(a, 2048, 128); }  // K-major
__device__ __forceinline__ uint64_t desc_k64(uint32_t a) { return ptx::make_smem_desc_noswz(a, 1024, 128); }
__device__ __forceinline__ uint64_t desc_mn64(uint32_t a) { return ptx::make_smem_desc_noswz(a, 128, 1024); }  // MN-major (N = head dim)
constexpr uint64_t KSTEP_K128 = 4096 >> 4, KSTEP_K64 = 2048 >> 4, KSTEP_MN64 = 256 >> 4;

// ---------------------------------------------------------------------------------------------------------
// forward: 256 threads, 128-key tiles, one TMEM score buffer; two CTAs per SM overlap each other's phases
// ---------------------------------------------------------------------------------------------------------
template <int HDP>
__global__ void __launch_bounds__(NC, 2) attn_fwd_tc_kernel(const Params p) {
  constexpr int TILE = 128 * HDP * 2;
  constexpr int HH = HDP / 2;  // O columns per thread
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + TILE;
  uint8_t* sV = sK + TILE;
  uint8_t* sP = sV + TILE;                                 // [128 q][128 keys] bf16, layout L1(128), 32 KB (reused as O staging)
  float* sBias = reinterpret_cast<float*>(sP + 32768);     // 128 additive key biases (0 / -inf)
  float* sX = sBias + 128;                                 // [2][128] exchange between the two column halves
  __shared__ uint64_t bar_s, bar_o;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, r = tid & 127, half = tid >> 7;
  const int warp_u = __shfl_sync(0xffffffffu, warp, 0);  // warp-uniform for the compiler: the MMA issue code below stays in uniform registers
  const int q0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
  ATTN_TRACE(0, 0);
  if (tid == 0) { ptx::mbar_init(&bar_s, 1); ptx::mbar_init(&bar_o, 1); ptx::fence_mbar_init(); }
  if (warp == 0) ptx::tmem_alloc<256>(&tmem_slot);
  load_tile<HDP, T_Q, 128>(sQ, p, b, h, q0, warp, lane);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, tmem_slot, 0);
  const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
  const uint32_t tS = tmem + lane_off, tO = tmem + 128 + lane_off;
  constexpr uint32_t idesc_s = ptx::make_idesc_bf16(128, 128, false, false);
  constexpr uint32_t idesc_o = ptx::make_idesc_bf16(128, HDP, false, true);
  const uint64_t dq0 = desc_k128(ptx::smem_u32(sQ)), dk0 = desc_k128(ptx::smem_u32(sK)), dp0 = desc_k128(ptx::smem_u32(sP));
  const uint64_t dv0 = ptx::make_smem_desc_noswz(ptx::smem_u32(sV), 128, 2048);  // V read MN-major (N = head dim)
  float o[HH];
#pragma unroll
  for (int i = 0; i < HH; ++i) o[i] = 0.f;
  float m = -INFINITY, l = 0.f;
  uint32_t phase = 0;
  const int cbase = half * 64;  // my S columns
  int it = 0;

  for (int kv0 = 0; kv0 < p.S; kv0 += 128, ++it) {
    const bool masked_tile = kv0 + 128 > p.S || (p.kmask != nullptr && kv0 < p.mask_len);  // uniform over the CTA
    load_tile<HDP, T_K, 128>(sK, p, b, h, kv0, warp, lane);
    load_tile<HDP, T_V, 128>(sV, p, b, h, kv0, warp, lane);
    if (masked_tile && half == 0) sBias[r] = key_bias(p, b, kv0 + r);
    if (it < 4) ATTN_TRACE(0, 4 + it * 8);
    cp_async_commit();
    cp_async_wait<0>();
    ptx::fence_proxy_async_smem();
    ptx::tc_fence_before();
    __syncthreads();
    if (it < 4) ATTN_TRACE(0, 5 + it * 8);
    if (warp_u == 0) {
      ptx::tc_fence_after();
#pragma unroll
      for (int ks = 0; ks < HDP / 16; ++ks) ptx::umma_bf16_elect(tmem, dq0 + ks * KSTEP_K128, dk0 + ks * KSTEP_K128, idesc_s, ks > 0);
      ptx::umma_commit_elect(&bar_s);
    }
    ptx::mbar_wait(&bar_s, phase);
    ptx::tc_fence_after();
    if (it < 4) ATTN_TRACE(0, 6 + it * 8);
    // pass 1: maximum over my 64 columns, then exchange with the other half of the row
    float mx = -INFINITY;
#pragma unroll
    for (int c = 0; c < 64; c += 32) {
      uint32_t v[32];
      ptx::tmem_ld32(tS + cbase + c, v);
      ptx::tmem_ld_wait();
      if (masked_tile) {
#pragma unroll
        for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(v[j]) + sBias[cbase + c + j]);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(v[j]));
      }
    }
    sX[half * 128 + r] = mx;
    pair_barrier(1 + (warp & 3));  // the two warps that share these 32 rows
    mx = fmaxf(sX[r], sX[128 + r]);
    const float m_new = fmaxf(m, mx);
    const float ms = (m_new == -INFINITY) ? 0.f : m_new * p.scale_log2;
    const float alpha = ex2(m * p.scale_log2 - ms);
    // pass 2: probabilities of my 64 columns -> bf16 A operand
    float lsum = 0.f;
    uint8_t* prow = sP + r * 16;
#pragma unroll
    for (int c = 0; c < 64; c += 32) {
      uint32_t v[32];
      ptx::tmem_ld32(tS + cbase + c, v);
      ptx::tmem_ld_wait();
      float pv[32];
      if (masked_tile) {
#pragma unroll
        for (int j = 0; j < 32; ++j) pv[j] = ex2((__uint_as_float(v[j]) + sBias[cbase + c + j]) * p.scale_log2 - ms);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) pv[j] = ex2(__uint_as_float(v[j]) * p.scale_log2 - ms);
      }
#pragma unroll
      for (int j = 0; j < 32; j += 4) lsum += (pv[j] + pv[j + 1]) + (pv[j + 2] + pv[j + 3]);
      store_bf16x32(prow, cbase + c, pv);
    }
    l = l * alpha + lsum;
    m = m_new;
    if (it < 4) ATTN_TRACE(0, 7 + it * 8);
    ptx::fence_proxy_async_smem();
    ptx::tc_fence_before();
    __syncthreads();
    if (it < 4) ATTN_TRACE(0, 8 + it * 8);
    if (warp_u == 0) {
      ptx::tc_fence_after();
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) ptx::umma_bf16_elect(tmem + 128, dp0 + ks * KSTEP_K128, dv0 + ks * (256 >> 4), idesc_o, ks > 0);
      ptx::umma_commit_elect(&bar_o);
    }
    ptx::mbar_wait(&bar_o, phase);
    ptx::tc_fence_after();
    if (it < 4) ATTN_TRACE(0, 9 + it * 8);
#pragma unroll
    for (int c8 = 0; c8 < HH / 8; ++c8) {  // o = o * alpha + O_tile (my half of the head dim)
      uint32_t v[8];
      ptx::tmem_ld8(tO + half * HH + c8 * 8, v);
      ptx::tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 8; ++j) o[c8 * 8 + j] = o[c8 * 8 + j] * alpha + __uint_as_float(v[j]);
    }
    phase ^= 1;
    ptx::tc_fence_before();
    __syncthreads();  // S / O tiles and sK / sV / sP are free again
    if (it < 4) ATTN_TRACE(0, 10 + it * 8);
  }
  // finalise: total row sum from both halves, normalise, stage, store
  sX[half * 128 + r] = l;
  pair_barrier(1 + (warp & 3));
  l = sX[r] + sX[128 + r];
  const float inv = l > 0.f ? 1.f / l : 0.f;
  const int row = q0 + r;
  if (half == 0 && p.lse && row < p.S) p.lse[((int64_t)b * p.H + h) * p.S + row] = m * p.scale + logf(l);
  bf16* stage = reinterpret_cast<bf16*>(sP);
#pragma unroll
  for (int c = 0; c < HH; c += 8) {
    float t8[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) t8[j] = o[c + j] * inv;
    st8(stage + r * HDP + half * HH + c, pack8(t8));
  }
  __syncthreads();
  ATTN_TRACE(0, 58);
  store_tile<HDP, O_OUT>(stage, p, b, h, q0, tid);
  ATTN_TRACE(0, 59);
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc<256>(tmem);
}

// ---------------------------------------------------------------------------------------------------------
// backward: dQ (and D = rowsum(dO o O))
// ---------------------------------------------------------------------------------------------------------
template <int HDP>
__global__ void __launch_bounds__(NT, 1) attn_bwd_dq_tc_kernel(const Params p) {
  constexpr int NST = 4;
  constexpr int TQ = 128 * HDP * 2, TK = KT * HDP * 2;
  constexpr int CPR = HDP / 8;
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* sQ = smem;
  uint8_t* sDO = sQ + TQ;
  uint8_t* sKV = sDO + TQ;                                 // NST stages of {K tile, V tile}
  uint8_t* sDS = sKV + NST * 2 * TK;                       // [128 q][64 keys] bf16, layout L1(128), 16 KB
  float* sBias = reinterpret_cast<float*>(sDS + 16384);    // [NST][64] key biases
  float* sDrow = sBias + NST * KT;                         // [128] D
  __shared__ uint64_t full[NST], bar1[2], ps_full, bar2;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
  const int T = (p.S + KT - 1) / KT;
  ATTN_TRACE(1, 0);
  if (tid == 0) {
    for (int i = 0; i < NST; ++i) ptx::mbar_init(&full[i], NC);
    ptx::mbar_init(&bar1[0], 1); ptx::mbar_init(&bar1[1], 1); ptx::mbar_init(&ps_full, NC); ptx::mbar_init(&bar2, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 8) ptx::tmem_alloc<512>(&tmem_slot);
  constexpr uint32_t idesc_s = ptx::make_idesc_bf16(128, KT, false, false);
  constexpr uint32_t idesc_dq = ptx::make_idesc_bf16(128, HDP, false, true);

  if (warp == 8) {
    // ---------------- MMA-issue warp ----------------
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = __shfl_sync(0xffffffffu, tmem_slot, 0);
    const uint64_t dq0 = desc_k128(ptx::smem_u32(sQ)), dd0 = desc_k128(ptx::smem_u32(sDO)), ds0 = desc_k128(ptx::smem_u32(sDS));
    const uint32_t kva = ptx::smem_u32(sKV);
    for (int j = -1; j < T; ++j) {
      if (j + 1 < T) {  // S(t) = Q K_t^T, dP(t) = dO V_t^T into TMEM buffer t & 1
        const int t = j + 1;
        ptx::mbar_wait(&full[t % NST], (t / NST) & 1);
        ptx::tc_fence_after();
        const uint64_t dk = desc_k64(kva + (t % NST) * 2 * TK), dv = desc_k64(kva + (t % NST) * 2 * TK + TK);
        const uint32_t ts = tmem + (t & 1) * 128;
#pragma unroll
        for (int ks = 0; ks < HDP / 16; ++ks) {
          ptx::umma_bf16_elect(ts, dq0 + ks * KSTEP_K128, dk + ks * KSTEP_K64, idesc_s, ks > 0);
          ptx::umma_bf16_elect(ts + KT, dd0 + ks * KSTEP_K128, dv + ks * KSTEP_K64, idesc_s, ks > 0);
        }
        ptx::umma_commit_elect(&bar1[t & 1]);
      }
      if (j >= 0) {  // dQ += dS K_j : A K-major over keys, B = K tile read MN-major (N = hd, K = keys)
        ptx::mbar_wait(&ps_full, j & 1);
        ptx::tc_fence_after();
        const uint64_t dk = desc_mn64(kva + (j % NST) * 2 * TK);
#pragma unroll
        for (int ks = 0; ks < KT / 16; ++ks) ptx::umma_bf16_elect(tmem + 256, ds0 + ks * KSTEP_K128, dk + ks * KSTEP_MN64, idesc_dq, (j > 0 || ks > 0) ? 1u : 0u);
        ptx::umma_commit_elect(&bar2);
      }
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tmem_dealloc<512>(tmem);
    return;
  }

  // ---------------- compute warps ----------------
  const int r = tid & 127, half = tid >> 7;
  auto issue_tile = [&](int t) {
    if (t < T) {
      uint8_t* st = sKV + (t % NST) * 2 * TK;
      load_tile<HDP, T_K, KT>(st, p, b, h, t * KT, warp, lane);
      load_tile<HDP, T_V, KT>(st + TK, p, b, h, t * KT, warp, lane);
      if (tid < KT && tile_may_be_masked(p, t * KT)) sBias[(t % NST) * KT + tid] = key_bias(p, b, t * KT + tid);
    }
    cp_async_commit();
  };
  // dO tile: staged through registers (same 8 rows x 4 chunks mapping) so that D = rowsum(dO o O) comes for free.
  // All global loads are issued first, ahead of the cp.async traffic they would otherwise queue behind.
  constexpr int NCG = (CPR + 3) / 4;
  bf16x8 dreg[2][NCG], oreg[2][NCG];
  {
    const int nvalid = p.hd >> 3;
    const int cl = lane & 3;
#pragma unroll
    for (int rg = 0; rg < 2; ++rg) {
      const int rr = (rg * 8 + warp) * 8 + (lane >> 2);
      const int s = q0 + rr;
      const bool rv = s < p.S;
      const bf16* dsrc = row_ptr(p, T_DO, b, h, rv ? s : 0);
      const bf16* osrc = row_ptr(p, T_O, b, h, rv ? s : 0);
#pragma unroll
      for (int g = 0; g < NCG; ++g) {
        const int c = g * 4 + cl;
        dreg[rg][g].u[0] = dreg[rg][g].u[1] = dreg[rg][g].u[2] = dreg[rg][g].u[3] = 0u;
        oreg[rg][g] = dreg[rg][g];
        if (rv && c < nvalid) { dreg[rg][g] = ld8(dsrc + c * 8); oreg[rg][g] = ld8(osrc + c * 8); }
      }
    }
  }
  load_tile<HDP, T_Q, 128>(sQ, p, b, h, q0, warp, lane);
#pragma unroll
  for (int t = 0; t < NST; ++t) issue_tile(t);
  ATTN_TRACE(1, 1);
  {
    const int cl = lane & 3;
#pragma unroll
    for (int rg = 0; rg < 2; ++rg) {
      const int rr = (rg * 8 + warp) * 8 + (lane >> 2);
      const int s = q0 + rr;
      uint8_t* dst = sDO + rr * 16;
      float dpart = 0.f;
#pragma unroll
      for (int g = 0; g < NCG; ++g) {
        const int c = g * 4 + cl;
        if (g * 4 + 3 < CPR || c < CPR) {
          float df[8], of[8];
          unpack8(dreg[rg][g], df); unpack8(oreg[rg][g], of);
#pragma unroll
          for (int i = 0; i < 8; ++i) dpart += df[i] * of[i];
          *reinterpret_cast<bf16x8*>(dst + c * 2048) = dreg[rg][g];
        }
      }
      dpart += __shfl_xor_sync(0xffffffffu, dpart, 1);
      dpart += __shfl_xor_sync(0xffffffffu, dpart, 2);
      if (cl == 0) {
        sDrow[rr] = dpart;
        if (s < p.S) p.dsum[((int64_t)b * p.H + h) * p.S + s] = dpart;
      }
    }
  }
  const int myrow = q0 + r;
  const float Lrow = myrow < p.S ? p.lse[((int64_t)b * p.H + h) * p.S + myrow] * LOG2E : INFINITY;
  ATTN_TRACE(1, 2);
  cp_async_wait<NST - 1>();  // Q and tile 0
  ptx::fence_proxy_async_smem();
  ptx::mbar_arrive(&full[0]);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  ATTN_TRACE(1, 3);
  const float Drow = sDrow[r];
  const uint32_t tmem = tmem_slot;
  const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
  const uint32_t tDQ = tmem + 256 + lane_off;

  for (int j = 0; j < T; ++j) {
    if (j + 1 < T) {
      cp_async_wait<NST - 3>();  // my part of tile j+1 has landed
      ptx::fence_proxy_async_smem();
      ptx::mbar_arrive(&full[(j + 1) % NST]);
    }
    if (j < 8) ATTN_TRACE(1, 4 + j * 6);
    ptx::mbar_wait(&bar1[j & 1], (j >> 1) & 1);
    ptx::tc_fence_after();
    if (j < 8) ATTN_TRACE(1, 5 + j * 6);
    if (j >= 1) {
      ptx::mbar_wait(&bar2, (j - 1) & 1);  // dQ += dS K_{j-1} finished: its stage and sDS are free
      issue_tile(j + NST - 1);
    }
    if (j < 8) ATTN_TRACE(1, 6 + j * 6);
    {
      const uint32_t ts = tmem + lane_off + (j & 1) * 128 + half * 32;
      uint32_t vs[32], vd[32];
      ptx::tmem_ld32(ts, vs);
      ptx::tmem_ld32(ts + KT, vd);
      ptx::tmem_ld_wait();
      float ds[32];
      if (tile_may_be_masked(p, j * KT)) {
        const float* bias = sBias + (j % NST) * KT + half * 32;
#pragma unroll
        for (int i = 0; i < 32; ++i) ds[i] = ex2((__uint_as_float(vs[i]) + bias[i]) * p.scale_log2 - Lrow) * (__uint_as_float(vd[i]) - Drow);
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) ds[i] = ex2(__uint_as_float(vs[i]) * p.scale_log2 - Lrow) * (__uint_as_float(vd[i]) - Drow);
      }
      store_bf16x32(sDS + r * 16, half * 32, ds);
    }
    ptx::fence_proxy_async_smem();
    ptx::tc_fence_before();
    ptx::mbar_arrive(&ps_full);
    if (j < 8) ATTN_TRACE(1, 7 + j * 6);
  }
  ATTN_TRACE(1, 56);
  ptx::mbar_wait(&bar2, (T - 1) & 1);
  ptx::tc_fence_after();
  ATTN_TRACE(1, 57);
  bf16* stage = reinterpret_cast<bf16*>(sKV);
  tmem_half_to_stage<HDP>(tDQ, stage, r, half, p.scale);
  compute_barrier();
  store_tile<HDP, O_DQ>(stage, p, b, h, q0, tid);
  ATTN_TRACE(1, 59);
  ptx::tc_fence_before();
  __syncthreads();
}

// ---------------------------------------------------------------------------------------------------------
// backward: dK, dV
// ---------------------------------------------------------------------------------------------------------
template <int HDP>
__global__ void __launch_bounds__(NT, 1) attn_bwd_dkv_tc_kernel(const Params p) {
  constexpr int NST = 4;
  constexpr int TQ = 128 * HDP * 2, TK = KT * HDP * 2;
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* sK = smem;
  uint8_t* sV = sK + TQ;
  uint8_t* sQD = sV + TQ;                                  // NST stages of {Q tile, dO tile}
  uint8_t* sPT = sQD + NST * 2 * TK;                       // P^T  [128 keys][64 queries] bf16, layout L1(128)
  uint8_t* sDST = sPT + 16384;                             // dS^T
  float* sL = reinterpret_cast<float*>(sDST + 16384);      // [NST][64] lse (natural log; +inf for rows >= S)
  float* sD = sL + NST * KT;                               // [NST][64] D
  __shared__ uint64_t full[NST], bar1[2], ps_full, bar2;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int kv0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
  const int T = (p.S + KT - 1) / KT;
  ATTN_TRACE(2, 0);
  if (tid == 0) {
    for (int i = 0; i < NST; ++i) ptx::mbar_init(&full[i], NC);
    ptx::mbar_init(&bar1[0], 1); ptx::mbar_init(&bar1[1], 1); ptx::mbar_init(&ps_full, NC); ptx::mbar_init(&bar2, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 8) ptx::tmem_alloc<512>(&tmem_slot);
  constexpr uint32_t idesc_s = ptx::make_idesc_bf16(128, KT, false, false);
  constexpr uint32_t idesc_o = ptx::make_idesc_bf16(128, HDP, false, true);

  if (warp == 8) {
    // ---------------- MMA-issue warp ----------------
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = __shfl_sync(0xffffffffu, tmem_slot, 0);
    const uint64_t dk0 = desc_k128(ptx::smem_u32(sK)), dv0 = desc_k128(ptx::smem_u32(sV));
    const uint64_t dp0 = desc_k128(ptx::smem_u32(sPT)), ds0 = desc_k128(ptx::smem_u32(sDST));
    const uint32_t qda = ptx::smem_u32(sQD);
    for (int j = -1; j < T; ++j) {
      if (j + 1 < T) {  // S^T(t) = K Q_t^T, dP^T(t) = V dO_t^T into TMEM buffer t & 1
        const int t = j + 1;
        ptx::mbar_wait(&full[t % NST], (t / NST) & 1);
        ptx::tc_fence_after();
        const uint64_t dq = desc_k64(qda + (t % NST) * 2 * TK), dd = desc_k64(qda + (t % NST) * 2 * TK + TK);
        const uint32_t ts = tmem + (t & 1) * 128;
#pragma unroll
        for (int ks = 0; ks < HDP / 16; ++ks) {
          ptx::umma_bf16_elect(ts, dk0 + ks * KSTEP_K128, dq + ks * KSTEP_K64, idesc_s, ks > 0);
          ptx::umma_bf16_elect(ts + KT, dv0 + ks * KSTEP_K128, dd + ks * KSTEP_K64, idesc_s, ks > 0);
        }
        ptx::umma_commit_elect(&bar1[t & 1]);
      }
      if (j >= 0) {  // dV += P^T dO_j, dK += dS^T Q_j: contraction over the 64 queries; Q / dO tiles read MN-major (N = hd)
        ptx::mbar_wait(&ps_full, j & 1);
        ptx::tc_fence_after();
        const uint64_t dq = desc_mn64(qda + (j % NST) * 2 * TK), dd = desc_mn64(qda + (j % NST) * 2 * TK + TK);
#pragma unroll
        for (int ks = 0; ks < KT / 16; ++ks) {
          const uint32_t acc = (j > 0 || ks > 0) ? 1u : 0u;
          ptx::umma_bf16_elect(tmem + 256 + HDP, dp0 + ks * KSTEP_K128, dd + ks * KSTEP_MN64, idesc_o, acc);
          ptx::umma_bf16_elect(tmem + 256, ds0 + ks * KSTEP_K128, dq + ks * KSTEP_MN64, idesc_o, acc);
        }
        ptx::umma_commit_elect(&bar2);
      }
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tmem_dealloc<512>(tmem);
    return;
  }

  // ---------------- compute warps ----------------
  const int r = tid & 127, half = tid >> 7;
  const float* lse = p.lse + ((int64_t)b * p.H + h) * p.S;
  const float* dsm = p.dsum + ((int64_t)b * p.H + h) * p.S;
  auto issue_tile = [&](int t) {
    if (t < T) {
      uint8_t* st = sQD + (t % NST) * 2 * TK;
      load_tile<HDP, T_Q, KT>(st, p, b, h, t * KT, warp, lane);
      load_tile<HDP, T_DO, KT>(st + TK, p, b, h, t * KT, warp, lane);
      if (tid < 2 * KT) {
        const int i = tid & (KT - 1), s = t * KT + i;
        float* dst = (tid < KT ? sL : sD) + (t % NST) * KT + i;
        if (s < p.S) cp_async4(dst, (tid < KT ? lse : dsm) + s);
        else *dst = tid < KT ? INFINITY : 0.f;
      }
    }
    cp_async_commit();
  };
  load_tile<HDP, T_K, 128>(sK, p, b, h, kv0, warp, lane);
  load_tile<HDP, T_V, 128>(sV, p, b, h, kv0, warp, lane);
#pragma unroll
  for (int t = 0; t < NST; ++t) issue_tile(t);
  ATTN_TRACE(2, 1);
  const float kbias = key_bias(p, b, kv0 + r);
  ATTN_TRACE(2, 2);
  cp_async_wait<NST - 1>();  // K, V and tile 0
  ptx::fence_proxy_async_smem();
  ptx::mbar_arrive(&full[0]);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  ATTN_TRACE(2, 3);
  const uint32_t tmem = tmem_slot;
  const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
  const uint32_t tDK = tmem + 256 + lane_off, tDV = tmem + 256 + HDP + lane_off;

  for (int j = 0; j < T; ++j) {
    if (j + 1 < T) {
      cp_async_wait<NST - 3>();  // my part of tile j+1 has landed
      ptx::fence_proxy_async_smem();
      ptx::mbar_arrive(&full[(j + 1) % NST]);
    }
    if (j < 8) ATTN_TRACE(2, 4 + j * 6);
    ptx::mbar_wait(&bar1[j & 1], (j >> 1) & 1);
    ptx::tc_fence_after();
    if (j < 8) ATTN_TRACE(2, 5 + j * 6);
    if (j >= 1) {
      ptx::mbar_wait(&bar2, (j - 1) & 1);  // dV / dK accumulation of tile j-1 finished: its stage, sPT and sDST are free
      issue_tile(j + NST - 1);
    }
    if (j < 8) ATTN_TRACE(2, 6 + j * 6);
    {
      const float* Lq = sL + (j % NST) * KT + half * 32;
      const float* Dq = sD + (j % NST) * KT + half * 32;
      const uint32_t ts = tmem + lane_off + (j & 1) * 128 + half * 32;
      uint32_t vs[32], vd[32];
      ptx::tmem_ld32(ts, vs);
      ptx::tmem_ld32(ts + KT, vd);
      ptx::tmem_ld_wait();
      float pv[32], ds[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        pv[i] = ex2(fmaf(Lq[i], -LOG2E, (__uint_as_float(vs[i]) + kbias) * p.scale_log2));
        ds[i] = pv[i] * (__uint_as_float(vd[i]) - Dq[i]);
      }
      store_bf16x32(sPT + r * 16, half * 32, pv);
      store_bf16x32(sDST + r * 16, half * 32, ds);
    }
    ptx::fence_proxy_async_smem();
    ptx::tc_fence_before();
    ptx::mbar_arrive(&ps_full);
    if (j < 8) ATTN_TRACE(2, 7 + j * 6);
  }
  ATTN_TRACE(2, 56);
  ptx::mbar_wait(&bar2, (T - 1) & 1);
  ptx::tc_fence_after();
  ATTN_TRACE(2, 57);
  bf16* stage = reinterpret_cast<bf16*>(sQD);
  tmem_half_to_stage<HDP>(tDK, stage, r, half, p.scale);
  compute_barrier();
  store_tile<HDP, O_DK>(stage, p, b, h, kv0, tid);
  compute_barrier();
  tmem_half_to_stage<HDP>(tDV, stage, r, half, 1.f);
  compute_barrier();
  store_tile<HDP, O_DV>(stage, p, b, h, kv0, tid);
  ATTN_TRACE(2, 59);
  ptx::tc_fence_before();
  __syncthreads();
}

}  // namespace attn_tc

// plain-C segment description shared with attention.cu (include/diffulab_b200.h: dlb_attn_seg)
struct dlb_attn_seg {
  const void* q; const void* k; const void* v;
  void* o;
  const void* dout;
  void* dq; void* dk; void* dv;
  int64_t ldq, ldk, ldv, ldo, lddo, lddq, lddk, lddv;
  int32_t len;
};

static long long* g_attn_trace = nullptr;
// Development aid: device buffer of 3 x 64 int64 that receives one CTA's SM-clock timeline per kernel (null = off).
DLB_EXPORT int dlb_attn_set_trace(long long* dev_buf) {
  g_attn_trace = dev_buf;
  return DLB_OK;
}

static int fill_tc_params(attn_tc::Params& p, const char* who, const dlb_attn_seg* segs, int nseg, float* lse, float* dsum,
                          const uint8_t* kmask, int mask_len, int B, int H, int hd, float scale, bool bwd) {
  using namespace attn_tc;
  DLB_REQUIRE(nseg == 1 || nseg == 2, DLB_ERR_SHAPE, "%s: 1 or 2 segments supported (got %d)", who, nseg);
  DLB_REQUIRE(B > 0 && H > 0 && hd > 0 && hd % 8 == 0 && hd <= 128, DLB_ERR_SHAPE, "%s: B=%d H=%d hd=%d", who, B, H, hd);
  int S = 0;
  for (int i = 0; i < nseg; ++i) {
    const dlb_attn_seg& s = segs[i];
    DLB_REQUIRE(s.len >= 0 && s.ldq % 8 == 0 && s.ldk % 8 == 0 && s.ldv % 8 == 0 && s.ldo % 8 == 0, DLB_ERR_ALIGN,
                "%s: strides must be multiples of 8", who);
    if (bwd)
      DLB_REQUIRE(s.lddo % 8 == 0 && s.lddq % 8 == 0 && s.lddk % 8 == 0 && s.lddv % 8 == 0, DLB_ERR_ALIGN, "%s: strides must be multiples of 8", who);
    p.seg[i] = Seg{(const bf16*)s.q, (const bf16*)s.k, (const bf16*)s.v, (const bf16*)s.o, (const bf16*)s.dout, (bf16*)s.o,
                   (bf16*)s.dq, (bf16*)s.dk, (bf16*)s.dv, s.ldq, s.ldk, s.ldv, s.ldo, s.lddo, s.lddq, s.lddk, s.lddv, s.len};
    S += s.len;
  }
  DLB_REQUIRE(S > 0 && mask_len >= 0 && mask_len <= S && (kmask != nullptr || mask_len == 0), DLB_ERR_SHAPE, "%s: bad sequence / mask", who);
  p.lse = lse; p.dsum = dsum; p.kmask = kmask; p.mask_len = mask_len; p.B = B; p.H = H; p.S = S; p.hd = hd;
  p.scale = scale; p.scale_log2 = scale * 1.4426950408889634f;
  p.trace = g_attn_trace;
  return DLB_OK;
}

#define HDP_SWITCH_TC(hd, ...)                                          \
  switch (((hd) + 15) / 16 * 16) {                                      \
    case 16: case 32: case 48: case 64: { constexpr int HDPV = 64; __VA_ARGS__; break; } \
    case 80: { constexpr int HDPV = 80; __VA_ARGS__; break; }           \
    case 96: { constexpr int HDPV = 96; __VA_ARGS__; break; }           \
    default: { constexpr int HDPV = 128; __VA_ARGS__; break; }          \
  }

// Same contract as dlb_attn_fwd (attention.cu); tcgen05 implementation.
DLB_EXPORT int dlb_attn_fwd_tc(const dlb_attn_seg* segs, int nseg, float* lse, const uint8_t* kmask, int mask_len, int B,
                               int H, int hd, float scale, cudaStream_t stream) {
  using namespace attn_tc;
  Params p{};
  int rc = fill_tc_params(p, "attn_fwd_tc", segs, nseg, lse, nullptr, kmask, mask_len, B, H, hd, scale, false);
  if (rc) return rc;
  dim3 grid((p.S + 127) / 128, H, B);
  HDP_SWITCH_TC(hd, {
    const size_t smem = (size_t)3 * 128 * HDPV * 2 + 32768 + 3 * 128 * 4;
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_tc_kernel<HDPV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    DLB_REQUIRE(e == cudaSuccess, (int)e, "attn_fwd_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attn_fwd_tc_kernel<HDPV><<<grid, NC, smem, stream>>>(p);
  });
  dlb_count_launch();
  return dlb_check_launch("attn_fwd_tc");
}

// Same contract as dlb_attn_bwd (attention.cu); tcgen05 implementation. dsum is written by the dq pass and read by
// the dkv pass (stream order).
DLB_EXPORT int dlb_attn_bwd_tc(const dlb_attn_seg* segs, int nseg, const float* lse, float* dsum, const uint8_t* kmask,
                               int mask_len, int B, int H, int hd, float scale, cudaStream_t stream) {
  using namespace attn_tc;
  DLB_REQUIRE(lse != nullptr && dsum != nullptr, DLB_ERR_SHAPE, "attn_bwd_tc: lse and dsum buffers are required");
  Params p{};
  int rc = fill_tc_params(p, "attn_bwd_tc", segs, nseg, const_cast<float*>(lse), dsum, kmask, mask_len, B, H, hd, scale, true);
  if (rc) return rc;
  dim3 grid((p.S + 127) / 128, H, B);
  HDP_SWITCH_TC(hd, {
    const size_t sm_dq = (size_t)2 * 128 * HDPV * 2 + 4 * 2 * 64 * HDPV * 2 + 16384 + 4 * 64 * 4 + 128 * 4;
    const size_t sm_dkv = (size_t)2 * 128 * HDPV * 2 + 4 * 2 * 64 * HDPV * 2 + 32768 + 2 * 4 * 64 * 4;
    cudaError_t e = cudaFuncSetAttribute(attn_bwd_dq_tc_kernel<HDPV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_dq);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_bwd_dkv_tc_kernel<HDPV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_dkv);
    DLB_REQUIRE(e == cudaSuccess, (int)e, "attn_bwd_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attn_bwd_dq_tc_kernel<HDPV><<<grid, NT, sm_dq, stream>>>(p);
    attn_bwd_dkv_tc_kernel<HDPV><<<grid, NT, sm_dkv, stream>>>(p);
  });
  dlb_count_launch(2);
  return dlb_check_launch("attn_bwd_tc");
}
