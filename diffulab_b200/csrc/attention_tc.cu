// tcgen05 joint attention, forward and backward, sm_100a.
//
// Replaces F.scaled_dot_product_attention as called by DiTAttention / MMDiTAttention (reference mmdit.py:92-98,
// 184-204): softmax(q k^T * hd^-1/2 + key_padding_mask) v over the concatenation of up to two segments (text rows
// first, then image rows), bf16 operands, fp32 softmax. Q / K arrive RMS-normalised and rotated (qknorm_rope.cu), V is
// read in place from the packed qkv projection.
//
// Structure (forward attn_fwd_ws_tc_kernel, backward attn_bwd_dq/dkv_tc_kernel; attn_fwd_tc_kernel is the earlier 256-thread
// forward kept for comparison): one CTA owns one 128-row tile of one (sample, head) — the "resident" tile, TMEM lane =
// resident row — and streams the other sequence axis past it in 64-row tiles through a ring of shared-memory stages.
// 352 threads: warps 0-7 are compute warps (two threads per resident row, each owning 32 of the 64 columns of a score
// tile), warp 8 only issues tcgen05.mma (warp-uniform code, elect.sync) so that instruction issue never sits on the
// softmax threads' path, warps 9-10 are producers: one elected lane issues TMA loads through 4-D tensor maps whose
// boxes land directly in the operand layout (segment lengths that are multiples of 128), otherwise both warps stage with
// cp.async (8 rows x 64 bytes per warp instruction) and arrive asynchronously (cp.async.mbarrier.arrive.noinc).
// CTAs are persistent over (tile, head, sample) work items and the load ring runs ahead across item boundaries.
// No CTA-wide barrier inside the loop; the hand-offs are mbarriers:
//   full[s]  (TMA bytes / producer arrivals) tile in stage s has landed       producers -> MMA warp
//   bar1[b]  (tcgen05.commit) score tiles in TMEM buffer b ready              MMA warp -> compute
//   ps_full  (256 arrivals) P / dS operand tile written, TMEM buffer drained  compute -> MMA warp
//   bar2     (tcgen05.commit) accumulating products of tile j done: operand tile (and O tile) free
//   empty[s] (tcgen05.commit) stage s may be refilled                          MMA warp -> producers
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

constexpr int NC = 256;      // compute threads (forward: the whole CTA)
constexpr int NPW = 2;       // backward: producer (cp.async) warps
constexpr int NT = NC + 32 + NPW * 32;  // backward CTA: compute warps 0-7, MMA-issue warp 8, producer warps 9..
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
__device__ __forceinline__ void cp_async4_zfill(void* dst, const void* src, bool valid) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
  const int sz = valid ? 4 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(d), "l"(src), "r"(sz));
}
// this thread arrives on the mbarrier (without raising its pending count) once all its earlier cp.async have landed
__device__ __forceinline__ void cp_async_arrive_noinc(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(ptx::smem_u32(bar)) : "memory");
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

// NW loader warps stage a ROWS-row tile into layout L1(ROWS). One warp instruction covers 8 rows x 4 chunks: 64
// contiguous bytes of each row on the global side, and 4 conflict-free 128-byte wavefronts on the shared side.
template <int HDP, int WHICH, int ROWS, int NW = 8>
__device__ __forceinline__ void load_tile(uint8_t* sm, const Params& p, int b, int h, int s0, int warp, int lane) {
  constexpr int CPR = HDP / 8;
  const int nvalid = p.hd >> 3;
  const int cl = lane & 3;
#pragma unroll
  for (int rg = 0; rg < ROWS / (8 * NW); ++rg) {
    const int r = (rg * NW + warp) * 8 + (lane >> 2);
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
  const int dq = NC / cpr, dm = NC - dq * cpr;  // idx += NC  <=>  (r, c) += (dq, dm) with carry
  int r = tid / cpr, c = tid - r * cpr;
  while (r < 128) {
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
    r += dq;
    c += dm;
    if (c >= cpr) { c -= cpr; ++r; }
  }
}

// this thread's half of a finished fp32 TMEM tile (HDP columns) -> scaled bf16 in the row-major staging tile
template <int HDP>
__device__ __forceinline__ void tmem_half_to_stage(uint32_t taddr, bf16* stage, int r, int half, float mul) {
  constexpr int HH = HDP / 2;  // 32, 40, 48 or 64 columns: all loads are issued before the single wait
  uint32_t v[HH];
  const uint32_t a = taddr + half * HH;
  ptx::tmem_ld32(a, v);
  if constexpr (HH == 40) ptx::tmem_ld8(a + 32, v + 32);
  if constexpr (HH == 48) ptx::tmem_ld16(a + 32, v + 32);
  if constexpr (HH == 64) ptx::tmem_ld32(a + 32, v + 32);
  ptx::tmem_ld_wait();
#pragma unroll
  for (int c = 0; c < HH; c += 8) {
    float t[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) t[j] = __uint_as_float(v[c + j]) * mul;
    st8(stage + r * HDP + half * HH + c, pack8(t));
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
// backward kernels: persistent CTAs (one per SM) walk a list of work items (128-row tile x, head h, sample b).
// Warp roles: 0-7 compute, 8 MMA issue, 9.. producers. The producers run the cp.async ring NST-1 streamed tiles ahead
// ACROSS item boundaries (resident tiles are double-buffered, RES = 2) and are the only threads that ever block on
// load-store-unit backpressure; each producer thread arrives on full[stage] asynchronously when its copies have landed
// (cp.async.mbarrier.arrive.noinc). The MMA warp likewise issues the next item's first score products during the
// current item's epilogue. RES = 1 (large head dims, or fewer than NST streamed tiles per item) = one item per CTA.
//   empty[s] (tcgen05.commit) accumulating products that read stage s are done        MMA warp -> producers
// ---------------------------------------------------------------------------------------------------------
// TMA fast path (every segment length a multiple of 128): 4-D tensor maps {8 elements, rows, 16-byte chunks of a head,
// heads} whose boxes land directly in layout L1(64) / L1(128) (pinned by scripts/probe_tma_gather.py); one elected
// producer lane issues them. Otherwise the producer warps use cp.async (ragged tiles, arbitrary segment lengths).
enum { M_Q64 = 0, M_Q128, M_K64, M_K128, M_V64, M_V128, M_DO64, M_DO128, M_O128, M_COUNT };
struct BwdMaps { CUtensorMap m[2][M_COUNT]; };
__device__ __forceinline__ void tma_tile(void* dst, const BwdMaps& maps, int which, const Params& p, int b, int h, int s0, uint64_t* bar) {
  const int sg = s0 < p.seg[0].len ? 0 : 1;
  const int row = b * p.seg[sg].len + (sg ? s0 - p.seg[0].len : s0);
  ptx::tma_load_4d(dst, &maps.m[sg][which], bar, 0, row, 0, h);
}

struct Item { int x, h, b; };
__device__ __forceinline__ Item decode_item(const Params& p, int w, int nx) {
  Item it;
  it.x = w % nx;
  const int t = w / nx;
  it.h = t % p.H;
  it.b = t / p.H;
  return it;
}
#define MMA_TRACE(KIND, item, i)                                                                   \
  do {                                                                                             \
    if (p.trace != nullptr && lane == 0 && blockIdx.x == gridDim.x / 2 && (item) == (n_my > 1 ? 1 : 0) && (i) < 56) p.trace[(KIND) * 64 + (i)] = clock64(); \
  } while (0)
#define BWD_TRACE(KIND, i)                                                                         \
  do {                                                                                             \
    if (p.trace != nullptr && tid == 0 && blockIdx.x == gridDim.x / 2 && k == trace_item) p.trace[(KIND) * 64 + (i)] = clock64(); \
  } while (0)

// ---------------------------------------------------------------------------------------------------------
// forward, persistent warp-specialised version (same roles and hand-offs as the backward kernels): resident Q tile
// (double-buffered), 64-key K / V tiles in a ring fed by the producer warps (TMA or cp.async), S = Q K_j^T double-
// buffered in TMEM so the MMA warp computes the scores of tile j+1 while the compute warps do the softmax of tile j;
// O_j = P V_j is read from TMEM one iteration later and accumulated (rescaled) in registers.
// ---------------------------------------------------------------------------------------------------------
template <int HDP, int RES, int NST, bool TMA>
__global__ void __launch_bounds__(NT, 1) attn_fwd_ws_tc_kernel(const Params p, const __grid_constant__ BwdMaps maps) {
  constexpr int TQ = 128 * HDP * 2, TK = KT * HDP * 2;
  constexpr int HH = HDP / 2;
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* sRes = smem;                                     // RES x Q tile, layout L1(128)
  uint8_t* sKV = sRes + RES * TQ;                           // NST stages of {K tile, V tile}
  uint8_t* sP = sKV + NST * 2 * TK;                         // [128 q][64 keys] bf16, layout L1(128), 16 KB
  bf16* sStage = reinterpret_cast<bf16*>(sP + 16384);       // [128][HDP] output staging
  float* sX = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(sStage) + TQ);  // [2][2][128] pair exchange
  __shared__ uint64_t full[NST], empty[NST], bar_s[2], ps_full, bar_o;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
  const int T = (p.S + KT - 1) / KT, nx = (p.S + 127) / 128;
  const int nitems = nx * p.H * p.B;
  const int n_my = (nitems - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int G = n_my * T;
  if (tid == 0) {
    for (int i = 0; i < NST; ++i) { ptx::mbar_init(&full[i], TMA ? 1 : NPW * 32); ptx::mbar_init(&empty[i], 1); }
    ptx::mbar_init(&bar_s[0], 1); ptx::mbar_init(&bar_s[1], 1); ptx::mbar_init(&ps_full, NC); ptx::mbar_init(&bar_o, 1);
    ptx::fence_mbar_init();
  }
  if (warp_u == 8) ptx::tmem_alloc<256>(&tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  constexpr uint32_t idesc_s = ptx::make_idesc_bf16(128, KT, false, false);
  constexpr uint32_t idesc_o = ptx::make_idesc_bf16(128, HDP, false, true);

  if (warp_u > 8) {
    // ---------------- producer warps ----------------
    const int pw = warp_u - 9;
    int uk = 0, uj = 0;
    for (int u = 0; u < G; ++u) {
      if (u >= NST) ptx::mbar_wait(&empty[u % NST], ((u / NST) - 1) & 1);
      const Item it = decode_item(p, blockIdx.x + uk * gridDim.x, nx);
      uint8_t* st = sKV + (u % NST) * 2 * TK;
      uint8_t* res = sRes + (uk % RES) * TQ;
      if constexpr (TMA) {
        if (pw == 0 && ptx::elect_one()) {
          uint64_t* bar = &full[u % NST];
          ptx::mbar_expect_tx(bar, 2 * TK + (uj == 0 ? TQ : 0));
          if (uj == 0) tma_tile(res, maps, M_Q128, p, it.b, it.h, it.x * 128, bar);
          tma_tile(st, maps, M_K64, p, it.b, it.h, uj * KT, bar);
          tma_tile(st + TK, maps, M_V64, p, it.b, it.h, uj * KT, bar);
        }
        __syncwarp();
        if (++uj == T) { uj = 0; ++uk; }
        continue;
      }
      if (uj == 0) load_tile<HDP, T_Q, 128, NPW>(res, p, it.b, it.h, it.x * 128, pw, lane);
      load_tile<HDP, T_K, KT, NPW>(st, p, it.b, it.h, uj * KT, pw, lane);
      load_tile<HDP, T_V, KT, NPW>(st + TK, p, it.b, it.h, uj * KT, pw, lane);
      cp_async_arrive_noinc(&full[u % NST]);
      if (++uj == T) { uj = 0; ++uk; }
    }
    cp_async_wait<0>();
  } else if (warp_u == 8) {
    // ---------------- MMA-issue warp ----------------
    const uint32_t tmem = __shfl_sync(0xffffffffu, tmem_slot, 0);
    const uint32_t resa = ptx::smem_u32(sRes), kva = ptx::smem_u32(sKV);
    const uint64_t dp0 = desc_k128(ptx::smem_u32(sP));
    int tk = 0, tj = 0;
    for (int g = -1; g < G; ++g) {
      const bool next_ready = g + 1 < G && (g < 0 || __shfl_sync(0xffffffffu, (int)ptx::mbar_test_wait(&full[(g + 1) % NST], ((g + 1) / NST) & 1), 0) != 0);
#pragma unroll 1
      for (int pass = 0; pass < 2; ++pass) {
        if ((pass == 0) == next_ready && g + 1 < G) {  // S(t) = Q K_t^T into TMEM buffer t & 1
          const int t = g + 1;
          ptx::mbar_wait(&full[t % NST], (t / NST) & 1);
          ptx::fence_proxy_async_smem();
          ptx::tc_fence_after();
          const uint64_t dq0 = desc_k128(resa + (tk % RES) * TQ);
          const uint64_t dk = desc_k64(kva + (t % NST) * 2 * TK);
#pragma unroll
          for (int ks = 0; ks < HDP / 16; ++ks) ptx::umma_bf16_elect(tmem + (t & 1) * KT, dq0 + ks * KSTEP_K128, dk + ks * KSTEP_K64, idesc_s, ks > 0);
          ptx::umma_commit_elect(&bar_s[t & 1]);
          if (++tj == T) { tj = 0; ++tk; }
        }
        if (pass == 0 && g >= 0) {  // O_g = P V_g (fresh tile: the compute warps accumulate in registers)
          ptx::mbar_wait(&ps_full, g & 1);
          ptx::tc_fence_after();
          const uint64_t dv = desc_mn64(kva + (g % NST) * 2 * TK + TK);
#pragma unroll
          for (int ks = 0; ks < KT / 16; ++ks) ptx::umma_bf16_elect(tmem + 128, dp0 + ks * KSTEP_K128, dv + ks * KSTEP_MN64, idesc_o, ks > 0);
          ptx::umma_commit_elect(&empty[g % NST]);
          ptx::umma_commit_elect(&bar_o);
        }
      }
    }
  } else {
    // ---------------- compute warps ----------------
    const int r = tid & 127, half = tid >> 7;
    const uint32_t tmem = tmem_slot;
    const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t tO = tmem + 128 + lane_off;
    int k = 0, j = 0;
    Item it = decode_item(p, blockIdx.x, nx);
    float o[HH];
    float m = -INFINITY, l = 0.f, alpha_prev = 0.f;
    auto add_o_tile = [&]() {  // o = o * alpha + O_tile (my half of the head dim)
      uint32_t v[HH];
      const uint32_t a = tO + half * HH;
      ptx::tmem_ld32(a, v);
      if constexpr (HH == 40) ptx::tmem_ld8(a + 32, v + 32);
      if constexpr (HH == 48) ptx::tmem_ld16(a + 32, v + 32);
      if constexpr (HH == 64) ptx::tmem_ld32(a + 32, v + 32);
      ptx::tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < HH; ++i) o[i] = fmaf(o[i], alpha_prev, __uint_as_float(v[i]));
    };
    for (int g = 0; g < G; ++g) {
      if (j == 0) {
        it = decode_item(p, blockIdx.x + k * gridDim.x, nx);
#pragma unroll
        for (int i = 0; i < HH; ++i) o[i] = 0.f;
        m = -INFINITY;
        l = 0.f;
        alpha_prev = 0.f;
      }
      ptx::mbar_wait(&bar_s[g & 1], (g >> 1) & 1);
      ptx::tc_fence_after();
      if (g >= 1) {
        ptx::mbar_wait(&bar_o, (g - 1) & 1);  // O_{g-1} finished: sP and the O tile are free
        ptx::tc_fence_after();
        if (j > 0) add_o_tile();  // (for j == 0 the previous item's epilogue already consumed it)
      }
      float s[32];
      {
        uint32_t v[32];
        ptx::tmem_ld32(tmem + lane_off + (g & 1) * KT + half * 32, v);
        ptx::tmem_ld_wait();
        if (tile_may_be_masked(p, j * KT)) {
#pragma unroll
          for (int i = 0; i < 32; ++i) s[i] = __uint_as_float(v[i]) + key_bias(p, it.b, j * KT + half * 32 + i);
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) s[i] = __uint_as_float(v[i]);
        }
      }
      float mx0 = fmaxf(fmaxf(s[0], s[1]), fmaxf(s[2], s[3])), mx1 = fmaxf(fmaxf(s[4], s[5]), fmaxf(s[6], s[7]));
#pragma unroll
      for (int i = 8; i < 32; i += 8) {
        mx0 = fmaxf(mx0, fmaxf(fmaxf(s[i], s[i + 1]), fmaxf(s[i + 2], s[i + 3])));
        mx1 = fmaxf(mx1, fmaxf(fmaxf(s[i + 4], s[i + 5]), fmaxf(s[i + 6], s[i + 7])));
      }
      float* sx = sX + (g & 1) * 256;  // double-buffered by tile parity: one pair barrier per tile
      sx[half * 128 + r] = fmaxf(mx0, mx1);
      pair_barrier(1 + (warp & 3));
      const float mx = fmaxf(sx[r], sx[128 + r]);
      const float m_new = fmaxf(m, mx);
      const float ms = (m_new == -INFINITY) ? 0.f : m_new * p.scale_log2;
      alpha_prev = ex2(m * p.scale_log2 - ms);
#pragma unroll
      for (int i = 0; i < 32; ++i) s[i] = ex2(s[i] * p.scale_log2 - ms);
      float ls0 = 0.f, ls1 = 0.f;
#pragma unroll
      for (int i = 0; i < 32; i += 8) {
        ls0 += (s[i] + s[i + 1]) + (s[i + 2] + s[i + 3]);
        ls1 += (s[i + 4] + s[i + 5]) + (s[i + 6] + s[i + 7]);
      }
      store_bf16x32(sP + r * 16, half * 32, s);
      l = l * alpha_prev + (ls0 + ls1);
      m = m_new;
      ptx::fence_proxy_async_smem();
      ptx::tc_fence_before();
      ptx::mbar_arrive(&ps_full);
      if (j == T - 1) {  // item epilogue
        ptx::mbar_wait(&bar_o, g & 1);
        ptx::tc_fence_after();
        add_o_tile();
        float* sl = sX + 512 + (k & 1) * 256;
        sl[half * 128 + r] = l;
        pair_barrier(1 + (warp & 3));
        const float lt = sl[r] + sl[128 + r];
        const float inv = lt > 0.f ? 1.f / lt : 0.f;
        const int row = it.x * 128 + r;
        if (half == 0 && p.lse && row < p.S) p.lse[((int64_t)it.b * p.H + it.h) * p.S + row] = m * p.scale + logf(lt);
        compute_barrier();  // the previous item's store_tile has finished reading the staging tile
#pragma unroll
        for (int c = 0; c < HH; c += 8) {
          float t8[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) t8[i] = o[c + i] * inv;
          st8(sStage + r * HDP + half * HH + c, pack8(t8));
        }
        ptx::tc_fence_before();
        compute_barrier();
        store_tile<HDP, O_OUT>(sStage, p, it.b, it.h, it.x * 128, tid);
        j = 0;
        ++k;
      } else {
        ++j;
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp_u == 8) ptx::tmem_dealloc<256>(__shfl_sync(0xffffffffu, tmem_slot, 0));
}

// ---------------------------------------------------------------------------------------------------------
// forward, whole-row version for sequences of at most 256 keys (DiT-XL/2: N = 256; cifar DiT: 64; head dims <= 80).
// The per-tile online softmax of the kernel above pays its fixed latencies (TMEM round trips, pair barrier, fences, mbarrier
// hand-offs: ~2000 cycles for ~160 cycles of tensor work and a 512-cycle MUFU floor) once per 64 keys. Here the WHOLE score row
// S = Q K^T (up to 256 fp32 columns of TMEM) is computed before the softmax starts, the softmax is a plain two-pass one over the
// thread's 128 scores held in registers (row maximum; exponentials, row sum, P -> shared memory), and O = P V is ONE accumulating
// MMA chain into TMEM — no rescaling, no per-tile register accumulation. O is double-buffered in TMEM and the epilogue of item k
// runs after the softmax of item k + 1, so the compute warps never wait for the P V products. K and V tiles travel
// through separate 4-stage rings so that the K tiles of item i+1 load during the softmax of item i (they are released as soon as
// S(i) is computed) and its V tiles during the epilogue of item i. Hand-offs per item: bar_s, ps_full, bar_o.
// TMEM: S at columns [0, 256), O buffers at [256, 256 + 2 HDP). Same warp roles, layouts and tensor maps as the kernels around it.
// ---------------------------------------------------------------------------------------------------------
template <int HDP, bool TMA, int FLAGS>
__global__ void __launch_bounds__(NT, 1) attn_fwd_row_tc_kernel(const Params p, const __grid_constant__ BwdMaps maps) {
  constexpr int TQ = 128 * HDP * 2, TK = KT * HDP * 2, NS = 4, PT = 128 * KT * 2;
  constexpr int HH = HDP / 2;
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* sQ = smem;                                       // 2 x Q tile, layout L1(128)
  uint8_t* sK = sQ + 2 * TQ;                                // NS K tiles, layout L1(64)
  uint8_t* sV = sK + NS * TK;                               // NS V tiles
  uint8_t* sP = sV + NS * TK;                               // 4 x [128 q][64 keys] bf16, layout L1(128)
  bf16* sStage = reinterpret_cast<bf16*>(sP + 4 * PT);      // [128][HDP] output staging
  float* sX = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(sStage) + TQ);  // [2][2][128] pair exchange (max, sum)
  __shared__ uint64_t fullK[NS], emptyK[NS], fullV[NS], emptyV[NS], bar_s, ps_full, bar_o[2];
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
  const int T = (p.S + KT - 1) / KT, nxq = (p.S + 127) / 128;  // T <= 4, nxq <= 2
  // FLAGS (compile time): 1 = the query tiles of a head share one copy of K / V, 2 = deferred epilogue, 4 = TMA store of the output
  // tile, 8 = two-pass softmax that re-reads the scores from TMEM (fewer registers) instead of keeping 128 scores per thread
  constexpr bool share = (FLAGS & 1) != 0, defer = (FLAGS & 2) != 0, tma_out = TMA && (FLAGS & 4) != 0, reread = (FLAGS & 8) != 0;
  const int nx = share ? nxq : 1;  // query tiles per work unit
  // work unit = one (sample, head): its nx query tiles are processed back to back against ONE copy of its K / V tiles in shared
  // memory (the head-slice gather, ~14 B/clk/SM with 16-byte boxes, is what bounds this kernel otherwise: 92 KB per query tile)
  const int nunits = p.H * p.B * (share ? 1 : nxq);
  const int n_my = (nunits - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int NI = n_my * nx;  // items (query tiles) of this CTA: item i = (unit i / nx, tile i % nx)
  auto item_of = [&](int i) {
    const int w = (int)blockIdx.x + (i / nx) * (int)gridDim.x;
    if (!share) return decode_item(p, w, nxq);
    return Item{i % nx, w % p.H, w / p.H};
  };
  if (tid == 0) {
    for (int i = 0; i < NS; ++i) {
      ptx::mbar_init(&fullK[i], TMA ? 1 : NPW * 32); ptx::mbar_init(&emptyK[i], 1);
      ptx::mbar_init(&fullV[i], TMA ? 1 : NPW * 32); ptx::mbar_init(&emptyV[i], 1);
    }
    ptx::mbar_init(&bar_s, 1); ptx::mbar_init(&ps_full, NC); ptx::mbar_init(&bar_o[0], 1); ptx::mbar_init(&bar_o[1], 1);
    ptx::fence_mbar_init();
  }
  if (warp_u == 8) ptx::tmem_alloc<512>(&tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  constexpr uint32_t idesc_s = ptx::make_idesc_bf16(128, KT, false, false);
  constexpr uint32_t idesc_o = ptx::make_idesc_bf16(128, HDP, false, true);

#define ROW_TRACE_C(i, slot) do { if (p.trace != nullptr && tid == 0 && blockIdx.x == gridDim.x / 2 && ((i) == 2 || (i) == 3)) p.trace[((i) - 2) * 10 + (slot)] = clock64(); } while (0)
#define ROW_TRACE_M(i, slot) do { if (p.trace != nullptr && lane == 0 && blockIdx.x == gridDim.x / 2 && ((i) == 2 || (i) == 3)) p.trace[32 + ((i) - 2) * 10 + (slot)] = clock64(); } while (0)
  if (warp_u > 8) {
    // ---------------- producer warps: per item Q + K tiles, then V tiles ----------------
    const int pw = warp_u - 9;
    for (int k = 0; k < n_my; ++k) {
      const Item it = item_of(k * nx);
      for (int j = 0; j < T; ++j) {
        const int u = k * T + j;
        if (u >= NS) ptx::mbar_wait(&emptyK[u % NS], ((u / NS) - 1) & 1);
        uint8_t* st = sK + (u % NS) * TK;
        // the unit's query tiles travel with its LAST key tile: that stage is only released once every score product of the
        // previous unit (which read the query buffers) has completed
        const bool with_q = j == T - 1;
        if constexpr (TMA) {
          if (pw == 0 && ptx::elect_one()) {
            uint64_t* bar = &fullK[u % NS];
            ptx::mbar_expect_tx(bar, TK + (with_q ? nx * TQ : 0));
            if (with_q)
              for (int x = 0; x < nx; ++x) tma_tile(sQ + ((share ? x : k) & 1) * TQ, maps, M_Q128, p, it.b, it.h, (share ? x : it.x) * 128, bar);
            tma_tile(st, maps, M_K64, p, it.b, it.h, j * KT, bar);
          }
          __syncwarp();
        } else {
          if (with_q)
            for (int x = 0; x < nx; ++x) load_tile<HDP, T_Q, 128, NPW>(sQ + ((share ? x : k) & 1) * TQ, p, it.b, it.h, (share ? x : it.x) * 128, pw, lane);
          load_tile<HDP, T_K, KT, NPW>(st, p, it.b, it.h, j * KT, pw, lane);
          cp_async_arrive_noinc(&fullK[u % NS]);
        }
      }
      for (int j = 0; j < T; ++j) {
        const int u = k * T + j;
        if (u >= NS) ptx::mbar_wait(&emptyV[u % NS], ((u / NS) - 1) & 1);
        uint8_t* st = sV + (u % NS) * TK;
        if constexpr (TMA) {
          if (pw == 0 && ptx::elect_one()) {
            uint64_t* bar = &fullV[u % NS];
            ptx::mbar_expect_tx(bar, TK);
            tma_tile(st, maps, M_V64, p, it.b, it.h, j * KT, bar);
          }
          __syncwarp();
        } else {
          load_tile<HDP, T_V, KT, NPW>(st, p, it.b, it.h, j * KT, pw, lane);
          cp_async_arrive_noinc(&fullV[u % NS]);
        }
      }
    }
    if constexpr (!TMA) cp_async_wait<0>();
  } else if (warp_u == 8) {
    // ---------------- MMA-issue warp: S(0); then per item: P V (after the softmax), S(next) ----------------
    const uint32_t tmem = __shfl_sync(0xffffffffu, tmem_slot, 0);
    const uint32_t qa = ptx::smem_u32(sQ), ka = ptx::smem_u32(sK), va = ptx::smem_u32(sV), pa = ptx::smem_u32(sP);
    auto issue_s = [&](int i) {  // S(i) = Q(i) K^T into TMEM columns [0, T * 64)
      const int k = i / nx, x = i % nx;
      const uint64_t dq0 = desc_k128(qa + ((share ? x : k) & 1) * TQ);
      if (x == 0) {  // the unit's query tiles arrive with its last key tile
        const int ul = k * T + T - 1;
        ptx::mbar_wait(&fullK[ul % NS], (ul / NS) & 1);
      }
      for (int j = 0; j < T; ++j) {
        const int u = k * T + j;
        if (x == 0) ptx::mbar_wait(&fullK[u % NS], (u / NS) & 1);
        ptx::fence_proxy_async_smem();
        ptx::tc_fence_after();
        const uint64_t dk = desc_k64(ka + (u % NS) * TK);
#pragma unroll
        for (int ks = 0; ks < HDP / 16; ++ks) ptx::umma_bf16_elect(tmem + j * KT, dq0 + ks * KSTEP_K128, dk + ks * KSTEP_K64, idesc_s, ks > 0);
        if (x == nx - 1) ptx::umma_commit_elect(&emptyK[u % NS]);  // refill as soon as the unit's last scores exist
      }
      ptx::umma_commit_elect(&bar_s);
      ROW_TRACE_M(i, 1);  // S(i) issued
    };
    if (NI > 0) issue_s(0);
    for (int i = 0; i < NI; ++i) {
      const int k = i / nx, x = i % nx;
      ptx::mbar_wait(&ps_full, i & 1);  // P(i) written, S(i) drained
      ptx::tc_fence_after();
      ROW_TRACE_M(i, 2);  // ps_full(i) seen
      for (int j = 0; j < T; ++j) {
        const int u = k * T + j;
        if (x == 0) ptx::mbar_wait(&fullV[u % NS], (u / NS) & 1);
        ptx::fence_proxy_async_smem();
        ptx::tc_fence_after();
        const uint64_t dp0 = desc_k128(pa + j * PT);
        const uint64_t dv = desc_mn64(va + (u % NS) * TK);
#pragma unroll
        for (int ks = 0; ks < KT / 16; ++ks)
          ptx::umma_bf16_elect(tmem + 256 + (i & 1) * HDP, dp0 + ks * KSTEP_K128, dv + ks * KSTEP_MN64, idesc_o, (j > 0 || ks > 0));
        if (x == nx - 1) ptx::umma_commit_elect(&emptyV[u % NS]);
      }
      ptx::umma_commit_elect(&bar_o[i & 1]);
      ROW_TRACE_M(i, 3);  // P V(i) issued
      if (i + 1 < NI) issue_s(i + 1);
    }
  } else {
    // ---------------- compute warps: two threads per row, each owns 32 of the 64 columns of every key tile ----------------
    const int r = tid & 127, half = tid >> 7;
    const uint32_t tmem = tmem_slot;
    const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t tS = tmem + lane_off + half * 32, tO = tmem + 256 + lane_off;
    // epilogue of item k (O(k) / l -> staging -> global). It runs one iteration LATE, after the softmax of item k + 1, so that the
    // compute warps never wait for the P V products (O is double-buffered in TMEM for that).
    auto epilogue = [&](int k, const Item& it, float inv) {
      ptx::mbar_wait(&bar_o[k & 1], (k >> 1) & 1);
      ptx::tc_fence_after();
      ROW_TRACE_C(k + 1, 6);  // O(k) ready
      if constexpr (tma_out) {
        // staging tile in layout L1(128) (16-byte chunk c of row r at c * 2048 + r * 16: conflict-free writes), sent by ONE TMA
        // store through the same kind of 4-D head-slice map the loads use; the padded chunk of a 72-wide head is clipped
        if (tid == 0) ptx::tma_wait_group_read<0>();  // the previous store has finished reading the staging tile
        compute_barrier();
        uint32_t v[HH];
        const uint32_t a = tO + (k & 1) * HDP + half * HH;
        ptx::tmem_ld32(a, v);
        if constexpr (HH == 40) ptx::tmem_ld8(a + 32, v + 32);
        ptx::tmem_ld_wait();
        uint8_t* st = reinterpret_cast<uint8_t*>(sStage);
        (void)maps;
#pragma unroll
        for (int c = 0; c < HH / 8; ++c) {
          float t8[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) t8[j] = __uint_as_float(v[c * 8 + j]) * inv;
          *reinterpret_cast<bf16x8*>(st + (half * (HH / 8) + c) * 2048 + r * 16) = pack8(t8);
        }
        ptx::fence_proxy_async_smem();
        ptx::tc_fence_before();
        compute_barrier();
        ROW_TRACE_C(k + 1, 7);  // staged
        if (tid == 0) {
          const int s0 = it.x * 128, sg = s0 < p.seg[0].len ? 0 : 1;
          const int grow = it.b * p.seg[sg].len + (sg ? s0 - p.seg[0].len : s0);
          ptx::tma_store_4d(&maps.m[sg][M_O128], st, 0, grow, 0, it.h);
          ptx::tma_commit_group();
        }
      } else {
        compute_barrier();  // the previous item's store_tile has finished reading the staging tile
        tmem_half_to_stage<HDP>(tO + (k & 1) * HDP, sStage, r, half, inv);
        ptx::tc_fence_before();
        compute_barrier();
        ROW_TRACE_C(k + 1, 7);  // staged
        store_tile<HDP, O_OUT>(sStage, p, it.b, it.h, it.x * 128, tid);
      }
      ROW_TRACE_C(k + 1, 8);  // stored
    };
    Item it_prev{0, 0, 0};
    float inv_prev = 0.f;
    for (int k = 0; k < NI; ++k) {
      const Item it = item_of(k);
      ROW_TRACE_C(k, 0);  // iteration start
      ptx::mbar_wait(&bar_s, k & 1);
      ptx::tc_fence_after();
      ROW_TRACE_C(k, 1);  // S(k) ready
      float mx = -INFINITY, l0 = 0.f, l1 = 0.f;
      float* sx = sX + (k & 1) * 512;  // double-buffered by item parity
      float m, ms;
      if constexpr (reread) {
        // pass 1: row maximum over my columns of every key tile
        for (int j = 0; j < T; ++j) {
          uint32_t v[32];
          ptx::tmem_ld32(tS + j * KT, v);
          ptx::tmem_ld_wait();
          float m0 = -INFINITY, m1 = -INFINITY;
          if (tile_may_be_masked(p, j * KT)) {
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              m0 = fmaxf(m0, __uint_as_float(v[i]) + key_bias(p, it.b, j * KT + half * 32 + i));
              m1 = fmaxf(m1, __uint_as_float(v[i + 1]) + key_bias(p, it.b, j * KT + half * 32 + i + 1));
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; i += 2) { m0 = fmaxf(m0, __uint_as_float(v[i])); m1 = fmaxf(m1, __uint_as_float(v[i + 1])); }
          }
          mx = fmaxf(mx, fmaxf(m0, m1));
        }
        ROW_TRACE_C(k, 2);
        sx[half * 128 + r] = mx;
        pair_barrier(1 + (warp & 3));
        m = fmaxf(sx[r], sx[128 + r]);
        ms = (m == -INFINITY) ? 0.f : m * p.scale_log2;
        ROW_TRACE_C(k, 3);  // row maximum exchanged
        // pass 2: exponentials, row sum, P -> shared memory (scores re-read from TMEM)
        for (int j = 0; j < T; ++j) {
          uint32_t v[32];
          float e[32];
          ptx::tmem_ld32(tS + j * KT, v);
          ptx::tmem_ld_wait();
          if (tile_may_be_masked(p, j * KT)) {
#pragma unroll
            for (int i = 0; i < 32; ++i) e[i] = ex2((__uint_as_float(v[i]) + key_bias(p, it.b, j * KT + half * 32 + i)) * p.scale_log2 - ms);
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) e[i] = ex2(__uint_as_float(v[i]) * p.scale_log2 - ms);
          }
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
            l0 += (e[i] + e[i + 1]) + (e[i + 2] + e[i + 3]);
            l1 += (e[i + 4] + e[i + 5]) + (e[i + 6] + e[i + 7]);
          }
          store_bf16x32(sP + j * PT + r * 16, half * 32, e);
        }
      } else {
      // the thread's 32 columns of every key tile, read from TMEM once (all loads in flight before the single wait)
      float sv[4][32];
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (j < T) ptx::tmem_ld32(tS + j * KT, reinterpret_cast<uint32_t*>(sv[j]));
      ptx::tmem_ld_wait();
      ROW_TRACE_C(k, 2);  // scores in registers
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (j < T) {
          if (tile_may_be_masked(p, j * KT)) {
#pragma unroll
            for (int i = 0; i < 32; ++i) sv[j][i] += key_bias(p, it.b, j * KT + half * 32 + i);
          }
          float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
          for (int i = 0; i < 32; i += 2) { m0 = fmaxf(m0, sv[j][i]); m1 = fmaxf(m1, sv[j][i + 1]); }
          mx = fmaxf(mx, fmaxf(m0, m1));
        }
      }
      sx[half * 128 + r] = mx;
      pair_barrier(1 + (warp & 3));
      m = fmaxf(sx[r], sx[128 + r]);
      ms = (m == -INFINITY) ? 0.f : m * p.scale_log2;
      ROW_TRACE_C(k, 3);  // row maximum exchanged
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (j < T) {
#pragma unroll
          for (int i = 0; i < 32; ++i) sv[j][i] = ex2(sv[j][i] * p.scale_log2 - ms);
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
            l0 += (sv[j][i] + sv[j][i + 1]) + (sv[j][i + 2] + sv[j][i + 3]);
            l1 += (sv[j][i + 4] + sv[j][i + 5]) + (sv[j][i + 6] + sv[j][i + 7]);
          }
          store_bf16x32(sP + j * PT + r * 16, half * 32, sv[j]);
        }
      }
      }
      sx[256 + half * 128 + r] = l0 + l1;
      ptx::fence_proxy_async_smem();
      ptx::tc_fence_before();
      ptx::mbar_arrive(&ps_full);
      ROW_TRACE_C(k, 4);  // P written, arrived
      pair_barrier(1 + (warp & 3));
      const float lt = sx[256 + r] + sx[256 + 128 + r];
      const float inv = lt > 0.f ? 1.f / lt : 0.f;
      const int row = it.x * 128 + r;
      if (half == 0 && p.lse && row < p.S) p.lse[((int64_t)it.b * p.H + it.h) * p.S + row] = m * p.scale + logf(lt);
      ROW_TRACE_C(k, 5);  // before the epilogue
      if (!defer) epilogue(k, it, inv);
      else if (k > 0) epilogue(k - 1, it_prev, inv_prev);
      ROW_TRACE_C(k, 9);  // after the epilogue
      it_prev = it;
      inv_prev = inv;
    }
    if (defer && NI > 0) epilogue(NI - 1, it_prev, inv_prev);
    if constexpr (tma_out) { if (tid == 0) ptx::tma_wait_group<0>(); }  // the last store has landed before the CTA exits
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp_u == 8) ptx::tmem_dealloc<512>(__shfl_sync(0xffffffffu, tmem_slot, 0));
}

// ---------------------------------------------------------------------------------------------------------
// backward: dQ (and D = rowsum(dO o O))
// ---------------------------------------------------------------------------------------------------------
template <int HDP, int RES, int NST, bool TMA>
__global__ void __launch_bounds__(NT, 1) attn_bwd_dq_tc_kernel(const Params p, const __grid_constant__ BwdMaps maps) {
  constexpr int TQ = 128 * HDP * 2, TK = KT * HDP * 2;
  constexpr int CPR = HDP / 8;
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* sRes = smem;                                    // RES x {Q, dO, O} resident tiles, layout L1(128)
  uint8_t* sKV = sRes + RES * 3 * TQ;                      // NST stages of {K tile, V tile}
  uint8_t* sDS = sKV + NST * 2 * TK;                       // [128 q][64 keys] bf16, layout L1(128), 16 KB
  float* sLrow = reinterpret_cast<float*>(sDS + 16384);    // [RES][128] lse rows (natural log)
  float* sX = sLrow + RES * 128;                           // [2][2][128] pair exchange
  __shared__ uint64_t full[NST], empty[NST], bar1[2], ps_full, bar2;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
  const int T = (p.S + KT - 1) / KT, nx = (p.S + 127) / 128;
  const int nitems = nx * p.H * p.B;
  const int n_my = (nitems - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int G = n_my * T;  // streamed tiles this CTA processes
  if (tid == 0) {
    for (int i = 0; i < NST; ++i) { ptx::mbar_init(&full[i], TMA ? 1 : NPW * 32); ptx::mbar_init(&empty[i], 1); }
    ptx::mbar_init(&bar1[0], 1); ptx::mbar_init(&bar1[1], 1); ptx::mbar_init(&ps_full, NC); ptx::mbar_init(&bar2, 1);
    ptx::fence_mbar_init();
  }
  if (warp_u == 8) ptx::tmem_alloc<512>(&tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  constexpr uint32_t idesc_s = ptx::make_idesc_bf16(128, KT, false, false);
  constexpr uint32_t idesc_dq = ptx::make_idesc_bf16(128, HDP, false, true);

  if (warp_u > 8) {
    // ---------------- producer warps ----------------
    const int pw = warp_u - 9;
    int uk = 0, uj = 0;
    for (int u = 0; u < G; ++u) {
      if (u >= NST) ptx::mbar_wait(&empty[u % NST], ((u / NST) - 1) & 1);
      const Item it = decode_item(p, blockIdx.x + uk * gridDim.x, nx);
      uint8_t* st = sKV + (u % NST) * 2 * TK;
      uint8_t* res = sRes + (uk % RES) * 3 * TQ;
      if constexpr (TMA) {
        if (pw == 0 && ptx::elect_one()) {
          uint64_t* bar = &full[u % NST];
          ptx::mbar_expect_tx(bar, 2 * TK + (uj == 0 ? 3 * TQ + 512 : 0));
          if (uj == 0) {
            tma_tile(res, maps, M_Q128, p, it.b, it.h, it.x * 128, bar);
            tma_tile(res + TQ, maps, M_DO128, p, it.b, it.h, it.x * 128, bar);
            tma_tile(res + 2 * TQ, maps, M_O128, p, it.b, it.h, it.x * 128, bar);
            ptx::bulk_load_1d(sLrow + (uk % RES) * 128, p.lse + ((int64_t)it.b * p.H + it.h) * p.S + it.x * 128, 512, bar);
          }
          tma_tile(st, maps, M_K64, p, it.b, it.h, uj * KT, bar);
          tma_tile(st + TK, maps, M_V64, p, it.b, it.h, uj * KT, bar);
        }
        __syncwarp();
        if (++uj == T) { uj = 0; ++uk; }
        continue;
      }
      if (uj == 0) {
        load_tile<HDP, T_Q, 128, NPW>(res, p, it.b, it.h, it.x * 128, pw, lane);
        load_tile<HDP, T_DO, 128, NPW>(res + TQ, p, it.b, it.h, it.x * 128, pw, lane);
        load_tile<HDP, T_O, 128, NPW>(res + 2 * TQ, p, it.b, it.h, it.x * 128, pw, lane);
#pragma unroll
        for (int i = pw * 32 + lane; i < 128; i += NPW * 32) {
          const int row = it.x * 128 + i;
          const bool v = row < p.S;
          cp_async4_zfill(sLrow + (uk % RES) * 128 + i, p.lse + ((int64_t)it.b * p.H + it.h) * p.S + (v ? row : 0), v);
        }
      }
      load_tile<HDP, T_K, KT, NPW>(st, p, it.b, it.h, uj * KT, pw, lane);
      load_tile<HDP, T_V, KT, NPW>(st + TK, p, it.b, it.h, uj * KT, pw, lane);
      cp_async_arrive_noinc(&full[u % NST]);
      if (++uj == T) { uj = 0; ++uk; }
    }
    cp_async_wait<0>();
  } else if (warp_u == 8) {
    // ---------------- MMA-issue warp ----------------
    const uint32_t tmem = __shfl_sync(0xffffffffu, tmem_slot, 0);
    const uint32_t resa = ptx::smem_u32(sRes), kva = ptx::smem_u32(sKV);
    const uint64_t ds0 = desc_k128(ptx::smem_u32(sDS));
    int tk = 0, tj = 0;  // item / tile-in-item of tile t = g + 1
    int gj = 0;          // tile-in-item of tile g
    for (int g = -1; g < G; ++g) {
      // Scores of tile g+1 go first (their operands are normally resident long before the dS of tile g is written) —
      // unless that tile has not landed yet (item boundary, load-bound phase): then the accumulation of tile g goes first.
      const bool next_ready = g + 1 < G && (g < 0 || __shfl_sync(0xffffffffu, (int)ptx::mbar_test_wait(&full[(g + 1) % NST], ((g + 1) / NST) & 1), 0) != 0);
#pragma unroll 1
      for (int pass = 0; pass < 2; ++pass) {
      if ((pass == 0) == next_ready && g + 1 < G) {  // S(t) = Q K_t^T, dP(t) = dO V_t^T into TMEM buffer t & 1
        const int t = g + 1;
        MMA_TRACE(1, tk, 32 + tj * 4);
        ptx::mbar_wait(&full[t % NST], (t / NST) & 1);
        MMA_TRACE(1, tk, 33 + tj * 4);
        ptx::fence_proxy_async_smem();
        ptx::tc_fence_after();
        const uint32_t ra = resa + (tk % RES) * 3 * TQ;
        const uint64_t dq0 = desc_k128(ra), dd0 = desc_k128(ra + TQ);
        const uint64_t dk = desc_k64(kva + (t % NST) * 2 * TK), dv = desc_k64(kva + (t % NST) * 2 * TK + TK);
        const uint32_t ts = tmem + (t & 1) * 128;
#pragma unroll
        for (int ks = 0; ks < HDP / 16; ++ks) {
          ptx::umma_bf16_elect(ts, dq0 + ks * KSTEP_K128, dk + ks * KSTEP_K64, idesc_s, ks > 0);
          ptx::umma_bf16_elect(ts + KT, dd0 + ks * KSTEP_K128, dv + ks * KSTEP_K64, idesc_s, ks > 0);
        }
        ptx::umma_commit_elect(&bar1[t & 1]);
        if (++tj == T) { tj = 0; ++tk; }
      }
      if (pass == 0 && g >= 0) {  // dQ += dS K_g : A K-major over keys, B = K tile read MN-major (N = hd, K = keys)
        MMA_TRACE(1, g / T, 34 + gj * 4);
        ptx::mbar_wait(&ps_full, g & 1);
        MMA_TRACE(1, g / T, 35 + gj * 4);
        ptx::tc_fence_after();
        const uint64_t dk = desc_mn64(kva + (g % NST) * 2 * TK);
#pragma unroll
        for (int ks = 0; ks < KT / 16; ++ks) ptx::umma_bf16_elect(tmem + 256, ds0 + ks * KSTEP_K128, dk + ks * KSTEP_MN64, idesc_dq, (gj > 0 || ks > 0) ? 1u : 0u);
        ptx::umma_commit_elect(&empty[g % NST]);
        ptx::umma_commit_elect(&bar2);
        if (++gj == T) gj = 0;
      }
      }
    }
  } else {
    // ---------------- compute warps ----------------
    const int r = tid & 127, half = tid >> 7;
    const int trace_item = n_my > 1 ? 1 : 0;
    const uint32_t tmem = tmem_slot;
    const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t tDQ = tmem + 256 + lane_off;
    int k = 0, j = 0;
    Item it = decode_item(p, blockIdx.x, nx);
    float Lrow = 0.f, Drow = 0.f;

    for (int g = 0; g < G; ++g) {
      if (j == 0) {  // per-item setup: D = rowsum(dO o O) from the resident tiles, lse row
        BWD_TRACE(1, 0);
        it = decode_item(p, blockIdx.x + k * gridDim.x, nx);
        ptx::mbar_wait(&full[g % NST], (g / NST) & 1);  // resident tiles (and streamed tile 0) of this item have landed
        const uint8_t* sdo = sRes + (k % RES) * 3 * TQ + TQ + r * 16;
        float dpart = 0.f;
#pragma unroll
        for (int c = half; c < CPR; c += 2) {
          float df[8], of[8];
          unpack8(*reinterpret_cast<const bf16x8*>(sdo + c * 2048), df);
          unpack8(*reinterpret_cast<const bf16x8*>(sdo + TQ + c * 2048), of);
#pragma unroll
          for (int i = 0; i < 8; ++i) dpart += df[i] * of[i];
        }
        float* sx = sX + (k & 1) * 256;  // double-buffered by item parity: one barrier is enough
        sx[half * 128 + r] = dpart;
        pair_barrier(1 + (warp & 3));
        Drow = sx[r] + sx[128 + r];
        const int myrow = it.x * 128 + r;
        Lrow = myrow < p.S ? sLrow[(k % RES) * 128 + r] * LOG2E : INFINITY;
        if (half == 0 && myrow < p.S) p.dsum[((int64_t)it.b * p.H + it.h) * p.S + myrow] = Drow;
        BWD_TRACE(1, 1);
      }
      if (j < 8) BWD_TRACE(1, 4 + j * 6);
      ptx::mbar_wait(&bar1[g & 1], (g >> 1) & 1);
      ptx::tc_fence_after();
      if (j < 8) BWD_TRACE(1, 5 + j * 6);
      if (g >= 1) ptx::mbar_wait(&bar2, (g - 1) & 1);  // dQ += dS K_{g-1} finished: sDS is free
      if (j < 8) BWD_TRACE(1, 6 + j * 6);
      {
        const uint32_t ts = tmem + lane_off + (g & 1) * 128 + half * 32;
        uint32_t vs[32], vd[32];
        ptx::tmem_ld32(ts, vs);
        ptx::tmem_ld32(ts + KT, vd);
        ptx::tmem_ld_wait();
        float ds[32];
        if (tile_may_be_masked(p, j * KT)) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            ds[i] = ex2((__uint_as_float(vs[i]) + key_bias(p, it.b, j * KT + half * 32 + i)) * p.scale_log2 - Lrow) * (__uint_as_float(vd[i]) - Drow);
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) ds[i] = ex2(__uint_as_float(vs[i]) * p.scale_log2 - Lrow) * (__uint_as_float(vd[i]) - Drow);
        }
        store_bf16x32(sDS + r * 16, half * 32, ds);
      }
      ptx::fence_proxy_async_smem();
      ptx::tc_fence_before();
      ptx::mbar_arrive(&ps_full);
      if (j < 8) BWD_TRACE(1, 7 + j * 6);
      if (j == T - 1) {  // item epilogue: dQ tile -> bf16 -> global
        BWD_TRACE(1, 56);
        ptx::mbar_wait(&bar2, g & 1);
        ptx::tc_fence_after();
        BWD_TRACE(1, 57);
        bf16* stage = reinterpret_cast<bf16*>(sRes + (k % RES) * 3 * TQ + 2 * TQ);  // this item's (now dead) O buffer
        tmem_half_to_stage<HDP>(tDQ, stage, r, half, p.scale);
        ptx::tc_fence_before();
        compute_barrier();
        store_tile<HDP, O_DQ>(stage, p, it.b, it.h, it.x * 128, tid);
        BWD_TRACE(1, 59);
        j = 0;
        ++k;
      } else {
        ++j;
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp_u == 8) ptx::tmem_dealloc<512>(__shfl_sync(0xffffffffu, tmem_slot, 0));
}

// ---------------------------------------------------------------------------------------------------------
// backward: dK, dV
// ---------------------------------------------------------------------------------------------------------
template <int HDP, int RES, int NST, bool TMA>
__global__ void __launch_bounds__(NT, 1) attn_bwd_dkv_tc_kernel(const Params p, const __grid_constant__ BwdMaps maps) {
  constexpr int TQ = 128 * HDP * 2, TK = KT * HDP * 2;
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* sRes = smem;                                    // RES x {K, V} resident tiles, layout L1(128)
  uint8_t* sQD = sRes + RES * 2 * TQ;                      // NST stages of {Q tile, dO tile}
  uint8_t* sPT = sQD + NST * 2 * TK;                       // P^T  [128 keys][64 queries] bf16, layout L1(128)
  uint8_t* sDST = sPT + 16384;                             // dS^T
  float* sL = reinterpret_cast<float*>(sDST + 16384);      // [NST][64] lse (natural log)
  float* sD = sL + NST * KT;                               // [NST][64] D
  __shared__ uint64_t full[NST], empty[NST], bar1[2], ps_full, bar2;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
  const int T = (p.S + KT - 1) / KT, nx = (p.S + 127) / 128;
  const int nitems = nx * p.H * p.B;
  const int n_my = (nitems - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int G = n_my * T;
  if (tid == 0) {
    for (int i = 0; i < NST; ++i) { ptx::mbar_init(&full[i], TMA ? 1 : NPW * 32); ptx::mbar_init(&empty[i], 1); }
    ptx::mbar_init(&bar1[0], 1); ptx::mbar_init(&bar1[1], 1); ptx::mbar_init(&ps_full, NC); ptx::mbar_init(&bar2, 1);
    ptx::fence_mbar_init();
  }
  if (warp_u == 8) ptx::tmem_alloc<512>(&tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  constexpr uint32_t idesc_s = ptx::make_idesc_bf16(128, KT, false, false);
  constexpr uint32_t idesc_o = ptx::make_idesc_bf16(128, HDP, false, true);

  if (warp_u > 8) {
    // ---------------- producer warps ----------------
    const int pw = warp_u - 9;
    int uk = 0, uj = 0;
    for (int u = 0; u < G; ++u) {
      if (u >= NST) ptx::mbar_wait(&empty[u % NST], ((u / NST) - 1) & 1);
      const Item it = decode_item(p, blockIdx.x + uk * gridDim.x, nx);
      uint8_t* st = sQD + (u % NST) * 2 * TK;
      uint8_t* res = sRes + (uk % RES) * 2 * TQ;
      const int64_t base = ((int64_t)it.b * p.H + it.h) * p.S;
      if constexpr (TMA) {
        if (pw == 0 && ptx::elect_one()) {
          uint64_t* bar = &full[u % NST];
          ptx::mbar_expect_tx(bar, 2 * TK + 2 * KT * 4 + (uj == 0 ? 2 * TQ : 0));
          if (uj == 0) {
            tma_tile(res, maps, M_K128, p, it.b, it.h, it.x * 128, bar);
            tma_tile(res + TQ, maps, M_V128, p, it.b, it.h, it.x * 128, bar);
          }
          tma_tile(st, maps, M_Q64, p, it.b, it.h, uj * KT, bar);
          tma_tile(st + TK, maps, M_DO64, p, it.b, it.h, uj * KT, bar);
          ptx::bulk_load_1d(sL + (u % NST) * KT, p.lse + base + uj * KT, KT * 4, bar);
          ptx::bulk_load_1d(sD + (u % NST) * KT, p.dsum + base + uj * KT, KT * 4, bar);
        }
        __syncwarp();
        if (++uj == T) { uj = 0; ++uk; }
        continue;
      }
      if (uj == 0) {
        load_tile<HDP, T_K, 128, NPW>(res, p, it.b, it.h, it.x * 128, pw, lane);
        load_tile<HDP, T_V, 128, NPW>(res + TQ, p, it.b, it.h, it.x * 128, pw, lane);
      }
      load_tile<HDP, T_Q, KT, NPW>(st, p, it.b, it.h, uj * KT, pw, lane);
      load_tile<HDP, T_DO, KT, NPW>(st + TK, p, it.b, it.h, uj * KT, pw, lane);
      {
#pragma unroll
        for (int i = pw * 32 + lane; i < 2 * KT; i += NPW * 32) {
          const int q = i & (KT - 1), s2 = uj * KT + q;
          const bool v = s2 < p.S;
          cp_async4_zfill((i < KT ? sL : sD) + (u % NST) * KT + q, (i < KT ? p.lse : p.dsum) + base + (v ? s2 : 0), v);
        }
      }
      cp_async_arrive_noinc(&full[u % NST]);
      if (++uj == T) { uj = 0; ++uk; }
    }
    cp_async_wait<0>();
  } else if (warp_u == 8) {
    // ---------------- MMA-issue warp ----------------
    const uint32_t tmem = __shfl_sync(0xffffffffu, tmem_slot, 0);
    const uint32_t resa = ptx::smem_u32(sRes), qda = ptx::smem_u32(sQD);
    const uint64_t dp0 = desc_k128(ptx::smem_u32(sPT)), ds0 = desc_k128(ptx::smem_u32(sDST));
    int tk = 0, tj = 0, gj = 0;
    for (int g = -1; g < G; ++g) {
      const bool next_ready = g + 1 < G && (g < 0 || __shfl_sync(0xffffffffu, (int)ptx::mbar_test_wait(&full[(g + 1) % NST], ((g + 1) / NST) & 1), 0) != 0);
#pragma unroll 1
      for (int pass = 0; pass < 2; ++pass) {
      if ((pass == 0) == next_ready && g + 1 < G) {  // S^T(t) = K Q_t^T, dP^T(t) = V dO_t^T into TMEM buffer t & 1
        const int t = g + 1;
        MMA_TRACE(2, tk, 32 + tj * 4);
        ptx::mbar_wait(&full[t % NST], (t / NST) & 1);
        MMA_TRACE(2, tk, 33 + tj * 4);
        ptx::fence_proxy_async_smem();
        ptx::tc_fence_after();
        const uint32_t ra = resa + (tk % RES) * 2 * TQ;
        const uint64_t dk0 = desc_k128(ra), dv0 = desc_k128(ra + TQ);
        const uint64_t dq = desc_k64(qda + (t % NST) * 2 * TK), dd = desc_k64(qda + (t % NST) * 2 * TK + TK);
        const uint32_t ts = tmem + (t & 1) * 128;
#pragma unroll
        for (int ks = 0; ks < HDP / 16; ++ks) {
          ptx::umma_bf16_elect(ts, dk0 + ks * KSTEP_K128, dq + ks * KSTEP_K64, idesc_s, ks > 0);
          ptx::umma_bf16_elect(ts + KT, dv0 + ks * KSTEP_K128, dd + ks * KSTEP_K64, idesc_s, ks > 0);
        }
        ptx::umma_commit_elect(&bar1[t & 1]);
        if (++tj == T) { tj = 0; ++tk; }
      }
      if (pass == 0 && g >= 0) {  // dV += P^T dO_g, dK += dS^T Q_g: contraction over the 64 queries; Q / dO tiles read MN-major (N = hd)
        MMA_TRACE(2, g / T, 34 + gj * 4);
        ptx::mbar_wait(&ps_full, g & 1);
        MMA_TRACE(2, g / T, 35 + gj * 4);
        ptx::tc_fence_after();
        const uint64_t dq = desc_mn64(qda + (g % NST) * 2 * TK), dd = desc_mn64(qda + (g % NST) * 2 * TK + TK);
#pragma unroll
        for (int ks = 0; ks < KT / 16; ++ks) {
          const uint32_t acc = (gj > 0 || ks > 0) ? 1u : 0u;
          ptx::umma_bf16_elect(tmem + 256 + HDP, dp0 + ks * KSTEP_K128, dd + ks * KSTEP_MN64, idesc_o, acc);
          ptx::umma_bf16_elect(tmem + 256, ds0 + ks * KSTEP_K128, dq + ks * KSTEP_MN64, idesc_o, acc);
        }
        ptx::umma_commit_elect(&empty[g % NST]);
        ptx::umma_commit_elect(&bar2);
        if (++gj == T) gj = 0;
      }
      }
    }
  } else {
    // ---------------- compute warps ----------------
    const int r = tid & 127, half = tid >> 7;
    const int trace_item = n_my > 1 ? 1 : 0;
    const uint32_t tmem = tmem_slot;
    const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t tDK = tmem + 256 + lane_off, tDV = tmem + 256 + HDP + lane_off;
    int k = 0, j = 0;
    Item it = decode_item(p, blockIdx.x, nx);
    float kbias = 0.f;

    for (int g = 0; g < G; ++g) {
      if (j == 0) {
        BWD_TRACE(2, 0);
        it = decode_item(p, blockIdx.x + k * gridDim.x, nx);
        kbias = (it.x * 128 + 128 > p.S || (p.kmask != nullptr && it.x * 128 < p.mask_len)) ? key_bias(p, it.b, it.x * 128 + r) : 0.f;
      }
      if (j < 8) BWD_TRACE(2, 4 + j * 6);
      ptx::mbar_wait(&bar1[g & 1], (g >> 1) & 1);
      ptx::tc_fence_after();
      if (j < 8) BWD_TRACE(2, 5 + j * 6);
      if (g >= 1) ptx::mbar_wait(&bar2, (g - 1) & 1);  // dV / dK accumulation of tile g-1 finished: sPT and sDST are free
      if (j < 8) BWD_TRACE(2, 6 + j * 6);
      {
        const float* Lq = sL + (g % NST) * KT + half * 32;
        const float* Dq = sD + (g % NST) * KT + half * 32;
        const uint32_t ts = tmem + lane_off + (g & 1) * 128 + half * 32;
        uint32_t vs[32], vd[32];
        ptx::tmem_ld32(ts, vs);
        ptx::tmem_ld32(ts + KT, vd);
        ptx::tmem_ld_wait();
        float pv[32], ds[32];
        const int nq = p.S - (j * KT + half * 32);  // queries of my 32 columns that exist
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          pv[i] = ex2(fmaf(Lq[i], -LOG2E, (__uint_as_float(vs[i]) + kbias) * p.scale_log2));
          if (i >= nq) pv[i] = 0.f;
          ds[i] = pv[i] * (__uint_as_float(vd[i]) - Dq[i]);
        }
        store_bf16x32(sPT + r * 16, half * 32, pv);
        store_bf16x32(sDST + r * 16, half * 32, ds);
      }
      ptx::fence_proxy_async_smem();
      ptx::tc_fence_before();
      ptx::mbar_arrive(&ps_full);
      if (j < 8) BWD_TRACE(2, 7 + j * 6);
      if (j == T - 1) {  // item epilogue: dK, dV tiles -> bf16 -> global
        BWD_TRACE(2, 56);
        ptx::mbar_wait(&bar2, g & 1);
        ptx::tc_fence_after();
        BWD_TRACE(2, 57);
        bf16* stage = reinterpret_cast<bf16*>(sPT);  // the (now idle) P^T / dS^T tiles
        tmem_half_to_stage<HDP>(tDK, stage, r, half, p.scale);
        compute_barrier();
        store_tile<HDP, O_DK>(stage, p, it.b, it.h, it.x * 128, tid);
        compute_barrier();
        tmem_half_to_stage<HDP>(tDV, stage, r, half, 1.f);
        ptx::tc_fence_before();
        compute_barrier();
        store_tile<HDP, O_DV>(stage, p, it.b, it.h, it.x * 128, tid);
        compute_barrier();  // the staging tile becomes P^T / dS^T again
        BWD_TRACE(2, 59);
        j = 0;
        ++k;
      } else {
        ++j;
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp_u == 8) ptx::tmem_dealloc<512>(__shfl_sync(0xffffffffu, tmem_slot, 0));
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

#include <cudaTypedefs.h>
#include <cstdlib>
#include <unordered_map>

namespace {
struct MapKey {
  const void* ptr; int64_t rows, ld; int H, hd, box_rows, cprb;
  bool operator==(const MapKey& o) const {
    return ptr == o.ptr && rows == o.rows && ld == o.ld && H == o.H && hd == o.hd && box_rows == o.box_rows && cprb == o.cprb;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    size_t h = std::hash<const void*>()(k.ptr);
    for (int64_t v : {k.rows, k.ld, (int64_t)k.H, (int64_t)k.hd, (int64_t)k.box_rows, (int64_t)k.cprb}) h = h * 1000003u ^ std::hash<int64_t>()(v);
    return h;
  }
};
// head-slice tensor map of a packed bf16 [rows, ld] activation (cached: the training loop reuses its buffers)
int head_map(CUtensorMap* out, const void* base, int64_t rows, int64_t ld, int H, int hd, int box_rows, int cprb) {
  static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
  static PFN_cuTensorMapEncodeTiled_v12000 enc = nullptr;
  const MapKey key{base, rows, ld, H, hd, box_rows, cprb};
  auto it = cache.find(key);
  if (it != cache.end()) { *out = it->second; return DLB_OK; }
  if (!enc) {
    void* fp = nullptr;
    cudaDriverEntryPointQueryResult q;
    const bool ok = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess;
    DLB_REQUIRE(ok, DLB_ERR_DRIVER, "attn_bwd_tc: cuTensorMapEncodeTiled unavailable");
    enc = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fp);
  }
  cuuint64_t dims[4] = {8, (cuuint64_t)rows, (cuuint64_t)(hd / 8), (cuuint64_t)H};
  cuuint64_t strides[3] = {(cuuint64_t)ld * 2, 16, (cuuint64_t)hd * 2};
  cuuint32_t box[4] = {8, (cuuint32_t)box_rows, (cuuint32_t)cprb, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DLB_REQUIRE(r == CUDA_SUCCESS, DLB_ERR_DRIVER, "attn_bwd_tc: cuTensorMapEncodeTiled failed (%d)", (int)r);
  if (cache.size() > 4096) cache.clear();
  cache.emplace(key, *out);
  return DLB_OK;
}
}  // namespace

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

// launch one compile-time variant of the whole-row forward kernel (uses the locals of dlb_attn_fwd_tc)
#define DLB_ROW_LAUNCH(FL)                                                                                                      \
  case FL: {                                                                                                                    \
    if (use_tma) {                                                                                                              \
      cudaError_t e = cudaFuncSetAttribute(attn_fwd_row_tc_kernel<HDPV, true, FL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
      DLB_REQUIRE(e == cudaSuccess, (int)e, "attn_fwd_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));                    \
      attn_fwd_row_tc_kernel<HDPV, true, FL><<<grid_r, NT, smem, stream>>>(p, maps_r);                                          \
    } else {                                                                                                                    \
      cudaError_t e = cudaFuncSetAttribute(attn_fwd_row_tc_kernel<HDPV, false, FL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
      DLB_REQUIRE(e == cudaSuccess, (int)e, "attn_fwd_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));                    \
      attn_fwd_row_tc_kernel<HDPV, false, FL><<<grid_r, NT, smem, stream>>>(p, maps_r);                                         \
    }                                                                                                                           \
    break;                                                                                                                      \
  }

// Same contract as dlb_attn_fwd (attention.cu); tcgen05 implementation.
DLB_EXPORT int dlb_attn_fwd_tc(const dlb_attn_seg* segs, int nseg, float* lse, const uint8_t* kmask, int mask_len, int B,
                               int H, int hd, float scale, cudaStream_t stream) {
  using namespace attn_tc;
  Params p{};
  int rc = fill_tc_params(p, "attn_fwd_tc", segs, nseg, lse, nullptr, kmask, mask_len, B, H, hd, scale, false);
  if (rc) return rc;
  static const bool simple = getenv("DLB_ATTN_FWD_SIMPLE") != nullptr;  // the 256-thread kernel (kept for comparison)
  if (!simple) {
    const int nitems = ((p.S + 127) / 128) * H * B, T = (p.S + 63) / 64;
    const int sms = dlb_num_sms();
    bool use_tma = true;
    for (int i = 0; i < nseg; ++i) use_tma = use_tma && segs[i].len > 0 && segs[i].len % 128 == 0;
    if (getenv("DLB_ATTN_NO_TMA") != nullptr) use_tma = false;  // tests: force the cp.async producers on TMA-eligible shapes
    static const bool no_row = getenv("DLB_ATTN_NO_ROW") != nullptr;  // tests / A-B: force the per-tile online-softmax kernel
    const int hdp = (hd + 15) / 16 * 16;
    if (!no_row && T <= 4 && hdp <= 80) {  // whole-row softmax kernel: at most 256 keys, head dim <= 80
      HDP_SWITCH_TC(hd, {
        if constexpr (HDPV <= 80) {
          const size_t smem = (size_t)2 * 128 * HDPV * 2 + (size_t)8 * 64 * HDPV * 2 + 4 * 16384 + (size_t)128 * HDPV * 2 + 1024 * 4;
          static const int row_flags = getenv("DLB_ATTN_ROW_FLAGS") ? atoi(getenv("DLB_ATTN_ROW_FLAGS")) : 10;
          const int nunits = (row_flags & 1) ? H * B : nitems;  // bit 0: one (sample, head) per work unit, its query tiles share K / V
          const int grid_r = (T == 4 && nunits > sms) ? sms : nunits;  // persistent only when a unit fills both rings exactly
          static BwdMaps maps_r;
          if (use_tma) {
            for (int i = 0; i < nseg; ++i) {
              const dlb_attn_seg& g = segs[i];
              const int64_t rows = (int64_t)B * g.len;
              rc = head_map(&maps_r.m[i][M_Q128], g.q, rows, g.ldq, H, hd, 128, HDPV / 8);
              if (!rc) rc = head_map(&maps_r.m[i][M_K64], g.k, rows, g.ldk, H, hd, 64, HDPV / 8);
              if (!rc) rc = head_map(&maps_r.m[i][M_V64], g.v, rows, g.ldv, H, hd, 64, HDPV / 8);
              if (!rc) rc = head_map(&maps_r.m[i][M_O128], g.o, rows, g.ldo, H, hd, 128, HDPV / 8);  // output tile (TMA store)
              if (rc) return rc;
            }
          }
          switch (row_flags) {
            // same-box sweep of all variants (profiles/attn_r2/row_kernel_variants.md): 0.129-0.149 ms at the DiT-XL/2 geometry against
            // 0.149 ms for the per-tile kernel; 10 (two-pass re-read + deferred epilogue, 79 registers) was the fastest
            DLB_ROW_LAUNCH(8) DLB_ROW_LAUNCH(10) DLB_ROW_LAUNCH(14) DLB_ROW_LAUNCH(15)
            default: DLB_REQUIRE(false, DLB_ERR_UNSUPPORTED, "attn_fwd_tc: row-kernel variant %d is not instantiated", row_flags);
          }
        }
      });
      dlb_count_launch();
      return dlb_check_launch("attn_fwd_row_tc");
    }
    HDP_SWITCH_TC(hd, {
      constexpr int RES_F = HDPV <= 80 ? 2 : 1, NST_F = HDPV <= 96 ? 4 : 3;
      const size_t smem = (size_t)RES_F * 128 * HDPV * 2 + (size_t)NST_F * 2 * 64 * HDPV * 2 + 16384 + (size_t)128 * HDPV * 2 + 1024 * 4;
      const int grid_f = (RES_F == 2 && T >= NST_F && nitems > sms) ? sms : nitems;
      static BwdMaps maps;
      if (use_tma) {
        for (int i = 0; i < nseg; ++i) {
          const dlb_attn_seg& g = segs[i];
          const int64_t rows = (int64_t)B * g.len;
          rc = head_map(&maps.m[i][M_Q128], g.q, rows, g.ldq, H, hd, 128, HDPV / 8);
          if (!rc) rc = head_map(&maps.m[i][M_K64], g.k, rows, g.ldk, H, hd, 64, HDPV / 8);
          if (!rc) rc = head_map(&maps.m[i][M_V64], g.v, rows, g.ldv, H, hd, 64, HDPV / 8);
          if (rc) return rc;
        }
        cudaError_t e = cudaFuncSetAttribute(attn_fwd_ws_tc_kernel<HDPV, RES_F, NST_F, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        DLB_REQUIRE(e == cudaSuccess, (int)e, "attn_fwd_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        attn_fwd_ws_tc_kernel<HDPV, RES_F, NST_F, true><<<grid_f, NT, smem, stream>>>(p, maps);
      } else {
        cudaError_t e = cudaFuncSetAttribute(attn_fwd_ws_tc_kernel<HDPV, RES_F, NST_F, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        DLB_REQUIRE(e == cudaSuccess, (int)e, "attn_fwd_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        attn_fwd_ws_tc_kernel<HDPV, RES_F, NST_F, false><<<grid_f, NT, smem, stream>>>(p, maps);
      }
    });
    dlb_count_launch();
    return dlb_check_launch("attn_fwd_ws_tc");
  }
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
  const int nitems = ((p.S + 127) / 128) * H * B, T = (p.S + 63) / 64;
  const int sms = dlb_num_sms();
  bool use_tma = true;  // TMA producer: every tile lies inside one segment and lse / dsum rows are 16-byte aligned
  for (int i = 0; i < nseg; ++i) use_tma = use_tma && segs[i].len > 0 && segs[i].len % 128 == 0;
    if (getenv("DLB_ATTN_NO_TMA") != nullptr) use_tma = false;  // tests: force the cp.async producers on TMA-eligible shapes
  HDP_SWITCH_TC(hd, {
    // resident double-buffering (persistent CTAs) where shared memory allows; otherwise one item per CTA
    constexpr int RES_DQ = HDPV <= 80 ? 2 : 1, NST_DQ = HDPV <= 96 ? 4 : 3;
    constexpr int RES_DKV = HDPV <= 96 ? 2 : 1, NST_DKV = 4;
    const size_t sm_dq = (size_t)RES_DQ * 3 * 128 * HDPV * 2 + (size_t)NST_DQ * 2 * 64 * HDPV * 2 + 16384 + RES_DQ * 128 * 4 + 512 * 4;
    const size_t sm_dkv = (size_t)RES_DKV * 2 * 128 * HDPV * 2 + (size_t)NST_DKV * 2 * 64 * HDPV * 2 + 32768 + 2 * NST_DKV * 64 * 4;
    const int grid_dq = (RES_DQ == 2 && T >= NST_DQ && nitems > sms) ? sms : nitems;
    const int grid_dkv = (RES_DKV == 2 && T >= NST_DKV && nitems > sms) ? sms : nitems;
    static BwdMaps maps;  // by-value kernel parameter; contents only read on the TMA path
    if (use_tma) {
      for (int i = 0; i < nseg; ++i) {
        const dlb_attn_seg& g = segs[i];
        const int64_t rows = (int64_t)B * g.len;
        const struct { int idx; const void* ptr; int64_t ld; int box; } want[M_COUNT] = {
            {M_Q64, g.q, g.ldq, 64}, {M_Q128, g.q, g.ldq, 128}, {M_K64, g.k, g.ldk, 64}, {M_K128, g.k, g.ldk, 128},
            {M_V64, g.v, g.ldv, 64}, {M_V128, g.v, g.ldv, 128}, {M_DO64, g.dout, g.lddo, 64}, {M_DO128, g.dout, g.lddo, 128},
            {M_O128, g.o, g.ldo, 128}};
        for (const auto& w : want) {
          rc = head_map(&maps.m[i][w.idx], w.ptr, rows, w.ld, H, hd, w.box, HDPV / 8);
          if (rc) return rc;
        }
      }
    }
    cudaError_t e = cudaSuccess;
    if (use_tma) {
      e = cudaFuncSetAttribute(attn_bwd_dq_tc_kernel<HDPV, RES_DQ, NST_DQ, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_dq);
      if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_bwd_dkv_tc_kernel<HDPV, RES_DKV, NST_DKV, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_dkv);
    } else {
      e = cudaFuncSetAttribute(attn_bwd_dq_tc_kernel<HDPV, RES_DQ, NST_DQ, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_dq);
      if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_bwd_dkv_tc_kernel<HDPV, RES_DKV, NST_DKV, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_dkv);
    }
    DLB_REQUIRE(e == cudaSuccess, (int)e, "attn_bwd_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    if (use_tma) {
      attn_bwd_dq_tc_kernel<HDPV, RES_DQ, NST_DQ, true><<<grid_dq, NT, sm_dq, stream>>>(p, maps);
      attn_bwd_dkv_tc_kernel<HDPV, RES_DKV, NST_DKV, true><<<grid_dkv, NT, sm_dkv, stream>>>(p, maps);
    } else {
      attn_bwd_dq_tc_kernel<HDPV, RES_DQ, NST_DQ, false><<<grid_dq, NT, sm_dq, stream>>>(p, maps);
      attn_bwd_dkv_tc_kernel<HDPV, RES_DKV, NST_DKV, false><<<grid_dkv, NT, sm_dkv, stream>>>(p, maps);
    }
  });
  dlb_count_launch(2);
  return dlb_check_launch("attn_bwd_tc");
}
