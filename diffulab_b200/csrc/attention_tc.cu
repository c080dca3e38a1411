// tcgen05 joint attention, forward and backward, sm_100a.
//
// Replaces F.scaled_dot_product_attention as called by DiTAttention / MMDiTAttention (reference mmdit.py:92-98,
// 184-204): softmax(q k^T * hd^-1/2 + key_padding_mask) v over the concatenation of up to two segments (text rows
// first, then image rows), bf16 operands, fp32 softmax. Q / K arrive RMS-normalised and rotated (qknorm_rope.cu), V is
// read in place from the packed qkv projection.
//
// Structure (all three kernels): one CTA = one 128-row tile of one (sample, head); TMEM lane = tile row; 256 threads,
// i.e. TWO threads per row, each owning one half of the columns of every TMEM tile (warps w and w+4 share a lane
// quarter). One elected thread issues the tcgen05.mma instructions; operands are staged by the threads (cp.async,
// 16-byte chunks) because sequences are segment concatenations with a key mask.
// Every [128 x HDP] operand tile uses ONE shared-memory layout ("L1": 16-byte chunk (row r, chunk c) at c*2048 + r*16),
// a valid non-swizzled UMMA layout both K-major (LBO 2048, SBO 128) and MN-major (LBO 128, SBO 2048) — pinned by
// tests/test_umma_probe_gpu.py — so Q / dO / K / V serve as row operands of one product and as transposed operands of
// another without data movement. P / dS tiles ([128 x 128] bf16) are written by their owning threads in the same layout.
//   forward : S = Q K^T -> online softmax (row max exchanged between the two column halves) -> P (smem) -> O_j = P V_j,
//             per-tile result read from TMEM and accumulated (rescaled) in registers
//   dq      : S = Q K^T, dP = dO V^T, dS = P o (dP - D), dQ += dS K         (also produces D = rowsum(dO o O))
//   dkv     : S^T = K Q^T, dP^T = V dO^T, dV += P^T dO, dK += dS^T Q
// Head dims that are not a multiple of 16 (DiT-XL/2: 72) are zero-padded in shared memory only.
#include "common.cuh"
#include "ptx.cuh"

namespace attn_tc {
typedef __nv_bfloat16 bf16;

constexpr int NT = 256;  // threads per CTA

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
};

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(src));
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory"); }

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

// Two threads stage one row of a 128-row tile into layout L1: thread (r, half) copies chunks half, half+2, ...
template <int HDP, int WHICH>
__device__ __forceinline__ void load_tile(uint8_t* sm, const Params& p, int b, int h, int s0, int r, int half) {
  constexpr int CPR = HDP / 8;
  uint8_t* dst = sm + r * 16;
  const int s = s0 + r;
  const int nvalid = p.hd >> 3;
  if (s < p.S) {
    const bf16* src = row_ptr(p, WHICH, b, h, s);
#pragma unroll
    for (int c0 = 0; c0 < CPR; c0 += 2) {
      const int c = c0 + half;
      if (c < CPR) {
        if (c < nvalid) cp_async16(dst + c * 2048, src + c * 8);
        else *reinterpret_cast<uint4*>(dst + c * 2048) = make_uint4(0, 0, 0, 0);
      }
    }
  } else {
#pragma unroll
    for (int c0 = 0; c0 < CPR; c0 += 2) {
      const int c = c0 + half;
      if (c < CPR) *reinterpret_cast<uint4*>(dst + c * 2048) = make_uint4(0, 0, 0, 0);
    }
  }
}

// row-major bf16 staging tile [128][HDP] -> global (valid rows / columns only), coalesced 16-byte stores
template <int HDP, int WHICH>
__device__ __forceinline__ void store_tile(const bf16* stage, const Params& p, int b, int h, int s0, int tid) {
  const int cpr = p.hd >> 3;
  for (int idx = tid; idx < 128 * cpr; idx += NT) {
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

__device__ __forceinline__ void store_bf16x32(uint8_t* rowbase, int c, const float* v) {  // 32 columns -> 4 chunks of L1
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

// ---------------------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------------------
template <int HDP>
__global__ void __launch_bounds__(NT) attn_fwd_tc_kernel(const Params p) {
  constexpr int TILE = 128 * HDP * 2;
  constexpr int HH = HDP / 2;  // O columns per thread
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + TILE;
  uint8_t* sV = sK + TILE;
  uint8_t* sP = sV + TILE;                                 // [128 q][128 keys] bf16, layout L1, 32 KB (reused as O staging)
  float* sBias = reinterpret_cast<float*>(sP + 32768);     // 128 additive key biases (0 / -inf)
  float* sX = sBias + 128;                                 // [2][128] exchange between the two column halves
  __shared__ uint64_t bar_s, bar_o;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, r = tid & 127, half = tid >> 7;
  const int q0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
  if (tid == 0) { ptx::mbar_init(&bar_s, 1); ptx::mbar_init(&bar_o, 1); ptx::fence_mbar_init(); }
  if (warp == 0) ptx::tmem_alloc<256>(&tmem_slot);
  load_tile<HDP, T_Q>(sQ, p, b, h, q0, r, half);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
  const uint32_t tS = tmem + lane_off, tO = tmem + 128 + lane_off;
  constexpr uint32_t idesc_s = ptx::make_idesc_bf16(128, 128, false, false);
  constexpr uint32_t idesc_o = ptx::make_idesc_bf16(128, HDP, false, true);
  float o[HH];
#pragma unroll
  for (int i = 0; i < HH; ++i) o[i] = 0.f;
  float m = -INFINITY, l = 0.f;
  uint32_t phase = 0;
  const int cbase = half * 64;  // my S columns

  for (int kv0 = 0; kv0 < p.S; kv0 += 128) {
    load_tile<HDP, T_K>(sK, p, b, h, kv0, r, half);
    load_tile<HDP, T_V>(sV, p, b, h, kv0, r, half);
    float my_bias = 0.f;
    if (half == 0) { my_bias = key_bias(p, b, kv0 + r); sBias[r] = my_bias; }
    cp_async_wait_all();
    ptx::fence_proxy_async_smem();
    ptx::tc_fence_before();
    const bool masked_tile = __syncthreads_or(my_bias != 0.f);  // uniform: any masked / out-of-range key in this tile
    if (tid == 0) {
      ptx::tc_fence_after();
      const uint32_t qa = ptx::smem_u32(sQ), ka = ptx::smem_u32(sK);
#pragma unroll
      for (int ks = 0; ks < HDP / 16; ++ks)
        ptx::umma_bf16(tmem, ptx::make_smem_desc_noswz(qa + ks * 4096, 2048, 128), ptx::make_smem_desc_noswz(ka + ks * 4096, 2048, 128), idesc_s, ks > 0);
      ptx::umma_commit(&bar_s);
    }
    ptx::mbar_wait(&bar_s, phase);
    ptx::tc_fence_after();
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
    __syncthreads();
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
    ptx::fence_proxy_async_smem();
    ptx::tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      ptx::tc_fence_after();
      const uint32_t pa = ptx::smem_u32(sP), va = ptx::smem_u32(sV);
#pragma unroll
      for (int ks = 0; ks < 8; ++ks)
        ptx::umma_bf16(tmem + 128, ptx::make_smem_desc_noswz(pa + ks * 4096, 2048, 128), ptx::make_smem_desc_noswz(va + ks * 256, 128, 2048), idesc_o, ks > 0);
      ptx::umma_commit(&bar_o);
    }
    ptx::mbar_wait(&bar_o, phase);
    ptx::tc_fence_after();
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
  }
  // finalise: total row sum from both halves, normalise, stage, store
  sX[half * 128 + r] = l;
  __syncthreads();
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
  store_tile<HDP, O_OUT>(stage, p, b, h, q0, tid);
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc<256>(tmem);
}

// ---------------------------------------------------------------------------------------------------------
// backward: dQ (and D = rowsum(dO o O))
// ---------------------------------------------------------------------------------------------------------
template <int HDP>
__global__ void __launch_bounds__(NT) attn_bwd_dq_tc_kernel(const Params p) {
  constexpr int TILE = 128 * HDP * 2;
  constexpr int CPR = HDP / 8;
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* sQ = smem;
  uint8_t* sDO = sQ + TILE;
  uint8_t* sK = sDO + TILE;
  uint8_t* sV = sK + TILE;
  uint8_t* sDS = sV + TILE;                                // [128 q][128 keys] bf16, layout L1, 32 KB
  float* sBias = reinterpret_cast<float*>(sDS + 32768);    // 128 key biases
  float* sX = sBias + 128;                                 // [2][128]
  __shared__ uint64_t bar1, bar2;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, r = tid & 127, half = tid >> 7;
  const int q0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
  if (tid == 0) { ptx::mbar_init(&bar1, 1); ptx::mbar_init(&bar2, 1); ptx::fence_mbar_init(); }
  if (warp == 0) ptx::tmem_alloc<512>(&tmem_slot);
  load_tile<HDP, T_Q>(sQ, p, b, h, q0, r, half);
  // dO row: staged through registers so that D = rowsum(dO o O) comes for free (each half sums its chunks)
  float dpart = 0.f;
  {
    const int s = q0 + r;
    const int nvalid = p.hd >> 3;
    uint8_t* dst = sDO + r * 16;
    const bool valid = s < p.S;
    const bf16* dsrc = valid ? row_ptr(p, T_DO, b, h, s) : nullptr;
    const bf16* osrc = valid ? row_ptr(p, T_O, b, h, s) : nullptr;
#pragma unroll
    for (int c0 = 0; c0 < CPR; c0 += 2) {
      const int c = c0 + half;
      if (c < CPR) {
        if (valid && c < nvalid) {
          const bf16x8 dv = ld8(dsrc + c * 8), ov = ld8(osrc + c * 8);
          float df[8], of[8];
          unpack8(dv, df); unpack8(ov, of);
#pragma unroll
          for (int j = 0; j < 8; ++j) dpart += df[j] * of[j];
          *reinterpret_cast<bf16x8*>(dst + c * 2048) = dv;
        } else {
          *reinterpret_cast<uint4*>(dst + c * 2048) = make_uint4(0, 0, 0, 0);
        }
      }
    }
  }
  sX[half * 128 + r] = dpart;
  const int myrow = q0 + r;
  const float Lrow = myrow < p.S ? p.lse[((int64_t)b * p.H + h) * p.S + myrow] * 1.4426950408889634f : INFINITY;
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const float Drow = sX[r] + sX[128 + r];
  if (half == 0 && myrow < p.S) p.dsum[((int64_t)b * p.H + h) * p.S + myrow] = Drow;
  const uint32_t tmem = tmem_slot;
  const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
  const uint32_t tS = tmem + lane_off, tDP = tmem + 128 + lane_off, tDQ = tmem + 256 + lane_off;
  constexpr uint32_t idesc_s = ptx::make_idesc_bf16(128, 128, false, false);
  constexpr uint32_t idesc_dq = ptx::make_idesc_bf16(128, HDP, false, true);
  const int cbase = half * 64;
  uint32_t phase = 0;
  int iter = 0;
  for (int kv0 = 0; kv0 < p.S; kv0 += 128, ++iter) {
    load_tile<HDP, T_K>(sK, p, b, h, kv0, r, half);
    load_tile<HDP, T_V>(sV, p, b, h, kv0, r, half);
    float my_bias = 0.f;
    if (half == 0) { my_bias = key_bias(p, b, kv0 + r); sBias[r] = my_bias; }
    cp_async_wait_all();
    ptx::fence_proxy_async_smem();
    ptx::tc_fence_before();
    const bool masked_tile = __syncthreads_or(my_bias != 0.f);
    if (tid == 0) {
      ptx::tc_fence_after();
      const uint32_t qa = ptx::smem_u32(sQ), da = ptx::smem_u32(sDO), ka = ptx::smem_u32(sK), va = ptx::smem_u32(sV);
#pragma unroll
      for (int ks = 0; ks < HDP / 16; ++ks) {
        ptx::umma_bf16(tmem, ptx::make_smem_desc_noswz(qa + ks * 4096, 2048, 128), ptx::make_smem_desc_noswz(ka + ks * 4096, 2048, 128), idesc_s, ks > 0);
        ptx::umma_bf16(tmem + 128, ptx::make_smem_desc_noswz(da + ks * 4096, 2048, 128), ptx::make_smem_desc_noswz(va + ks * 4096, 2048, 128), idesc_s, ks > 0);
      }
      ptx::umma_commit(&bar1);
    }
    ptx::mbar_wait(&bar1, phase);
    ptx::tc_fence_after();
    uint8_t* dsrow = sDS + r * 16;
#pragma unroll
    for (int c = 0; c < 64; c += 32) {
      uint32_t vs[32], vd[32];
      ptx::tmem_ld32(tS + cbase + c, vs);
      ptx::tmem_ld32(tDP + cbase + c, vd);
      ptx::tmem_ld_wait();
      float ds[32];
      if (masked_tile) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          ds[j] = ex2((__uint_as_float(vs[j]) + sBias[cbase + c + j]) * p.scale_log2 - Lrow) * (__uint_as_float(vd[j]) - Drow);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) ds[j] = ex2(__uint_as_float(vs[j]) * p.scale_log2 - Lrow) * (__uint_as_float(vd[j]) - Drow);
      }
      store_bf16x32(dsrow, cbase + c, ds);
    }
    ptx::fence_proxy_async_smem();
    ptx::tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      ptx::tc_fence_after();
      const uint32_t sa = ptx::smem_u32(sDS), ka = ptx::smem_u32(sK);
#pragma unroll
      for (int ks = 0; ks < 8; ++ks)  // dQ += dS K : A K-major over keys, B = K tile read MN-major (N = hd, K = keys)
        ptx::umma_bf16(tmem + 256, ptx::make_smem_desc_noswz(sa + ks * 4096, 2048, 128), ptx::make_smem_desc_noswz(ka + ks * 256, 128, 2048),
                       idesc_dq, (iter > 0 || ks > 0) ? 1u : 0u);
      ptx::umma_commit(&bar2);
    }
    ptx::mbar_wait(&bar2, phase);  // K / V / dS tiles are free again
    ptx::tc_fence_after();
    phase ^= 1;
  }
  bf16* stage = reinterpret_cast<bf16*>(sDS);
  tmem_half_to_stage<HDP>(tDQ, stage, r, half, p.scale);
  __syncthreads();
  store_tile<HDP, O_DQ>(stage, p, b, h, q0, tid);
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc<512>(tmem);
}

// ---------------------------------------------------------------------------------------------------------
// backward: dK, dV
// ---------------------------------------------------------------------------------------------------------
template <int HDP>
__global__ void __launch_bounds__(NT) attn_bwd_dkv_tc_kernel(const Params p) {
  constexpr int TILE = 128 * HDP * 2;
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* sK = smem;
  uint8_t* sV = sK + TILE;
  uint8_t* sQ = sV + TILE;
  uint8_t* sDO = sQ + TILE;
  uint8_t* sPT = sDO + TILE;                               // P^T  [128 keys][128 queries] bf16, layout L1
  uint8_t* sDST = sPT + 32768;                             // dS^T
  float* sL = reinterpret_cast<float*>(sDST + 32768);      // 128 lse (log2 units; +inf for rows >= S)
  float* sD = sL + 128;
  __shared__ uint64_t bar1, bar2;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, r = tid & 127, half = tid >> 7;
  const int kv0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
  if (tid == 0) { ptx::mbar_init(&bar1, 1); ptx::mbar_init(&bar2, 1); ptx::fence_mbar_init(); }
  if (warp == 0) ptx::tmem_alloc<512>(&tmem_slot);
  load_tile<HDP, T_K>(sK, p, b, h, kv0, r, half);
  load_tile<HDP, T_V>(sV, p, b, h, kv0, r, half);
  const float kbias = key_bias(p, b, kv0 + r);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
  const uint32_t tST = tmem + lane_off, tDPT = tmem + 128 + lane_off, tDK = tmem + 256 + lane_off, tDV = tmem + 256 + HDP + lane_off;
  constexpr uint32_t idesc_s = ptx::make_idesc_bf16(128, 128, false, false);
  constexpr uint32_t idesc_o = ptx::make_idesc_bf16(128, HDP, false, true);
  const float* lse = p.lse + ((int64_t)b * p.H + h) * p.S;
  const float* dsm = p.dsum + ((int64_t)b * p.H + h) * p.S;
  const int cbase = half * 64;
  uint32_t phase = 0;
  int iter = 0;
  for (int q0 = 0; q0 < p.S; q0 += 128, ++iter) {
    load_tile<HDP, T_Q>(sQ, p, b, h, q0, r, half);
    load_tile<HDP, T_DO>(sDO, p, b, h, q0, r, half);
    if (half == 0) {
      const int s = q0 + r;
      sL[r] = s < p.S ? lse[s] * 1.4426950408889634f : INFINITY;
      sD[r] = s < p.S ? dsm[s] : 0.f;
    }
    cp_async_wait_all();
    ptx::fence_proxy_async_smem();
    ptx::tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      ptx::tc_fence_after();
      const uint32_t ka = ptx::smem_u32(sK), va = ptx::smem_u32(sV), qa = ptx::smem_u32(sQ), da = ptx::smem_u32(sDO);
#pragma unroll
      for (int ks = 0; ks < HDP / 16; ++ks) {
        ptx::umma_bf16(tmem, ptx::make_smem_desc_noswz(ka + ks * 4096, 2048, 128), ptx::make_smem_desc_noswz(qa + ks * 4096, 2048, 128), idesc_s, ks > 0);
        ptx::umma_bf16(tmem + 128, ptx::make_smem_desc_noswz(va + ks * 4096, 2048, 128), ptx::make_smem_desc_noswz(da + ks * 4096, 2048, 128), idesc_s, ks > 0);
      }
      ptx::umma_commit(&bar1);
    }
    ptx::mbar_wait(&bar1, phase);
    ptx::tc_fence_after();
    uint8_t* prow = sPT + r * 16;
    uint8_t* dsrow = sDST + r * 16;
#pragma unroll
    for (int c = 0; c < 64; c += 32) {
      uint32_t vs[32], vd[32];
      ptx::tmem_ld32(tST + cbase + c, vs);
      ptx::tmem_ld32(tDPT + cbase + c, vd);
      ptx::tmem_ld_wait();
      float pv[32], ds[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        pv[j] = ex2((__uint_as_float(vs[j]) + kbias) * p.scale_log2 - sL[cbase + c + j]);
        ds[j] = pv[j] * (__uint_as_float(vd[j]) - sD[cbase + c + j]);
      }
      store_bf16x32(prow, cbase + c, pv);
      store_bf16x32(dsrow, cbase + c, ds);
    }
    ptx::fence_proxy_async_smem();
    ptx::tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      ptx::tc_fence_after();
      const uint32_t pa = ptx::smem_u32(sPT), sa = ptx::smem_u32(sDST), qa = ptx::smem_u32(sQ), da = ptx::smem_u32(sDO);
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {  // contraction over the 128 queries; Q / dO tiles read MN-major (N = hd)
        const uint32_t acc = (iter > 0 || ks > 0) ? 1u : 0u;
        ptx::umma_bf16(tmem + 256 + HDP, ptx::make_smem_desc_noswz(pa + ks * 4096, 2048, 128), ptx::make_smem_desc_noswz(da + ks * 256, 128, 2048), idesc_o, acc);
        ptx::umma_bf16(tmem + 256, ptx::make_smem_desc_noswz(sa + ks * 4096, 2048, 128), ptx::make_smem_desc_noswz(qa + ks * 256, 128, 2048), idesc_o, acc);
      }
      ptx::umma_commit(&bar2);
    }
    ptx::mbar_wait(&bar2, phase);
    ptx::tc_fence_after();
    phase ^= 1;
  }
  bf16* stage = reinterpret_cast<bf16*>(sPT);
  tmem_half_to_stage<HDP>(tDK, stage, r, half, p.scale);
  __syncthreads();
  store_tile<HDP, O_DK>(stage, p, b, h, kv0, tid);
  __syncthreads();
  tmem_half_to_stage<HDP>(tDV, stage, r, half, 1.f);
  __syncthreads();
  store_tile<HDP, O_DV>(stage, p, b, h, kv0, tid);
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc<512>(tmem);
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
    attn_fwd_tc_kernel<HDPV><<<grid, NT, smem, stream>>>(p);
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
    const size_t sm_dq = (size_t)4 * 128 * HDPV * 2 + 32768 + 3 * 128 * 4, sm_dkv = (size_t)4 * 128 * HDPV * 2 + 65536 + 2 * 128 * 4;
    cudaError_t e = cudaFuncSetAttribute(attn_bwd_dq_tc_kernel<HDPV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_dq);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_bwd_dkv_tc_kernel<HDPV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_dkv);
    DLB_REQUIRE(e == cudaSuccess, (int)e, "attn_bwd_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attn_bwd_dq_tc_kernel<HDPV><<<grid, NT, sm_dq, stream>>>(p);
    attn_bwd_dkv_tc_kernel<HDPV><<<grid, NT, sm_dkv, stream>>>(p);
  });
  dlb_count_launch(2);
  return dlb_check_launch("attn_bwd_tc");
}
