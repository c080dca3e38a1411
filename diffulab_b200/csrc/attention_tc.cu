// tcgen05 joint attention (forward), sm_100a. One CTA = 128 query rows of one (sample, head); TMEM lane = query row.
//   S = Q K^T   : tcgen05.mma M=128 N=128 K=16 x (HDP/16), A = Q tile, B = K tile (both K-major, non-swizzled smem)
//   softmax     : thread = row (no shuffles): two passes over the 128 S columns with tcgen05.ld, online max/sum
//   O_j = P V   : P (bf16) staged in smem as the K-major A operand; B = V tile read MN-major straight from its
//                 [key][hd] row layout; per-tile result read back from TMEM and accumulated (rescaled) in registers
// Operands are staged by the threads (cp.async, 16-byte chunks) into the canonical 8x8 core-matrix layout, because the
// sequence is a concatenation of up to two segments with a key-padding mask (reference mmdit.py:184-204).
// Head dims that are not a multiple of 16 (DiT-XL/2: 72) are zero-padded in shared memory only.
#include "common.cuh"
#include "ptx.cuh"

namespace attn_tc {
typedef __nv_bfloat16 bf16;

struct Seg {
  const bf16* q; const bf16* k; const bf16* v; bf16* out;
  int64_t ldq, ldk, ldv, ldo;
  int len;
};
struct Params {
  Seg seg[2];
  float* lse;
  const uint8_t* kmask;
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

enum { T_Q = 0, T_K = 1, T_V = 2 };

// 128-row tile of tokens s0.. -> smem; thread r stages row r (its 16-byte chunks are contiguous in global memory).
// K-major core layout (Q, K): chunk (row r, hd-chunk c) at (r/8)*SBO + c*128 + (r%8)*16, SBO = (HDP/8)*128.
// MN-major layout (V as the B operand of P*V): chunk (key r, hd-chunk c) at c*2048 + (r/8)*128 + (r%8)*16.
template <int HDP, int WHICH>
__device__ __forceinline__ void load_tile(uint8_t* sm, const Params& p, int b, int h, int s0, int r) {
  constexpr int CPR = HDP / 8;
  const int s = s0 + r;
  uint8_t* dst = (WHICH == T_V) ? sm + (r >> 3) * 128 + (r & 7) * 16 : sm + (r >> 3) * (CPR * 128) + (r & 7) * 16;
  constexpr int CSTRIDE = (WHICH == T_V) ? 2048 : 128;
  const int nvalid = p.hd >> 3;
  if (s < p.S) {
    const int sg = s < p.seg[0].len ? 0 : 1;
    const Seg& g = p.seg[sg];
    const int64_t row = (int64_t)b * g.len + (sg ? s - p.seg[0].len : s);
    const bf16* src = (WHICH == T_Q ? g.q + row * g.ldq : (WHICH == T_K ? g.k + row * g.ldk : g.v + row * g.ldv)) + (int64_t)h * p.hd;
#pragma unroll
    for (int c = 0; c < CPR; ++c) {
      if (c < nvalid) cp_async16(dst + c * CSTRIDE, src + c * 8);
      else *reinterpret_cast<uint4*>(dst + c * CSTRIDE) = make_uint4(0, 0, 0, 0);
    }
  } else {
#pragma unroll
    for (int c = 0; c < CPR; ++c) *reinterpret_cast<uint4*>(dst + c * CSTRIDE) = make_uint4(0, 0, 0, 0);
  }
}

template <int HDP>
__global__ void __launch_bounds__(128) attn_fwd_tc_kernel(const Params p) {
  constexpr int CPR = HDP / 8;
  constexpr int QK_BYTES = 128 * HDP * 2;  // 20480 for HDP = 80
  constexpr uint32_t SBO_QK = CPR * 128;   // stride between 8-row groups of a K-major [128][HDP] tile
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + QK_BYTES;
  uint8_t* sV = sK + QK_BYTES;
  uint8_t* sP = sV + QK_BYTES;                            // [128 q][128 keys] bf16, K-major core layout, 32 KB
  float* sBias = reinterpret_cast<float*>(sP + 32768);    // 128 additive key biases (0 / -inf)
  __shared__ uint64_t bar_s, bar_o;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int q0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;

  if (tid == 0) {
    ptx::mbar_init(&bar_s, 1);
    ptx::mbar_init(&bar_o, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 0) ptx::tmem_alloc<256>(&tmem_slot);
  load_tile<HDP, T_Q>(sQ, p, b, h, q0, tid);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t tS = tmem, tO = tmem + 128;  // S: 128 fp32 columns, O tile: HDP columns
  const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
  constexpr uint32_t idesc_s = ptx::make_idesc_bf16(128, 128, false, false);
  constexpr uint32_t idesc_o = ptx::make_idesc_bf16(128, HDP, false, true);

  float o[HDP];
#pragma unroll
  for (int i = 0; i < HDP; ++i) o[i] = 0.f;
  float m = -INFINITY, l = 0.f;
  uint32_t phase = 0;

  for (int kv0 = 0; kv0 < p.S; kv0 += 128) {
    load_tile<HDP, T_K>(sK, p, b, h, kv0, tid);
    load_tile<HDP, T_V>(sV, p, b, h, kv0, tid);
    float my_bias = 0.f;
    {
      const int key = kv0 + tid;
      if (key >= p.S) my_bias = -INFINITY;
      else if (p.kmask && key < p.mask_len && p.kmask[(int64_t)b * p.mask_len + key] == 0) my_bias = -INFINITY;
      sBias[tid] = my_bias;
    }
    cp_async_wait_all();
    ptx::fence_proxy_async_smem();
    ptx::tc_fence_before();
    const bool masked_tile = __syncthreads_or(my_bias != 0.f);  // uniform: any masked / out-of-range key in this tile
    if (tid == 0) {
      ptx::tc_fence_after();
      const uint32_t qa = ptx::smem_u32(sQ), ka = ptx::smem_u32(sK);
#pragma unroll
      for (int ks = 0; ks < HDP / 16; ++ks)
        ptx::umma_bf16(tS, ptx::make_smem_desc_noswz(qa + ks * 256, 128, SBO_QK), ptx::make_smem_desc_noswz(ka + ks * 256, 128, SBO_QK),
                       idesc_s, ks > 0);
      ptx::umma_commit(&bar_s);
    }
    ptx::mbar_wait(&bar_s, phase);
    ptx::tc_fence_after();

    // pass 1: row maximum
    float mx = -INFINITY;
#pragma unroll 1
    for (int c = 0; c < 128; c += 32) {
      uint32_t r[32];
      ptx::tmem_ld32(tS + lane_off + c, r);
      ptx::tmem_ld_wait();
      if (masked_tile) {
#pragma unroll
        for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(r[j]) + sBias[c + j]);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(r[j]));
      }
    }
    const float m_new = fmaxf(m, mx);
    const float ms = (m_new == -INFINITY) ? 0.f : m_new * p.scale_log2;
    const float alpha = ex2(m * p.scale_log2 - ms);
    // pass 2: probabilities -> bf16 A operand in smem
    float lsum = 0.f;
    uint8_t* prow = sP + (tid >> 3) * 2048 + (tid & 7) * 16;
#pragma unroll 1
    for (int c = 0; c < 128; c += 32) {
      uint32_t r[32];
      ptx::tmem_ld32(tS + lane_off + c, r);
      ptx::tmem_ld_wait();
      float pv[32];
      if (masked_tile) {
#pragma unroll
        for (int j = 0; j < 32; ++j) pv[j] = ex2((__uint_as_float(r[j]) + sBias[c + j]) * p.scale_log2 - ms);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) pv[j] = ex2(__uint_as_float(r[j]) * p.scale_log2 - ms);
      }
#pragma unroll
      for (int j = 0; j < 32; j += 4) lsum += (pv[j] + pv[j + 1]) + (pv[j + 2] + pv[j + 3]);
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        uint4 v;
        v.x = pack_bf16x2(pv[8 * q4 + 0], pv[8 * q4 + 1]);
        v.y = pack_bf16x2(pv[8 * q4 + 2], pv[8 * q4 + 3]);
        v.z = pack_bf16x2(pv[8 * q4 + 4], pv[8 * q4 + 5]);
        v.w = pack_bf16x2(pv[8 * q4 + 6], pv[8 * q4 + 7]);
        *reinterpret_cast<uint4*>(prow + ((c >> 3) + q4) * 128) = v;
      }
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
        ptx::umma_bf16(tO, ptx::make_smem_desc_noswz(pa + ks * 256, 128, 2048), ptx::make_smem_desc_noswz(va + ks * 256, 128, 2048),
                       idesc_o, ks > 0);
      ptx::umma_commit(&bar_o);
    }
    ptx::mbar_wait(&bar_o, phase);
    ptx::tc_fence_after();
    // o = o * alpha + O_tile
#pragma unroll
    for (int c = 0; c < HDP; c += 16) {
      uint32_t r[16];
      ptx::tmem_ld16(tO + lane_off + c, r);
      ptx::tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) o[c + j] = o[c + j] * alpha + __uint_as_float(r[j]);
    }
    phase ^= 1;
    ptx::tc_fence_before();
    __syncthreads();  // every thread is done with S / O tiles and sK / sV / sP before they are overwritten
  }

  // finalise: normalise, stage rows in smem (row-major [128][HDP]), coalesced 16-byte stores
  const float inv = l > 0.f ? 1.f / l : 0.f;
  const int row = q0 + tid;
  if (p.lse && row < p.S) p.lse[((int64_t)b * p.H + h) * p.S + row] = m * p.scale + logf(l);
  bf16* stage = reinterpret_cast<bf16*>(sP);
#pragma unroll
  for (int c = 0; c < HDP; c += 8) {
    float t8[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) t8[j] = o[c + j] * inv;
    st8(stage + tid * HDP + c, pack8(t8));
  }
  __syncthreads();
  const int cpr = p.hd >> 3;
  for (int idx = tid; idx < 128 * cpr; idx += 128) {
    const int r = idx / cpr, c = idx - r * cpr;
    const int s = q0 + r;
    if (s < p.S) {
      const int sg = s < p.seg[0].len ? 0 : 1;
      const Seg& g = p.seg[sg];
      const int64_t grow = (int64_t)b * g.len + (sg ? s - p.seg[0].len : s);
      *reinterpret_cast<uint4*>(g.out + grow * g.ldo + (int64_t)h * p.hd + c * 8) = *reinterpret_cast<const uint4*>(stage + r * HDP + c * 8);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc<256>(tmem);
}


// ---------------------------------------------------------------------------------------------------------
// backward (tcgen05). Two kernels, both with TMEM lane = row of the tile that owns the output:
//   dq  : CTA = 128 query rows, loops over 128-key tiles:  S = Q K^T, dP = dO V^T, dS = P o (dP - D), dQ += dS K
//   dkv : CTA = 128 key rows,   loops over 128-query tiles: S^T = K Q^T, dP^T = V dO^T, dV += P^T dO, dK += dS^T Q
// Every [128 x HDP] operand tile lives in ONE shared-memory layout ("L1": 16-byte chunk (row r, chunk c) at
// c*2048 + r*16) that is a valid non-swizzled UMMA layout both K-major (LBO 2048, SBO 128) and MN-major (LBO 128,
// SBO 2048), so Q / dO / K serve as row operands of one product and as transposed operands of another without any
// data movement. P^T / dS^T / dS tiles ([128 x 128] bf16) are written by their owning threads in the same layout.
// ---------------------------------------------------------------------------------------------------------
struct BwdSeg {
  const bf16* q; const bf16* k; const bf16* v; const bf16* o; const bf16* dout;
  bf16* dq; bf16* dk; bf16* dv;
  int64_t ldq, ldk, ldv, ldo, lddo, lddq, lddk, lddv;
  int len;
};
struct BwdParams {
  BwdSeg seg[2];
  const float* lse;  // [B,H,S] natural log
  float* dsum;       // [B,H,S] rowsum(dO * O): written by the dq kernel, read by the dkv kernel
  const uint8_t* kmask;
  int mask_len, B, H, S, hd;
  float scale, scale_log2;
};

enum { B_Q = 0, B_K = 1, B_V = 2, B_DO = 3 };

__device__ __forceinline__ const bf16* bwd_row_ptr(const BwdParams& p, int which, int b, int h, int s) {
  const int sg = s < p.seg[0].len ? 0 : 1;
  const BwdSeg& g = p.seg[sg];
  const int64_t row = (int64_t)b * g.len + (sg ? s - p.seg[0].len : s);
  const bf16* base = which == B_Q ? g.q + row * g.ldq : (which == B_K ? g.k + row * g.ldk : (which == B_V ? g.v + row * g.ldv : g.dout + row * g.lddo));
  return base + (int64_t)h * p.hd;
}

// thread r stages row r of a 128-row tile into layout L1 (zero fill for rows >= S and the head-dim padding)
template <int HDP, int WHICH>
__device__ __forceinline__ void bwd_load_tile(uint8_t* sm, const BwdParams& p, int b, int h, int s0, int r) {
  constexpr int CPR = HDP / 8;
  uint8_t* dst = sm + r * 16;
  const int s = s0 + r;
  const int nvalid = p.hd >> 3;
  if (s < p.S) {
    const bf16* src = bwd_row_ptr(p, WHICH, b, h, s);
#pragma unroll
    for (int c = 0; c < CPR; ++c) {
      if (c < nvalid) cp_async16(dst + c * 2048, src + c * 8);
      else *reinterpret_cast<uint4*>(dst + c * 2048) = make_uint4(0, 0, 0, 0);
    }
  } else {
#pragma unroll
    for (int c = 0; c < CPR; ++c) *reinterpret_cast<uint4*>(dst + c * 2048) = make_uint4(0, 0, 0, 0);
  }
}

// stage a finished [128 x HDP] fp32-in-registers tile (one row per thread) as bf16 row-major in smem, then store the
// valid rows / columns with coalesced 16-byte writes
enum { O_DQ = 0, O_DK = 1, O_DV = 2 };
template <int HDP, int WHICH>
__device__ __forceinline__ void bwd_store_tile(bf16* stage, const BwdParams& p, int b, int h, int s0, int tid) {
  const int cpr = p.hd >> 3;
  for (int idx = tid; idx < 128 * cpr; idx += 128) {
    const int r = idx / cpr, c = idx - r * cpr;
    const int s = s0 + r;
    if (s < p.S) {
      const int sg = s < p.seg[0].len ? 0 : 1;
      const BwdSeg& g = p.seg[sg];
      const int64_t row = (int64_t)b * g.len + (sg ? s - p.seg[0].len : s);
      bf16* base = WHICH == O_DQ ? g.dq + row * g.lddq : (WHICH == O_DK ? g.dk + row * g.lddk : g.dv + row * g.lddv);
      *reinterpret_cast<uint4*>(base + (int64_t)h * p.hd + c * 8) = *reinterpret_cast<const uint4*>(stage + r * HDP + c * 8);
    }
  }
}

template <int HDP>
__global__ void __launch_bounds__(128) attn_bwd_dq_tc_kernel(const BwdParams p) {
  constexpr int TILE = 128 * HDP * 2;
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* sQ = smem;
  uint8_t* sDO = sQ + TILE;
  uint8_t* sK = sDO + TILE;
  uint8_t* sV = sK + TILE;
  uint8_t* sDS = sV + TILE;                                // [128 q][128 keys] bf16, layout L1 (16 chunks), 32 KB
  float* sBias = reinterpret_cast<float*>(sDS + 32768);    // 128 key biases
  __shared__ uint64_t bar1, bar2;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int q0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
  if (tid == 0) { ptx::mbar_init(&bar1, 1); ptx::mbar_init(&bar2, 1); ptx::fence_mbar_init(); }
  if (warp == 0) ptx::tmem_alloc<512>(&tmem_slot);
  bwd_load_tile<HDP, B_Q>(sQ, p, b, h, q0, tid);
  // dO row: staged through registers so that D = rowsum(dO * O) comes for free
  float Drow = 0.f;
  {
    const int s = q0 + tid;
    const int nvalid = p.hd >> 3;
    uint8_t* dst = sDO + tid * 16;
    if (s < p.S) {
      const bf16* dsrc = bwd_row_ptr(p, B_DO, b, h, s);
      const int sg = s < p.seg[0].len ? 0 : 1;
      const BwdSeg& g = p.seg[sg];
      const bf16* osrc = g.o + ((int64_t)b * g.len + (sg ? s - p.seg[0].len : s)) * g.ldo + (int64_t)h * p.hd;
#pragma unroll
      for (int c = 0; c < HDP / 8; ++c) {
        if (c < nvalid) {
          const bf16x8 dv = ld8(dsrc + c * 8), ov = ld8(osrc + c * 8);
          float df[8], of[8];
          unpack8(dv, df); unpack8(ov, of);
#pragma unroll
          for (int j = 0; j < 8; ++j) Drow += df[j] * of[j];
          *reinterpret_cast<bf16x8*>(dst + c * 2048) = dv;
        } else {
          *reinterpret_cast<uint4*>(dst + c * 2048) = make_uint4(0, 0, 0, 0);
        }
      }
      p.dsum[((int64_t)b * p.H + h) * p.S + s] = Drow;
    } else {
#pragma unroll
      for (int c = 0; c < HDP / 8; ++c) *reinterpret_cast<uint4*>(dst + c * 2048) = make_uint4(0, 0, 0, 0);
    }
  }
  const int myrow = q0 + tid;
  const float Lrow = myrow < p.S ? p.lse[((int64_t)b * p.H + h) * p.S + myrow] * 1.4426950408889634f : INFINITY;
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t tS = tmem, tDP = tmem + 128, tDQ = tmem + 256;
  const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
  constexpr uint32_t idesc_s = ptx::make_idesc_bf16(128, 128, false, false);
  constexpr uint32_t idesc_dq = ptx::make_idesc_bf16(128, HDP, false, true);
  uint32_t phase = 0;
  int iter = 0;
  for (int kv0 = 0; kv0 < p.S; kv0 += 128, ++iter) {
    bwd_load_tile<HDP, B_K>(sK, p, b, h, kv0, tid);
    bwd_load_tile<HDP, B_V>(sV, p, b, h, kv0, tid);
    {
      const int key = kv0 + tid;
      float bias = 0.f;
      if (key >= p.S) bias = -INFINITY;
      else if (p.kmask && key < p.mask_len && p.kmask[(int64_t)b * p.mask_len + key] == 0) bias = -INFINITY;
      sBias[tid] = bias;
    }
    cp_async_wait_all();
    ptx::fence_proxy_async_smem();
    ptx::tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      ptx::tc_fence_after();
      const uint32_t qa = ptx::smem_u32(sQ), da = ptx::smem_u32(sDO), ka = ptx::smem_u32(sK), va = ptx::smem_u32(sV);
#pragma unroll
      for (int ks = 0; ks < HDP / 16; ++ks) {
        ptx::umma_bf16(tS, ptx::make_smem_desc_noswz(qa + ks * 4096, 2048, 128), ptx::make_smem_desc_noswz(ka + ks * 4096, 2048, 128), idesc_s, ks > 0);
        ptx::umma_bf16(tDP, ptx::make_smem_desc_noswz(da + ks * 4096, 2048, 128), ptx::make_smem_desc_noswz(va + ks * 4096, 2048, 128), idesc_s, ks > 0);
      }
      ptx::umma_commit(&bar1);
    }
    ptx::mbar_wait(&bar1, phase);
    ptx::tc_fence_after();
    uint8_t* dsrow = sDS + tid * 16;
#pragma unroll 1
    for (int c = 0; c < 128; c += 32) {
      uint32_t rs[32], rd[32];
      ptx::tmem_ld32(tS + lane_off + c, rs);
      ptx::tmem_ld32(tDP + lane_off + c, rd);
      ptx::tmem_ld_wait();
      float ds[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float pj = ex2((__uint_as_float(rs[j]) + sBias[c + j]) * p.scale_log2 - Lrow);
        ds[j] = pj * (__uint_as_float(rd[j]) - Drow);
      }
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        uint4 v;
        v.x = pack_bf16x2(ds[8 * q4 + 0], ds[8 * q4 + 1]);
        v.y = pack_bf16x2(ds[8 * q4 + 2], ds[8 * q4 + 3]);
        v.z = pack_bf16x2(ds[8 * q4 + 4], ds[8 * q4 + 5]);
        v.w = pack_bf16x2(ds[8 * q4 + 6], ds[8 * q4 + 7]);
        *reinterpret_cast<uint4*>(dsrow + ((c >> 3) + q4) * 2048) = v;
      }
    }
    ptx::fence_proxy_async_smem();
    ptx::tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      ptx::tc_fence_after();
      const uint32_t sa = ptx::smem_u32(sDS), ka = ptx::smem_u32(sK);
#pragma unroll
      for (int ks = 0; ks < 8; ++ks)  // dQ += dS K : A K-major over keys, B = K tile read MN-major (N = hd, K = keys)
        ptx::umma_bf16(tDQ, ptx::make_smem_desc_noswz(sa + ks * 4096, 2048, 128), ptx::make_smem_desc_noswz(ka + ks * 256, 128, 2048),
                       idesc_dq, (iter > 0 || ks > 0) ? 1u : 0u);
      ptx::umma_commit(&bar2);
    }
    ptx::mbar_wait(&bar2, phase);  // K / V / dS tiles are free again
    ptx::tc_fence_after();
    phase ^= 1;
  }
  // dQ * scale -> bf16 -> global
  bf16* stage = reinterpret_cast<bf16*>(sDS);
#pragma unroll
  for (int c = 0; c < HDP; c += 16) {
    uint32_t r[16];
    ptx::tmem_ld16(tDQ + lane_off + c, r);
    ptx::tmem_ld_wait();
    float t[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) t[j] = __uint_as_float(r[j]) * p.scale;
    st8(stage + tid * HDP + c, pack8(t));
    st8(stage + tid * HDP + c + 8, pack8(t + 8));
  }
  __syncthreads();
  bwd_store_tile<HDP, O_DQ>(stage, p, b, h, q0, tid);
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc<512>(tmem);
}

template <int HDP>
__global__ void __launch_bounds__(128) attn_bwd_dkv_tc_kernel(const BwdParams p) {
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
  const int tid = threadIdx.x, warp = tid >> 5;
  const int kv0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
  if (tid == 0) { ptx::mbar_init(&bar1, 1); ptx::mbar_init(&bar2, 1); ptx::fence_mbar_init(); }
  if (warp == 0) ptx::tmem_alloc<512>(&tmem_slot);
  bwd_load_tile<HDP, B_K>(sK, p, b, h, kv0, tid);
  bwd_load_tile<HDP, B_V>(sV, p, b, h, kv0, tid);
  float kbias = 0.f;
  {
    const int key = kv0 + tid;
    if (key >= p.S) kbias = -INFINITY;
    else if (p.kmask && key < p.mask_len && p.kmask[(int64_t)b * p.mask_len + key] == 0) kbias = -INFINITY;
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t tST = tmem, tDPT = tmem + 128, tDK = tmem + 256, tDV = tmem + 256 + HDP;
  const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
  constexpr uint32_t idesc_s = ptx::make_idesc_bf16(128, 128, false, false);
  constexpr uint32_t idesc_o = ptx::make_idesc_bf16(128, HDP, false, true);
  const float* lse = p.lse + ((int64_t)b * p.H + h) * p.S;
  const float* dsm = p.dsum + ((int64_t)b * p.H + h) * p.S;
  uint32_t phase = 0;
  int iter = 0;
  for (int q0 = 0; q0 < p.S; q0 += 128, ++iter) {
    bwd_load_tile<HDP, B_Q>(sQ, p, b, h, q0, tid);
    bwd_load_tile<HDP, B_DO>(sDO, p, b, h, q0, tid);
    {
      const int r = q0 + tid;
      sL[tid] = r < p.S ? lse[r] * 1.4426950408889634f : INFINITY;
      sD[tid] = r < p.S ? dsm[r] : 0.f;
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
        ptx::umma_bf16(tST, ptx::make_smem_desc_noswz(ka + ks * 4096, 2048, 128), ptx::make_smem_desc_noswz(qa + ks * 4096, 2048, 128), idesc_s, ks > 0);
        ptx::umma_bf16(tDPT, ptx::make_smem_desc_noswz(va + ks * 4096, 2048, 128), ptx::make_smem_desc_noswz(da + ks * 4096, 2048, 128), idesc_s, ks > 0);
      }
      ptx::umma_commit(&bar1);
    }
    ptx::mbar_wait(&bar1, phase);
    ptx::tc_fence_after();
    uint8_t* prow = sPT + tid * 16;
    uint8_t* dsrow = sDST + tid * 16;
#pragma unroll 1
    for (int c = 0; c < 128; c += 32) {
      uint32_t rs[32], rd[32];
      ptx::tmem_ld32(tST + lane_off + c, rs);
      ptx::tmem_ld32(tDPT + lane_off + c, rd);
      ptx::tmem_ld_wait();
      float pv[32], ds[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        pv[j] = ex2((__uint_as_float(rs[j]) + kbias) * p.scale_log2 - sL[c + j]);
        ds[j] = pv[j] * (__uint_as_float(rd[j]) - sD[c + j]);
      }
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        uint4 v, w;
        v.x = pack_bf16x2(pv[8 * q4 + 0], pv[8 * q4 + 1]); w.x = pack_bf16x2(ds[8 * q4 + 0], ds[8 * q4 + 1]);
        v.y = pack_bf16x2(pv[8 * q4 + 2], pv[8 * q4 + 3]); w.y = pack_bf16x2(ds[8 * q4 + 2], ds[8 * q4 + 3]);
        v.z = pack_bf16x2(pv[8 * q4 + 4], pv[8 * q4 + 5]); w.z = pack_bf16x2(ds[8 * q4 + 4], ds[8 * q4 + 5]);
        v.w = pack_bf16x2(pv[8 * q4 + 6], pv[8 * q4 + 7]); w.w = pack_bf16x2(ds[8 * q4 + 6], ds[8 * q4 + 7]);
        *reinterpret_cast<uint4*>(prow + ((c >> 3) + q4) * 2048) = v;
        *reinterpret_cast<uint4*>(dsrow + ((c >> 3) + q4) * 2048) = w;
      }
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
        ptx::umma_bf16(tDV, ptx::make_smem_desc_noswz(pa + ks * 4096, 2048, 128), ptx::make_smem_desc_noswz(da + ks * 256, 128, 2048), idesc_o, acc);
        ptx::umma_bf16(tDK, ptx::make_smem_desc_noswz(sa + ks * 4096, 2048, 128), ptx::make_smem_desc_noswz(qa + ks * 256, 128, 2048), idesc_o, acc);
      }
      ptx::umma_commit(&bar2);
    }
    ptx::mbar_wait(&bar2, phase);
    ptx::tc_fence_after();
    phase ^= 1;
  }
  bf16* stage = reinterpret_cast<bf16*>(sPT);
  for (int which = 0; which < 2; ++which) {
    const uint32_t tsrc = which ? tDV : tDK;
    const float mul = which ? 1.f : p.scale;
#pragma unroll
    for (int c = 0; c < HDP; c += 16) {
      uint32_t r[16];
      ptx::tmem_ld16(tsrc + lane_off + c, r);
      ptx::tmem_ld_wait();
      float t[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) t[j] = __uint_as_float(r[j]) * mul;
      st8(stage + tid * HDP + c, pack8(t));
      st8(stage + tid * HDP + c + 8, pack8(t + 8));
    }
    __syncthreads();
    if (which == 0) bwd_store_tile<HDP, O_DK>(stage, p, b, h, kv0, tid);
    else bwd_store_tile<HDP, O_DV>(stage, p, b, h, kv0, tid);
    __syncthreads();
  }
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

// Same contract as dlb_attn_fwd (attention.cu); tcgen05 implementation.
DLB_EXPORT int dlb_attn_fwd_tc(const dlb_attn_seg* segs, int nseg, float* lse, const uint8_t* kmask, int mask_len, int B,
                               int H, int hd, float scale, cudaStream_t stream) {
  using namespace attn_tc;
  DLB_REQUIRE(nseg == 1 || nseg == 2, DLB_ERR_SHAPE, "attn_fwd_tc: 1 or 2 segments supported (got %d)", nseg);
  DLB_REQUIRE(B > 0 && H > 0 && hd > 0 && hd % 8 == 0 && hd <= 128, DLB_ERR_SHAPE, "attn_fwd_tc: B=%d H=%d hd=%d", B, H, hd);
  Params p{};
  int S = 0;
  for (int i = 0; i < nseg; ++i) {
    const dlb_attn_seg& s = segs[i];
    DLB_REQUIRE(s.len >= 0 && s.ldq % 8 == 0 && s.ldk % 8 == 0 && s.ldv % 8 == 0 && s.ldo % 8 == 0, DLB_ERR_ALIGN,
                "attn_fwd_tc: strides must be multiples of 8");
    p.seg[i] = Seg{(const bf16*)s.q, (const bf16*)s.k, (const bf16*)s.v, (bf16*)s.o, s.ldq, s.ldk, s.ldv, s.ldo, s.len};
    S += s.len;
  }
  DLB_REQUIRE(S > 0 && mask_len >= 0 && mask_len <= S && (kmask != nullptr || mask_len == 0), DLB_ERR_SHAPE, "attn_fwd_tc: bad sequence / mask");
  p.lse = lse; p.kmask = kmask; p.mask_len = mask_len; p.B = B; p.H = H; p.S = S; p.hd = hd;
  p.scale = scale; p.scale_log2 = scale * 1.4426950408889634f;
  dim3 grid((S + 127) / 128, H, B);
  const int hdp = (hd + 15) / 16 * 16;
#define LAUNCH_FWD(HDPV)                                                                                          \
  {                                                                                                               \
    const size_t smem = (size_t)3 * 128 * HDPV * 2 + 32768 + 512;                                                 \
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_tc_kernel<HDPV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    DLB_REQUIRE(e == cudaSuccess, (int)e, "attn_fwd_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));        \
    attn_fwd_tc_kernel<HDPV><<<grid, 128, smem, stream>>>(p);                                                     \
  }
  switch (hdp) {
    case 16: case 32: case 48: case 64: LAUNCH_FWD(64); break;
    case 80: LAUNCH_FWD(80); break;
    case 96: LAUNCH_FWD(96); break;
    default: LAUNCH_FWD(128); break;
  }
#undef LAUNCH_FWD
  dlb_count_launch();
  return dlb_check_launch("attn_fwd_tc");
}

// Same contract as dlb_attn_bwd (attention.cu); tcgen05 implementation. dsum is written by the dq pass and read by
// the dkv pass (stream order). Head dims above 128 are not supported.
DLB_EXPORT int dlb_attn_bwd_tc(const dlb_attn_seg* segs, int nseg, const float* lse, float* dsum, const uint8_t* kmask,
                               int mask_len, int B, int H, int hd, float scale, cudaStream_t stream) {
  using namespace attn_tc;
  DLB_REQUIRE(nseg == 1 || nseg == 2, DLB_ERR_SHAPE, "attn_bwd_tc: 1 or 2 segments supported (got %d)", nseg);
  DLB_REQUIRE(B > 0 && H > 0 && hd > 0 && hd % 8 == 0 && hd <= 128, DLB_ERR_SHAPE, "attn_bwd_tc: B=%d H=%d hd=%d", B, H, hd);
  DLB_REQUIRE(lse != nullptr && dsum != nullptr, DLB_ERR_SHAPE, "attn_bwd_tc: lse and dsum buffers are required");
  BwdParams p{};
  int S = 0;
  for (int i = 0; i < nseg; ++i) {
    const dlb_attn_seg& s = segs[i];
    DLB_REQUIRE(s.len >= 0 && s.ldq % 8 == 0 && s.ldk % 8 == 0 && s.ldv % 8 == 0 && s.ldo % 8 == 0 && s.lddo % 8 == 0 &&
                    s.lddq % 8 == 0 && s.lddk % 8 == 0 && s.lddv % 8 == 0,
                DLB_ERR_ALIGN, "attn_bwd_tc: strides must be multiples of 8");
    p.seg[i] = BwdSeg{(const bf16*)s.q, (const bf16*)s.k, (const bf16*)s.v, (const bf16*)s.o, (const bf16*)s.dout,
                      (bf16*)s.dq, (bf16*)s.dk, (bf16*)s.dv, s.ldq, s.ldk, s.ldv, s.ldo, s.lddo, s.lddq, s.lddk, s.lddv, s.len};
    S += s.len;
  }
  DLB_REQUIRE(S > 0 && mask_len >= 0 && mask_len <= S && (kmask != nullptr || mask_len == 0), DLB_ERR_SHAPE, "attn_bwd_tc: bad sequence / mask");
  p.lse = lse; p.dsum = dsum; p.kmask = kmask; p.mask_len = mask_len; p.B = B; p.H = H; p.S = S; p.hd = hd;
  p.scale = scale; p.scale_log2 = scale * 1.4426950408889634f;
  dim3 grid((S + 127) / 128, H, B);
  const int hdp = (hd + 15) / 16 * 16;
#define LAUNCH_BWD(HDPV)                                                                                               \
  {                                                                                                                    \
    const size_t sm_dq = (size_t)4 * 128 * HDPV * 2 + 32768 + 512, sm_dkv = (size_t)4 * 128 * HDPV * 2 + 65536 + 1024; \
    cudaError_t e = cudaFuncSetAttribute(attn_bwd_dq_tc_kernel<HDPV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_dq); \
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_bwd_dkv_tc_kernel<HDPV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_dkv); \
    DLB_REQUIRE(e == cudaSuccess, (int)e, "attn_bwd_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));             \
    attn_bwd_dq_tc_kernel<HDPV><<<grid, 128, sm_dq, stream>>>(p);                                                      \
    attn_bwd_dkv_tc_kernel<HDPV><<<grid, 128, sm_dkv, stream>>>(p);                                                    \
  }
  switch (hdp) {
    case 16: case 32: case 48: case 64: LAUNCH_BWD(64); break;
    case 80: LAUNCH_BWD(80); break;
    case 96: LAUNCH_BWD(96); break;
    default: LAUNCH_BWD(128); break;
  }
#undef LAUNCH_BWD
  dlb_count_launch(2);
  return dlb_check_launch("attn_bwd_tc");
}
