// Joint (text + image) flash-style attention, forward and backward, for short sequences (S <= a few thousand)
// and head dims that are multiples of 8 up to 128 (DiT-XL/2: 72, shipped configs: 64).
//
// Replaces F.scaled_dot_product_attention as called by DiTAttention / MMDiTAttention
// (reference mmdit.py:92-98, 184-204): softmax(q k^T * hd^-1/2 + key_padding_mask) v, bf16 operands, fp32
// softmax. Q/K arrive already RMS-normalised and rotated (qknorm_rope.cu); V is read in place from the packed
// qkv projection. The sequence is the concatenation of up to two segments living in different buffers (text rows
// first, then image rows, mmdit.py:185-187), so the `torch.cat`s of the reference are never materialised.
//
// Round-1 implementation: warp-level mma.sync.m16n8k16 (bf16 -> fp32) tiles, 64 query rows x 64 keys per step,
// online softmax in registers. Attention is 2.7% of the step's FLOPs; the tcgen05 version
// is the next optimisation step.
#include "common.cuh"

namespace {
typedef __nv_bfloat16 bf16;

struct Seg {
  const bf16* q; const bf16* k; const bf16* v; const bf16* o; const bf16* dout;
  bf16* out; bf16* dq; bf16* dk; bf16* dv;
  int64_t ldq, ldk, ldv, ldo, lddo, lddq, lddk, lddv;
  int len;
};

struct AttnParams {
  Seg seg[2];
  float* lse;            // [B, H, S] natural-log LSE of the scaled scores
  float* dsum;           // [B, H, S] rowsum(dO * O)
  const uint8_t* kmask;  // [B, mask_len] 1 = attend; keys >= mask_len always attend
  int mask_len;
  int B, H, S, hd;
  float scale, scale_log2;
};

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(src));
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory"); }
__device__ __forceinline__ void ldsm4(uint32_t* r, const bf16* p) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ldsm4t(uint32_t* r, const bf16* p) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void mma16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v + __shfl_xor_sync(0xffffffffu, v, 2);
}

enum { T_Q = 0, T_K = 1, T_V = 2, T_DO = 3 };

// Load a 64-row tile (tokens s0..s0+63 of sample b, head h) into smem [64][HDP+8]; zero-fills rows >= S and the
// head-dim padding (zeros are required on both operands of every contraction over the head dim).
template <int HDP, int WHICH>
__device__ __forceinline__ void load_tile(bf16* sm, const AttnParams& p, int b, int h, int s0, int tid) {
  constexpr int LDS = HDP + 8, CPR = HDP / 8;
  for (int idx = tid; idx < 64 * CPR; idx += 128) {
    const int r = idx / CPR, c = idx - r * CPR;
    const int s = s0 + r;
    bf16* dst = sm + r * LDS + c * 8;
    if (s < p.S && c * 8 < p.hd) {
      const int sg = s < p.seg[0].len ? 0 : 1;
      const Seg& g = p.seg[sg];
      const int64_t row = (int64_t)b * g.len + (sg ? s - p.seg[0].len : s);
      const bf16* base; int64_t ld;
      if (WHICH == T_Q) { base = g.q; ld = g.ldq; }
      else if (WHICH == T_K) { base = g.k; ld = g.ldk; }
      else if (WHICH == T_V) { base = g.v; ld = g.ldv; }
      else { base = g.dout; ld = g.lddo; }
      cp_async16(dst, base + row * ld + (int64_t)h * p.hd + c * 8);
    } else {
      *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
    }
  }
}

// Store a warp's 16 x hd tile staged in smem rows [r0, r0+16) to the per-segment destination (coalesced 16 B).
enum { D_OUT = 0, D_DQ = 1, D_DK = 2, D_DV = 3 };
template <int HDP, int WHICH>
__device__ __forceinline__ void store_rows16(const bf16* sm, const AttnParams& p, int b, int h, int s0, int r0, int lane) {
  constexpr int LDS = HDP + 8;
  const int cpr = p.hd >> 3;
  for (int idx = lane; idx < 16 * cpr; idx += 32) {
    const int r = idx / cpr, c = idx - r * cpr;
    const int s = s0 + r0 + r;
    if (s < p.S) {
      const int sg = s < p.seg[0].len ? 0 : 1;
      const Seg& g = p.seg[sg];
      const int64_t row = (int64_t)b * g.len + (sg ? s - p.seg[0].len : s);
      bf16* base; int64_t ld;
      if (WHICH == D_OUT) { base = g.out; ld = g.ldo; }
      else if (WHICH == D_DQ) { base = g.dq; ld = g.lddq; }
      else if (WHICH == D_DK) { base = g.dk; ld = g.lddk; }
      else { base = g.dv; ld = g.lddv; }
      *reinterpret_cast<uint4*>(base + row * ld + (int64_t)h * p.hd + c * 8) =
          *reinterpret_cast<const uint4*>(sm + (r0 + r) * LDS + c * 8);
    }
  }
}

__device__ __forceinline__ float key_bias(const AttnParams& p, int b, int key) {
  if (key >= p.S) return -INFINITY;
  if (p.kmask && key < p.mask_len && p.kmask[(int64_t)b * p.mask_len + key] == 0) return -INFINITY;
  return 0.f;
}

// ---------------------------------------------------------------------------------------------------------
// forward: grid (ceil(S/64), H, B), 4 warps, each warp 16 query rows
// ---------------------------------------------------------------------------------------------------------
template <int HDP>
__global__ void __launch_bounds__(128) attn_fwd_kernel(const AttnParams p) {
  constexpr int LDS = HDP + 8, KS = HDP / 16, NT = HDP / 8;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  bf16* sQ = reinterpret_cast<bf16*>(smem_raw);
  bf16* sK = sQ + 64 * LDS;
  bf16* sV = sK + 64 * LDS;
  float* sB = reinterpret_cast<float*>(sV + 64 * LDS);  // 64 additive key biases (0 / -inf)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int q0 = blockIdx.x * 64, h = blockIdx.y, b = blockIdx.z;

  load_tile<HDP, T_Q>(sQ, p, b, h, q0, tid);
  cp_async_wait_all();
  __syncthreads();
  uint32_t qf[KS][4];
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) ldsm4(qf[ks], sQ + (warp * 16 + (lane & 15)) * LDS + ks * 16 + (lane >> 4) * 8);

  float o[NT][4];
#pragma unroll
  for (int j = 0; j < NT; ++j) { o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f; }
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;

  for (int kv0 = 0; kv0 < p.S; kv0 += 64) {
    __syncthreads();
    load_tile<HDP, T_K>(sK, p, b, h, kv0, tid);
    load_tile<HDP, T_V>(sV, p, b, h, kv0, tid);
    if (tid < 64) sB[tid] = key_bias(p, b, kv0 + tid);
    cp_async_wait_all();
    __syncthreads();

    float s[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) { s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f; }
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        uint32_t kf[4];
        ldsm4(kf, sK + (np * 16 + (lane & 7) + (lane >> 4) * 8) * LDS + ks * 16 + ((lane >> 3) & 1) * 8);
        mma16816(s[2 * np], qf[ks], kf[0], kf[1]);
        mma16816(s[2 * np + 1], qf[ks], kf[2], kf[3]);
      }
    }
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float b0 = sB[8 * j + 2 * t], b1 = sB[8 * j + 2 * t + 1];
      s[j][0] += b0; s[j][1] += b1; s[j][2] += b0; s[j][3] += b1;
      mx0 = fmaxf(mx0, fmaxf(s[j][0], s[j][1]));
      mx1 = fmaxf(mx1, fmaxf(s[j][2], s[j][3]));
    }
    mx0 = quad_max(mx0); mx1 = quad_max(mx1);
    const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
    const float ms0 = (mn0 == -INFINITY) ? 0.f : mn0 * p.scale_log2;
    const float ms1 = (mn1 == -INFINITY) ? 0.f : mn1 * p.scale_log2;
    const float al0 = ex2(m0 * p.scale_log2 - ms0), al1 = ex2(m1 * p.scale_log2 - ms1);
    float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s[j][0] = ex2(s[j][0] * p.scale_log2 - ms0);
      s[j][1] = ex2(s[j][1] * p.scale_log2 - ms0);
      s[j][2] = ex2(s[j][2] * p.scale_log2 - ms1);
      s[j][3] = ex2(s[j][3] * p.scale_log2 - ms1);
      rs0 += s[j][0] + s[j][1];
      rs1 += s[j][2] + s[j][3];
    }
    l0 = l0 * al0 + rs0; l1 = l1 * al1 + rs1;
    m0 = mn0; m1 = mn1;
#pragma unroll
    for (int j = 0; j < NT; ++j) { o[j][0] *= al0; o[j][1] *= al0; o[j][2] *= al1; o[j][3] *= al1; }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t pa[4];
      pa[0] = pack_bf16x2(s[2 * kk][0], s[2 * kk][1]);
      pa[1] = pack_bf16x2(s[2 * kk][2], s[2 * kk][3]);
      pa[2] = pack_bf16x2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
      pa[3] = pack_bf16x2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
      for (int np = 0; np < NT / 2; ++np) {
        uint32_t vf[4];
        ldsm4t(vf, sV + (kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * LDS + np * 16 + (lane >> 4) * 8);
        mma16816(o[2 * np], pa, vf[0], vf[1]);
        mma16816(o[2 * np + 1], pa, vf[2], vf[3]);
      }
    }
  }
  l0 = quad_sum(l0); l1 = quad_sum(l1);
  const float i0 = l0 > 0.f ? 1.f / l0 : 0.f, i1 = l1 > 0.f ? 1.f / l1 : 0.f;
  if (t == 0 && p.lse) {
    const int r0 = q0 + warp * 16 + g, r1 = r0 + 8;
    float* lse = p.lse + ((int64_t)b * p.H + h) * p.S;
    if (r0 < p.S) lse[r0] = m0 * p.scale + logf(l0);
    if (r1 < p.S) lse[r1] = m1 * p.scale + logf(l1);
  }
  // stage through this warp's own Q rows, then coalesced stores
#pragma unroll
  for (int j = 0; j < NT; ++j) {
    *reinterpret_cast<uint32_t*>(sQ + (warp * 16 + g) * LDS + 8 * j + 2 * t) = pack_bf16x2(o[j][0] * i0, o[j][1] * i0);
    *reinterpret_cast<uint32_t*>(sQ + (warp * 16 + g + 8) * LDS + 8 * j + 2 * t) = pack_bf16x2(o[j][2] * i1, o[j][3] * i1);
  }
  __syncwarp();
  store_rows16<HDP, D_OUT>(sQ, p, b, h, q0, warp * 16, lane);
}

// ---------------------------------------------------------------------------------------------------------
// backward pre-pass: dsum[b,h,s] = sum_c dO * O   (one warp per (token, head))
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) attn_bwd_prep_kernel(const AttnParams p) {
  const int lane = threadIdx.x & 31;
  const int64_t wid = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t total = (int64_t)p.B * p.S * p.H;
  if (wid >= total) return;
  const int h = (int)(wid % p.H);
  const int64_t bs = wid / p.H;
  const int s = (int)(bs % p.S), b = (int)(bs / p.S);
  const int sg = s < p.seg[0].len ? 0 : 1;
  const Seg& g = p.seg[sg];
  const int64_t row = (int64_t)b * g.len + (sg ? s - p.seg[0].len : s);
  float acc = 0.f;
  for (int c = lane * 8; c < p.hd; c += 256) {
    float a[8], d[8];
    unpack8(ld8(g.o + row * g.ldo + (int64_t)h * p.hd + c), a);
    unpack8(ld8(g.dout + row * g.lddo + (int64_t)h * p.hd + c), d);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc += a[j] * d[j];
  }
  acc = warp_sum(acc);
  if (lane == 0) p.dsum[((int64_t)b * p.H + h) * p.S + s] = acc;
}

// ---------------------------------------------------------------------------------------------------------
// backward dQ: grid (ceil(S/64), H, B); each warp 16 query rows, loops over key blocks
// ---------------------------------------------------------------------------------------------------------
template <int HDP>
__global__ void __launch_bounds__(128) attn_bwd_dq_kernel(const AttnParams p) {
  constexpr int LDS = HDP + 8, KS = HDP / 16, NT = HDP / 8;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  bf16* sQ = reinterpret_cast<bf16*>(smem_raw);   // Q tile, later dQ staging
  bf16* sO = sQ + 64 * LDS;                       // dO tile
  bf16* sK = sO + 64 * LDS;
  bf16* sV = sK + 64 * LDS;
  float* sB = reinterpret_cast<float*>(sV + 64 * LDS);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int q0 = blockIdx.x * 64, h = blockIdx.y, b = blockIdx.z;

  load_tile<HDP, T_Q>(sQ, p, b, h, q0, tid);
  load_tile<HDP, T_DO>(sO, p, b, h, q0, tid);
  cp_async_wait_all();
  __syncthreads();
  uint32_t qf[KS][4], df[KS][4];
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) {
    ldsm4(qf[ks], sQ + (warp * 16 + (lane & 15)) * LDS + ks * 16 + (lane >> 4) * 8);
    ldsm4(df[ks], sO + (warp * 16 + (lane & 15)) * LDS + ks * 16 + (lane >> 4) * 8);
  }
  const int r0 = q0 + warp * 16 + g, r1 = r0 + 8;
  const float* lse = p.lse + ((int64_t)b * p.H + h) * p.S;
  const float* dsm = p.dsum + ((int64_t)b * p.H + h) * p.S;
  const float L0 = r0 < p.S ? lse[r0] * 1.4426950408889634f : INFINITY;
  const float L1 = r1 < p.S ? lse[r1] * 1.4426950408889634f : INFINITY;
  const float D0 = r0 < p.S ? dsm[r0] : 0.f, D1 = r1 < p.S ? dsm[r1] : 0.f;

  float dq[NT][4];
#pragma unroll
  for (int j = 0; j < NT; ++j) { dq[j][0] = dq[j][1] = dq[j][2] = dq[j][3] = 0.f; }

  for (int kv0 = 0; kv0 < p.S; kv0 += 64) {
    __syncthreads();
    load_tile<HDP, T_K>(sK, p, b, h, kv0, tid);
    load_tile<HDP, T_V>(sV, p, b, h, kv0, tid);
    if (tid < 64) sB[tid] = key_bias(p, b, kv0 + tid);
    cp_async_wait_all();
    __syncthreads();
    float s[8][4], dp[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
      dp[j][0] = dp[j][1] = dp[j][2] = dp[j][3] = 0.f;
    }
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        uint32_t kf[4], vf[4];
        const int off = (np * 16 + (lane & 7) + (lane >> 4) * 8) * LDS + ks * 16 + ((lane >> 3) & 1) * 8;
        ldsm4(kf, sK + off);
        ldsm4(vf, sV + off);
        mma16816(s[2 * np], qf[ks], kf[0], kf[1]);
        mma16816(s[2 * np + 1], qf[ks], kf[2], kf[3]);
        mma16816(dp[2 * np], df[ks], vf[0], vf[1]);
        mma16816(dp[2 * np + 1], df[ks], vf[2], vf[3]);
      }
    }
    // P = exp(S*scale - lse); dS = P * (dP - D)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float b0 = sB[8 * j + 2 * t], b1 = sB[8 * j + 2 * t + 1];
      const float p00 = ex2((s[j][0] + b0) * p.scale_log2 - L0), p01 = ex2((s[j][1] + b1) * p.scale_log2 - L0);
      const float p10 = ex2((s[j][2] + b0) * p.scale_log2 - L1), p11 = ex2((s[j][3] + b1) * p.scale_log2 - L1);
      s[j][0] = p00 * (dp[j][0] - D0); s[j][1] = p01 * (dp[j][1] - D0);
      s[j][2] = p10 * (dp[j][2] - D1); s[j][3] = p11 * (dp[j][3] - D1);
    }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t pa[4];
      pa[0] = pack_bf16x2(s[2 * kk][0], s[2 * kk][1]);
      pa[1] = pack_bf16x2(s[2 * kk][2], s[2 * kk][3]);
      pa[2] = pack_bf16x2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
      pa[3] = pack_bf16x2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
      for (int np = 0; np < NT / 2; ++np) {
        uint32_t kf[4];
        ldsm4t(kf, sK + (kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * LDS + np * 16 + (lane >> 4) * 8);
        mma16816(dq[2 * np], pa, kf[0], kf[1]);
        mma16816(dq[2 * np + 1], pa, kf[2], kf[3]);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < NT; ++j) {
    *reinterpret_cast<uint32_t*>(sQ + (warp * 16 + g) * LDS + 8 * j + 2 * t) =
        pack_bf16x2(dq[j][0] * p.scale, dq[j][1] * p.scale);
    *reinterpret_cast<uint32_t*>(sQ + (warp * 16 + g + 8) * LDS + 8 * j + 2 * t) =
        pack_bf16x2(dq[j][2] * p.scale, dq[j][3] * p.scale);
  }
  __syncwarp();
  store_rows16<HDP, D_DQ>(sQ, p, b, h, q0, warp * 16, lane);
}

// ---------------------------------------------------------------------------------------------------------
// backward dK/dV: grid (ceil(S/64), H, B) over key blocks; each warp 16 keys, loops over query blocks.
// Works on transposed tiles: S^T = K Q^T, dP^T = V dO^T, dV += P^T dO, dK += dS^T Q.
// ---------------------------------------------------------------------------------------------------------
template <int HDP>
__global__ void __launch_bounds__(128) attn_bwd_dkv_kernel(const AttnParams p) {
  constexpr int LDS = HDP + 8, KS = HDP / 16, NT = HDP / 8;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  bf16* sK = reinterpret_cast<bf16*>(smem_raw);   // K tile, later dK staging
  bf16* sV = sK + 64 * LDS;                       // V tile, later dV staging
  bf16* sQ = sV + 64 * LDS;
  bf16* sO = sQ + 64 * LDS;                       // dO tile
  float* sL = reinterpret_cast<float*>(sO + 64 * LDS);  // 64 lse (log2 units, +inf for rows >= S)
  float* sD = sL + 64;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int kv0 = blockIdx.x * 64, h = blockIdx.y, b = blockIdx.z;

  load_tile<HDP, T_K>(sK, p, b, h, kv0, tid);
  load_tile<HDP, T_V>(sV, p, b, h, kv0, tid);
  cp_async_wait_all();
  __syncthreads();
  uint32_t kf[KS][4], vf[KS][4];
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) {
    ldsm4(kf[ks], sK + (warp * 16 + (lane & 15)) * LDS + ks * 16 + (lane >> 4) * 8);
    ldsm4(vf[ks], sV + (warp * 16 + (lane & 15)) * LDS + ks * 16 + (lane >> 4) * 8);
  }
  const float kb0 = key_bias(p, b, kv0 + warp * 16 + g), kb1 = key_bias(p, b, kv0 + warp * 16 + g + 8);
  const float* lse = p.lse + ((int64_t)b * p.H + h) * p.S;
  const float* dsm = p.dsum + ((int64_t)b * p.H + h) * p.S;

  float dk[NT][4], dv[NT][4];
#pragma unroll
  for (int j = 0; j < NT; ++j) {
    dk[j][0] = dk[j][1] = dk[j][2] = dk[j][3] = 0.f;
    dv[j][0] = dv[j][1] = dv[j][2] = dv[j][3] = 0.f;
  }

  for (int q0 = 0; q0 < p.S; q0 += 64) {
    __syncthreads();
    load_tile<HDP, T_Q>(sQ, p, b, h, q0, tid);
    load_tile<HDP, T_DO>(sO, p, b, h, q0, tid);
    if (tid < 64) {
      const int r = q0 + tid;
      sL[tid] = r < p.S ? lse[r] * 1.4426950408889634f : INFINITY;
      sD[tid] = r < p.S ? dsm[r] : 0.f;
    }
    cp_async_wait_all();
    __syncthreads();
    float s[8][4], dp[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
      dp[j][0] = dp[j][1] = dp[j][2] = dp[j][3] = 0.f;
    }
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        uint32_t qf[4], of[4];
        const int off = (np * 16 + (lane & 7) + (lane >> 4) * 8) * LDS + ks * 16 + ((lane >> 3) & 1) * 8;
        ldsm4(qf, sQ + off);
        ldsm4(of, sO + off);
        mma16816(s[2 * np], kf[ks], qf[0], qf[1]);        // S^T: rows = keys, cols = queries
        mma16816(s[2 * np + 1], kf[ks], qf[2], qf[3]);
        mma16816(dp[2 * np], vf[ks], of[0], of[1]);       // dP^T
        mma16816(dp[2 * np + 1], vf[ks], of[2], of[3]);
      }
    }
    uint32_t pT[4][4], dsT[4][4];  // A fragments over the query (k) dimension
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c0 = 8 * j + 2 * t;
      const float l0 = sL[c0], l1 = sL[c0 + 1], d0 = sD[c0], d1 = sD[c0 + 1];
      const float p00 = ex2((s[j][0] + kb0) * p.scale_log2 - l0), p01 = ex2((s[j][1] + kb0) * p.scale_log2 - l1);
      const float p10 = ex2((s[j][2] + kb1) * p.scale_log2 - l0), p11 = ex2((s[j][3] + kb1) * p.scale_log2 - l1);
      const int kk = j >> 1, hi = (j & 1) * 2;
      pT[kk][hi] = pack_bf16x2(p00, p01);
      pT[kk][hi + 1] = pack_bf16x2(p10, p11);
      dsT[kk][hi] = pack_bf16x2(p00 * (dp[j][0] - d0), p01 * (dp[j][1] - d1));
      dsT[kk][hi + 1] = pack_bf16x2(p10 * (dp[j][2] - d0), p11 * (dp[j][3] - d1));
    }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
      for (int np = 0; np < NT / 2; ++np) {
        uint32_t of[4], qf[4];
        const int off = (kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * LDS + np * 16 + (lane >> 4) * 8;
        ldsm4t(of, sO + off);
        ldsm4t(qf, sQ + off);
        mma16816(dv[2 * np], pT[kk], of[0], of[1]);
        mma16816(dv[2 * np + 1], pT[kk], of[2], of[3]);
        mma16816(dk[2 * np], dsT[kk], qf[0], qf[1]);
        mma16816(dk[2 * np + 1], dsT[kk], qf[2], qf[3]);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < NT; ++j) {
    *reinterpret_cast<uint32_t*>(sK + (warp * 16 + g) * LDS + 8 * j + 2 * t) = pack_bf16x2(dk[j][0] * p.scale, dk[j][1] * p.scale);
    *reinterpret_cast<uint32_t*>(sK + (warp * 16 + g + 8) * LDS + 8 * j + 2 * t) = pack_bf16x2(dk[j][2] * p.scale, dk[j][3] * p.scale);
    *reinterpret_cast<uint32_t*>(sV + (warp * 16 + g) * LDS + 8 * j + 2 * t) = pack_bf16x2(dv[j][0], dv[j][1]);
    *reinterpret_cast<uint32_t*>(sV + (warp * 16 + g + 8) * LDS + 8 * j + 2 * t) = pack_bf16x2(dv[j][2], dv[j][3]);
  }
  __syncwarp();
  store_rows16<HDP, D_DK>(sK, p, b, h, kv0, warp * 16, lane);
  store_rows16<HDP, D_DV>(sV, p, b, h, kv0, warp * 16, lane);
}

int hdp_for(int hd) { return hd <= 64 ? 64 : (hd <= 80 ? 80 : (hd <= 96 ? 96 : 128)); }

template <typename K>
int set_smem(K kern, size_t bytes) {
  if (bytes > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) { dlb_set_error("attention: cudaFuncSetAttribute(%zu): %s", bytes, cudaGetErrorString(e)); return (int)e; }
  }
  return 0;
}

#define HDP_SWITCH(hd, ...)                                        \
  switch (hdp_for(hd)) {                                           \
    case 64: { constexpr int HDP = 64; __VA_ARGS__; break; }       \
    case 80: { constexpr int HDP = 80; __VA_ARGS__; break; }       \
    case 96: { constexpr int HDP = 96; __VA_ARGS__; break; }       \
    default: { constexpr int HDP = 128; __VA_ARGS__; break; }      \
  }

}  // namespace

// Plain-C description of one sequence segment (see include/diffulab_b200.h: dlb_attn_seg).
struct dlb_attn_seg {
  const void* q; const void* k; const void* v;      // forward inputs  [B*len, ...] with row strides ldq/ldk/ldv
  void* o;                                          // forward output / backward input
  const void* dout;                                 // backward: grad of o
  void* dq; void* dk; void* dv;                     // backward outputs
  int64_t ldq, ldk, ldv, ldo, lddo, lddq, lddk, lddv;
  int32_t len;
};

static int fill_params(AttnParams& p, const dlb_attn_seg* segs, int nseg, float* lse, float* dsum, const uint8_t* kmask,
                       int mask_len, int B, int H, int hd, float scale, bool bwd) {
  DLB_REQUIRE(nseg == 1 || nseg == 2, DLB_ERR_SHAPE, "attention: 1 or 2 segments supported (got %d)", nseg);
  DLB_REQUIRE(B > 0 && H > 0 && hd > 0 && hd % 8 == 0 && hd <= 128, DLB_ERR_SHAPE, "attention: B=%d H=%d hd=%d", B, H, hd);
  int S = 0;
  for (int i = 0; i < 2; ++i) {
    Seg& g = p.seg[i];
    if (i < nseg) {
      const dlb_attn_seg& s = segs[i];
      DLB_REQUIRE(s.len >= 0, DLB_ERR_SHAPE, "attention: negative segment length");
      DLB_REQUIRE(s.ldq % 8 == 0 && s.ldk % 8 == 0 && s.ldv % 8 == 0 && s.ldo % 8 == 0, DLB_ERR_ALIGN, "attention: strides must be multiples of 8");
      g.q = (const bf16*)s.q; g.k = (const bf16*)s.k; g.v = (const bf16*)s.v; g.o = (const bf16*)s.o; g.out = (bf16*)s.o;
      g.dout = (const bf16*)s.dout; g.dq = (bf16*)s.dq; g.dk = (bf16*)s.dk; g.dv = (bf16*)s.dv;
      g.ldq = s.ldq; g.ldk = s.ldk; g.ldv = s.ldv; g.ldo = s.ldo; g.lddo = s.lddo; g.lddq = s.lddq; g.lddk = s.lddk; g.lddv = s.lddv;
      g.len = s.len;
      if (bwd) DLB_REQUIRE(s.lddo % 8 == 0 && s.lddq % 8 == 0 && s.lddk % 8 == 0 && s.lddv % 8 == 0, DLB_ERR_ALIGN, "attention bwd: strides must be multiples of 8");
      S += s.len;
    } else {
      g = Seg{};
      g.len = 0;
    }
  }
  DLB_REQUIRE(S > 0, DLB_ERR_SHAPE, "attention: empty sequence");
  DLB_REQUIRE(mask_len >= 0 && mask_len <= S && (kmask != nullptr || mask_len == 0), DLB_ERR_SHAPE, "attention: bad mask_len %d", mask_len);
  p.lse = lse; p.dsum = dsum; p.kmask = kmask; p.mask_len = mask_len;
  p.B = B; p.H = H; p.S = S; p.hd = hd;
  p.scale = scale; p.scale_log2 = scale * 1.4426950408889634f;
  return DLB_OK;
}

DLB_EXPORT int dlb_attn_fwd(const dlb_attn_seg* segs, int nseg, float* lse, const uint8_t* kmask, int mask_len, int B,
                            int H, int hd, float scale, cudaStream_t stream) {
  AttnParams p;
  int rc = fill_params(p, segs, nseg, lse, nullptr, kmask, mask_len, B, H, hd, scale, false);
  if (rc) return rc;
  dim3 grid((p.S + 63) / 64, H, B);
  HDP_SWITCH(hd, {
    const size_t smem = (size_t)3 * 64 * (HDP + 8) * 2 + 64 * 4;
    if ((rc = set_smem(attn_fwd_kernel<HDP>, smem))) return rc;
    attn_fwd_kernel<HDP><<<grid, 128, smem, stream>>>(p);
  });
  dlb_count_launch();
  return dlb_check_launch("attn_fwd");
}

DLB_EXPORT int dlb_attn_bwd(const dlb_attn_seg* segs, int nseg, const float* lse, float* dsum, const uint8_t* kmask,
                            int mask_len, int B, int H, int hd, float scale, cudaStream_t stream) {
  AttnParams p;
  int rc = fill_params(p, segs, nseg, const_cast<float*>(lse), dsum, kmask, mask_len, B, H, hd, scale, true);
  if (rc) return rc;
  DLB_REQUIRE(lse != nullptr && dsum != nullptr, DLB_ERR_SHAPE, "attention bwd: lse and dsum buffers are required");
  const int64_t nrows = (int64_t)B * p.S * H;
  attn_bwd_prep_kernel<<<(unsigned)((nrows + 7) / 8), 256, 0, stream>>>(p);
  dim3 grid((p.S + 63) / 64, H, B);
  HDP_SWITCH(hd, {
    const size_t smem = (size_t)4 * 64 * (HDP + 8) * 2 + 128 * 4;
    if ((rc = set_smem(attn_bwd_dq_kernel<HDP>, smem))) return rc;
    if ((rc = set_smem(attn_bwd_dkv_kernel<HDP>, smem))) return rc;
    attn_bwd_dq_kernel<HDP><<<grid, 128, smem, stream>>>(p);
    attn_bwd_dkv_kernel<HDP><<<grid, 128, smem, stream>>>(p);
  });
  dlb_count_launch(3);
  return dlb_check_launch("attn_bwd");
}
