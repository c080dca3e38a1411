// Lean row-streaming kernels for the bandwidth-class ops of a transformer block (per-SAMPLE modulation fast path).
//
// ncu on the round-1 kernels (profiles/ncu_elementwise_r2_before.txt) showed them ISSUE-bound, not memory-bound: 770 (LayerNorm
// forward), 750 (gated residual) and 1830 (QK-norm + RoPE forward) warp instructions per 1152-wide row at 55-68 % issue-slot
// utilisation, and their time scaled with the SM clock (x1.23 slower inside the power-capped train step). These versions cut the
// instruction count 2.5-10x so that HBM becomes the limit:
//   * one WARP per row, the lane's column units fixed for the whole kernel, so every per-column quantity (gate, RMS scales,
//     rotary pair indices) is loaded / computed ONCE per kernel into registers instead of once per element per row;
//   * no 64-bit divisions or per-element index arithmetic: rows advance by a constant pointer stride;
//   * units of 4 or 8 channels chosen so that all 32 lanes own the same number of units (d = 1152 -> 9 units of 4);
//   * the next row's loads are issued before the current row is processed (register double buffering), so no thread waits on
//     its own load and no shared-memory staging / barrier sits between load and use;
//   * bf16 arithmetic that the reference performs in bf16 runs as packed HMUL2 / HADD2 (bit-identical rounding), everything else
//     in fp32 with one final rounding.
// Eligibility (else the general kernels in norm_mod.cu / qknorm_rope.cu run): d / U == 32 * UPL for U in {8, 4}, rows_per_mod a
// multiple of 8 (per-sample modulation), 16-byte aligned rows.
#pragma once
#include "common.cuh"

namespace lean {
typedef __nv_bfloat16 bf16;
constexpr int WARPS = 8;  // per CTA

__device__ __forceinline__ float blo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bhi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }
__device__ __forceinline__ uint32_t hmul2(uint32_t a, uint32_t b) {
  __nv_bfloat162 r = __hmul2(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b));
  return *reinterpret_cast<uint32_t*>(&r);
}
__device__ __forceinline__ uint32_t hadd2(uint32_t a, uint32_t b) {
  __nv_bfloat162 r = __hadd2(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b));
  return *reinterpret_cast<uint32_t*>(&r);
}

// a unit = U consecutive bf16 channels = U/2 packed words
template <int U>
__device__ __forceinline__ void ldu(const bf16* p, uint32_t* r) {
  if constexpr (U == 8) { const uint4 v = *reinterpret_cast<const uint4*>(p); r[0] = v.x; r[1] = v.y; r[2] = v.z; r[3] = v.w; }
  else { const uint2 v = *reinterpret_cast<const uint2*>(p); r[0] = v.x; r[1] = v.y; }
}
template <int U>
__device__ __forceinline__ void stu(bf16* p, const uint32_t* r) {
  if constexpr (U == 8) *reinterpret_cast<uint4*>(p) = make_uint4(r[0], r[1], r[2], r[3]);
  else *reinterpret_cast<uint2*>(p) = make_uint2(r[0], r[1]);
}
template <int U>
__device__ __forceinline__ void ldf(const float* p, float* f) {  // U consecutive floats (16-byte aligned)
#pragma unroll
  for (int i = 0; i < U; i += 4) *reinterpret_cast<float4*>(f + i) = *reinterpret_cast<const float4*>(p + i);
}
// packed row: UPL units per lane, unit u of lane l covers channels (l + 32 u) * U ...
template <int U, int UPL>
struct Row {
  uint32_t w[UPL][U / 2];
  __device__ __forceinline__ void load(const bf16* row, int lane) {
#pragma unroll
    for (int u = 0; u < UPL; ++u) ldu<U>(row + (lane + 32 * u) * U, w[u]);
  }
  __device__ __forceinline__ void store(bf16* row, int lane) const {
#pragma unroll
    for (int u = 0; u < UPL; ++u) stu<U>(row + (lane + 32 * u) * U, w[u]);
  }
  __device__ __forceinline__ void unpack(float (&f)[UPL][U]) const {
#pragma unroll
    for (int u = 0; u < UPL; ++u)
#pragma unroll
      for (int j = 0; j < U / 2; ++j) { f[u][2 * j] = blo(w[u][j]); f[u][2 * j + 1] = bhi(w[u][j]); }
  }
};

// ---------------------------------------------------------------------------------------------------------
// gated residual forward: out = x + bf16((a1 [+ a2]) * gate)          (reference mmdit.py:300, 307, 529)
// ---------------------------------------------------------------------------------------------------------
template <int U, int UPL, bool TWO>
__global__ void __launch_bounds__(WARPS * 32, 3)
gate_residual_fwd_lean(const bf16* __restrict__ x, const bf16* __restrict__ a1, const bf16* __restrict__ a2, const bf16* __restrict__ gate,
                       int64_t gate_ld, int rows_per_mod, bf16* __restrict__ out, int R, int rows_per_cta) {
  constexpr int d = U * UPL * 32;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int row0 = blockIdx.x * rows_per_cta, row1 = min(R, row0 + rows_per_cta);
  Row<U, UPL> g;
  int cur_sample = -1;
  for (int r = row0 + warp; r < row1; r += WARPS) {
    const int sample = r / rows_per_mod;
    if (sample != cur_sample) { g.load(gate + (int64_t)sample * gate_ld, lane); cur_sample = sample; }
    Row<U, UPL> xv, av, bv;
    const int64_t off = (int64_t)r * d;
    xv.load(x + off, lane);
    av.load(a1 + off, lane);
    if constexpr (TWO) bv.load(a2 + off, lane);
#pragma unroll
    for (int u = 0; u < UPL; ++u)
#pragma unroll
      for (int j = 0; j < U / 2; ++j) {
        uint32_t a = av.w[u][j];
        if constexpr (TWO) a = hadd2(a, bv.w[u][j]);
        xv.w[u][j] = hadd2(xv.w[u][j], hmul2(a, g.w[u][j]));
      }
    xv.store(out + off, lane);
  }
}

// ---------------------------------------------------------------------------------------------------------
// LayerNorm (affine or not) + modulate forward: y = LN(x) * w (1 + scale) + (b (1 + scale) + shift)     (mmdit.py:296-298, 542-548)
// P = w (1 + scale), Q = b (1 + scale) + shift per sample in shared memory (rebuilt when the CTA's sample changes).
// ---------------------------------------------------------------------------------------------------------
template <int U, int UPL>
__global__ void __launch_bounds__(WARPS * 32, 2)
ln_modulate_fwd_lean(const bf16* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b, const bf16* __restrict__ scale,
                     const bf16* __restrict__ shift, int64_t mod_ld, int rows_per_mod, bf16* __restrict__ y, float* __restrict__ mean_out,
                     float* __restrict__ rstd_out, int R, float eps, int rows_per_cta) {
  constexpr int d = U * UPL * 32;
  __shared__ __align__(16) float sP[d];
  __shared__ __align__(16) float sQ[d];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int row0 = blockIdx.x * rows_per_cta, row1 = min(R, row0 + rows_per_cta);  // row0 % 8 == 0, rows_per_mod % 8 == 0
  int cur_sample = -1;
  // Three row buffers per warp used in rotation (the loop is unrolled by three, so no buffer is ever COPIED: a register move of
  // a row whose load is still in flight would stall on it and cancel the prefetch). The row being processed was requested two
  // iterations ago; two further rows are in flight per warp.
  auto step = [&](Row<U, UPL>& buf, int base) {
    const int sample = base / rows_per_mod;  // uniform over the CTA: all 8 rows of this step belong to one sample
    if (sample != cur_sample) {
      __syncthreads();
      const bf16* sc = scale + (int64_t)sample * mod_ld;
      const bf16* sh = shift + (int64_t)sample * mod_ld;
      for (int j = tid; j < d; j += WARPS * 32) {
        const float s1 = bf16_round(1.f + __bfloat162float(sc[j]));  // `1 + scale` is evaluated in bf16 (autocast)
        sP[j] = w ? w[j] * s1 : s1;
        sQ[j] = (w ? b[j] * s1 : 0.f) + __bfloat162float(sh[j]);
      }
      __syncthreads();
      cur_sample = sample;
    }
    const int r = base + warp;
    if (r < row1) {
      // packed fp32 pairs (FADD2 / FFMA2): one issue slot per two channels; two-pass statistics
      f32x2 f[UPL][U / 2];
      f32x2 s2 = make_f32x2(0.f, 0.f);
#pragma unroll
      for (int u = 0; u < UPL; ++u)
#pragma unroll
        for (int j = 0; j < U / 2; ++j) { f[u][j] = unpack2(buf.w[u][j]); s2 = add2(s2, f[u][j]); }
      if (r + 3 * WARPS < row1) buf.load(x + (int64_t)(r + 3 * WARPS) * d, lane);  // refill this buffer: consumed above
      float s_lo, s_hi;
      split_f32x2(s2, s_lo, s_hi);
      const float mean = warp_sum(s_lo + s_hi) * (1.f / d);
      const f32x2 nmean2 = make_f32x2(-mean, -mean);
      f32x2 q2 = make_f32x2(0.f, 0.f);
#pragma unroll
      for (int u = 0; u < UPL; ++u)
#pragma unroll
        for (int j = 0; j < U / 2; ++j) { f[u][j] = add2(f[u][j], nmean2); q2 = fma2(f[u][j], f[u][j], q2); }
      float q_lo, q_hi;
      split_f32x2(q2, q_lo, q_hi);
      const float rstd = rsqrtf(warp_sum(q_lo + q_hi) * (1.f / d) + eps);
      if (lane == 0 && mean_out) { mean_out[r] = mean; rstd_out[r] = rstd; }
      const f32x2 rstd2 = make_f32x2(rstd, rstd);
      Row<U, UPL> o;
#pragma unroll
      for (int u = 0; u < UPL; ++u) {
        float pv[U], qv[U];
        ldf<U>(sP + (lane + 32 * u) * U, pv);
        ldf<U>(sQ + (lane + 32 * u) * U, qv);
#pragma unroll
        for (int j = 0; j < U / 2; ++j)
          o.w[u][j] = pack2(fma2(f[u][j], mul2(rstd2, make_f32x2(pv[2 * j], pv[2 * j + 1])), make_f32x2(qv[2 * j], qv[2 * j + 1])));
      }
      o.store(y + (int64_t)r * d, lane);
    }
  };
  Row<U, UPL> b0, b1, b2;
  if (row0 + warp < row1) b0.load(x + (int64_t)(row0 + warp) * d, lane);
  if (row0 + warp + WARPS < row1) b1.load(x + (int64_t)(row0 + warp + WARPS) * d, lane);
  if (row0 + warp + 2 * WARPS < row1) b2.load(x + (int64_t)(row0 + warp + 2 * WARPS) * d, lane);
  for (int base = row0; base < row1; base += 3 * WARPS) {  // (all warps run every step: the sample switch holds CTA barriers)
    step(b0, base);
    if (base + WARPS < row1) step(b1, base + WARPS);
    if (base + 2 * WARPS < row1) step(b2, base + 2 * WARPS);
  }
}

// ---------------------------------------------------------------------------------------------------------
// QK-RMSNorm (over the whole inner dim, learnable per-channel scale) + N-D interleaved-pair RoPE, forward.
// Reference: RMSNorm / QKNorm nn.py:423-475, RotaryPositionalEmbeddingNDim nn.py:331-400 (mmdit.py:81-89).
// A warp owns one HALF-row (the q part or the k part of a token): the lane's RMS scales and rotary pair indices are loop
// invariants in registers; cos / sin arrive as ONE vector load per unit (pairs of a unit are adjacent in the table).
// fp32 arithmetic with a single final rounding (the reference rounds to bf16 after the normalisation, after the scale and
// inside the rotation; tests bound the difference against the fp32 restatement at 1e-2, this version is closer to it).
// (A packed-fp32 / three-buffer variant like ln_modulate_fwd_lean measured 5 % SLOWER here, same box: the register pairs cost
// this kernel its third resident CTA's worth of latency hiding; the scalar form stays.)
// ---------------------------------------------------------------------------------------------------------
constexpr int QK_WARPS = 4;  // 2 q-warps + 2 k-warps per CTA (the per-lane scale registers make this kernel register-heavy)
template <int U, int UPL>
__global__ void __launch_bounds__(QK_WARPS * 32, 3)
qknorm_rope_fwd_lean(const bf16* __restrict__ qkv, int64_t ld_in, const float* __restrict__ sq, const float* __restrict__ sk,
                     const uint32_t* __restrict__ cs_t, const int32_t* __restrict__ pos_idx, int rot_half, int pos_offset, int tokens_per_sample,
                     int hd, bf16* __restrict__ out, int64_t ld_out, float* __restrict__ rrms_out, int R, float eps, int rows_per_cta) {
  constexpr int d = U * UPL * 32;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int which = warp & 1;  // 0: q half-rows, 1: k half-rows
  const int row0 = blockIdx.x * rows_per_cta, row1 = min(R, row0 + rows_per_cta);
  // loop invariants of this lane: learnable scales and the rotary pair index of every unit (-1: the unit does not rotate)
  float sc[UPL][U];
  int pj[UPL];
#pragma unroll
  for (int u = 0; u < UPL; ++u) {
    const int c = (lane + 32 * u) * U;
    ldf<U>((which ? sk : sq) + c, sc[u]);
    const int cl = c % hd;  // hd % 8 == 0 keeps a unit inside one head
    pj[u] = (cl >> 1) < rot_half ? (cl >> 1) : -1;
  }
  const int col0 = which * d;
  constexpr int STEP = QK_WARPS / 2;  // token rows advanced per iteration by the warps of each half
  Row<U, UPL> cur, nxt;
  int r = row0 + (warp >> 1);
  if (r < row1) cur.load(qkv + (int64_t)r * ld_in + col0, lane);
  for (; r < row1; r += STEP) {
    if (r + STEP < row1) nxt.load(qkv + (int64_t)(r + STEP) * ld_in + col0, lane);
    const uint32_t* csr = cs_t + (int64_t)(pos_idx ? pos_idx[r] : pos_offset + r % tokens_per_sample) * rot_half;
    float f[UPL][U];
    cur.unpack(f);
    float ss = 0.f;
#pragma unroll
    for (int u = 0; u < UPL; ++u) {
      float t = 0.f;
#pragma unroll
      for (int j = 0; j < U; ++j) t = fmaf(f[u][j], f[u][j], t);
      ss += t;
    }
    const float rrms = rsqrtf(warp_sum(ss) * (1.f / d) + eps);
    if (lane == 0 && rrms_out) rrms_out[(int64_t)r * 2 + which] = rrms;
    Row<U, UPL> o;
#pragma unroll
    for (int u = 0; u < UPL; ++u) {
      uint32_t cs[U / 2];
      if (pj[u] >= 0) {
        if constexpr (U == 8) { const uint4 v = __ldg(reinterpret_cast<const uint4*>(csr + pj[u])); cs[0] = v.x; cs[1] = v.y; cs[2] = v.z; cs[3] = v.w; }
        else { const uint2 v = __ldg(reinterpret_cast<const uint2*>(csr + pj[u])); cs[0] = v.x; cs[1] = v.y; }
      }
#pragma unroll
      for (int j = 0; j < U / 2; ++j) {
        float e = f[u][2 * j] * rrms * sc[u][2 * j], od = f[u][2 * j + 1] * rrms * sc[u][2 * j + 1];
        if (pj[u] >= 0) {  // (cos, sin) packed as bf16x2: the reference casts the tables to the activation dtype
          const float c = blo(cs[j]), s = bhi(cs[j]);
          const float re = e * c - od * s, ro = fmaf(e, s, od * c);
          e = re;
          od = ro;
        }
        o.w[u][j] = pack_bf16x2(e, od);
      }
    }
    o.store(out + (int64_t)r * ld_out + col0, lane);
    cur = nxt;
  }
}

// =========================================================================================================
// Backward kernels. They keep per-column gradient accumulators in registers (2 x d / 32 floats per lane for the LayerNorm),
// which leaves no room for register double buffering: every warp streams ITS rows through a private two-stage cp.async ring in
// shared memory instead (global -> shared without registers, the next row in flight while the current one is processed). There
// is no CTA-wide barrier and no producer warp in the row loop; the CTA only synchronises when the per-sample accumulators are
// flushed (shared-memory reduction over the 8 warps, then one vector atomic per 4 columns).
// =========================================================================================================
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }
// the warp copies one contiguous row of BYTES bytes (16-byte aligned on both sides)
template <int BYTES>
__device__ __forceinline__ void copy_row(uint8_t* dst, const void* src, int lane) {
#pragma unroll
  for (int i = 0; i < (BYTES + 511) / 512; ++i) {
    const int o = (lane + 32 * i) * 16;
    if (o < BYTES) cp_async16(dst + o, reinterpret_cast<const uint8_t*>(src) + o);
  }
}
template <int U>
__device__ __forceinline__ void ldsu(const uint8_t* row, int unit, uint32_t* r) {  // unit `unit` of a staged row
  if constexpr (U == 8) { const uint4 v = *reinterpret_cast<const uint4*>(row + unit * 16); r[0] = v.x; r[1] = v.y; r[2] = v.z; r[3] = v.w; }
  else { const uint2 v = *reinterpret_cast<const uint2*>(row + unit * 8); r[0] = v.x; r[1] = v.y; }
}
// CTA-collective flush of per-lane column accumulators acc[UPL][U] (one set per warp) into global fp32 row `dst`:
// shared-memory atomics reduce the 8 warps, then each thread adds whole 4-column groups with one vector atomic.
template <int U, int UPL>
__device__ __forceinline__ void flush_columns(float (&acc)[UPL][U], float* sAcc, float* dst, int tid, int lane) {
  constexpr int d = U * UPL * 32;
  for (int c = tid; c < d; c += WARPS * 32) sAcc[c] = 0.f;
  __syncthreads();
#pragma unroll
  for (int u = 0; u < UPL; ++u)
#pragma unroll
    for (int j = 0; j < U; ++j) { atomicAdd(sAcc + (lane + 32 * u) * U + j, acc[u][j]); acc[u][j] = 0.f; }
  __syncthreads();
  for (int c4 = tid; c4 < d / 4; c4 += WARPS * 32) atomicAdd(reinterpret_cast<float4*>(dst) + c4, *reinterpret_cast<const float4*>(sAcc + c4 * 4));
  __syncthreads();
}

// ---------------------------------------------------------------------------------------------------------
// LayerNorm + modulate backward, per-sample modulation (algebra: DESIGN.md "LN backward algebra"): ONE pass produces
//   dx = rstd (q - mean(q) - xhat mean(q xhat)) (+ dres),  q = dy G,  G = w (1 + scale)
// and the per-sample column sums S1 = sum dy, S2 = sum dy xhat (finalised by ln_modulate_bwd_finalize_kernel).
// ---------------------------------------------------------------------------------------------------------
constexpr int LNB_NS = 3;  // ring depth: two rows (x 2-3 tensors) in flight per warp
template <int U, int UPL, bool HAS_RES>
__global__ void __launch_bounds__(WARPS * 32, 1)
ln_modulate_bwd_lean(const bf16* __restrict__ dy, const bf16* __restrict__ x, const float* __restrict__ mean_in, const float* __restrict__ rstd_in,
                     const float* __restrict__ w, const bf16* __restrict__ scale, int64_t mod_ld, int rows_per_mod, const bf16* __restrict__ dres,
                     bf16* __restrict__ dx, float* __restrict__ acc1, float* __restrict__ acc2, int64_t acc_ld, int R, int rows_per_cta) {
  constexpr int d = U * UPL * 32, ROWB = d * 2, NIN = HAS_RES ? 3 : 2;
  extern __shared__ __align__(16) uint8_t smem[];
  float* sG = reinterpret_cast<float*>(smem);            // [d] G of the current sample
  float* sAcc = sG + d;                                  // [d] flush buffer
  uint8_t* ring = reinterpret_cast<uint8_t*>(sAcc + d);  // [WARPS][LNB_NS][NIN][ROWB]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint8_t* my = ring + (size_t)warp * LNB_NS * NIN * ROWB;
  const int row0 = blockIdx.x * rows_per_cta, row1 = min(R, row0 + rows_per_cta);
  // column accumulators and all row arithmetic as packed fp32 pairs (FADD2 / FMUL2 / FFMA2: one issue slot per two channels)
  f32x2 S1[UPL][U / 2], S2[UPL][U / 2];
#pragma unroll
  for (int u = 0; u < UPL; ++u)
#pragma unroll
    for (int j = 0; j < U / 2; ++j) { S1[u][j] = make_f32x2(0.f, 0.f); S2[u][j] = make_f32x2(0.f, 0.f); }
  auto flush = [&](f32x2 (&S)[UPL][U / 2], float* dst) {
    float acc[UPL][U];
#pragma unroll
    for (int u = 0; u < UPL; ++u)
#pragma unroll
      for (int j = 0; j < U / 2; ++j) { split_f32x2(S[u][j], acc[u][2 * j], acc[u][2 * j + 1]); S[u][j] = make_f32x2(0.f, 0.f); }
    flush_columns<U, UPL>(acc, sAcc, dst, tid, lane);
  };
  auto prefetch = [&](int r, int stage) {
    if (r < row1) {
      uint8_t* sb = my + stage * NIN * ROWB;
      copy_row<ROWB>(sb, dy + (int64_t)r * d, lane);
      copy_row<ROWB>(sb + ROWB, x + (int64_t)r * d, lane);
      if constexpr (HAS_RES) copy_row<ROWB>(sb + 2 * ROWB, dres + (int64_t)r * d, lane);
    }
    cp_async_commit();
  };
  prefetch(row0 + warp, 0);
  prefetch(row0 + warp + WARPS, 1);
  int cur_sample = -1, stage = 0;
  for (int base = row0; base < row1; base += WARPS, stage = stage == LNB_NS - 1 ? 0 : stage + 1) {
    const int sample = base / rows_per_mod;  // uniform over the CTA (row0 and rows_per_mod are multiples of 8)
    if (sample != cur_sample) {
      if (cur_sample >= 0) {
        flush(S1, acc1 + (int64_t)cur_sample * acc_ld);
        flush(S2, acc2 + (int64_t)cur_sample * acc_ld);
      }
      const bf16* sc = scale + (int64_t)sample * mod_ld;
      for (int j = tid; j < d; j += WARPS * 32) {
        const float s1 = bf16_round(1.f + __bfloat162float(sc[j]));
        sG[j] = w ? w[j] * s1 : s1;
      }
      __syncthreads();
      cur_sample = sample;
    }
    const int r = base + warp;
    prefetch(r + 2 * WARPS, stage >= 1 ? stage - 1 : LNB_NS - 1);  // (stage + 2) % 3: the stage released one iteration ago
    float mean = 0.f, rstd = 0.f;
    if (r < row1) { mean = __ldg(mean_in + r); rstd = __ldg(rstd_in + r); }
    cp_async_wait<2>();
    __syncwarp();
    if (r < row1) {
      const uint8_t* sb = my + stage * NIN * ROWB;
      const float nmr = -mean * rstd;
      const f32x2 rstd2 = make_f32x2(rstd, rstd), nmr2 = make_f32x2(nmr, nmr);
      f32x2 q[UPL][U / 2], xh[UPL][U / 2];
      f32x2 p1 = make_f32x2(0.f, 0.f), p2 = make_f32x2(0.f, 0.f);
#pragma unroll
      for (int u = 0; u < UPL; ++u) {
        const int unit = lane + 32 * u;
        uint32_t gw[U / 2], xw[U / 2];
        float G[U];
        ldsu<U>(sb, unit, gw);
        ldsu<U>(sb + ROWB, unit, xw);
        ldf<U>(sG + unit * U, G);
#pragma unroll
        for (int j = 0; j < U / 2; ++j) {
          const f32x2 g = unpack2(gw[j]);
          xh[u][j] = fma2(unpack2(xw[j]), rstd2, nmr2);
          S1[u][j] = add2(S1[u][j], g);
          S2[u][j] = fma2(g, xh[u][j], S2[u][j]);
          q[u][j] = mul2(g, make_f32x2(G[2 * j], G[2 * j + 1]));
          p1 = add2(p1, q[u][j]);
          p2 = fma2(q[u][j], xh[u][j], p2);
        }
      }
      float p1a, p1b, p2a, p2b;
      split_f32x2(p1, p1a, p1b);
      split_f32x2(p2, p2a, p2b);
      float s1 = p1a + p1b, s2 = p2a + p2b;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
      const float c1 = rstd * s1 * (1.f / d), c2 = rstd * s2 * (1.f / d);  // rstd * mean(q), rstd * mean(q xhat)
      const f32x2 nc1 = make_f32x2(-c1, -c1), nc2 = make_f32x2(-c2, -c2);
      bf16* out = dx + (int64_t)r * d;
#pragma unroll
      for (int u = 0; u < UPL; ++u) {
        const int unit = lane + 32 * u;
        uint32_t rw[U / 2], ow[U / 2];
        if constexpr (HAS_RES) ldsu<U>(sb + 2 * ROWB, unit, rw);
#pragma unroll
        for (int j = 0; j < U / 2; ++j) {
          f32x2 t = fma2(xh[u][j], nc2, fma2(q[u][j], rstd2, nc1));
          if constexpr (HAS_RES) t = add2(t, unpack2(rw[j]));
          ow[j] = pack2(t);
        }
        stu<U>(out + unit * U, ow);
      }
    }
    __syncwarp();  // every lane is done with this stage before the warp refills it (next iteration's prefetch)
  }
  cp_async_wait<0>();
  if (cur_sample >= 0) {
    flush(S1, acc1 + (int64_t)cur_sample * acc_ld);
    flush(S2, acc2 + (int64_t)cur_sample * acc_ld);
  }
}

// ---------------------------------------------------------------------------------------------------------
// gated residual backward, per-sample gate: da = bf16(dout * gate), dgate[sample] += sum_rows dout * (a1 [+ a2])
// ---------------------------------------------------------------------------------------------------------
template <int U, int UPL, bool TWO>
__global__ void __launch_bounds__(WARPS * 32, 2)
gate_residual_bwd_lean(const bf16* __restrict__ dout, const bf16* __restrict__ a1, const bf16* __restrict__ a2, const bf16* __restrict__ gate,
                       int64_t gate_ld, int rows_per_mod, bf16* __restrict__ da, float* __restrict__ dgate, int64_t dgate_ld, int R,
                       int rows_per_cta) {
  constexpr int d = U * UPL * 32;
  __shared__ __align__(16) float sAcc[d];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int row0 = blockIdx.x * rows_per_cta, row1 = min(R, row0 + rows_per_cta);
  float S[UPL][U];
#pragma unroll
  for (int u = 0; u < UPL; ++u)
#pragma unroll
    for (int j = 0; j < U; ++j) S[u][j] = 0.f;
  Row<U, UPL> g;
  int cur_sample = -1;
  // 16 resident warps per SM each keep one row pair (2 x d x 2 bytes) in flight: no explicit prefetch needed
  for (int base = row0; base < row1; base += WARPS) {
    const int sample = base / rows_per_mod;
    if (sample != cur_sample) {
      if (cur_sample >= 0) flush_columns<U, UPL>(S, sAcc, dgate + (int64_t)cur_sample * dgate_ld, tid, lane);
      g.load(gate + (int64_t)sample * gate_ld, lane);
      cur_sample = sample;
    }
    const int r = base + warp;
    if (r < row1) {
      Row<U, UPL> dd, aa, bb;
      dd.load(dout + (int64_t)r * d, lane);
      aa.load(a1 + (int64_t)r * d, lane);
      if constexpr (TWO) bb.load(a2 + (int64_t)r * d, lane);
#pragma unroll
      for (int u = 0; u < UPL; ++u)
#pragma unroll
        for (int j = 0; j < U / 2; ++j) {
          uint32_t a = aa.w[u][j];
          if constexpr (TWO) a = hadd2(a, bb.w[u][j]);
          const uint32_t dv = dd.w[u][j];
          S[u][2 * j] = fmaf(blo(dv), blo(a), S[u][2 * j]);
          S[u][2 * j + 1] = fmaf(bhi(dv), bhi(a), S[u][2 * j + 1]);
          dd.w[u][j] = hmul2(dv, g.w[u][j]);
        }
      dd.store(da + (int64_t)r * d, lane);
    }
  }
  if (cur_sample >= 0) flush_columns<U, UPL>(S, sAcc, dgate + (int64_t)cur_sample * dgate_ld, tid, lane);
}

// ---------------------------------------------------------------------------------------------------------
// QK-RMSNorm + RoPE backward: one warp per half-row (q or k part of a token). Input: gradient wrt the normalised, scaled,
// rotated half-row; output: gradient wrt the raw projection half-row; the gradient of the learnable scale accumulates per lane.
//   gz = R^T g (rotation transposed),  dscale += gz * bf16(x rrms),  gn = gz * scale,  dx = rrms (gn - x rrms mean(gn x rrms))
// ---------------------------------------------------------------------------------------------------------
constexpr int QB_NS = 4;  // ring depth: three half-rows (x 2 tensors) in flight per warp
template <int U, int UPL>
__global__ void __launch_bounds__(WARPS * 32, 1)
qknorm_rope_bwd_lean(const bf16* __restrict__ dqk, int64_t ld_dqk, const bf16* __restrict__ qkv, int64_t ld_in, const float* __restrict__ sq,
                     const float* __restrict__ sk, const uint32_t* __restrict__ cs_t, const int32_t* __restrict__ pos_idx, int rot_half, int pos_offset,
                     int tokens_per_sample, int hd, const float* __restrict__ rrms_in, bf16* __restrict__ dqkv, int64_t ld_out,
                     float* __restrict__ dsq, float* __restrict__ dsk, int R, int rows_per_cta) {
  constexpr int d = U * UPL * 32, ROWB = d * 2;
  extern __shared__ __align__(16) uint8_t smem[];
  float* sAcc = reinterpret_cast<float*>(smem);              // [2][d] flush buffers (q, k)
  uint8_t* ring = reinterpret_cast<uint8_t*>(sAcc + 2 * d);  // [WARPS][QB_NS stages][2 tensors][ROWB]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int which = warp & 1;
  uint8_t* my = ring + (size_t)warp * QB_NS * 2 * ROWB;
  const int row0 = blockIdx.x * rows_per_cta, row1 = min(R, row0 + rows_per_cta);
  const int col0 = which * d;
  // Packed fp32 pairs over two adjacent rotary pairs, as in the forward kernel: E = (even_j, even_j+1), O = (odd_j, odd_j+1).
  f32x2 sce[UPL][U / 4], sco[UPL][U / 4], SE[UPL][U / 4], SO[UPL][U / 4];
  int pj[UPL];
#pragma unroll
  for (int u = 0; u < UPL; ++u) {
    const int c = (lane + 32 * u) * U;
    float sc[U];
    ldf<U>((which ? sk : sq) + c, sc);
    const int cl = c % hd;
    pj[u] = (cl >> 1) < rot_half ? (cl >> 1) : -1;
#pragma unroll
    for (int j = 0; j < U / 4; ++j) {
      sce[u][j] = make_f32x2(sc[4 * j], sc[4 * j + 2]);
      sco[u][j] = make_f32x2(sc[4 * j + 1], sc[4 * j + 3]);
      SE[u][j] = make_f32x2(0.f, 0.f);
      SO[u][j] = make_f32x2(0.f, 0.f);
    }
  }
  constexpr int STEP = WARPS / 2;
  auto prefetch = [&](int r, int stage) {
    if (r < row1) {
      uint8_t* sb = my + stage * 2 * ROWB;
      copy_row<ROWB>(sb, dqk + (int64_t)r * ld_dqk + col0, lane);
      copy_row<ROWB>(sb + ROWB, qkv + (int64_t)r * ld_in + col0, lane);
    }
    cp_async_commit();
  };
  int r = row0 + (warp >> 1), stage = 0;
  prefetch(r, 0);
  prefetch(r + STEP, 1);
  prefetch(r + 2 * STEP, 2);
  for (; r < row1; r += STEP, stage = stage == QB_NS - 1 ? 0 : stage + 1) {
    prefetch(r + 3 * STEP, stage >= 1 ? stage - 1 : QB_NS - 1);  // (stage + 3) % 4: the stage released one iteration ago
    const float rrms = __ldg(rrms_in + (int64_t)r * 2 + which);
    const f32x2 rr2 = make_f32x2(rrms, rrms);
    const uint32_t* csr = cs_t + (int64_t)(pos_idx ? pos_idx[r] : pos_offset + r % tokens_per_sample) * rot_half;
    cp_async_wait<3>();
    __syncwarp();
    const uint8_t* sb = my + stage * 2 * ROWB;
    f32x2 gne[UPL][U / 4], gno[UPL][U / 4], xe[UPL][U / 4], xo[UPL][U / 4];
    f32x2 dot2 = make_f32x2(0.f, 0.f);
#pragma unroll
    for (int u = 0; u < UPL; ++u) {
      const int unit = lane + 32 * u;
      uint32_t gw[U / 2], xw[U / 2], cs[U / 2];
      ldsu<U>(sb, unit, gw);
      ldsu<U>(sb + ROWB, unit, xw);
      if (pj[u] >= 0) {
        if constexpr (U == 8) { const uint4 v = __ldg(reinterpret_cast<const uint4*>(csr + pj[u])); cs[0] = v.x; cs[1] = v.y; cs[2] = v.z; cs[3] = v.w; }
        else { const uint2 v = __ldg(reinterpret_cast<const uint2*>(csr + pj[u])); cs[0] = v.x; cs[1] = v.y; }
      }
#pragma unroll
      for (int j = 0; j < U / 4; ++j) {
        f32x2 ge = make_f32x2(blo(gw[2 * j]), blo(gw[2 * j + 1])), go = make_f32x2(bhi(gw[2 * j]), bhi(gw[2 * j + 1]));
        if (pj[u] >= 0) {  // transposed rotation: e' = e c + o s,  o' = o c - e s
          const uint32_t c0 = cs[2 * j], c1 = cs[2 * j + 1];
          const f32x2 C = make_f32x2(blo(c0), blo(c1)), S = make_f32x2(bhi(c0), bhi(c1));
          const f32x2 nS = make_f32x2(__uint_as_float((c0 & 0xffff0000u) ^ 0x80000000u), __uint_as_float((c1 & 0xffff0000u) ^ 0x80000000u));
          const f32x2 e2 = fma2(go, S, mul2(ge, C)), o2 = fma2(ge, nS, mul2(go, C));
          ge = e2;
          go = o2;
        }
        xe[u][j] = mul2(make_f32x2(blo(xw[2 * j]), blo(xw[2 * j + 1])), rr2);
        xo[u][j] = mul2(make_f32x2(bhi(xw[2 * j]), bhi(xw[2 * j + 1])), rr2);
        // the reference multiplies the bf16-rounded normalised value into the scale gradient
        SE[u][j] = fma2(ge, unpack2(pack2(xe[u][j])), SE[u][j]);
        SO[u][j] = fma2(go, unpack2(pack2(xo[u][j])), SO[u][j]);
        gne[u][j] = mul2(ge, sce[u][j]);
        gno[u][j] = mul2(go, sco[u][j]);
        dot2 = fma2(gne[u][j], xe[u][j], dot2);
        dot2 = fma2(gno[u][j], xo[u][j], dot2);
      }
    }
    float dot_a, dot_b;
    split_f32x2(dot2, dot_a, dot_b);
    const float dot = warp_sum(dot_a + dot_b) * (1.f / d);
    const f32x2 ndot2 = make_f32x2(-dot, -dot);
    bf16* out = dqkv + (int64_t)r * ld_out + col0;
#pragma unroll
    for (int u = 0; u < UPL; ++u) {
      uint32_t ow[U / 2];
#pragma unroll
      for (int j = 0; j < U / 4; ++j) {
        float e_a, e_b, o_a, o_b;
        split_f32x2(mul2(rr2, fma2(xe[u][j], ndot2, gne[u][j])), e_a, e_b);
        split_f32x2(mul2(rr2, fma2(xo[u][j], ndot2, gno[u][j])), o_a, o_b);
        ow[2 * j] = pack_bf16x2(e_a, o_a);
        ow[2 * j + 1] = pack_bf16x2(e_b, o_b);
      }
      stu<U>(out + (lane + 32 * u) * U, ow);
    }
    __syncwarp();
  }
  cp_async_wait<0>();
  if (dsq != nullptr) {  // flush: warps of each half reduce through shared memory, one vector atomic per 4 columns
    float* mine = sAcc + which * d;
    for (int c = tid; c < 2 * d; c += WARPS * 32) sAcc[c] = 0.f;
    __syncthreads();
#pragma unroll
    for (int u = 0; u < UPL; ++u)
#pragma unroll
      for (int j = 0; j < U / 4; ++j) {
        float e_a, e_b, o_a, o_b;
        split_f32x2(SE[u][j], e_a, e_b);
        split_f32x2(SO[u][j], o_a, o_b);
        float* col = mine + (lane + 32 * u) * U + 4 * j;
        atomicAdd(col, e_a);
        atomicAdd(col + 1, o_a);
        atomicAdd(col + 2, e_b);
        atomicAdd(col + 3, o_b);
      }
    __syncthreads();
    for (int c4 = tid; c4 < 2 * d / 4; c4 += WARPS * 32) {
      float* dst = c4 < d / 4 ? dsq + c4 * 4 : dsk + (c4 - d / 4) * 4;
      atomicAdd(reinterpret_cast<float4*>(dst), *reinterpret_cast<const float4*>(sAcc + c4 * 4));
    }
  }
}

// U / UPL for a channel count, or false when no exact per-lane split exists (general kernels take over)
inline bool pick_units(int d, int& U, int& UPL) {
  if (d % 256 == 0 && d / 256 <= 8) { U = 8; UPL = d / 256; return true; }
  if (d % 128 == 0 && d / 128 <= 12) { U = 4; UPL = d / 128; return true; }
  return false;
}
inline int rows_per_cta_for(int R, int target_ctas) {
  int rpc = (R + target_ctas - 1) / target_ctas;
  return (rpc + WARPS - 1) / WARPS * WARPS;
}

// dispatch over the (U, UPL) pairs that occur in practice: d = 256 .. 2048 in steps of 256 (U = 8) and the multiples of 128 in
// between up to 1536 (U = 4): 384, 640, 896, 1152, 1408
#define DLB_LEAN_SWITCH(d, ...)                                                                   \
  do {                                                                                            \
    int U_ = 0, UPL_ = 0;                                                                         \
    if (!lean::pick_units(d, U_, UPL_)) break;                                                    \
    if (U_ == 8) {                                                                                \
      constexpr int U = 8;                                                                        \
      switch (UPL_) {                                                                             \
        case 1: { constexpr int UPL = 1; __VA_ARGS__; } break;                                    \
        case 2: { constexpr int UPL = 2; __VA_ARGS__; } break;                                    \
        case 3: { constexpr int UPL = 3; __VA_ARGS__; } break;                                    \
        case 4: { constexpr int UPL = 4; __VA_ARGS__; } break;                                    \
        case 5: { constexpr int UPL = 5; __VA_ARGS__; } break;                                    \
        case 6: { constexpr int UPL = 6; __VA_ARGS__; } break;                                    \
        default: break;                                                                           \
      }                                                                                           \
    } else {                                                                                      \
      constexpr int U = 4;                                                                        \
      switch (UPL_) {                                                                             \
        case 1: { constexpr int UPL = 1; __VA_ARGS__; } break;                                    \
        case 3: { constexpr int UPL = 3; __VA_ARGS__; } break;                                    \
        case 5: { constexpr int UPL = 5; __VA_ARGS__; } break;                                    \
        case 7: { constexpr int UPL = 7; __VA_ARGS__; } break;                                    \
        case 9: { constexpr int UPL = 9; __VA_ARGS__; } break;                                    \
        case 11: { constexpr int UPL = 11; __VA_ARGS__; } break;                                  \
        default: break;                                                                           \
      }                                                                                           \
    }                                                                                             \
  } while (0)

}  // namespace lean
