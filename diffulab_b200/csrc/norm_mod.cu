// Fused LayerNorm + adaLN modulate, gated residual, packed SwiGLU: the bandwidth-bound token-stream kernels.
//
// Reference semantics (bf16 autocast on CUDA):
//   ln_modulate : modulate(nn.LayerNorm(x), scale, shift)      mmdit.py:257-259,299,305 ; nn.py:539-540
//   gate_res    : x + branch * gate                            mmdit.py:296-307
//   swiglu      : silu(x1) * x3 on the packed up-projection    nn.py:478-486
//
// Shapes of the kernels (all accesses are 128-bit, coalesced). The per-sample-modulation hot path uses the TILE kernels
// (cp.async.bulk ring -> shared memory -> compute warps -> bulk store; backward = one pass with per-lane register column
// accumulators); the row / col / stream kernels below remain for per-token modulation (DDT decoder), ragged group sizes
// and unaligned views:
//   row kernels  : one warp owns one token row (d <= 2048); the row is kept PACKED (bf16x8 vectors) in registers so
//                  the kernels stay at <= 64-80 registers (6-8 CTAs of 4 warps per SM); statistics via warp shuffles.
//   col kernels  : one thread owns one 8-channel vector and marches down a chunk of rows in batches of 4 rows (8+
//                  independent 16-byte loads in flight per thread); per-sample sums land in the fp32 adaLN-gradient
//                  buffer with one low-contention atomic per column per block.
//   stream kernels: grid-stride over vectors, one vector per thread-iteration, <= 42 registers (occupancy beats ILP
//                  here: measured, profiles/elementwise_microbench_r1.jsonl).
#include "common.cuh"
#include "ptx.cuh"
#include "rowwise_lean.cuh"
#include <cstdlib>

namespace {

typedef __nv_bfloat16 bf16;

__device__ __forceinline__ void ldf8(const float* p, float* f) {
  *reinterpret_cast<float4*>(f) = __ldg(reinterpret_cast<const float4*>(p));
  *reinterpret_cast<float4*>(f + 4) = __ldg(reinterpret_cast<const float4*>(p + 4));
}

// ---------------------------------------------------------------------------------------------------------
// LayerNorm (+affine) + modulate forward
// ---------------------------------------------------------------------------------------------------------
template <int VPL>
__global__ void __launch_bounds__(128, 6)
ln_modulate_fwd_kernel(const bf16* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                       const bf16* __restrict__ scale, const bf16* __restrict__ shift, int64_t mod_ld,
                       int rows_per_mod, bf16* __restrict__ y, float* __restrict__ mean_out,
                       float* __restrict__ rstd_out, int64_t R, int d, float eps) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= R) return;
  const int nv = d >> 3;
  const bf16* xr = x + row * d;
  bf16x8 xp[VPL];
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int v = lane + 32 * i;
    if (v < nv) xp[i] = ld8(xr + v * 8);
  }
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    if (lane + 32 * i < nv) {
      float f[8];
      unpack8(xp[i], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) sum += f[j];
    }
  }
  const float mean = warp_sum(sum) / d;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    if (lane + 32 * i < nv) {
      float f[8];
      unpack8(xp[i], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) { const float t = f[j] - mean; sq += t * t; }
    }
  }
  const float rstd = rsqrtf(warp_sum(sq) / d + eps);
  if (lane == 0 && mean_out) { mean_out[row] = mean; rstd_out[row] = rstd; }
  const int64_t mrow = row / rows_per_mod;
  const bf16* sc = scale + mrow * mod_ld;
  const bf16* sh = shift + mrow * mod_ld;
  bf16* yr = y + row * d;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int v = lane + 32 * i;
    if (v < nv) {
      float f[8], s[8], t[8], o[8];
      unpack8(xp[i], f);
      unpack8(ld8(sc + v * 8), s);
      unpack8(ld8(sh + v * 8), t);
      // reference: `1 + scale` is evaluated in bf16 before meeting the fp32 LayerNorm output
      if (w) {
        float wv[8], bv[8];
        ldf8(w + v * 8, wv);
        ldf8(b + v * 8, bv);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = ((f[j] - mean) * rstd * wv[j] + bv[j]) * bf16_round(1.f + s[j]) + t[j];
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = (f[j] - mean) * rstd * bf16_round(1.f + s[j]) + t[j];
      }
      st8(yr + v * 8, pack8(o));
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Tiled LayerNorm + modulate forward (per-sample modulation): the streaming structure that replaces "one warp loads its
// own row". A producer lane moves 4-row tiles (one contiguous 4*d*2-byte block) global -> shared with cp.async.bulk into
// a 3-stage ring (four CTAs per SM); four compute warps take one row each from shared memory, and a finished output tile leaves with one
// cp.async.bulk store. The SM keeps ~150 KB of loads in flight with no thread ever waiting on its own global load; the
// per-sample vectors P = w*(1+scale), Q = b*(1+scale)+shift are rebuilt in shared memory only when the sample changes,
// so a row costs 2 FMA-class ops per element after the statistics. Measured (profiles/elementwise_microbench_r1.jsonl):
// 0.049 ms per XL/2 call vs 0.056-0.061 for the row-per-warp kernel; a copy-only run of the same pipeline takes 0.037 ms,
// the arithmetic adds ~110 SM-cycles per row = its instruction-issue cost (~440 warp instructions per row over 4 schedulers).
// ---------------------------------------------------------------------------------------------------------
constexpr int LT_ROWS = 4, LT_STAGES = 3, LT_WARPS = 4;
template <int VPL>
__global__ void __launch_bounds__((LT_WARPS + 1) * 32, 4)
ln_modulate_fwd_tile_kernel(const bf16* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                            const bf16* __restrict__ scale, const bf16* __restrict__ shift, int64_t mod_ld, int rows_per_mod,
                            bf16* __restrict__ y, float* __restrict__ mean_out, float* __restrict__ rstd_out, int64_t R, int d,
                            float eps, int tiles_per_cta, int dbg_mode) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int tile_bytes = LT_ROWS * d * 2;
  bf16* sIn = reinterpret_cast<bf16*>(smem);                                   // [LT_STAGES][LT_ROWS][d]
  bf16* sOut = reinterpret_cast<bf16*>(smem + LT_STAGES * tile_bytes);        // [2][LT_ROWS][d]
  float* sP = reinterpret_cast<float*>(smem + (LT_STAGES + 2) * tile_bytes);  // [d]
  float* sQ = sP + d;
  __shared__ uint64_t full[LT_STAGES], empty[LT_STAGES];
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int64_t ntiles = (R + LT_ROWS - 1) / LT_ROWS;
  const int64_t t0 = (int64_t)blockIdx.x * tiles_per_cta;
  const int64_t t1 = t0 + tiles_per_cta < ntiles ? t0 + tiles_per_cta : ntiles;
  if (tid == 0) {
    for (int i = 0; i < LT_STAGES; ++i) { ptx::mbar_init(&full[i], 1); ptx::mbar_init(&empty[i], LT_WARPS); }
    ptx::fence_mbar_init();
  }
  __syncthreads();
  if (warp == LT_WARPS) {
    // ---------------- producer ----------------
    int st = 0;
    uint32_t ph = 0;
    for (int64_t t = t0; t < t1; ++t) {
      ptx::mbar_wait(&empty[st], ph ^ 1);
      if (ptx::elect_one()) {
        const int64_t r0 = t * LT_ROWS;
        const int rows = (int)(R - r0 < LT_ROWS ? R - r0 : LT_ROWS);
        const uint32_t bytes = (uint32_t)rows * d * 2;
        ptx::mbar_expect_tx(&full[st], bytes);
        ptx::bulk_load_1d(reinterpret_cast<uint8_t*>(sIn) + st * tile_bytes, x + r0 * d, bytes, &full[st]);
      }
      __syncwarp();
      if (++st == LT_STAGES) { st = 0; ph ^= 1; }
    }
    return;
  }
  // ---------------- compute warps ----------------
  const int nv = d >> 3;
  int st = 0, ob = 0;
  uint32_t ph = 0;
  int64_t cur_sample = -1;
  for (int64_t t = t0; t < t1; ++t) {
    const int64_t r0 = t * LT_ROWS;
    const int64_t sample = r0 / rows_per_mod;
    if (sample != cur_sample) {  // rebuild P, Q for this sample (uniform over the compute warps)
      asm volatile("bar.sync 1, %0;\n" ::"n"(LT_WARPS * 32) : "memory");
      const bf16* sc = scale + sample * mod_ld;
      const bf16* sh = shift + sample * mod_ld;
      for (int j = tid; j < d; j += LT_WARPS * 32) {
        const float s1 = bf16_round(1.f + __bfloat162float(sc[j]));  // `1 + scale` is evaluated in bf16 (autocast)
        sP[j] = w ? w[j] * s1 : s1;
        sQ[j] = (w ? b[j] * s1 : 0.f) + __bfloat162float(sh[j]);
      }
      asm volatile("bar.sync 1, %0;\n" ::"n"(LT_WARPS * 32) : "memory");
      cur_sample = sample;
    }
    // the output staging buffer `ob` was handed to a bulk store two tiles ago: wait until that store has read it
    if (tid == 0) ptx::tma_wait_group_read<1>();
    asm volatile("bar.sync 1, %0;\n" ::"n"(LT_WARPS * 32) : "memory");
    ptx::mbar_wait(&full[st], ph);
    const bf16* tin = sIn + (size_t)st * LT_ROWS * d;
    bf16* tout = sOut + (size_t)ob * LT_ROWS * d;
    if (dbg_mode == 1) {  // development probe: pure shared-memory copy of my rows (no arithmetic)
      for (int rr = 0; rr < LT_ROWS / LT_WARPS; ++rr)
        for (int v = lane; v < nv; v += 32)
          *reinterpret_cast<bf16x8*>(tout + (size_t)(warp + rr * LT_WARPS) * d + v * 8) = *reinterpret_cast<const bf16x8*>(tin + (size_t)(warp + rr * LT_WARPS) * d + v * 8);
    } else {
      // this warp's two rows are processed together, every reduction as independent partial sums: the dependent chain
      // per row is ~10 operations deep instead of ~80 (with two warps per scheduler, chain latency is what binds)
      constexpr int RPW = LT_ROWS / LT_WARPS;
      bf16x8 xp[RPW][VPL];
      float mean[RPW], rstd[RPW];
#pragma unroll
      for (int rr = 0; rr < RPW; ++rr) {
        const bf16* xr = tin + (size_t)(warp + rr * LT_WARPS) * d;
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
          const int v = lane + 32 * i;
          if (v < nv) xp[rr][i] = *reinterpret_cast<const bf16x8*>(xr + v * 8);
        }
      }
      float part[RPW][VPL];
#pragma unroll
      for (int rr = 0; rr < RPW; ++rr)
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
          part[rr][i] = 0.f;
          if (lane + 32 * i < nv) {
            float f[8];
            unpack8(xp[rr][i], f);
            part[rr][i] = ((f[0] + f[1]) + (f[2] + f[3])) + ((f[4] + f[5]) + (f[6] + f[7]));
          }
        }
#pragma unroll
      for (int rr = 0; rr < RPW; ++rr) {
        float sum = part[rr][0];
#pragma unroll
        for (int i = 1; i < VPL; ++i) sum += part[rr][i];
        mean[rr] = sum;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int rr = 0; rr < RPW; ++rr) mean[rr] += __shfl_xor_sync(0xffffffffu, mean[rr], o);
#pragma unroll
      for (int rr = 0; rr < RPW; ++rr) mean[rr] /= d;
#pragma unroll
      for (int rr = 0; rr < RPW; ++rr)
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
          part[rr][i] = 0.f;
          if (lane + 32 * i < nv) {
            float f[8];
            unpack8(xp[rr][i], f);
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] -= mean[rr];
            part[rr][i] = ((f[0] * f[0] + f[1] * f[1]) + (f[2] * f[2] + f[3] * f[3])) + ((f[4] * f[4] + f[5] * f[5]) + (f[6] * f[6] + f[7] * f[7]));
          }
        }
#pragma unroll
      for (int rr = 0; rr < RPW; ++rr) {
        float sq = part[rr][0];
#pragma unroll
        for (int i = 1; i < VPL; ++i) sq += part[rr][i];
        rstd[rr] = sq;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int rr = 0; rr < RPW; ++rr) rstd[rr] += __shfl_xor_sync(0xffffffffu, rstd[rr], o);
#pragma unroll
      for (int rr = 0; rr < RPW; ++rr) {
        rstd[rr] = rsqrtf(rstd[rr] / d + eps);
        const int64_t row = r0 + warp + rr * LT_WARPS;
        if (lane == 0 && mean_out && row < R) { mean_out[row] = mean[rr]; rstd_out[row] = rstd[rr]; }
      }
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        const int v = lane + 32 * i;
        if (v < nv) {
          float pv[8], qv[8];
          *reinterpret_cast<float4*>(pv) = *reinterpret_cast<const float4*>(sP + v * 8);
          *reinterpret_cast<float4*>(pv + 4) = *reinterpret_cast<const float4*>(sP + v * 8 + 4);
          *reinterpret_cast<float4*>(qv) = *reinterpret_cast<const float4*>(sQ + v * 8);
          *reinterpret_cast<float4*>(qv + 4) = *reinterpret_cast<const float4*>(sQ + v * 8 + 4);
#pragma unroll
          for (int rr = 0; rr < RPW; ++rr) {
            float f[8], o[8];
            unpack8(xp[rr][i], f);
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = fmaf((f[j] - mean[rr]) * rstd[rr], pv[j], qv[j]);
            *reinterpret_cast<bf16x8*>(tout + (size_t)(warp + rr * LT_WARPS) * d + v * 8) = pack8(o);
          }
        }
      }
    }
    __syncwarp();
    if (lane == 0) ptx::mbar_arrive(&empty[st]);  // this warp is done with the input stage
    ptx::fence_proxy_async_smem();
    asm volatile("bar.sync 1, %0;\n" ::"n"(LT_WARPS * 32) : "memory");
    if (tid == 0) {
      const int rows = (int)(R - r0 < LT_ROWS ? R - r0 : LT_ROWS);
      ptx::bulk_store_1d(y + r0 * d, tout, (uint32_t)rows * d * 2);
      ptx::tma_commit_group();
    }
    ob ^= 1;
    if (++st == LT_STAGES) { st = 0; ph ^= 1; }
  }
  if (tid == 0) ptx::tma_wait_group<0>();
}

// ---------------------------------------------------------------------------------------------------------
// LayerNorm + modulate backward:
//  (1) rows kernel: dx (+dres); per-token mode also writes the dscale/dshift rows.
//  (2) cols kernel: S1 = sum_rows dy, S2 = sum_rows dy * xhat per modulation group (per-token mode: weighted by
//      (1+scale_row), reduced over all rows straight into db/dw).
//  (3) finalize kernel (per-sample mode): dshift = S1, dscale = w S2 + b S1, dw += sum_g (1+s_g) S2_g,
//      db += sum_g (1+s_g) S1_g.            ("LN backward algebra")
// ---------------------------------------------------------------------------------------------------------
template <int VPL, bool PER_TOKEN>
__global__ void __launch_bounds__(128, 5)
ln_modulate_bwd_rows_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x, const float* __restrict__ mean_in,
                            const float* __restrict__ rstd_in, const float* __restrict__ w, const float* __restrict__ b,
                            const bf16* __restrict__ scale, int64_t mod_ld, int64_t rows_per_mod,
                            const bf16* __restrict__ dres, bf16* __restrict__ dx, bf16* __restrict__ dscale_tok,
                            bf16* __restrict__ dshift_tok, int64_t dtok_ld, int64_t R, int d) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= R) return;
  const int nv = d >> 3;
  const bf16* sc = scale + (row / rows_per_mod) * mod_ld;
  bf16x8 xp[VPL], gp[VPL];
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int v = lane + 32 * i;
    if (v < nv) {
      xp[i] = ld8(x + row * d + v * 8);
      gp[i] = ld8(dy + row * d + v * 8);
    }
  }
  const float mean = mean_in[row], rstd = rstd_in[row];
  float m1 = 0.f, m2 = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int v = lane + 32 * i;
    if (v < nv) {
      float xf[8], g[8], s[8], wv[8];
      unpack8(xp[i], xf);
      unpack8(gp[i], g);
      unpack8(ld8(sc + v * 8), s);
      if (w) ldf8(w + v * 8, wv);
      float ds_tok[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float xh = (xf[j] - mean) * rstd;
        float gq = g[j] * bf16_round(1.f + s[j]);  // grad wrt the affine LayerNorm output u
        if (w) gq *= wv[j];
        m1 += gq;
        m2 += gq * xh;
        if (PER_TOKEN) ds_tok[j] = g[j] * (w ? xh * wv[j] + __ldg(b + v * 8 + j) : xh);
      }
      if (PER_TOKEN) {
        st8(dscale_tok + row * dtok_ld + v * 8, pack8(ds_tok));
        st8(dshift_tok + row * dtok_ld + v * 8, gp[i]);
      }
    }
  }
  bf16x8 rp[VPL];  // residual-branch gradient: fetched while the row statistics are being reduced
  if (dres) {
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int v = lane + 32 * i;
      if (v < nv) rp[i] = ld8(dres + row * d + v * 8);
    }
  }
  m1 = warp_sum(m1) / d;
  m2 = warp_sum(m2) / d;
  // pass 2 recomputes g_w from the packed registers (scale / w come from L1) instead of keeping 40 more floats live
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int v = lane + 32 * i;
    if (v < nv) {
      float xf[8], g[8], s[8], wv[8], o[8];
      unpack8(xp[i], xf);
      unpack8(gp[i], g);
      unpack8(ld8(sc + v * 8), s);
      if (w) ldf8(w + v * 8, wv);
      if (dres) unpack8(rp[i], o);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float xh = (xf[j] - mean) * rstd;
        float gq = g[j] * bf16_round(1.f + s[j]);
        if (w) gq *= wv[j];
        const float t = rstd * (gq - m1 - xh * m2);
        o[j] = dres ? o[j] + t : t;
      }
      st8(dx + row * d + v * 8, pack8(o));
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Tiled LayerNorm + modulate backward (per-sample modulation): ONE pass over dy / x produces dx AND the per-sample
// column sums S1 = sum dy, S2 = sum dy*xhat (the separate column kernel re-read both tensors). Same cp.async.bulk ring as
// the forward tile kernel; because no thread waits on its own global loads, the kernel can afford 80 registers of
// per-lane column accumulators (2 CTAs / SM) — the reason the row-per-warp version could not be fused. Accumulators are
// flushed (warp -> shared -> one atomicAdd per column per CTA) when the sample changes and at the end.
// ---------------------------------------------------------------------------------------------------------
constexpr int LB_ROWS = 4, LB_STAGES = 3, LB_WARPS = 4;
template <int VPL, bool HAS_RES>
__global__ void __launch_bounds__((LB_WARPS + 1) * 32, 2)
ln_modulate_bwd_tile_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x, const float* __restrict__ mean_in,
                            const float* __restrict__ rstd_in, const float* __restrict__ w, const bf16* __restrict__ scale,
                            int64_t mod_ld, int rows_per_mod, const bf16* __restrict__ dres, bf16* __restrict__ dx,
                            float* __restrict__ acc1, float* __restrict__ acc2, int64_t acc_ld, int64_t R, int d, int tiles_per_cta) {
  extern __shared__ __align__(128) uint8_t smem[];
  constexpr int NIN = HAS_RES ? 3 : 2;
  const int tile_bytes = LB_ROWS * d * 2;
  const int stage_bytes = NIN * tile_bytes + 32;  // {dy, x, [dres]} tiles + mean[4] + rstd[4]
  uint8_t* sOut = smem + LB_STAGES * stage_bytes;                       // [2][LB_ROWS][d] bf16
  float* sG = reinterpret_cast<float*>(sOut + 2 * tile_bytes);          // [d]  (1 + scale) * w of the current sample
  float* sAcc = sG + d;                                                  // [2][d] flush buffer
  __shared__ uint64_t full[LB_STAGES], empty[LB_STAGES];
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int64_t ntiles = (R + LB_ROWS - 1) / LB_ROWS;
  const int64_t t0 = (int64_t)blockIdx.x * tiles_per_cta;
  const int64_t t1 = t0 + tiles_per_cta < ntiles ? t0 + tiles_per_cta : ntiles;
  if (tid == 0) {
    for (int i = 0; i < LB_STAGES; ++i) { ptx::mbar_init(&full[i], 1); ptx::mbar_init(&empty[i], LB_WARPS); }
    ptx::fence_mbar_init();
  }
  __syncthreads();
  if (warp == LB_WARPS) {
    // ---------------- producer ----------------
    int st = 0;
    uint32_t ph = 0;
    for (int64_t t = t0; t < t1; ++t) {
      ptx::mbar_wait(&empty[st], ph ^ 1);
      if (ptx::elect_one()) {
        const int64_t r0 = t * LB_ROWS;
        const int rows = (int)(R - r0 < LB_ROWS ? R - r0 : LB_ROWS);
        const uint32_t bytes = (uint32_t)rows * d * 2;
        uint8_t* sb = smem + st * stage_bytes;
        // the statistics ride along: 16 bytes each when the whole tile exists (R % 4 == 0 is required by the launcher)
        ptx::mbar_expect_tx(&full[st], NIN * bytes + 32);
        ptx::bulk_load_1d(sb, dy + r0 * d, bytes, &full[st]);
        ptx::bulk_load_1d(sb + tile_bytes, x + r0 * d, bytes, &full[st]);
        if constexpr (HAS_RES) ptx::bulk_load_1d(sb + 2 * tile_bytes, dres + r0 * d, bytes, &full[st]);
        ptx::bulk_load_1d(sb + NIN * tile_bytes, mean_in + r0, 16, &full[st]);
        ptx::bulk_load_1d(sb + NIN * tile_bytes + 16, rstd_in + r0, 16, &full[st]);
      }
      __syncwarp();
      if (++st == LB_STAGES) { st = 0; ph ^= 1; }
    }
    return;
  }
  // ---------------- compute warps: one row of the tile each ----------------
  const int nv = d >> 3;
  int st = 0, ob = 0;
  uint32_t ph = 0;
  int64_t cur_sample = -1;
  float S1[VPL][8], S2[VPL][8];
#pragma unroll
  for (int i = 0; i < VPL; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) { S1[i][j] = 0.f; S2[i][j] = 0.f; }
  auto flush = [&](int64_t sample) {  // compute-warp collective: column sums of `sample` -> global accumulators
    for (int wv = 0; wv < LB_WARPS; ++wv) {
      if (warp == wv) {
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
          const int v = lane + 32 * i;
          if (v < nv) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float* a = sAcc + v * 8 + j;
              if (wv == 0) { a[0] = S1[i][j]; a[d] = S2[i][j]; }
              else { a[0] += S1[i][j]; a[d] += S2[i][j]; }
              S1[i][j] = 0.f;
              S2[i][j] = 0.f;
            }
          }
        }
      }
      asm volatile("bar.sync 1, %0;\n" ::"n"(LB_WARPS * 32) : "memory");
    }
    for (int c = tid; c < d; c += LB_WARPS * 32) {
      atomicAdd(acc1 + sample * acc_ld + c, sAcc[c]);
      atomicAdd(acc2 + sample * acc_ld + c, sAcc[d + c]);
    }
    asm volatile("bar.sync 1, %0;\n" ::"n"(LB_WARPS * 32) : "memory");
  };
  for (int64_t t = t0; t < t1; ++t) {
    const int64_t r0 = t * LB_ROWS;
    const int64_t sample = r0 / rows_per_mod;
    if (sample != cur_sample) {
      if (cur_sample >= 0) flush(cur_sample);
      else asm volatile("bar.sync 1, %0;\n" ::"n"(LB_WARPS * 32) : "memory");
      const bf16* sc = scale + sample * mod_ld;
      for (int j = tid; j < d; j += LB_WARPS * 32) {
        const float s1 = bf16_round(1.f + __bfloat162float(sc[j]));
        sG[j] = w ? w[j] * s1 : s1;
      }
      asm volatile("bar.sync 1, %0;\n" ::"n"(LB_WARPS * 32) : "memory");
      cur_sample = sample;
    }
    if (tid == 0) ptx::tma_wait_group_read<1>();
    asm volatile("bar.sync 1, %0;\n" ::"n"(LB_WARPS * 32) : "memory");
    ptx::mbar_wait(&full[st], ph);
    const uint8_t* sb = smem + st * stage_bytes;
    bf16* tout = reinterpret_cast<bf16*>(sOut + ob * tile_bytes);
    const int64_t row = r0 + warp;
    if (row < R) {
      const bf16* gr = reinterpret_cast<const bf16*>(sb) + (size_t)warp * d;
      const bf16* xr = reinterpret_cast<const bf16*>(sb + tile_bytes) + (size_t)warp * d;
      const float mean = reinterpret_cast<const float*>(sb + NIN * tile_bytes)[warp];
      const float rstd = reinterpret_cast<const float*>(sb + NIN * tile_bytes + 16)[warp];
      const float nmr = -mean * rstd;
      bf16x8 xp[VPL], gp[VPL];
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        const int v = lane + 32 * i;
        if (v < nv) {
          xp[i] = *reinterpret_cast<const bf16x8*>(xr + v * 8);
          gp[i] = *reinterpret_cast<const bf16x8*>(gr + v * 8);
        }
      }
      float p1[VPL], p2[VPL];
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        p1[i] = 0.f;
        p2[i] = 0.f;
        const int v = lane + 32 * i;
        if (v < nv) {
          float xf[8], g[8], G[8], q[8], qx[8];
          unpack8(xp[i], xf);
          unpack8(gp[i], g);
          *reinterpret_cast<float4*>(G) = *reinterpret_cast<const float4*>(sG + v * 8);
          *reinterpret_cast<float4*>(G + 4) = *reinterpret_cast<const float4*>(sG + v * 8 + 4);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float xh = fmaf(xf[j], rstd, nmr);
            S1[i][j] += g[j];
            S2[i][j] = fmaf(g[j], xh, S2[i][j]);
            q[j] = g[j] * G[j];
            qx[j] = q[j] * xh;
          }
          p1[i] = ((q[0] + q[1]) + (q[2] + q[3])) + ((q[4] + q[5]) + (q[6] + q[7]));
          p2[i] = ((qx[0] + qx[1]) + (qx[2] + qx[3])) + ((qx[4] + qx[5]) + (qx[6] + qx[7]));
        }
      }
      float m1 = p1[0], m2 = p2[0];
#pragma unroll
      for (int i = 1; i < VPL; ++i) { m1 += p1[i]; m2 += p2[i]; }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        m1 += __shfl_xor_sync(0xffffffffu, m1, o);
        m2 += __shfl_xor_sync(0xffffffffu, m2, o);
      }
      m1 /= d;
      m2 /= d;
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        const int v = lane + 32 * i;
        if (v < nv) {
          float xf[8], g[8], G[8], o[8];
          unpack8(xp[i], xf);
          unpack8(gp[i], g);
          *reinterpret_cast<float4*>(G) = *reinterpret_cast<const float4*>(sG + v * 8);
          *reinterpret_cast<float4*>(G + 4) = *reinterpret_cast<const float4*>(sG + v * 8 + 4);
          if constexpr (HAS_RES) unpack8(*reinterpret_cast<const bf16x8*>(reinterpret_cast<const bf16*>(sb + 2 * tile_bytes) + (size_t)warp * d + v * 8), o);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float xh = fmaf(xf[j], rstd, nmr);
            const float tt = rstd * (g[j] * G[j] - m1 - xh * m2);
            o[j] = HAS_RES ? o[j] + tt : tt;
          }
          *reinterpret_cast<bf16x8*>(tout + (size_t)warp * d + v * 8) = pack8(o);
        }
      }
    }
    __syncwarp();
    if (lane == 0) ptx::mbar_arrive(&empty[st]);
    ptx::fence_proxy_async_smem();
    asm volatile("bar.sync 1, %0;\n" ::"n"(LB_WARPS * 32) : "memory");
    if (tid == 0) {
      const int rows = (int)(R - r0 < LB_ROWS ? R - r0 : LB_ROWS);
      ptx::bulk_store_1d(dx + r0 * d, tout, (uint32_t)rows * d * 2);
      ptx::tma_commit_group();
    }
    ob ^= 1;
    if (++st == LB_STAGES) { st = 0; ph ^= 1; }
  }
  if (cur_sample >= 0) flush(cur_sample);
  if (tid == 0) ptx::tma_wait_group<0>();
}

// grid (col chunks, row chunks, groups); block = one thread per 8-channel vector
template <bool PER_TOKEN>
__global__ void __launch_bounds__(256, 3)
ln_modulate_bwd_cols_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x, const float* __restrict__ mean_in,
                            const float* __restrict__ rstd_in, const bf16* __restrict__ scale, int64_t mod_ld,
                            int64_t rows_per_group, float* __restrict__ acc1, float* __restrict__ acc2, int64_t acc_ld,
                            int d, int rows_per_block) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= (d >> 3)) return;
  const int c = v * 8;
  const int64_t group = blockIdx.z;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_block;
  const int64_t r1 = min(r0 + rows_per_block, rows_per_group);
  const int64_t base = group * rows_per_group;
  float S1[8], S2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { S1[j] = 0.f; S2[j] = 0.f; }
  for (int64_t rb = r0; rb < r1; rb += 4) {
    bf16x8 xa[4], ga[4], sa[4];
    float mu[4], rs[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t row = base + rb + u;
      if (rb + u < r1) {
        xa[u] = ld8(x + row * d + c);
        ga[u] = ld8(dy + row * d + c);
        if (PER_TOKEN) sa[u] = ld8(scale + row * mod_ld + c);
        mu[u] = __ldg(mean_in + row);
        rs[u] = __ldg(rstd_in + row);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (rb + u < r1) {
        float xf[8], g[8];
        unpack8(xa[u], xf);
        unpack8(ga[u], g);
        if (PER_TOKEN) {
          float s[8];
          unpack8(sa[u], s);
#pragma unroll
          for (int j = 0; j < 8; ++j) g[j] *= bf16_round(1.f + s[j]);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          S1[j] += g[j];
          S2[j] += g[j] * (xf[j] - mu[u]) * rs[u];
        }
      }
    }
  }
  // per-sample: acc1/acc2 = dshift/dscale rows of this group (raw sums, finalised later); per-token: db/dw
  float* a1 = acc1 + (PER_TOKEN ? 0 : group * acc_ld) + c;
  float* a2 = acc2 + (PER_TOKEN ? 0 : group * acc_ld) + c;
#pragma unroll
  for (int j = 0; j < 8; ++j) { atomicAdd(a1 + j, S1[j]); atomicAdd(a2 + j, S2[j]); }
}

// grid (col blocks, group chunks): thread = one column, loops over a chunk of groups
__global__ void __launch_bounds__(256)
ln_modulate_bwd_finalize_kernel(const float* __restrict__ w, const float* __restrict__ b, const bf16* __restrict__ scale,
                                int64_t mod_ld, float* __restrict__ dscale, float* __restrict__ dshift, int64_t dmod_ld,
                                float* __restrict__ dw, float* __restrict__ db, int d, int64_t groups, int groups_per_block) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= d) return;
  const int64_t g0 = (int64_t)blockIdx.y * groups_per_block;
  const int64_t g1 = min(g0 + groups_per_block, groups);
  const float wc = w ? w[c] : 1.f, bc = w ? b[c] : 0.f;
  float aw = 0.f, ab = 0.f;
#pragma unroll 4
  for (int64_t g = g0; g < g1; ++g) {
    const float s1 = dshift[g * dmod_ld + c], s2 = dscale[g * dmod_ld + c];
    const float one_s = bf16_round(1.f + __bfloat162float(scale[g * mod_ld + c]));
    dscale[g * dmod_ld + c] = wc * s2 + bc * s1;
    aw += one_s * s2;
    ab += one_s * s1;
  }
  if (dw) { atomicAdd(dw + c, aw); atomicAdd(db + c, ab); }
}

// ---------------------------------------------------------------------------------------------------------
// gated residual: out = x + (a1 [+ a2]) * gate        (bf16 roundings placed where the reference rounds)
// ---------------------------------------------------------------------------------------------------------
template <bool TWO>
__global__ void __launch_bounds__(256, 6)
gate_residual_fwd_kernel(const bf16* __restrict__ x, const bf16* __restrict__ a1, const bf16* __restrict__ a2,
                         const bf16* __restrict__ gate, int64_t gate_ld, int rows_per_mod, bf16* __restrict__ out,
                         int64_t R, int d) {
  const int nv = d >> 3;
  const int64_t total = R * nv;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / nv;
    const int v = (int)(i - row * nv);
    const int64_t off = row * d + v * 8;
    const bf16x8 xv = ld8(x + off), av = ld8(a1 + off), gv = ld8(gate + (row / rows_per_mod) * gate_ld + v * 8);
    float xf[8], a[8], g[8], o[8];
    unpack8(av, a);
    if (TWO) {
      float b2[8];
      unpack8(ld8(a2 + off), b2);
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] = bf16_round(a[j] + b2[j]);
    }
    unpack8(xv, xf);
    unpack8(gv, g);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = xf[j] + bf16_round(a[j] * g[j]);
    st8(out + off, pack8(o));
  }
}

// backward part 1 (stream): da = dout * gate  (per-token mode also writes dgate rows = dout * a)
template <bool TWO, bool PER_TOKEN>
__global__ void __launch_bounds__(256, 6)
gate_residual_bwd_stream_kernel(const bf16* __restrict__ dout, const bf16* __restrict__ a1, const bf16* __restrict__ a2,
                                const bf16* __restrict__ gate, int64_t gate_ld, int64_t rows_per_mod,
                                bf16* __restrict__ da, bf16* __restrict__ dgate_tok, int64_t dtok_ld, int64_t R, int d) {
  const int nv = d >> 3;
  const int64_t total = R * nv;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / nv;
    const int v = (int)(i - row * nv);
    const int64_t off = row * d + v * 8;
    float g[8], dd[8], o[8];
    unpack8(ld8(dout + off), dd);
    unpack8(ld8(gate + (row / rows_per_mod) * gate_ld + v * 8), g);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = dd[j] * g[j];
    st8(da + off, pack8(o));
    if (PER_TOKEN) {
      float a[8];
      unpack8(ld8(a1 + off), a);
      if (TWO) {
        float b2[8];
        unpack8(ld8(a2 + off), b2);
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] = bf16_round(a[j] + b2[j]);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = dd[j] * a[j];
      st8(dgate_tok + row * dtok_ld + v * 8, pack8(o));
    }
  }
}

// backward part 2 (cols): dgate[group] += sum_rows dout * (a1 [+ a2])
template <bool TWO>
__global__ void __launch_bounds__(256, 3)
gate_residual_bwd_cols_kernel(const bf16* __restrict__ dout, const bf16* __restrict__ a1, const bf16* __restrict__ a2,
                              int64_t rows_per_group, float* __restrict__ dgate, int64_t dgate_ld, int d, int rows_per_block) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= (d >> 3)) return;
  const int c = v * 8;
  const int64_t group = blockIdx.z;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_block;
  const int64_t r1 = min(r0 + rows_per_block, rows_per_group);
  const int64_t base = group * rows_per_group;
  float S[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) S[j] = 0.f;
  for (int64_t rb = r0; rb < r1; rb += 4) {
    bf16x8 da_[4], aa[4], ba[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t row = base + rb + u;
      if (rb + u < r1) {
        da_[u] = ld8(dout + row * d + c);
        aa[u] = ld8(a1 + row * d + c);
        if (TWO) ba[u] = ld8(a2 + row * d + c);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (rb + u < r1) {
        float g[8], a[8];
        unpack8(da_[u], g);
        unpack8(aa[u], a);
        if (TWO) {
          float b2[8];
          unpack8(ba[u], b2);
#pragma unroll
          for (int j = 0; j < 8; ++j) a[j] = bf16_round(a[j] + b2[j]);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) S[j] += g[j] * a[j];
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) atomicAdd(dgate + group * dgate_ld + c + j, S[j]);
}

// ---------------------------------------------------------------------------------------------------------
// One-pass tiled gate-residual backward (per-sample gate): da = dout * gate AND dgate[sample] += sum_rows dout * (a1 [+ a2])
// from a single read of dout / a1 (the separate column kernel re-read them). Same cp.async.bulk ring + register column
// accumulators as the LayerNorm backward tile kernel.
// ---------------------------------------------------------------------------------------------------------
template <int VPL, bool TWO>
__global__ void __launch_bounds__((LB_WARPS + 1) * 32, 2)
gate_residual_bwd_tile_kernel(const bf16* __restrict__ dout, const bf16* __restrict__ a1, const bf16* __restrict__ a2,
                              const bf16* __restrict__ gate, int64_t gate_ld, int rows_per_mod, bf16* __restrict__ da,
                              float* __restrict__ dgate, int64_t dgate_ld, int64_t R, int d, int tiles_per_cta) {
  extern __shared__ __align__(128) uint8_t smem[];
  constexpr int NIN = TWO ? 3 : 2;
  const int tile_bytes = LB_ROWS * d * 2;
  const int stage_bytes = NIN * tile_bytes;
  uint8_t* sOut = smem + LB_STAGES * stage_bytes;
  float* sAcc = reinterpret_cast<float*>(sOut + 2 * tile_bytes);  // [d] flush buffer
  __shared__ uint64_t full[LB_STAGES], empty[LB_STAGES];
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int64_t ntiles = (R + LB_ROWS - 1) / LB_ROWS;
  const int64_t t0 = (int64_t)blockIdx.x * tiles_per_cta;
  const int64_t t1 = t0 + tiles_per_cta < ntiles ? t0 + tiles_per_cta : ntiles;
  if (tid == 0) {
    for (int i = 0; i < LB_STAGES; ++i) { ptx::mbar_init(&full[i], 1); ptx::mbar_init(&empty[i], LB_WARPS); }
    ptx::fence_mbar_init();
  }
  __syncthreads();
  if (warp == LB_WARPS) {
    int st = 0;
    uint32_t ph = 0;
    for (int64_t t = t0; t < t1; ++t) {
      ptx::mbar_wait(&empty[st], ph ^ 1);
      if (ptx::elect_one()) {
        const int64_t r0 = t * LB_ROWS;
        const int rows = (int)(R - r0 < LB_ROWS ? R - r0 : LB_ROWS);
        const uint32_t bytes = (uint32_t)rows * d * 2;
        uint8_t* sb = smem + st * stage_bytes;
        ptx::mbar_expect_tx(&full[st], NIN * bytes);
        ptx::bulk_load_1d(sb, dout + r0 * d, bytes, &full[st]);
        ptx::bulk_load_1d(sb + tile_bytes, a1 + r0 * d, bytes, &full[st]);
        if constexpr (TWO) ptx::bulk_load_1d(sb + 2 * tile_bytes, a2 + r0 * d, bytes, &full[st]);
      }
      __syncwarp();
      if (++st == LB_STAGES) { st = 0; ph ^= 1; }
    }
    return;
  }
  const int nv = d >> 3;
  int st = 0, ob = 0;
  uint32_t ph = 0;
  int64_t cur_sample = -1;
  float S[VPL][8];
  bf16x8 gv[VPL];  // this lane's slice of the current sample's gate
#pragma unroll
  for (int i = 0; i < VPL; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) S[i][j] = 0.f;
  auto flush = [&](int64_t sample) {
    for (int wv = 0; wv < LB_WARPS; ++wv) {
      if (warp == wv) {
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
          const int v = lane + 32 * i;
          if (v < nv) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float* a = sAcc + v * 8 + j;
              if (wv == 0) *a = S[i][j];
              else *a += S[i][j];
              S[i][j] = 0.f;
            }
          }
        }
      }
      asm volatile("bar.sync 1, %0;\n" ::"n"(LB_WARPS * 32) : "memory");
    }
    for (int c = tid; c < d; c += LB_WARPS * 32) atomicAdd(dgate + sample * dgate_ld + c, sAcc[c]);
    asm volatile("bar.sync 1, %0;\n" ::"n"(LB_WARPS * 32) : "memory");
  };
  for (int64_t t = t0; t < t1; ++t) {
    const int64_t r0 = t * LB_ROWS;
    const int64_t sample = r0 / rows_per_mod;
    if (sample != cur_sample) {
      if (cur_sample >= 0) flush(cur_sample);
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        const int v = lane + 32 * i;
        if (v < nv) gv[i] = ld8(gate + sample * gate_ld + v * 8);
      }
      cur_sample = sample;
    }
    if (tid == 0) ptx::tma_wait_group_read<1>();
    asm volatile("bar.sync 1, %0;\n" ::"n"(LB_WARPS * 32) : "memory");
    ptx::mbar_wait(&full[st], ph);
    const uint8_t* sb = smem + st * stage_bytes;
    bf16* tout = reinterpret_cast<bf16*>(sOut + ob * tile_bytes);
    if (r0 + warp < R) {
      const bf16* dr = reinterpret_cast<const bf16*>(sb) + (size_t)warp * d;
      const bf16* ar = reinterpret_cast<const bf16*>(sb + tile_bytes) + (size_t)warp * d;
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        const int v = lane + 32 * i;
        if (v < nv) {
          float dd[8], a[8], g[8], o[8];
          unpack8(*reinterpret_cast<const bf16x8*>(dr + v * 8), dd);
          unpack8(*reinterpret_cast<const bf16x8*>(ar + v * 8), a);
          unpack8(gv[i], g);
          if constexpr (TWO) {
            float b2[8];
            unpack8(*reinterpret_cast<const bf16x8*>(reinterpret_cast<const bf16*>(sb + 2 * tile_bytes) + (size_t)warp * d + v * 8), b2);
#pragma unroll
            for (int j = 0; j < 8; ++j) a[j] = bf16_round(a[j] + b2[j]);
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            o[j] = dd[j] * g[j];
            S[i][j] = fmaf(dd[j], a[j], S[i][j]);
          }
          *reinterpret_cast<bf16x8*>(tout + (size_t)warp * d + v * 8) = pack8(o);
        }
      }
    }
    __syncwarp();
    if (lane == 0) ptx::mbar_arrive(&empty[st]);
    ptx::fence_proxy_async_smem();
    asm volatile("bar.sync 1, %0;\n" ::"n"(LB_WARPS * 32) : "memory");
    if (tid == 0) {
      const int rows = (int)(R - r0 < LB_ROWS ? R - r0 : LB_ROWS);
      ptx::bulk_store_1d(da + r0 * d, tout, (uint32_t)rows * d * 2);
      ptx::tma_commit_group();
    }
    ob ^= 1;
    if (++st == LB_STAGES) { st = 0; ph ^= 1; }
  }
  if (cur_sample >= 0) flush(cur_sample);
  if (tid == 0) ptx::tma_wait_group<0>();
}

// ---------------------------------------------------------------------------------------------------------
// packed SwiGLU (one vector per thread-iteration; <= 40 registers so 6+ CTAs of 256 threads stay resident)
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 6)
swiglu_fwd_kernel(const bf16* __restrict__ h, bf16* __restrict__ out, int64_t R, int F) {
  const int nv = F >> 3;
  const int64_t total = R * nv;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / nv;
    const int v = (int)(i - row * nv);
    const bf16x8 av = ld8(h + row * 2 * F + v * 8), gv = ld8(h + row * 2 * F + F + v * 8);
    float a[8], g[8], o[8];
    unpack8(av, a);
    unpack8(gv, g);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = bf16_round(silu_fast(a[j])) * g[j];
    st8(out + row * F + v * 8, pack8(o));
  }
}
__global__ void __launch_bounds__(256, 6)
swiglu_bwd_kernel(const bf16* __restrict__ dout, const bf16* __restrict__ h, bf16* __restrict__ dh, int64_t R, int F) {
  const int nv = F >> 3;
  const int64_t total = R * nv;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / nv;
    const int v = (int)(i - row * nv);
    const int64_t ho = row * 2 * F + v * 8;
    const bf16x8 av = ld8(h + ho), gv = ld8(h + ho + F), dv = ld8(dout + row * F + v * 8);
    float a[8], g[8], go[8], da[8], dg[8];
    unpack8(av, a);
    unpack8(gv, g);
    unpack8(dv, go);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float sg = 1.f / (1.f + __expf(-a[j]));
      da[j] = go[j] * g[j] * sg * (1.f + a[j] * (1.f - sg));
      dg[j] = go[j] * a[j] * sg;
    }
    st8(dh + ho, pack8(da));
    st8(dh + ho + F, pack8(dg));
  }
}

int vpl_for(int d) { return (d / 8 + 31) / 32; }
int stream_grid(int64_t total_vec, int per_thread) {
  int64_t blocks = (total_vec + 256 * per_thread - 1) / (256 * per_thread);
  const int64_t cap = (int64_t)dlb_num_sms() * 24;
  return (int)(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}
// blocks of `threads` vectors x rows_per_block rows: aim for ~6 blocks per SM, rows per block a multiple of 8
void cols_grid(int d, int64_t groups, int64_t rows_per_group, int& threads, dim3& grid, int& rpb) {
  const int nv = d / 8;
  threads = nv < 256 ? (nv + 31) / 32 * 32 : 256;
  const int col_chunks = (nv + threads - 1) / threads;
  int64_t want = (int64_t)dlb_num_sms() * 6 / (col_chunks * groups);
  if (want < 1) want = 1;
  int64_t r = (rows_per_group + want - 1) / want;
  r = (r + 7) / 8 * 8;
  if (r < 8) r = 8;
  rpb = (int)r;
  grid = dim3(col_chunks, (unsigned)((rows_per_group + r - 1) / r), (unsigned)groups);
}

}  // namespace

#define VPL_SWITCH(d, ...)                                                                       \
  switch (vpl_for(d)) {                                                                          \
    case 1: { constexpr int VPL = 1; __VA_ARGS__; break; }                                       \
    case 2: { constexpr int VPL = 2; __VA_ARGS__; break; }                                       \
    case 3: { constexpr int VPL = 3; __VA_ARGS__; break; }                                       \
    case 4: { constexpr int VPL = 4; __VA_ARGS__; break; }                                       \
    case 5: { constexpr int VPL = 5; __VA_ARGS__; break; }                                       \
    case 6: { constexpr int VPL = 6; __VA_ARGS__; break; }                                       \
    case 7: case 8: { constexpr int VPL = 8; __VA_ARGS__; break; }                               \
    default: dlb_set_error("channel count %d unsupported (max 2048)", d); return DLB_ERR_SHAPE;  \
  }

DLB_EXPORT int dlb_ln_modulate_fwd(const void* x, const float* w, const float* b, const void* scale, const void* shift,
                                   int64_t mod_ld, int64_t rows_per_mod, void* y, float* mean, float* rstd, int64_t R,
                                   int d, float eps, cudaStream_t stream) {
  DLB_REQUIRE(R > 0 && d > 0 && d % 8 == 0 && mod_ld % 8 == 0, DLB_ERR_SHAPE, "ln_modulate_fwd: R=%lld d=%d mod_ld=%lld",
              (long long)R, d, (long long)mod_ld);
  DLB_REQUIRE((w == nullptr) == (b == nullptr), DLB_ERR_SHAPE, "ln_modulate_fwd: weight and bias must both be set or null");
  DLB_REQUIRE(rows_per_mod >= 1 && (mean == nullptr) == (rstd == nullptr), DLB_ERR_SHAPE, "ln_modulate_fwd: bad args");
  // lean warp-per-row kernel (rowwise_lean.cuh): per-sample modulation, exact per-lane unit split of the channels
  static const bool no_lean = getenv("DLB_NO_LEAN") != nullptr;  // tests: force the general kernels
  if (!no_lean && rows_per_mod % lean::WARPS == 0 && R < (1ll << 31) && ((uintptr_t)x % 16) == 0 && ((uintptr_t)y % 16) == 0) {
    const int rpc = lean::rows_per_cta_for((int)R, dlb_num_sms() * 2);  // 2 CTAs / SM resident, 3 rows in flight per warp
    const int grid_l = (int)((R + rpc - 1) / rpc);
    DLB_LEAN_SWITCH(d, {
      lean::ln_modulate_fwd_lean<U, UPL><<<grid_l, lean::WARPS * 32, 0, stream>>>((const bf16*)x, w, b, (const bf16*)scale, (const bf16*)shift, mod_ld,
                                                                                  (int)rows_per_mod, (bf16*)y, mean, rstd, (int)R, eps, rpc);
      dlb_count_launch();
      return dlb_check_launch("ln_modulate_fwd_lean");
    });
  }
  // tiled bulk-async kernel: per-sample modulation whose groups are whole 8-row tiles, contiguous 16-byte aligned rows
  if (rows_per_mod % LT_ROWS == 0 && ((uintptr_t)x % 16) == 0 && ((uintptr_t)y % 16) == 0) {  // (no size threshold: the kernel choice must not depend on the batch size)
    const int64_t ntiles = (R + LT_ROWS - 1) / LT_ROWS;
    const int max_ctas = dlb_num_sms() * 4;
    const int tiles_per_cta = (int)((ntiles + max_ctas - 1) / max_ctas);
    const int grid_t = (int)((ntiles + tiles_per_cta - 1) / tiles_per_cta);
    const size_t smem = (size_t)(LT_STAGES + 2) * LT_ROWS * d * 2 + 2 * (size_t)d * 4;
    static const int dbg_mode = getenv("DLB_LN_DBG") ? atoi(getenv("DLB_LN_DBG")) : 0;  // 1 = copy-only probe (development)
    VPL_SWITCH(d, {
      static bool attr_set = false;
      if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(ln_modulate_fwd_tile_kernel<VPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 56 * 1024);
        DLB_REQUIRE(e == cudaSuccess, (int)e, "ln_modulate_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        attr_set = true;
      }
      if (smem <= 56 * 1024) {
        ln_modulate_fwd_tile_kernel<VPL><<<grid_t, (LT_WARPS + 1) * 32, smem, stream>>>(
            (const bf16*)x, w, b, (const bf16*)scale, (const bf16*)shift, mod_ld, (int)rows_per_mod, (bf16*)y, mean, rstd, R, d, eps, tiles_per_cta,
            dbg_mode);
        dlb_count_launch();
        return dlb_check_launch("ln_modulate_fwd_tile");
      }
    });
  }
  const int warps = 4;
  const int grid = (int)((R + warps - 1) / warps);
  VPL_SWITCH(d, (ln_modulate_fwd_kernel<VPL><<<grid, warps * 32, 0, stream>>>(
                    (const bf16*)x, w, b, (const bf16*)scale, (const bf16*)shift, mod_ld, (int)rows_per_mod, (bf16*)y,
                    mean, rstd, R, d, eps)));
  dlb_count_launch();
  return dlb_check_launch("ln_modulate_fwd");
}

// groups * rows_per_group rows. Per-sample mode (per_token == 0): dscale/dshift are fp32 rows (stride dmod_ld), one per
// group, that MUST BE ZERO on entry and are owned by this call. Per-token mode: scale is per row, dscale/dshift are
// written per row (bf16) into dscale_tok/dshift_tok and `groups` must be 1. dres (optional) is added to dx.
// dw/db (fp32 [d], accumulated) optional (must be NULL for the affine-free norm).
DLB_EXPORT int dlb_ln_modulate_bwd(const void* dy, const void* x, const float* mean, const float* rstd, const float* w,
                                   const float* b, const void* scale, int64_t mod_ld, int64_t groups,
                                   int64_t rows_per_group, int per_token, const void* dres, void* dx, float* dscale,
                                   float* dshift, int64_t dmod_ld, void* dscale_tok, void* dshift_tok, int64_t dtok_ld,
                                   float* dw, float* db, int d, cudaStream_t stream) {
  DLB_REQUIRE(groups > 0 && rows_per_group > 0 && d > 0 && d % 8 == 0, DLB_ERR_SHAPE, "ln_modulate_bwd: bad shape");
  DLB_REQUIRE((w == nullptr) == (b == nullptr) && (w != nullptr || dw == nullptr) && (dw == nullptr) == (db == nullptr),
              DLB_ERR_SHAPE, "ln_modulate_bwd: inconsistent affine arguments");
  DLB_REQUIRE(!per_token || groups == 1, DLB_ERR_SHAPE, "ln_modulate_bwd: per-token mode takes a single group");
  const int64_t R = groups * rows_per_group;
  const int64_t rows_per_mod = per_token ? 1 : rows_per_group;
  // lean one-pass kernel (rowwise_lean.cuh), then the finalize kernel
  static const bool no_lean = getenv("DLB_NO_LEAN") != nullptr;
  if (!no_lean && !per_token && rows_per_group % lean::WARPS == 0 && R < (1ll << 31) && ((uintptr_t)dy % 16) == 0 && ((uintptr_t)x % 16) == 0 &&
      ((uintptr_t)dx % 16) == 0 && ((uintptr_t)dres % 16) == 0 && ((uintptr_t)dscale % 16) == 0 && ((uintptr_t)dshift % 16) == 0 && dmod_ld % 4 == 0) {
    const int rpc = lean::rows_per_cta_for((int)R, dlb_num_sms());
    const int grid_l = (int)((R + rpc - 1) / rpc);
    const size_t smem_l = (size_t)2 * d * 4 + (size_t)lean::WARPS * lean::LNB_NS * (dres ? 3 : 2) * d * 2;
    if (smem_l <= 220 * 1024) {
      DLB_LEAN_SWITCH(d, {
        if (dres) {
          cudaFuncSetAttribute(lean::ln_modulate_bwd_lean<U, UPL, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_l);
          lean::ln_modulate_bwd_lean<U, UPL, true><<<grid_l, lean::WARPS * 32, smem_l, stream>>>(
              (const bf16*)dy, (const bf16*)x, mean, rstd, w, (const bf16*)scale, mod_ld, (int)rows_per_group, (const bf16*)dres, (bf16*)dx, dshift, dscale,
              dmod_ld, (int)R, rpc);
        } else {
          cudaFuncSetAttribute(lean::ln_modulate_bwd_lean<U, UPL, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_l);
          lean::ln_modulate_bwd_lean<U, UPL, false><<<grid_l, lean::WARPS * 32, smem_l, stream>>>(
              (const bf16*)dy, (const bf16*)x, mean, rstd, w, (const bf16*)scale, mod_ld, (int)rows_per_group, nullptr, (bf16*)dx, dshift, dscale, dmod_ld,
              (int)R, rpc);
        }
        dlb_count_launch();
        int rcl = dlb_check_launch("ln_modulate_bwd_lean");
        if (rcl) return rcl;
        const int gpb = 8;
        dim3 fgrid((d + 255) / 256, (unsigned)((groups + gpb - 1) / gpb));
        ln_modulate_bwd_finalize_kernel<<<fgrid, 256, 0, stream>>>(w, b, (const bf16*)scale, mod_ld, dscale, dshift, dmod_ld, dw, db, d, groups, gpb);
        dlb_count_launch();
        return dlb_check_launch("ln_modulate_bwd_finalize");
      });
    }
  }
  // one-pass tiled kernel (dx + column sums), then the finalize kernel
  if (!per_token && rows_per_group % LB_ROWS == 0 && ((uintptr_t)dy % 16) == 0 && ((uintptr_t)x % 16) == 0 && ((uintptr_t)dx % 16) == 0 &&
      (dres == nullptr || ((uintptr_t)dres % 16) == 0) && ((uintptr_t)mean % 16) == 0 && ((uintptr_t)rstd % 16) == 0) {
    const int nin = dres ? 3 : 2;
    const size_t smem_t = (size_t)LB_STAGES * (nin * LB_ROWS * d * 2 + 32) + 2 * (size_t)LB_ROWS * d * 2 + 3 * (size_t)d * 4;
    if (smem_t <= 113 * 1024) {
      const int64_t ntiles = R / LB_ROWS;
      const int max_ctas = dlb_num_sms() * 2;
      const int tiles_per_cta = (int)((ntiles + max_ctas - 1) / max_ctas);
      const int grid_t = (int)((ntiles + tiles_per_cta - 1) / tiles_per_cta);
      VPL_SWITCH(d, {
        if (dres) {
          cudaFuncSetAttribute(ln_modulate_bwd_tile_kernel<VPL, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024);
          ln_modulate_bwd_tile_kernel<VPL, true><<<grid_t, (LB_WARPS + 1) * 32, smem_t, stream>>>(
              (const bf16*)dy, (const bf16*)x, mean, rstd, w, (const bf16*)scale, mod_ld, (int)rows_per_group, (const bf16*)dres, (bf16*)dx,
              dshift, dscale, dmod_ld, R, d, tiles_per_cta);
        } else {
          cudaFuncSetAttribute(ln_modulate_bwd_tile_kernel<VPL, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024);
          ln_modulate_bwd_tile_kernel<VPL, false><<<grid_t, (LB_WARPS + 1) * 32, smem_t, stream>>>(
              (const bf16*)dy, (const bf16*)x, mean, rstd, w, (const bf16*)scale, mod_ld, (int)rows_per_group, nullptr, (bf16*)dx,
              dshift, dscale, dmod_ld, R, d, tiles_per_cta);
        }
      });
      dlb_count_launch();
      int rcl = dlb_check_launch("ln_modulate_bwd_tile");
      if (rcl) return rcl;
      const int gpb = 8;
      dim3 fgrid((d + 255) / 256, (unsigned)((groups + gpb - 1) / gpb));
      ln_modulate_bwd_finalize_kernel<<<fgrid, 256, 0, stream>>>(w, b, (const bf16*)scale, mod_ld, dscale, dshift, dmod_ld, dw, db, d, groups, gpb);
      dlb_count_launch();
      return dlb_check_launch("ln_modulate_bwd_finalize");
    }
  }
  const int warps = 4;
  const int grid_rows = (int)((R + warps - 1) / warps);
  if (per_token) {
    VPL_SWITCH(d, (ln_modulate_bwd_rows_kernel<VPL, true><<<grid_rows, warps * 32, 0, stream>>>(
                      (const bf16*)dy, (const bf16*)x, mean, rstd, w, b, (const bf16*)scale, mod_ld, rows_per_mod,
                      (const bf16*)dres, (bf16*)dx, (bf16*)dscale_tok, (bf16*)dshift_tok, dtok_ld, R, d)));
  } else {
    VPL_SWITCH(d, (ln_modulate_bwd_rows_kernel<VPL, false><<<grid_rows, warps * 32, 0, stream>>>(
                      (const bf16*)dy, (const bf16*)x, mean, rstd, w, b, (const bf16*)scale, mod_ld, rows_per_mod,
                      (const bf16*)dres, (bf16*)dx, nullptr, nullptr, 0, R, d)));
  }
  dlb_count_launch();
  int threads, rpb;
  dim3 grid;
  cols_grid(d, groups, rows_per_group, threads, grid, rpb);
  if (per_token) {
    if (dw != nullptr) {
      ln_modulate_bwd_cols_kernel<true><<<grid, threads, 0, stream>>>((const bf16*)dy, (const bf16*)x, mean, rstd,
                                                                     (const bf16*)scale, mod_ld, rows_per_group, db, dw, 0, d, rpb);
      dlb_count_launch();
    }
  } else {
    ln_modulate_bwd_cols_kernel<false><<<grid, threads, 0, stream>>>((const bf16*)dy, (const bf16*)x, mean, rstd,
                                                                    (const bf16*)scale, mod_ld, rows_per_group, dshift, dscale,
                                                                    dmod_ld, d, rpb);
    const int gpb = 16;
    dim3 fgrid((d + 255) / 256, (unsigned)((groups + gpb - 1) / gpb));
    ln_modulate_bwd_finalize_kernel<<<fgrid, 256, 0, stream>>>(w, b, (const bf16*)scale, mod_ld, dscale, dshift, dmod_ld, dw, db,
                                                               d, groups, gpb);
    dlb_count_launch(2);
  }
  return dlb_check_launch("ln_modulate_bwd");
}

DLB_EXPORT int dlb_gate_residual_fwd(const void* x, const void* a1, const void* a2, const void* gate, int64_t gate_ld,
                                     int64_t rows_per_mod, void* out, int64_t R, int d, cudaStream_t stream) {
  DLB_REQUIRE(R > 0 && d > 0 && d % 8 == 0 && gate_ld % 8 == 0 && rows_per_mod >= 1, DLB_ERR_SHAPE,
              "gate_residual_fwd: bad shape R=%lld d=%d", (long long)R, d);
  static const bool no_lean = getenv("DLB_NO_LEAN") != nullptr;
  if (!no_lean && R < (1ll << 31) && ((uintptr_t)x % 16) == 0 && ((uintptr_t)a1 % 16) == 0 && ((uintptr_t)a2 % 16) == 0 && ((uintptr_t)out % 16) == 0 &&
      ((uintptr_t)gate % 16) == 0 && gate_ld % 8 == 0) {
    const int rpc = lean::rows_per_cta_for((int)R, dlb_num_sms() * 3);
    const int grid_l = (int)((R + rpc - 1) / rpc);
    DLB_LEAN_SWITCH(d, {
      if (a2)
        lean::gate_residual_fwd_lean<U, UPL, true><<<grid_l, lean::WARPS * 32, 0, stream>>>((const bf16*)x, (const bf16*)a1, (const bf16*)a2, (const bf16*)gate,
                                                                                          gate_ld, (int)rows_per_mod, (bf16*)out, (int)R, rpc);
      else
        lean::gate_residual_fwd_lean<U, UPL, false><<<grid_l, lean::WARPS * 32, 0, stream>>>((const bf16*)x, (const bf16*)a1, nullptr, (const bf16*)gate,
                                                                                           gate_ld, (int)rows_per_mod, (bf16*)out, (int)R, rpc);
      dlb_count_launch();
      return dlb_check_launch("gate_residual_fwd_lean");
    });
  }
  const int g = stream_grid(R * (d / 8), 1);
  if (a2)
    gate_residual_fwd_kernel<true><<<g, 256, 0, stream>>>((const bf16*)x, (const bf16*)a1, (const bf16*)a2, (const bf16*)gate,
                                                         gate_ld, (int)rows_per_mod, (bf16*)out, R, d);
  else
    gate_residual_fwd_kernel<false><<<g, 256, 0, stream>>>((const bf16*)x, (const bf16*)a1, nullptr, (const bf16*)gate, gate_ld,
                                                          (int)rows_per_mod, (bf16*)out, R, d);
  dlb_count_launch();
  return dlb_check_launch("gate_residual_fwd");
}

// da = dout * gate; per-sample mode: dgate fp32 rows (stride dgate_ld) accumulated; per-token: dgate_tok bf16 rows written
DLB_EXPORT int dlb_gate_residual_bwd(const void* dout, const void* a1, const void* a2, const void* gate, int64_t gate_ld,
                                     int64_t groups, int64_t rows_per_group, int per_token, void* da, float* dgate,
                                     int64_t dgate_ld, void* dgate_tok, int64_t dtok_ld, int d, cudaStream_t stream) {
  DLB_REQUIRE(groups > 0 && rows_per_group > 0 && d > 0 && d % 8 == 0, DLB_ERR_SHAPE, "gate_residual_bwd: bad shape");
  DLB_REQUIRE(!per_token || groups == 1, DLB_ERR_SHAPE, "gate_residual_bwd: per-token mode takes a single group");
  const int64_t R = groups * rows_per_group;
  const int64_t rows_per_mod = per_token ? 1 : rows_per_group;
  static const bool no_lean_b = getenv("DLB_NO_LEAN") != nullptr;
  if (!no_lean_b && !per_token && rows_per_group % lean::WARPS == 0 && R < (1ll << 31) && ((uintptr_t)dout % 16) == 0 && ((uintptr_t)a1 % 16) == 0 &&
      ((uintptr_t)a2 % 16) == 0 && ((uintptr_t)da % 16) == 0 && ((uintptr_t)gate % 16) == 0 && gate_ld % 8 == 0 && ((uintptr_t)dgate % 16) == 0 &&
      dgate_ld % 4 == 0) {
    const int rpc = lean::rows_per_cta_for((int)R, dlb_num_sms() * 2);
    const int grid_l = (int)((R + rpc - 1) / rpc);
    DLB_LEAN_SWITCH(d, {
      if (a2)
        lean::gate_residual_bwd_lean<U, UPL, true><<<grid_l, lean::WARPS * 32, 0, stream>>>((const bf16*)dout, (const bf16*)a1, (const bf16*)a2, (const bf16*)gate,
                                                                                          gate_ld, (int)rows_per_group, (bf16*)da, dgate, dgate_ld, (int)R, rpc);
      else
        lean::gate_residual_bwd_lean<U, UPL, false><<<grid_l, lean::WARPS * 32, 0, stream>>>((const bf16*)dout, (const bf16*)a1, nullptr, (const bf16*)gate,
                                                                                           gate_ld, (int)rows_per_group, (bf16*)da, dgate, dgate_ld, (int)R, rpc);
      dlb_count_launch();
      return dlb_check_launch("gate_residual_bwd_lean");
    });
  }
  if (!per_token && rows_per_group % LB_ROWS == 0 && ((uintptr_t)dout % 16) == 0 && ((uintptr_t)a1 % 16) == 0 &&
      (a2 == nullptr || ((uintptr_t)a2 % 16) == 0) && ((uintptr_t)da % 16) == 0) {
    const int nin = a2 ? 3 : 2;
    const size_t smem_t = (size_t)LB_STAGES * nin * LB_ROWS * d * 2 + 2 * (size_t)LB_ROWS * d * 2 + (size_t)d * 4;
    if (smem_t <= 113 * 1024) {
      const int64_t ntiles = R / LB_ROWS;
      const int max_ctas = dlb_num_sms() * 2;
      const int tiles_per_cta = (int)((ntiles + max_ctas - 1) / max_ctas);
      const int grid_t = (int)((ntiles + tiles_per_cta - 1) / tiles_per_cta);
      VPL_SWITCH(d, {
        if (a2) {
          cudaFuncSetAttribute(gate_residual_bwd_tile_kernel<VPL, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024);
          gate_residual_bwd_tile_kernel<VPL, true><<<grid_t, (LB_WARPS + 1) * 32, smem_t, stream>>>(
              (const bf16*)dout, (const bf16*)a1, (const bf16*)a2, (const bf16*)gate, gate_ld, (int)rows_per_group, (bf16*)da, dgate, dgate_ld, R, d, tiles_per_cta);
        } else {
          cudaFuncSetAttribute(gate_residual_bwd_tile_kernel<VPL, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024);
          gate_residual_bwd_tile_kernel<VPL, false><<<grid_t, (LB_WARPS + 1) * 32, smem_t, stream>>>(
              (const bf16*)dout, (const bf16*)a1, nullptr, (const bf16*)gate, gate_ld, (int)rows_per_group, (bf16*)da, dgate, dgate_ld, R, d, tiles_per_cta);
        }
      });
      dlb_count_launch();
      return dlb_check_launch("gate_residual_bwd_tile");
    }
  }
  const int g = stream_grid(R * (d / 8), 1);
#define GATE_BWD_STREAM(TWO, PT)                                                                                         \
  gate_residual_bwd_stream_kernel<TWO, PT><<<g, 256, 0, stream>>>((const bf16*)dout, (const bf16*)a1, (const bf16*)a2,  \
                                                                  (const bf16*)gate, gate_ld, rows_per_mod, (bf16*)da,   \
                                                                  (bf16*)dgate_tok, dtok_ld, R, d)
  if (per_token) { if (a2) GATE_BWD_STREAM(true, true); else GATE_BWD_STREAM(false, true); }
  else { if (a2) GATE_BWD_STREAM(true, false); else GATE_BWD_STREAM(false, false); }
#undef GATE_BWD_STREAM
  dlb_count_launch();
  if (!per_token) {
    int threads, rpb;
    dim3 grid;
    cols_grid(d, groups, rows_per_group, threads, grid, rpb);
    if (a2)
      gate_residual_bwd_cols_kernel<true><<<grid, threads, 0, stream>>>((const bf16*)dout, (const bf16*)a1, (const bf16*)a2,
                                                                       rows_per_group, dgate, dgate_ld, d, rpb);
    else
      gate_residual_bwd_cols_kernel<false><<<grid, threads, 0, stream>>>((const bf16*)dout, (const bf16*)a1, nullptr,
                                                                        rows_per_group, dgate, dgate_ld, d, rpb);
    dlb_count_launch();
  }
  return dlb_check_launch("gate_residual_bwd");
}

DLB_EXPORT int dlb_swiglu_fwd(const void* h, void* out, int64_t R, int F, cudaStream_t stream) {
  DLB_REQUIRE(R > 0 && F > 0 && F % 8 == 0, DLB_ERR_SHAPE, "swiglu_fwd: bad shape R=%lld F=%d", (long long)R, F);
  swiglu_fwd_kernel<<<stream_grid(R * (F / 8), 1), 256, 0, stream>>>((const bf16*)h, (bf16*)out, R, F);
  dlb_count_launch();
  return dlb_check_launch("swiglu_fwd");
}
DLB_EXPORT int dlb_swiglu_bwd(const void* dout, const void* h, void* dh, int64_t R, int F, cudaStream_t stream) {
  DLB_REQUIRE(R > 0 && F > 0 && F % 8 == 0, DLB_ERR_SHAPE, "swiglu_bwd: bad shape R=%lld F=%d", (long long)R, F);
  swiglu_bwd_kernel<<<stream_grid(R * (F / 8), 1), 256, 0, stream>>>((const bf16*)dout, (const bf16*)h, (bf16*)dh, R, F);
  dlb_count_launch();
  return dlb_check_launch("swiglu_bwd");
}
