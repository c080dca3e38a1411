// Fused LayerNorm + adaLN modulate, gated residual, packed SwiGLU: the bandwidth-bound token-stream kernels.
// One warp owns one token row (d <= 2048 channels kept in registers as 16-byte bf16x8 vectors); statistics
// via warp shuffles; every global access is a 128-bit coalesced vector.
//
// Reference semantics (bf16 autocast on CUDA):
//   ln_modulate : modulate(nn.LayerNorm(x), scale, shift)      mmdit.py:257-259,299,305 ; nn.py:539-540
//   gate_res    : x + branch * gate                            mmdit.py:296-307
//   swiglu      : silu(x1) * x3 on the packed up-projection    nn.py:478-486
#include "common.cuh"

namespace {

typedef __nv_bfloat16 bf16;

// ---------------------------------------------------------------------------------------------------------
// LayerNorm (+affine) + modulate forward
// ---------------------------------------------------------------------------------------------------------
template <int VPL>
__global__ void __launch_bounds__(256)
ln_modulate_fwd_kernel(const bf16* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                       const bf16* __restrict__ scale, const bf16* __restrict__ shift, int64_t mod_ld,
                       int rows_per_mod, bf16* __restrict__ y, float* __restrict__ mean_out,
                       float* __restrict__ rstd_out, int64_t R, int d, float eps) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= R) return;
  const int nv = d >> 3;
  const bf16* xr = x + row * d;
  float xv[VPL][8];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int v = lane + 32 * i;
    if (v < nv) {
      unpack8(ld8(xr + v * 8), xv[i]);
#pragma unroll
      for (int j = 0; j < 8; ++j) sum += xv[i][j];
    }
  }
  const float mean = warp_sum(sum) / d;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    if (lane + 32 * i < nv) {
#pragma unroll
      for (int j = 0; j < 8; ++j) { const float t = xv[i][j] - mean; sq += t * t; }
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
      float s[8], t[8], o[8];
      unpack8(ld8(sc + v * 8), s);
      unpack8(ld8(sh + v * 8), t);
      float wv[8], bv[8];
      if (w) {
        *reinterpret_cast<float4*>(wv) = __ldg(reinterpret_cast<const float4*>(w + v * 8));
        *reinterpret_cast<float4*>(wv + 4) = __ldg(reinterpret_cast<const float4*>(w + v * 8 + 4));
        *reinterpret_cast<float4*>(bv) = __ldg(reinterpret_cast<const float4*>(b + v * 8));
        *reinterpret_cast<float4*>(bv + 4) = __ldg(reinterpret_cast<const float4*>(b + v * 8 + 4));
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float u = (xv[i][j] - mean) * rstd;
        if (w) u = u * wv[j] + bv[j];
        // reference: `1 + scale` is evaluated in bf16 before meeting the fp32 LayerNorm output
        o[j] = u * bf16_round(1.f + s[j]) + t[j];
      }
      st8(yr + v * 8, pack8(o));
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// LayerNorm + modulate backward, two bandwidth-shaped kernels:
//  (1) rows: one warp per token row -> dx (+dres); per-token mode also writes dscale/dshift rows.
//  (2) cols: one thread per 8-channel vector, marching down a chunk of rows -> two column accumulators
//        per-sample mode : S1 = sum dy           S2 = sum dy * xhat
//        per-token  mode : S1 = sum dy*(1+scale) S2 = sum dy*(1+scale)*xhat
//      which are enough for dshift, dscale, dw and db ("LN backward algebra"); one atomic per column
//      per block. Keeping the column sums out of kernel (1) keeps it at ~100 registers (2+ CTAs per SM).
// ---------------------------------------------------------------------------------------------------------
template <int VPL, bool PER_TOKEN>
__global__ void __launch_bounds__(128)
ln_modulate_bwd_rows_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x, const float* __restrict__ mean_in,
                            const float* __restrict__ rstd_in, const float* __restrict__ w, const float* __restrict__ b,
                            const bf16* __restrict__ scale, int64_t mod_ld, int64_t rows_per_mod,
                            const bf16* __restrict__ dres, bf16* __restrict__ dx, bf16* __restrict__ dscale_tok,
                            bf16* __restrict__ dshift_tok, int64_t dtok_ld, int64_t R, int d) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= R) return;
  const int nv = d >> 3;
  const float mean = mean_in[row], rstd = rstd_in[row];
  const bf16* sc = scale + (row / rows_per_mod) * mod_ld;
  float xh[VPL][8], gw[VPL][8];
  float m1 = 0.f, m2 = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int v = lane + 32 * i;
    if (v < nv) {
      float g[8], s[8];
      unpack8(ld8(x + row * d + v * 8), xh[i]);
      unpack8(ld8(dy + row * d + v * 8), g);
      unpack8(ld8(sc + v * 8), s);
      float wv[8], bv[8];
      if (w) {
        *reinterpret_cast<float4*>(wv) = __ldg(reinterpret_cast<const float4*>(w + v * 8));
        *reinterpret_cast<float4*>(wv + 4) = __ldg(reinterpret_cast<const float4*>(w + v * 8 + 4));
        if (PER_TOKEN) {
          *reinterpret_cast<float4*>(bv) = __ldg(reinterpret_cast<const float4*>(b + v * 8));
          *reinterpret_cast<float4*>(bv + 4) = __ldg(reinterpret_cast<const float4*>(b + v * 8 + 4));
        }
      }
      float ds_tok[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        xh[i][j] = (xh[i][j] - mean) * rstd;
        const float gu = g[j] * bf16_round(1.f + s[j]);  // grad wrt the affine LayerNorm output u
        if (PER_TOKEN) ds_tok[j] = g[j] * (w ? xh[i][j] * wv[j] + bv[j] : xh[i][j]);
        gw[i][j] = w ? gu * wv[j] : gu;
        m1 += gw[i][j];
        m2 += gw[i][j] * xh[i][j];
      }
      if (PER_TOKEN) {
        st8(dscale_tok + row * dtok_ld + v * 8, pack8(ds_tok));
        st8(dshift_tok + row * dtok_ld + v * 8, pack8(g));
      }
    }
  }
  m1 = warp_sum(m1) / d;
  m2 = warp_sum(m2) / d;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int v = lane + 32 * i;
    if (v < nv) {
      float o[8];
      if (dres) unpack8(ld8(dres + row * d + v * 8), o);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float t = rstd * (gw[i][j] - m1 - xh[i][j] * m2);
        o[j] = dres ? o[j] + t : t;
      }
      st8(dx + row * d + v * 8, pack8(o));
    }
  }
}

// grid (col chunks, row chunks, groups); thread = one 8-channel vector, marching down its row chunk
template <bool PER_TOKEN>
__global__ void __launch_bounds__(256)
ln_modulate_bwd_cols_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x, const float* __restrict__ mean_in,
                            const float* __restrict__ rstd_in, const float* __restrict__ w, const float* __restrict__ b,
                            const bf16* __restrict__ scale, int64_t mod_ld, int64_t rows_per_group,
                            float* __restrict__ dscale, float* __restrict__ dshift, int64_t dmod_ld,
                            float* __restrict__ dw, float* __restrict__ db, int d, int rows_per_block) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= (d >> 3)) return;
  const int c = v * 8;
  const int64_t group = blockIdx.z;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_block;
  const int64_t r1 = min(r0 + rows_per_block, rows_per_group);
  float S1[8], S2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { S1[j] = 0.f; S2[j] = 0.f; }
  const int64_t base = group * rows_per_group;
#pragma unroll 4
  for (int64_t rl = r0; rl < r1; ++rl) {
    const int64_t row = base + rl;
    float xv[8], g[8];
    unpack8(ld8(x + row * d + c), xv);
    unpack8(ld8(dy + row * d + c), g);
    const float mean = __ldg(mean_in + row), rstd = __ldg(rstd_in + row);
    if (PER_TOKEN) {
      float s[8];
      unpack8(ld8(scale + row * mod_ld + c), s);
#pragma unroll
      for (int j = 0; j < 8; ++j) g[j] *= bf16_round(1.f + s[j]);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      S1[j] += g[j];
      S2[j] += g[j] * (xv[j] - mean) * rstd;
    }
  }
  if (PER_TOKEN) {
    if (dw) {
#pragma unroll
      for (int j = 0; j < 8; ++j) { atomicAdd(dw + c + j, S2[j]); atomicAdd(db + c + j, S1[j]); }
    }
  } else {
    float s[8];
    unpack8(ld8(scale + group * mod_ld + c), s);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float wc = w ? __ldg(w + c + j) : 1.f, bc = w ? __ldg(b + c + j) : 0.f;
      atomicAdd(dshift + group * dmod_ld + c + j, S1[j]);
      atomicAdd(dscale + group * dmod_ld + c + j, wc * S2[j] + bc * S1[j]);
      if (dw) {
        const float one_s = bf16_round(1.f + s[j]);
        atomicAdd(dw + c + j, one_s * S2[j]);
        atomicAdd(db + c + j, one_s * S1[j]);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// gated residual: out = x + (a1 [+ a2]) * gate        (bf16 roundings placed where the reference rounds)
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gate_residual_fwd_kernel(const bf16* __restrict__ x, const bf16* __restrict__ a1, const bf16* __restrict__ a2,
                         const bf16* __restrict__ gate, int64_t gate_ld, int rows_per_mod, bf16* __restrict__ out,
                         int64_t R, int d) {
  const int nv = d >> 3;
  const int64_t total = R * nv;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / nv;
    const int v = (int)(i - row * nv);
    float xv[8], av[8], gv[8], o[8];
    unpack8(ld8(x + row * d + v * 8), xv);
    unpack8(ld8(a1 + row * d + v * 8), av);
    if (a2) {
      float bv[8];
      unpack8(ld8(a2 + row * d + v * 8), bv);
#pragma unroll
      for (int j = 0; j < 8; ++j) av[j] = bf16_round(av[j] + bv[j]);
    }
    unpack8(ld8(gate + (row / rows_per_mod) * gate_ld + v * 8), gv);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = xv[j] + bf16_round(av[j] * gv[j]);
    st8(out + row * d + v * 8, pack8(o));
  }
}

// backward: da = dout * gate ; dgate[group] += sum_rows dout * a      (dx = dout is an alias, no kernel)
template <int VPL, bool PER_TOKEN>
__global__ void __launch_bounds__(256)
gate_residual_bwd_kernel(const bf16* __restrict__ dout, const bf16* __restrict__ a1, const bf16* __restrict__ a2,
                         const bf16* __restrict__ gate, int64_t gate_ld, int64_t rows_per_group,
                         bf16* __restrict__ da, float* __restrict__ dgate, int64_t dgate_ld,
                         bf16* __restrict__ dgate_tok, int64_t dtok_ld, int d, int rows_per_warp) {
  extern __shared__ float red[];  // [warps][d]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int nv = d >> 3;
  const int64_t group = blockIdx.y;
  const int64_t r_begin = ((int64_t)blockIdx.x * nwarps + warp) * rows_per_warp;
  const int64_t r_end = min(r_begin + rows_per_warp, rows_per_group);
  float S[VPL][8];
#pragma unroll
  for (int i = 0; i < VPL; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) S[i][j] = 0.f;
  for (int64_t rl = r_begin; rl < r_end; ++rl) {
    const int64_t row = group * rows_per_group + rl;
    const bf16* gp = gate + (PER_TOKEN ? row : group) * gate_ld;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int v = lane + 32 * i;
      if (v < nv) {
        float g[8], av[8], gv[8], o[8], dg[8];
        unpack8(ld8(dout + row * d + v * 8), g);
        unpack8(ld8(a1 + row * d + v * 8), av);
        if (a2) {
          float bv[8];
          unpack8(ld8(a2 + row * d + v * 8), bv);
#pragma unroll
          for (int j = 0; j < 8; ++j) av[j] = bf16_round(av[j] + bv[j]);
        }
        unpack8(ld8(gp + v * 8), gv);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          o[j] = g[j] * gv[j];
          dg[j] = g[j] * av[j];
          S[i][j] += dg[j];
        }
        st8(da + row * d + v * 8, pack8(o));
        if (PER_TOKEN) st8(dgate_tok + row * dtok_ld + v * 8, pack8(dg));
      }
    }
  }
  if (PER_TOKEN) return;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int v = lane + 32 * i;
    if (v < nv) {
#pragma unroll
      for (int j = 0; j < 8; ++j) red[(size_t)warp * d + v * 8 + j] = S[i][j];
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    float a = 0.f;
    for (int wi = 0; wi < nwarps; ++wi) a += red[(size_t)wi * d + c];
    atomicAdd(dgate + group * dgate_ld + c, a);
  }
}

// ---------------------------------------------------------------------------------------------------------
// packed SwiGLU
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
swiglu_fwd_kernel(const bf16* __restrict__ h, bf16* __restrict__ out, int64_t R, int F) {
  const int nv = F >> 3;
  const int64_t total = R * nv;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / nv;
    const int v = (int)(i - row * nv);
    float a[8], g[8], o[8];
    unpack8(ld8(h + row * 2 * F + v * 8), a);
    unpack8(ld8(h + row * 2 * F + F + v * 8), g);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = bf16_round(silu_f(a[j])) * g[j];
    st8(out + row * F + v * 8, pack8(o));
  }
}
__global__ void __launch_bounds__(256)
swiglu_bwd_kernel(const bf16* __restrict__ dout, const bf16* __restrict__ h, bf16* __restrict__ dh, int64_t R, int F) {
  const int nv = F >> 3;
  const int64_t total = R * nv;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / nv;
    const int v = (int)(i - row * nv);
    float a[8], g[8], go[8], da[8], dg[8];
    unpack8(ld8(h + row * 2 * F + v * 8), a);
    unpack8(ld8(h + row * 2 * F + F + v * 8), g);
    unpack8(ld8(dout + row * F + v * 8), go);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      da[j] = go[j] * g[j] * dsilu_f(a[j]);
      dg[j] = go[j] * silu_f(a[j]);
    }
    st8(dh + row * 2 * F + v * 8, pack8(da));
    st8(dh + row * 2 * F + F + v * 8, pack8(dg));
  }
}

int vpl_for(int d) { return (d / 8 + 31) / 32; }
int ew_grid(int64_t total_vec) {
  int64_t blocks = (total_vec + 255) / 256;
  const int64_t cap = (int64_t)dlb_num_sms() * 16;
  return (int)(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
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
  const int warps = 8;
  const int grid = (int)((R + warps - 1) / warps);
  VPL_SWITCH(d, (ln_modulate_fwd_kernel<VPL><<<grid, warps * 32, 0, stream>>>(
                    (const bf16*)x, w, b, (const bf16*)scale, (const bf16*)shift, mod_ld, (int)rows_per_mod, (bf16*)y,
                    mean, rstd, R, d, eps)));
  dlb_count_launch();
  return dlb_check_launch("ln_modulate_fwd");
}

// groups * rows_per_group rows. per_token != 0: scale is per row, dscale/dshift are written per row (bf16) into
// dscale_tok/dshift_tok and `groups` must be 1. dres (optional) is added to dx. dw/db optional (non-affine LN).
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
  if (!per_token || dw != nullptr) {
    const int nv = d / 8;
    int threads = nv < 256 ? (nv + 31) / 32 * 32 : 256;
    const int col_chunks = (nv + threads - 1) / threads;
    // ~4 blocks per SM in total, at least 16 rows per block so the final atomics stay a small fraction
    int64_t want_blocks = (int64_t)dlb_num_sms() * 4 / (col_chunks * groups);
    if (want_blocks < 1) want_blocks = 1;
    int64_t rpb = (rows_per_group + want_blocks - 1) / want_blocks;
    if (rpb < 16) rpb = 16;
    dim3 grid(col_chunks, (unsigned)((rows_per_group + rpb - 1) / rpb), (unsigned)groups);
    if (per_token)
      ln_modulate_bwd_cols_kernel<true><<<grid, threads, 0, stream>>>((const bf16*)dy, (const bf16*)x, mean, rstd, w, b,
                                                                     (const bf16*)scale, mod_ld, rows_per_group, dscale,
                                                                     dshift, dmod_ld, dw, db, d, (int)rpb);
    else
      ln_modulate_bwd_cols_kernel<false><<<grid, threads, 0, stream>>>((const bf16*)dy, (const bf16*)x, mean, rstd, w, b,
                                                                      (const bf16*)scale, mod_ld, rows_per_group, dscale,
                                                                      dshift, dmod_ld, dw, db, d, (int)rpb);
    dlb_count_launch();
  }
  dlb_count_launch();
  return dlb_check_launch("ln_modulate_bwd");
}

DLB_EXPORT int dlb_gate_residual_fwd(const void* x, const void* a1, const void* a2, const void* gate, int64_t gate_ld,
                                     int64_t rows_per_mod, void* out, int64_t R, int d, cudaStream_t stream) {
  DLB_REQUIRE(R > 0 && d > 0 && d % 8 == 0 && gate_ld % 8 == 0 && rows_per_mod >= 1, DLB_ERR_SHAPE,
              "gate_residual_fwd: bad shape R=%lld d=%d", (long long)R, d);
  gate_residual_fwd_kernel<<<ew_grid(R * (d / 8)), 256, 0, stream>>>((const bf16*)x, (const bf16*)a1, (const bf16*)a2,
                                                                    (const bf16*)gate, gate_ld, (int)rows_per_mod,
                                                                    (bf16*)out, R, d);
  dlb_count_launch();
  return dlb_check_launch("gate_residual_fwd");
}

DLB_EXPORT int dlb_gate_residual_bwd(const void* dout, const void* a1, const void* a2, const void* gate, int64_t gate_ld,
                                     int64_t groups, int64_t rows_per_group, int per_token, void* da, float* dgate,
                                     int64_t dgate_ld, void* dgate_tok, int64_t dtok_ld, int d, cudaStream_t stream) {
  DLB_REQUIRE(groups > 0 && rows_per_group > 0 && d > 0 && d % 8 == 0, DLB_ERR_SHAPE, "gate_residual_bwd: bad shape");
  DLB_REQUIRE(!per_token || groups == 1, DLB_ERR_SHAPE, "gate_residual_bwd: per-token mode takes a single group");
  const int warps = 8;
  int64_t rpw = rows_per_group * groups / ((int64_t)dlb_num_sms() * 4 * warps);
  rpw = rpw < 1 ? 1 : (rpw > 8 ? 8 : rpw);
  dim3 grid((unsigned)((rows_per_group + warps * rpw - 1) / (warps * rpw)), (unsigned)groups);
  const size_t smem = (size_t)warps * d * sizeof(float);
  if (per_token) {
    VPL_SWITCH(d, (gate_residual_bwd_kernel<VPL, true><<<grid, warps * 32, smem, stream>>>(
                      (const bf16*)dout, (const bf16*)a1, (const bf16*)a2, (const bf16*)gate, gate_ld, rows_per_group,
                      (bf16*)da, dgate, dgate_ld, (bf16*)dgate_tok, dtok_ld, d, (int)rpw)));
  } else {
    VPL_SWITCH(d, {
      auto k = gate_residual_bwd_kernel<VPL, false>;
      if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      k<<<grid, warps * 32, smem, stream>>>((const bf16*)dout, (const bf16*)a1, (const bf16*)a2, (const bf16*)gate,
                                           gate_ld, rows_per_group, (bf16*)da, dgate, dgate_ld, (bf16*)dgate_tok,
                                           dtok_ld, d, (int)rpw);
    });
  }
  dlb_count_launch();
  return dlb_check_launch("gate_residual_bwd");
}

DLB_EXPORT int dlb_swiglu_fwd(const void* h, void* out, int64_t R, int F, cudaStream_t stream) {
  DLB_REQUIRE(R > 0 && F > 0 && F % 8 == 0, DLB_ERR_SHAPE, "swiglu_fwd: bad shape R=%lld F=%d", (long long)R, F);
  swiglu_fwd_kernel<<<ew_grid(R * (F / 8)), 256, 0, stream>>>((const bf16*)h, (bf16*)out, R, F);
  dlb_count_launch();
  return dlb_check_launch("swiglu_fwd");
}
DLB_EXPORT int dlb_swiglu_bwd(const void* dout, const void* h, void* dh, int64_t R, int F, cudaStream_t stream) {
  DLB_REQUIRE(R > 0 && F > 0 && F % 8 == 0, DLB_ERR_SHAPE, "swiglu_bwd: bad shape R=%lld F=%d", (long long)R, F);
  swiglu_bwd_kernel<<<ew_grid(R * (F / 8)), 256, 0, stream>>>((const bf16*)dout, (const bf16*)h, (bf16*)dh, R, F);
  dlb_count_launch();
  return dlb_check_launch("swiglu_bwd");
}
