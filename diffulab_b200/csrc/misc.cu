// Small glue kernels of the denoiser: dtype casts, SiLU, sinusoidal timestep embedding, label-embedding add,
// patchify (im2col for the k=s=p convolution) / unpatchify, column sums (bias gradients), embedding scatter-add.
//
// Reference: timestep_embedding networks/utils/nn.py:91-114; LabelEmbed nn.py:117-164; patchify / unpatchify
// denoisers/mmdit.py:747-787 (Conv2d(k=p, s=p, bias=False) + "b c h w -> b (h w) c"; inverse rearrange
// "b (h w) (p1 p2 c) -> b c (h p1) (w p2)").
#include "common.cuh"

namespace {
typedef __nv_bfloat16 bf16;

int grid_for(int64_t n, int per_thread = 1) {
  int64_t blocks = (n + 256 * per_thread - 1) / (256 * per_thread);
  const int64_t cap = (int64_t)dlb_num_sms() * 16;
  return (int)(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}

// fp32 -> bf16 cast of a [rows, cols] matrix into a [rows, ld_out] buffer (ld_out >= cols; padding zero-filled).
__global__ void cast_f32_bf16_kernel(const float* __restrict__ in, bf16* __restrict__ out, int64_t rows, int64_t cols,
                                     int64_t ld_out) {
  const int64_t total = rows * ld_out;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / ld_out, c = i - r * ld_out;
    out[i] = __float2bfloat16_rn(c < cols ? in[r * cols + c] : 0.f);
  }
}
__global__ void cast_f32_bf16_vec_kernel(const float* __restrict__ in, bf16* __restrict__ out, int64_t n8) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
    float f[8];
    *reinterpret_cast<float4*>(f) = *reinterpret_cast<const float4*>(in + i * 8);
    *reinterpret_cast<float4*>(f + 4) = *reinterpret_cast<const float4*>(in + i * 8 + 4);
    st8(out + i * 8, pack8(f));
  }
}
__global__ void cast_bf16_f32_kernel(const bf16* __restrict__ in, float* __restrict__ out, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = __bfloat162float(in[i]);
}

__global__ void add_bf16_kernel(const bf16* __restrict__ a, const bf16* __restrict__ b, bf16* __restrict__ out, int64_t n8) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
    float x[8], y[8];
    unpack8(ld8(a + i * 8), x);
    unpack8(ld8(b + i * 8), y);
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] += y[j];
    st8(out + i * 8, pack8(x));
  }
}

// DDT decoder conditioning (ddt.py:421-422): out[b,n,:] = silu(bf16(x[b,n,:] + v[b,:]))
__global__ void bias_silu_fwd_kernel(const bf16* __restrict__ x, const bf16* __restrict__ v, bf16* __restrict__ out,
                                     int64_t R, int rows_per_sample, int d) {
  const int nv = d >> 3;
  const int64_t total = R * nv;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / nv;
    const int c = (int)(i - row * nv) * 8;
    float a[8], b[8];
    unpack8(ld8(x + row * d + c), a);
    unpack8(ld8(v + (row / rows_per_sample) * d + c), b);
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = silu_f(bf16_round(a[j] + b[j]));
    st8(out + row * d + c, pack8(a));
  }
}
// dx = dy * silu'(x + v); dv[b,:] += sum_n dx   (one block per (sample, row chunk); threads own channel vectors)
__global__ void __launch_bounds__(256)
bias_silu_bwd_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x, const bf16* __restrict__ v,
                     bf16* __restrict__ dx, float* __restrict__ dv, int rows_per_sample, int rows_per_block, int d) {
  const int nv = d >> 3;
  const int b = blockIdx.y;
  const int r0 = blockIdx.x * rows_per_block, r1 = min(r0 + rows_per_block, rows_per_sample);
  for (int c8 = threadIdx.x; c8 < nv; c8 += blockDim.x) {
    const int c = c8 * 8;
    float vv[8], acc[8];
    unpack8(ld8(v + (int64_t)b * d + c), vv);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int r = r0; r < r1; ++r) {
      const int64_t row = (int64_t)b * rows_per_sample + r;
      float a[8], g[8];
      unpack8(ld8(x + row * d + c), a);
      unpack8(ld8(dy + row * d + c), g);
#pragma unroll
      for (int j = 0; j < 8; ++j) { g[j] *= dsilu_f(bf16_round(a[j] + vv[j])); acc[j] += g[j]; }
      st8(dx + row * d + c, pack8(g));
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) atomicAdd(dv + (int64_t)b * d + c + j, acc[j]);
  }
}

// y = silu(x); input fp32 or bf16, output bf16 (the GEMM operand dtype under autocast).
template <typename TIn>
__global__ void silu_fwd_kernel(const TIn* __restrict__ x, bf16* __restrict__ y, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    y[i] = __float2bfloat16_rn(silu_f((float)x[i]));
}
// dx = dy * silu'(x); dy fp32 or bf16; dx fp32 or bf16
template <typename TIn, typename TG, typename TOut>
__global__ void silu_bwd_kernel(const TG* __restrict__ dy, const TIn* __restrict__ x, TOut* __restrict__ dx, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    dx[i] = (TOut)((float)dy[i] * dsilu_f((float)x[i]));
}

// exact (erf) GELU on bf16 rows, 8 channels per thread: y = x Phi(x); backward dx = dy (Phi(x) + x phi(x))   (nn.GELU default,
// the PerceiverResampler feed-forward, reference networks/repa/perceiver_resampler.py:75-77)
__global__ void gelu_fwd_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, int64_t n8) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
    float f[8];
    unpack8(ld8(x + i * 8), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = 0.5f * f[j] * (1.f + erff(f[j] * 0.70710678118654752f));
    st8(y + i * 8, pack8(f));
  }
}
__global__ void gelu_bwd_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x, bf16* __restrict__ dx, int64_t n8) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
    float f[8], g[8];
    unpack8(ld8(x + i * 8), f);
    unpack8(ld8(dy + i * 8), g);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float cdf = 0.5f * (1.f + erff(f[j] * 0.70710678118654752f));
      const float pdf = 0.3989422804014327f * __expf(-0.5f * f[j] * f[j]);
      g[j] *= cdf + f[j] * pdf;
    }
    st8(dx + i * 8, pack8(g));
  }
}

// N-D interleaved-pair RoPE applied to the heads of a packed bf16 [R, ld] tensor (H heads of hd channels from column 0), in
// place or out of place; `inverse` applies the transposed rotation (the backward pass). No normalisation, no scale: the
// key-only rotation of the PerceiverResampler (reference perceiver_resampler.py:13-56). One thread per 8-channel vector.
__global__ void rope_apply_kernel(const bf16* __restrict__ x, int64_t ld_in, bf16* __restrict__ y, int64_t ld_out, const uint32_t* __restrict__ cs_t,
                                  int rot_half, const int32_t* __restrict__ pos_idx, int pos_offset, int tokens_per_sample, int hd, int d,
                                  int64_t R, int inverse) {
  const int nv = d >> 3;
  const int64_t total = R * nv;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / nv;
    const int c = (int)(i - row * nv) * 8;
    const int cl = c % hd;
    const uint32_t* csr = cs_t + (int64_t)(pos_idx ? pos_idx[row] : pos_offset + (int)(row % tokens_per_sample)) * rot_half;
    float f[8];
    unpack8(ld8(x + row * ld_in + c), f);
#pragma unroll
    for (int j = 0; j < 8; j += 2) {
      const int pj = (cl + j) >> 1;
      if (pj < rot_half) {
        const uint32_t cs = __ldg(csr + pj);
        const float co = __uint_as_float(cs << 16);
        float sn = __uint_as_float(cs & 0xffff0000u);
        if (inverse) sn = -sn;
        const float e = f[j], o = f[j + 1];
        f[j] = e * co - o * sn;
        f[j + 1] = e * sn + o * co;
      }
    }
    st8(y + row * ld_out + c, pack8(f));
  }
}

// te[b, :] = [cos(t_b f_i) | sin(t_b f_i)], f_i = exp(-ln(max_period) i / half); bf16 output (GEMM operand).
__global__ void timestep_embed_kernel(const float* __restrict__ t, bf16* __restrict__ out, int B, int dim,
                                      float max_period) {
  const int half = dim / 2;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * dim) return;
  const int b = i / dim, c = i - b * dim;
  float v = 0.f;
  if (c < 2 * half) {
    const int k = c < half ? c : c - half;
    const float freq = expf(-logf(max_period) * (float)k / (float)half);
    const float arg = t[b] * freq;
    v = c < half ? cosf(arg) : sinf(arg);
  }
  out[i] = __float2bfloat16_rn(v);
}

// emb = float(emb_time_bf16) [+ table[label]]; also emits silu(emb) in bf16 (input of every adaLN linear).
__global__ void cond_combine_kernel(const bf16* __restrict__ te, const float* __restrict__ table,
                                    const int64_t* __restrict__ labels, float* __restrict__ emb,
                                    bf16* __restrict__ emb_silu, int B, int E) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)B * E) return;
  const int b = (int)(i / E), c = (int)(i - (int64_t)b * E);
  float v = __bfloat162float(te[i]);
  if (table) v += table[labels[b] * E + c];
  emb[i] = v;
  emb_silu[i] = __float2bfloat16_rn(silu_f(v));
}

// dtable[label[b], :] += g[b, :]
__global__ void embedding_bwd_kernel(const float* __restrict__ g, const int64_t* __restrict__ labels,
                                     float* __restrict__ dtable, int B, int E) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)B * E) return;
  const int b = (int)(i / E), c = (int)(i - (int64_t)b * E);
  atomicAdd(dtable + labels[b] * E + c, g[i]);
}

// x [B,C,H,W] fp32 -> patches [B*Hp*Wp, Kp] bf16, column (c, p1, p2) (conv weight order), zero padded to Kp.
__global__ void patchify_kernel(const float* __restrict__ x, bf16* __restrict__ out, int B, int C, int H, int W, int p,
                                int Kp) {
  const int Hp = H / p, Wp = W / p;
  const int64_t total = (int64_t)B * Hp * Wp * Kp;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int col = (int)(i % Kp);
    const int64_t tok = i / Kp;
    float v = 0.f;
    if (col < C * p * p) {
      const int c = col / (p * p), r = col - c * p * p, p1 = r / p, p2 = r - p1 * p;
      const int w = (int)(tok % Wp), h = (int)((tok / Wp) % Hp);
      const int64_t b = tok / ((int64_t)Wp * Hp);
      v = x[((b * C + c) * H + (h * p + p1)) * W + (w * p + p2)];
    }
    out[i] = __float2bfloat16_rn(v);
  }
}

// tokens [B*Hp*Wp, p*p*C] (column (p1 p2 c)) -> image [B,C,H,W]; dtype preserved (bf16) or to fp32.
template <typename TOut>
__global__ void unpatchify_kernel(const bf16* __restrict__ tok, int64_t ld, TOut* __restrict__ img, int B, int C, int H,
                                  int W, int p) {
  const int Hp = H / p, Wp = W / p;
  const int64_t total = (int64_t)B * C * H * W;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % W), y = (int)((i / W) % H), c = (int)((i / ((int64_t)W * H)) % C);
    const int64_t b = i / ((int64_t)W * H * C);
    const int h = y / p, p1 = y - h * p, w = x / p, p2 = x - w * p;
    const int64_t row = (b * Hp + h) * Wp + w;
    img[i] = (TOut)tok[row * ld + (p1 * p + p2) * C + c];
  }
}
// image-layout gradient [B,C,H,W] (fp32 or bf16) -> token layout [B*Hp*Wp, p*p*C] bf16
template <typename TIn>
__global__ void patchify_grad_kernel(const TIn* __restrict__ img, bf16* __restrict__ tok, int64_t ld, int B, int C, int H,
                                     int W, int p) {
  const int Hp = H / p, Wp = W / p;
  const int ppc = p * p * C;
  const int64_t total = (int64_t)B * Hp * Wp * ppc;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int col = (int)(i % ppc);
    const int64_t row = i / ppc;
    const int c = col % C, pp = col / C, p1 = pp / p, p2 = pp - p1 * p;
    const int w = (int)(row % Wp), h = (int)((row / Wp) % Hp);
    const int64_t b = row / ((int64_t)Wp * Hp);
    tok[row * ld + col] = __float2bfloat16_rn((float)img[((b * C + c) * H + (h * p + p1)) * W + (w * p + p2)]);
  }
}

// out[c] += sum_r in[r, c]; grid (col tiles of 32*VEC?, row chunks). Simple and bandwidth friendly: each block
// owns 64 columns x a row chunk, threads (32 x 8): x -> column pair, y -> row stride.
template <typename TIn>
__global__ void __launch_bounds__(256) colsum_kernel(const TIn* __restrict__ in, int64_t ld, float* __restrict__ out,
                                                     int64_t R, int Cn, int64_t rows_per_block) {
  __shared__ float red[8][64];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c0 = blockIdx.x * 64 + tx * 2;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_block;
  const int64_t r1 = min(r0 + rows_per_block, R);
  float a0 = 0.f, a1 = 0.f;
  if (c0 < Cn) {
    for (int64_t r = r0 + ty; r < r1; r += 8) {
      a0 += (float)in[r * ld + c0];
      if (c0 + 1 < Cn) a1 += (float)in[r * ld + c0 + 1];
    }
  }
  red[ty][tx * 2] = a0;
  red[ty][tx * 2 + 1] = a1;
  __syncthreads();
  if (threadIdx.x < 64) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += red[i][threadIdx.x];
    const int c = blockIdx.x * 64 + threadIdx.x;
    if (c < Cn) atomicAdd(out + c, s);
  }
}

}  // namespace

DLB_EXPORT int dlb_cast_f32_bf16(const float* in, void* out, int64_t rows, int64_t cols, int64_t ld_out,
                                 cudaStream_t stream) {
  DLB_REQUIRE(rows > 0 && cols > 0 && ld_out >= cols, DLB_ERR_SHAPE, "cast_f32_bf16: bad shape");
  const int64_t n = rows * cols;
  if (ld_out == cols && n % 8 == 0 && ((uintptr_t)in % 16) == 0 && ((uintptr_t)out % 16) == 0)
    cast_f32_bf16_vec_kernel<<<grid_for(n / 8), 256, 0, stream>>>(in, (bf16*)out, n / 8);
  else
    cast_f32_bf16_kernel<<<grid_for(rows * ld_out), 256, 0, stream>>>(in, (bf16*)out, rows, cols, ld_out);
  dlb_count_launch();
  return dlb_check_launch("cast_f32_bf16");
}
DLB_EXPORT int dlb_cast_bf16_f32(const void* in, float* out, int64_t n, cudaStream_t stream) {
  DLB_REQUIRE(n > 0, DLB_ERR_SHAPE, "cast_bf16_f32: empty");
  cast_bf16_f32_kernel<<<grid_for(n), 256, 0, stream>>>((const bf16*)in, out, n);
  dlb_count_launch();
  return dlb_check_launch("cast_bf16_f32");
}

// out = a + b (bf16, n % 8 == 0): sum of two gradient branches that read the same activation
DLB_EXPORT int dlb_add_bf16(const void* a, const void* b, void* out, int64_t n, cudaStream_t stream) {
  DLB_REQUIRE(n > 0 && n % 8 == 0, DLB_ERR_SHAPE, "add_bf16: n must be a positive multiple of 8");
  add_bf16_kernel<<<grid_for(n / 8), 256, 0, stream>>>((const bf16*)a, (const bf16*)b, (bf16*)out, n / 8);
  dlb_count_launch();
  return dlb_check_launch("add_bf16");
}

DLB_EXPORT int dlb_bias_silu_fwd(const void* x, const void* v, void* out, int64_t B, int64_t rows_per_sample, int d,
                                 cudaStream_t stream) {
  DLB_REQUIRE(B > 0 && rows_per_sample > 0 && d > 0 && d % 8 == 0, DLB_ERR_SHAPE, "bias_silu_fwd: bad shape");
  bias_silu_fwd_kernel<<<grid_for(B * rows_per_sample * (d / 8)), 256, 0, stream>>>((const bf16*)x, (const bf16*)v, (bf16*)out,
                                                                                   B * rows_per_sample, (int)rows_per_sample, d);
  dlb_count_launch();
  return dlb_check_launch("bias_silu_fwd");
}
// dv: fp32 [B, d], accumulated
DLB_EXPORT int dlb_bias_silu_bwd(const void* dy, const void* x, const void* v, void* dx, float* dv, int64_t B,
                                 int64_t rows_per_sample, int d, cudaStream_t stream) {
  DLB_REQUIRE(B > 0 && rows_per_sample > 0 && d > 0 && d % 8 == 0, DLB_ERR_SHAPE, "bias_silu_bwd: bad shape");
  const int rpb = 32;
  dim3 grid((unsigned)((rows_per_sample + rpb - 1) / rpb), (unsigned)B);
  bias_silu_bwd_kernel<<<grid, 256, 0, stream>>>((const bf16*)dy, (const bf16*)x, (const bf16*)v, (bf16*)dx, dv,
                                                 (int)rows_per_sample, rpb, d);
  dlb_count_launch();
  return dlb_check_launch("bias_silu_bwd");
}

// in_dtype: 0 = bf16, 1 = fp32
DLB_EXPORT int dlb_silu_fwd(const void* x, int in_dtype, void* y, int64_t n, cudaStream_t stream) {
  DLB_REQUIRE(n > 0, DLB_ERR_SHAPE, "silu_fwd: empty");
  if (in_dtype == 0) silu_fwd_kernel<bf16><<<grid_for(n), 256, 0, stream>>>((const bf16*)x, (bf16*)y, n);
  else silu_fwd_kernel<float><<<grid_for(n), 256, 0, stream>>>((const float*)x, (bf16*)y, n);
  dlb_count_launch();
  return dlb_check_launch("silu_fwd");
}
// dx = dy * silu'(x). (x_dtype, dy_dtype, dx_dtype) in {0 = bf16, 1 = fp32}; supported: (0,0,0), (1,1,1), (1,0,1), (0,1,0)
DLB_EXPORT int dlb_silu_bwd(const void* dy, int dy_dtype, const void* x, int x_dtype, void* dx, int dx_dtype, int64_t n,
                            cudaStream_t stream) {
  DLB_REQUIRE(n > 0, DLB_ERR_SHAPE, "silu_bwd: empty");
  const int g = grid_for(n);
  if (x_dtype == 0 && dy_dtype == 0 && dx_dtype == 0)
    silu_bwd_kernel<bf16, bf16, bf16><<<g, 256, 0, stream>>>((const bf16*)dy, (const bf16*)x, (bf16*)dx, n);
  else if (x_dtype == 1 && dy_dtype == 1 && dx_dtype == 1)
    silu_bwd_kernel<float, float, float><<<g, 256, 0, stream>>>((const float*)dy, (const float*)x, (float*)dx, n);
  else if (x_dtype == 1 && dy_dtype == 0 && dx_dtype == 1)
    silu_bwd_kernel<float, bf16, float><<<g, 256, 0, stream>>>((const bf16*)dy, (const float*)x, (float*)dx, n);
  else if (x_dtype == 0 && dy_dtype == 1 && dx_dtype == 0)
    silu_bwd_kernel<bf16, float, bf16><<<g, 256, 0, stream>>>((const float*)dy, (const bf16*)x, (bf16*)dx, n);
  else {
    dlb_set_error("silu_bwd: unsupported dtype combination (%d,%d,%d)", x_dtype, dy_dtype, dx_dtype);
    return DLB_ERR_UNSUPPORTED;
  }
  dlb_count_launch();
  return dlb_check_launch("silu_bwd");
}

DLB_EXPORT int dlb_timestep_embed(const float* t, void* out, int B, int dim, float max_period, cudaStream_t stream) {
  DLB_REQUIRE(B > 0 && dim > 0, DLB_ERR_SHAPE, "timestep_embed: bad shape");
  timestep_embed_kernel<<<(B * dim + 255) / 256, 256, 0, stream>>>(t, (bf16*)out, B, dim, max_period);
  dlb_count_launch();
  return dlb_check_launch("timestep_embed");
}

DLB_EXPORT int dlb_cond_combine(const void* te, const float* table, const int64_t* labels, float* emb, void* emb_silu,
                                int B, int E, cudaStream_t stream) {
  DLB_REQUIRE(B > 0 && E > 0 && ((table == nullptr) == (labels == nullptr)), DLB_ERR_SHAPE, "cond_combine: bad args");
  cond_combine_kernel<<<(unsigned)(((int64_t)B * E + 255) / 256), 256, 0, stream>>>((const bf16*)te, table, labels, emb,
                                                                                   (bf16*)emb_silu, B, E);
  dlb_count_launch();
  return dlb_check_launch("cond_combine");
}

DLB_EXPORT int dlb_embedding_bwd(const float* g, const int64_t* labels, float* dtable, int B, int E, cudaStream_t stream) {
  DLB_REQUIRE(B > 0 && E > 0, DLB_ERR_SHAPE, "embedding_bwd: bad shape");
  embedding_bwd_kernel<<<(unsigned)(((int64_t)B * E + 255) / 256), 256, 0, stream>>>(g, labels, dtable, B, E);
  dlb_count_launch();
  return dlb_check_launch("embedding_bwd");
}

DLB_EXPORT int dlb_patchify(const float* x, void* out, int B, int C, int H, int W, int p, int Kp, cudaStream_t stream) {
  DLB_REQUIRE(B > 0 && C > 0 && p > 0 && H % p == 0 && W % p == 0 && Kp >= C * p * p && Kp % 8 == 0, DLB_ERR_SHAPE,
              "patchify: bad shape B=%d C=%d H=%d W=%d p=%d Kp=%d", B, C, H, W, p, Kp);
  patchify_kernel<<<grid_for((int64_t)B * (H / p) * (W / p) * Kp), 256, 0, stream>>>(x, (bf16*)out, B, C, H, W, p, Kp);
  dlb_count_launch();
  return dlb_check_launch("patchify");
}

// out_dtype: 0 = bf16, 1 = fp32
DLB_EXPORT int dlb_unpatchify(const void* tok, int64_t ld, void* img, int out_dtype, int B, int C, int H, int W, int p,
                              cudaStream_t stream) {
  DLB_REQUIRE(B > 0 && C > 0 && p > 0 && H % p == 0 && W % p == 0 && ld >= p * p * C, DLB_ERR_SHAPE, "unpatchify: bad shape");
  const int g = grid_for((int64_t)B * C * H * W);
  if (out_dtype == 0) unpatchify_kernel<bf16><<<g, 256, 0, stream>>>((const bf16*)tok, ld, (bf16*)img, B, C, H, W, p);
  else unpatchify_kernel<float><<<g, 256, 0, stream>>>((const bf16*)tok, ld, (float*)img, B, C, H, W, p);
  dlb_count_launch();
  return dlb_check_launch("unpatchify");
}
// in_dtype: 0 = bf16, 1 = fp32
DLB_EXPORT int dlb_patchify_grad(const void* img, int in_dtype, void* tok, int64_t ld, int B, int C, int H, int W, int p,
                                 cudaStream_t stream) {
  DLB_REQUIRE(B > 0 && C > 0 && p > 0 && H % p == 0 && W % p == 0 && ld >= p * p * C, DLB_ERR_SHAPE, "patchify_grad: bad shape");
  const int g = grid_for((int64_t)B * C * H * W);
  if (in_dtype == 0) patchify_grad_kernel<bf16><<<g, 256, 0, stream>>>((const bf16*)img, (bf16*)tok, ld, B, C, H, W, p);
  else patchify_grad_kernel<float><<<g, 256, 0, stream>>>((const float*)img, (bf16*)tok, ld, B, C, H, W, p);
  dlb_count_launch();
  return dlb_check_launch("patchify_grad");
}

// out[c] += sum over rows of in[r, c]  (in_dtype 0 = bf16, 1 = fp32); out is fp32 and must be initialised.
DLB_EXPORT int dlb_colsum(const void* in, int in_dtype, int64_t ld, float* out, int64_t R, int Cn, cudaStream_t stream) {
  DLB_REQUIRE(R > 0 && Cn > 0 && ld >= Cn, DLB_ERR_SHAPE, "colsum: bad shape");
  const int col_blocks = (Cn + 63) / 64;
  int64_t row_blocks = ((int64_t)dlb_num_sms() * 8 + col_blocks - 1) / col_blocks;
  int64_t rpb = (R + row_blocks - 1) / row_blocks;
  if (rpb < 64) rpb = 64;
  row_blocks = (R + rpb - 1) / rpb;
  dim3 grid(col_blocks, (unsigned)row_blocks);
  if (in_dtype == 0) colsum_kernel<bf16><<<grid, 256, 0, stream>>>((const bf16*)in, ld, out, R, Cn, rpb);
  else colsum_kernel<float><<<grid, 256, 0, stream>>>((const float*)in, ld, out, R, Cn, rpb);
  dlb_count_launch();
  return dlb_check_launch("colsum");
}

DLB_EXPORT int dlb_gelu_fwd(const void* x, void* y, int64_t n, cudaStream_t stream) {
  DLB_REQUIRE(n > 0 && n % 8 == 0, DLB_ERR_SHAPE, "gelu_fwd: element count must be a positive multiple of 8");
  gelu_fwd_kernel<<<grid_for(n / 8), 256, 0, stream>>>((const bf16*)x, (bf16*)y, n / 8);
  dlb_count_launch();
  return dlb_check_launch("gelu_fwd");
}
DLB_EXPORT int dlb_gelu_bwd(const void* dy, const void* x, void* dx, int64_t n, cudaStream_t stream) {
  DLB_REQUIRE(n > 0 && n % 8 == 0, DLB_ERR_SHAPE, "gelu_bwd: element count must be a positive multiple of 8");
  gelu_bwd_kernel<<<grid_for(n / 8), 256, 0, stream>>>((const bf16*)dy, (const bf16*)x, (bf16*)dx, n / 8);
  dlb_count_launch();
  return dlb_check_launch("gelu_bwd");
}
// x, y: bf16 [R, ld] (y may alias x); cs_t: packed bf16x2 (cos, sin) table [positions, rot_half] as written by dlb_rope_table
DLB_EXPORT int dlb_rope_apply(const void* x, int64_t ld_in, void* y, int64_t ld_out, const uint32_t* cs_t, int rot_half, const int32_t* pos_idx,
                              int pos_offset, int tokens_per_sample, int hd, int d, int64_t R, int inverse, cudaStream_t stream) {
  DLB_REQUIRE(R > 0 && d > 0 && d % 8 == 0 && hd > 0 && hd % 8 == 0 && d % hd == 0 && 2 * rot_half <= hd && ld_in % 8 == 0 && ld_out % 8 == 0,
              DLB_ERR_SHAPE, "rope_apply: bad shape R=%lld d=%d hd=%d rot_half=%d", (long long)R, d, hd, rot_half);
  DLB_REQUIRE(pos_idx != nullptr || tokens_per_sample > 0, DLB_ERR_SHAPE, "rope_apply: tokens_per_sample required without pos_idx");
  rope_apply_kernel<<<grid_for(R * (d / 8)), 256, 0, stream>>>((const bf16*)x, ld_in, (bf16*)y, ld_out, cs_t, rot_half, pos_idx, pos_offset,
                                                               tokens_per_sample > 0 ? tokens_per_sample : 1, hd, d, R, inverse);
  dlb_count_launch();
  return dlb_check_launch("rope_apply");
}
