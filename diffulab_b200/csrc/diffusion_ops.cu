// Formalisation-side kernels: rectified-flow / DDPM interpolation, velocity-target MSE (+ its gradient), REPA
// cosine loss, SPRINT token selection / gather / restore, Euler (+CFG) sampler update, fused AdamW.
//
// Reference: Flow.add_noise / compute_loss diffuse/modelizations/flow.py:262-315,382-408;
// GaussianDiffusion.add_noise gaussian_diffusion.py:313-341 (same kernel with per-sample (a,b));
// RepaLoss.forward training/losses/repa.py:159-186; SprintDiT.drop_tokens / restore_tokens
// networks/denoisers/sprint.py:317-387; Euler.step samplers/flow/euler.py:22-41 and the CFG combine
// flow.py:256-260; torch.optim.AdamW as used by BaseTrainer.training_step base_trainer.py:149.
#include "common.cuh"

namespace {
typedef __nv_bfloat16 bf16;

int grid_for(int64_t n) {
  int64_t blocks = (n + 255) / 256;
  const int64_t cap = (int64_t)dlb_num_sms() * 16;
  return (int)(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}

// x_t = a_b * x0 + b_b * eps with per-sample coefficients. Flow: a = 1 - t, b = t.
__global__ void interp_kernel(const float* __restrict__ x0, const float* __restrict__ eps, const float* __restrict__ a,
                              const float* __restrict__ b, float* __restrict__ xt, int64_t per_sample, int64_t total) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t s = i / per_sample;
    xt[i] = a[s] * x0[i] + b[s] * eps[i];
  }
}

// loss_sum += sum (target - v)^2 with target = eps - x0 (flow) or eps (ddpm: x0 == nullptr);
// v = pred (v-prediction) or (x_t - pred) / t_b (x-prediction, flow.py:300-303).
template <typename TP>
__global__ void __launch_bounds__(256)
mse_fwd_kernel(const TP* __restrict__ pred, const float* __restrict__ x0, const float* __restrict__ eps,
               const float* __restrict__ xt, const float* __restrict__ t, int64_t per_sample, int64_t total,
               float inv_total, float* __restrict__ loss) {
  __shared__ float red[32];
  float acc = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    float v = (float)pred[i];
    if (xt) v = (xt[i] - v) / t[i / per_sample];
    const float tgt = x0 ? eps[i] - x0[i] : eps[i];
    const float d = tgt - v;
    acc += d * d;
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) atomicAdd(loss, acc * inv_total);
}
// dpred = gout * d loss / d pred
template <typename TP>
__global__ void mse_bwd_kernel(const TP* __restrict__ pred, const float* __restrict__ x0, const float* __restrict__ eps,
                               const float* __restrict__ xt, const float* __restrict__ t, int64_t per_sample,
                               int64_t total, float inv_total, const float* __restrict__ gout, TP* __restrict__ dpred) {
  const float g = gout ? *gout : 1.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    float v = (float)pred[i];
    float dv_dpred = 1.f;
    if (xt) {
      const float ti = t[i / per_sample];
      v = (xt[i] - v) / ti;
      dv_dpred = -1.f / ti;
    }
    const float tgt = x0 ? eps[i] - x0[i] : eps[i];
    dpred[i] = (TP)(g * (-2.f * inv_total) * (tgt - v) * dv_dpred);
  }
}

// REPA: one warp per token row. cos = <s,z> / (max(|s|,eps) max(|z|,eps)); loss += coeff * (1 - mean cos).
__global__ void __launch_bounds__(256)
repa_cos_fwd_kernel(const bf16* __restrict__ s, const float* __restrict__ z, int64_t R, int E, float coeff_over_R,
                    float* __restrict__ loss) {
  __shared__ float red[32];
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  float contrib = 0.f;
  if (row < R) {
    float dot = 0.f, ss = 0.f, zz = 0.f;
    for (int c = lane * 8; c < E; c += 256) {
      float a[8], b[8];
      unpack8(ld8(s + row * E + c), a);
      *reinterpret_cast<float4*>(b) = *reinterpret_cast<const float4*>(z + row * E + c);
      *reinterpret_cast<float4*>(b + 4) = *reinterpret_cast<const float4*>(z + row * E + c + 4);
#pragma unroll
      for (int j = 0; j < 8; ++j) { dot += a[j] * b[j]; ss += a[j] * a[j]; zz += b[j] * b[j]; }
    }
    dot = warp_sum(dot); ss = warp_sum(ss); zz = warp_sum(zz);
    const float ns = fmaxf(sqrtf(ss), 1e-8f), nz = fmaxf(sqrtf(zz), 1e-8f);
    if (lane == 0) contrib = coeff_over_R * (1.f - dot / (ns * nz));
  }
  contrib = block_sum(contrib, red);
  if (threadIdx.x == 0) atomicAdd(loss, contrib);
}
// ds = gout * (-coeff/R) * ( z / (|s||z|) - cos * s / |s|^2 )
__global__ void __launch_bounds__(256)
repa_cos_bwd_kernel(const bf16* __restrict__ s, const float* __restrict__ z, int64_t R, int E, float coeff_over_R,
                    const float* __restrict__ gout, bf16* __restrict__ ds) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= R) return;
  const float g = (gout ? *gout : 1.f) * (-coeff_over_R);
  float dot = 0.f, ss = 0.f, zz = 0.f;
  for (int c = lane * 8; c < E; c += 256) {
    float a[8], b[8];
    unpack8(ld8(s + row * E + c), a);
    *reinterpret_cast<float4*>(b) = *reinterpret_cast<const float4*>(z + row * E + c);
    *reinterpret_cast<float4*>(b + 4) = *reinterpret_cast<const float4*>(z + row * E + c + 4);
#pragma unroll
    for (int j = 0; j < 8; ++j) { dot += a[j] * b[j]; ss += a[j] * a[j]; zz += b[j] * b[j]; }
  }
  dot = warp_sum(dot); ss = warp_sum(ss); zz = warp_sum(zz);
  const float ns_raw = sqrtf(ss);
  const float ns = fmaxf(ns_raw, 1e-8f), nz = fmaxf(sqrtf(zz), 1e-8f);
  const float inv = 1.f / (ns * nz);
  // d/ds of max(|s|, eps) vanishes when the clamp is active
  const float k2 = ns_raw > 1e-8f ? dot * inv / (ns * ns) : 0.f;
  for (int c = lane * 8; c < E; c += 256) {
    float a[8], b[8], o[8];
    unpack8(ld8(s + row * E + c), a);
    *reinterpret_cast<float4*>(b) = *reinterpret_cast<const float4*>(z + row * E + c);
    *reinterpret_cast<float4*>(b + 4) = *reinterpret_cast<const float4*>(z + row * E + c + 4);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = g * (b[j] * inv - k2 * a[j]);
    st8(ds + row * E + c, pack8(o));
  }
}

// SPRINT selection: one block per sample. kept = indices of the k largest scores, ascending; ties broken towards
// the larger index (matches the CPU torch.topk probe in SURVEY.md A.10; exactness is defined on tie-free draws).
// inv[b, s] = slot of token s among the kept ones, or -1.
__global__ void __launch_bounds__(256)
sprint_select_kernel(const float* __restrict__ scores, int S, int k, int64_t* __restrict__ kept, int32_t* __restrict__ kept32,
                     int32_t* __restrict__ inv) {
  extern __shared__ float sc[];              // S scores, then S flags (as int)
  int* flag = reinterpret_cast<int*>(sc + S);
  __shared__ int warp_tot[8];
  const int b = blockIdx.x;
  for (int i = threadIdx.x; i < S; i += blockDim.x) sc[i] = scores[(int64_t)b * S + i];
  __syncthreads();
  for (int i = threadIdx.x; i < S; i += blockDim.x) {
    const float v = sc[i];
    int rank = 0;
    for (int j = 0; j < S; ++j) {
      const float u = sc[j];
      rank += (u > v) || (u == v && j > i);
    }
    flag[i] = rank < k;
  }
  __syncthreads();
  // ordered compaction: chunks of blockDim.x tokens, ballot-based exclusive scan
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  int base = 0;
  for (int c0 = 0; c0 < S; c0 += blockDim.x) {
    const int i = c0 + threadIdx.x;
    const int f = (i < S) ? flag[i] : 0;
    const unsigned bal = __ballot_sync(0xffffffffu, f);
    if (lane == 0) warp_tot[warp] = __popc(bal);
    __syncthreads();
    int off = base;
    for (int w = 0; w < warp; ++w) off += warp_tot[w];
    const int slot = off + __popc(bal & ((1u << lane) - 1u));
    if (i < S) {
      if (f) {
        kept[(int64_t)b * k + slot] = i;
        if (kept32) kept32[(int64_t)b * k + slot] = i;
      }
      inv[(int64_t)b * S + i] = f ? slot : -1;
    }
    int tot = 0;
    for (int w = 0; w < nw; ++w) tot += warp_tot[w];
    base += tot;
    __syncthreads();
  }
}

// out[b, j, :] = x[b, idx[b, j], :]
__global__ void gather_rows_kernel(const bf16* __restrict__ x, const int64_t* __restrict__ idx, bf16* __restrict__ out,
                                   int B, int S, int k, int d) {
  const int nv = d >> 3;
  const int64_t total = (int64_t)B * k * nv;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int v = (int)(i % nv);
    const int64_t bj = i / nv;
    const int64_t b = bj / k;
    const int64_t src = b * S + idx[bj];
    st8(out + bj * d + v * 8, ld8(x + src * d + v * 8));
  }
}
// out[b, s, :] = inv[b,s] >= 0 && !drop[b] ? xk[b, inv[b,s], :] : fill[:]     (fill == nullptr -> zeros)
__global__ void restore_rows_kernel(const bf16* __restrict__ xk, const int32_t* __restrict__ inv,
                                    const float* __restrict__ fill, const uint8_t* __restrict__ drop,
                                    bf16* __restrict__ out, int B, int S, int k, int d) {
  const int nv = d >> 3;
  const int64_t total = (int64_t)B * S * nv;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int v = (int)(i % nv);
    const int64_t bs = i / nv;
    const int64_t b = bs / S;
    const int slot = inv[bs];
    if (slot >= 0 && !(drop && drop[b])) {
      st8(out + bs * d + v * 8, ld8(xk + (b * k + slot) * d + v * 8));
    } else {
      float f[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = fill ? fill[v * 8 + j] : 0.f;
      st8(out + bs * d + v * 8, pack8(f));
    }
  }
}
// restore backward: dxk[b, j, :] = drop[b] ? 0 : dy[b, idx[b,j], :];  dfill[:] += sum over filled positions of dy
__global__ void __launch_bounds__(256)
restore_bwd_kernel(const bf16* __restrict__ dy, const int64_t* __restrict__ idx, const int32_t* __restrict__ inv,
                   const uint8_t* __restrict__ drop, bf16* __restrict__ dxk, float* __restrict__ dfill, int B, int S,
                   int k, int d) {
  // part 1: gather
  const int nv = d >> 3;
  const int64_t total = (int64_t)B * k * nv;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int v = (int)(i % nv);
    const int64_t bj = i / nv;
    const int64_t b = bj / k;
    bf16x8 val;
    if (drop && drop[b]) { val.u[0] = val.u[1] = val.u[2] = val.u[3] = 0u; }
    else val = ld8(dy + (b * S + idx[bj]) * d + v * 8);
    st8(dxk + bj * d + v * 8, val);
  }
  // part 2: mask-token gradient; each thread owns channels, strides over (b, s) rows assigned to this block
  if (!dfill) return;
  const int64_t rows = (int64_t)B * S;
  const int64_t rows_per_block = (rows + gridDim.x - 1) / gridDim.x;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block, r1 = min(r0 + rows_per_block, rows);
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    float acc = 0.f;
    for (int64_t r = r0; r < r1; ++r) {
      const int64_t b = r / S;
      if (inv[r] < 0 || (drop && drop[b])) acc += __bfloat162float(dy[r * d + c]);
    }
    if (acc != 0.f) atomicAdd(dfill + c, acc);
  }
}

// Euler step with optional classifier-free guidance: v = vu + g (vc - vu); x_prev = x - v dt; x0_est = x - v t.
template <typename TV>
__global__ void euler_step_kernel(const float* __restrict__ x, const TV* __restrict__ vc, const TV* __restrict__ vu,
                                  float guidance, float dt, float t_curr, float* __restrict__ x_prev,
                                  float* __restrict__ x0_est, float* __restrict__ v_out, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float v = (float)vc[i];
    if (vu) { const float u = (float)vu[i]; v = u + guidance * (v - u); }
    const float xi = x[i];
    x_prev[i] = xi - v * dt;
    if (x0_est) x0_est[i] = xi - v * t_curr;
    if (v_out) v_out[i] = v;
  }
}

// One reverse step of the Gaussian samplers (reference samplers/gaussian_diffusion/ddpm.py `step`, ddim.py `step`):
// x0 from the model output (epsilon / xstart / xprev parameterisation, optional clamp), posterior mean, x_{t-1} =
// mean + mask * std * noise and the per-element log-probability, in one pass. All schedule-dependent scalars come
// from a per-timestep fp32 table built on the host from the float64 schedule exactly as `extract_into_tensor` yields
// them (GS_* columns); arithmetic keeps the reference's operation order (no FMA contraction) so results agree to ~1 ulp.
enum { GS_RSAB = 0, GS_CEPS, GS_RC1, GS_C2C1, GS_C1, GS_C2, GS_VAR, GS_STD, GS_MASK, GS_EDEN, GS_SABP, GS_SA, GS_SB, GS_ABP, GS_MINLOG, GS_MAXLOG, GS_COLS = 16 };
// var_mode 0: variance from the table (fixed_small / fixed_large). 1 ("learned") / 2 ("learned_range"): the model output holds
// 2C channels per sample, [mean prediction | variance head] (torch.chunk(.., 2, dim=1), ddpm.py:268-270): sample b's prediction
// starts at pred + b * pred_stride and its variance head per_sample elements later; the per-element variance follows
// ddpm.py:213-223 and the per-element std (x_prev_std of DDPM.step) goes to std_out.
template <typename TP>
__global__ void gaussian_step_kernel(const TP* __restrict__ pred, const float* __restrict__ xt, const float* __restrict__ noise,
                                     const float* __restrict__ table, const int* __restrict__ t, int sampler, int mean_type,
                                     int var_mode, int clamp, float eta, int64_t per_sample, int64_t pred_stride,
                                     float* __restrict__ x_prev, float* __restrict__ x0_out, float* __restrict__ mean_out,
                                     float* __restrict__ logprob, float* __restrict__ std_out) {
  const int b = blockIdx.y;
  const float* c = table + (int64_t)t[b] * GS_COLS;
  const float r_sab = c[GS_RSAB], c_eps = c[GS_CEPS], r_c1 = c[GS_RC1], c2c1 = c[GS_C2C1], c1 = c[GS_C1], c2 = c[GS_C2];
  const float mask = c[GS_MASK], e_den = c[GS_EDEN], sabp = c[GS_SABP], abp = c[GS_ABP];
  const float min_log = c[GS_MINLOG], max_log = c[GS_MAXLOG];
  const float sigma = __fmul_rn(__fmul_rn(eta, c[GS_SA]), c[GS_SB]);
  const float dir = sqrtf(__fsub_rn(__fsub_rn(1.f, abp), __fmul_rn(sigma, sigma)));
  float stdv = c[GS_STD], vs = fmaxf(c[GS_VAR], 1e-20f);
  float two_vs = __fmul_rn(2.f, vs), lconst = __fmul_rn(logf(__fmul_rn(6.283185307179586f, vs)), 0.5f);
  const float two_s2 = __fmul_rn(2.f, __fmul_rn(sigma, sigma)), lsig = logf(sigma), lhalf = __fmul_rn(0.5f, logf(6.283185307179586f));
  const int64_t base = (int64_t)b * per_sample;
  const TP* pb = pred + (int64_t)b * pred_stride;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < per_sample; i += (int64_t)gridDim.x * blockDim.x) {
    const float p = (float)pb[i], x = xt[base + i], nz = noise[base + i];
    float x0;
    if (mean_type == 0) x0 = __fsub_rn(__fmul_rn(r_sab, x), __fmul_rn(c_eps, p));
    else if (mean_type == 1) x0 = p;
    else x0 = __fsub_rn(__fmul_rn(r_c1, p), __fmul_rn(c2c1, x));
    if (clamp && x0 == x0) x0 = fminf(fmaxf(x0, -1.f), 1.f);  // NaN propagates, like torch.clamp
    float mean, xp, lp;
    if (sampler == 0) {
      if (var_mode != 0) {
        float lv = (float)pb[per_sample + i];
        if (var_mode == 2) {
          const float w = __fdiv_rn(__fadd_rn(lv, 1.f), 2.f);
          lv = __fadd_rn(__fmul_rn(w, max_log), __fmul_rn(__fsub_rn(1.f, w), min_log));
        }
        vs = fmaxf(expf(lv), 1e-20f);
        stdv = expf(__fmul_rn(0.5f, lv));
        two_vs = __fmul_rn(2.f, vs);
        lconst = __fmul_rn(logf(__fmul_rn(6.283185307179586f, vs)), 0.5f);
        std_out[base + i] = sqrtf(vs);
      }
      mean = __fadd_rn(__fmul_rn(c1, x0), __fmul_rn(c2, x));
      xp = __fadd_rn(mean, __fmul_rn(__fmul_rn(mask, nz), stdv));
      const float d = __fsub_rn(xp, mean);
      lp = __fmul_rn(__fsub_rn(__fdiv_rn(-__fmul_rn(d, d), two_vs), lconst), mask);
    } else {
      const float eps = __fdiv_rn(__fsub_rn(__fmul_rn(r_sab, x), x0), e_den);
      mean = __fadd_rn(__fmul_rn(x0, sabp), __fmul_rn(dir, eps));
      xp = __fadd_rn(mean, __fmul_rn(__fmul_rn(mask, sigma), nz));
      const float d = __fsub_rn(xp, mean);
      lp = -__fadd_rn(__fadd_rn(__fdiv_rn(__fmul_rn(d, d), two_s2), lsig), lhalf);
    }
    x_prev[base + i] = xp;
    x0_out[base + i] = x0;
    mean_out[base + i] = mean;
    if (logprob) logprob[base + i] = lp;
  }
}

// Euler-Maruyama flow step (reference samplers/flow/euler_meruyama.py:24-57):
//   mean = x - (v + c (x + (1 - t) v)) dt,  c = sigma^2 / (2 t);  x_prev = mean + std * noise (or a given x_prev);
//   x0_est = x - v t;  logprob = -((x_prev - mean)^2 / (2 std^2) + log std + 0.5 log 2 pi).  fp32, reference operation order.
template <typename TV>
__global__ void euler_maruyama_kernel(const float* __restrict__ x, const TV* __restrict__ v, const float* __restrict__ noise,
                                      const float* __restrict__ x_prev_in, float c, float one_minus_t, float dt, float t_curr,
                                      float stdv, float* __restrict__ x_prev, float* __restrict__ mean_out,
                                      float* __restrict__ x0_est, float* __restrict__ logprob, int64_t n) {
  const float two_s2 = __fmul_rn(2.f, __fmul_rn(stdv, stdv)), lstd = logf(stdv), lhalf = __fmul_rn(0.5f, logf(6.283185307179586f));
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float xi = x[i], vi = (float)v[i];
    const float inner = __fmul_rn(c, __fadd_rn(xi, __fmul_rn(one_minus_t, vi)));
    const float mean = __fsub_rn(xi, __fmul_rn(__fadd_rn(vi, inner), dt));
    const float xp = x_prev_in ? x_prev_in[i] : __fadd_rn(mean, __fmul_rn(stdv, noise[i]));
    const float d = __fsub_rn(xp, mean);
    x_prev[i] = xp;
    mean_out[i] = mean;
    x0_est[i] = __fsub_rn(xi, __fmul_rn(vi, t_curr));
    logprob[i] = -__fadd_rn(__fadd_rn(__fdiv_rn(__fmul_rn(d, d), two_s2), lstd), lhalf);
  }
}

// torch.optim.AdamW (no amsgrad, no maximize): p *= 1 - lr*wd; m,v update; p -= lr/bc1 * m / (sqrt(v)/sqrt(bc2) + eps).
// Also refreshes the bf16 shadow copy used by the GEMMs and (optionally) the EMA copy.
// omb1 / omb2 = 1 - beta evaluated in DOUBLE on the host (1.f - 0.999f is off by 1.3e-5 relative, visible in exp_avg_sq)
__device__ __forceinline__ void adamw_one(float& pi, float gi, float& mi, float& vi, float decay, float beta1, float beta2, float omb1,
                                          float omb2, float eps, float step_size, float bc2_sqrt) {
  pi *= decay;  // decoupled weight decay: p *= 1 - lr * wd
  mi = beta1 * mi + omb1 * gi;
  vi = beta2 * vi + omb2 * gi * gi;
  pi -= step_size * (mi / (sqrtf(vi) / bc2_sqrt + eps));
}
// 4 parameters per thread per iteration (128-bit loads / stores; n4 = n / 4), scalar tail handled by the last block.
// chunk_active (nullable): one byte per 64-element chunk of the flat buffer; a zero byte skips the chunk entirely —
// torch.optim.AdamW skips parameters whose .grad is None (no decay, no moment update, no step).
__global__ void __launch_bounds__(256)
adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
             bf16* __restrict__ shadow, float* __restrict__ ema, float ema_decay, const uint8_t* __restrict__ chunk_active,
             int64_t n, float decay, float beta1, float beta2, float omb1, float omb2, float eps, float step_size, float bc2_sqrt, float grad_scale) {
  const int64_t n4 = n >> 2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    if (chunk_active != nullptr && chunk_active[i >> 4] == 0) continue;
    float4 pv = reinterpret_cast<float4*>(p)[i];
    const float4 gv = reinterpret_cast<const float4*>(g)[i];
    float4 mv = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    adamw_one(pv.x, gv.x * grad_scale, mv.x, vv.x, decay, beta1, beta2, omb1, omb2, eps, step_size, bc2_sqrt);
    adamw_one(pv.y, gv.y * grad_scale, mv.y, vv.y, decay, beta1, beta2, omb1, omb2, eps, step_size, bc2_sqrt);
    adamw_one(pv.z, gv.z * grad_scale, mv.z, vv.z, decay, beta1, beta2, omb1, omb2, eps, step_size, bc2_sqrt);
    adamw_one(pv.w, gv.w * grad_scale, mv.w, vv.w, decay, beta1, beta2, omb1, omb2, eps, step_size, bc2_sqrt);
    reinterpret_cast<float4*>(p)[i] = pv;
    reinterpret_cast<float4*>(m)[i] = mv;
    reinterpret_cast<float4*>(v)[i] = vv;
    if (shadow) {
      uint2 s2;
      s2.x = pack_bf16x2(pv.x, pv.y);
      s2.y = pack_bf16x2(pv.z, pv.w);
      reinterpret_cast<uint2*>(shadow)[i] = s2;
    }
    if (ema) {  // ema.lerp_(p, 1 - decay) (ema_pytorch update_moving_average); decay 0 = copy
      float4 ev = reinterpret_cast<float4*>(ema)[i];
      ev.x = ev.x * ema_decay + pv.x * (1.f - ema_decay);
      ev.y = ev.y * ema_decay + pv.y * (1.f - ema_decay);
      ev.z = ev.z * ema_decay + pv.z * (1.f - ema_decay);
      ev.w = ev.w * ema_decay + pv.w * (1.f - ema_decay);
      reinterpret_cast<float4*>(ema)[i] = ev;
    }
  }
  if (blockIdx.x == gridDim.x - 1) {
    for (int64_t i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) {
      if (chunk_active != nullptr && chunk_active[i >> 6] == 0) continue;
      float pi = p[i], mi = m[i], vi = v[i];
      adamw_one(pi, g[i] * grad_scale, mi, vi, decay, beta1, beta2, omb1, omb2, eps, step_size, bc2_sqrt);
      p[i] = pi; m[i] = mi; v[i] = vi;
      if (shadow) shadow[i] = __float2bfloat16_rn(pi);
      if (ema) ema[i] = ema[i] * ema_decay + pi * (1.f - ema_decay);
    }
  }
}

// ema = ema * decay + p * (1 - decay) over a flat buffer (EMA updates that do not coincide with an optimizer step)
__global__ void __launch_bounds__(256) ema_lerp_kernel(float* __restrict__ ema, const float* __restrict__ p, float decay, int64_t n) {
  const int64_t n4 = n >> 2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 ev = reinterpret_cast<float4*>(ema)[i];
    const float4 pv = reinterpret_cast<const float4*>(p)[i];
    ev.x = ev.x * decay + pv.x * (1.f - decay);
    ev.y = ev.y * decay + pv.y * (1.f - decay);
    ev.z = ev.z * decay + pv.z * (1.f - decay);
    ev.w = ev.w * decay + pv.w * (1.f - decay);
    reinterpret_cast<float4*>(ema)[i] = ev;
  }
  if (blockIdx.x == gridDim.x - 1)
    for (int64_t i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) ema[i] = ema[i] * decay + p[i] * (1.f - decay);
}

}  // namespace

DLB_EXPORT int dlb_interp(const float* x0, const float* eps, const float* a, const float* b, float* xt, int64_t B,
                          int64_t per_sample, cudaStream_t stream) {
  DLB_REQUIRE(B > 0 && per_sample > 0, DLB_ERR_SHAPE, "interp: bad shape");
  interp_kernel<<<grid_for(B * per_sample), 256, 0, stream>>>(x0, eps, a, b, xt, per_sample, B * per_sample);
  dlb_count_launch();
  return dlb_check_launch("interp");
}

// pred_dtype 0 = bf16, 1 = fp32. x0 may be null (target = eps). xt/t non-null selects x-prediction.
// loss must be zero-initialised; receives mean over all elements.
DLB_EXPORT int dlb_mse_fwd(const void* pred, int pred_dtype, const float* x0, const float* eps, const float* xt,
                           const float* t, int64_t B, int64_t per_sample, float* loss, cudaStream_t stream) {
  DLB_REQUIRE(B > 0 && per_sample > 0 && ((xt == nullptr) == (t == nullptr)), DLB_ERR_SHAPE, "mse_fwd: bad args");
  const int64_t total = B * per_sample;
  int g = grid_for(total);
  if (g > dlb_num_sms() * 4) g = dlb_num_sms() * 4;
  if (pred_dtype == 0)
    mse_fwd_kernel<bf16><<<g, 256, 0, stream>>>((const bf16*)pred, x0, eps, xt, t, per_sample, total, 1.f / (float)total, loss);
  else
    mse_fwd_kernel<float><<<g, 256, 0, stream>>>((const float*)pred, x0, eps, xt, t, per_sample, total, 1.f / (float)total, loss);
  dlb_count_launch();
  return dlb_check_launch("mse_fwd");
}
DLB_EXPORT int dlb_mse_bwd(const void* pred, int pred_dtype, const float* x0, const float* eps, const float* xt,
                           const float* t, int64_t B, int64_t per_sample, const float* gout, void* dpred,
                           cudaStream_t stream) {
  DLB_REQUIRE(B > 0 && per_sample > 0 && ((xt == nullptr) == (t == nullptr)), DLB_ERR_SHAPE, "mse_bwd: bad args");
  const int64_t total = B * per_sample;
  if (pred_dtype == 0)
    mse_bwd_kernel<bf16><<<grid_for(total), 256, 0, stream>>>((const bf16*)pred, x0, eps, xt, t, per_sample, total, 1.f / (float)total, gout, (bf16*)dpred);
  else
    mse_bwd_kernel<float><<<grid_for(total), 256, 0, stream>>>((const float*)pred, x0, eps, xt, t, per_sample, total, 1.f / (float)total, gout, (float*)dpred);
  dlb_count_launch();
  return dlb_check_launch("mse_bwd");
}

DLB_EXPORT int dlb_repa_cos_fwd(const void* s, const float* z, int64_t R, int E, float coeff, float* loss,
                                cudaStream_t stream) {
  DLB_REQUIRE(R > 0 && E > 0 && E % 8 == 0, DLB_ERR_SHAPE, "repa_cos_fwd: bad shape R=%lld E=%d", (long long)R, E);
  repa_cos_fwd_kernel<<<(unsigned)((R + 7) / 8), 256, 0, stream>>>((const bf16*)s, z, R, E, coeff / (float)R, loss);
  dlb_count_launch();
  return dlb_check_launch("repa_cos_fwd");
}
DLB_EXPORT int dlb_repa_cos_bwd(const void* s, const float* z, int64_t R, int E, float coeff, const float* gout, void* ds,
                                cudaStream_t stream) {
  DLB_REQUIRE(R > 0 && E > 0 && E % 8 == 0, DLB_ERR_SHAPE, "repa_cos_bwd: bad shape");
  repa_cos_bwd_kernel<<<(unsigned)((R + 7) / 8), 256, 0, stream>>>((const bf16*)s, z, R, E, coeff / (float)R, gout, (bf16*)ds);
  dlb_count_launch();
  return dlb_check_launch("repa_cos_bwd");
}

DLB_EXPORT int dlb_sprint_select(const float* scores, int B, int S, int k, int64_t* kept, int32_t* kept32, int32_t* inv,
                                 cudaStream_t stream) {
  DLB_REQUIRE(B > 0 && S > 0 && k > 0 && k <= S && S <= 8192, DLB_ERR_SHAPE, "sprint_select: B=%d S=%d k=%d", B, S, k);
  sprint_select_kernel<<<B, 256, (size_t)S * 8, stream>>>(scores, S, k, kept, kept32, inv);
  dlb_count_launch();
  return dlb_check_launch("sprint_select");
}
DLB_EXPORT int dlb_gather_rows(const void* x, const int64_t* idx, void* out, int B, int S, int k, int d,
                               cudaStream_t stream) {
  DLB_REQUIRE(B > 0 && S > 0 && k > 0 && d > 0 && d % 8 == 0, DLB_ERR_SHAPE, "gather_rows: bad shape");
  gather_rows_kernel<<<grid_for((int64_t)B * k * (d / 8)), 256, 0, stream>>>((const bf16*)x, idx, (bf16*)out, B, S, k, d);
  dlb_count_launch();
  return dlb_check_launch("gather_rows");
}
DLB_EXPORT int dlb_restore_rows(const void* xk, const int32_t* inv, const float* fill, const uint8_t* drop, void* out,
                                int B, int S, int k, int d, cudaStream_t stream) {
  DLB_REQUIRE(B > 0 && S > 0 && k > 0 && d > 0 && d % 8 == 0, DLB_ERR_SHAPE, "restore_rows: bad shape");
  restore_rows_kernel<<<grid_for((int64_t)B * S * (d / 8)), 256, 0, stream>>>((const bf16*)xk, inv, fill, drop, (bf16*)out, B, S, k, d);
  dlb_count_launch();
  return dlb_check_launch("restore_rows");
}
DLB_EXPORT int dlb_restore_rows_bwd(const void* dy, const int64_t* idx, const int32_t* inv, const uint8_t* drop, void* dxk,
                                    float* dfill, int B, int S, int k, int d, cudaStream_t stream) {
  DLB_REQUIRE(B > 0 && S > 0 && k > 0 && d > 0 && d % 8 == 0, DLB_ERR_SHAPE, "restore_rows_bwd: bad shape");
  int g = grid_for((int64_t)B * k * (d / 8));
  if (g > dlb_num_sms() * 2) g = dlb_num_sms() * 2;
  restore_bwd_kernel<<<g, 256, 0, stream>>>((const bf16*)dy, idx, inv, drop, (bf16*)dxk, dfill, B, S, k, d);
  dlb_count_launch();
  return dlb_check_launch("restore_rows_bwd");
}

// v_dtype 0 = bf16, 1 = fp32; vu may be null (no guidance); x0_est and v_out optional.
DLB_EXPORT int dlb_euler_step(const float* x, const void* vc, const void* vu, int v_dtype, float guidance, float t_curr,
                              float t_prev, float* x_prev, float* x0_est, float* v_out, int64_t n, cudaStream_t stream) {
  DLB_REQUIRE(n > 0, DLB_ERR_SHAPE, "euler_step: empty");
  const float dt = t_curr - t_prev;
  if (v_dtype == 0)
    euler_step_kernel<bf16><<<grid_for(n), 256, 0, stream>>>(x, (const bf16*)vc, (const bf16*)vu, guidance, dt, t_curr, x_prev, x0_est, v_out, n);
  else
    euler_step_kernel<float><<<grid_for(n), 256, 0, stream>>>(x, (const float*)vc, (const float*)vu, guidance, dt, t_curr, x_prev, x0_est, v_out, n);
  dlb_count_launch();
  return dlb_check_launch("euler_step");
}

// table: [n_steps, 16] fp32 (columns GS_*), t: [B] int32 timestep indices into it. sampler 0 = DDPM, 1 = DDIM (eta);
// mean_type 0 = epsilon, 1 = xstart, 2 = xprev; var_mode 0 = table, 1 = learned, 2 = learned_range (pred then holds 2 * per_sample
// elements per sample and std_out receives the per-element std; DDIM ignores the variance head); pred_dtype 0 = bf16, 1 = fp32.
DLB_EXPORT int dlb_gaussian_step(const void* pred, int pred_dtype, const float* xt, const float* noise, const float* table,
                                 const int* t, int sampler, int mean_type, int var_mode, int clamp, float eta, int64_t B,
                                 int64_t per_sample, float* x_prev, float* x0, float* mean, float* logprob, float* std_out,
                                 cudaStream_t stream) {
  DLB_REQUIRE(B > 0 && B < 65536 && per_sample > 0, DLB_ERR_SHAPE, "gaussian_step: bad shape B=%lld per_sample=%lld", (long long)B, (long long)per_sample);
  DLB_REQUIRE((sampler == 0 || sampler == 1) && mean_type >= 0 && mean_type <= 2 && var_mode >= 0 && var_mode <= 2, DLB_ERR_UNSUPPORTED,
              "gaussian_step: bad sampler / mean_type / var_mode");
  DLB_REQUIRE(var_mode == 0 || sampler == 1 || std_out != nullptr, DLB_ERR_UNSUPPORTED, "gaussian_step: learned variance needs std_out");
  const int64_t pred_stride = var_mode == 0 ? per_sample : 2 * per_sample;
  int gx = (int)((per_sample + 255) / 256);
  if (gx > 1024) gx = 1024;
  dim3 grid(gx, (unsigned)B);
  if (pred_dtype == 0)
    gaussian_step_kernel<bf16><<<grid, 256, 0, stream>>>((const bf16*)pred, xt, noise, table, t, sampler, mean_type, var_mode, clamp, eta, per_sample,
                                                         pred_stride, x_prev, x0, mean, logprob, std_out);
  else
    gaussian_step_kernel<float><<<grid, 256, 0, stream>>>((const float*)pred, xt, noise, table, t, sampler, mean_type, var_mode, clamp, eta, per_sample,
                                                          pred_stride, x_prev, x0, mean, logprob, std_out);
  dlb_count_launch();
  return dlb_check_launch("gaussian_step");
}

// exactly one of noise / x_prev_in is non-null; v_dtype 0 = bf16, 1 = fp32; scalars as computed by the caller in double
// and rounded to float (python-scalar semantics of the reference expression)
DLB_EXPORT int dlb_euler_maruyama_step(const float* x, const void* v, int v_dtype, const float* noise, const float* x_prev_in,
                                       float c, float one_minus_t, float dt, float t_curr, float stdv, float* x_prev,
                                       float* mean, float* x0_est, float* logprob, int64_t n, cudaStream_t stream) {
  DLB_REQUIRE(n > 0 && ((noise == nullptr) != (x_prev_in == nullptr)), DLB_ERR_SHAPE, "euler_maruyama_step: need exactly one of noise / x_prev");
  if (v_dtype == 0)
    euler_maruyama_kernel<bf16><<<grid_for(n), 256, 0, stream>>>(x, (const bf16*)v, noise, x_prev_in, c, one_minus_t, dt, t_curr, stdv, x_prev, mean, x0_est, logprob, n);
  else
    euler_maruyama_kernel<float><<<grid_for(n), 256, 0, stream>>>(x, (const float*)v, noise, x_prev_in, c, one_minus_t, dt, t_curr, stdv, x_prev, mean, x0_est, logprob, n);
  dlb_count_launch();
  return dlb_check_launch("euler_maruyama_step");
}

// Hyper-parameters arrive as doubles (python floats) and the bias corrections are evaluated in double exactly as
// torch.optim.AdamW's single-tensor path does (step_size = lr / (1 - beta1^step), sqrt(1 - beta2^step)).
DLB_EXPORT int dlb_adamw_step(float* p, const float* g, float* m, float* v, void* shadow, float* ema, float ema_decay,
                              const uint8_t* chunk_active, int64_t n, double lr, double beta1, double beta2, double eps, double wd,
                              int64_t step, float grad_scale, cudaStream_t stream) {
  DLB_REQUIRE(n > 0 && step >= 1, DLB_ERR_SHAPE, "adamw_step: bad args");
  const double bc1 = 1.0 - pow(beta1, (double)step);
  const double bc2_sqrt = sqrt(1.0 - pow(beta2, (double)step));
  DLB_REQUIRE(((uintptr_t)p % 16) == 0 && ((uintptr_t)g % 16) == 0 && ((uintptr_t)m % 16) == 0 && ((uintptr_t)v % 16) == 0 &&
                  ((uintptr_t)shadow % 8) == 0 && ((uintptr_t)ema % 16) == 0,
              DLB_ERR_ALIGN, "adamw_step: buffers must be 16-byte aligned");
  adamw_kernel<<<grid_for((n + 3) / 4), 256, 0, stream>>>(p, g, m, v, (bf16*)shadow, ema, ema_decay, chunk_active, n, (float)(1.0 - lr * wd),
                                                          (float)beta1, (float)beta2, (float)(1.0 - beta1), (float)(1.0 - beta2), (float)eps, (float)(lr / bc1), (float)bc2_sqrt, grad_scale);
  dlb_count_launch();
  return dlb_check_launch("adamw_step");
}

DLB_EXPORT int dlb_ema_lerp(float* ema, const float* p, float decay, int64_t n, cudaStream_t stream) {
  DLB_REQUIRE(n > 0 && ((uintptr_t)ema % 16) == 0 && ((uintptr_t)p % 16) == 0, DLB_ERR_ALIGN, "ema_lerp: bad args");
  ema_lerp_kernel<<<grid_for((n + 3) / 4), 256, 0, stream>>>(ema, p, decay, n);
  dlb_count_launch();
  return dlb_check_launch("ema_lerp");
}
