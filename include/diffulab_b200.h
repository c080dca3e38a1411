/*
 * libdiffulab_b200 — C ABI of the B200-native (sm_100a) kernels behind DiffuLab's denoiser training / sampling
 * hot path. Plain pointers and sizes only; no torch types. The caller owns every buffer (device pointers unless
 * stated), passes the CUDA stream to launch on, and keeps tensors contiguous in the documented row-major layout.
 *
 * The reference (LouisRouss/DiffuLab) is pure Python: each entry point below replaces a PyTorch library-call site
 * of the reference, cited as file:line relative to /root/reference/src/diffulab. INTEGRATION.md shows the ctypes
 * binding a maintainer adds on the reference side.
 *
 * Conventions
 *   return value : 0 on success; negative DLB_ERR_* on bad shape / alignment / unsupported request; positive =
 *                  cudaError_t of a failed launch. Never throws. dlb_last_error() returns the message (thread local).
 *   dtypes       : "bf16" = __nv_bfloat16 storage (void*), fp32 = float. dtype selector ints: 0 = bf16, 1 = fp32.
 *   streams      : every call enqueues on `stream` and returns immediately; no allocation, no host sync
 *                  (CUDA-graph capturable). The GEMM encodes its TMA descriptors on the host per call.
 *   no fallback  : there is no CPU path; non-sm_100 devices are rejected by dlb_device_check().
 */
#ifndef DIFFULAB_B200_H
#define DIFFULAB_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* dlb_stream_t; /* cudaStream_t */

enum {
  DLB_OK = 0,
  DLB_ERR_SHAPE = -1,
  DLB_ERR_ALIGN = -2,
  DLB_ERR_UNSUPPORTED = -3,
  DLB_ERR_DRIVER = -4
};

/* ---- library ------------------------------------------------------------------------------------------- */
int dlb_version(void);                 /* 100 = 0.1.0 */
const char* dlb_last_error(void);
int dlb_device_check(void);            /* 0 iff the current device is compute capability 10.x */
long long dlb_launch_count(void);      /* kernels launched by this library since load / last reset */
void dlb_reset_launch_count(void);
/* Number of SMs the persistent kernels (GEMMs, attention) size their grids to; 0 = all. The data-parallel reducer lowers it
 * by the collective's CTA count while gradient buckets are in flight behind backward (the role DDP's bucketed overlap plays
 * under accelerate in the reference, base_trainer.py:111-123) so that NCCL does not time-slice with a persistent CTA.
 * Returns the previous budget. */
int dlb_set_sm_budget(int sms);

/* ---- dense contractions: tcgen05 / TMEM / TMA GEMM -------------------------------------------------------
 * C[M,N] (+)= A * B^T (+ bias[N]); bf16 operands, fp32 accumulate.
 *   a_mn_major = 0: A is row-major [M,K] (ld = lda);   1: A is row-major [K,M]  (read "MN-major").
 *   b_mn_major = 0: B is row-major [N,K] (ld = ldb);   1: B is row-major [K,N].
 *   out_mode 0: C bf16 store; 1: C fp32 store; 2: C fp32 += (TMA reduce-add; required for split_k > 1).
 *   split_k / tile_n: 0 = chosen by the library's cost model; tile_n in {64,128,192,256}.
 * Replaces nn.Linear / nn.Conv2d(k=s=p) forward (networks/denoisers/mmdit.py:70-73, 260-264, 539, 697-699;
 * networks/utils/nn.py:528) and their autograd dgrad (b_mn_major=1) and wgrad (both MN-major, out_mode 2).
 * Requirements: pointers 16-byte aligned; lda, ldb multiples of 8 elements; ldc * sizeof(C elem) multiple of 16. */
int dlb_gemm_bf16(const void* A, const void* B, void* C, const float* bias, int64_t M, int64_t N, int64_t K,
                  int64_t lda, int64_t ldb, int64_t ldc, int a_mn_major, int b_mn_major, int out_mode, int split_k,
                  int tile_n, dlb_stream_t stream);

/* CTA-pair (tcgen05 cta_group::2, 256 x tile_n tiles, cluster of two CTAs) version of dlb_gemm_bf16: same contract and
 * results; tile_n in {128, 256} (192 as well when B is K-major); split_k >= 1 (no automatic choice). dlb_gemm_bf16 dispatches to it where it is faster. */
int dlb_gemm2_bf16(const void* A, const void* B, void* C, const float* bias, int64_t M, int64_t N, int64_t K, int64_t lda,
                   int64_t ldb, int64_t ldc, int a_mn_major, int b_mn_major, int out_mode, int split_k, int tile_n,
                   dlb_stream_t stream);

/* fc1 of the packed-SwiGLU MLP with the activation fused into the epilogue: H[M,2F] = A[M,K] W[2F,K]^T (+ bias[2F]) in bf16
 * (the pre-activation saved for backward) and ACT[M,F] = silu(H[:, :F]) * H[:, F:], one launch. Replaces nn.Linear +
 * PackedSwiGLU.forward (reference networks/denoisers/mmdit.py:260-264, networks/utils/nn.py:478-486). F % 128 == 0. */
int dlb_gemm_swiglu_bf16(const void* A, const void* W, const float* bias, void* H, void* ACT, int64_t M, int64_t F, int64_t K,
                         int64_t lda, int64_t ldw, int64_t ldh, int64_t ldact, dlb_stream_t stream);

/* dgrad of the MLP down-projection with the SwiGLU backward fused into the epilogue: d(act) = dY[M,D] W2[D,F] stays in
 * TMEM / registers; dH[M,2F] = [d(act) * g * silu'(a) | d(act) * silu(a)] with (a, g) = the saved H[M,2F] (TMA-prefetched
 * per tile). Replaces the autograd mirror of nn.Linear(F, D) + PackedSwiGLU (reference nn.py:478-486). F % 128 == 0. */
int dlb_gemm_swiglu_bwd_bf16(const void* dY, const void* W2, const void* H, void* dH, int64_t M, int64_t F, int64_t D,
                             int64_t lddy, int64_t ldw2, int64_t ldh, int64_t lddh, dlb_stream_t stream);

/* ---- LayerNorm (+affine) + adaLN modulate ----------------------------------------------------------------
 * y = (LN(x) * w + b) * (1 + scale) + shift; x,y bf16 [R,d]; w,b fp32 [d] or both NULL; scale/shift bf16 rows of
 * a [G, k*d] adaLN output (row stride mod_ld), row r uses modulation row r / rows_per_mod (1 = per token).
 * mean/rstd fp32 [R] are saved for backward. Replaces modulate(nn.LayerNorm(x), ...) mmdit.py:257-259,299,305,538-547;
 * nn.py:539-540. */
int dlb_ln_modulate_fwd(const void* x, const float* w, const float* b, const void* scale, const void* shift,
                        int64_t mod_ld, int64_t rows_per_mod, void* y, float* mean, float* rstd, int64_t R, int d,
                        float eps, dlb_stream_t stream);
/* dx = dLN(dy) (+ dres); per-sample mode (per_token = 0): dscale/dshift fp32 rows (stride dmod_ld) accumulated
 * atomically per group; per-token mode: written as bf16 rows (stride dtok_ld). dw/db fp32 [d] accumulated. */
int dlb_ln_modulate_bwd(const void* dy, const void* x, const float* mean, const float* rstd, const float* w,
                        const float* b, const void* scale, int64_t mod_ld, int64_t groups, int64_t rows_per_group,
                        int per_token, const void* dres, void* dx, float* dscale, float* dshift, int64_t dmod_ld,
                        void* dscale_tok, void* dshift_tok, int64_t dtok_ld, float* dw, float* db, int d,
                        dlb_stream_t stream);

/* ---- gated residual: out = x + (a1 [+ a2]) * gate   (mmdit.py:296-307, 435-457, 524-531) ------------------- */
int dlb_gate_residual_fwd(const void* x, const void* a1, const void* a2, const void* gate, int64_t gate_ld,
                          int64_t rows_per_mod, void* out, int64_t R, int d, dlb_stream_t stream);
int dlb_gate_residual_bwd(const void* dout, const void* a1, const void* a2, const void* gate, int64_t gate_ld,
                          int64_t groups, int64_t rows_per_group, int per_token, void* da, float* dgate,
                          int64_t dgate_ld, void* dgate_tok, int64_t dtok_ld, int d, dlb_stream_t stream);

/* ---- packed SwiGLU: out[R,F] = silu(h[:, :F]) * h[:, F:]   (nn.py:478-486) --------------------------------- */
int dlb_swiglu_fwd(const void* h, void* out, int64_t R, int F, dlb_stream_t stream);
int dlb_swiglu_bwd(const void* dout, const void* h, void* dh, int64_t R, int F, dlb_stream_t stream);

/* ---- QK-RMSNorm over the full inner dim + N-D interleaved-pair RoPE (nn.py:262-307, 331-400, 423-475) ------
 * qkv bf16 [R, >=2d] packed (q | k | ...), out bf16 [R, 2d] (q | k) normalised, scaled and rotated.
 * cs_t uint32 [P, rot_half]: packed bf16x2 (low = cos, high = sin) written by dlb_rope_table (the reference casts
 * its fp32 tables to the activation dtype); table row of token r = pos_idx[r] if given else
 * pos_offset + r % tokens_per_sample.
 * rrms fp32 [R,2] (reciprocal RMS of q and k rows) is written by the forward and consumed by the backward;
 * dsq/dsk (fp32 [d], accumulated) may both be NULL when the scales are frozen. */
int dlb_qknorm_rope_fwd(const void* qkv, int64_t ld_in, const float* sq, const float* sk, const uint32_t* cs_t,
                        int rot_half, const int32_t* pos_idx, int pos_offset,
                        int tokens_per_sample, int hd, void* out, int64_t ld_out, float* rrms, int64_t R, int d,
                        float eps, dlb_stream_t stream);
int dlb_qknorm_rope_bwd(const void* dqk, int64_t ld_dqk, const void* qkv, int64_t ld_in, const float* sq,
                        const float* sk, const uint32_t* cs_t, int rot_half, const int32_t* pos_idx,
                        int pos_offset, int tokens_per_sample, int hd, const float* rrms, void* dqkv, int64_t ld_out,
                        float* dsq, float* dsk, int64_t R, int d, dlb_stream_t stream);
/* cos/sin tables of get_cos_sin_ndim_grid (fp64 angles -> fp32) and/or the packed bf16x2 table; each output optional
 * (cos_t and sin_t together). pos int32 [P, n_axes]. */
int dlb_rope_table(const int32_t* pos, int n_axes, const int32_t* axis_of_pair, const int32_t* local_of_pair,
                   const int32_t* axis_dim, double base, float* cos_t, float* sin_t, uint32_t* cs_t, int64_t P,
                   int rot_half, dlb_stream_t stream);

/* ---- joint attention over 1 or 2 segments (text rows first, then image rows) ------------------------------
 * softmax(q k^T * scale + key_padding_mask) v per (sample, head); replaces F.scaled_dot_product_attention and the
 * torch.cat's around it (mmdit.py:92-98, 184-204). kmask uint8 [B, mask_len] covers the first mask_len keys. */
typedef struct dlb_attn_seg {
  const void* q; const void* k; const void* v;  /* bf16 rows [B*len, ...], row strides ldq / ldk / ldv           */
  void* o;                                      /* bf16 [B*len, H*hd] (ldo): forward output, backward input      */
  const void* dout;                             /* backward: grad of o (lddo)                                    */
  void* dq; void* dk; void* dv;                 /* backward outputs (lddq / lddk / lddv)                         */
  int64_t ldq, ldk, ldv, ldo, lddo, lddq, lddk, lddv;
  int32_t len;                                  /* tokens of this segment per sample                             */
} dlb_attn_seg;
int dlb_attn_fwd(const dlb_attn_seg* segs, int nseg, float* lse, const uint8_t* kmask, int mask_len, int B, int H,
                 int hd, float scale, dlb_stream_t stream);
/* tcgen05 / TMEM implementation of dlb_attn_fwd (same contract) */
int dlb_attn_fwd_tc(const dlb_attn_seg* segs, int nseg, float* lse, const uint8_t* kmask, int mask_len, int B, int H,
                    int hd, float scale, dlb_stream_t stream);
int dlb_attn_bwd(const dlb_attn_seg* segs, int nseg, const float* lse, float* dsum, const uint8_t* kmask,
                 int mask_len, int B, int H, int hd, float scale, dlb_stream_t stream);

/* tcgen05 / TMEM implementation of dlb_attn_bwd (same contract; head_dim <= 128) */
int dlb_attn_bwd_tc(const dlb_attn_seg* segs, int nseg, const float* lse, float* dsum, const uint8_t* kmask,
                    int mask_len, int B, int H, int hd, float scale, dlb_stream_t stream);

/* development aid: dev_buf = device buffer of 3 x 64 int64 receiving one CTA's SM-clock timeline per tcgen05 attention
 * kernel (slots 0..63 forward, 64..127 dq, 128..191 dkv); NULL switches tracing off (the default) */
int dlb_attn_set_trace(long long* dev_buf);

/* exact (erf) GELU forward / backward on bf16 (nn.GELU default; PerceiverResampler feed-forward, perceiver_resampler.py:75-77) */
int dlb_gelu_fwd(const void* x, void* y, int64_t n, dlb_stream_t stream);
int dlb_gelu_bwd(const void* dy, const void* x, void* dx, int64_t n, dlb_stream_t stream);
/* N-D interleaved-pair RoPE on the heads of a packed bf16 [R, ld] tensor, no norm / scale (key-only rotation of the
 * PerceiverResampler, perceiver_resampler.py:13-56); inverse = 1 applies the transposed rotation (backward) */
int dlb_rope_apply(const void* x, int64_t ld_in, void* y, int64_t ld_out, const uint32_t* cs_t, int rot_half, const int32_t* pos_idx,
                   int pos_offset, int tokens_per_sample, int hd, int d, int64_t R, int inverse, dlb_stream_t stream);

/* ---- glue ------------------------------------------------------------------------------------------------ */
int dlb_cast_f32_bf16(const float* in, void* out, int64_t rows, int64_t cols, int64_t ld_out, dlb_stream_t stream);
int dlb_cast_bf16_f32(const void* in, float* out, int64_t n, dlb_stream_t stream);
int dlb_add_bf16(const void* a, const void* b, void* out, int64_t n, dlb_stream_t stream);
int dlb_silu_fwd(const void* x, int in_dtype, void* y, int64_t n, dlb_stream_t stream);
int dlb_silu_bwd(const void* dy, int dy_dtype, const void* x, int x_dtype, void* dx, int dx_dtype, int64_t n,
                 dlb_stream_t stream);
/* DDT decoder conditioning silu(x[b,n,:] + v[b,:]) (denoisers/ddt.py:421-422) */
int dlb_bias_silu_fwd(const void* x, const void* v, void* out, int64_t B, int64_t rows_per_sample, int d,
                      dlb_stream_t stream);
int dlb_bias_silu_bwd(const void* dy, const void* x, const void* v, void* dx, float* dv, int64_t B,
                      int64_t rows_per_sample, int d, dlb_stream_t stream);
/* timestep_embedding (nn.py:91-114): [cos(t f) | sin(t f)] -> bf16 [B, dim] */
int dlb_timestep_embed(const float* t, void* out, int B, int dim, float max_period, dlb_stream_t stream);
/* emb = float(te) + table[labels]; emb_silu = bf16(silu(emb))   (mmdit.py:866-868; nn.py:531) */
int dlb_cond_combine(const void* te, const float* table, const int64_t* labels, float* emb, void* emb_silu, int B,
                     int E, dlb_stream_t stream);
int dlb_embedding_bwd(const float* g, const int64_t* labels, float* dtable, int B, int E, dlb_stream_t stream);
/* patchify = im2col of Conv2d(k=s=p, bias=False) (mmdit.py:697-699, 757-765); unpatchify mmdit.py:778-786 */
int dlb_patchify(const float* x, void* out, int B, int C, int H, int W, int p, int Kp, dlb_stream_t stream);
int dlb_unpatchify(const void* tok, int64_t ld, void* img, int out_dtype, int B, int C, int H, int W, int p,
                   dlb_stream_t stream);
int dlb_patchify_grad(const void* img, int in_dtype, void* tok, int64_t ld, int B, int C, int H, int W, int p,
                      dlb_stream_t stream);
int dlb_colsum(const void* in, int in_dtype, int64_t ld, float* out, int64_t R, int Cn, dlb_stream_t stream);

/* ---- formalisation ----------------------------------------------------------------------------------------
 * x_t = a_b x0 + b_b eps (Flow.add_noise diffuse/modelizations/flow.py:382-408 with a = 1-t, b = t;
 * GaussianDiffusion.add_noise gaussian_diffusion.py:313-341 with a = sqrt(alpha_bar), b = sqrt(1-alpha_bar)). */
int dlb_interp(const float* x0, const float* eps, const float* a, const float* b, float* xt, int64_t B,
               int64_t per_sample, dlb_stream_t stream);
/* loss += mean((target - v)^2), target = eps - x0 (x0 != NULL) or eps; v = pred or (xt - pred)/t (flow.py:300-308) */
int dlb_mse_fwd(const void* pred, int pred_dtype, const float* x0, const float* eps, const float* xt, const float* t,
                int64_t B, int64_t per_sample, float* loss, dlb_stream_t stream);
int dlb_mse_bwd(const void* pred, int pred_dtype, const float* x0, const float* eps, const float* xt, const float* t,
                int64_t B, int64_t per_sample, const float* gout, void* dpred, dlb_stream_t stream);
/* REPA: loss += coeff * (1 - mean_rows cos(s, z)) (training/losses/repa.py:183-185); s bf16, z fp32 [R,E] */
int dlb_repa_cos_fwd(const void* s, const float* z, int64_t R, int E, float coeff, float* loss, dlb_stream_t stream);
int dlb_repa_cos_bwd(const void* s, const float* z, int64_t R, int E, float coeff, const float* gout, void* ds,
                     dlb_stream_t stream);
/* SPRINT (networks/denoisers/sprint.py:317-387): kept = ascending indices of the k largest scores per row
 * (ties -> larger index); inv[b,s] = slot or -1. gather / restore (mask-token fill, per-sample path drop). */
int dlb_sprint_select(const float* scores, int B, int S, int k, int64_t* kept, int32_t* kept32, int32_t* inv,
                      dlb_stream_t stream);
int dlb_gather_rows(const void* x, const int64_t* idx, void* out, int B, int S, int k, int d, dlb_stream_t stream);
int dlb_restore_rows(const void* xk, const int32_t* inv, const float* fill, const uint8_t* drop, void* out, int B,
                     int S, int k, int d, dlb_stream_t stream);
int dlb_restore_rows_bwd(const void* dy, const int64_t* idx, const int32_t* inv, const uint8_t* drop, void* dxk,
                         float* dfill, int B, int S, int k, int d, dlb_stream_t stream);
/* Euler step + optional CFG combine (samplers/flow/euler.py:37-39; flow.py:256-260) */
int dlb_euler_step(const float* x, const void* vc, const void* vu, int v_dtype, float guidance, float t_curr,
                   float t_prev, float* x_prev, float* x0_est, float* v_out, int64_t n, dlb_stream_t stream);

/* One reverse step of the Gaussian samplers, fused: replaces DDPM.step / DDIM.step
 * (reference diffuse/samplers/gaussian_diffusion/ddpm.py `_get_p_mean_var` + `step`, ddim.py:27-103).
 * table [n_steps,16] fp32 per-timestep coefficients (columns: 1/sqrt(ab), sqrt(1-ab)/sqrt(ab), 1/c1, c2/c1, c1, c2, var,
 * exp(0.5 logvar), [t>0], sqrt(1/ab-1), sqrt(ab_prev), sqrt((1-ab_prev)/(1-ab)), sqrt(1-ab/ab_prev), ab_prev,
 * posterior_log_variance_clipped, log(beta));
 * t [B] int32; sampler 0 DDPM / 1 DDIM(eta); mean_type 0 epsilon / 1 xstart / 2 xprev; var_mode 0 fixed (table) /
 * 1 learned / 2 learned_range (ddpm.py:213-223: pred holds [mean prediction | variance head], 2*per_sample elements per
 * sample as torch.chunk(.., 2, dim=1) splits it; std_out [B*per_sample] receives sqrt(max(var, 1e-20)), DDPM only);
 * logprob nullable; std_out nullable when var_mode == 0. */
int dlb_gaussian_step(const void* pred, int pred_dtype, const float* xt, const float* noise, const float* table,
                      const int* t, int sampler, int mean_type, int var_mode, int clamp, float eta, int64_t B,
                      int64_t per_sample, float* x_prev, float* x0, float* mean, float* logprob, float* std_out,
                      dlb_stream_t stream);

/* Euler-Maruyama flow step with log-probability (reference diffuse/samplers/flow/euler_meruyama.py:24-57), one launch.
 * c = sigma^2/(2 t_curr), stdv = sigma sqrt(t_curr - t_prev); exactly one of noise / x_prev_in is non-null. */
int dlb_euler_maruyama_step(const float* x, const void* v, int v_dtype, const float* noise, const float* x_prev_in, float c,
                            float one_minus_t, float dt, float t_curr, float stdv, float* x_prev, float* mean,
                            float* x0_est, float* logprob, int64_t n, dlb_stream_t stream);
/* ---- data-parallel gradient reduction over NVLink peer memory (replaces the NCCL all-reduce kernels DDP runs under the
 * reference's Accelerate wrapper, training/trainers/common.py:103-109; orchestration in diffulab_b200/training.py GradReducer) ----
 * dlb_reduce_pieces: own[0..n) = scale * sum over ranks (fixed order) of this rank's piece; the world-1 peer copies were pulled by
 * the copy engines into `staged` (slot s, stride stage_stride floats, holds rank (rank + 1 + s) % world). max_ctas 0 = no cap.
 * dlb_multimem_allreduce: mc_piece = MULTICAST address of this rank's piece of a symmetric buffer; multimem.ld_reduce (sum in the
 * NVSwitch) -> * scale -> multimem.st to every rank, on `ctas` CTAs. Cross-rank ordering is the caller's (stream barriers). */
int dlb_reduce_pieces(float* own, const float* staged, int64_t stage_stride, int world, int rank, int64_t n, float scale,
                      int max_ctas, dlb_stream_t stream);
int dlb_multimem_allreduce(void* mc_piece, int64_t n, float scale, int ctas, dlb_stream_t stream);

/* torch.optim.AdamW step over a flat buffer (+ bf16 shadow, + optional EMA lerp ema = ema*decay + p*(1-decay))
 * (base_trainer.py:149-153). Hyper-parameters are doubles (python floats); bias corrections are evaluated in double as
 * torch does. chunk_active (nullable): one byte per 64 elements, 0 = skip (parameters whose .grad is None in torch). */
int dlb_adamw_step(float* p, const float* g, float* m, float* v, void* shadow, float* ema, float ema_decay,
                   const uint8_t* chunk_active, int64_t n, double lr, double beta1, double beta2, double eps, double wd,
                   int64_t step, float grad_scale, dlb_stream_t stream);
/* ema_pytorch EMA.update_moving_average over a flat buffer: ema.lerp_(p, 1 - decay) (base_trainer.py:152-153) */
int dlb_ema_lerp(float* ema, const float* p, float decay, int64_t n, dlb_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* DIFFULAB_B200_H */
