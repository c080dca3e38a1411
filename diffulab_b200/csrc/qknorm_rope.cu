// QK-RMSNorm over the full inner dimension + N-D interleaved-pair RoPE, applied to the packed qkv projection.
// One warp per token row; q and k rows live in registers; pairs (2j, 2j+1) sit in the same 16-byte vector.
//
// Reference: QKNorm/RMSNorm networks/utils/nn.py:423-475 (norm over all heads jointly, eps 1e-6, fp32 math, cast
// back, learnable scale, cast to v's dtype) and RotaryPositionalEmbeddingNDim nn.py:331-400 (cos/sin cast to the
// activation dtype, rotation evaluated in that dtype) as called from DiTAttention.forward mmdit.py:81-89.
#include "common.cuh"
#include "ptx.cuh"
#include "rowwise_lean.cuh"
#include <cstdlib>

namespace {
typedef __nv_bfloat16 bf16;

struct RopeArgs {
  const uint32_t* cs_t;  // [positions, rot_half] packed bf16x2: low half = cos, high half = sin (the reference casts the
                         // fp32 tables to the activation dtype before rotating, nn.py:378-379)
  const int32_t* pos_idx;  // optional [R]: table row per token (SPRINT-gathered sequences); else offset + row % tps
  int rot_half;            // rotary pairs per head
  int pos_offset;
  int tokens_per_sample;
  int hd;
};

__device__ __forceinline__ int rope_pos(const RopeArgs& ra, int64_t row) {
  return ra.pos_idx ? ra.pos_idx[row] : ra.pos_offset + (int)(row % ra.tokens_per_sample);
}

__device__ __forceinline__ float2 cs_unpack(uint32_t cs) {  // (cos, sin) as fp32 (exactly the bf16 values)
  return make_float2(__uint_as_float(cs << 16), __uint_as_float(cs & 0xffff0000u));
}

template <int VPL>
__global__ void __launch_bounds__(128, 6)
qknorm_rope_fwd_kernel(const bf16* __restrict__ qkv, int64_t ld_in, const float* __restrict__ sq,
                       const float* __restrict__ sk, RopeArgs ra, bf16* __restrict__ out, int64_t ld_out,
                       float* __restrict__ rrms_out, int64_t R, int d, float eps) {
  const int lane = threadIdx.x & 31;
  const int nv = d >> 3;
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < R; row += nwarps) {
    const uint32_t* csr = ra.cs_t + (int64_t)rope_pos(ra, row) * ra.rot_half;
    // q and k rows are fetched together (2 * VPL independent 16-byte loads in flight) and kept packed
    bf16x8 xp[2][VPL];
#pragma unroll
    for (int which = 0; which < 2; ++which)
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        const int v = lane + 32 * i;
        if (v < nv) xp[which][i] = ld8(qkv + row * ld_in + which * d + v * 8);
      }
#pragma unroll
    for (int which = 0; which < 2; ++which) {
      const float* sc = which ? sk : sq;
      float ss = 0.f;
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        if (lane + 32 * i < nv) {
          float f[8];
          unpack8(xp[which][i], f);
#pragma unroll
          for (int j = 0; j < 8; ++j) ss += f[j] * f[j];
        }
      }
      const float rrms = rsqrtf(warp_sum(ss) / d + eps);
      if (lane == 0 && rrms_out) rrms_out[row * 2 + which] = rrms;
      bf16* dst = out + row * ld_out + which * d;
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        const int v = lane + 32 * i;
        if (v < nv) {
          const int c = v * 8;
          const int cl = c % ra.hd;  // channel inside the head; hd % 8 == 0 keeps a vector inside one head
          float y[8], scv[8];
          unpack8(xp[which][i], y);
          *reinterpret_cast<float4*>(scv) = __ldg(reinterpret_cast<const float4*>(sc + c));
          *reinterpret_cast<float4*>(scv + 4) = __ldg(reinterpret_cast<const float4*>(sc + c + 4));
          bf16x8 o;
#pragma unroll
          for (int j = 0; j < 8; j += 2) {
            // bf16(bf16(x * rrms) * scale) on a pair, then the rotation in native bf16x2 arithmetic: every product
            // and the final add/sub round to bf16 exactly like the reference's bf16 tensor ops
            float2 xn = unpack_bf16x2(pack_bf16x2(y[j] * rrms, y[j + 1] * rrms));
            uint32_t yp = pack_bf16x2(xn.x * scv[j], xn.y * scv[j + 1]);
            const int pj = (cl + j) >> 1;
            if (pj < ra.rot_half) {
              const uint32_t cs = __ldg(csr + pj);
              const uint32_t cc = __byte_perm(cs, cs, 0x1010), sn = __byte_perm(cs, cs, 0x3232);  // (c,c) (s,s)
              const uint32_t ysw = __byte_perm(yp, yp, 0x1032);                                   // (o, e)
              __nv_bfloat162 t1 = __hmul2(*reinterpret_cast<__nv_bfloat162*>(&yp), *reinterpret_cast<const __nv_bfloat162*>(&cc));
              __nv_bfloat162 t2 = __hmul2(*reinterpret_cast<const __nv_bfloat162*>(&ysw), *reinterpret_cast<const __nv_bfloat162*>(&sn));
              uint32_t t2n = *reinterpret_cast<uint32_t*>(&t2) ^ 0x00008000u;  // (-o*s, e*s)
              __nv_bfloat162 r = __hadd2(t1, *reinterpret_cast<__nv_bfloat162*>(&t2n));
              yp = *reinterpret_cast<uint32_t*>(&r);
            }
            o.u[j >> 1] = yp;
          }
          st8(dst + c, o);
        }
      }
    }
  }
}

// backward, two kernels (same split as the LayerNorm backward):
//  rows: dqk (grad wrt normalised+rotated q,k) -> dq, dk written into the packed dqkv buffer (one warp per token);
//  cols: column-accumulated gradients of the learnable RMS scales (one thread per 8-channel vector).
template <int VPL>
__global__ void __launch_bounds__(128, 5)
qknorm_rope_bwd_rows_kernel(const bf16* __restrict__ dqk, int64_t ld_dqk, const bf16* __restrict__ qkv, int64_t ld_in,
                            const float* __restrict__ sq, const float* __restrict__ sk, RopeArgs ra,
                            const float* __restrict__ rrms_in, bf16* __restrict__ dqkv, int64_t ld_out, int64_t R, int d) {
  const int lane = threadIdx.x & 31;
  const int nv = d >> 3;
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < R; row += nwarps) {
    const uint32_t* csr = ra.cs_t + (int64_t)rope_pos(ra, row) * ra.rot_half;
#pragma unroll 1
    for (int which = 0; which < 2; ++which) {
      const float* sc = which ? sk : sq;
      const bf16* src = qkv + row * ld_in + which * d;
      const bf16* gsrc = dqk + row * ld_dqk + which * d;
      const float rrms = rrms_in[row * 2 + which];
      bf16x8 xp[VPL], gp[VPL];
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        const int v = lane + 32 * i;
        if (v < nv) { xp[i] = ld8(src + v * 8); gp[i] = ld8(gsrc + v * 8); }
      }
      // grad wrt the normalised value: transpose of the rotation, times the learnable scale
      auto grad_norm = [&](int i, float* gn) {
        const int c = (lane + 32 * i) * 8;
        const int cl = c % ra.hd;
        float scv[8];
        unpack8(gp[i], gn);
        *reinterpret_cast<float4*>(scv) = __ldg(reinterpret_cast<const float4*>(sc + c));
        *reinterpret_cast<float4*>(scv + 4) = __ldg(reinterpret_cast<const float4*>(sc + c + 4));
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
          const int pj = (cl + j) >> 1;
          if (pj < ra.rot_half) {
            const float2 cs = cs_unpack(__ldg(csr + pj));
            const float ge = gn[j], go = gn[j + 1];
            gn[j] = ge * cs.x + go * cs.y;
            gn[j + 1] = go * cs.x - ge * cs.y;
          }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) gn[j] *= scv[j];
      };
      float dot = 0.f;
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        if (lane + 32 * i < nv) {
          float xn[8], gn[8];
          unpack8(xp[i], xn);
          grad_norm(i, gn);
#pragma unroll
          for (int j = 0; j < 8; ++j) dot += gn[j] * (xn[j] * rrms);
        }
      }
      dot = warp_sum(dot) / d;
      bf16* dst = dqkv + row * ld_out + which * d;
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        const int v = lane + 32 * i;
        if (v < nv) {
          float xn[8], gn[8], o[8];
          unpack8(xp[i], xn);
          grad_norm(i, gn);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = rrms * (gn[j] - xn[j] * rrms * dot);
          st8(dst + v * 8, pack8(o));
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// One-pass tiled backward: a cp.async.bulk ring feeds 2-row tiles {dqk rows, qkv rows, rrms}; warp (row, q|k) owns one
// half-row, produces the raw-projection gradient AND accumulates the gradient of the learnable RMS scale for its fixed
// columns in registers (the separate column kernel re-read both tensors: 302 MB per call). Flushed once per CTA.
// ---------------------------------------------------------------------------------------------------------
constexpr int QB_ROWS = 2, QB_STAGES = 3, QB_WARPS = 4;
template <int VPL>
__global__ void __launch_bounds__((QB_WARPS + 1) * 32, 2)
qknorm_rope_bwd_tile_kernel(const bf16* __restrict__ dqk, int64_t ld_dqk, const bf16* __restrict__ qkv, int64_t ld_in,
                            const float* __restrict__ sq, const float* __restrict__ sk, RopeArgs ra, const float* __restrict__ rrms_in,
                            bf16* __restrict__ dqkv, int64_t ld_out, float* __restrict__ dsq, float* __restrict__ dsk, int64_t R, int d,
                            int tiles_per_cta) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int row_bytes = 2 * d * 2;                          // q | k halves of one token
  const int stage_bytes = 2 * QB_ROWS * row_bytes + 16;    // {dqk rows, qkv rows} + rrms[QB_ROWS][2]
  uint8_t* sOut = smem + QB_STAGES * stage_bytes;          // [2][QB_ROWS][2d] bf16
  float* sScale = reinterpret_cast<float*>(sOut + 2 * QB_ROWS * row_bytes);  // [2][d] learnable scales (q, k)
  float* sAcc = sScale + 2 * d;                             // [2][d] flush buffer
  __shared__ uint64_t full[QB_STAGES], empty[QB_STAGES];
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int64_t ntiles = (R + QB_ROWS - 1) / QB_ROWS;
  const int64_t t0 = (int64_t)blockIdx.x * tiles_per_cta;
  const int64_t t1 = t0 + tiles_per_cta < ntiles ? t0 + tiles_per_cta : ntiles;
  if (tid == 0) {
    for (int i = 0; i < QB_STAGES; ++i) { ptx::mbar_init(&full[i], 1); ptx::mbar_init(&empty[i], QB_WARPS); }
    ptx::fence_mbar_init();
  }
  if (tid < QB_WARPS * 32)
    for (int c = tid; c < 2 * d; c += QB_WARPS * 32) sScale[c] = c < d ? sq[c] : sk[c - d];
  __syncthreads();
  if (warp == QB_WARPS) {
    // ---------------- producer ----------------
    int st = 0;
    uint32_t ph = 0;
    for (int64_t t = t0; t < t1; ++t) {
      ptx::mbar_wait(&empty[st], ph ^ 1);
      if (ptx::elect_one()) {
        const int64_t r0 = t * QB_ROWS;  // R % QB_ROWS == 0 (launcher)
        uint8_t* sb = smem + st * stage_bytes;
        ptx::mbar_expect_tx(&full[st], 2 * QB_ROWS * row_bytes + 16);
#pragma unroll
        for (int rr = 0; rr < QB_ROWS; ++rr) {
          ptx::bulk_load_1d(sb + rr * row_bytes, dqk + (r0 + rr) * ld_dqk, row_bytes, &full[st]);
          ptx::bulk_load_1d(sb + (QB_ROWS + rr) * row_bytes, qkv + (r0 + rr) * ld_in, row_bytes, &full[st]);
        }
        ptx::bulk_load_1d(sb + 2 * QB_ROWS * row_bytes, rrms_in + r0 * 2, 16, &full[st]);
      }
      __syncwarp();
      if (++st == QB_STAGES) { st = 0; ph ^= 1; }
    }
    return;
  }
  // ---------------- compute warps: warp = (row of the tile, q | k) ----------------
  const int nv = d >> 3;
  const int rr = warp >> 1, which = warp & 1;
  const float* scl = sScale + which * d;
  int st = 0, ob = 0;
  uint32_t ph = 0;
  float S[VPL][8];
#pragma unroll
  for (int i = 0; i < VPL; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) S[i][j] = 0.f;
  for (int64_t t = t0; t < t1; ++t) {
    const int64_t row = t * QB_ROWS + rr;
    const uint32_t* csr = ra.cs_t + (int64_t)rope_pos(ra, row) * ra.rot_half;
    if (tid == 0) ptx::tma_wait_group_read<1>();
    asm volatile("bar.sync 1, %0;\n" ::"n"(QB_WARPS * 32) : "memory");
    ptx::mbar_wait(&full[st], ph);
    const uint8_t* sb = smem + st * stage_bytes;
    const bf16* gsrc = reinterpret_cast<const bf16*>(sb + rr * row_bytes) + which * d;
    const bf16* src = reinterpret_cast<const bf16*>(sb + (QB_ROWS + rr) * row_bytes) + which * d;
    const float rrms = reinterpret_cast<const float*>(sb + 2 * QB_ROWS * row_bytes)[rr * 2 + which];
    bf16* dst = reinterpret_cast<bf16*>(sOut + (ob * QB_ROWS + rr) * row_bytes) + which * d;
    float gn[VPL][8];  // grad wrt the normalised value (rotation transposed, times the learnable scale)
    bf16x8 xp[VPL];
    float part[VPL];
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      part[i] = 0.f;
      const int v = lane + 32 * i;
      if (v < nv) {
        const int c = v * 8, cl = c % ra.hd;
        xp[i] = *reinterpret_cast<const bf16x8*>(src + c);
        float xn[8], scv[8];
        unpack8(xp[i], xn);
        unpack8(*reinterpret_cast<const bf16x8*>(gsrc + c), gn[i]);
        *reinterpret_cast<float4*>(scv) = *reinterpret_cast<const float4*>(scl + c);
        *reinterpret_cast<float4*>(scv + 4) = *reinterpret_cast<const float4*>(scl + c + 4);
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
          const int pj = (cl + j) >> 1;
          if (pj < ra.rot_half) {
            const float2 cs = cs_unpack(__ldg(csr + pj));
            const float ge = gn[i][j], go = gn[i][j + 1];
            gn[i][j] = ge * cs.x + go * cs.y;
            gn[i][j + 1] = go * cs.x - ge * cs.y;
          }
        }
        float q[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float xr = xn[j] * rrms;
          S[i][j] = fmaf(gn[i][j], bf16_round(xr), S[i][j]);  // the reference multiplies the bf16-rounded normalised value
          gn[i][j] *= scv[j];
          q[j] = gn[i][j] * xr;
        }
        part[i] = ((q[0] + q[1]) + (q[2] + q[3])) + ((q[4] + q[5]) + (q[6] + q[7]));
      }
    }
    float dot = part[0];
#pragma unroll
    for (int i = 1; i < VPL; ++i) dot += part[i];
    dot = warp_sum(dot) / d;
    const float rd = rrms * dot;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int v = lane + 32 * i;
      if (v < nv) {
        float xn[8], o[8];
        unpack8(xp[i], xn);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = rrms * (gn[i][j] - xn[j] * rd);
        *reinterpret_cast<bf16x8*>(dst + v * 8) = pack8(o);
      }
    }
    __syncwarp();
    if (lane == 0) ptx::mbar_arrive(&empty[st]);
    ptx::fence_proxy_async_smem();
    asm volatile("bar.sync 1, %0;\n" ::"n"(QB_WARPS * 32) : "memory");
    if (tid == 0) {
#pragma unroll
      for (int r2 = 0; r2 < QB_ROWS; ++r2) ptx::bulk_store_1d(dqkv + (t * QB_ROWS + r2) * ld_out, sOut + (ob * QB_ROWS + r2) * row_bytes, row_bytes);
      ptx::tma_commit_group();
    }
    ob ^= 1;
    if (++st == QB_STAGES) { st = 0; ph ^= 1; }
  }
  // flush the scale gradients: rows 0 / 1 of the tile held by warps (0,1) / (2,3)
  if (dsq != nullptr) {
    for (int round = 0; round < 2; ++round) {
      if (rr == round) {
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
          const int v = lane + 32 * i;
          if (v < nv) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float* a = sAcc + which * d + v * 8 + j;
              if (round == 0) *a = S[i][j];
              else *a += S[i][j];
            }
          }
        }
      }
      asm volatile("bar.sync 1, %0;\n" ::"n"(QB_WARPS * 32) : "memory");
    }
    for (int c = tid; c < 2 * d; c += QB_WARPS * 32) atomicAdd((c < d ? dsq : dsk) + (c < d ? c : c - d), sAcc[c]);
  }
  if (tid == 0) ptx::tma_wait_group<0>();
}

// grid (col chunks, row chunks, 2 = q/k); thread = one 8-channel vector marching down rows in batches of 4
__global__ void __launch_bounds__(256, 3)
qknorm_rope_bwd_cols_kernel(const bf16* __restrict__ dqk, int64_t ld_dqk, const bf16* __restrict__ qkv, int64_t ld_in,
                            RopeArgs ra, const float* __restrict__ rrms_in, float* __restrict__ dsq,
                            float* __restrict__ dsk, int64_t R, int d, int rows_per_block) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= (d >> 3)) return;
  const int c = v * 8;
  const int cl = c % ra.hd;
  const int which = blockIdx.z;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_block;
  const int64_t r1 = min(r0 + rows_per_block, R);
  float S[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) S[j] = 0.f;
  for (int64_t rb = r0; rb < r1; rb += 4) {
    bf16x8 xa[4], ga[4];
    float rr[4];
    int ps[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t row = rb + u;
      if (row < r1) {
        xa[u] = ld8(qkv + row * ld_in + which * d + c);
        ga[u] = ld8(dqk + row * ld_dqk + which * d + c);
        rr[u] = __ldg(rrms_in + row * 2 + which);
        ps[u] = rope_pos(ra, row);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (rb + u < r1) {
        float xq[8], g[8];
        unpack8(xa[u], xq);
        unpack8(ga[u], g);
        const uint32_t* csr = ra.cs_t + (int64_t)ps[u] * ra.rot_half;
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
          const int pj = (cl + j) >> 1;
          if (pj < ra.rot_half) {
            const float2 cs = cs_unpack(__ldg(csr + pj));
            const float ge = g[j], go = g[j + 1];
            g[j] = ge * cs.x + go * cs.y;
            g[j + 1] = go * cs.x - ge * cs.y;
          }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) S[j] += g[j] * bf16_round(xq[j] * rr[u]);  // the reference multiplies the bf16-rounded value
      }
    }
  }
  float* dsc = which ? dsk : dsq;
#pragma unroll
  for (int j = 0; j < 8; ++j) atomicAdd(dsc + c + j, S[j]);
}

int vpl_for(int d) { return (d / 8 + 31) / 32; }
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

static int check_rope(const char* who, int d, int hd, int rot_half, int tokens_per_sample, const int32_t* pos_idx) {
  DLB_REQUIRE(d > 0 && d % 8 == 0 && hd > 0 && hd % 8 == 0 && d % hd == 0, DLB_ERR_SHAPE, "%s: d=%d hd=%d", who, d, hd);
  DLB_REQUIRE(rot_half >= 0 && 2 * rot_half <= hd, DLB_ERR_SHAPE, "%s: rotary dim %d exceeds head dim %d", who, 2 * rot_half, hd);
  DLB_REQUIRE(pos_idx != nullptr || tokens_per_sample > 0, DLB_ERR_SHAPE, "%s: need pos_idx or tokens_per_sample", who);
  return DLB_OK;
}

// qkv: [R, >=2d] packed (q | k | ...) with row stride ld_in; out: [R, 2d] rotated (q | k) with row stride ld_out.
DLB_EXPORT int dlb_qknorm_rope_fwd(const void* qkv, int64_t ld_in, const float* sq, const float* sk, const uint32_t* cs_t,
                                   int rot_half, const int32_t* pos_idx, int pos_offset,
                                   int tokens_per_sample, int hd, void* out, int64_t ld_out, float* rrms, int64_t R,
                                   int d, float eps, cudaStream_t stream) {
  int rc = check_rope("qknorm_rope_fwd", d, hd, rot_half, tokens_per_sample, pos_idx);
  if (rc) return rc;
  DLB_REQUIRE(R > 0 && ld_in % 8 == 0 && ld_out % 8 == 0, DLB_ERR_SHAPE, "qknorm_rope_fwd: bad strides");
  RopeArgs ra{cs_t, pos_idx, rot_half, pos_offset, tokens_per_sample > 0 ? tokens_per_sample : 1, hd};
  // lean warp-per-half-row kernel (rowwise_lean.cuh): needs an exact per-lane unit split and whole units inside the rotary span
  static const bool no_lean = getenv("DLB_NO_LEAN") != nullptr;
  if (!no_lean && R < (1ll << 31) && ((uintptr_t)qkv % 16) == 0 && ((uintptr_t)out % 16) == 0 && ((uintptr_t)cs_t % 16) == 0 && ((uintptr_t)sq % 16) == 0 &&
      ((uintptr_t)sk % 16) == 0) {
    int Uc = 0, UPLc = 0;
    if (lean::pick_units(d, Uc, UPLc) && rot_half % (Uc / 2) == 0 && hd % Uc == 0) {
      const int rpc = lean::rows_per_cta_for((int)R, dlb_num_sms() * 3);
      const int grid_l = (int)((R + rpc - 1) / rpc);
      DLB_LEAN_SWITCH(d, {
        lean::qknorm_rope_fwd_lean<U, UPL><<<grid_l, lean::QK_WARPS * 32, 0, stream>>>((const bf16*)qkv, ld_in, sq, sk, cs_t, pos_idx, rot_half, pos_offset,
                                                                                      ra.tokens_per_sample, hd, (bf16*)out, ld_out, rrms, (int)R, eps, rpc);
        dlb_count_launch();
        return dlb_check_launch("qknorm_rope_fwd_lean");
      });
    }
  }
  // (a cp.async.bulk tile version of this forward measured SLOWER, 0.100 vs 0.088 ms: its packed-bf16 arithmetic is light
  //  enough that 24 resident row-warps per SM already hide the load latency; the backward kernels are the opposite case)
  const int warps = 4;
  int64_t grid64 = (R + warps - 1) / warps;
  const int grid = (int)(grid64 < (int64_t)dlb_num_sms() * 6 ? grid64 : (int64_t)dlb_num_sms() * 6);
  VPL_SWITCH(d, (qknorm_rope_fwd_kernel<VPL><<<grid, warps * 32, 0, stream>>>((const bf16*)qkv, ld_in, sq, sk, ra,
                                                                             (bf16*)out, ld_out, rrms, R, d, eps)));
  dlb_count_launch();
  return dlb_check_launch("qknorm_rope_fwd");
}

DLB_EXPORT int dlb_qknorm_rope_bwd(const void* dqk, int64_t ld_dqk, const void* qkv, int64_t ld_in, const float* sq,
                                   const float* sk, const uint32_t* cs_t, int rot_half,
                                   const int32_t* pos_idx, int pos_offset, int tokens_per_sample, int hd,
                                   const float* rrms, void* dqkv, int64_t ld_out, float* dsq, float* dsk, int64_t R,
                                   int d, cudaStream_t stream) {
  int rc = check_rope("qknorm_rope_bwd", d, hd, rot_half, tokens_per_sample, pos_idx);
  if (rc) return rc;
  DLB_REQUIRE(R > 0 && ld_in % 8 == 0 && ld_out % 8 == 0 && ld_dqk % 8 == 0, DLB_ERR_SHAPE, "qknorm_rope_bwd: bad strides");
  DLB_REQUIRE(rrms != nullptr, DLB_ERR_SHAPE, "qknorm_rope_bwd: the rrms buffer saved by the forward pass is required");
  RopeArgs ra{cs_t, pos_idx, rot_half, pos_offset, tokens_per_sample > 0 ? tokens_per_sample : 1, hd};
  static const bool no_lean = getenv("DLB_NO_LEAN") != nullptr;
  if (!no_lean && R < (1ll << 31) && ((uintptr_t)dqk % 16) == 0 && ((uintptr_t)qkv % 16) == 0 && ((uintptr_t)dqkv % 16) == 0 && ((uintptr_t)cs_t % 16) == 0 &&
      ((uintptr_t)sq % 16) == 0 && ((uintptr_t)sk % 16) == 0 && ((uintptr_t)dsq % 16) == 0 && ((uintptr_t)dsk % 16) == 0 && (dsq == nullptr) == (dsk == nullptr)) {
    int Uc = 0, UPLc = 0;
    const size_t smem_l = (size_t)2 * d * 4 + (size_t)lean::WARPS * lean::QB_NS * 2 * d * 2;
    if (lean::pick_units(d, Uc, UPLc) && rot_half % (Uc / 2) == 0 && hd % Uc == 0 && smem_l <= 220 * 1024) {
      const int rpc = lean::rows_per_cta_for((int)R, dlb_num_sms());
      const int grid_l = (int)((R + rpc - 1) / rpc);
      DLB_LEAN_SWITCH(d, {
        cudaFuncSetAttribute(lean::qknorm_rope_bwd_lean<U, UPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_l);
        lean::qknorm_rope_bwd_lean<U, UPL><<<grid_l, lean::WARPS * 32, smem_l, stream>>>(
            (const bf16*)dqk, ld_dqk, (const bf16*)qkv, ld_in, sq, sk, cs_t, pos_idx, rot_half, pos_offset, ra.tokens_per_sample, hd, rrms, (bf16*)dqkv, ld_out, dsq,
            dsk, (int)R, rpc);
        dlb_count_launch();
        return dlb_check_launch("qknorm_rope_bwd_lean");
      });
    }
  }
  {  // one-pass tiled kernel (raw gradient + scale gradients from a single read)
    const size_t smem_t = (size_t)QB_STAGES * (2 * QB_ROWS * 4 * d + 16) + 2 * (size_t)QB_ROWS * 4 * d + 4 * (size_t)d * 4;
    const bool aligned = ((uintptr_t)dqk % 16) == 0 && ((uintptr_t)qkv % 16) == 0 && ((uintptr_t)dqkv % 16) == 0 && ((uintptr_t)rrms % 16) == 0;
    if (R % QB_ROWS == 0 && aligned && smem_t <= 113 * 1024 && (dsq == nullptr) == (dsk == nullptr)) {
      const int64_t ntiles = R / QB_ROWS;
      const int max_ctas = dlb_num_sms() * 2;
      const int tiles_per_cta = (int)((ntiles + max_ctas - 1) / max_ctas);
      const int grid_t = (int)((ntiles + tiles_per_cta - 1) / tiles_per_cta);
      VPL_SWITCH(d, {
        cudaFuncSetAttribute(qknorm_rope_bwd_tile_kernel<VPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024);
        qknorm_rope_bwd_tile_kernel<VPL><<<grid_t, (QB_WARPS + 1) * 32, smem_t, stream>>>(
            (const bf16*)dqk, ld_dqk, (const bf16*)qkv, ld_in, sq, sk, ra, rrms, (bf16*)dqkv, ld_out, dsq, dsk, R, d, tiles_per_cta);
      });
      dlb_count_launch();
      return dlb_check_launch("qknorm_rope_bwd_tile");
    }
  }
  const int warps = 4;
  int64_t grid64 = (R + warps - 1) / warps;
  const int grid = (int)(grid64 < (int64_t)dlb_num_sms() * 5 ? grid64 : (int64_t)dlb_num_sms() * 5);
  VPL_SWITCH(d, (qknorm_rope_bwd_rows_kernel<VPL><<<grid, warps * 32, 0, stream>>>(
                    (const bf16*)dqk, ld_dqk, (const bf16*)qkv, ld_in, sq, sk, ra, rrms, (bf16*)dqkv, ld_out, R, d)));
  dlb_count_launch();
  if (dsq != nullptr && dsk != nullptr) {
    const int nv = d / 8;
    const int threads = nv < 256 ? (nv + 31) / 32 * 32 : 256;
    const int col_chunks = (nv + threads - 1) / threads;
    int64_t want_blocks = (int64_t)dlb_num_sms() * 6 / (col_chunks * 2);
    if (want_blocks < 1) want_blocks = 1;
    int64_t rpb = (R + want_blocks - 1) / want_blocks;
    rpb = (rpb + 3) / 4 * 4;
    if (rpb < 16) rpb = 16;
    dim3 g2(col_chunks, (unsigned)((R + rpb - 1) / rpb), 2);
    qknorm_rope_bwd_cols_kernel<<<g2, threads, 0, stream>>>((const bf16*)dqk, ld_dqk, (const bf16*)qkv, ld_in, ra, rrms, dsq,
                                                          dsk, R, d, (int)rpb);
    dlb_count_launch();
  }
  return dlb_check_launch("qknorm_rope_bwd");
}

// cos/sin table of get_cos_sin_ndim_grid (nn.py:262-307): fp64 angles, stored fp32. pos: [P, n_axes] int32.
__global__ void rope_table_kernel(const int32_t* __restrict__ pos, int n_axes, const int32_t* __restrict__ axis_of_pair,
                                  const int32_t* __restrict__ local_of_pair, const int32_t* __restrict__ axis_dim,
                                  double base, float* __restrict__ cos_t, float* __restrict__ sin_t,
                                  uint32_t* __restrict__ cs_t, int64_t P, int rot_half) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P * rot_half) return;
  const int64_t p = i / rot_half;
  const int j = (int)(i - p * rot_half);
  const int ax = axis_of_pair[j];
  const double freq = 1.0 / pow(base, (double)(2 * local_of_pair[j]) / (double)axis_dim[ax]);
  const double ang = (double)pos[p * n_axes + ax] * freq;
  const float c = (float)cos(ang), sn = (float)sin(ang);
  if (cos_t) { cos_t[i] = c; sin_t[i] = sn; }
  if (cs_t) cs_t[i] = pack_bf16x2(c, sn);
}

DLB_EXPORT int dlb_rope_table(const int32_t* pos, int n_axes, const int32_t* axis_of_pair, const int32_t* local_of_pair,
                              const int32_t* axis_dim, double base, float* cos_t, float* sin_t, uint32_t* cs_t, int64_t P,
                              int rot_half, cudaStream_t stream) {
  DLB_REQUIRE(P > 0 && rot_half > 0 && n_axes > 0, DLB_ERR_SHAPE, "rope_table: bad shape");
  const int64_t total = P * rot_half;
  rope_table_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(pos, n_axes, axis_of_pair, local_of_pair,
                                                                        axis_dim, base, cos_t, sin_t, cs_t, P, rot_half);
  dlb_count_launch();
  return dlb_check_launch("rope_table");
}
