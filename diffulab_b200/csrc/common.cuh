// Shared host/device helpers for libdiffulab_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#define DLB_EXPORT extern "C" __attribute__((visibility("default")))

// Error codes returned across the C ABI (see include/diffulab_b200.h).
enum {
  DLB_OK = 0,
  DLB_ERR_SHAPE = -1,
  DLB_ERR_ALIGN = -2,
  DLB_ERR_UNSUPPORTED = -3,
  DLB_ERR_DRIVER = -4,
};

void dlb_set_error(const char* fmt, ...);
int dlb_check_launch(const char* what);   // returns 0 or positive cudaError_t
void dlb_count_launch(int n = 1);
int dlb_num_sms();

#define DLB_REQUIRE(cond, code, ...)      \
  do {                                    \
    if (!(cond)) {                        \
      dlb_set_error(__VA_ARGS__);         \
      return (code);                      \
    }                                     \
  } while (0)

#ifdef __CUDACC__

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Block-wide sum of one float per thread; `red` needs >= 32 floats of shared memory.
// All threads receive the total. Safe to call repeatedly (leading barrier protects reuse).
__device__ __forceinline__ float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  float t = (lane < nw) ? red[lane] : 0.f;
  t = warp_sum(t);
  return t;
}

__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 p = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&p);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
  __nv_bfloat162 p = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(p);
}

// 8 bf16 values <-> 8 floats through one 16-byte vector.
struct alignas(16) bf16x8 { uint32_t u[4]; };
// Packed fp32 pairs (sm_100: FADD2 / FMUL2 / FFMA2 issue ONE instruction for two IEEE-rn operations, same results as the scalar
// forms). Used where a kernel is bound by issue slots rather than by the FP32 pipe (GEMM epilogue math, row kernels).
struct f32x2 { unsigned long long v; };
__device__ __forceinline__ f32x2 make_f32x2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void split_f32x2(f32x2 a, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a.v)); }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
  return r;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
  return r;
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v));
  return r;
}
// bf16x2 word -> (lo, hi) as an fp32 pair
__device__ __forceinline__ f32x2 unpack2(uint32_t u) { return make_f32x2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u)); }
// fp32 pair -> bf16x2 word (round to nearest even), one CVT
__device__ __forceinline__ uint32_t pack2(f32x2 a) {
  float lo, hi;
  split_f32x2(a, lo, hi);
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

__device__ __forceinline__ void unpack8(const bf16x8& p, float* f) {
#pragma unroll
  for (int i = 0; i < 4; ++i) { float2 t = unpack_bf16x2(p.u[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
}
__device__ __forceinline__ bf16x8 pack8(const float* f) {
  bf16x8 p;
#pragma unroll
  for (int i = 0; i < 4; ++i) p.u[i] = pack_bf16x2(f[2 * i], f[2 * i + 1]);
  return p;
}
__device__ __forceinline__ bf16x8 ld8(const __nv_bfloat16* p) { return *reinterpret_cast<const bf16x8*>(p); }
__device__ __forceinline__ void st8(__nv_bfloat16* p, const bf16x8& v) { *reinterpret_cast<bf16x8*>(p) = v; }

__device__ __forceinline__ float silu_f(float x) { return x / (1.f + __expf(-x)); }
// silu(x) = x * sigmoid(x) = h * tanh(h) + h with h = x / 2: one MUFU op and no division (the GEMM-epilogue version; the
// result is rounded to bf16 by every caller, tanh.approx's 2^-11 relative error sits below that rounding)
__device__ __forceinline__ float silu_fast(float x) {
  const float h = 0.5f * x;
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  return fmaf(h, t, h);
}
// d/dx silu(x) = s + x s (1 - s), s = sigmoid(x)
__device__ __forceinline__ float dsilu_f(float x) {
  float s = 1.f / (1.f + __expf(-x));
  return s * (1.f + x * (1.f - s));
}

#endif  // __CUDACC__
