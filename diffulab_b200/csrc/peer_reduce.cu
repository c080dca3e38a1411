// Data-parallel gradient reduction over NVLink peer memory, without NCCL kernels on the SMs the GEMMs need.
//
// The reference trains under HF Accelerate's DDP (training/trainers/common.py:103-109): fp32 gradient buckets all-reduced (mean)
// behind backward by NCCL kernels that take 16-32 SMs while they run. Here the flat gradient buffer lives in symmetric (peer
// mapped) memory and a bucket [off, off + n) is reduced in pieces, rank p owning piece p:
//
//   copy-engine mode  : rank p PULLS piece p of every peer's bucket with plain device-to-device copies (copy engines over NVLink,
//                       zero SMs), dlb_reduce_pieces averages them into its own piece (a short HBM-bound kernel), then every
//                       rank pulls the reduced pieces back (copy engines again).
//   multicast mode    : dlb_multimem_allreduce — multimem.ld_reduce sums piece p over all ranks INSIDE the NVSwitch, the mean is
//                       written to all ranks with multimem.st; a handful of CTAs (the wire needs ~40 GB/s), one kernel per bucket.
//
// Cross-rank ordering (peer gradients complete before the pulls, reduced pieces complete before they are read) is provided by the
// caller with stream-ordered symmetric-memory barriers (diffulab_b200/training.py GradReducer).
#include "common.cuh"

namespace {

// own[i] = scale * sum over ranks q = 0..world-1 (fixed order: run-to-run deterministic) of piece_q[i];
// piece_rank = own, piece_q (q != rank) = staged + slot(q) * stage_stride with slot(q) = (q - rank - 1 + world) % world.
__global__ void __launch_bounds__(256)
reduce_pieces_kernel(float* __restrict__ own, const float* __restrict__ staged, int64_t stage_stride, int world, int rank, int64_t n4, float scale) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int q = 0; q < world; ++q) {
      const int slot = q > rank ? q - rank - 1 : q - rank - 1 + world;  // (q - rank - 1) mod world
      const float4* src = q == rank ? reinterpret_cast<const float4*>(own) : reinterpret_cast<const float4*>(staged + (int64_t)slot * stage_stride);
      const float4 v = src[i];
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    reinterpret_cast<float4*>(own)[i] = make_float4(acc.x * scale, acc.y * scale, acc.z * scale, acc.w * scale);
  }
}

// mc: MULTICAST address of this rank's piece (the same offset in every rank's buffer). One 16-byte multimem.ld_reduce returns the
// sum over all ranks (reduced in the switch), multimem.st writes the scaled value to every rank.
__global__ void __launch_bounds__(512)
multimem_allreduce_kernel(float* __restrict__ mc, int64_t n4, float scale) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  constexpr int UNROLL = 4;  // independent 16-byte reductions in flight per thread (NVLink round trip ~2-3 us)
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + (UNROLL - 1) * stride < n4; i += UNROLL * stride) {
    float4 v[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u)
      asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                   : "=f"(v[u].x), "=f"(v[u].y), "=f"(v[u].z), "=f"(v[u].w) : "l"(mc + 4 * (i + u * stride)) : "memory");
#pragma unroll
    for (int u = 0; u < UNROLL; ++u)
      asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc + 4 * (i + u * stride)), "f"(v[u].x * scale),
                   "f"(v[u].y * scale), "f"(v[u].z * scale), "f"(v[u].w * scale) : "memory");
  }
  for (; i < n4; i += stride) {
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(mc + 4 * i) : "memory");
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc + 4 * i), "f"(v.x * scale), "f"(v.y * scale), "f"(v.z * scale),
                 "f"(v.w * scale) : "memory");
  }
}

}  // namespace

DLB_EXPORT int dlb_reduce_pieces(float* own, const float* staged, int64_t stage_stride, int world, int rank, int64_t n, float scale, int max_ctas,
                                 cudaStream_t stream) {
  DLB_REQUIRE(world >= 1 && rank >= 0 && rank < world && n >= 0, DLB_ERR_SHAPE, "reduce_pieces: bad world / rank / n");
  DLB_REQUIRE(n % 4 == 0 && stage_stride % 4 == 0 && ((uintptr_t)own & 15) == 0 && ((uintptr_t)staged & 15) == 0, DLB_ERR_ALIGN,
              "reduce_pieces: pieces must be 16-byte aligned multiples of 4 floats");
  if (n == 0) return 0;
  int64_t grid = (n / 4 + 255) / 256;
  const int cap = max_ctas > 0 ? max_ctas : 4 * dlb_num_sms();
  if (grid > cap) grid = cap;
  reduce_pieces_kernel<<<(unsigned)grid, 256, 0, stream>>>(own, staged, stage_stride, world, rank, n / 4, scale);
  dlb_count_launch();
  return dlb_check_launch("reduce_pieces");
}

DLB_EXPORT int dlb_multimem_allreduce(void* mc_piece, int64_t n, float scale, int ctas, cudaStream_t stream) {
  DLB_REQUIRE(n >= 0 && n % 4 == 0 && ((uintptr_t)mc_piece & 15) == 0, DLB_ERR_ALIGN, "multimem_allreduce: piece must be a 16-byte aligned multiple of 4 floats");
  DLB_REQUIRE(ctas > 0 && ctas <= 64, DLB_ERR_SHAPE, "multimem_allreduce: ctas must be in 1..64");
  if (n == 0) return 0;
  multimem_allreduce_kernel<<<(unsigned)ctas, 512, 0, stream>>>(reinterpret_cast<float*>(mc_piece), n / 4, scale);
  dlb_count_launch();
  return dlb_check_launch("multimem_allreduce");
}
