/* Development probes of libdiffulab_b200_probes.so (diffulab_b200/csrc/probes/): hardware-layout checks used by tests/ and
 * scripts/ only. NOT part of the product ABI (include/diffulab_b200.h) and never loaded by the diffulab_b200 package's hot path. */
#ifndef DIFFULAB_B200_PROBES_H
#define DIFFULAB_B200_PROBES_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct CUstream_st* dlb_probe_stream_t;

/* D[128,N] = A*B^T from thread-staged non-swizzled UMMA operands (pins the LBO/SBO descriptor semantics) */
int dlb_umma_probe(const void* A, const void* B, float* D, int N, int K, int a_mn, int b_mn, int swap_lbo_sbo, dlb_probe_stream_t stream);
/* stream [64 rows x hd] head slices with the round-1 4-D tensor map (16-byte inner boxes); dump (nullable) = CTA 0's first tile */
int dlb_tma_gather_probe(const void* base, int64_t rows, int64_t ld, int H, int hd, int grid, int tiles_per_cta, void* dump,
                         dlb_probe_stream_t stream);
/* swizzled head-slice tiles (csrc/attn_sw.cuh): D1[128,64] = X Y^T (K-major x K-major), D2[128,HDP] = P Y (Y read MN-major) */
int dlb_attn_sw_probe(const void* X, const void* Y, const void* P, float* D1, float* D2, int64_t rows, int64_t ld, int H, int hd, int h,
                      int xrow, int yrow, void* dump, dlb_probe_stream_t stream);
/* stream 64-row swizzled head-slice tiles through a 4-stage ring on `grid` CTAs (TMA delivery rate of the new layout) */
int dlb_attn_sw_stream_probe(const void* X, int64_t rows, int64_t ld, int H, int hd, int grid, int tiles_per_cta, dlb_probe_stream_t stream);
#ifdef __cplusplus
}
#endif
#endif
