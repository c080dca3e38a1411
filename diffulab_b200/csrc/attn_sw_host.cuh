// Host side of the swizzled head-slice tiles (attn_sw.cuh): cached 3-D tensor maps {hd, heads, tokens} of a packed bf16
// activation [tokens, ld]; main = box {64, 1, 64} SWIZZLE_128B, tail = box {16, 1, 64} SWIZZLE_32B.
#pragma once
#include "common.cuh"
#include <cudaTypedefs.h>
#include <mutex>
#include <unordered_map>

namespace attn_sw_host {
struct Key {
  const void* ptr; int64_t rows, ld; int H, hd, tail;
  bool operator==(const Key& o) const { return ptr == o.ptr && rows == o.rows && ld == o.ld && H == o.H && hd == o.hd && tail == o.tail; }
};
struct KeyHash {
  size_t operator()(const Key& k) const {
    uint64_t h = reinterpret_cast<uintptr_t>(k.ptr) * 0x9E3779B97F4A7C15ull;
    for (uint64_t v : {(uint64_t)k.rows, (uint64_t)k.ld, (uint64_t)k.H << 32 | (uint64_t)k.hd, (uint64_t)k.tail}) h = (h ^ v) * 0x100000001B3ull + (h >> 29);
    return (size_t)h;
  }
};
// tail = 0: main map (64-element boxes, 128-byte swizzle); tail = 1: 16-element boxes, 32-byte swizzle
inline int head_map3(CUtensorMap* out, const void* base, int64_t rows, int64_t ld, int H, int hd, int tail) {
  static std::unordered_map<Key, CUtensorMap, KeyHash> cache;
  static std::mutex mu;
  static PFN_cuTensorMapEncodeTiled_v12000 enc = nullptr;
  const Key key{base, rows, ld, H, hd, tail};
  {
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find(key);
    if (it != cache.end()) { *out = it->second; return DLB_OK; }
  }
  if (!enc) {
    void* fp = nullptr;
    cudaDriverEntryPointQueryResult q;
    const bool ok = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess;
    DLB_REQUIRE(ok, DLB_ERR_DRIVER, "attention: cuTensorMapEncodeTiled unavailable");
    enc = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fp);
  }
  cuuint64_t dims[3] = {(cuuint64_t)hd, (cuuint64_t)H, (cuuint64_t)rows};
  cuuint64_t strides[2] = {(cuuint64_t)hd * 2, (cuuint64_t)ld * 2};
  cuuint32_t box[3] = {tail ? 16u : 64u, 1u, 64u};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   tail ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DLB_REQUIRE(r == CUDA_SUCCESS, DLB_ERR_DRIVER, "attention: cuTensorMapEncodeTiled(3d head map) failed (%d): rows=%lld ld=%lld H=%d hd=%d tail=%d", (int)r,
              (long long)rows, (long long)ld, H, hd, tail);
  std::lock_guard<std::mutex> lock(mu);
  if (cache.size() > 8192) cache.clear();
  cache.emplace(key, *out);
  return DLB_OK;
}
}  // namespace attn_sw_host
