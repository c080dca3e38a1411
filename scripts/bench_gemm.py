"""Micro-benchmark of the tcgen05 GEMM on the DiT-XL/2 shapes (per-GPU batch 128 -> M = 32768), all three
operand-major combinations, each tile_n, next to torch.matmul (cuBLAS) on the same inputs.
Timing: CUDA events, 3 warm-up + 10 timed launches, L2 flushed between launches (256 MB write)."""
import json
import sys

import torch

sys.path.insert(0, ".")
from diffulab_b200 import ops  # noqa: E402

flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


M = 32768
res = []
for name, N, K in [("qkv", 3456, 1152), ("proj", 1152, 1152), ("mlp_up", 9216, 1152), ("mlp_down", 1152, 4608)]:
    x = torch.randn(M, K, device="cuda").bfloat16()
    w = (torch.randn(N, K, device="cuda") * 0.02).bfloat16()
    dy = torch.randn(M, N, device="cuda").bfloat16()
    flops = 2.0 * M * N * K
    row = {"shape": name, "M": M, "N": N, "K": K}
    row["cublas_fwd_tflops"] = flops / timeit(lambda: torch.matmul(x, w.t())) / 1e9
    for tn in (128, 192, 256):
        row[f"fwd_tn{tn}_tflops"] = flops / timeit(lambda: ops.gemm(x, w, tile_n=tn)) / 1e9
    # dgrad: dX[M,K] = dY[M,N] @ W[N,K]  -> B operand MN-major
    row["cublas_dgrad_tflops"] = flops / timeit(lambda: torch.matmul(dy, w)) / 1e9
    for tn in (128, 192, 256):
        row[f"dgrad_tn{tn}_tflops"] = flops / timeit(lambda: ops.gemm(dy, w, b_mn=True, tile_n=tn)) / 1e9
    # wgrad: dW[N,K] = dY^T @ X -> both MN-major, fp32 accumulate with split-K
    dw = torch.zeros(N, K, device="cuda")
    row["cublas_wgrad_tflops"] = flops / timeit(lambda: torch.matmul(dy.t(), x)) / 1e9
    for tn in (128, 192, 256):
        for sk in (1, 2, 4, 8):
            row[f"wgrad_tn{tn}_sk{sk}_tflops"] = flops / timeit(lambda: ops.gemm(dy, x, a_mn=True, b_mn=True, out=dw, accumulate=True, split_k=sk, tile_n=tn)) / 1e9
    res.append(row)
    print(json.dumps(row), flush=True)
