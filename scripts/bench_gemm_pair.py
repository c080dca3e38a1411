"""CTA-pair (cta_group::2) GEMM vs the single-CTA kernel vs cuBLAS on the DiT-XL/2 shapes (same protocol as bench_gemm.py)."""
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
for name, N, K in [("qkv", 3456, 1152), ("proj", 1152, 1152), ("mlp_up", 9216, 1152), ("mlp_down", 1152, 4608)]:
    x = torch.randn(M, K, device="cuda").bfloat16()
    w = (torch.randn(N, K, device="cuda") * 0.02).bfloat16()
    dy = torch.randn(M, N, device="cuda").bfloat16()
    dw = torch.zeros(N, K, device="cuda")
    flops = 2.0 * M * N * K
    tf = lambda ms: round(flops / ms / 1e9)  # noqa: E731
    row = {"shape": name, "N": N, "K": K}
    row["fwd"] = {"cublas": tf(timeit(lambda: torch.matmul(x, w.t()))), "single_auto": tf(timeit(lambda: ops.gemm(x, w))),
                  "pair128": tf(timeit(lambda: ops.gemm(x, w, tile_n=128, pair=True))), "pair256": tf(timeit(lambda: ops.gemm(x, w, tile_n=256, pair=True)))}
    row["dgrad"] = {"cublas": tf(timeit(lambda: torch.matmul(dy, w))), "single_auto": tf(timeit(lambda: ops.gemm(dy, w, b_mn=True))),
                    "pair128": tf(timeit(lambda: ops.gemm(dy, w, b_mn=True, tile_n=128, pair=True))),
                    "pair256": tf(timeit(lambda: ops.gemm(dy, w, b_mn=True, tile_n=256, pair=True)))}
    wg = {"cublas": tf(timeit(lambda: torch.matmul(dy.t(), x))),
          "single_auto": tf(timeit(lambda: ops.gemm(dy, x, a_mn=True, b_mn=True, out=dw, accumulate=True)))}
    for tn in (128, 256):
        for sk in (1, 2, 4, 8):
            wg[f"pair{tn}_sk{sk}"] = tf(timeit(lambda: ops.gemm(dy, x, a_mn=True, b_mn=True, out=dw, accumulate=True, split_k=sk, tile_n=tn, pair=True)))
    row["wgrad"] = wg
    print(json.dumps(row), flush=True)
