"""Attention micro-benchmark at the DiT-XL/2 geometry (B=128, H=16, N=256, hd=72) and the SPRINT/MMDiT one
(B=64, H=12, L=128 text + 256 image, hd=64): tcgen05 forward vs the mma.sync forward, plus the backward."""
import json
import sys

import torch

sys.path.insert(0, ".")
from diffulab_b200 import ops  # noqa: E402


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


CASES = [("dit_xl2", 128, 16, 72, 0, 256), ("sprint_mm", 64, 12, 64, 128, 256), ("hd128", 32, 8, 128, 0, 1024)]
ONLY = sys.argv[1:]  # optional case names
for name, B, H, hd, L, N in [c for c in CASES if not ONLY or c[0] in ONLY]:
    d = H * hd
    lens = [L, N] if L else [N]
    qks = [torch.randn(B * l, 2 * d, device="cuda").bfloat16() for l in lens]
    qkvs = [torch.randn(B * l, 3 * d, device="cuda").bfloat16() for l in lens]
    specs = [ops.AttnSegSpec(a, b, l) for a, b, l in zip(qks, qkvs, lens)]
    S = L + N
    flops = 4.0 * B * H * S * S * hd
    row = {"case": name, "B": B, "H": H, "hd": hd, "S": S, "fwd_gflop": flops / 1e9}
    for impl in ("dlb_attn_fwd_tc", "dlb_attn_fwd"):
        ms = timeit(lambda: ops.attn_fwd(specs, B, H, hd, hd ** -0.5, None, impl=impl))
        row[impl + "_ms"] = round(ms, 4)
        row[impl + "_tflops"] = round(flops / ms / 1e9, 1)
    outs, lse = ops.attn_fwd(specs, B, H, hd, hd ** -0.5, None)
    douts = [torch.randn_like(o) for o in outs]
    dqkvs = [torch.empty_like(q) for q in qkvs]
    for impl in ("dlb_attn_bwd_tc", "dlb_attn_bwd"):
        ms = timeit(lambda: ops.attn_bwd(specs, outs, douts, lse, B, H, hd, hd ** -0.5, dqkvs, None, impl=impl))
        row[impl + "_ms"] = round(ms, 4)
        row[impl + "_tflops"] = round(2.5 * flops / ms / 1e9, 1)
    print(json.dumps(row), flush=True)
