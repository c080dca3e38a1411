"""Attention kernels timed the way they run inside the train step: every timed call is preceded by the block's QKV-sized GEMM and
QK-norm pass on fresh data (power-capped clocks, producer-written inputs in L2, cold instruction cache), CUDA events around the
attention call only. Compare with scripts/bench_attn.py (back-to-back launches of the same kernel at boost clocks)."""
import json
import sys

import torch

sys.path.insert(0, ".")
from diffulab_b200 import ops  # noqa: E402

B, H, hd, N = 128, 16, 72, 256
d = H * hd
R = B * N
x = torch.randn(R, d, device="cuda").bfloat16()
w = (torch.randn(3 * d, d, device="cuda") * 0.03).bfloat16()
wu = (torch.randn(8 * d, d, device="cuda") * 0.03).bfloat16()
pos = torch.stack([torch.arange(N) // 16, torch.arange(N) % 16], -1).int().cuda()
rope = ops.rope_table(pos, [36, 36], 10000.0)
sq = torch.ones(d, device="cuda")
fw, bw = [], []
for it in range(12):
    u = ops.gemm(x, wu)  # MLP-up sized GEMM first: keeps the chip at its power-capped clock
    qkv = ops.gemm(x, w)
    qk, rrms = ops.qknorm_rope_fwd(qkv, sq, sq, rope, hd, tokens_per_sample=N)
    specs = [ops.AttnSegSpec(qk, qkv, N)]
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    outs, lse = ops.attn_fwd(specs, B, H, hd, hd ** -0.5, None)
    e1.record()
    dout = ops.gemm(outs[0], w[:d, :d].contiguous())  # a GEMM between forward and backward as in the step
    dqkv = [torch.empty_like(qkv)]
    e3, e4 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e3.record()
    ops.attn_bwd(specs, outs, [dout], lse, B, H, hd, hd ** -0.5, dqkv, None)
    e4.record()
    torch.cuda.synchronize()
    if it >= 2:
        fw.append(e0.elapsed_time(e1))
        bw.append(e3.elapsed_time(e4))
fw.sort()
bw.sort()
flops = 4.0 * B * H * N * N * hd
print(json.dumps({"case": "dit_xl2 in-step conditions", "fwd_ms_median": round(fw[len(fw) // 2], 4), "bwd_ms_median": round(bw[len(bw) // 2], 4),
                  "fwd_tflops": round(flops / fw[len(fw) // 2] / 1e9, 1), "bwd_tflops": round(2.5 * flops / bw[len(bw) // 2] / 1e9, 1),
                  "fwd_all": [round(v, 3) for v in fw], "bwd_all": [round(v, 3) for v in bw]}))
