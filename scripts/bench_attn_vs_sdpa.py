"""Same-box table: the tcgen05 attention kernels against torch.nn.functional.scaled_dot_product_attention (flash and cuDNN
backends, bf16) at the DiT-XL/2 geometry (B 128, H 16, hd 72, N 256) and at hd 64 / hd 128. SDPA gets what the reference hands it
(`[B, H, N, hd]` views of the packed projection, reference mmdit.py:92-98); forward, and forward + backward."""
import json
import sys

import torch
import torch.nn.functional as F
from torch.nn.attention import SDPBackend, sdpa_kernel

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


for name, B, H, hd, N in [("dit_xl2", 128, 16, 72, 256), ("hd64", 64, 12, 64, 256), ("hd128", 32, 8, 128, 1024)]:
    d = H * hd
    qk = torch.randn(B * N, 2 * d, device="cuda").bfloat16()
    qkv = torch.randn(B * N, 3 * d, device="cuda").bfloat16()
    specs = [ops.AttnSegSpec(qk, qkv, N)]
    flops = 4.0 * B * H * N * N * hd
    row = {"case": name, "B": B, "H": H, "hd": hd, "N": N}
    row["ours_fwd_ms"] = round(timeit(lambda: ops.attn_fwd(specs, B, H, hd, hd ** -0.5, None)), 4)
    outs, lse = ops.attn_fwd(specs, B, H, hd, hd ** -0.5, None)
    douts = [torch.randn_like(o) for o in outs]
    dqkv = [torch.empty_like(qkv)]
    row["ours_bwd_ms"] = round(timeit(lambda: ops.attn_bwd(specs, outs, douts, lse, B, H, hd, hd ** -0.5, dqkv, None)), 4)
    q = qk[:, :d].view(B, N, H, hd).transpose(1, 2)
    k = qk[:, d:].view(B, N, H, hd).transpose(1, 2)
    v = qkv[:, 2 * d:].view(B, N, H, hd).transpose(1, 2)
    for tag, backend in (("flash", SDPBackend.FLASH_ATTENTION), ("cudnn", SDPBackend.CUDNN_ATTENTION), ("efficient", SDPBackend.EFFICIENT_ATTENTION)):
        try:
            with sdpa_kernel(backend):
                row[f"sdpa_{tag}_fwd_ms"] = round(timeit(lambda: F.scaled_dot_product_attention(q, k, v)), 4)
                qg, kg, vg = (t.detach().clone().requires_grad_(True) for t in (q, k, v))
                o = F.scaled_dot_product_attention(qg, kg, vg)
                do = torch.randn_like(o)
                row[f"sdpa_{tag}_bwd_ms"] = round(timeit(lambda: torch.autograd.grad(o, (qg, kg, vg), do, retain_graph=True)), 4)
        except Exception as e:  # noqa: BLE001 - a backend may refuse the shape (head dim 72)
            row[f"sdpa_{tag}"] = f"unavailable: {str(e)[:80]}"
    row["fwd_tflops_ours"] = round(flops / row["ours_fwd_ms"] / 1e9, 1)
    row["bwd_tflops_ours"] = round(2.5 * flops / row["ours_bwd_ms"] / 1e9, 1)
    print(json.dumps(row), flush=True)
