"""SM-clock timeline of one CTA of each tcgen05 attention kernel (dlb_attn_set_trace) at the DiT-XL/2 geometry.
Prints, per kernel, the cycle deltas between the instrumented points (see ATTN_TRACE in csrc/attention_tc.cu)."""
import sys

import torch

sys.path.insert(0, ".")
from diffulab_b200 import _lib, ops  # noqa: E402

B, H, hd, N = 128, 16, 72, 256
d = H * hd
qk = torch.randn(B * N, 2 * d, device="cuda").bfloat16()
qkv = torch.randn(B * N, 3 * d, device="cuda").bfloat16()
specs = [ops.AttnSegSpec(qk, qkv, N)]
lib = _lib.load()
trace = torch.zeros(3 * 64, dtype=torch.int64, device="cuda")
for _ in range(3):
    outs, lse = ops.attn_fwd(specs, B, H, hd, hd ** -0.5, None)
douts = [torch.randn_like(o) for o in outs]
dqkv = [torch.empty_like(qkv)]
ops.attn_bwd(specs, outs, douts, lse, B, H, hd, hd ** -0.5, dqkv, None)
torch.cuda.synchronize()
lib.dlb_attn_set_trace(trace.data_ptr())
outs, lse = ops.attn_fwd(specs, B, H, hd, hd ** -0.5, None)
ops.attn_bwd(specs, outs, douts, lse, B, H, hd, hd ** -0.5, dqkv, None)
torch.cuda.synchronize()
lib.dlb_attn_set_trace(None)
t = trace.cpu().tolist()
for k, name in enumerate(("fwd", "dq", "dkv")):
    pts = sorted([(v, i) for i, v in enumerate(t[k * 64:(k + 1) * 64]) if v])
    t0 = pts[0][0]
    print(name, "total cycles", pts[-1][0] - t0, "(slots 32-55: MMA warp; others: compute thread 0)")
    prev = t0
    for v, i in pts:
        who = "mma " if 32 <= i < 56 else "cmp "
        print(f"   {who} slot {i:2d}  +{v - prev:6d}   @{v - t0:7d}")
        prev = v
