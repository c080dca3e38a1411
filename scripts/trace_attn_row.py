"""SM-clock timeline of items 2 and 3 of one CTA of the whole-row attention forward kernel (attn_fwd_row_tc_kernel, ROW_TRACE_*).
compute thread 0: slots 0 start, 1 S ready, 2 scores in registers, 3 max exchanged, 4 P written + arrived, 5 before the deferred
epilogue, 6 O(prev) ready, 7 staged, 8 stored, 9 end; MMA warp (+32): 1 S issued, 2 ps_full seen, 3 P V issued."""
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
    ops.attn_fwd(specs, B, H, hd, hd ** -0.5, None)
torch.cuda.synchronize()
lib.dlb_attn_set_trace(trace.data_ptr())
ops.attn_fwd(specs, B, H, hd, hd ** -0.5, None)
torch.cuda.synchronize()
lib.dlb_attn_set_trace(None)
t = trace.cpu().tolist()
pts = sorted([(v, i) for i, v in enumerate(t[:64]) if v])
t0 = pts[0][0]
for v, i in pts:
    who, j = ("mma", i - 32) if i >= 32 else ("cmp", i)
    print(f"{who} item {2 + j // 10} slot {j % 10}  @{v - t0:7d}")
