"""CTA timelines (dlb_attn_set_trace) of the backward attention kernels, round-1 library vs current, same box, DiT-XL/2 geometry."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, ".")
from diffulab_b200 import _lib, ops  # noqa: E402

new = _lib.load()
old = C.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "ab", "libdiffulab_b200_r1.so"))
p, i32, f32 = C.c_void_p, C.c_int, C.c_float
old.dlb_attn_fwd_tc.argtypes = [p, i32, p, p, i32, i32, i32, i32, f32, p]
old.dlb_attn_bwd_tc.argtypes = [p, i32, p, p, p, i32, i32, i32, i32, f32, p]
old.dlb_attn_set_trace.argtypes = [p]
st = torch.cuda.current_stream().cuda_stream
B, H, hd, N = 128, 16, int(sys.argv[1]) if len(sys.argv) > 1 else 72, 256
d = H * hd
qk = torch.randn(B * N, 2 * d, device="cuda").bfloat16()
qkv = torch.randn(B * N, 3 * d, device="cuda").bfloat16()
specs = [ops.AttnSegSpec(qk, qkv, N)]
outs = [torch.empty(B * N, d, device="cuda", dtype=torch.bfloat16)]
lse = torch.empty(B, H, N, device="cuda")
douts = [torch.randn_like(outs[0])]
dqks = [torch.empty_like(qk)]
dqkvs = [torch.empty_like(qkv)]
dsum = torch.empty_like(lse)
af = C.cast(ops._seg_array(specs, outs), C.c_void_p)
arrb = ops._seg_array(specs, outs, douts, dqks, dqkvs)
ab = C.cast(arrb, C.c_void_p)
scale = hd ** -0.5
for name, lib in (("r1", old), ("r2", new)):
    trace = torch.zeros(3 * 64, dtype=torch.int64, device="cuda")
    for _ in range(3):
        lib.dlb_attn_fwd_tc(af, 1, lse.data_ptr(), None, 0, B, H, hd, scale, st)
        lib.dlb_attn_bwd_tc(ab, 1, lse.data_ptr(), dsum.data_ptr(), None, 0, B, H, hd, scale, st)
    torch.cuda.synchronize()
    lib.dlb_attn_set_trace(trace.data_ptr())
    lib.dlb_attn_fwd_tc(af, 1, lse.data_ptr(), None, 0, B, H, hd, scale, st)
    lib.dlb_attn_bwd_tc(ab, 1, lse.data_ptr(), dsum.data_ptr(), None, 0, B, H, hd, scale, st)
    torch.cuda.synchronize()
    lib.dlb_attn_set_trace(None)
    t = trace.cpu().tolist()
    for k, kn in enumerate(("fwd", "dq", "dkv")):
        pts = sorted([(v, i) for i, v in enumerate(t[k * 64:(k + 1) * 64]) if v])
        if not pts:
            continue
        t0 = pts[0][0]
        print(f"{name} {kn} hd={hd}: total {pts[-1][0] - t0}  " + " ".join(f"{'m' if 32 <= i < 56 else 'c'}{i}@{v - t0}" for v, i in pts))
