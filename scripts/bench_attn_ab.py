"""Same-box A/B of the tcgen05 attention kernels: the round-1 library (scripts/ab/libdiffulab_b200_r1.so, built from commit 9d02e24:
4-D tensor maps with 16-byte inner boxes, unswizzled tiles, single P / O buffers) against the current one, alternating launches on
the same tensors so that box-to-box clock differences cancel. DiT-XL/2 and SPRINT/MMDiT geometries."""
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, ".")
from diffulab_b200 import _lib, ops  # noqa: E402

new = _lib.load()
old_path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ab", "libdiffulab_b200_r1.so")
old = C.CDLL(old_path)
p, i32, f32 = C.c_void_p, C.c_int, C.c_float
for lib in (old,):
    lib.dlb_attn_fwd_tc.argtypes = [p, i32, p, p, i32, i32, i32, i32, f32, p]
    lib.dlb_attn_bwd_tc.argtypes = [p, i32, p, p, p, i32, i32, i32, i32, f32, p]
st = torch.cuda.current_stream().cuda_stream


def time_pair(fa, fb, iters=10):
    for _ in range(3):
        fa(); fb()
    torch.cuda.synchronize()
    ta, tb = [], []
    for _ in range(iters):
        for f, acc in ((fa, ta), (fb, tb)):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); f(); e1.record()
            torch.cuda.synchronize()
            acc.append(e0.elapsed_time(e1))
    ta.sort(); tb.sort()
    return ta[len(ta) // 2], tb[len(tb) // 2]


for name, B, H, hd, L, N in (("dit_xl2", 128, 16, 72, 0, 256), ("sprint_mm", 64, 12, 64, 128, 256), ("hd128", 32, 8, 128, 0, 1024)):
    d = H * hd
    lens = [L, N] if L else [N]
    qks = [torch.randn(B * l, 2 * d, device="cuda").bfloat16() for l in lens]
    qkvs = [torch.randn(B * l, 3 * d, device="cuda").bfloat16() for l in lens]
    specs = [ops.AttnSegSpec(a, b, l) for a, b, l in zip(qks, qkvs, lens)]
    S = L + N
    outs = [torch.empty(B * s.len, s.d, device="cuda", dtype=torch.bfloat16) for s in specs]
    lse = torch.empty(B, H, S, device="cuda")
    arr = ops._seg_array(specs, outs)
    ap = C.cast(arr, C.c_void_p)
    scale = hd ** -0.5
    f_old = lambda: old.dlb_attn_fwd_tc(ap, len(specs), lse.data_ptr(), None, 0, B, H, hd, scale, st)
    f_new = lambda: new.dlb_attn_fwd_tc(ap, len(specs), lse.data_ptr(), None, 0, B, H, hd, scale, st)
    assert f_old() == 0 and f_new() == 0
    t_old, t_new = time_pair(f_old, f_new)
    douts = [torch.randn_like(o) for o in outs]
    dqks = [torch.empty(B * s.len, 2 * s.d, device="cuda", dtype=torch.bfloat16) for s in specs]
    dqkvs = [torch.empty_like(q) for q in qkvs]
    dsum = torch.empty_like(lse)
    arrb = ops._seg_array(specs, outs, douts, dqks, dqkvs)
    bp = C.cast(arrb, C.c_void_p)
    b_old = lambda: old.dlb_attn_bwd_tc(bp, len(specs), lse.data_ptr(), dsum.data_ptr(), None, 0, B, H, hd, scale, st)
    b_new = lambda: new.dlb_attn_bwd_tc(bp, len(specs), lse.data_ptr(), dsum.data_ptr(), None, 0, B, H, hd, scale, st)
    assert b_old() == 0 and b_new() == 0
    tb_old, tb_new = time_pair(b_old, b_new)
    flops = 4.0 * B * H * S * S * hd
    print(json.dumps({"case": name, "fwd_ms_r1": round(t_old, 4), "fwd_ms_r2": round(t_new, 4), "fwd_speedup": round(t_old / t_new, 3),
                      "bwd_ms_r1": round(tb_old, 4), "bwd_ms_r2": round(tb_new, 4), "bwd_speedup": round(tb_old / tb_new, 3),
                      "fwd_tflops_r2": round(flops / t_new / 1e9, 1), "bwd_tflops_r2": round(2.5 * flops / tb_new / 1e9, 1)}), flush=True)
