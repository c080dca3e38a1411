"""Does a memory-bound elementwise kernel overlap with a tensor-bound GEMM when they run on two streams?
Times (a) the wgrad GEMM of the MLP up-projection, (b) swiglu_bwd / qknorm-sized streaming work, (c) both concurrently."""
import json
import sys

import torch

sys.path.insert(0, ".")
from diffulab_b200 import ops  # noqa: E402

M, D, F = 32768, 1152, 4608
x = torch.randn(M, D, device="cuda").bfloat16()
dh = torch.randn(M, 2 * F, device="cuda").bfloat16()
dw = torch.zeros(2 * F, D, device="cuda", dtype=torch.float32)
h = torch.randn(M, 2 * F, device="cuda").bfloat16()
dact = torch.randn(M, F, device="cuda").bfloat16()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def gemm_():
    ops.gemm(dh, x, a_mn=True, b_mn=True, out=dw, accumulate=True)  # dW[2F, D] += dh^T x


def ew_():
    ops.swiglu_bwd(dact, h)


def run(which, iters=20):
    for _ in range(3):
        gemm_(); ew_()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    s1.wait_event(e0); s2.wait_event(e0)
    for _ in range(iters):
        if which in ("gemm", "both"):
            with torch.cuda.stream(s1):
                gemm_()
        if which in ("ew", "both"):
            with torch.cuda.stream(s2):
                ew_()
    d1, d2 = torch.cuda.Event(), torch.cuda.Event()
    d1.record(s1); d2.record(s2)
    torch.cuda.current_stream().wait_event(d1); torch.cuda.current_stream().wait_event(d2)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


res = {k: round(run(k), 4) for k in ("gemm", "ew", "both")}
res["sum"] = round(res["gemm"] + res["ew"], 4)
print(json.dumps(res))
