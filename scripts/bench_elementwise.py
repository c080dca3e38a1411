"""Micro-benchmark of the bandwidth-bound kernels at the DiT-XL/2 shape (B=128, N=256, d=1152): CUDA-event time per
C-ABI call, algorithmic bytes (the one-pass backward kernels read each tensor once: LN bwd 4T, gate bwd 3T,
QK-norm bwd 6T elements) and the resulting fraction of the measured HBM peak."""
import json
import os
import sys

import torch

sys.path.insert(0, ".")
from diffulab_b200 import ops  # noqa: E402

BF = torch.bfloat16
B, N, d, H, hd = 128, 256, 1152, 16, 72
R, T = B * N, B * N * d
peaks = json.load(open("MEASURED_PEAKS.json")) if os.path.exists("MEASURED_PEAKS.json") else {"hbm_gbs": 6650.0}
PEAK = peaks["hbm_gbs"]
flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device="cuda")


def timeit(fn, iters=5):
    for _ in range(2):
        fn()
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


def rnd(*s, dt=BF):
    return torch.randn(*s, device="cuda").to(dt)


x, dy, dres = rnd(B, N, d), rnd(B, N, d), rnd(B, N, d)
mod = rnd(B, 6 * d)
w, b = torch.ones(d, device="cuda"), torch.zeros(d, device="cuda")
y, mean, rstd = ops.ln_modulate_fwd(x, w, b, mod[:, :d], mod[:, d:2 * d], 1e-5)
dmod = torch.zeros(B, 6 * d, device="cuda")
dw, db = torch.zeros(d, device="cuda"), torch.zeros(d, device="cuda")
qkv = rnd(R, 3 * d)
pos = torch.stack([torch.arange(N) // 16, torch.arange(N) % 16], -1).int().cuda()
rope = ops.rope_table(pos, [36, 36], 10000.0)
sq, sk = torch.ones(d, device="cuda"), torch.ones(d, device="cuda")
qk, rrms = ops.qknorm_rope_fwd(qkv, sq, sk, rope, hd, tokens_per_sample=N)
dqk, dqkv = rnd(R, 2 * d), torch.empty(R, 3 * d, device="cuda", dtype=BF)
u, ds = rnd(R, 8 * d), rnd(R, 4 * d)
n_par = 823_400_000 // 8  # an eighth of the model is enough to saturate
p32, g32, m32, v32 = (torch.randn(n_par, device="cuda") for _ in range(4))
v32.abs_()
shadow = torch.empty(n_par, device="cuda", dtype=BF)

cases = [
    ("ln_modulate_fwd", 2 * T * 2, lambda: ops.ln_modulate_fwd(x, w, b, mod[:, :d], mod[:, d:2 * d], 1e-5)),
    ("ln_modulate_bwd", 4 * T * 2, lambda: ops.ln_modulate_bwd(dy, x, mean, rstd, w, b, mod[:, :d], dres, dmod[:, :d], dmod[:, d:2 * d], dw, db)),
    ("gate_residual_fwd", 3 * T * 2, lambda: ops.gate_residual_fwd(x, dy, None, mod[:, 2 * d:3 * d])),
    ("gate_residual_bwd", 3 * T * 2, lambda: ops.gate_residual_bwd(dy, x, None, mod[:, 2 * d:3 * d], dmod[:, 2 * d:3 * d])),
    ("swiglu_fwd", 12 * T * 2, lambda: ops.swiglu_fwd(u)),
    ("swiglu_bwd", 20 * T * 2, lambda: ops.swiglu_bwd(ds, u)),
    ("qknorm_rope_fwd", 4 * T * 2, lambda: ops.qknorm_rope_fwd(qkv, sq, sk, rope, hd, tokens_per_sample=N)),
    ("qknorm_rope_bwd", 6 * T * 2, lambda: ops.qknorm_rope_bwd(dqk, qkv, rrms, sq, sk, rope, hd, dqkv, dw, db, tokens_per_sample=N)),
    ("adamw_step", 30 * n_par, lambda: ops.adamw_step(p32, g32, m32, v32, shadow, lr=1e-4, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.01, step=3)),
]
for name, nbytes, fn in cases:
    ms = timeit(fn)
    gbs = nbytes / ms / 1e6
    print(json.dumps({"kernel": name, "ms": round(ms, 4), "algorithmic_MB": round(nbytes / 1e6, 1), "GB_per_s": round(gbs, 1),
                      "frac_of_measured_hbm_peak": round(gbs / PEAK, 3)}), flush=True)
