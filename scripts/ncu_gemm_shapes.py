"""Every GEMM shape of one DiT-XL/2 block at the bench geometry (M = 128 x 256 = 32768), ONE launch each in a fixed order between
cudaProfilerStart / Stop, with the dispatcher's production defaults (tile width, split-K, CTA pair). The manifest (label, M, N, K,
algorithmic bytes, FLOPs) is written next to the capture so that scripts/ncu_gemm_traffic.py can attribute the k-th GEMM launch
of the ncu report to its shape:

    ncu --profile-from-start off --set full --clock-control none -k regex:gemm -o gpurun_out/ncu_gemm_shapes_r2 -f \\
        python scripts/ncu_gemm_shapes.py gpurun_out/ncu_gemm_shapes_manifest.json
"""
import json
import sys

import torch

sys.path.insert(0, ".")
from diffulab_b200 import ops  # noqa: E402

M, d, F = 32768, 1152, 4608
BF = torch.bfloat16
manifest = []
calls = []


def add(label, m, n, k, fn, extra_bytes=0, out_bytes_per_el=2):
    manifest.append({"label": label, "M": m, "N": n, "K": k, "flops": 2.0 * m * n * k,
                     "algorithmic_bytes": 2 * m * k + 2 * n * k + out_bytes_per_el * m * n + extra_bytes})
    calls.append(fn)


for name, N, K in [("qkv", 3 * d, d), ("proj", d, d), ("mlp_down", d, F)]:
    x = torch.randn(M, K, device="cuda").to(BF)
    w = (torch.randn(N, K, device="cuda") * 0.02).to(BF)
    dy = torch.randn(M, N, device="cuda").to(BF)
    dw = torch.zeros(N, K, device="cuda")
    add(f"{name}_fwd", M, N, K, lambda x=x, w=w: ops.gemm(x, w))
    if name != "mlp_down":  # the down-projection dgrad runs fused with the SwiGLU backward (below)
        add(f"{name}_dgrad", M, K, N, lambda dy=dy, w=w: ops.gemm(dy, w, b_mn=True))
    # wgrad accumulates into fp32 (read + write of the [N, K] fp32 gradient: 8 bytes per element)
    manifest.append({"label": f"{name}_wgrad", "M": N, "N": K, "K": M, "flops": 2.0 * M * N * K, "algorithmic_bytes": 2 * M * N + 2 * M * K + 8 * N * K})
    calls.append(lambda dy=dy, x=x, dw=dw: ops.gemm(dy, x, a_mn=True, b_mn=True, out=dw, accumulate=True))
x = torch.randn(M, d, device="cuda").to(BF)
w1 = (torch.randn(2 * F, d, device="cuda") * 0.02).to(BF)
w2 = (torch.randn(d, F, device="cuda") * 0.02).to(BF)
dy = torch.randn(M, d, device="cuda").to(BF)
h, act = ops.gemm_swiglu(x, w1)
dh = torch.randn(M, 2 * F, device="cuda").to(BF)
dw1 = torch.zeros(2 * F, d, device="cuda")
# fused fc1: writes H [M, 2F] and ACT [M, F]
manifest.append({"label": "mlp_up_swiglu_fwd", "M": M, "N": 2 * F, "K": d, "flops": 2.0 * M * 2 * F * d, "algorithmic_bytes": 2 * M * d + 2 * 2 * F * d + 2 * M * 2 * F + 2 * M * F})
calls.append(lambda: ops.gemm_swiglu(x, w1))
# fused fc2 dgrad: reads dY, W2, H [M, 2F]; writes dH [M, 2F]
manifest.append({"label": "mlp_down_dgrad_swiglu_bwd", "M": M, "N": F, "K": d, "flops": 2.0 * M * F * d, "algorithmic_bytes": 2 * M * d + 2 * d * F + 2 * 2 * M * 2 * F})
calls.append(lambda: ops.gemm_swiglu_bwd(dy, w2, h))
add("mlp_up_dgrad", M, d, 2 * F, lambda: ops.gemm(dh, w1, b_mn=True))
manifest.append({"label": "mlp_up_wgrad", "M": 2 * F, "N": d, "K": M, "flops": 2.0 * M * 2 * F * d, "algorithmic_bytes": 2 * M * 2 * F + 2 * M * d + 8 * 2 * F * d})
calls.append(lambda: ops.gemm(dh, x, a_mn=True, b_mn=True, out=dw1, accumulate=True))

for fn in calls:  # warm-up (tensor-map cache, attribute setup)
    fn()
torch.cuda.synchronize()
flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device="cuda")
torch.cuda.profiler.start()
for fn in calls:
    flush.zero_()  # (a fill kernel, not matched by -k regex:gemm)
    fn()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
with open(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/ncu_gemm_shapes_manifest.json", "w") as f:
    json.dump(manifest, f, indent=1)
print(len(manifest), "gemm launches profiled")
