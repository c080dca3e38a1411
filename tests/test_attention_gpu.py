"""Joint attention kernels vs an fp32 PyTorch restatement of the reference call
(F.scaled_dot_product_attention with scale=hd^-0.5 and a key-padding mask; mmdit.py:92-98, 184-204).
Tolerance: bf16 operands / bf16 P and outputs -> relL2 <= 1e-2 forward, 2e-2 backward."""
import pytest
import torch

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


def rel_l2(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / b.norm().clamp_min(1e-20)).item()


def ref_attention(q, k, v, mask, scale):
    """q,k,v fp32 [B,S,H,hd]; mask bool [B,S] or None -> [B,S,H,hd]"""
    s = torch.einsum("bqhd,bkhd->bhqk", q, k) * scale
    if mask is not None:
        s = s.masked_fill(~mask[:, None, None, :], float("-inf"))
    p = s.softmax(-1)
    return torch.einsum("bhqk,bkhd->bqhd", p, v)


CASES = [
    # B, H, hd, L_text, N_img, masked
    (2, 2, 64, 0, 64, False),
    (2, 16, 72, 0, 256, False),
    (2, 4, 64, 24, 64, True),
    (3, 12, 64, 128, 256, True),
    (2, 3, 72, 7, 50, True),      # ragged lengths: partial tiles on both axes
    (1, 2, 128, 0, 100, False),
    (2, 2, 32, 5, 20, True),
    (2, 4, 72, 100, 300, True),   # S = 400: four 128-key tiles, ragged tail
    (11, 8, 72, 0, 256, False),   # 176 work items > 148 SMs: persistent backward CTAs, uneven items per CTA
    (7, 6, 64, 40, 300, True),    # 126 items x ragged S = 340 (6 streamed tiles): one item per CTA, long ring
    (10, 6, 64, 72, 256, True),   # 180 items, S = 328: persistent + mask + partial tiles
    (2, 4, 128, 0, 256, False),   # swizzled TMA path, head dim 128 (two 128-byte blocks per row)
    (2, 2, 32, 128, 128, True),   # swizzled TMA path, head dim < 64 (out-of-bounds columns arrive as zeros), masked text segment
    (5, 16, 72, 128, 256, True),  # swizzled TMA path, hd 72 (32-byte-swizzled tail block), two segments, 240 items
]


@pytest.mark.parametrize("impl", ["dlb_attn_fwd_tc", "dlb_attn_fwd"])
@pytest.mark.parametrize("bwd_impl", ["dlb_attn_bwd_tc", "dlb_attn_bwd"])
@pytest.mark.parametrize("B,H,hd,L,N,masked", CASES)
def test_attention_fwd_bwd(cuda_device, B, H, hd, L, N, masked, impl, bwd_impl):
    from diffulab_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(B * 100 + H * 10 + hd + L)
    d = H * hd
    S = L + N
    segs_len = [L, N] if L > 0 else [N]
    qks = [(torch.randn(B * l, 2 * d, device="cuda", generator=g)).to(BF) for l in segs_len]
    qkvs = [(torch.randn(B * l, 3 * d, device="cuda", generator=g)).to(BF) for l in segs_len]
    kmask = None
    mask_full = None
    if masked:
        lens = torch.randint(1, L + 1, (B,), device="cuda", generator=g)
        m = torch.arange(L, device="cuda")[None, :] < lens[:, None]
        kmask = m.to(torch.uint8).contiguous()
        mask_full = torch.cat([m, torch.ones(B, N, device="cuda", dtype=torch.bool)], 1)
    scale = hd ** -0.5
    specs = [ops.AttnSegSpec(qk, qkv, l) for qk, qkv, l in zip(qks, qkvs, segs_len)]
    outs, lse = ops.attn_fwd(specs, B, H, hd, scale, kmask, impl=impl)

    def cat(parts, lo, hi):
        return torch.cat([p.view(B, l, -1)[..., lo:hi] for p, l in zip(parts, segs_len)], 1).float()

    q = cat(qks, 0, d).view(B, S, H, hd).requires_grad_(True)
    k = cat(qks, d, 2 * d).view(B, S, H, hd).requires_grad_(True)
    v = cat(qkvs, 2 * d, 3 * d).view(B, S, H, hd).requires_grad_(True)
    ref = ref_attention(q, k, v, mask_full, scale)
    out = torch.cat([o.view(B, l, d) for o, l in zip(outs, segs_len)], 1)
    assert rel_l2(out, ref.reshape(B, S, d)) < 1e-2

    douts = [torch.randn(B * l, d, device="cuda", generator=g).to(BF) for l in segs_len]
    dref = torch.cat([o.view(B, l, d) for o, l in zip(douts, segs_len)], 1).float().view(B, S, H, hd)
    ref.backward(dref)
    dqkvs = [torch.zeros(B * l, 3 * d, device="cuda", dtype=BF) for l in segs_len]
    dqks = ops.attn_bwd(specs, outs, douts, lse, B, H, hd, scale, dqkvs, kmask, impl=bwd_impl)
    dq = torch.cat([x.view(B, l, 2 * d)[..., :d] for x, l in zip(dqks, segs_len)], 1)
    dk = torch.cat([x.view(B, l, 2 * d)[..., d:] for x, l in zip(dqks, segs_len)], 1)
    dv = torch.cat([x.view(B, l, 3 * d)[..., 2 * d :] for x, l in zip(dqkvs, segs_len)], 1)
    assert rel_l2(dq, q.grad.reshape(B, S, d)) < 2e-2
    assert rel_l2(dk, k.grad.reshape(B, S, d)) < 2e-2
    assert rel_l2(dv, v.grad.reshape(B, S, d)) < 2e-2
    for x in dqkvs:
        assert x[:, : 2 * d].abs().max().item() == 0.0  # attention bwd only owns the v columns


@pytest.mark.parametrize("B,H,hd,L,N", [(3, 4, 72, 0, 256), (2, 6, 64, 128, 256), (10, 16, 72, 0, 256)])
def test_tma_and_cp_async_producers_agree_bitwise(cuda_device, B, H, hd, L, N):
    """The TMA (4-D tensor map) and cp.async producer paths stage identical shared-memory images, so forward and backward
    results must be bit-identical; run in a subprocess with DLB_ATTN_NO_TMA=1 to force the fallback path."""
    import os
    import subprocess
    import sys
    import tempfile

    code = f"""
import sys, torch
sys.path.insert(0, {os.getcwd()!r})
from diffulab_b200 import ops
B, H, hd, L, N = {B}, {H}, {hd}, {L}, {N}
g = torch.Generator(device="cuda").manual_seed(7)
d = H * hd
lens = [L, N] if L else [N]
qks = [torch.randn(B * l, 2 * d, device="cuda", generator=g).bfloat16() for l in lens]
qkvs = [torch.randn(B * l, 3 * d, device="cuda", generator=g).bfloat16() for l in lens]
specs = [ops.AttnSegSpec(a, b, l) for a, b, l in zip(qks, qkvs, lens)]
kmask = None
if L:
    kl = torch.randint(1, L + 1, (B,), device="cuda", generator=g)
    kmask = (torch.arange(L, device="cuda")[None, :] < kl[:, None]).to(torch.uint8).contiguous()
outs, lse = ops.attn_fwd(specs, B, H, hd, hd ** -0.5, kmask)
douts = [torch.randn(o.shape, device="cuda", generator=g).bfloat16() for o in outs]
dqkvs = [torch.zeros_like(q) for q in qkvs]
ops.attn_bwd(specs, outs, douts, lse, B, H, hd, hd ** -0.5, dqkvs, kmask)
torch.save({{"outs": [o.cpu() for o in outs], "lse": lse.cpu(), "dqkvs": [t.cpu() for t in dqkvs]}}, sys.argv[1])
"""
    res = []
    for env_extra in ({}, {"DLB_ATTN_NO_TMA": "1"}):
        with tempfile.NamedTemporaryFile(suffix=".pt") as f:
            subprocess.run([sys.executable, "-c", code, f.name], check=True, env={**os.environ, **env_extra}, timeout=300)
            res.append(torch.load(f.name))
    a, b = res
    assert torch.equal(a["lse"], b["lse"])
    for x, y in zip(a["outs"] + a["dqkvs"], b["outs"] + b["dqkvs"]):
        assert torch.equal(x, y)
