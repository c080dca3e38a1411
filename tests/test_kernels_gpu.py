"""Unit parity of every fused elementwise / reduction kernel against a plain PyTorch fp32 restatement of the
reference op (file:line cited per test). Tolerances: outputs are bf16, so relL2 <= 6e-3 (one or two bf16
roundings, 2^-8 each); fp32 reductions <= 1e-4 relative unless stated."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


def rel_l2(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / b.norm().clamp_min(1e-20)).item()


def mk(gen, *shape, scale=1.0, dtype=BF):
    return (torch.randn(*shape, device="cuda", generator=gen) * scale).to(dtype)


@pytest.fixture()
def gen(cuda_device):
    return torch.Generator(device="cuda").manual_seed(1234)


# ---- LayerNorm + modulate (mmdit.py:257-259,299 ; nn.py:539-540) ------------------------------------------
def ref_ln_mod(x, w, b, scale, shift, eps):
    xf = x.float()
    u = torch.nn.functional.layer_norm(xf, (x.shape[-1],), w, b, eps)
    one_s = (1 + scale).float()  # bf16 add as in the reference, then promoted
    return u * one_s + shift.float()


@pytest.mark.parametrize("B,N,d", [(2, 16, 64), (3, 50, 1152), (2, 64, 768), (4, 33, 512), (1, 7, 2048),
                                   # lean warp-per-row kernels (rowwise_lean.cuh): exact per-lane unit split, rows_per_mod % 8 == 0;
                                   # (200, 8, 256): a CTA's row range crosses sample boundaries (accumulator flush + P/Q rebuild)
                                   (3, 256, 1152), (2, 128, 640), (200, 8, 256), (5, 72, 1152)])
@pytest.mark.parametrize("affine", [True, False])
def test_ln_modulate_fwd_bwd(gen, B, N, d, affine):
    from diffulab_b200 import ops

    x = mk(gen, B, N, d)
    mod = mk(gen, B, 6 * d, scale=0.5)
    scale, shift = mod[:, 0:d], mod[:, d : 2 * d]
    w = (1 + 0.1 * torch.randn(d, device="cuda", generator=gen)) if affine else None
    b = (0.1 * torch.randn(d, device="cuda", generator=gen)) if affine else None
    eps = 1e-5 if affine else 1e-6
    y, mean, rstd = ops.ln_modulate_fwd(x, w, b, scale, shift, eps)
    ref = ref_ln_mod(x, w, b, scale[:, None, :], shift[:, None, :], eps)
    assert rel_l2(y, ref) < 6e-3

    # backward vs autograd of the fp32 restatement
    xr = x.float().requires_grad_(True)
    sr = scale.float().requires_grad_(True)
    hr = shift.float().requires_grad_(True)
    wr = w.clone().requires_grad_(True) if affine else None
    br = b.clone().requires_grad_(True) if affine else None
    u = torch.nn.functional.layer_norm(xr, (d,), wr, br, eps)
    out = u * (1 + sr)[:, None, :] + hr[:, None, :]
    dy = mk(gen, B, N, d)
    dres = mk(gen, B, N, d)
    out.backward(dy.float())
    dmod = torch.zeros(B, 6 * d, device="cuda")
    dw = torch.zeros(d, device="cuda") if affine else None
    db = torch.zeros(d, device="cuda") if affine else None
    dx = ops.ln_modulate_bwd(dy, x, mean, rstd, w, b, scale, dres, dmod[:, 0:d], dmod[:, d : 2 * d], dw, db)
    assert rel_l2(dx, xr.grad + dres.float()) < 8e-3
    assert rel_l2(dmod[:, 0:d], sr.grad) < 2e-2  # (1+scale) is bf16-rounded in the kernel as in the reference
    assert rel_l2(dmod[:, d : 2 * d], hr.grad) < 1e-4
    assert dmod[:, 2 * d :].abs().max().item() == 0.0
    if affine:
        assert rel_l2(dw, wr.grad) < 1e-2
        assert rel_l2(db, br.grad) < 1e-2


def test_ln_modulate_per_token(gen):
    """DDT decoder: modulation is per token (ddt.py:421-459, Modulation with a 3-D input nn.py:530-536)."""
    from diffulab_b200 import ops

    B, N, d = 2, 24, 640
    x = mk(gen, B, N, d)
    mod = mk(gen, B, N, 6 * d, scale=0.5)
    scale, shift = mod[..., 0:d], mod[..., d : 2 * d]
    w = 1 + 0.1 * torch.randn(d, device="cuda", generator=gen)
    b = 0.1 * torch.randn(d, device="cuda", generator=gen)
    y, mean, rstd = ops.ln_modulate_fwd(x, w, b, scale, shift, 1e-5)
    assert rel_l2(y, ref_ln_mod(x, w, b, scale, shift, 1e-5)) < 6e-3
    xr = x.float().requires_grad_(True)
    sr = scale.float().requires_grad_(True)
    hr = shift.float().requires_grad_(True)
    wr, br = w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    out = torch.nn.functional.layer_norm(xr, (d,), wr, br, 1e-5) * (1 + sr) + hr
    dy = mk(gen, B, N, d)
    out.backward(dy.float())
    dmod = torch.zeros(B, N, 6 * d, device="cuda", dtype=BF)
    dw, db = torch.zeros(d, device="cuda"), torch.zeros(d, device="cuda")
    dx = ops.ln_modulate_bwd(dy, x, mean, rstd, w, b, scale, None, dmod[..., 0:d], dmod[..., d : 2 * d], dw, db)
    assert rel_l2(dx, xr.grad) < 8e-3
    assert rel_l2(dmod[..., 0:d], sr.grad) < 8e-3
    assert rel_l2(dmod[..., d : 2 * d], hr.grad) < 6e-3
    assert rel_l2(dw, wr.grad) < 1e-2 and rel_l2(db, br.grad) < 1e-2


# ---- gated residual (mmdit.py:296-307, 524-531) -----------------------------------------------------------
@pytest.mark.parametrize("B,N,d,two", [(2, 16, 64, False), (3, 50, 1152, False), (2, 40, 768, True),
                                       (3, 256, 1152, False), (200, 8, 256, True), (2, 128, 640, True), (2, 64, 1152, True)])
def test_gate_residual(gen, B, N, d, two):
    from diffulab_b200 import ops

    x, a1 = mk(gen, B, N, d), mk(gen, B, N, d)
    a2 = mk(gen, B, N, d) if two else None
    mod = mk(gen, B, 3 * d, scale=0.5)
    gate = mod[:, 2 * d :]
    out = ops.gate_residual_fwd(x, a1, a2, gate)
    a = a1.float() + (a2.float() if two else 0)
    ref = x.float() + a * gate.float()[:, None, :]
    assert rel_l2(out, ref) < 6e-3
    dout = mk(gen, B, N, d)
    dmod = torch.zeros(B, 3 * d, device="cuda")
    da = ops.gate_residual_bwd(dout, a1, a2, gate, dmod[:, 2 * d :])
    assert rel_l2(da, dout.float() * gate.float()[:, None, :]) < 6e-3
    assert rel_l2(dmod[:, 2 * d :], (dout.float() * a).sum(1)) < 6e-3
    assert dmod[:, : 2 * d].abs().max().item() == 0.0


# ---- SwiGLU (nn.py:478-486) -------------------------------------------------------------------------------
@pytest.mark.parametrize("R,F", [(64, 256), (300, 4608), (17, 3072)])
def test_swiglu(gen, R, F):
    from diffulab_b200 import ops

    h = mk(gen, R, 2 * F)
    out = ops.swiglu_fwd(h)
    hr = h.float().requires_grad_(True)
    a, g = hr.chunk(2, dim=-1)
    ref = torch.nn.functional.silu(a) * g
    assert rel_l2(out, ref) < 6e-3
    dout = mk(gen, R, F)
    ref.backward(dout.float())
    assert rel_l2(ops.swiglu_bwd(dout, h), hr.grad) < 6e-3


# ---- QK RMSNorm + N-D RoPE (nn.py:262-400, 423-475 ; mmdit.py:81-89) --------------------------------------
def ref_rope_tables(pos, axes_dim, base):
    cs, sn = [], []
    for i, ad in enumerate(axes_dim):
        p = pos[:, i].double()
        freqs = 1.0 / (base ** (torch.arange(0, ad, 2, dtype=torch.float64, device=pos.device) / ad))
        ang = p[:, None] * freqs[None]
        cs.append(ang.cos().float())
        sn.append(ang.sin().float())
    return torch.cat(cs, -1), torch.cat(sn, -1)


def ref_qknorm_rope(x, s, cos, sin, H, hd):
    """x [B,S,d] fp32 -> normalised, scaled, rotated [B,S,d] (fp32 math)."""
    B, S, d = x.shape
    rr = torch.rsqrt((x * x).mean(-1, keepdim=True) + 1e-6)
    y = (x * rr) * s
    y = y.view(B, S, H, hd)
    R = cos.shape[-1] * 2
    yr, yp = y[..., :R], y[..., R:]
    e, o = yr[..., 0::2], yr[..., 1::2]
    c, sn_ = cos[None, :, None, :], sin[None, :, None, :]
    rot = torch.stack([e * c - o * sn_, e * sn_ + o * c], -1).flatten(-2)
    return torch.cat([rot, yp], -1).reshape(B, S, d)


@pytest.mark.parametrize("B,S,H,hd,axes", [(2, 16, 2, 32, [16, 16]), (2, 64, 16, 72, [36, 36]), (2, 40, 12, 64, [16, 24, 24]), (1, 9, 8, 64, [8, 8]),
                                           (3, 256, 16, 72, [36, 36]), (2, 32, 4, 64, [16, 16]), (2, 129, 10, 64, [20, 22, 22])])
def test_qknorm_rope(gen, B, S, H, hd, axes):
    from diffulab_b200 import ops

    d = H * hd
    n_ax = len(axes)
    pos = torch.stack([torch.randint(0, 17, (S,), device="cuda", generator=gen) for _ in range(n_ax)], -1).int()
    rope = ops.rope_table(pos, axes, 10000.0)
    cr, sr_ = ref_rope_tables(pos, axes, 10000.0)
    assert (rope.cos - cr).abs().max().item() < 2e-6 and (rope.sin - sr_).abs().max().item() < 2e-6
    packed = torch.stack([rope.cos.to(BF), rope.sin.to(BF)], -1).view(torch.int32).squeeze(-1)
    assert torch.equal(rope.cs, packed)
    qkv = mk(gen, B * S, 3 * d)
    sq = 1 + 0.2 * torch.randn(d, device="cuda", generator=gen)
    sk = 1 + 0.2 * torch.randn(d, device="cuda", generator=gen)
    out, rrms = ops.qknorm_rope_fwd(qkv, sq, sk, rope, hd, tokens_per_sample=S)
    x = qkv.float().view(B, S, 3 * d).requires_grad_(True)
    sqr, skr = sq.clone().requires_grad_(True), sk.clone().requires_grad_(True)
    cb, sb = cr.to(BF).float(), sr_.to(BF).float()  # the reference casts cos/sin to the activation dtype
    qr = ref_qknorm_rope(x[..., :d], sqr, cb, sb, H, hd)
    kr = ref_qknorm_rope(x[..., d : 2 * d], skr, cb, sb, H, hd)
    ref = torch.cat([qr, kr], -1).view(B * S, 2 * d)
    assert rel_l2(out, ref) < 1e-2  # three chained bf16 roundings in the reference's own dtype flow
    dqk = mk(gen, B * S, 2 * d)
    ref.backward(dqk.float())
    dqkv = torch.zeros(B * S, 3 * d, device="cuda", dtype=BF)
    dsq, dsk = torch.zeros(d, device="cuda"), torch.zeros(d, device="cuda")
    ops.qknorm_rope_bwd(dqk, qkv, rrms, sq, sk, rope, hd, dqkv, dsq, dsk, tokens_per_sample=S)
    g = x.grad.view(B * S, 3 * d)
    assert rel_l2(dqkv[:, : 2 * d], g[:, : 2 * d]) < 1e-2
    assert dqkv[:, 2 * d :].abs().max().item() == 0.0
    assert rel_l2(dsq, sqr.grad) < 1e-2 and rel_l2(dsk, skr.grad) < 1e-2


def test_qknorm_rope_pos_idx(gen):
    """SPRINT: RoPE rows gathered by kept indices (sprint.py:348-352, 460-465) == table lookup through pos_idx."""
    from diffulab_b200 import ops

    B, S, H, hd, k = 2, 32, 4, 32, 8
    d = H * hd
    pos = torch.stack([torch.arange(S, device="cuda") // 8, torch.arange(S, device="cuda") % 8], -1).int()
    rope = ops.rope_table(pos, [16, 16], 10000.0)
    kept = torch.stack([torch.randperm(S, device="cuda", generator=gen)[:k].sort().values for _ in range(B)]).int()
    qkv = mk(gen, B * k, 3 * d)
    ones = torch.ones(d, device="cuda")
    a, _ = ops.qknorm_rope_fwd(qkv, ones, ones, rope, hd, tokens_per_sample=k, pos_idx=kept.reshape(-1).contiguous())
    for b in range(B):
        single, _ = ops.qknorm_rope_fwd(qkv[b * k : (b + 1) * k].contiguous(), ones, ones, rope.rows(kept[b].long()), hd, tokens_per_sample=k)
        assert torch.equal(a[b * k : (b + 1) * k], single)


# ---- glue kernels -----------------------------------------------------------------------------------------
def test_timestep_embed_and_cond(gen):
    """timestep_embedding nn.py:91-114 ; LabelEmbed add mmdit.py:866-868."""
    from diffulab_b200 import ops

    B, dim, E = 8, 256, 128
    t = torch.rand(B, device="cuda", generator=gen)
    te = ops.timestep_embed(t, dim)
    half = dim // 2
    freqs = torch.exp(-math.log(10000) * torch.arange(half, dtype=torch.float32, device="cuda") / half)
    args = t[:, None] * freqs[None]
    ref = torch.cat([args.cos(), args.sin()], -1)
    assert rel_l2(te, ref) < 4e-3
    tb = mk(gen, B, E)
    table = torch.randn(11, E, device="cuda", generator=gen)
    labels = torch.randint(0, 11, (B,), device="cuda", generator=gen)
    emb, es = ops.cond_combine(tb, table, labels)
    refe = tb.float() + table[labels]
    assert rel_l2(emb, refe) < 1e-6
    assert rel_l2(es, torch.nn.functional.silu(refe)) < 4e-3
    g = torch.randn(B, E, device="cuda", generator=gen)
    dt = torch.zeros_like(table)
    ops.embedding_bwd(g, labels, dt)
    assert rel_l2(dt, torch.zeros_like(table).index_add_(0, labels, g)) < 1e-6


@pytest.mark.parametrize("B,C,H,W,p", [(2, 4, 32, 32, 2), (3, 3, 32, 32, 2), (2, 128, 16, 16, 1), (1, 1, 8, 8, 4)])
def test_patchify_unpatchify(gen, B, C, H, W, p):
    """patchify == im2col of Conv2d(k=s=p) (mmdit.py:697-699,760-763); unpatchify mmdit.py:778-786."""
    from einops import rearrange

    from diffulab_b200 import ops

    x = torch.randn(B, C, H, W, device="cuda", generator=gen)
    P = ops.patchify(x, p)
    w = torch.randn(40, C, p, p, device="cuda", generator=gen)
    conv = torch.nn.functional.conv2d(x.to(BF).float(), w.to(BF).float(), stride=p)
    ref = rearrange(conv, "b c h w -> (b h w) c")
    wk = torch.zeros(40, P.shape[1], device="cuda")
    wk[:, : C * p * p] = w.to(BF).float().reshape(40, -1)
    assert rel_l2(P.float() @ wk.t(), ref) < 1e-5
    tok = mk(gen, B * (H // p) * (W // p), p * p * C)
    img = ops.unpatchify(tok, B, C, H, W, p)
    ref_img = rearrange(tok.view(B, -1, p * p * C), "b (h w) (p1 p2 c) -> b c (h p1) (w p2)", h=H // p, w=W // p, p1=p, p2=p, c=C)
    assert torch.equal(img, ref_img)
    back = ops.patchify_grad(img, p, ld=p * p * C)
    assert torch.equal(back, tok)


def test_colsum_silu_cast(gen):
    from diffulab_b200 import ops

    x = mk(gen, 1000, 200)
    out = torch.ones(200, device="cuda")
    ops.colsum_(x, out)
    assert rel_l2(out, 1 + x.float().sum(0)) < 1e-4
    xf = torch.randn(4096, 24, device="cuda", generator=gen)
    out = torch.zeros(16, device="cuda")
    ops.colsum_(xf[:, 8:], out)
    assert rel_l2(out, xf[:, 8:].sum(0)) < 1e-4
    assert rel_l2(ops.silu_fwd(xf), torch.nn.functional.silu(xf)) < 4e-3
    xr = xf.clone().requires_grad_(True)
    dy = torch.randn_like(xf)
    torch.nn.functional.silu(xr).backward(dy)
    assert rel_l2(ops.silu_bwd(dy, xf, torch.float32), xr.grad) < 1e-5
    assert torch.equal(ops.cast_bf16(xf), xf.to(BF))
    padded = ops.cast_bf16(xf[:, :12].contiguous(), ld_out=16)
    assert torch.equal(padded[:, :12], xf[:, :12].to(BF)) and padded[:, 12:].abs().max().item() == 0


# ---- flow matching (flow.py:262-315, 382-408) -------------------------------------------------------------
@pytest.mark.parametrize("pred_dtype", [BF, torch.float32])
def test_flow_interp_and_loss(gen, pred_dtype):
    from diffulab_b200 import ops

    B, C, H, W = 6, 4, 32, 32
    x0 = torch.randn(B, C, H, W, device="cuda", generator=gen)
    eps = torch.randn(B, C, H, W, device="cuda", generator=gen)
    t = torch.rand(B, device="cuda", generator=gen)
    xt = ops.interp(x0, eps, 1 - t, t)
    ref_xt = (1 - t).view(-1, 1, 1, 1) * x0 + t.view(-1, 1, 1, 1) * eps
    assert rel_l2(xt, ref_xt) < 1e-6
    pred = mk(gen, B, C, H, W, dtype=pred_dtype)
    loss = ops.mse_fwd(pred, x0, eps)
    pr = pred.float().clone().requires_grad_(True)
    ref = (((eps - x0) - pr) ** 2).reshape(B, -1).mean(-1).mean()
    assert abs(loss.item() - ref.item()) < 1e-5 * abs(ref.item())
    ref.backward()
    gout = torch.tensor(0.5, device="cuda")
    dp = ops.mse_bwd(pred, x0, eps, gout)
    assert rel_l2(dp, 0.5 * pr.grad) < (6e-3 if pred_dtype == BF else 1e-6)
    # x-prediction variant (flow.py:300-303)
    tc = t.clamp(min=0.05)
    loss_x = ops.mse_fwd(pred, x0, eps, xt=xt, t=tc)
    pr2 = pred.float().clone().requires_grad_(True)
    v = (xt - pr2) / tc.view(-1, 1, 1, 1)
    ref2 = (((eps - x0) - v) ** 2).reshape(B, -1).mean(-1).mean()
    assert abs(loss_x.item() - ref2.item()) < 1e-4 * abs(ref2.item())
    ref2.backward()
    assert rel_l2(ops.mse_bwd(pred, x0, eps, None, xt=xt, t=tc), pr2.grad) < (6e-3 if pred_dtype == BF else 1e-5)
    # DDPM-style: target = eps (gaussian_diffusion.py:268-311)
    loss_e = ops.mse_fwd(pred, None, eps)
    assert abs(loss_e.item() - ((eps - pred.float()) ** 2).mean().item()) < 1e-5


# ---- REPA cosine (repa.py:183-185) ------------------------------------------------------------------------
@pytest.mark.parametrize("R,E", [(64, 384), (500, 1024), (8, 768)])
def test_repa_cosine(gen, R, E):
    from diffulab_b200 import ops

    s = mk(gen, R, E)
    z = torch.randn(R, E, device="cuda", generator=gen)
    loss = ops.repa_cos_fwd(s, z, 0.5)
    sr = s.float().requires_grad_(True)
    ref = 0.5 * (1 - torch.nn.functional.cosine_similarity(sr, z, dim=-1).mean())
    assert abs(loss.item() - ref.item()) < 1e-5
    ref.backward()
    assert rel_l2(ops.repa_cos_bwd(s, z, 0.5, None), sr.grad) < 6e-3


# ---- SPRINT (sprint.py:317-387) ---------------------------------------------------------------------------
@pytest.mark.parametrize("B,S,k", [(4, 256, 64), (3, 100, 25), (2, 64, 1), (2, 300, 300), (1, 1024, 256)])
def test_sprint_select_bit_exact(gen, B, S, k):
    from diffulab_b200 import ops

    scores = torch.rand(B, S, device="cuda", generator=gen)
    kept, kept32, inv = ops.sprint_select(scores, k)
    idx = torch.topk(scores, k=k, dim=1, largest=True, sorted=False).indices.to(torch.long)
    ref = torch.gather(idx, 1, torch.argsort(idx, dim=1))
    assert torch.equal(kept, ref)  # int64, bit exact
    assert torch.equal(kept32.long(), ref)
    inv_ref = torch.full((B, S), -1, device="cuda", dtype=torch.int32)
    inv_ref.scatter_(1, ref, torch.arange(k, device="cuda", dtype=torch.int32).expand(B, k))
    assert torch.equal(inv, inv_ref)


def test_sprint_gather_restore(gen):
    from diffulab_b200 import ops

    B, S, k, d = 3, 64, 16, 128
    x = mk(gen, B, S, d)
    scores = torch.rand(B, S, device="cuda", generator=gen)
    kept, _, inv = ops.sprint_select(scores, k)
    xk = ops.gather_rows(x, kept)
    assert torch.equal(xk, torch.gather(x, 1, kept.unsqueeze(-1).expand(B, k, d)))
    mask_token = torch.randn(d, device="cuda", generator=gen)
    drop = torch.tensor([0, 1, 0], device="cuda", dtype=torch.uint8)
    full = ops.restore_rows(xk, inv, mask_token, drop)
    ref = mask_token.to(BF).expand(B, S, d).clone()
    ref.scatter_(1, kept.unsqueeze(-1).expand(-1, -1, d), xk)
    ref = torch.where(drop.bool()[:, None, None], mask_token.to(BF).expand_as(ref), ref)
    assert torch.equal(full, ref)
    dy = mk(gen, B, S, d)
    dfill = torch.zeros(d, device="cuda")
    dxk = ops.restore_rows_bwd(dy, kept, inv, drop, dfill)
    ref_dxk = torch.gather(dy, 1, kept.unsqueeze(-1).expand(B, k, d)) * (1 - drop.float())[:, None, None].to(BF)
    assert torch.equal(dxk, ref_dxk)
    filled = (inv < 0) | drop.bool()[:, None]
    assert rel_l2(dfill, (dy.float() * filled[..., None]).sum((0, 1))) < 1e-4
    # gather backward == restore with zero fill
    dz = ops.restore_rows(xk, inv, None, None)
    ref0 = torch.zeros(B, S, d, device="cuda", dtype=BF).scatter_(1, kept.unsqueeze(-1).expand(-1, -1, d), xk)
    assert torch.equal(dz, ref0)


# ---- Euler + CFG (euler.py:37-39 ; flow.py:256-260) -------------------------------------------------------
@pytest.mark.parametrize("vdt", [BF, torch.float32])
def test_euler_step(gen, vdt):
    from diffulab_b200 import ops

    x = torch.randn(4, 4, 32, 32, device="cuda", generator=gen)
    vc, vu = mk(gen, 4, 4, 32, 32, dtype=vdt), mk(gen, 4, 4, 32, 32, dtype=vdt)
    xp, x0 = ops.euler_step(x, vc, None, 0.0, 0.8, 0.78)
    assert rel_l2(xp, x - vc.float() * (0.8 - 0.78)) < 1e-6 and rel_l2(x0, x - vc.float() * 0.8) < 1e-6
    xp, x0 = ops.euler_step(x, vc, vu, 4.0, 0.8, 0.78)
    v = vu.float() + 4.0 * (vc.float() - vu.float())
    assert rel_l2(xp, x - v * (0.8 - 0.78)) < 1e-6 and rel_l2(x0, x - v * 0.8) < 1e-6


def test_adamw_matches_torch(gen):
    """torch.optim.AdamW(lr 1e-4, wd 0.01) as configured in configs/optimizer/adamw.yaml."""
    from diffulab_b200 import ops

    n = 10000
    p0 = torch.randn(n, device="cuda", generator=gen)
    pt = p0.clone().requires_grad_(True)
    opt = torch.optim.AdamW([pt], lr=1e-3, weight_decay=0.01, betas=(0.9, 0.999), eps=1e-8)
    p, m, v = p0.clone(), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    shadow = torch.empty(n, device="cuda", dtype=BF)
    for step in range(1, 4):
        g = torch.randn(n, device="cuda", generator=gen)
        pt.grad = g.clone()
        opt.step()
        ops.adamw_step(p, g, m, v, shadow, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.01, step=step)
    assert rel_l2(p, pt.detach()) < 1e-6
    assert torch.equal(shadow, p.to(BF))
