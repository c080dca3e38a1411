"""Block-level forward/backward of the denoiser hot path, written as explicit kernel sequences.

Each transformer block is ONE `torch.autograd.Function` whose backward is hand-scheduled over the C-ABI kernels
(ops.py): the reference gets these backward passes from autograd over ~60 ATen ops per block
(networks/denoisers/mmdit.py:288-309, 416-459, 499-532); here every step is a fused sm_100a kernel and the
adaLN gradients of a block are reduced into one fp32 [B, k*d] buffer by the kernels themselves.

Parameter handling: parameters stay fp32 `nn.Parameter`s with the reference's names (so `state_dict`, AdamW and
EMA work unchanged); GEMMs read a bf16 shadow (what CUDA autocast does in the reference,
training/trainers/base_trainer.py:309). Weight gradients are accumulated by the wgrad GEMM / reduction kernels
straight into `param.grad` (fp32), so the Functions return None for parameters.
"""

from __future__ import annotations

from typing import Callable

import torch
from torch import Tensor, nn
from torch.utils.weak import WeakIdKeyDictionary

from . import ops

BF16 = torch.bfloat16
F32 = torch.float32

# ---------------------------------------------------------------------------------------------------------
# bf16 weight shadows + gradient sinks
# ---------------------------------------------------------------------------------------------------------
_shadow: WeakIdKeyDictionary = WeakIdKeyDictionary()  # param -> (version, epoch, bf16 copy)
_shadow_epoch = 0
_grad_ready_hook: Callable[[Tensor], None] | None = None


def bump_shadow_epoch() -> None:
    """Invalidate every cached bf16 shadow (call after parameters were modified behind autograd's back)."""
    global _shadow_epoch
    _shadow_epoch += 1


def install_shadow(p: Tensor, shadow: Tensor) -> None:
    """Register an externally maintained bf16 copy of `p` (a view of FlatParams.flat_shadow, rewritten by the fused
    AdamW kernel every step). If `p` is later modified through the torch API (load_state_dict, copy_), its version
    counter changes and the shadow is re-synchronised on next use."""
    _shadow[p] = (p._version, -2, shadow)


def set_grad_ready_hook(fn: Callable[[Tensor], None] | None) -> None:
    """Called with each parameter right after its gradient for this backward pass has been fully accumulated
    (used by the data-parallel gradient reducer to launch bucket all-reduces during backward)."""
    global _grad_ready_hook
    _grad_ready_hook = fn


def wb(p: Tensor, pad_cols: int | None = None) -> Tensor:
    """bf16 shadow of an fp32 parameter as a 2-D [out, in] matrix (conv kernels flattened, optionally K-padded)."""
    ent = _shadow.get(p)
    if ent is not None:
        if ent[1] == -2:  # installed flat shadow
            if ent[0] != p._version:
                src = p.detach().contiguous()
                ops._lib_call("dlb_cast_f32_bf16", src.data_ptr(), ent[2].data_ptr(), 1, src.numel(), src.numel(), ops._stream())
                _shadow[p] = (p._version, -2, ent[2])
            return ent[2]
        if ent[0] == p._version and ent[1] == _shadow_epoch:
            return ent[2]
    p2 = p.detach().reshape(p.shape[0], -1) if p.dim() > 1 else p.detach().reshape(1, -1)
    cols = p2.shape[1]
    ld = pad_cols if pad_cols is not None else cols
    s = ops.cast_bf16(p2.contiguous(), ld_out=ld)
    s = s.view(p2.shape[0], ld)
    _shadow[p] = (p._version, _shadow_epoch, s)
    return s


# ids of the parameters whose gradient buffer was handed to a kernel since the last `clear_touched()` (= the parameters
# whose .grad would be non-None in the reference after this backward; the optimizer skips the others like torch does)
_touched: set[int] = set()


def clear_touched() -> None:
    _touched.clear()


def gbuf(p: Tensor) -> Tensor:
    """fp32 gradient accumulator of a parameter (allocated zeroed on first use)."""
    _touched.add(id(p))
    if p.grad is None:
        p.grad = torch.zeros_like(p, memory_format=torch.contiguous_format)
    return p.grad


def _ready(p: Tensor) -> None:
    if _grad_ready_hook is not None:
        _grad_ready_hook(p)


def wgrad_(p: Tensor, dy: Tensor, x: Tensor) -> None:
    """p.grad[N,K] += dy[R,N]^T @ x[R,K] — both operands read MN-major straight from the activations."""
    if not p.requires_grad:
        return
    N, K = dy.shape[1], x.shape[1]
    g = gbuf(p)
    if g.numel() == N * K:
        ops.gemm(dy, x, a_mn=True, b_mn=True, out=g.view(N, K), accumulate=True)
    else:  # K-padded operand (conv kernels whose C*p*p is not a multiple of 8)
        tmp = torch.zeros(N, K, device=g.device, dtype=F32)
        ops.gemm(dy, x, a_mn=True, b_mn=True, out=tmp, accumulate=True)
        g.view(N, -1).add_(tmp[:, : g.numel() // N])
    _ready(p)


def bgrad_(p: Tensor | None, dy: Tensor) -> None:
    if p is None or not p.requires_grad:
        return
    ops.colsum_(dy, gbuf(p))
    _ready(p)


def vgrad(p: Tensor | None) -> Tensor | None:
    """Vector-parameter gradient buffer handed to kernels that accumulate atomically (LN affine, RMS scales)."""
    if p is None or not p.requires_grad:
        return None
    return gbuf(p)


def linear_fwd(x2: Tensor, w: Tensor, b: Tensor | None = None) -> Tensor:
    return ops.gemm(x2, wb(w), bias=b.detach() if b is not None else None)


def _dgrad(dy2: Tensor, w: Tensor) -> Tensor:
    """dx = dy @ W. A short-and-wide dgrad (the adaLN linears: 128 x 1152 from K = 6912) has a handful of output tiles and a
    very long reduction; it goes through the fp32 split-K accumulate path (all SMs busy) and one cast instead of a
    nine-CTA launch (32 -> ~11 us per block)."""
    M, K = dy2.shape
    if M <= 256 and K >= 2048:
        acc = torch.zeros(M, w.shape[1], device=dy2.device, dtype=F32)
        ops.gemm(dy2, wb(w), b_mn=True, out=acc, accumulate=True)
        return ops.cast_bf16(acc)
    return ops.gemm(dy2, wb(w), b_mn=True)


def linear_bwd(dy2: Tensor, x2: Tensor | None, w: Tensor, b: Tensor | None, need_dx: bool = True) -> Tensor | None:
    dx = _dgrad(dy2, w) if need_dx else None
    if x2 is not None:
        wgrad_(w, dy2, x2)
    bgrad_(b, dy2)
    return dx


# ---------------------------------------------------------------------------------------------------------
# generic Linear (+bias) on bf16 activations: modulation / adaLN linears, context_embed, fuse, time MLP pieces
# ---------------------------------------------------------------------------------------------------------
class LinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x: Tensor, w: Tensor, b: Tensor | None):
        x2 = x.reshape(-1, x.shape[-1])
        y = linear_fwd(x2, w, b)
        ctx.save_for_backward(x2)
        ctx.w, ctx.b, ctx.xshape = w, b, x.shape
        return y.view(*x.shape[:-1], y.shape[-1])

    @staticmethod
    def backward(ctx, dy: Tensor):
        (x2,) = ctx.saved_tensors
        dy2 = dy.reshape(-1, dy.shape[-1])
        if not dy2.is_contiguous():
            dy2 = dy2.contiguous()
        dx = linear_bwd(dy2, x2, ctx.w, ctx.b, need_dx=ctx.needs_input_grad[0])
        return (dx.view(ctx.xshape) if dx is not None else None), None, None


def linear(x: Tensor, w: Tensor, b: Tensor | None = None) -> Tensor:
    return LinearFn.apply(x, w, b)


# ---------------------------------------------------------------------------------------------------------
# conditioning vector: timestep MLP (+ label embedding) -> silu(emb) in bf16   (mmdit.py:866-868; nn.py:531)
# ---------------------------------------------------------------------------------------------------------
class CondFn(torch.autograd.Function):
    """Returns (emb_silu bf16 [B,E], te bf16 [B,E]): silu of the full conditioning vector (what every adaLN linear
    consumes) and the bare time embedding (DDT decoder conditioning, ddt.py:421)."""

    @staticmethod
    def forward(ctx, t: Tensor, labels: Tensor | None, freq_dim: int, w0, b0, w2, b2, table):
        ctx.set_materialize_grads(False)
        f = ops.timestep_embed(t.to(F32).contiguous(), freq_dim)
        h = linear_fwd(f, w0, b0)
        hs = ops.silu_fwd(h)
        te = linear_fwd(hs, w2, b2)
        emb, es = ops.cond_combine(te, table.detach() if table is not None else None, labels)
        ctx.save_for_backward(f, h, hs, emb)
        ctx.p = (w0, b0, w2, b2, table)
        ctx.labels = labels
        return es, te

    @staticmethod
    def backward(ctx, d_es: Tensor | None, d_te: Tensor | None):
        f, h, hs, emb = ctx.saved_tensors
        w0, b0, w2, b2, table = ctx.p
        demb = None
        if d_es is not None:
            demb = ops.silu_bwd(d_es.contiguous(), emb, F32)  # fp32 [B,E]
            if table is not None and table.requires_grad:
                ops.embedding_bwd(demb, ctx.labels, gbuf(table))
                _ready(table)
        if demb is not None:
            dte = ops.cast_bf16(demb)
            if d_te is not None:
                dte = dte + d_te  # tiny [B,E]
        else:
            dte = d_te.contiguous()
        dhs = linear_bwd(dte, hs, w2, b2)
        dh = ops.silu_bwd(dhs, h, BF16)
        linear_bwd(dh, f, w0, b0, need_dx=False)
        return None, None, None, None, None, None, None, None


# ---------------------------------------------------------------------------------------------------------
# patch embedding (Conv2d k=s=p as im2col + GEMM)   mmdit.py:697-699, 757-765
# ---------------------------------------------------------------------------------------------------------
class PatchEmbedFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x: Tensor, w: Tensor, p: int):
        patches = ops.patchify(x.to(F32).contiguous(), p)
        tok = ops.gemm(patches, wb(w, pad_cols=patches.shape[1]))
        ctx.save_for_backward(patches)
        ctx.w = w
        B, _, H, W = x.shape
        return tok.view(B, (H // p) * (W // p), -1)

    @staticmethod
    def backward(ctx, dtok: Tensor):
        (patches,) = ctx.saved_tensors
        d2 = dtok.reshape(-1, dtok.shape[-1])
        wgrad_(ctx.w, d2 if d2.is_contiguous() else d2.contiguous(), patches)
        return None, None, None


# ---------------------------------------------------------------------------------------------------------
# attention sub-layer over 1 or 2 streams (qkv GEMM -> QK-norm + RoPE -> joint attention -> out proj)
# ---------------------------------------------------------------------------------------------------------
RopeCtx = ops.RopeTable  # RoPE tables for one forward (fp32 cos / sin + the packed bf16x2 table the kernels read)


class StreamSpec:
    """One token stream entering attention: its projection weights and its position mapping."""

    def __init__(self, w_qkv, sq, sk, w_out, length: int, pos_offset: int = 0, pos_idx: Tensor | None = None):
        self.w_qkv, self.sq, self.sk, self.w_out = w_qkv, sq, sk, w_out
        self.len, self.pos_offset, self.pos_idx = length, pos_offset, pos_idx


def attn_fwd(hs: list[Tensor], streams: list[StreamSpec], rope: RopeCtx, B: int, H: int, kmask: Tensor | None, save: dict):
    """hs[i]: [B*len_i, d] modulated inputs. Returns per-stream attention outputs BEFORE the out projection when
    w_out is None (single-stream block handles the projection itself), else after it."""
    d = hs[0].shape[-1]
    hd = d // H
    qkvs, qks, rrmss, specs = [], [], [], []
    for h, s in zip(hs, streams):
        qkv = ops.gemm(h, wb(s.w_qkv))
        qk, rrms = ops.qknorm_rope_fwd(qkv, s.sq.detach(), s.sk.detach(), rope, hd, tokens_per_sample=s.len,
                                       pos_offset=s.pos_offset, pos_idx=s.pos_idx)
        qkvs.append(qkv)
        qks.append(qk)
        rrmss.append(rrms)
        specs.append(ops.AttnSegSpec(qk, qkv, s.len))
    outs, lse = ops.attn_fwd(specs, B, H, hd, hd**-0.5, kmask)
    projs = [ops.gemm(o, wb(s.w_out)) for o, s in zip(outs, streams)]
    save.update(hs=hs, qkvs=qkvs, qks=qks, rrmss=rrmss, outs=outs, lse=lse)
    return projs


def attn_bwd(dprojs: list[Tensor | None], streams: list[StreamSpec], rope: RopeCtx, B: int, H: int, kmask: Tensor | None,
             save: dict) -> list[Tensor]:
    d = save["hs"][0].shape[-1]
    hd = d // H
    douts = []
    for dp, o, s in zip(dprojs, save["outs"], streams):
        if dp is None:  # stream whose projected output is unused (last dual block of DDT encoder / Sprint decoder)
            douts.append(torch.zeros_like(o))
            continue
        douts.append(ops.gemm(dp, wb(s.w_out), b_mn=True))
        wgrad_(s.w_out, dp, o)
    specs = [ops.AttnSegSpec(qk, qkv, s.len) for qk, qkv, s in zip(save["qks"], save["qkvs"], streams)]
    dqkvs = [torch.empty_like(q) for q in save["qkvs"]]
    dqks = ops.attn_bwd(specs, save["outs"], douts, save["lse"], B, H, hd, hd**-0.5, dqkvs, kmask)
    dhs = []
    for dqk, dqkv, qkv, rrms, h, s in zip(dqks, dqkvs, save["qkvs"], save["rrmss"], save["hs"], streams):
        ops.qknorm_rope_bwd(dqk, qkv, rrms, s.sq.detach(), s.sk.detach(), rope, hd, dqkv, vgrad(s.sq), vgrad(s.sk),
                            tokens_per_sample=s.len, pos_offset=s.pos_offset, pos_idx=s.pos_idx)
        _ready(s.sq)
        _ready(s.sk)
        dhs.append(ops.gemm(dqkv, wb(s.w_qkv), b_mn=True))
        wgrad_(s.w_qkv, dqkv, h)
    return dhs


def mlp_fwd(h: Tensor, w_up: Tensor, w_down: Tensor, save: dict) -> Tensor:
    if (w_up.shape[0] // 2) % 128 == 0:
        u, s = ops.gemm_swiglu(h, wb(w_up))  # SwiGLU in the fc1 epilogue: u is written once and never re-read in forward
    else:
        u = ops.gemm(h, wb(w_up))
        s = ops.swiglu_fwd(u)
    m = ops.gemm(s, wb(w_down))
    save.update(mlp_h=h, mlp_u=u, mlp_s=s)
    return m


def mlp_bwd(dm: Tensor, w_up: Tensor, w_down: Tensor, save: dict) -> Tensor:
    if w_down.shape[1] % 128 == 0:
        du = ops.gemm_swiglu_bwd(dm, wb(w_down), save["mlp_u"])  # d(act) never leaves the SM
    else:
        du = ops.swiglu_bwd(ops.gemm(dm, wb(w_down), b_mn=True), save["mlp_u"])
    wgrad_(w_down, dm, save["mlp_s"])
    dh = ops.gemm(du, wb(w_up), b_mn=True)
    wgrad_(w_up, du, save["mlp_h"])
    return dh


def _chunks(mod: Tensor, n: int) -> list[Tensor]:
    d = mod.shape[-1] // n
    return [mod[..., i * d : (i + 1) * d] for i in range(n)]


def _dmod_like(mod: Tensor) -> Tensor:
    """fp32 accumulator for per-sample modulation gradients; bf16 written-once buffer for per-token modulation."""
    if mod.dim() == 3 and mod.shape[1] > 1:
        return torch.empty_like(mod)
    return torch.zeros(mod.shape, device=mod.device, dtype=F32)


def _dmod_out(dmod: Tensor) -> Tensor:
    return dmod if dmod.dtype == BF16 else ops.cast_bf16(dmod)


class _Norm:
    """LayerNorm parameters of a sub-layer (either may be None for the affine-free final norm)."""

    def __init__(self, w: Tensor | None, b: Tensor | None, eps: float):
        self.w, self.b, self.eps = w, b, eps


def ln_mod_fwd(x: Tensor, n: _Norm, scale: Tensor, shift: Tensor, save: dict, key: str) -> Tensor:
    y, mean, rstd = ops.ln_modulate_fwd(x, n.w.detach() if n.w is not None else None, n.b.detach() if n.b is not None else None,
                                        scale, shift, n.eps)
    save[key] = (x, mean, rstd)
    return y


def ln_mod_bwd(dy: Tensor, n: _Norm, scale: Tensor, dres: Tensor | None, dscale: Tensor, dshift: Tensor, save: dict, key: str) -> Tensor:
    x, mean, rstd = save[key]
    dx = ops.ln_modulate_bwd(dy, x, mean, rstd, n.w.detach() if n.w is not None else None,
                             n.b.detach() if n.b is not None else None, scale, dres, dscale, dshift, vgrad(n.w), vgrad(n.b))
    if n.w is not None:
        _ready(n.w)
        _ready(n.b)
    return dx


# ---------------------------------------------------------------------------------------------------------
# DiT block (single stream, 6-way adaLN-Zero)       mmdit.py:288-309
# ---------------------------------------------------------------------------------------------------------
class DiTBlockFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x: Tensor, mod: Tensor, blk, rope: RopeCtx, pos_idx: Tensor | None, pos_offset: int, *params):
        B, N, d = x.shape
        x = x.contiguous()
        a, b, g, dl, e, z = _chunks(mod, 6)
        sv: dict = {}
        st = StreamSpec(blk.attention.qkv.weight, blk.attention.qk_norm.query_norm.scale, blk.attention.qk_norm.key_norm.scale,
                        blk.attention.proj_out.weight, N, pos_offset, pos_idx)
        n1 = _Norm(blk.norm_1.weight, blk.norm_1.bias, blk.norm_1.eps)
        n2 = _Norm(blk.norm_2.weight, blk.norm_2.bias, blk.norm_2.eps)
        h1 = ln_mod_fwd(x.view(-1, d), n1, a, b, sv, "ln1")
        (att,) = attn_fwd([h1], [st], rope, B, blk.num_heads, None, sv)
        x1 = ops.gate_residual_fwd(x.view(-1, d), att, None, g)
        h2 = ln_mod_fwd(x1, n2, dl, e, sv, "ln2")
        m = mlp_fwd(h2, blk.mlp_input[0].weight, blk.mlp_input[2].weight, sv)
        x2 = ops.gate_residual_fwd(x1, m, None, z)
        if any(ctx.needs_input_grad):
            sv.update(att=att, m=m, mod=mod)
            ctx.sv, ctx.blk, ctx.rope, ctx.st, ctx.n1, ctx.n2, ctx.shape = sv, blk, rope, st, n1, n2, (B, N, d)
        return x2.view(B, N, d)

    @staticmethod
    def backward(ctx, dx2: Tensor):
        sv, blk, rope, st, n1, n2 = ctx.sv, ctx.blk, ctx.rope, ctx.st, ctx.n1, ctx.n2
        B, N, d = ctx.shape
        mod = sv["mod"]
        a, b, g, dl, e, z = _chunks(mod, 6)
        dmod = _dmod_like(mod)
        da_, db_, dg_, ddl_, de_, dz_ = _chunks(dmod, 6)
        dx2 = dx2.contiguous().view(-1, d)
        dm = ops.gate_residual_bwd(dx2, sv["m"], None, z, dz_)
        dh2 = mlp_bwd(dm, blk.mlp_input[0].weight, blk.mlp_input[2].weight, sv)
        dx1 = ln_mod_bwd(dh2, n2, dl, dx2, ddl_, de_, sv, "ln2")
        datt = ops.gate_residual_bwd(dx1, sv["att"], None, g, dg_)
        (dh1,) = attn_bwd([datt], [st], rope, B, blk.num_heads, None, sv)
        dx = ln_mod_bwd(dh1, n1, a, dx1, da_, db_, sv, "ln1")
        ctx.sv = None
        return (dx.view(B, N, d), _dmod_out(dmod), None, None, None, None) + (None,) * (len(ctx.needs_input_grad) - 6)


# ---------------------------------------------------------------------------------------------------------
# MMDiT dual-stream block       mmdit.py:416-459
# ---------------------------------------------------------------------------------------------------------
class MMDiTBlockFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x: Tensor, c: Tensor, mod_x: Tensor, mod_c: Tensor, blk, rope: RopeCtx, kmask: Tensor | None,
                pos_idx: Tensor | None, *params):
        ctx.set_materialize_grads(False)
        B, N, d = x.shape
        L = c.shape[1]
        x, c = x.contiguous().view(-1, d), c.contiguous().view(-1, d)
        mx, mc = _chunks(mod_x, 6), _chunks(mod_c, 6)
        at = blk.attention
        sc = StreamSpec(at.qkv_context.weight, at.qk_norm_context.query_norm.scale, at.qk_norm_context.key_norm.scale,
                        at.context_proj_out.weight, L, 0, None)
        sx = StreamSpec(at.qkv_input.weight, at.qk_norm_input.query_norm.scale, at.qk_norm_input.key_norm.scale,
                        at.input_proj_out.weight, N, L, pos_idx)
        nx1 = _Norm(blk.input_norm_1.weight, blk.input_norm_1.bias, blk.input_norm_1.eps)
        nc1 = _Norm(blk.context_norm_1.weight, blk.context_norm_1.bias, blk.context_norm_1.eps)
        nx2 = _Norm(blk.input_norm_2.weight, blk.input_norm_2.bias, blk.input_norm_2.eps)
        nc2 = _Norm(blk.context_norm_2.weight, blk.context_norm_2.bias, blk.context_norm_2.eps)
        sv: dict = {}
        svx: dict = {}
        svc: dict = {}
        hx = ln_mod_fwd(x, nx1, mx[0], mx[1], sv, "lnx1")
        hc = ln_mod_fwd(c, nc1, mc[0], mc[1], sv, "lnc1")
        ac, ax = attn_fwd([hc, hx], [sc, sx], rope, B, blk.num_heads, kmask, sv)
        x1 = ops.gate_residual_fwd(x, ax, None, mx[2])
        c1 = ops.gate_residual_fwd(c, ac, None, mc[2])
        hx2 = ln_mod_fwd(x1, nx2, mx[3], mx[4], sv, "lnx2")
        m_x = mlp_fwd(hx2, blk.mlp_input[0].weight, blk.mlp_input[2].weight, svx)
        x2 = ops.gate_residual_fwd(x1, m_x, None, mx[5])
        hc2 = ln_mod_fwd(c1, nc2, mc[3], mc[4], sv, "lnc2")
        m_c = mlp_fwd(hc2, blk.mlp_context[0].weight, blk.mlp_context[2].weight, svc)
        c2 = ops.gate_residual_fwd(c1, m_c, None, mc[5])
        if any(ctx.needs_input_grad):
            sv.update(ax=ax, ac=ac, m_x=m_x, m_c=m_c, mod_x=mod_x, mod_c=mod_c, svx=svx, svc=svc)
            ctx.sv, ctx.blk, ctx.rope, ctx.kmask = sv, blk, rope, kmask
            ctx.streams, ctx.norms, ctx.shape = (sc, sx), (nx1, nc1, nx2, nc2), (B, N, L, d)
        return x2.view(B, N, d), c2.view(B, L, d)

    @staticmethod
    def backward(ctx, dx2: Tensor | None, dc2: Tensor | None):
        sv, blk, rope = ctx.sv, ctx.blk, ctx.rope
        (sc, sx), (nx1, nc1, nx2, nc2) = ctx.streams, ctx.norms
        B, N, L, d = ctx.shape
        mod_x, mod_c = sv["mod_x"], sv["mod_c"]
        mx, mc = _chunks(mod_x, 6), _chunks(mod_c, 6)
        dmod_x, dmod_c = _dmod_like(mod_x), _dmod_like(mod_c)
        dmx, dmc = _chunks(dmod_x, 6), _chunks(dmod_c, 6)
        # image stream MLP sub-layer
        dx2 = dx2.contiguous().view(-1, d)
        dm = ops.gate_residual_bwd(dx2, sv["m_x"], None, mx[5], dmx[5])
        dh = mlp_bwd(dm, blk.mlp_input[0].weight, blk.mlp_input[2].weight, sv["svx"])
        dx1 = ln_mod_bwd(dh, nx2, mx[3], dx2, dmx[3], dmx[4], sv, "lnx2")
        # text stream MLP sub-layer (its output is unused in the last dual block: gradient is None)
        if dc2 is not None:
            dc2 = dc2.contiguous().view(-1, d)
            dmc_ = ops.gate_residual_bwd(dc2, sv["m_c"], None, mc[5], dmc[5])
            dhc = mlp_bwd(dmc_, blk.mlp_context[0].weight, blk.mlp_context[2].weight, sv["svc"])
            dc1 = ln_mod_bwd(dhc, nc2, mc[3], dc2, dmc[3], dmc[4], sv, "lnc2")
            dac = ops.gate_residual_bwd(dc1, sv["ac"], None, mc[2], dmc[2])
        else:
            dc1, dac = None, None
        dax = ops.gate_residual_bwd(dx1, sv["ax"], None, mx[2], dmx[2])
        dhc1, dhx1 = attn_bwd([dac, dax], [sc, sx], rope, B, blk.num_heads, ctx.kmask, sv)
        dx = ln_mod_bwd(dhx1, nx1, mx[0], dx1, dmx[0], dmx[1], sv, "lnx1")
        dc = ln_mod_bwd(dhc1, nc1, mc[0], dc1, dmc[0], dmc[1], sv, "lnc1")
        ctx.sv = None
        return (dx.view(B, N, d), dc.view(B, L, d), _dmod_out(dmod_x), _dmod_out(dmod_c), None, None, None, None) + (None,) * (
            len(ctx.needs_input_grad) - 8
        )


# ---------------------------------------------------------------------------------------------------------
# MMDiT single-stream block (parallel attention + MLP on cat[text, image])      mmdit.py:499-532
# ---------------------------------------------------------------------------------------------------------
class SingleStreamBlockFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z: Tensor, mod: Tensor, blk, rope: RopeCtx, kmask: Tensor | None, pos_idx: Tensor | None, *params):
        B, S, d = z.shape
        z = z.contiguous().view(-1, d)
        a, b, g = _chunks(mod, 3)
        at = blk.attention
        st = StreamSpec(at.qkv.weight, at.qk_norm.query_norm.scale, at.qk_norm.key_norm.scale, at.proj_out.weight, S, 0, pos_idx)
        n = _Norm(blk.norm.weight, blk.norm.bias, blk.norm.eps)
        sv: dict = {}
        h = ln_mod_fwd(z, n, a, b, sv, "ln")
        (att,) = attn_fwd([h], [st], rope, B, blk.num_heads, kmask, sv)
        m = mlp_fwd(h, blk.mlp[0].weight, blk.mlp[2].weight, sv)
        z2 = ops.gate_residual_fwd(z, att, m, g)
        if any(ctx.needs_input_grad):
            sv.update(att=att, m=m, mod=mod)
            ctx.sv, ctx.blk, ctx.rope, ctx.kmask, ctx.st, ctx.n, ctx.shape = sv, blk, rope, kmask, st, n, (B, S, d)
        return z2.view(B, S, d)

    @staticmethod
    def backward(ctx, dz2: Tensor):
        sv, blk, rope, st, n = ctx.sv, ctx.blk, ctx.rope, ctx.st, ctx.n
        B, S, d = ctx.shape
        mod = sv["mod"]
        a, b, g = _chunks(mod, 3)
        dmod = _dmod_like(mod)
        da_, db_, dg_ = _chunks(dmod, 3)
        dz2 = dz2.contiguous().view(-1, d)
        dbr = ops.gate_residual_bwd(dz2, sv["att"], sv["m"], g, dg_)  # grad of (attn + mlp)
        dh_m = mlp_bwd(dbr, blk.mlp[0].weight, blk.mlp[2].weight, sv)
        (dh_a,) = attn_bwd([dbr], [st], rope, B, blk.num_heads, ctx.kmask, sv)
        dh = ops.add(dh_a, dh_m)  # two branches read the same modulated input
        dz = ln_mod_bwd(dh, n, a, dz2, da_, db_, sv, "ln")
        ctx.sv = None
        return (dz.view(B, S, d), _dmod_out(dmod), None, None, None, None) + (None,) * (len(ctx.needs_input_grad) - 6)


# ---------------------------------------------------------------------------------------------------------
# final layer: LN (no affine) + 2-way modulate + Linear(d -> p*p*C) + unpatchify      mmdit.py:535-549, 767-787
# ---------------------------------------------------------------------------------------------------------
class FinalLayerFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x: Tensor, mod: Tensor, w: Tensor, bias: Tensor, geom: tuple):
        B, C, Himg, Wimg, p = geom
        N, d = x.shape[1], x.shape[2]
        x = x.contiguous().view(-1, d)
        a, b = _chunks(mod, 2)
        sv: dict = {}
        n = _Norm(None, None, 1e-6)
        h = ln_mod_fwd(x, n, a, b, sv, "ln")
        ppc = p * p * C
        ld = (ppc + 7) // 8 * 8
        out = torch.empty(x.shape[0], ld, device=x.device, dtype=BF16)
        ops.gemm(h, wb(w), bias=bias.detach(), out=out[:, :ppc])
        img = ops.unpatchify(out, B, C, Himg, Wimg, p, BF16)
        if any(ctx.needs_input_grad):
            sv.update(h=h, mod=mod)
            ctx.sv, ctx.w, ctx.bias, ctx.n, ctx.geom, ctx.shape, ctx.ld = sv, w, bias, n, geom, (B, N, d), ld
        return img

    @staticmethod
    def backward(ctx, dimg: Tensor):
        sv, w, bias, n = ctx.sv, ctx.w, ctx.bias, ctx.n
        B, C, Himg, Wimg, p = ctx.geom
        _, N, d = ctx.shape
        ppc = p * p * C
        mod = sv["mod"]
        a, b = _chunks(mod, 2)
        dmod = _dmod_like(mod)
        da_, db_ = _chunks(dmod, 2)
        dout = ops.patchify_grad(dimg.contiguous(), p, ld=ctx.ld)[:, :ppc]
        dh = ops.gemm(dout, wb(w), b_mn=True)
        wgrad_(w, dout, sv["h"])
        bgrad_(bias, dout)
        dx = ln_mod_bwd(dh, n, a, None, da_, db_, sv, "ln")
        ctx.sv = None
        return dx.view(B, N, d), _dmod_out(dmod), None, None, None


# ---------------------------------------------------------------------------------------------------------
# SPRINT token drop / restore       sprint.py:317-387
# ---------------------------------------------------------------------------------------------------------
class GatherTokensFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x: Tensor, kept: Tensor, inv: Tensor):
        ctx.save_for_backward(inv)
        return ops.gather_rows(x.contiguous(), kept)

    @staticmethod
    def backward(ctx, dxk: Tensor):
        (inv,) = ctx.saved_tensors
        return ops.restore_rows(dxk.contiguous(), inv, None, None), None, None


class RestoreTokensFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xk: Tensor, mask_token: Tensor, kept: Tensor, inv: Tensor, drop: Tensor | None):
        ctx.save_for_backward(kept, inv)
        ctx.drop, ctx.mask_token = drop, mask_token
        fill = ops.cast_f32(ops.cast_bf16(mask_token.detach().reshape(-1).contiguous()))  # mask_token.to(bf16) as in the reference
        return ops.restore_rows(xk.contiguous(), inv, fill, drop)

    @staticmethod
    def backward(ctx, dy: Tensor):
        kept, inv = ctx.saved_tensors
        mt = ctx.mask_token
        dfill = gbuf(mt).view(-1) if mt.requires_grad else None
        dxk = ops.restore_rows_bwd(dy.contiguous(), kept, inv, ctx.drop, dfill)
        if mt.requires_grad:
            _ready(mt)
        return dxk, None, None, None, None


class MaskFillFn(torch.autograd.Function):
    """x_restored = mask_token.expand(B,S,d) (p >= 1: deep path skipped, sprint.py:474-475)."""

    @staticmethod
    def forward(ctx, mask_token: Tensor, B: int, S: int):
        ctx.mask_token = mask_token
        d = mask_token.numel()
        fill = ops.cast_f32(ops.cast_bf16(mask_token.detach().reshape(-1).contiguous()))
        inv = torch.full((B, S), -1, device=mask_token.device, dtype=torch.int32)
        dummy = torch.empty(B, 1, d, device=mask_token.device, dtype=BF16)
        return ops.restore_rows(dummy, inv, fill, None)

    @staticmethod
    def backward(ctx, dy: Tensor):
        mt = ctx.mask_token
        if mt.requires_grad:
            ops.colsum_(dy.reshape(-1, dy.shape[-1]), gbuf(mt).view(-1))
            _ready(mt)
        return None, None, None


# ---------------------------------------------------------------------------------------------------------
# SiLU on the token stream (DDT decoder conditioning silu(enc + emb), ddt.py:422)
# ---------------------------------------------------------------------------------------------------------
class SiluFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x: Tensor):
        x = x.contiguous()
        ctx.save_for_backward(x)
        return ops.silu_fwd(x)

    @staticmethod
    def backward(ctx, dy: Tensor):
        (x,) = ctx.saved_tensors
        return ops.silu_bwd(dy.contiguous(), x, x.dtype)


class BiasSiluFn(torch.autograd.Function):
    """silu(x + v[:, None, :]) with x [B,N,d] and v [B,d] (DDT decoder conditioning, ddt.py:421-422)."""

    @staticmethod
    def forward(ctx, x: Tensor, v: Tensor):
        x, v = x.contiguous(), v.contiguous()
        ctx.save_for_backward(x, v)
        return ops.bias_silu_fwd(x, v)

    @staticmethod
    def backward(ctx, dy: Tensor):
        x, v = ctx.saved_tensors
        dx, dv = ops.bias_silu_bwd(dy.contiguous(), x, v)
        return dx, ops.cast_bf16(dv)


def block_params(blk: nn.Module) -> list[Tensor]:
    return list(blk.parameters())
