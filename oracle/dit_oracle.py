"""TEST INFRASTRUCTURE ONLY — CPU oracle for the denoiser training / sampling hot path.

A plain, functional, fp32 restatement (torch CPU tensor arithmetic only: matmul, elementwise, softmax) of the
reference algorithm, driven by a `state_dict` with the reference's own parameter names.
It exists to CHECK the CUDA path; only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` leg may import it. The product package (`diffulab_b200/`) never does.

Pinning: the reference ships no tests or golden vectors of its own ("parity unpinned by the reference",
SURVEY.md 4.2). The oracle is therefore pinned against OUTPUTS OF THE REFERENCE ITSELF, generated in the build
container by `oracle/make_golden.py` (imports the unmodified reference through `oracle/ref_shim.py`) and committed
under `tests/golden/`; `tests/test_oracle_golden.py` replays them, `tests/test_oracle_vs_reference.py` compares
live where /root/reference is mounted.

Every function cites the reference file:line it restates (paths relative to /root/reference/src/diffulab).
`set_round(fn)` optionally inserts bf16 round-trips at the points where CUDA bf16 autocast rounds in the
reference (Linear / SDPA outputs, elementwise results on bf16 tensors), to tighten GPU comparisons.
"""

from __future__ import annotations

import math
from typing import Any, Callable

import torch
import torch.nn.functional as F
from torch import Tensor

SD = dict[str, Tensor]

_round: Callable[[Tensor], Tensor] = lambda x: x  # noqa: E731


def set_round(mode: str | None) -> None:
    """mode None/'fp32': exact fp32 restatement. 'bf16': emulate autocast rounding points."""
    global _round
    if mode in (None, "fp32"):
        _round = lambda x: x  # noqa: E731
    elif mode == "bf16":
        _round = lambda x: x.to(torch.bfloat16).to(torch.float32)  # noqa: E731
    else:
        raise ValueError(mode)


def r(x: Tensor) -> Tensor:
    return _round(x)


def linear(x: Tensor, w: Tensor, b: Tensor | None = None) -> Tensor:
    """nn.Linear under autocast: operands rounded to bf16, fp32 accumulate, bf16 result."""
    return r(F.linear(r(x), r(w), r(b) if b is not None else None))


# ---------------------------------------------------------------------------------------------------------
# conditioning
# ---------------------------------------------------------------------------------------------------------
def timestep_embedding(t: Tensor, dim: int, max_period: int = 10000) -> Tensor:
    """networks/utils/nn.py:91-114"""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(half, dtype=torch.float32, device=t.device) / half)
    args = t[:, None].float() * freqs[None]
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    if dim % 2:
        emb = torch.cat([emb, torch.zeros_like(emb[:, :1])], dim=-1)
    return emb


def time_embed(sd: SD, t: Tensor, freq_dim: int, prefix: str = "time_embed") -> Tensor:
    """denoisers/mmdit.py:691-695, 866"""
    h = linear(timestep_embedding(t, freq_dim), sd[f"{prefix}.0.weight"], sd[f"{prefix}.0.bias"])
    h = r(F.silu(h))
    return linear(h, sd[f"{prefix}.2.weight"], sd[f"{prefix}.2.bias"])


def label_embed(sd: SD, y: Tensor, n_classes: int, p: float, drop_u: Tensor | None) -> Tensor:
    """nn.py:135-164: labels replaced by the extra class where rand < p"""
    if p > 0:
        assert drop_u is not None, "label dropout needs the uniform draw"
        y = torch.where(drop_u.to(y.device) < p, torch.full_like(y, n_classes), y)
    return sd["label_embed.embedding.weight"][y]


def rope_tables(pos_ids: Tensor, axes_dim: list[int], base: float) -> tuple[Tensor, Tensor]:
    """nn.py:262-307 (fp64 angles -> fp32 tables). pos_ids [S, n_axes] -> cos, sin [S, sum(axes)/2]"""
    cs, sn = [], []
    for i, ad in enumerate(axes_dim):
        pos = pos_ids[..., i].to(torch.float64)
        freqs = 1.0 / (base ** (torch.arange(0, ad, 2, dtype=torch.float64, device=pos_ids.device) / ad))
        ang = pos[..., None] * freqs
        cs.append(ang.cos().float())
        sn.append(ang.sin().float())
    return torch.cat(cs, -1), torch.cat(sn, -1)


def pos_ids_2d(hp: int, wp: int) -> Tensor:
    """mmdit.py:871-886"""
    return torch.stack(torch.meshgrid(torch.arange(hp), torch.arange(wp), indexing="ij"), -1).view(-1, 2)


def pos_ids_joint(L: int, hp: int, wp: int) -> Tensor:
    """mmdit.py:815-835: text (l,0,0), l = 1..L, then image (0,h,w)"""
    text = torch.stack([torch.arange(1, L + 1), torch.zeros(L, dtype=torch.long), torch.zeros(L, dtype=torch.long)], -1)
    img = torch.stack(torch.meshgrid(torch.zeros(1, dtype=torch.long), torch.arange(hp), torch.arange(wp), indexing="ij"), -1).view(-1, 3)
    return torch.cat([text, img], 0)


# ---------------------------------------------------------------------------------------------------------
# block pieces
# ---------------------------------------------------------------------------------------------------------
def modulate(x: Tensor, scale: Tensor, shift: Tensor) -> Tensor:
    """nn.py:539-540; `1 + scale` is a bf16 op in the reference when scale is bf16"""
    return x * r(1 + scale) + shift


def layer_norm(x: Tensor, w: Tensor | None, b: Tensor | None, eps: float) -> Tensor:
    return F.layer_norm(x, (x.shape[-1],), w, b, eps)


def modulation(sd: SD, prefix: str, vec: Tensor, n: int) -> list[Tensor]:
    """nn.py:526-536 (Modulation, prefix.lin) / the nn.Sequential(SiLU, Linear) variants (prefix.1)"""
    key = f"{prefix}.lin" if f"{prefix}.lin.weight" in sd else f"{prefix}.1"
    out = linear(F.silu(vec), sd[f"{key}.weight"], sd[f"{key}.bias"])
    if out.dim() == 2:
        out = out[:, None, :]
    return list(out.chunk(n, dim=-1))


def rms_norm(x: Tensor, scale: Tensor) -> Tensor:
    """nn.py:427-431, 473-475"""
    xf = x.float()  # RMSNorm.forward upcasts, casts back to the input dtype, then multiplies by the fp32 scale
    rr = torch.rsqrt(torch.mean(xf * xf, dim=-1, keepdim=True) + 1e-6)
    return r(r((xf * rr).to(x.dtype)) * scale)


def apply_rope(x: Tensor, cos: Tensor, sin: Tensor) -> Tensor:
    """nn.py:331-353, 377-400. x [B,S,H,hd]; cos/sin [S,R/2] or [B,S,R/2] (cast to the activation dtype)."""
    R = cos.shape[-1] * 2
    cos, sin = cos.to(x.dtype), sin.to(x.dtype)  # nn.py:377-378
    c = r(cos)[..., None, :] if cos.dim() == 3 else r(cos)[None, :, None, :]
    s = r(sin)[..., None, :] if sin.dim() == 3 else r(sin)[None, :, None, :]
    xr, xp = x[..., :R], x[..., R:]
    e, o = xr[..., 0::2], xr[..., 1::2]
    re = r(r(e * c) - r(o * s))
    ro = r(r(e * s) + r(o * c))
    rot = torch.stack([re, ro], -1).flatten(-2)
    return torch.cat([rot, xp], -1)


_fused_sdpa = False


def set_fused_sdpa(on: bool) -> None:
    """True: call F.scaled_dot_product_attention exactly as the reference does (mmdit.py:92-98) instead of the explicit
    softmax(QK^T)V restatement — used by bench.py's reference-GPU arm so that the timed attention is the library kernel
    the reference would run (flash / cuDNN / math backend chosen by torch)."""
    global _fused_sdpa
    _fused_sdpa = bool(on)


def sdpa(q: Tensor, k: Tensor, v: Tensor, key_mask: Tensor | None) -> Tensor:
    """F.scaled_dot_product_attention(scale=hd^-0.5, attn_mask=key padding) mmdit.py:92-98. q,k,v [B,S,H,hd]"""
    hd = q.shape[-1]
    if _fused_sdpa:
        mask = key_mask[:, None, None, :].bool() if key_mask is not None else None
        o = F.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2), attn_mask=mask, scale=hd**-0.5)
        return o.transpose(1, 2)
    s = torch.einsum("bqhd,bkhd->bhqk", q, k) * hd**-0.5
    if key_mask is not None:
        s = s.masked_fill(~key_mask[:, None, None, :].bool(), float("-inf"))
    p = r(s.softmax(-1))
    return r(torch.einsum("bhqk,bkhd->bqhd", p, v))


def qkv_heads(sd: SD, wkey: str, nkey: str, x: Tensor, H: int) -> tuple[Tensor, Tensor, Tensor]:
    B, S, d = x.shape
    q, k, v = linear(x, sd[f"{wkey}.weight"]).chunk(3, dim=-1)
    q = rms_norm(q, sd[f"{nkey}.query_norm.scale"]).to(v.dtype)  # QKNorm.forward: q.to(v), k.to(v) (nn.py:475)
    k = rms_norm(k, sd[f"{nkey}.key_norm.scale"]).to(v.dtype)
    return q.view(B, S, H, -1), k.view(B, S, H, -1), v.view(B, S, H, -1)


def dit_attention(sd: SD, prefix: str, x: Tensor, cos: Tensor, sin: Tensor, H: int, key_mask: Tensor | None = None) -> Tensor:
    """DiTAttention.forward mmdit.py:75-104"""
    B, S, d = x.shape
    q, k, v = qkv_heads(sd, f"{prefix}.qkv", f"{prefix}.qk_norm", x, H)
    o = sdpa(apply_rope(q, cos, sin), apply_rope(k, cos, sin), v, key_mask)
    return linear(o.reshape(B, S, d), sd[f"{prefix}.proj_out.weight"])


def mmdit_attention(sd: SD, prefix: str, x: Tensor, ctx: Tensor, cos: Tensor, sin: Tensor, H: int, ctx_mask: Tensor | None):
    """MMDiTAttention.forward mmdit.py:171-210 (text rows first)"""
    B, N, d = x.shape
    L = ctx.shape[1]
    qi, ki, vi = qkv_heads(sd, f"{prefix}.qkv_input", f"{prefix}.qk_norm_input", x, H)
    qc, kc, vc = qkv_heads(sd, f"{prefix}.qkv_context", f"{prefix}.qk_norm_context", ctx, H)
    q, k, v = torch.cat([qc, qi], 1), torch.cat([kc, ki], 1), torch.cat([vc, vi], 1)
    mask = None
    if ctx_mask is not None:
        mask = torch.cat([ctx_mask.bool(), torch.ones(B, N, dtype=torch.bool, device=ctx_mask.device)], 1)
    o = sdpa(apply_rope(q, cos, sin), apply_rope(k, cos, sin), v, mask).reshape(B, L + N, d)
    return linear(o[:, L:], sd[f"{prefix}.input_proj_out.weight"]), linear(o[:, :L], sd[f"{prefix}.context_proj_out.weight"])


def swiglu_mlp(sd: SD, prefix: str, x: Tensor) -> Tensor:
    """nn.Sequential(Linear(d,8d), PackedSwiGLU, Linear(4d,d)) mmdit.py:260-264 ; nn.py:485-486"""
    h = linear(x, sd[f"{prefix}.0.weight"])
    a, g = h.chunk(2, dim=-1)
    return linear(r(r(F.silu(a)) * g), sd[f"{prefix}.2.weight"])


def gate_res(x: Tensor, branch: Tensor, gate: Tensor) -> Tensor:
    return r(x + r(branch * gate))


def dit_block(sd: SD, p: str, x: Tensor, cond: Tensor, cos: Tensor, sin: Tensor, H: int) -> Tensor:
    """DiTBlock._forward mmdit.py:288-309"""
    a, b, g, d_, e, z = modulation(sd, f"{p}.modulation", cond, 6)
    h = modulate(layer_norm(x, sd[f"{p}.norm_1.weight"], sd[f"{p}.norm_1.bias"], 1e-5), a, b)
    x = gate_res(x, dit_attention(sd, f"{p}.attention", h, cos, sin, H), g)
    h = modulate(layer_norm(x, sd[f"{p}.norm_2.weight"], sd[f"{p}.norm_2.bias"], 1e-5), d_, e)
    return gate_res(x, swiglu_mlp(sd, f"{p}.mlp_input", h), z)


def mmdit_block(sd: SD, p: str, x: Tensor, cond: Tensor, ctx: Tensor, cos: Tensor, sin: Tensor, H: int, ctx_mask):
    """MMDiTBlock._forward mmdit.py:416-459"""
    mi = modulation(sd, f"{p}.modulation_input", cond, 6)
    mc = modulation(sd, f"{p}.modulation_context", cond, 6)
    hx = modulate(layer_norm(x, sd[f"{p}.input_norm_1.weight"], sd[f"{p}.input_norm_1.bias"], 1e-5), mi[0], mi[1])
    hc = modulate(layer_norm(ctx, sd[f"{p}.context_norm_1.weight"], sd[f"{p}.context_norm_1.bias"], 1e-5), mc[0], mc[1])
    ax, ac = mmdit_attention(sd, f"{p}.attention", hx, hc, cos, sin, H, ctx_mask)
    x = gate_res(x, ax, mi[2])
    ctx = gate_res(ctx, ac, mc[2])
    hx = modulate(layer_norm(x, sd[f"{p}.input_norm_2.weight"], sd[f"{p}.input_norm_2.bias"], 1e-5), mi[3], mi[4])
    x = gate_res(x, swiglu_mlp(sd, f"{p}.mlp_input", hx), mi[5])
    hc = modulate(layer_norm(ctx, sd[f"{p}.context_norm_2.weight"], sd[f"{p}.context_norm_2.bias"], 1e-5), mc[3], mc[4])
    ctx = gate_res(ctx, swiglu_mlp(sd, f"{p}.mlp_context", hc), mc[5])
    return x, ctx


def single_stream_block(sd: SD, p: str, x: Tensor, cond: Tensor, ctx: Tensor, cos: Tensor, sin: Tensor, H: int, ctx_mask):
    """MMDiTSingleStreamBlock._forward mmdit.py:499-532"""
    L = ctx.shape[1]
    z = torch.cat([ctx, x], 1)
    mask = None
    if ctx_mask is not None:
        mask = torch.cat([ctx_mask.bool(), torch.ones(x.shape[0], x.shape[1], dtype=torch.bool, device=x.device)], 1)
    a, b, g = modulation(sd, f"{p}.modulation", cond, 3)
    h = modulate(layer_norm(z, sd[f"{p}.norm.weight"], sd[f"{p}.norm.bias"], 1e-5), a, b)
    branch = r(dit_attention(sd, f"{p}.attention", h, cos, sin, H, mask) + swiglu_mlp(sd, f"{p}.mlp", h))
    z = gate_res(z, branch, g)
    return z[:, L:], z[:, :L]


def last_layer(sd: SD, x: Tensor, cond: Tensor) -> Tensor:
    """ModulatedLastLayer.forward mmdit.py:542-549"""
    a, b = modulation(sd, "last_layer.adaLN_modulation", cond, 2)
    h = modulate(layer_norm(x, None, None, 1e-6), a, b)
    return linear(h, sd["last_layer.linear.weight"], sd["last_layer.linear.bias"])


def patchify(x: Tensor, w: Tensor) -> tuple[Tensor, int, int]:
    """Conv2d(k=s=p, bias=False) + 'b c h w -> b (h w) c' (mmdit.py:697-699, 757-765) written as an explicit matmul"""
    B, C, H, W = x.shape
    p = w.shape[-1]
    hp, wp = H // p, W // p
    patches = x.view(B, C, hp, p, wp, p).permute(0, 2, 4, 1, 3, 5).reshape(B, hp * wp, C * p * p)
    return linear(patches, w.reshape(w.shape[0], -1)), hp, wp


def unpatchify(x: Tensor, hp: int, wp: int, p: int, c: int) -> Tensor:
    """'b (h w) (p1 p2 c) -> b c (h p1) (w p2)' mmdit.py:778-786"""
    B = x.shape[0]
    return x.view(B, hp, wp, p, p, c).permute(0, 5, 1, 3, 2, 4).reshape(B, c, hp * p, wp * p)


def block_kind(sd: SD, prefix: str) -> str:
    if f"{prefix}.modulation_input.lin.weight" in sd:
        return "mmdit"
    if f"{prefix}.modulation.lin.weight" in sd:
        return "dit"
    if f"{prefix}.modulation.1.weight" in sd:
        return "single"
    raise KeyError(prefix)


def n_blocks(sd: SD, list_name: str) -> int:
    idx = {int(k[len(list_name) + 1 :].split(".")[0]) for k in sd if k.startswith(list_name + ".")}
    return max(idx) + 1 if idx else 0


def run_block(sd, prefix, x, cond, ctx, cos, sin, H, ctx_mask):
    kind = block_kind(sd, prefix)
    if kind == "dit":
        return dit_block(sd, prefix, x, cond, cos, sin, H), ctx
    if kind == "mmdit":
        return mmdit_block(sd, prefix, x, cond, ctx, cos, sin, H, ctx_mask)
    return single_stream_block(sd, prefix, x, cond, ctx, cos, sin, H, ctx_mask)


def drop_context(context: dict[str, Tensor], null_emb: Tensor, null_mask: Tensor, p: float, drop_u: Tensor | None):
    """PrecomputedEmbedder.drop_conditions embedders/precomputed.py:20-39"""
    emb, mask = context["embeddings"], context["attn_mask"]
    if drop_u is None:
        drop_u = torch.ones(emb.shape[0], device=emb.device)  # rand < 0 never true; p == 0 path
    dm = drop_u.to(emb.device) < p
    emb = torch.where(dm[:, None, None], null_emb.to(emb.device)[None].expand_as(emb), emb)
    mask = torch.where(dm[:, None], null_mask.to(mask.device)[None].expand_as(mask), mask)
    return emb, mask


# ---------------------------------------------------------------------------------------------------------
# denoisers
# ---------------------------------------------------------------------------------------------------------
def mmdit_forward(sd: SD, cfg: dict[str, Any], x: Tensor, t: Tensor, y: Tensor | None = None, context: dict | None = None,
                  p: float = 0.0, draws: dict[str, Tensor] | None = None, capture: dict | None = None) -> Tensor:
    """MMDiT.forward mmdit.py:903-928 (+ simple_dit_forward :853-901, mmdit_forward :789-851).
    cfg: num_heads, patch_size, output_channels, rope_axes_dim, rope_base, frequency_embedding, n_classes,
    null_embedding/null_mask (MM mode). draws: explicit uniform draws replacing torch.rand ('label', 'context')."""
    draws = draws or {}
    H, ps = cfg["num_heads"], cfg["patch_size"]
    tok, hp, wp = patchify(x, sd["conv_proj.weight"])
    emb = time_embed(sd, t, cfg.get("frequency_embedding", 256))
    ctx, ctx_mask = None, None
    if context is None:
        if "label_embed.embedding.weight" in sd:
            emb = emb + label_embed(sd, y, cfg["n_classes"], p, draws.get("label"))
        cos, sin = rope_tables(pos_ids_2d(hp, wp).to(x.device), cfg["rope_axes_dim"], cfg.get("rope_base", 10000))
    else:
        ce, ctx_mask = drop_context(context, cfg["null_embedding"], cfg["null_mask"], p, draws.get("context"))
        ctx = linear(ce, sd["context_embed.weight"])
        cos, sin = rope_tables(pos_ids_joint(ctx.shape[1], hp, wp).to(x.device), cfg["rope_axes_dim"], cfg.get("rope_base", 10000))
    for i in range(n_blocks(sd, "layers")):
        tok, ctx = run_block(sd, f"layers.{i}", tok, emb, ctx, cos, sin, H, ctx_mask)
        if capture is not None:
            capture[f"layers.{i}"] = tok
    out = last_layer(sd, tok, emb)
    return unpatchify(out, hp, wp, ps, cfg["output_channels"])


def sprint_select(scores: Tensor, k: int) -> Tensor:
    """sprint.py:343-346: indices of the k largest scores, ascending. Plain loops, tie -> larger index
    (the build's documented tie rule; exactness vs torch.topk is defined on tie-free draws)."""
    B, S = scores.shape
    if scores.is_cuda:  # the reference's own device expression (bench timing arm only; tie order is implementation-defined)
        return torch.topk(scores, k=k, dim=1, largest=True, sorted=False).indices.sort(dim=1).values
    out = torch.empty(B, k, dtype=torch.long)
    for b in range(B):
        row = scores[b].tolist()
        order = sorted(range(S), key=lambda i: (row[i], i), reverse=True)[:k]
        out[b] = torch.tensor(sorted(order), dtype=torch.long)
    return out


def sprint_restore(xk: Tensor, kept: Tensor, S: int, mask_token: Tensor, path_drop: Tensor | None) -> Tensor:
    """sprint.py:371-387"""
    B, k, d = xk.shape
    full = r(mask_token).to(xk.dtype).reshape(1, 1, d).expand(B, S, d).clone()
    full = full.scatter(1, kept.unsqueeze(-1).expand(-1, -1, d), xk)
    if path_drop is not None:
        full = torch.where(path_drop.to(xk.device).bool()[:, None, None], r(mask_token).to(xk.dtype).reshape(1, 1, d).expand(B, S, d), full)
    return full


def sprint_forward(sd: SD, cfg: dict[str, Any], x: Tensor, t: Tensor, y: Tensor | None = None, context: dict | None = None,
                   p: float = 0.0, training: bool = True, draws: dict[str, Tensor] | None = None, capture: dict | None = None) -> Tensor:
    """SprintDiT._forward_mmdit / _forward_dit sprint.py:389-585. draws: 'label'|'context', 'scores' [B,S], 'path' [B]."""
    draws = draws or {}
    H, ps = cfg["num_heads"], cfg["patch_size"]
    tok, hp, wp = patchify(x, sd["conv_proj.weight"])
    B, S, d = tok.shape
    emb = time_embed(sd, t, cfg.get("frequency_embedding", 256))
    ctx, ctx_mask, L = None, None, 0
    if context is None:
        if "label_embed.embedding.weight" in sd:
            emb = emb + label_embed(sd, y, cfg["n_classes"], p, draws.get("label"))
        pos = pos_ids_2d(hp, wp)
    else:
        ce, ctx_mask = drop_context(context, cfg["null_embedding"], cfg["null_mask"], p, draws.get("context"))
        ctx = linear(ce, sd["context_embed.weight"])
        L = ctx.shape[1]
        pos = pos_ids_joint(L, hp, wp)
    cos, sin = rope_tables(pos.to(x.device), cfg["rope_axes_dim"], cfg.get("rope_base", 10000))
    for i in range(n_blocks(sd, "layers")):
        tok, ctx = run_block(sd, f"layers.{i}", tok, emb, ctx, cos, sin, H, ctx_mask)
        if capture is not None:
            capture[f"layers.{i}"] = tok
    enc_ctx = ctx
    if training:
        k = max(1, int(S * (1.0 - float(cfg["drop_rate"]))))
        sc = draws["scores"] if cfg.get("device_select") else draws["scores"].cpu()  # parity: documented tie rule on the host
        kept = sprint_select(sc, k).to(x.device)
    else:
        kept = torch.arange(S, device=x.device).expand(B, S)
    if capture is not None:
        capture["kept_indices"] = kept
    xk = torch.gather(tok, 1, kept.unsqueeze(-1).expand(-1, -1, d))
    # per-sample RoPE rows: text rows, then the kept image rows (sprint.py:460-465)
    cos_b = torch.cat([cos[:L].expand(B, L, -1), cos[L:][kept]], 1)
    sin_b = torch.cat([sin[:L].expand(B, L, -1), sin[L:][kept]], 1)
    if p < 1:
        for i in range(n_blocks(sd, "deep_layers")):
            xk, ctx = run_block(sd, f"deep_layers.{i}", xk, emb, ctx, cos_b, sin_b, H, ctx_mask)
        path = (draws["path"].cpu() < p) if p > 0 else None
        restored = sprint_restore(xk, kept, S, sd["mask_token"], path)
    else:
        restored = r(sd["mask_token"]).expand(B, S, d).clone()
    fused = linear(torch.cat([restored, tok], -1), sd["fuse.weight"])
    if ctx is not None:
        ctx = linear(torch.cat([ctx, enc_ctx], -1), sd["fuse_context.weight"])
    for i in range(n_blocks(sd, "decoder_layers")):
        fused, ctx = run_block(sd, f"decoder_layers.{i}", fused, emb, ctx, cos, sin, H, ctx_mask)
    out = last_layer(sd, fused, emb)
    return unpatchify(out, hp, wp, ps, cfg["output_channels"])


def ddt_forward(sd: SD, cfg: dict[str, Any], x: Tensor, t: Tensor, y: Tensor | None = None, context: dict | None = None,
                p: float = 0.0, draws: dict[str, Tensor] | None = None, capture: dict | None = None) -> Tensor:
    """DDT.forward / encode_* / decode denoisers/ddt.py:274-512"""
    draws = draws or {}
    H, ps = cfg["num_heads"], cfg["patch_size"]
    tok, hp, wp = patchify(x, sd["conv_proj_encoder.weight"])
    emb = time_embed(sd, t, cfg.get("frequency_embedding", 256))
    ctx, ctx_mask = None, None
    if context is None:
        if "label_embed.embedding.weight" in sd:
            emb = emb + label_embed(sd, y, cfg["n_classes"], p, draws.get("label"))
        cos, sin = rope_tables(pos_ids_2d(hp, wp).to(x.device), cfg["rope_axes_dim"], cfg.get("rope_base", 10000))
        cos_d, sin_d = cos, sin
    else:
        ce, ctx_mask = drop_context(context, cfg["null_embedding"], cfg["null_mask"], p, draws.get("context"))
        ctx = linear(ce, sd["context_embed.weight"])
        L = ctx.shape[1]
        cos, sin = rope_tables(pos_ids_joint(L, hp, wp).to(x.device), cfg["rope_axes_dim"], cfg.get("rope_base", 10000))
        cos_d, sin_d = cos[L:], sin[L:]  # decoder: image positions (0,h,w) only (ddt.py:425-449)
    for i in range(n_blocks(sd, "layers")):
        tok, ctx = run_block(sd, f"layers.{i}", tok, emb, ctx, cos, sin, H, ctx_mask)
        if capture is not None:
            capture[f"layers.{i}"] = tok
    # decoder: per-token conditioning silu(enc + time_emb) (ddt.py:421-422); time_embed evaluated again (same value)
    cond = r(F.silu(tok + time_embed(sd, t, cfg.get("frequency_embedding", 256))[:, None, :]))
    z, _, _ = patchify(x, sd["conv_proj_decoder.weight"])
    for i in range(n_blocks(sd, "decoder_layers")):
        z = dit_block(sd, f"decoder_layers.{i}", z, cond, cos_d, sin_d, H)
    out = last_layer(sd, z, cond)
    return unpatchify(out, hp, wp, ps, cfg["output_channels"])


# ---------------------------------------------------------------------------------------------------------
# formalisation: flow matching, REPA, Euler sampling
# ---------------------------------------------------------------------------------------------------------
def shift_timestep(t, alpha: float):
    """flow.py:84-99"""
    return alpha * t / (1 + (alpha - 1) * t)


def flow_timesteps(n_steps: int, shift: float | None) -> list[float]:
    """Flow.set_steps flow.py:125-131"""
    ts = torch.linspace(1, 0, n_steps + 1).tolist()
    if shift is not None:
        ts = [shift_timestep(v, shift) for v in ts]
    return ts


def flow_add_noise(x0: Tensor, t: Tensor, eps: Tensor) -> Tensor:
    """flow.py:405-407"""
    shape = (-1,) + (1,) * (x0.dim() - 1)
    return (torch.ones_like(t) - t).view(shape) * x0 + t.view(shape) * eps


def flow_loss(pred: Tensor, x0: Tensor, eps: Tensor, x_t: Tensor | None = None, t: Tensor | None = None) -> Tensor:
    """flow.py:300-308 (x_t/t given -> x-prediction)"""
    if x_t is not None:
        pred = (x_t - pred) / t.view((-1,) + (1,) * (x_t.dim() - 1))
    losses = ((eps - x0) - pred) ** 2
    return losses.reshape(losses.shape[0], -1).mean(dim=-1).mean()


def repa_loss(sd_repa: SD, feats: Tensor, dst: Tensor, coeff: float) -> Tensor:
    """RepaLoss.forward training/losses/repa.py:176-186 (projector proj.{0,2,4})"""
    h = r(F.silu(linear(feats, sd_repa["proj.0.weight"], sd_repa["proj.0.bias"])))
    h = r(F.silu(linear(h, sd_repa["proj.2.weight"], sd_repa["proj.2.bias"])))
    s = linear(h, sd_repa["proj.4.weight"], sd_repa["proj.4.bias"]).float()  # F.cosine_similarity runs in fp32 under autocast
    dot = (s * dst).sum(-1)
    ns = s.norm(dim=-1).clamp_min(1e-8)
    nz = dst.norm(dim=-1).clamp_min(1e-8)
    return coeff * (1 - (dot / (ns * nz)).mean())


def perceiver_resampler(sd: SD, x: Tensor, num_heads: int, head_dim: int, rope_axes_dim: list[int], rope_base: float, prefix: str = "") -> Tensor:
    """PerceiverResampler.forward networks/repa/perceiver_resampler.py:90-252 (PerceiverAttention :119-166, FeedForward :59-77).
    The reference builds un-batched position ids and then indexes the tables with a batch dimension (IndexError, SURVEY.md
    4.3-3); the tables are sample-independent, so they are used un-batched here (golden fixtures call the reference with
    explicitly batched `cos_sin`, which is the same arithmetic). Key-only RoPE on the input tokens' keys; latents attend over
    cat(x keys, latent keys)."""
    B, N, dim = x.shape
    H, hd = num_heads, head_dim
    hw = int(N**0.5)
    cos, sin = rope_tables(pos_ids_2d(hw, hw).to(x.device), rope_axes_dim, rope_base)
    lat = sd[f"{prefix}latents"].to(x.dtype)[None].expand(B, -1, -1)
    M = lat.shape[1]
    n_layers = n_blocks(sd, f"{prefix}layers") if prefix else max(int(k.split(".")[1]) for k in sd if k.startswith("layers.")) + 1
    for i in range(n_layers):
        a, f = f"{prefix}layers.{i}.0", f"{prefix}layers.{i}.1"
        xn = r(layer_norm(x, sd[f"{a}.norm_x.weight"], sd[f"{a}.norm_x.bias"], 1e-5))
        ln = r(layer_norm(lat, sd[f"{a}.norm_latents.weight"], sd[f"{a}.norm_latents.bias"], 1e-5))
        q = linear(ln, sd[f"{a}.to_q.weight"]).view(B, M, H, hd)
        kx, vx = linear(xn, sd[f"{a}.to_kv.weight"]).chunk(2, dim=-1)
        kl, vl = linear(ln, sd[f"{a}.to_kv.weight"]).chunk(2, dim=-1)
        kx = apply_rope(kx.reshape(B, N, H, hd), cos, sin)
        k = torch.cat([kx, kl.reshape(B, M, H, hd)], 1)
        v = torch.cat([vx.reshape(B, N, H, hd), vl.reshape(B, M, H, hd)], 1)
        sim = torch.einsum("bihd,bjhd->bhij", r(q * hd**-0.5), k)
        attn = r((sim - sim.amax(dim=-1, keepdim=True)).softmax(-1))
        o = r(torch.einsum("bhij,bjhd->bihd", attn, v)).reshape(B, M, H * hd)
        lat = r(linear(o, sd[f"{a}.to_out.weight"]) + lat)
        h = r(layer_norm(lat, sd[f"{f}.0.weight"], sd[f"{f}.0.bias"], 1e-5))
        h = r(F.gelu(linear(h, sd[f"{f}.1.weight"])))
        lat = r(linear(h, sd[f"{f}.3.weight"]) + lat)
    return r(layer_norm(lat, sd[f"{prefix}norm.weight"], sd[f"{prefix}norm.bias"], 1e-5))


def repa_loss_resampled(sd_repa: SD, feats: Tensor, dst: Tensor, coeff: float, resampler_kw: dict) -> Tensor:
    """RepaLoss.forward with use_resampler=True (training/losses/repa.py:176-186): proj -> PerceiverResampler -> cosine."""
    h = r(F.silu(linear(feats, sd_repa["proj.0.weight"], sd_repa["proj.0.bias"])))
    h = r(F.silu(linear(h, sd_repa["proj.2.weight"], sd_repa["proj.2.bias"])))
    s = linear(h, sd_repa["proj.4.weight"], sd_repa["proj.4.bias"])
    s = perceiver_resampler(sd_repa, s, resampler_kw["num_heads"], resampler_kw["head_dim"], resampler_kw["rope_axes_dim"],
                            resampler_kw.get("rope_base", 10000), prefix="resampler.").float()
    dot = (s * dst).sum(-1)
    ns = s.norm(dim=-1).clamp_min(1e-8)
    nz = dst.norm(dim=-1).clamp_min(1e-8)
    return coeff * (1 - (dot / (ns * nz)).mean())


def euler_step(x: Tensor, v: Tensor, t_curr: float, t_prev: float) -> tuple[Tensor, Tensor]:
    """Euler.step samplers/flow/euler.py:37-39"""
    return x - v * (t_curr - t_prev), x - v * t_curr


def cfg_combine(v: Tensor, v_dropped: Tensor, g: float) -> Tensor:
    """flow.py:259"""
    return v_dropped + g * (v - v_dropped)


def flow_denoise(velocity: Callable[[Tensor, float, float], Tensor], x: Tensor, n_steps: int, shift: float | None,
                 guidance_scale: float = 0.0, method: str = "euler") -> Tensor:
    """Flow.denoise flow.py:484-499. velocity(x, t, p) -> model(x, t, p)['x'] (v-prediction).
    method 'heun' has no reference counterpart: the textbook trapezoidal predictor-corrector built from two
    reference-style velocity evaluations (each with the reference's CFG combine), last step (t_prev = 0) plain Euler."""

    def guided(xx: Tensor, t: float) -> Tensor:
        v = velocity(xx, t, 0.0)
        if guidance_scale > 0:
            v = cfg_combine(v, velocity(xx, t, 1.0), guidance_scale)
        return v

    ts = flow_timesteps(n_steps, shift)
    for t_curr, t_prev in zip(ts[:-1], ts[1:]):
        v = guided(x, t_curr)
        if method == "heun" and t_prev > 0:
            x_pred, _ = euler_step(x, v, t_curr, t_prev)
            v = 0.5 * (v + guided(x_pred, t_prev))
        x, _ = euler_step(x, v, t_curr, t_prev)
    return x
