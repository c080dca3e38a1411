"""PerceiverResampler for REPA: drop-in for reference networks/repa/perceiver_resampler.py:90-252 (same constructor, parameter
names and shapes: `latents`, `layers.{i}.0.{norm_x,norm_latents,to_q,to_kv,to_out}`, `layers.{i}.1.{0,1,3}`, `norm`), executed by
the kernels of libdiffulab_b200.so: LayerNorm (the LN + modulate kernel with zero modulation), tcgen05 GEMMs, key-only N-D RoPE
(`dlb_rope_apply`), the joint-attention kernel, exact GELU (`dlb_gelu_*`).

Cross-attention on the self-attention kernel: queries come from the M latents only, keys / values from cat(x tokens, latents)
(perceiver_resampler.py:143-158). The joint kernel attends over two segments [x (N rows), latents (M rows)] with ZERO queries for
the x segment; its outputs for those rows are discarded and their output gradient is zero, which makes every x-query
contribution to dK / dV vanish exactly (dO = 0 => dP = 0 and D = 0 => dS = 0).

Reference defect kept observable: `PerceiverResampler.forward` builds un-batched position ids `[S, 2]` but the
rotary helper indexes the tables as `[B, S, D/2]`, so the reference raises IndexError unless `cos_sin` is passed with a batch
dimension. The tables do not depend on the sample; this implementation uses them un-batched (the evident intent) and accepts the
reference's optional `cos_sin` argument for signature compatibility only when it is None.
"""

from __future__ import annotations

import torch
from torch import Tensor, nn

from .. import blocks as K
from .. import ops

BF16 = torch.bfloat16
F32 = torch.float32


class _LayerNormFn(torch.autograd.Function):
    """nn.LayerNorm(dim) (affine, eps 1e-5) on bf16 rows: the LN + modulate kernels with zero scale / shift."""

    @staticmethod
    def forward(ctx, x: Tensor, w: Tensor, b: Tensor, eps: float):
        x2 = x.reshape(-1, x.shape[-1]).contiguous()
        d = x2.shape[-1]
        zero = torch.zeros(1, 2 * d, device=x.device, dtype=BF16)
        y, mean, rstd = ops.ln_modulate_fwd(x2, w.detach(), b.detach(), zero[:, :d], zero[:, d:], eps)
        ctx.save_for_backward(x2, mean, rstd, zero)
        ctx.w, ctx.b, ctx.shape = w, b, x.shape
        return y.view(x.shape)

    @staticmethod
    def backward(ctx, dy: Tensor):
        x2, mean, rstd, zero = ctx.saved_tensors
        d = x2.shape[-1]
        scratch = torch.zeros(1, 2 * d, device=x2.device, dtype=F32)  # dscale / dshift of the (absent) modulation
        dx = ops.ln_modulate_bwd(dy.reshape(-1, d).contiguous(), x2, mean, rstd, ctx.w.detach(), ctx.b.detach(), zero[:, :d], None,
                                 scratch[:, :d], scratch[:, d:], K.vgrad(ctx.w), K.vgrad(ctx.b))
        K._ready(ctx.w)
        K._ready(ctx.b)
        return dx.view(ctx.shape), None, None, None


class _GeluFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x: Tensor):
        x = x.contiguous()
        ctx.save_for_backward(x)
        return ops.gelu_fwd(x)

    @staticmethod
    def backward(ctx, dy: Tensor):
        (x,) = ctx.saved_tensors
        return ops.gelu_bwd(dy.contiguous(), x)


class _AddFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a: Tensor, b: Tensor):
        return ops.add(a.contiguous(), b.contiguous())

    @staticmethod
    def backward(ctx, dy: Tensor):
        return dy, dy


class _LatentsFn(torch.autograd.Function):
    """repeat(latents, 'n d -> b n d') in bf16; gradient = sum over the batch, accumulated by the column-sum kernel."""

    @staticmethod
    def forward(ctx, latents: Tensor, B: int):
        ctx.latents = latents
        lat = ops.cast_bf16(latents.detach().contiguous())
        return lat.unsqueeze(0).expand(B, -1, -1).contiguous()  # layout plumbing only

    @staticmethod
    def backward(ctx, dy: Tensor):
        lat = ctx.latents
        if lat.requires_grad:
            ops.colsum_(dy.reshape(dy.shape[0], -1).contiguous(), K.gbuf(lat).view(-1))
            K._ready(lat)
        return None, None


class _CrossAttnFn(torch.autograd.Function):
    """softmax(q_lat [k_x_rot ; k_lat]^T * hd^-1/2) [v_x ; v_lat] per head (perceiver_resampler.py:143-164).
    q [B*M, inner]; kv_x [B*N, 2*inner] (k | v); kv_l [B*M, 2*inner]. Returns [B*M, inner]."""

    @staticmethod
    def forward(ctx, q: Tensor, kv_x: Tensor, kv_l: Tensor, rope, B: int, N: int, M: int, H: int, hd: int):
        inner = H * hd
        kx_rot = ops.rope_apply(kv_x[:, :inner], rope, hd, tokens_per_sample=N)
        qx0 = torch.zeros(B * N, inner, device=q.device, dtype=BF16)
        specs = [ops.AttnSegViews(qx0, kx_rot, kv_x[:, inner:], N), ops.AttnSegViews(q, kv_l[:, :inner], kv_l[:, inner:], M)]
        outs, lse = ops.attn_fwd(specs, B, H, hd, hd**-0.5, None)
        ctx.save_for_backward(q, kv_x, kv_l, kx_rot, qx0, outs[0], outs[1], lse)
        ctx.meta = (rope, B, N, M, H, hd)
        return outs[1]

    @staticmethod
    def backward(ctx, dout: Tensor):
        q, kv_x, kv_l, kx_rot, qx0, o_x, o_l, lse = ctx.saved_tensors
        rope, B, N, M, H, hd = ctx.meta
        inner = H * hd
        specs = [ops.AttnSegViews(qx0, kx_rot, kv_x[:, inner:], N), ops.AttnSegViews(q, kv_l[:, :inner], kv_l[:, inner:], M)]
        grads = ops.attn_bwd_views(specs, [o_x, o_l], [torch.zeros_like(o_x), dout.contiguous()], lse, B, H, hd, hd**-0.5, None)
        (_, dkx_rot, dvx), (dq, dkl, dvl) = grads
        dkv_x = torch.empty_like(kv_x)
        ops.rope_apply(dkx_rot, rope, hd, tokens_per_sample=N, inverse=True, out=dkv_x[:, :inner])
        dkv_x[:, inner:].copy_(dvx)  # layout plumbing: pack (dk | dv) for the to_kv weight-gradient GEMM
        dkv_l = torch.cat([dkl, dvl], dim=1)
        return dq, dkv_x, dkv_l, None, None, None, None, None, None


class PerceiverAttention(nn.Module):
    """reference perceiver_resampler.py:80-166 (parameter holder + forward over the kernels above)"""

    def __init__(self, dim: int, axes_dim: list[int], head_dim: int = 64, num_heads: int = 8) -> None:
        super().__init__()
        self.scale = head_dim**-0.5
        self.num_heads, self.head_dim = num_heads, head_dim
        inner_dim = head_dim * num_heads
        self.norm_x = nn.LayerNorm(dim)
        self.norm_latents = nn.LayerNorm(dim)
        self.to_q = nn.Linear(dim, inner_dim, bias=False)
        self.to_kv = nn.Linear(dim, inner_dim * 2, bias=False)
        self.to_out = nn.Linear(inner_dim, dim, bias=False)
        self.axes_dim = list(axes_dim)

    def forward(self, x: Tensor, latents: Tensor, rope) -> Tensor:
        B, N, dim = x.shape
        M = latents.shape[1]
        xn = _LayerNormFn.apply(x, self.norm_x.weight, self.norm_x.bias, self.norm_x.eps)
        ln = _LayerNormFn.apply(latents, self.norm_latents.weight, self.norm_latents.bias, self.norm_latents.eps)
        q = K.linear(ln.view(B * M, dim), self.to_q.weight, None)
        kv_x = K.linear(xn.view(B * N, dim), self.to_kv.weight, None)
        kv_l = K.linear(ln.view(B * M, dim), self.to_kv.weight, None)
        o = _CrossAttnFn.apply(q, kv_x, kv_l, rope, B, N, M, self.num_heads, self.head_dim)
        return K.linear(o, self.to_out.weight, None).view(B, M, dim)


def FeedForward(dim: int, mult: float = 4) -> nn.Sequential:
    """reference perceiver_resampler.py:59-77 (index 2 is the parameter-free GELU)"""
    inner_dim = int(dim * mult)
    return nn.Sequential(nn.LayerNorm(dim), nn.Linear(dim, inner_dim, bias=False), nn.GELU(), nn.Linear(inner_dim, dim, bias=False))


class PerceiverResampler(nn.Module):
    def __init__(self, dim: int, depth: int, rope_axes_dim: list[int] | None = None, head_dim: int = 64, num_heads: int = 8, ff_mult: int = 4,
                 num_latents: int = 16, rope_base: int = 10_000):
        super().__init__()
        assert head_dim % 8 == 0 and head_dim <= 128 and dim % 8 == 0, "head_dim must be a multiple of 8 (<= 128), dim a multiple of 8"
        self.latents = nn.Parameter(torch.randn(num_latents, dim))
        self.rope_base = rope_base
        if rope_axes_dim is None:
            rope_axes_dim = [int(head_dim // 2), int(head_dim // 2)]
        self.layers = nn.ModuleList(
            [nn.ModuleList([PerceiverAttention(dim=dim, axes_dim=rope_axes_dim, head_dim=head_dim, num_heads=num_heads), FeedForward(dim=dim, mult=ff_mult)])
             for _ in range(depth)]
        )
        self.rope_axes_dim = list(rope_axes_dim)
        self.norm = nn.LayerNorm(dim)
        self._rope_cache: dict = {}

    def _rope(self, n_tokens: int, device) -> ops.RopeTable:
        key = (n_tokens, str(device))
        if key not in self._rope_cache:
            hw = int(n_tokens**0.5)
            assert hw * hw == n_tokens, "PerceiverResampler expects a square token grid (reference perceiver_resampler.py:236)"
            hh, ww = torch.meshgrid(torch.arange(hw), torch.arange(hw), indexing="ij")
            pos = torch.stack([hh.reshape(-1), ww.reshape(-1)], -1).to(torch.int32).to(device).contiguous()
            self._rope_cache[key] = ops.rope_table(pos, self.rope_axes_dim, float(self.rope_base))
        return self._rope_cache[key]

    @staticmethod
    def _ff(ff: nn.Sequential, x: Tensor) -> Tensor:
        B, M, dim = x.shape
        h = _LayerNormFn.apply(x, ff[0].weight, ff[0].bias, ff[0].eps)
        h = K.linear(h.view(B * M, dim), ff[1].weight, None)
        h = _GeluFn.apply(h)
        return K.linear(h, ff[3].weight, None).view(B, M, dim)

    def forward(self, x: Tensor, cos_sin=None) -> Tensor:
        assert cos_sin is None, "precomputed cos / sin tables are not accepted: the tables are built (and cached) from the token grid"
        if x.dtype != BF16:
            x = ops.cast_bf16(x.float().contiguous())
        B, N, _ = x.shape
        rope = self._rope(N, x.device)
        latents = _LatentsFn.apply(self.latents, B)
        for attn, ff in self.layers:
            latents = _AddFn.apply(attn(x, latents, rope), latents)
            latents = _AddFn.apply(self._ff(ff, latents), latents)
        return _LayerNormFn.apply(latents, self.norm.weight, self.norm.bias, self.norm.eps)
