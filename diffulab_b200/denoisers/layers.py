"""Parameter-holding modules with the reference's names and initialisation; their forward passes launch the
block-level autograd Functions of `diffulab_b200.blocks` (hand-written sm_100a kernels behind the C ABI).

`nn.Linear` / `nn.LayerNorm` / `nn.Conv2d` / `nn.Embedding` are used purely as parameter containers so that
`state_dict()` keys and shapes equal the reference's; their own forward is never called.
"""

from __future__ import annotations

import torch
import torch.nn as nn
from torch import Tensor

from .. import blocks as K
from .. import ops


class RMSNorm(nn.Module):
    """reference networks/utils/nn.py:403-431 (parameter holder)"""

    def __init__(self, dim: int):
        super().__init__()
        self.scale = nn.Parameter(torch.ones(dim))


class QKNorm(nn.Module):
    """reference nn.py:434-475"""

    def __init__(self, dim: int):
        super().__init__()
        self.query_norm = RMSNorm(dim)
        self.key_norm = RMSNorm(dim)


class Modulation(nn.Module):
    """reference nn.py:499-536: lin(silu(vec)) chunked 6-way. `cond_silu` is silu(vec) already (bf16)."""

    def __init__(self, embedding_dim: int, input_dim: int):
        super().__init__()
        self.lin = nn.Linear(embedding_dim, 6 * input_dim, bias=True)

    def forward(self, cond_silu: Tensor) -> Tensor:
        return K.linear(cond_silu, self.lin.weight, self.lin.bias)


class LabelEmbed(nn.Module):
    """reference nn.py:117-164 (label dropout draws `torch.rand(labels.size()) < p` exactly as the reference)"""

    def __init__(self, num_classes: int, embed_dim: int, classifier_free_guidance: bool = False) -> None:
        super().__init__()
        self.num_classes = num_classes
        self.embed_dim = embed_dim
        self.classifier_free_guidance = classifier_free_guidance
        self.embedding = nn.Embedding(num_classes + 1 if classifier_free_guidance else num_classes, embed_dim)

    def drop_labels(self, labels: Tensor, p: float) -> Tensor:
        return torch.where(torch.rand(labels.size(), device=labels.device) < p, self.num_classes, labels)

    def labels_for(self, labels: Tensor, p: float = 0) -> Tensor:
        if p > 0:
            assert self.classifier_free_guidance, "Label dropout is only supported with classifier-free guidance."
            labels = self.drop_labels(labels, p)
        return labels.reshape(-1).to(torch.int64).contiguous()


class DiTAttention(nn.Module):
    """reference denoisers/mmdit.py:29-104 (parameters only; executed inside the block Functions)"""

    def __init__(self, inner_dim: int, num_heads: int, rope_axes_dim: list[int]) -> None:
        super().__init__()
        self.num_heads = num_heads
        self.head_dim = inner_dim // num_heads
        self.scale = self.head_dim**-0.5
        self.rope_axes_dim = rope_axes_dim
        self.qkv = nn.Linear(inner_dim, 3 * inner_dim, bias=False)
        self.qk_norm = QKNorm(inner_dim)
        self.proj_out = nn.Linear(inner_dim, inner_dim, bias=False)


class MMDiTAttention(nn.Module):
    """reference mmdit.py:107-210"""

    def __init__(self, inner_dim: int, num_heads: int, rope_axes_dim: list[int]):
        super().__init__()
        self.num_heads = num_heads
        self.head_dim = inner_dim // num_heads
        self.scale = self.head_dim**-0.5
        self.rope_axes_dim = rope_axes_dim
        self.qkv_input = nn.Linear(inner_dim, 3 * inner_dim, bias=False)
        self.qkv_context = nn.Linear(inner_dim, 3 * inner_dim, bias=False)
        self.qk_norm_input = QKNorm(inner_dim)
        self.qk_norm_context = QKNorm(inner_dim)
        self.input_proj_out = nn.Linear(inner_dim, inner_dim, bias=False)
        self.context_proj_out = nn.Linear(inner_dim, inner_dim, bias=False)


def _mlp(inner_dim: int, mlp_ratio: int) -> nn.Sequential:
    # index 1 is the parameter-free PackedSwiGLU of the reference (nn.py:478-486)
    return nn.Sequential(
        nn.Linear(inner_dim, mlp_ratio * inner_dim * 2, bias=False),
        nn.Identity(),
        nn.Linear(mlp_ratio * inner_dim, inner_dim, bias=False),
    )


class DiTBlock(nn.Module):
    """reference mmdit.py:213-309. forward(x, cond_silu, rope, pos_idx): cond_silu is silu(conditioning) in bf16,
    [B,E] (per sample) or [B,N,E] (per token, DDT decoder)."""

    def __init__(self, inner_dim: int, embedding_dim: int, num_heads: int, mlp_ratio: int, rope_axes_dim: list[int],
                 use_checkpoint: bool = False):
        super().__init__()
        self.num_heads = num_heads
        self.modulation = Modulation(embedding_dim, inner_dim)
        self.norm_1 = nn.LayerNorm(inner_dim)
        self.attention = DiTAttention(inner_dim, num_heads, rope_axes_dim=rope_axes_dim)
        self.norm_2 = nn.LayerNorm(inner_dim)
        self.mlp_input = _mlp(inner_dim, mlp_ratio)
        self.use_checkpoint = use_checkpoint  # accepted for config compatibility; activations are always kept

    def forward(self, x: Tensor, cond_silu: Tensor, rope: K.RopeCtx, pos_idx: Tensor | None = None, pos_offset: int = 0) -> Tensor:
        mod = self.modulation(cond_silu)
        return K.DiTBlockFn.apply(x, mod, self, rope, pos_idx, pos_offset, *K.block_params(self))


class MMDiTBlock(nn.Module):
    """reference mmdit.py:312-459. forward(x, cond_silu, context, rope, kmask, pos_idx) -> (x, context)"""

    def __init__(self, inner_dim: int, embedding_dim: int, num_heads: int, mlp_ratio: int, rope_axes_dim: list[int],
                 use_checkpoint: bool = False):
        super().__init__()
        self.num_heads = num_heads
        self.modulation_context = Modulation(embedding_dim, inner_dim)
        self.modulation_input = Modulation(embedding_dim, inner_dim)
        self.context_norm_1 = nn.LayerNorm(inner_dim)
        self.input_norm_1 = nn.LayerNorm(inner_dim)
        self.attention = MMDiTAttention(inner_dim, num_heads, rope_axes_dim=rope_axes_dim)
        self.context_norm_2 = nn.LayerNorm(inner_dim)
        self.input_norm_2 = nn.LayerNorm(inner_dim)
        self.mlp_context = _mlp(inner_dim, mlp_ratio)
        self.mlp_input = _mlp(inner_dim, mlp_ratio)
        self.use_checkpoint = use_checkpoint

    def forward(self, x: Tensor, cond_silu: Tensor, context: Tensor, rope: K.RopeCtx, kmask: Tensor | None = None,
                pos_idx: Tensor | None = None) -> tuple[Tensor, Tensor]:
        mod_x = self.modulation_input(cond_silu)
        mod_c = self.modulation_context(cond_silu)
        return K.MMDiTBlockFn.apply(x, context, mod_x, mod_c, self, rope, kmask, pos_idx, *K.block_params(self))


class MMDiTSingleStreamBlock(nn.Module):
    """reference mmdit.py:462-532. Returns (x, context) like the dual-stream block."""

    def __init__(self, inner_dim: int, embedding_dim: int, num_heads: int, mlp_ratio: int, rope_axes_dim: list[int],
                 use_checkpoint: bool = False):
        super().__init__()
        self.num_heads = num_heads
        self.mlp = _mlp(inner_dim, mlp_ratio)
        self.attention = DiTAttention(inner_dim, num_heads, rope_axes_dim=rope_axes_dim)
        self.modulation = nn.Sequential(nn.SiLU(), nn.Linear(embedding_dim, 3 * inner_dim))
        self.norm = nn.LayerNorm(inner_dim)
        self.use_checkpoint = use_checkpoint

    def forward(self, x: Tensor, cond_silu: Tensor, context: Tensor, rope: K.RopeCtx, kmask: Tensor | None = None,
                pos_idx: Tensor | None = None) -> tuple[Tensor, Tensor]:
        L = context.shape[1]
        z = torch.cat([context, x], dim=1)  # layout plumbing only (text rows first, mmdit.py:507)
        mod = K.linear(cond_silu, self.modulation[1].weight, self.modulation[1].bias)
        if pos_idx is not None:  # per-row table rows for the whole joint sequence
            B = x.shape[0]
            text = torch.arange(L, device=x.device, dtype=torch.int32).expand(B, L)
            pos_idx = torch.cat([text, pos_idx.view(B, -1)], dim=1).reshape(-1).contiguous()
        z = K.SingleStreamBlockFn.apply(z, mod, self, rope, kmask, pos_idx, *K.block_params(self))
        return z[:, L:, :], z[:, :L, :]


class ModulatedLastLayer(nn.Module):
    """reference mmdit.py:535-549 fused with unpatchify (mmdit.py:767-787)."""

    def __init__(self, embedding_dim: int, hidden_size: int, patch_size: int, out_channels: int):
        super().__init__()
        self.norm_final = nn.LayerNorm(hidden_size, elementwise_affine=False, eps=1e-6)
        self.linear = nn.Linear(hidden_size, patch_size * patch_size * out_channels)
        self.adaLN_modulation = nn.Sequential(nn.SiLU(), nn.Linear(embedding_dim, 2 * hidden_size))
        self.patch_size, self.out_channels = patch_size, out_channels

    def forward(self, x: Tensor, cond_silu: Tensor, image_hw: tuple[int, int]) -> Tensor:
        mod = K.linear(cond_silu, self.adaLN_modulation[1].weight, self.adaLN_modulation[1].bias)
        geom = (x.shape[0], self.out_channels, image_hw[0], image_hw[1], self.patch_size)
        return K.FinalLayerFn.apply(x, mod, self.linear.weight, self.linear.bias, geom)


def zero_module(module: nn.Module) -> nn.Module:
    """reference networks/utils/utils.py"""
    for p in module.parameters():
        p.detach().zero_()
    return module


def init_weights(module: nn.Module) -> None:
    """reference mmdit.py:737-745"""
    if isinstance(module, (nn.Linear, nn.Conv2d)):
        nn.init.xavier_uniform_(module.weight)
        if module.bias is not None:
            nn.init.constant_(module.bias, 0)
    if isinstance(module, Modulation):
        zero_module(module)
    if isinstance(module, ModulatedLastLayer):
        zero_module(module.adaLN_modulation)


# ---------------------------------------------------------------------------------------------------------
# RoPE table cache (reference recomputes get_cos_sin_ndim_grid every forward, nn.py:262-307)
# ---------------------------------------------------------------------------------------------------------
_rope_cache: dict[tuple, K.RopeCtx] = {}


def rope_for(device: torch.device, L_text: int, hp: int, wp: int, axes_dim: list[int], base: float, joint: bool) -> K.RopeCtx:
    """joint=False: 2 axes (h, w), row-major tokens (mmdit.py:871-886). joint=True: 3 axes, text rows (l,0,0) with
    l = 1..L first, then image rows (0,h,w) (mmdit.py:815-835)."""
    key = (str(device), L_text, hp, wp, tuple(axes_dim), float(base), joint)
    ctx = _rope_cache.get(key)
    if ctx is None:
        hh, ww = torch.meshgrid(torch.arange(hp), torch.arange(wp), indexing="ij")
        if joint:
            text = torch.stack([torch.arange(1, L_text + 1), torch.zeros(L_text, dtype=torch.long), torch.zeros(L_text, dtype=torch.long)], -1)
            img = torch.stack([torch.zeros(hp * wp, dtype=torch.long), hh.reshape(-1), ww.reshape(-1)], -1)
            pos = torch.cat([text, img], 0)
        else:
            pos = torch.stack([hh.reshape(-1), ww.reshape(-1)], -1)
        ctx = ops.rope_table(pos.to(torch.int32).to(device).contiguous(), list(axes_dim), float(base))
        _rope_cache[key] = ctx
    return ctx
