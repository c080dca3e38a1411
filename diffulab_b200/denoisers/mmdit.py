"""MMDiT / DiT denoiser: drop-in for reference networks/denoisers/mmdit.py:552-928 (same constructor arguments,
same `forward(x, timesteps, initial_context, p, y, x_context, intermediate_features)` signature, same state_dict
keys), executed by hand-written sm_100a kernels. bf16 compute with fp32 statistics/accumulation, i.e. the
reference's CUDA bf16-autocast numerics; the returned "x" is bf16 [B, C_out, H, W]."""

from __future__ import annotations

import logging
from typing import Any

import torch
import torch.nn as nn
from torch import Tensor

from .. import blocks as K
from .. import ops
from ..embedders.common import ContextEmbedder
from .common import Denoiser, ModelOutput
from .layers import DiTBlock, LabelEmbed, MMDiTBlock, MMDiTSingleStreamBlock, ModulatedLastLayer, init_weights, rope_for


class _DenoiserBase(Denoiser):
    """Shared conditioning / context plumbing of MMDiT, SprintDiT and DDT."""

    simple: bool
    patch_size: int
    frequency_embedding: int
    rope_base: float
    rope_axes_dim: list[int]
    label_embed: LabelEmbed | None
    context_embedder: ContextEmbedder | None
    # True when the drop probability `p` touches nothing but the label / context (so Flow may batch the conditional and
    # unconditional guidance passes into one forward); SprintDiT's p also drives path-drop and opts out.
    cfg_batchable: bool = True

    def _conditioning(self, timesteps: Tensor, y: Tensor | None, p: float) -> tuple[Tensor, Tensor]:
        labels, table = None, None
        if self.simple and getattr(self, "label_embed", None) is not None:
            assert y is not None, "class labels `y` are required by a label-conditioned model"
            labels = self.label_embed.labels_for(y, p)
            table = self.label_embed.embedding.weight
        te = self.time_embed
        return K.CondFn.apply(timesteps, labels, self.frequency_embedding, te[0].weight, te[0].bias, te[2].weight, te[2].bias, table)

    def _context(self, initial_context: Any, p: float, cond: tuple[Tensor, Tensor]) -> tuple[Tensor, Tensor | None]:
        assert self.context_embedder is not None, "for MMDiT context embedder must be provided"
        out = self.context_embedder(initial_context, p)
        assert not self.pooled_embedding, "pooled context embeddings are outside the accelerated hot path"
        emb = out["embeddings"]
        emb_b = emb if emb.dtype == torch.bfloat16 else ops.cast_bf16(emb.float().contiguous())
        ctx = K.linear(emb_b, self.context_embed.weight, None)
        mask = out.get("attn_mask", None)
        kmask = mask.to(torch.uint8).contiguous() if mask is not None else None
        return ctx, kmask

    def _setup_mm(self, context_embedder: ContextEmbedder, inner_dim: int, embedding_dim: int) -> None:
        assert isinstance(context_embedder.output_size, tuple) and all(isinstance(i, int) for i in context_embedder.output_size), (
            "context_embedder.output_size must be a tuple of integers"
        )
        self.pooled_embedding = False
        self.mlp_pooled_context = None
        if context_embedder.n_output == 2:
            self.pooled_embedding = True
            self.mlp_pooled_context = nn.Sequential(
                nn.Linear(context_embedder.output_size[0], embedding_dim * 2), nn.SiLU(), nn.Linear(embedding_dim * 2, embedding_dim)
            )
            self.context_embed = nn.Linear(context_embedder.output_size[1], inner_dim, bias=False)
        else:
            assert context_embedder.n_output == 1
            self.context_embed = nn.Linear(context_embedder.output_size[0], inner_dim, bias=False)

    def _run_layers(self, layers, x, cond_silu, context, rope, kmask, features, pos_idx=None):
        for layer in layers:
            if isinstance(layer, DiTBlock):
                x = layer(x, cond_silu, rope, pos_idx)
            else:
                x, context = layer(x, cond_silu, context, rope, kmask, pos_idx)
            if features is not None:
                features.append(x)
        return x, context


def _default_axes(simple: bool, heads_dim: int, partial_rotary_factor: float) -> list[int]:
    n = 2 if simple else 3
    return [int((partial_rotary_factor * heads_dim) // n)] * n


class MMDiT(_DenoiserBase):
    def __init__(
        self,
        simple_dit: bool = False,
        input_channels: int = 3,
        output_channels: int | None = None,
        inner_dim: int = 4096,
        embedding_dim: int = 4096,
        num_heads: int = 16,
        mlp_ratio: int = 4,
        patch_size: int = 16,
        depth: int = 38,
        n_single_stream_blocks: int = 0,
        rope_base: int = 10_000,
        partial_rotary_factor: float = 1,
        rope_axes_dim: list[int] | None = None,
        frequency_embedding: int = 256,
        n_classes: int | None = None,
        classifier_free: bool = False,
        context_embedder: ContextEmbedder | None = None,
        use_checkpoint: bool = False,
    ):
        super().__init__()
        assert not (n_classes is not None and context_embedder is not None), "n_classes and context_embedder cannot both be specified"
        self.simple_dit = self.simple = simple_dit
        self.patch_size = patch_size
        self.input_channels = input_channels
        self.output_channels = output_channels or input_channels
        self.context_embedder = context_embedder
        self.frequency_embedding = frequency_embedding
        self.rope_base = rope_base
        self.n_classes = n_classes
        self.classifier_free = classifier_free
        self.inner_dim, self.num_heads = inner_dim, num_heads
        heads_dim = inner_dim // num_heads
        assert heads_dim % 8 == 0 and heads_dim <= 128, "head_dim must be a multiple of 8 and <= 128"
        if not simple_dit:
            assert context_embedder is not None, "for MMDiT context embedder must be provided"
            self._setup_mm(context_embedder, inner_dim, embedding_dim)
        else:
            self.label_embed = LabelEmbed(n_classes, embedding_dim, classifier_free) if n_classes is not None else None
            if n_single_stream_blocks > 0:
                logging.warning("n_single_stream_blocks is ignored when simple_dit=True. All blocks are single-stream DiT blocks.")
                n_single_stream_blocks = depth
        if rope_axes_dim is None:
            rope_axes_dim = _default_axes(simple_dit, heads_dim, partial_rotary_factor)
        for a in rope_axes_dim:
            assert a % 2 == 0, f"Each axis_dim must be even, got {a}"
        self.rope_axes_dim = list(rope_axes_dim)
        self.last_layer = ModulatedLastLayer(embedding_dim, inner_dim, patch_size, self.output_channels)
        self.time_embed = nn.Sequential(nn.Linear(frequency_embedding, embedding_dim), nn.SiLU(), nn.Linear(embedding_dim, embedding_dim))
        self.conv_proj = nn.Conv2d(input_channels, inner_dim, kernel_size=patch_size, stride=patch_size, bias=False)
        mk = dict(inner_dim=inner_dim, embedding_dim=embedding_dim, num_heads=num_heads, mlp_ratio=mlp_ratio,
                  rope_axes_dim=self.rope_axes_dim, use_checkpoint=use_checkpoint)
        # as in the reference (mmdit.py:701-733): with simple_dit the "single stream" tail is empty DiT blocks only
        n_main = depth - n_single_stream_blocks if not simple_dit else depth
        n_tail = n_single_stream_blocks if not simple_dit else 0
        self.layers = nn.ModuleList(
            [(MMDiTBlock(**mk) if not simple_dit else DiTBlock(**mk)) for _ in range(n_main)]
            + [MMDiTSingleStreamBlock(**mk) for _ in range(n_tail)]
        )
        self.apply(init_weights)

    def forward(
        self,
        x: Tensor,
        timesteps: Tensor,
        initial_context: Any | None = None,
        p: float = 0.0,
        y: Tensor | None = None,
        x_context: Tensor | None = None,
        intermediate_features: bool = False,
    ) -> ModelOutput:
        assert not (initial_context is not None and y is not None), "initial_context and y cannot both be specified"
        if p > 0:
            assert self.classifier_free, (
                "probability of dropping for classifier free guidance is only available if model is set up to be classifier free"
            )
        if x_context is not None:
            x = torch.cat([x, x_context], dim=1)
        B, _, H, W = x.shape
        ps = self.patch_size
        self.original_size, self.grid_size = (H, W), (H // ps, W // ps)
        tok = K.PatchEmbedFn.apply(x, self.conv_proj.weight, ps)
        if self.simple_dit and p > 0:
            assert self.n_classes, "probability of dropping for classifier free guidance is only available if a number of classes is set"
        cond_silu, _ = self._conditioning(timesteps, y, p)
        context, kmask = None, None
        if self.simple_dit:
            rope = rope_for(tok.device, 0, H // ps, W // ps, self.rope_axes_dim, self.rope_base, joint=False)
        else:
            context, kmask = self._context(initial_context, p, (cond_silu, _))
            rope = rope_for(tok.device, context.shape[1], H // ps, W // ps, self.rope_axes_dim, self.rope_base, joint=True)
        features: list[Tensor] | None = [] if intermediate_features else None
        tok, context = self._run_layers(self.layers, tok, cond_silu, context, rope, kmask, features)
        out = self.last_layer(tok, cond_silu, (H, W))
        model_output: ModelOutput = {"x": out}
        if features is not None:
            # NOTE: the reference never returns features here (empty-list truthiness bug);
            # this implementation returns them as documented.
            model_output["features"] = features
        return model_output
