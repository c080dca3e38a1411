"""DDT (decoupled diffusion transformer): drop-in for reference networks/denoisers/ddt.py:19-512. Encoder of
MMDiT / DiT blocks on `conv_proj_encoder(x)`; decoder of DiT blocks on `conv_proj_decoder(x)` whose adaLN
modulation is PER TOKEN: conditioning silu(encoder_output + time_emb) (ddt.py:421-422), so each decoder
modulation is a real [B*N, d] x [6d, d] GEMM."""

from __future__ import annotations

import logging
from typing import Any

import torch
import torch.nn as nn
from torch import Tensor

from .. import blocks as K
from ..embedders.common import ContextEmbedder
from .common import ModelOutput
from .layers import DiTBlock, LabelEmbed, MMDiTBlock, MMDiTSingleStreamBlock, ModulatedLastLayer, init_weights, rope_for
from .mmdit import _DenoiserBase, _default_axes


class DDT(_DenoiserBase):
    def __init__(
        self,
        simple_ddt: bool = False,
        input_channels: int = 3,
        output_channels: int | None = None,
        inner_dim: int = 768,
        num_heads: int = 12,
        mlp_ratio: int = 4,
        patch_size: int = 16,
        encoder_depth: int = 8,
        n_single_stream_blocks: int = 0,
        decoder_depth: int = 4,
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
        assert n_single_stream_blocks < encoder_depth, "n_single_stream_blocks must be less than encoder_depth"
        self.simple_ddt = self.simple = simple_ddt
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
        if not simple_ddt:
            assert context_embedder is not None, "for ddt with text context embedder must be provided"
            self._setup_mm(context_embedder, inner_dim, inner_dim)
        else:
            self.label_embed = LabelEmbed(n_classes, inner_dim, classifier_free) if n_classes is not None else None
            if n_single_stream_blocks > 0:
                logging.warning("n_single_stream_blocks is ignored when simple_ddt=True. All blocks are single-stream DiT blocks.")
                n_single_stream_blocks = 0
        if rope_axes_dim is None:
            rope_axes_dim = _default_axes(simple_ddt, heads_dim, partial_rotary_factor)
        self.rope_axes_dim = list(rope_axes_dim)
        self.last_layer = ModulatedLastLayer(inner_dim, inner_dim, patch_size, self.output_channels)
        self.time_embed = nn.Sequential(nn.Linear(frequency_embedding, inner_dim), nn.SiLU(), nn.Linear(inner_dim, inner_dim))
        self.conv_proj_encoder = nn.Conv2d(input_channels, inner_dim, kernel_size=patch_size, stride=patch_size, bias=False)
        self.conv_proj_decoder = nn.Conv2d(input_channels, inner_dim, kernel_size=patch_size, stride=patch_size, bias=False)
        mk = dict(inner_dim=inner_dim, embedding_dim=inner_dim, num_heads=num_heads, mlp_ratio=mlp_ratio,
                  rope_axes_dim=self.rope_axes_dim, use_checkpoint=use_checkpoint)
        self.layers = nn.ModuleList(
            [(MMDiTBlock(**mk) if not simple_ddt else DiTBlock(**mk)) for _ in range(encoder_depth - n_single_stream_blocks)]
            + [MMDiTSingleStreamBlock(**mk) for _ in range(n_single_stream_blocks)]
        )
        self.decoder_layers = nn.ModuleList([DiTBlock(**mk) for _ in range(decoder_depth)])
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
        hp, wp = H // ps, W // ps
        self.original_size, self.grid_size = (H, W), (hp, wp)
        tok = K.PatchEmbedFn.apply(x, self.conv_proj_encoder.weight, ps)
        cond_silu, te = self._conditioning(timesteps, y, p)
        context, kmask, L = None, None, 0
        if self.simple_ddt:
            rope = rope_for(tok.device, 0, hp, wp, self.rope_axes_dim, self.rope_base, joint=False)
        else:
            context, kmask = self._context(initial_context, p, (cond_silu, te))
            L = context.shape[1]
            rope = rope_for(tok.device, L, hp, wp, self.rope_axes_dim, self.rope_base, joint=True)
        features: list[Tensor] | None = [] if intermediate_features else None
        enc, _ = self._run_layers(self.layers, tok, cond_silu, context, rope, kmask, features)
        # decoder conditioning: silu(enc + time_emb) per token, then the Modulation's own silu (nn.py:531)
        cond_tok = K.SiluFn.apply(K.BiasSiluFn.apply(enc, te))
        z = K.PatchEmbedFn.apply(x, self.conv_proj_decoder.weight, ps)
        for layer in self.decoder_layers:
            z = layer(z, cond_tok, rope, None, L)  # image rows (0,h,w) start at table row L (ddt.py:425-449)
            if features is not None:
                features.append(z)
        out = self.last_layer(z, cond_tok, (H, W))
        model_output: ModelOutput = {"x": out}
        if features is not None:
            model_output["features"] = features
        return model_output
