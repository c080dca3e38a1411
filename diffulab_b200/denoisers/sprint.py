"""SprintDiT: drop-in for reference networks/denoisers/sprint.py:19-624 (encoder -> token drop -> deep layers on
the kept tokens -> restore with mask token -> fuse -> decoder), executed by hand-written sm_100a kernels.
Token selection consumes the same `torch.rand((B, S))` draw as the reference (sprint.py:343) and the kept
indices are bit-exact w.r.t. `topk(sorted=False)` + `argsort` on tie-free draws (ties -> larger index)."""

from __future__ import annotations

import logging
from typing import Any

import torch
import torch.nn as nn
from torch import Tensor

from .. import blocks as K
from .. import ops
from ..embedders.common import ContextEmbedder
from .common import ModelOutput
from .layers import DiTBlock, LabelEmbed, MMDiTBlock, MMDiTSingleStreamBlock, ModulatedLastLayer, init_weights, rope_for
from .mmdit import _DenoiserBase, _default_axes


class SprintDiT(_DenoiserBase):
    cfg_batchable = False  # p = 1 also skips the deep layers (path-drop guidance, sprint.py:474-475)

    def __init__(
        self,
        simple_dit: bool = False,
        input_channels: int = 3,
        output_channels: int | None = None,
        inner_dim: int = 768,
        embedding_dim: int = 768,
        num_heads: int = 12,
        mlp_ratio: int = 4,
        patch_size: int = 16,
        encoder_depth: int = 2,
        deep_layers_depth: int = 8,
        n_single_stream_blocks: int = 0,
        decoder_depth: int = 2,
        rope_base: int = 10_000,
        partial_rotary_factor: float = 1,
        rope_axes_dim: list[int] | None = None,
        frequency_embedding: int = 256,
        n_classes: int | None = None,
        classifier_free: bool = False,
        context_embedder: ContextEmbedder | None = None,
        use_checkpoint: bool = False,
        drop_rate: float = 0.75,
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
        self.mask_token = nn.Parameter(torch.zeros(1, 1, inner_dim))
        self.drop_rate = drop_rate
        self.inner_dim, self.num_heads = inner_dim, num_heads
        heads_dim = inner_dim // num_heads
        assert heads_dim % 8 == 0 and heads_dim <= 128, "head_dim must be a multiple of 8 and <= 128"
        if not simple_dit:
            assert context_embedder is not None, "for dit with text context embedder must be provided"
            self._setup_mm(context_embedder, inner_dim, embedding_dim)
        else:
            self.label_embed = LabelEmbed(n_classes, embedding_dim, classifier_free) if n_classes is not None else None
            if n_single_stream_blocks > 0:
                logging.warning("n_single_stream_blocks is ignored when simple_dit=True. All blocks are single-stream DiT blocks.")
                n_single_stream_blocks = 0
        if rope_axes_dim is None:
            rope_axes_dim = _default_axes(simple_dit, heads_dim, partial_rotary_factor)
        self.rope_axes_dim = list(rope_axes_dim)
        self.time_embed = nn.Sequential(nn.Linear(frequency_embedding, embedding_dim), nn.SiLU(), nn.Linear(embedding_dim, embedding_dim))
        self.conv_proj = nn.Conv2d(input_channels, inner_dim, kernel_size=patch_size, stride=patch_size, bias=False)
        self.fuse = nn.Linear(inner_dim * 2, inner_dim, bias=False)
        if not simple_dit:
            self.fuse_context = nn.Linear(2 * inner_dim, inner_dim, bias=False)
        self.last_layer = ModulatedLastLayer(embedding_dim, inner_dim, patch_size, self.output_channels)
        mk = dict(inner_dim=inner_dim, embedding_dim=embedding_dim, num_heads=num_heads, mlp_ratio=mlp_ratio,
                  rope_axes_dim=self.rope_axes_dim, use_checkpoint=use_checkpoint)

        def main_block():
            return MMDiTBlock(**mk) if not simple_dit else DiTBlock(**mk)

        self.layers = nn.ModuleList([main_block() for _ in range(encoder_depth)])  # name kept for REPA (sprint.py:178)
        self.deep_layers = nn.ModuleList(
            [main_block() for _ in range(deep_layers_depth - n_single_stream_blocks)]
            + [MMDiTSingleStreamBlock(**mk) for _ in range(n_single_stream_blocks)]
        )
        self.decoder_layers = nn.ModuleList([main_block() for _ in range(decoder_depth)])
        self.apply(init_weights)

    # -- token drop / restore (sprint.py:317-387) ------------------------------------------------------------
    def drop_tokens(self, x: Tensor) -> tuple[Tensor, Tensor, Tensor | None, Tensor | None]:
        """Returns (x_kept, kept_indices int64 [B,k], kept32 int32 [B,k] or None, inv int32 [B,S] or None)."""
        B, S, _ = x.shape
        if not self.training:
            return x, torch.arange(S, device=x.device).expand(B, S), None, None
        k = max(1, int(S * (1.0 - float(self.drop_rate))))
        scores = torch.rand((B, S), device=x.device, dtype=torch.float32)  # same draw as the reference
        kept, kept32, inv = ops.sprint_select(scores, k)
        return K.GatherTokensFn.apply(x, kept, inv), kept, kept32, inv

    def restore_tokens(self, x_kept: Tensor, kept: Tensor, inv: Tensor | None, path_drop_p: float = 0.0) -> Tensor:
        B, S = x_kept.shape[0], self.grid_size[0] * self.grid_size[1]
        drop = None
        if path_drop_p > 0:
            drop = (torch.rand(B, device=x_kept.device) < path_drop_p).to(torch.uint8)
        if inv is None:  # eval: every token kept, identity scatter
            if drop is None:
                return x_kept
            inv = torch.arange(S, device=x_kept.device, dtype=torch.int32).expand(B, S).contiguous()
            kept = kept.contiguous()
        return K.RestoreTokensFn.apply(x_kept, self.mask_token, kept.contiguous(), inv, drop)

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
        tok = K.PatchEmbedFn.apply(x, self.conv_proj.weight, ps)
        cond_silu, _ = self._conditioning(timesteps, y, p)
        context, kmask, L = None, None, 0
        if self.simple_dit:
            rope = rope_for(tok.device, 0, hp, wp, self.rope_axes_dim, self.rope_base, joint=False)
        else:
            context, kmask = self._context(initial_context, p, (cond_silu, _))
            L = context.shape[1]
            rope = rope_for(tok.device, L, hp, wp, self.rope_axes_dim, self.rope_base, joint=True)
        features: list[Tensor] | None = [] if intermediate_features else None
        tok, context = self._run_layers(self.layers, tok, cond_silu, context, rope, kmask, features)
        enc_context = context

        x_kept, kept, kept32, inv = self.drop_tokens(tok)
        # RoPE rows of the kept image tokens: table row = L + original token index (sprint.py:460-465)
        pos_idx = (kept32 + L).reshape(-1).contiguous() if kept32 is not None else None
        if p < 1:
            x_kept, context = self._run_layers(self.deep_layers, x_kept, cond_silu, context, rope, kmask, features, pos_idx)
            restored = self.restore_tokens(x_kept, kept, inv, p)
        else:
            restored = K.MaskFillFn.apply(self.mask_token, B, hp * wp)
        fused = K.linear(torch.cat([restored, tok], dim=-1), self.fuse.weight, None)
        if context is not None:
            context = K.linear(torch.cat([context, enc_context], dim=-1), self.fuse_context.weight, None)
        fused, context = self._run_layers(self.decoder_layers, fused, cond_silu, context, rope, kmask, features)
        out = self.last_layer(fused, cond_silu, (H, W))
        model_output: ModelOutput = {"x": out}
        if features is not None:
            # the reference appends the post-last_layer token tensor [B, N, p*p*C] as the final feature (sprint.py:492-494,
            # 574-576); the fused last layer returns the image, so the token view is rebuilt (layout plumbing only)
            C_out = self.output_channels
            features.append(out.view(B, C_out, hp, ps, wp, ps).permute(0, 2, 4, 3, 5, 1).reshape(B, hp * wp, ps * ps * C_out))
            model_output["features"] = features
        return model_output
