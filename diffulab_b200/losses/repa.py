"""REPA loss: drop-in for reference training/losses/repa.py:24-186 (forward-hook capture of a block's output,
3-layer SiLU projector, coeff * (1 - mean cosine similarity) against precomputed encoder features).
Projector GEMMs run on the tcgen05 kernel; the cosine similarity + mean + its gradient are single fused kernels.
Frozen vision towers (DINOv2) are outside the hot path: `load_dino=True` is rejected, pass `dst_features`."""

from __future__ import annotations

from typing import Any

import torch
from torch import Tensor, nn
from torch.utils.hooks import RemovableHandle

from .. import blocks as K
from .. import ops
from .common import LossFunction


class _RepaProjCosFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feats: Tensor, dst: Tensor, coeff: float, w0, b0, w1, b1, w2, b2):
        f2 = feats.reshape(-1, feats.shape[-1])
        if not f2.is_contiguous():
            f2 = f2.contiguous()
        z1 = K.linear_fwd(f2, w0, b0)
        a1 = ops.silu_fwd(z1)
        z2 = K.linear_fwd(a1, w1, b1)
        a2 = ops.silu_fwd(z2)
        s = K.linear_fwd(a2, w2, b2)
        d2 = dst.reshape(-1, dst.shape[-1]).to(torch.float32).contiguous()
        loss = ops.repa_cos_fwd(s, d2, coeff)
        ctx.save_for_backward(f2, z1, a1, z2, a2, s, d2)
        ctx.p, ctx.coeff, ctx.fshape = (w0, b0, w1, b1, w2, b2), coeff, feats.shape
        return loss

    @staticmethod
    def backward(ctx, gout: Tensor):
        f2, z1, a1, z2, a2, s, d2 = ctx.saved_tensors
        w0, b0, w1, b1, w2, b2 = ctx.p
        ds = ops.repa_cos_bwd(s, d2, ctx.coeff, gout.to(torch.float32).contiguous())
        da2 = K.linear_bwd(ds, a2, w2, b2)
        dz2 = ops.silu_bwd(da2, z2, torch.bfloat16)
        da1 = K.linear_bwd(dz2, a1, w1, b1)
        dz1 = ops.silu_bwd(da1, z1, torch.bfloat16)
        df = K.linear_bwd(dz1, f2, w0, b0, need_dx=ctx.needs_input_grad[0])
        return (df.view(ctx.fshape) if df is not None else None), None, None, None, None, None, None, None, None


class _RepaProjFn(torch.autograd.Function):
    """The 3-layer SiLU projector alone (used when a PerceiverResampler sits between projector and cosine)."""

    @staticmethod
    def forward(ctx, feats: Tensor, w0, b0, w1, b1, w2, b2):
        f2 = feats.reshape(-1, feats.shape[-1])
        if not f2.is_contiguous():
            f2 = f2.contiguous()
        z1 = K.linear_fwd(f2, w0, b0)
        a1 = ops.silu_fwd(z1)
        z2 = K.linear_fwd(a1, w1, b1)
        a2 = ops.silu_fwd(z2)
        s = K.linear_fwd(a2, w2, b2)
        ctx.save_for_backward(f2, z1, a1, z2, a2)
        ctx.p, ctx.fshape = (w0, b0, w1, b1, w2, b2), feats.shape
        return s.view(*feats.shape[:-1], s.shape[-1])

    @staticmethod
    def backward(ctx, ds: Tensor):
        f2, z1, a1, z2, a2 = ctx.saved_tensors
        w0, b0, w1, b1, w2, b2 = ctx.p
        ds2 = ds.reshape(-1, ds.shape[-1]).contiguous()
        da2 = K.linear_bwd(ds2, a2, w2, b2)
        dz2 = ops.silu_bwd(da2, z2, torch.bfloat16)
        da1 = K.linear_bwd(dz2, a1, w1, b1)
        dz1 = ops.silu_bwd(da1, z1, torch.bfloat16)
        df = K.linear_bwd(dz1, f2, w0, b0, need_dx=ctx.needs_input_grad[0])
        return (df.view(ctx.fshape) if df is not None else None), None, None, None, None, None, None


class _CosFn(torch.autograd.Function):
    """coeff * (1 - mean cosine similarity) between bf16 projected features and fp32 targets (fused kernels)."""

    @staticmethod
    def forward(ctx, s: Tensor, dst: Tensor, coeff: float):
        s2 = s.reshape(-1, s.shape[-1]).contiguous()
        d2 = dst.reshape(-1, dst.shape[-1]).to(torch.float32).contiguous()
        ctx.save_for_backward(s2, d2)
        ctx.coeff, ctx.shape = coeff, s.shape
        return ops.repa_cos_fwd(s2, d2, coeff)

    @staticmethod
    def backward(ctx, gout: Tensor):
        s2, d2 = ctx.saved_tensors
        return ops.repa_cos_bwd(s2, d2, ctx.coeff, gout.to(torch.float32).contiguous()).view(ctx.shape), None, None


class RepaLoss(LossFunction):
    name: str = "RepaLoss"

    def __init__(
        self,
        repa_encoder: str = "dinov2",
        encoder_args: dict[str, Any] = {},
        alignment_layer: int = 8,
        denoiser_dimension: int = 256,
        hidden_dim: int = 1024,
        load_dino: bool = True,
        embedding_dim: int = 768,
        use_resampler: bool = False,
        resampler_params: dict[str, Any] | None = None,
        coeff: float = 1.0,
    ) -> None:
        super().__init__()
        if load_dino:
            raise NotImplementedError(
                "diffulab_b200.RepaLoss aligns against precomputed features (`dst_features`); the frozen DINO tower is "
                "outside the accelerated hot path (pass load_dino=False)"
            )
        self.repa_encoder = None
        self.proj = nn.Sequential(
            nn.Linear(denoiser_dimension, hidden_dim), nn.SiLU(), nn.Linear(hidden_dim, hidden_dim), nn.SiLU(),
            nn.Linear(hidden_dim, embedding_dim),
        )
        self.resampler = None
        if use_resampler:
            assert resampler_params is not None, "Resampler parameters must be provided when using the perceiver resampler."
            from .perceiver import PerceiverResampler

            self.resampler = PerceiverResampler(**resampler_params)
        self.alignment_layer = alignment_layer
        self._handles: dict[int, RemovableHandle] = {}
        self._captured_features: dict[int, Tensor] = {}
        self._active_model_id: int | None = None
        self._hook_layer_idx = self.alignment_layer - 1
        self.coeff = coeff

    def _make_hook(self, model_id: int):
        def _hook(_mod: nn.Module, _inp: tuple[Any, ...], out: Any) -> None:
            self._captured_features[model_id] = out

        return _hook

    def set_model(self, model: nn.Module) -> None:
        model_id = id(model)
        if model_id not in self._handles:
            layer = model.layers[self._hook_layer_idx]
            self._handles[model_id] = layer.register_forward_hook(self._make_hook(model_id))
        self._active_model_id = model_id

    def _unregister_all(self) -> None:
        for handle in self._handles.values():
            handle.remove()
        self._handles.clear()
        self._captured_features.clear()
        self._active_model_id = None

    def forward(self, x0: Tensor | None = None, dst_features: Tensor | None = None) -> Tensor:
        if self._active_model_id is None or self._active_model_id not in self._captured_features:
            raise RuntimeError("REPA: no captured features for the active model. Did you call set_model(...) and run a forward pass?")
        assert dst_features is not None, "dst_features must be provided (the DINO tower is not instantiated)."
        src = self._captured_features[self._active_model_id]
        if isinstance(src, tuple):
            src = src[0]
        p = self.proj
        if self.resampler is not None:  # proj -> PerceiverResampler -> cosine (reference repa.py:180-186)
            s = _RepaProjFn.apply(src, p[0].weight, p[0].bias, p[2].weight, p[2].bias, p[4].weight, p[4].bias)
            s = self.resampler(s)
            return _CosFn.apply(s, dst_features, float(self.coeff))
        return _RepaProjCosFn.apply(src, dst_features, float(self.coeff), p[0].weight, p[0].bias, p[2].weight, p[2].bias,
                                    p[4].weight, p[4].bias)
