"""Heun (2nd-order, trapezoidal) flow sampler. The reference has no Heun sampler (its registry is
{"euler", "euler_maruyama"}, diffuse/modelizations/flow.py:54-57); BASELINE.json's north star names a fused
Euler/Heun update, so it is defined here as the textbook predictor-corrector over two reference `get_v` evaluations
(SURVEY.md 8(f)-2):

    x_pred = x_t - v1 * (t_curr - t_prev)                      (Euler predictor, v1 = v(x_t, t_curr))
    x_prev = x_t - 0.5 * (v1 + v2) * (t_curr - t_prev)         (corrector, v2 = v(x_pred, t_prev))

and, as in EDM, the last step (t_prev == 0) is a plain Euler step. Both halves are single fused launches of
`dlb_euler_step` (the corrector is its guidance form v1 + 0.5 (v2 - v1)); `Flow.one_step_denoise` drives the second
model evaluation. `step()` alone (one velocity) is the Euler step, so the FlowSampler interface still holds."""

from __future__ import annotations

import torch
from torch import Tensor

from ... import ops
from .common import FlowSampler, StepResult


class Heun(FlowSampler):
    name = "heun"
    second_order = True

    def __init__(self) -> None:
        super().__init__()

    def set_steps(self, timesteps: list[float]) -> None:
        pass

    def needs_corrector(self, t_prev: float) -> bool:
        return t_prev > 0.0

    @staticmethod
    def _f32(x: Tensor) -> Tensor:
        return (x if x.dtype == torch.float32 else x.float()).contiguous()

    def step(self, x_t: Tensor, v: Tensor, t_curr: float, t_prev: float) -> StepResult:
        return self.step_cfg(x_t, v, None, 0.0, t_curr, t_prev)

    def step_cfg(self, x_t: Tensor, v_cond: Tensor, v_uncond: Tensor | None, guidance_scale: float, t_curr: float,
                 t_prev: float) -> StepResult:
        x_prev, x0 = ops.euler_step(self._f32(x_t), v_cond.contiguous(), v_uncond.contiguous() if v_uncond is not None else None,
                                    float(guidance_scale), float(t_curr), float(t_prev), want_x0=True)
        return StepResult(x_prev=x_prev, estimated_x0=x0)

    def predict_cfg(self, x_t: Tensor, v_cond: Tensor, v_uncond: Tensor | None, guidance_scale: float, t_curr: float,
                    t_prev: float) -> tuple[Tensor, Tensor]:
        """-> (x_pred, v1): Euler predictor; v1 = v_cond, or the guided combination v_u + g (v_c - v_u) (fp32)."""
        if v_uncond is None:
            x_pred, _ = ops.euler_step(self._f32(x_t), v_cond.contiguous(), None, 0.0, float(t_curr), float(t_prev), want_x0=False)
            return x_pred, v_cond
        x_pred, _, v1 = ops.euler_step(self._f32(x_t), v_cond.contiguous(), v_uncond.contiguous(), float(guidance_scale), float(t_curr),
                                       float(t_prev), want_x0=False, want_v=True)
        return x_pred, v1

    def combine(self, x_like: Tensor, v_cond: Tensor, v_uncond: Tensor, guidance_scale: float) -> Tensor:
        """v_u + g (v_c - v_u) in fp32 (the update outputs of the launch are discarded: dt = 0)."""
        _, _, v = ops.euler_step(self._f32(x_like), v_cond.contiguous(), v_uncond.contiguous(), float(guidance_scale), 0.0, 0.0,
                                 want_x0=False, want_v=True)
        return v

    def correct(self, x_t: Tensor, v1: Tensor, v2: Tensor, t_curr: float, t_prev: float) -> StepResult:
        v1, v2 = v1.contiguous(), v2.contiguous()
        if v1.dtype != v2.dtype:
            v1, v2 = v1.float(), v2.float()
        # v1 + 0.5 (v2 - v1) = (v1 + v2) / 2 combined with the update in one launch
        x_prev, x0 = ops.euler_step(self._f32(x_t), v2, v1, 0.5, float(t_curr), float(t_prev), want_x0=True)
        return StepResult(x_prev=x_prev, estimated_x0=x0)
