"""Euler flow sampler (reference diffuse/samplers/flow/euler.py:8-41) as one fused kernel: x_prev = x_t - v*dt and
estimated_x0 = x_t - v*t_curr in a single pass; optionally folds the classifier-free-guidance combine
v = v_u + g (v_c - v_u) (reference flow.py:259) into the same launch via `step_cfg`."""

from __future__ import annotations

import torch
from torch import Tensor

from ... import ops
from .common import FlowSampler, StepResult


class Euler(FlowSampler):
    name = "euler"

    def __init__(self) -> None:
        super().__init__()

    def set_steps(self, timesteps: list[float]) -> None:
        pass

    def step(self, x_t: Tensor, v: Tensor, t_curr: float, t_prev: float) -> StepResult:
        return self.step_cfg(x_t, v, None, 0.0, t_curr, t_prev)

    def step_cfg(self, x_t: Tensor, v_cond: Tensor, v_uncond: Tensor | None, guidance_scale: float, t_curr: float,
                 t_prev: float) -> StepResult:
        x = x_t if x_t.dtype == torch.float32 else x_t.float()
        x_prev, x0 = ops.euler_step(x.contiguous(), v_cond.contiguous(), v_uncond.contiguous() if v_uncond is not None else None,
                                    float(guidance_scale), float(t_curr), float(t_prev), want_x0=True)
        return StepResult(x_prev=x_prev, estimated_x0=x0)
