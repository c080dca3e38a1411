"""Euler-Maruyama flow sampler (reference diffuse/samplers/flow/euler_meruyama.py:8-57): stochastic step with the
per-element log-probability needed by the GRPO-style objectives, as ONE fused kernel (dlb_euler_maruyama_step). The scalar
schedule terms are evaluated in Python floats exactly as the reference does; the noise is drawn with torch.randn_like at
the same point, so seeded runs consume the same Philox stream."""

from __future__ import annotations

import torch
from torch import Tensor

from ... import ops
from .common import FlowSampler, StepResult


class EulerMaruyama(FlowSampler):
    name = "euler_maruyama"

    def __init__(self, eta: float = 0.7) -> None:
        super().__init__()
        self.eta = eta
        self.tmax: float | None = None

    def set_steps(self, timesteps: list[float]) -> None:
        self.tmax = timesteps[1]

    def step(self, x_t: Tensor, v: Tensor, t_curr: float, t_prev: float, x_prev: Tensor | None = None) -> StepResult:
        assert self.tmax is not None, "set_steps must be called before step"
        sigma: float = ((t_curr / (1 - min(t_curr, self.tmax))) ** 0.5) * self.eta
        std = sigma * (t_curr - t_prev) ** 0.5
        x = (x_t if x_t.dtype == torch.float32 else x_t.float()).contiguous()
        noise = torch.randn_like(x) if x_prev is None else None
        given = None if x_prev is None else x_prev.detach().float().contiguous()
        xp, mean, x0, logprob = ops.euler_maruyama_step(x, v, noise, given, sigma**2 / (2 * t_curr), 1 - t_curr, t_curr - t_prev, t_curr, std)
        return StepResult(x_prev=xp if x_prev is None else x_prev, x_prev_mean=mean, x_prev_std=torch.tensor(std, device=x.device),
                          estimated_x0=x0, logprob=logprob)
