"""DDPM / DDIM samplers: drop-ins for reference diffuse/samplers/gaussian_diffusion/{common,ddpm,ddim}.py (same
constructor arguments, `set_steps(betas)`, `step(model_prediction, timesteps, xt, clamp_x[, eta]) -> StepResult`).

`set_steps` builds the reference's float64 schedule tables and, from them, ONE float32 coefficient table with a row
per timestep — each entry computed in float32 from the float64 table value exactly as the reference does per sample
after `extract_into_tensor` (diffuse/utils.py:6-19). `step` is then a single fused CUDA launch (dlb_gaussian_step):
x0 from the model output, optional clamp, posterior mean, x_{t-1} = mean + [t>0] * std * noise and the log-probability.
The noise is drawn with `torch.randn_like` at the same point as in the reference, so seeded runs consume the same
Philox stream. Learned-variance parameterisations ("learned", "learned_range", ddpm.py:213-223): the model prediction
holds 2C channels [mean prediction | variance head]; the same launch reads both halves in place (no torch.chunk copies)
and also writes the per-element standard deviation. DDIM ignores the variance head, as the reference does.
"""

from __future__ import annotations

from typing import Any

import torch
from torch import Tensor

from ... import ops
from .common import Sampler, StepResult

_MEAN_TYPES = {"epsilon": 0, "xstart": 1, "xprev": 2}
_VAR_TYPES = ("learned", "fixed_small", "fixed_large", "learned_range")
# columns of the per-timestep table (csrc/diffusion_ops.cu GS_*)
_COLS = 16


class GaussianSampler(Sampler):
    name: str

    def set_steps(self, betas: Tensor) -> None:  # pragma: no cover - interface
        raise NotImplementedError

    def step(self, model_prediction: Tensor, timesteps: Tensor, xt: Tensor, clamp_x: bool = False, *args: Any, **kwargs: Any) -> StepResult:  # pragma: no cover
        raise NotImplementedError


class DDPM(GaussianSampler):
    name = "ddpm"
    _sampler_id = 0

    def __init__(self, mean_type: str = "epsilon", var_type: str = "fixed_small") -> None:
        super().__init__()
        if mean_type not in _MEAN_TYPES:
            raise ValueError(f"mean_type must be one of {list(_MEAN_TYPES)}")
        if var_type not in _VAR_TYPES:
            raise ValueError(f"variance_type must be one of {list(_VAR_TYPES)}")
        self.mean_type = mean_type
        self.var_type = var_type
        self._dev_tables: dict[torch.device, tuple[Tensor, Tensor]] = {}

    def set_steps(self, betas: Tensor) -> None:
        """Reference ddpm.py `set_steps` (float64 tables) + the derived float32 per-timestep coefficient table."""
        one = torch.ones_like
        self.betas = betas
        self.alphas = one(betas) - betas
        self.alphas_bar = self.alphas.cumprod(dim=0)
        self.alphas_bar_prev = torch.cat([torch.tensor([1.0], dtype=torch.float64), self.alphas_bar[:-1]])
        self.alphas_bar_next = torch.cat([self.alphas_bar[1:], torch.tensor([0.0], dtype=torch.float64)])
        self.sqrt_alphas_bar = self.alphas_bar.sqrt()
        self.posterior_variance = betas * (one(self.alphas_bar_prev) - self.alphas_bar_prev) / (one(self.alphas_bar) - self.alphas_bar)
        self.posterior_log_variance_clipped = torch.log(torch.cat([self.posterior_variance[1:2], self.posterior_variance[1:]]))
        self.posterior_mean_coef1 = betas * self.alphas_bar_prev.sqrt() / (one(self.alphas_bar) - self.alphas_bar)
        self.posterior_mean_coef2 = (one(self.alphas_bar_prev) - self.alphas_bar_prev) * self.alphas.sqrt() / (one(self.alphas_bar) - self.alphas_bar)

        f = lambda a: a.float()  # noqa: E731  (what extract_into_tensor hands to the fp32 arithmetic)
        ab, sab, abp = f(self.alphas_bar), f(self.sqrt_alphas_bar), f(self.alphas_bar_prev)
        c1, c2 = f(self.posterior_mean_coef1), f(self.posterior_mean_coef2)
        if self.var_type != "fixed_large":  # learned types: columns 6 / 7 are unused by the kernel
            var, lv = f(self.posterior_variance), f(self.posterior_log_variance_clipped)
        else:
            seq = torch.cat([self.posterior_variance[1:2], self.betas[1:]])
            var, lv = f(seq), f(torch.log(seq))
        n = betas.numel()
        tab = torch.zeros(n, _COLS, dtype=torch.float32)
        tab[:, 0] = 1.0 / sab                                 # x0 <- eps : (1/sqrt_ab) * xt
        tab[:, 1] = (torch.ones_like(ab) - ab).sqrt() / sab   #            - (sqrt(1-ab)/sqrt_ab) * eps
        tab[:, 2] = 1.0 / c1                                  # x0 <- xprev
        tab[:, 3] = c2 / c1
        tab[:, 4] = c1
        tab[:, 5] = c2
        tab[:, 6] = var
        tab[:, 7] = torch.exp(0.5 * lv)
        tab[:, 8] = (torch.arange(n) > 0).float()
        tab[:, 9] = (1 / ab - 1).sqrt()                        # eps <- x0 denominator (ddpm.py `_get_eps_from_xstart`)
        tab[:, 10] = abp.sqrt()
        tab[:, 11] = ((torch.ones_like(abp) - abp) / (torch.ones_like(ab) - ab)).sqrt()   # DDIM sigma = eta * [11] * [12]
        tab[:, 12] = (torch.ones_like(ab) - ab / abp).sqrt()
        tab[:, 13] = abp
        tab[:, 14] = f(self.posterior_log_variance_clipped)    # learned_range: min_log
        tab[:, 15] = f(self.betas).log()                       #                max_log (fp32 log of the fp32 beta, ddpm.py:219)
        self._table = tab
        self._std = var.clamp_min(1e-20).sqrt()               # x_prev_std of DDPM.step
        self._dev_tables = {}

    def _tables(self, device: torch.device) -> tuple[Tensor, Tensor]:
        if device not in self._dev_tables:
            self._dev_tables[device] = (self._table.to(device).contiguous(), self._std.to(device))
        return self._dev_tables[device]

    def _run(self, model_prediction: Tensor, timesteps: Tensor, xt: Tensor, clamp_x: bool, eta: float, want_logprob: bool):
        x = (xt if xt.dtype == torch.float32 else xt.float()).contiguous()
        table, std = self._tables(x.device)
        t = timesteps.to(device=x.device, dtype=torch.int32).contiguous()
        noise = torch.randn_like(x)
        var_mode = {"learned": 1, "learned_range": 2}.get(self.var_type, 0)
        if var_mode and (model_prediction.dim() < 2 or model_prediction.shape[1] != 2 * x.shape[1]):
            raise ValueError(f"var_type {self.var_type!r}: the model must emit 2 x {x.shape[1]} channels, got {tuple(model_prediction.shape)}")
        *out, std_el = ops.gaussian_step(model_prediction, x, noise, table, t, self._sampler_id, _MEAN_TYPES[self.mean_type], clamp_x, eta,
                                         want_logprob, var_mode)
        return out, t, (std, std_el)

    def step(self, model_prediction: Tensor, timesteps: Tensor, xt: Tensor, clamp_x: bool = False) -> StepResult:
        (x_prev, x0, mean, logprob), t, (std, std_el) = self._run(model_prediction, timesteps, xt, clamp_x, 0.0, True)
        shape = (-1,) + (1,) * (xt.dim() - 1)
        assert logprob is not None
        x_prev_std = std_el if std_el is not None else std[t.long()].view(shape).expand_as(x_prev)
        return StepResult(x_prev=x_prev, estimated_x0=x0, x_prev_mean=mean, x_prev_std=x_prev_std, logprob=logprob)


class DDIM(DDPM):
    name = "ddim"
    _sampler_id = 1

    def step(self, model_prediction: Tensor, timesteps: Tensor, xt: Tensor, clamp_x: bool = False, eta: float = 0.0) -> StepResult:
        (x_prev, x0, mean, logprob), t, _ = self._run(model_prediction, timesteps, xt, clamp_x, float(eta), eta > 0)
        out = StepResult(x_prev=x_prev, estimated_x0=x0, x_prev_mean=mean)
        if eta > 0:
            table, _ = self._tables(x_prev.device)
            rows = table[t.long()]
            sigma = (eta * rows[:, 11] * rows[:, 12]).view((-1,) + (1,) * (xt.dim() - 1)).expand_as(x_prev)
            out["x_prev_std"] = sigma
            assert logprob is not None
            out["logprob"] = logprob
        return out
