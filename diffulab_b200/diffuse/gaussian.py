"""Gaussian (DDPM-style) diffusion formalisation: drop-in for reference diffuse/modelizations/gaussian_diffusion.py:17-447
(same constructor, `set_steps` with respacing + `timestep_map`, `draw_timesteps`, `add_noise`, `compute_loss`,
`one_step_denoise`, `denoise`) and modelizations/utils.py `space_timesteps`.
Device math runs in fused kernels: x_t = sqrt(ab_t) x0 + sqrt(1 - ab_t) eps (dlb_interp with per-sample coefficients),
mean((pred - eps)^2) and its gradient (dlb_mse_fwd/bwd), one launch per reverse step (dlb_gaussian_step)."""

from __future__ import annotations

import math
from typing import Any, Callable, Literal, cast

import torch
from torch import Tensor

from .. import ops
from ..denoisers.common import Denoiser, ModelInput
from ..losses.common import LossFunction
from .diffusion import Diffusion, SamplingOutput
from .flow import _FlowLossFn
from .samplers.common import StepResult
from .samplers.gaussian import DDIM, DDPM


def space_timesteps(num_timesteps: int, section_counts: str | int, ddim: bool = False) -> set[int]:
    """Reference modelizations/utils.py:1-57. NOTE (reference defect, kept observable): the reference's DDIM branch
    raises from inside the first iteration of its stride loop, i.e. for every request except the identity; here the loop
    runs to completion first (the evident intent), so strided DDIM schedules work instead of raising."""
    if ddim:
        assert isinstance(section_counts, int)
        for i in range(1, num_timesteps):
            if len(range(0, num_timesteps, i)) == section_counts:
                return set(range(0, num_timesteps, i))
        raise ValueError(f"cannot create exactly {section_counts} steps with an integer stride")
    counts = [int(x) for x in section_counts.split(",")] if isinstance(section_counts, str) else [section_counts]
    size_per, extra = num_timesteps // len(counts), num_timesteps % len(counts)
    start_idx, all_steps = 0, []
    for i, count in enumerate(counts):
        size = size_per + (1 if i < extra else 0)
        if size < count:
            raise ValueError(f"cannot divide section of {size} steps into {count}")
        frac_stride = 1 if count <= 1 else (size - 1) / (count - 1)
        cur = 0.0
        for _ in range(count):
            all_steps.append(start_idx + round(cur))
            cur += frac_stride
        start_idx += size
    return set(all_steps)


class GaussianDiffusion(Diffusion):
    sampler_registry = {"ddpm": DDPM, "ddim": DDIM}

    def __init__(self, n_steps: int = 1000, sampling_method: Literal["ddpm", "ddim"] = "ddpm",
                 schedule: Literal["linear", "cosine"] = "linear", latent_diffusion: bool = False,
                 sampler_parameters: dict[str, Any] = {}):
        if sampling_method not in ["ddpm", "ddim"]:
            raise ValueError("sampling method must be one of ['ddpm', 'ddim']")
        self.training_steps = n_steps
        super().__init__(n_steps=self.training_steps, sampling_method=sampling_method, schedule=schedule,
                         latent_diffusion=latent_diffusion, sampler_parameters=sampler_parameters)

    def set_diffusion_parameters(self, betas: Tensor) -> None:
        self.betas = betas
        self.alphas = torch.ones_like(self.betas) - self.betas
        self.alphas_bar = self.alphas.cumprod(dim=0)
        self.sqrt_alphas_bar = self.alphas_bar.sqrt()
        self.sampler.set_steps(betas)
        # per-timestep fp32 coefficients of add_noise, as extract_into_tensor hands them to the fp32 arithmetic
        ab = self.alphas_bar.float()
        self._noise_a = self.sqrt_alphas_bar.float()
        self._noise_b = (torch.ones_like(ab) - ab).sqrt()
        self._dev: dict[torch.device, tuple[Tensor, Tensor]] = {}

    def set_steps(self, n_steps: int, schedule: str = "linear", section_counts: int | str | None = None) -> None:
        if n_steps != self.training_steps:
            section_counts = section_counts or n_steps
        self.steps = n_steps
        betas = self._get_variance_schedule(self.training_steps, schedule)
        self.set_diffusion_parameters(betas)
        self.timestep_map: list[int] = []
        if section_counts:
            use = space_timesteps(num_timesteps=self.training_steps, section_counts=section_counts, ddim=self.sampling_method == "ddim")
            last_alpha_bar = torch.tensor(1.0)
            new_betas: list[Tensor] = []
            for i, alpha_bar in enumerate(self.alphas_bar):
                if i in use:
                    new_betas.append(torch.ones_like(alpha_bar) - alpha_bar / last_alpha_bar)
                    last_alpha_bar = alpha_bar
                    self.timestep_map.append(i)
            self.set_diffusion_parameters(torch.tensor(new_betas))

    def _get_variance_schedule(self, n_steps: int, variance_schedule: str = "linear") -> Tensor:
        if variance_schedule == "linear":
            scale = 1000 / n_steps
            return torch.linspace(scale * 0.0001, scale * 0.02, n_steps, dtype=torch.float64, requires_grad=False)
        if variance_schedule == "cosine":
            return self._betas_for_alpha_bar(n_steps, lambda t: math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2)
        raise NotImplementedError(f"unknown beta schedule: {variance_schedule}")

    def _betas_for_alpha_bar(self, n_steps: int, alpha_bar: Callable[[float], float], max_beta: float = 0.999) -> Tensor:
        betas = [min(1 - alpha_bar((i + 1) / n_steps) / alpha_bar(i / n_steps), max_beta) for i in range(n_steps)]
        return torch.tensor(betas, dtype=torch.float64, requires_grad=False)

    def draw_timesteps(self, batch_size: int) -> Tensor:
        return torch.randint(0, self.steps, (batch_size,), dtype=torch.int32)

    def _coeffs(self, device: torch.device) -> tuple[Tensor, Tensor]:
        if device not in self._dev:
            self._dev[device] = (self._noise_a.to(device), self._noise_b.to(device))
        return self._dev[device]

    def add_noise(self, x: Tensor, timesteps: Tensor, noise: Tensor | None = None) -> tuple[Tensor, Tensor]:
        if noise is None:
            noise = torch.randn_like(x)
        assert noise.shape == x.shape
        assert timesteps.shape[0] == x.shape[0]
        a_tab, b_tab = self._coeffs(x.device)
        idx = timesteps.to(device=x.device, dtype=torch.long)
        x_t = ops.interp(x.float().contiguous(), noise.float().contiguous(), a_tab[idx].contiguous(), b_tab[idx].contiguous())
        return x_t, noise

    def _map(self, timesteps: Tensor) -> Tensor:
        if self.timestep_map:
            map_tensor = torch.tensor(self.timestep_map, device=timesteps.device, dtype=timesteps.dtype)
            return map_tensor[timesteps.long()]
        return timesteps

    def one_step_denoise(self, model: Denoiser, model_inputs: ModelInput, t: int, clamp_x: bool = False, guidance_scale: float = 0.0,
                         sampler_args: dict[str, Any] = {}) -> StepResult:
        device = next(model.parameters()).device
        timesteps = torch.full((model_inputs["x"].shape[0],), t, device=device, dtype=torch.int32)
        timesteps_model = self._map(timesteps)
        prediction = model(**{**model_inputs, "p": 0}, timesteps=timesteps_model)["x"]
        if guidance_scale > 0:
            prediction_uncond = model(**{**model_inputs, "p": 1}, timesteps=timesteps_model)["x"]
            prediction = prediction_uncond + guidance_scale * (prediction - prediction_uncond)
        return self.sampler.step(model_prediction=prediction, timesteps=timesteps, xt=model_inputs["x"], clamp_x=clamp_x, **sampler_args)

    def compute_loss(self, model: Denoiser, model_inputs: ModelInput, timesteps: Tensor, noise: Tensor | None = None,
                     extra_losses: list[LossFunction] = [], extra_args: dict[str, Any] = {}) -> dict[str, Tensor]:
        model_inputs["x"], noise = self.add_noise(model_inputs["x"], timesteps, noise)  # mutates the caller's dict like the reference
        timesteps = self._map(timesteps.to(model_inputs["x"].device))
        prediction = model(**model_inputs, timesteps=timesteps)["x"]
        loss = _FlowLossFn.apply(prediction, None, noise.float().contiguous(), None, None)  # mean((pred - eps)^2)
        loss_dict = {"loss": loss}
        for extra_loss in extra_losses:
            loss_dict[extra_loss.name] = cast(Tensor, extra_loss(**extra_args))
        return loss_dict

    def denoise(self, model: Denoiser, model_inputs: ModelInput, data_shape: tuple[int, ...] | None = None, use_tqdm: bool = True,
                clamp_x: bool = False, guidance_scale: float = 0, sampler_args: dict[str, Any] = {},
                return_intermediates: bool = False) -> SamplingOutput:
        if "x" not in model_inputs:
            assert data_shape is not None, "'data_shape' must be provided if 'x' is not in model_inputs"
            p0 = next(model.parameters())
            model_inputs["x"] = torch.randn(data_shape, device=p0.device, dtype=p0.dtype)
        keep: dict[str, list[Tensor]] = {"estimated_x0": [], "xt": [model_inputs["x"]], "xt_mean": [], "xt_std": [], "logprob": []}
        for t in list(range(self.steps))[::-1]:
            step = self.one_step_denoise(model=model, model_inputs=model_inputs, t=t, clamp_x=clamp_x, guidance_scale=guidance_scale,
                                         sampler_args=sampler_args)
            model_inputs["x"] = step["x_prev"]
            if return_intermediates:
                keep["estimated_x0"].append(step["estimated_x0"])
                keep["xt"].append(step["x_prev"])
                for src, dst in (("x_prev_mean", "xt_mean"), ("x_prev_std", "xt_std"), ("logprob", "logprob")):
                    if src in step:
                        keep[dst].append(cast(Tensor, step[src]))
        out = SamplingOutput(x=model_inputs["x"])
        if return_intermediates:
            for k, v in keep.items():
                if v:
                    out[k] = torch.stack(v, dim=1)  # type: ignore[literal-required]
        return out
