"""Diffuser facade (mirrors reference diffuse/diffuser.py:17-239: registries, compute_loss, generate, set_steps)."""

from __future__ import annotations

from typing import Any

from torch import Tensor

from ..denoisers.common import Denoiser, ModelInput
from ..losses.common import LossFunction
from .diffusion import Diffusion, SamplingOutput
from .flow import Flow
from .gaussian import GaussianDiffusion


class Diffuser:
    model_registry: dict[str, type[Diffusion]] = {"rectified_flow": Flow, "gaussian_diffusion": GaussianDiffusion}

    def __init__(self, denoiser: Denoiser, sampling_method: str, model_type: str = "rectified_flow", n_steps: int = 1000,
                 vision_tower: Any | None = None, extra_args: dict[str, Any] = {}, extra_losses: list[LossFunction] = []):
        self.model_type = model_type
        self.denoiser = denoiser
        self.n_steps = n_steps
        self.vision_tower = vision_tower
        self.extra_losses = extra_losses
        if self.vision_tower:
            self.latent_scale = self.vision_tower.latent_scale
            self.latent_bias = self.vision_tower.latent_bias
        if self.model_type in self.model_registry:
            self.diffusion = self.model_registry[self.model_type](
                n_steps=n_steps, sampling_method=sampling_method, latent_diffusion=self.vision_tower is not None, **extra_args
            )
        else:
            raise NotImplementedError(f"Model type {self.model_type} is not implemented")

    def eval(self) -> None:
        self.denoiser.eval()

    def train(self) -> None:
        self.denoiser.train()

    def draw_timesteps(self, batch_size: int) -> Tensor:
        return self.diffusion.draw_timesteps(batch_size=batch_size)

    def compute_loss(self, model_inputs: ModelInput, timesteps: Tensor | None = None, noise: Tensor | None = None,
                     extra_args: dict[str, Any] = {}, grpo: bool = False, grpo_args: dict[str, Any] = {}) -> dict[str, Tensor]:
        if grpo:
            raise NotImplementedError("GRPO fine-tuning is outside the accelerated hot path (SURVEY.md 2.1 #15)")
        assert timesteps is not None, "timesteps must be provided for loss computation"
        return self.diffusion.compute_loss(self.denoiser, model_inputs, timesteps, noise, self.extra_losses, extra_args)

    def set_steps(self, n_steps: int, schedule: str = "linear", *args: Any, **kwargs: Any) -> None:
        self.diffusion.set_steps(n_steps, schedule=schedule, *args, **kwargs)

    def generate(self, model_inputs: ModelInput, data_shape: tuple[int, ...] | None = None, use_tqdm: bool = True,
                 clamp_x: bool = False, guidance_scale: float = 0, sampler_args: dict[str, Any] = {},
                 return_intermediates: bool = False, return_latents: bool = False) -> SamplingOutput:
        out = self.diffusion.denoise(self.denoiser, model_inputs=model_inputs, data_shape=data_shape, use_tqdm=use_tqdm,
                                     clamp_x=clamp_x, guidance_scale=guidance_scale, sampler_args=sampler_args,
                                     return_intermediates=return_intermediates)
        if self.vision_tower and not return_latents:
            latent_scale, latent_bias = self.latent_scale, self.latent_bias
            if isinstance(latent_scale, Tensor):
                latent_scale = latent_scale.to(out["x"].device)
            if isinstance(latent_bias, Tensor):
                latent_bias = latent_bias.to(out["x"].device)
            out["x"] = self.vision_tower.decode(out["x"] / latent_scale + latent_bias)
        return out
