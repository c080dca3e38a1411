"""Diffusion ABC (mirrors reference diffuse/modelizations/diffusion.py:13-244)."""

from __future__ import annotations

from abc import ABC, abstractmethod
from typing import Any, NotRequired, Required, TypedDict

from torch import Tensor

from ..denoisers.common import Denoiser, ModelInput
from ..losses.common import LossFunction
from .samplers.common import Sampler, StepResult


class SamplingOutput(TypedDict, total=False):
    x: Required[Tensor]
    estimated_x0: NotRequired[Tensor]
    xt: NotRequired[Tensor]
    xt_mean: NotRequired[Tensor]
    xt_std: NotRequired[Tensor]
    logprob: NotRequired[Tensor]


class Diffusion(ABC):
    sampler_registry: dict[str, type[Sampler]]

    def __init__(self, n_steps: int, sampling_method: str = "euler", schedule: str = "linear", latent_diffusion: bool = False,
                 sampler_parameters: dict[str, Any] = {}):
        assert sampling_method in self.sampler_registry, (
            f"Unknown sampling method '{sampling_method}'. Available methods: {list(self.sampler_registry.keys())}"
        )
        self.sampler = self.sampler_registry[sampling_method](**sampler_parameters)
        self.timesteps: list[float] = []
        self.steps: int = n_steps
        self.sampling_method = sampling_method
        self.schedule = schedule
        self.latent_diffusion = latent_diffusion
        self.set_steps(n_steps, schedule=schedule)

    @abstractmethod
    def set_steps(self, n_steps: int, schedule: str) -> None: ...

    @abstractmethod
    def one_step_denoise(self, model: Denoiser, model_inputs: ModelInput, guidance_scale: float, *args: Any, **kwargs: Any) -> StepResult: ...

    @abstractmethod
    def compute_loss(self, model: Denoiser, model_inputs: ModelInput, timesteps: Tensor, noise: Tensor | None = None,
                     extra_losses: list[LossFunction] = [], extra_args: dict[str, Any] = {}) -> dict[str, Tensor]: ...

    @abstractmethod
    def add_noise(self, x: Tensor, timesteps: Tensor, noise: Tensor | None = None) -> tuple[Tensor, Tensor]: ...

    @abstractmethod
    def denoise(self, model: Denoiser, model_inputs: ModelInput, *args: Any, **kwargs: Any) -> SamplingOutput: ...

    @abstractmethod
    def draw_timesteps(self, batch_size: int) -> Tensor: ...
