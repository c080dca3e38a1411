"""Sampler interfaces (mirror reference diffuse/samplers/common.py:7-32 and samplers/flow/common.py:9-43)."""

from __future__ import annotations

from abc import ABC, abstractmethod
from typing import Any, NotRequired, Required, TypedDict

from torch import Tensor


class StepResult(TypedDict):
    x_prev: Required[Tensor]
    estimated_x0: Required[Tensor]
    x_prev_mean: NotRequired[Tensor]
    x_prev_std: NotRequired[Tensor]
    logprob: NotRequired[Tensor]


class Sampler(ABC):
    name: str

    def __init__(self) -> None:
        pass

    @abstractmethod
    def set_steps(self, *args: Any, **kwargs: Any) -> None: ...

    @abstractmethod
    def step(self, *args: Any, **kwargs: Any) -> StepResult: ...


class FlowSampler(Sampler, ABC):
    name: str

    @abstractmethod
    def set_steps(self, timesteps: list[float]) -> None: ...

    @abstractmethod
    def step(self, x_t: Tensor, v: Tensor, t_curr: float, t_prev: float, *args: Any, **kwargs: Any) -> StepResult: ...
