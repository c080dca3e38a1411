"""Plugin surface of the denoisers (mirrors reference networks/denoisers/common.py:8-46)."""

from __future__ import annotations

from abc import ABC, abstractmethod
from typing import Any, NotRequired, Required, TypedDict

import torch.nn as nn
from torch import Tensor


class ModelInput(TypedDict, total=False):
    x: Required[Tensor]
    p: NotRequired[float]
    y: NotRequired[Tensor]
    initial_context: NotRequired[Any]
    x_context: NotRequired[Tensor]


class ModelOutput(TypedDict, total=False):
    x: Required[Tensor]
    features: NotRequired[list[Tensor]]
    repa_features: NotRequired[list[Tensor]]


class Denoiser(nn.Module, ABC):
    classifier_free: bool

    def __init__(self) -> None:
        super().__init__()

    @abstractmethod
    def forward(self, x: Tensor, timesteps: Tensor, *args: Any, **kwargs: Any) -> ModelOutput: ...
