"""LossFunction interface (mirrors reference training/losses/common.py:10-24)."""

from __future__ import annotations

from abc import ABC
from typing import TYPE_CHECKING

import torch.nn as nn

if TYPE_CHECKING:
    from ..denoisers import Denoiser


class LossFunction(ABC, nn.Module):
    name: str = "extra_loss"

    def __init__(self) -> None:
        super().__init__()

    def set_model(self, model: "Denoiser") -> None:
        pass
