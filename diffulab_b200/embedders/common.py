"""ContextEmbedder ABC (mirrors reference networks/embedders/common.py:14-64)."""

from __future__ import annotations

from abc import ABC, abstractmethod
from typing import Any, NotRequired, Required, TypedDict

import torch.nn as nn
from torch import Tensor


class ContextEmbedderOutput(TypedDict):
    embeddings: Required[Tensor]
    pooled_embeddings: NotRequired[Tensor]
    attn_mask: NotRequired[Tensor]


class ContextEmbedder(nn.Module, ABC):
    _n_output: int
    _output_size: tuple[int, ...]

    def __init__(self) -> None:
        super().__init__()

    @property
    def n_output(self) -> int:
        return self._n_output

    @property
    def output_size(self) -> tuple[int, ...]:
        return self._output_size

    @abstractmethod
    def drop_conditions(self, context: Any, p: float) -> Any: ...

    @abstractmethod
    def forward(self, context: Any, p: float = 0) -> ContextEmbedderOutput: ...
