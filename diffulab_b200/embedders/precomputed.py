"""PrecomputedEmbedder: tensor passthrough with per-sample null-embedding swap for classifier-free guidance
(mirrors reference networks/embedders/precomputed.py:8-43; same `torch.rand(B) < p` draw so that seeded runs
drop the same samples). Accepts either a path (as the reference) or the null-embedding tensor itself."""

from __future__ import annotations

from pathlib import Path

import torch
from torch import Tensor

from .common import ContextEmbedder, ContextEmbedderOutput


class PrecomputedEmbedder(ContextEmbedder):
    def __init__(self, path_null_embedding: Path | str | Tensor, null_embedding_seq_len: int) -> None:
        super().__init__()
        if isinstance(path_null_embedding, Tensor):
            null = path_null_embedding
        else:
            null = torch.load(path_null_embedding)
        self.null_embedding = null.squeeze()
        self.null_embedding_mask = torch.cat(
            [
                torch.ones(null_embedding_seq_len, dtype=torch.bool),
                torch.zeros(self.null_embedding.shape[0] - null_embedding_seq_len, dtype=torch.bool),
            ],
            dim=0,
        )
        self._output_size = (self.null_embedding.shape[-1],)
        self._n_output = 1
        self._resident: dict = {}  # (device, dtype) -> device copies, so that sampling can be captured in a CUDA graph

    def _null(self, device, dtype) -> tuple[Tensor, Tensor]:
        key = (str(device), dtype)
        if key not in self._resident:
            self._resident[key] = (self.null_embedding.to(device=device, dtype=dtype), self.null_embedding_mask.to(device=device))
        return self._resident[key]

    def drop_conditions(self, context: ContextEmbedderOutput, p: float) -> ContextEmbedderOutput:
        emb = context["embeddings"]
        batch_size = emb.shape[0]
        device, dtype = emb.device, emb.dtype
        drop_mask = torch.rand(batch_size, device=device) < p
        null_emb, null_mask = self._null(device, dtype)
        embeddings = torch.where(drop_mask[:, None, None], null_emb.unsqueeze(0).expand(batch_size, -1, -1), emb)
        attn_mask = torch.where(drop_mask[:, None], null_mask.unsqueeze(0).expand(batch_size, -1), context["attn_mask"])
        return {"embeddings": embeddings, "attn_mask": attn_mask}

    def forward(self, context: ContextEmbedderOutput, p: float = 0) -> ContextEmbedderOutput:
        return self.drop_conditions(context, p)
