"""Synthetic workloads of the BASELINE configs: build the denoiser (+ REPA loss) from one of the
`configs/train_*.yaml` files and generate seeded inputs of that config's shape — latents / pixels, labels or
text-embedding contexts with ragged key masks, REPA target features. Used by `bench.py --config ...` and by the
real-shape parity tests; there is no dataset or network access on the GPU box, so this replaces the reference's
dataloaders (datasets/*.py) for measurement only.

Per-config FLOP model (`flops_per_image`): matmul FLOPs only (2 M N K), SURVEY.md 8(d) "ALGORITHMIC flops".
"""

from __future__ import annotations

import os
from typing import Any

import torch
from torch import Tensor

from .config import instantiate, load_config
from .embedders.precomputed import PrecomputedEmbedder

CONFIG_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "configs")
CONFIGS = {
    "cifar10": "train_cifar10_flow_matching.yaml",
    "imagenet_repa": "train_imagenet_flow_matching_repa.yaml",
    "txt_to_img": "train_imagenet_repa_txt_to_img.yaml",
    "sprint": "train_imagenet_repa_txt_to_img_sprint.yaml",
}


def config_path(name_or_path: str) -> str:
    if name_or_path in CONFIGS:
        return os.path.join(CONFIG_DIR, CONFIGS[name_or_path])
    return name_or_path


class Workload:
    """cfg (composed YAML), model, repa (or None), text spec; `.batch(B, generator)` -> host tensors of one step."""

    def __init__(self, cfg: dict, model, repa, null_embedding: Tensor | None):
        self.cfg, self.model, self.repa, self.null_embedding = cfg, model, repa, null_embedding
        self.shape = tuple(cfg["synthetic"]["image_shape"])
        self.text = cfg["synthetic"].get("text")
        self.n_classes = cfg["model"].get("n_classes")
        self.p_cfg = float(cfg["trainer"].get("p_classifier_free_guidance", 0.0))

    @property
    def mm(self) -> bool:
        return self.text is not None

    def batch(self, B: int, g: torch.Generator) -> dict[str, Any]:
        """Host (CPU) tensors of one step: x0 ~ N(0,1) fp32, labels or {embeddings, attn_mask}, REPA targets."""
        out: dict[str, Any] = {"x": torch.randn(B, *self.shape, generator=g)}
        if self.mm:
            L, D = int(self.text["length"]), int(self.text["dim"])
            lens = torch.randint(8, L + 1, (B,), generator=g)  # ragged prompts: len_b ~ U{8..L}
            out["context"] = {"embeddings": torch.randn(B, L, D, generator=g), "attn_mask": torch.arange(L)[None, :] < lens[:, None]}
        elif self.n_classes:
            out["y"] = torch.randint(0, int(self.n_classes), (B,), generator=g)
        if self.repa is not None:
            out["dst"] = torch.randn(B, int(self.cfg["synthetic"]["repa_tokens"]), int(self.cfg["repa"]["embedding_dim"]), generator=g)
        return out

    @staticmethod
    def to_step(b: dict[str, Any], device, non_blocking: bool = False) -> dict[str, Any]:
        """-> the `batch` dict `training_step` takes ({"model_inputs": ..., "extra": ...}) with tensors on `device`."""
        def mv(t):
            return t.to(device, non_blocking=non_blocking)
        mi: dict[str, Any] = {"x": mv(b["x"])}
        if "context" in b:
            mi["initial_context"] = {k: mv(v) for k, v in b["context"].items()}
        if "y" in b:
            mi["y"] = mv(b["y"])
        extra = {"dst_features": mv(b["dst"])} if "dst" in b else {}
        return {"model_inputs": mi, "extra": extra}

    @staticmethod
    def pin(b: dict[str, Any]) -> dict[str, Any]:
        return {k: ({kk: vv.pin_memory() for kk, vv in v.items()} if isinstance(v, dict) else v.pin_memory()) for k, v in b.items()}

    @staticmethod
    def nbytes(b: dict[str, Any]) -> int:
        n = 0
        for v in b.values():
            for t in (v.values() if isinstance(v, dict) else [v]):
                n += t.numel() * t.element_size()
        return n

    # ---- FLOP model ----------------------------------------------------------------------
    def flops_per_image(self, train: bool = True, kept_tokens: int | None = None) -> float:
        """Forward matmul FLOPs of one image (x3 for training: forward + dgrad + wgrad), incl. the REPA projector when
        training. Blocks are counted by kind: DiT (N tokens), dual-stream MMDiT (N + L tokens, separate linears per
        stream, joint attention), single-stream (joint linears over N + L)."""
        m = self.cfg["model"]
        d = int(m["inner_dim"])
        E = int(m.get("embedding_dim", d))
        C, H, W = self.shape
        ps = int(m["patch_size"])
        N = (H // ps) * (W // ps)
        L = int(self.text["length"]) if self.mm else 0
        C_out = int(m.get("output_channels") or C)
        mlp = int(m.get("mlp_ratio", 4))
        lin_per_tok = 2 * d * (3 * d) + 2 * d * d + 2 * d * (2 * mlp * d) + 2 * (mlp * d) * d  # qkv + out + up + down

        def dit(n_tok: int, mod_rows: int, cond_dim: int) -> float:
            return n_tok * lin_per_tok + 4.0 * n_tok * n_tok * d + 2.0 * mod_rows * cond_dim * 6 * d

        def dual(n_img: int) -> float:
            return (n_img + L) * lin_per_tok + 4.0 * (n_img + L) ** 2 * d + 2 * 2.0 * E * 6 * d

        def single(n_img: int) -> float:
            return (n_img + L) * lin_per_tok + 4.0 * (n_img + L) ** 2 * d + 2.0 * E * 3 * d

        target = m["_target_"].rsplit(".", 1)[-1]
        total = 2.0 * N * (C * ps * ps) * d + 2.0 * N * d * (ps * ps * C_out) + 2.0 * 256 * E + 2.0 * E * E  # embed, last, time MLP
        if self.mm:
            total += 2.0 * L * int(self.text["dim"]) * d
        if target == "MMDiT":
            total += int(m["depth"]) * (dual(N) if self.mm else dit(N, 1, E)) + 2.0 * E * 2 * d
        elif target == "DDT":
            n_single = int(m.get("n_single_stream_blocks", 0))
            enc = int(m["encoder_depth"])
            total += (enc - n_single) * (dual(N) if self.mm else dit(N, 1, d)) + n_single * single(N)
            total += 2.0 * N * (C * ps * ps) * d  # conv_proj_decoder
            total += int(m["decoder_depth"]) * dit(N, N, d) + 2.0 * N * d * 2 * d  # per-token modulation GEMMs
        elif target == "SprintDiT":
            k = kept_tokens if kept_tokens is not None else (max(1, int(N * (1.0 - float(m.get("drop_rate", 0.75))))) if train else N)
            n_single = int(m.get("n_single_stream_blocks", 0))
            blk = dual if self.mm else (lambda n: dit(n, 1, E))
            total += int(m["encoder_depth"]) * blk(N) + int(m["decoder_depth"]) * blk(N)
            total += (int(m["deep_layers_depth"]) - n_single) * blk(k) + n_single * (single(k) if self.mm else dit(k, 1, E))
            total += 2.0 * N * 2 * d * d + (2.0 * L * 2 * d * d if self.mm else 0.0) + 2.0 * E * 2 * d  # fuse, fuse_context, last adaLN
        else:
            raise ValueError(target)
        if not train:
            return total
        total *= 3.0
        if self.repa is not None:
            r = self.cfg["repa"]
            hid, Er = int(r["hidden_dim"]), int(r["embedding_dim"])
            total += 3.0 * 2.0 * N * (int(r["denoiser_dimension"]) * hid + hid * hid + hid * Er)
        return total


def rerandomize_zero_init(model: torch.nn.Module, seed: int, std: float = 0.02) -> None:
    """adaLN-Zero makes every block the identity at initialisation (zero modulation linears, zero mask token): redraw the
    all-zero parameters with N(0, std) so that parity checks and benchmarks exercise every kernel with live gates
    (the same treatment the reference-generated test fixtures receive)."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for p in model.parameters():
            if p.abs().sum() == 0:
                p.copy_((torch.randn(p.shape, generator=g) * std).to(p.device))


def build_workload(name_or_path: str, overrides: list[str] | None = None, device=None, seed: int = 1234, live_gates: bool = True) -> Workload:
    cfg = load_config(config_path(name_or_path), overrides or [])
    text = cfg["synthetic"].get("text")
    torch.manual_seed(seed)
    null = None
    extra: dict[str, Any] = {}
    if text is not None:
        g = torch.Generator().manual_seed(seed + 17)
        null = torch.randn(int(text["length"]), int(text["dim"]), generator=g)
        extra["context_embedder"] = PrecomputedEmbedder(null, int(text["null_valid"]))
    model = instantiate(cfg["model"], **extra)
    repa = instantiate(cfg["repa"]) if "repa" in cfg else None
    if live_gates:
        rerandomize_zero_init(model, seed + 1)
    if device is not None:
        model = model.to(device)
        repa = repa.to(device) if repa is not None else None
    if repa is not None:
        repa.set_model(model)
    return Workload(cfg, model, repa, null)
