"""Helpers shared by the CPU oracle tests and the GPU parity tests: load a golden fixture (generated from the
unmodified reference by oracle/make_golden.py) and run the oracle on it."""
from __future__ import annotations

import glob
import os

import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def fixture_names() -> list[str]:
    """Denoiser fixtures (gaussian.pt holds the model-free Gaussian-diffusion vectors and is loaded separately)."""
    names = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.pt")))
    return [n for n in names if n not in ("gaussian", "gaussian_learned", "flow_misc", "perceiver")]


def load_fixture(name: str) -> dict:
    return torch.load(os.path.join(GOLDEN_DIR, f"{name}.pt"), map_location="cpu", weights_only=False)


def model_kind(fx: dict) -> str:
    sd = fx["state_dict"]
    if "mask_token" in sd:
        return "sprint"
    if "conv_proj_encoder.weight" in sd:
        return "ddt"
    return "mmdit"


def oracle_cfg(fx: dict) -> dict:
    kw = fx["kwargs"]
    d = kw["inner_dim"]
    H = kw["num_heads"]
    hd = d // H
    axes = kw.get("rope_axes_dim")
    if axes is None:
        n = 3 if fx["mm"] else 2
        axes = [int(hd // n)] * n
    cfg = dict(num_heads=H, patch_size=kw["patch_size"], output_channels=kw.get("output_channels") or kw["input_channels"],
               rope_axes_dim=axes, rope_base=kw.get("rope_base", 10000), frequency_embedding=kw.get("frequency_embedding", 256),
               n_classes=kw.get("n_classes"), drop_rate=kw.get("drop_rate", 0.75))
    if fx["mm"]:
        null = fx["null_embedding"]
        L = null.shape[0]
        cfg["null_embedding"] = null
        cfg["null_mask"] = torch.arange(L) < fx["null_valid"]
    return cfg


def oracle_forward(fx: dict, sd: dict, x_t: torch.Tensor, t: torch.Tensor, p: float, draws: dict, training: bool, capture=None):
    from oracle import dit_oracle as O

    cfg = oracle_cfg(fx)
    kind = model_kind(fx)
    common = dict(y=fx["y"], context=fx["context"], p=p, draws=draws, capture=capture)
    if kind == "sprint":
        return O.sprint_forward(sd, cfg, x_t, t, training=training, **common)
    if kind == "ddt":
        return O.ddt_forward(sd, cfg, x_t, t, **common)
    return O.mmdit_forward(sd, cfg, x_t, t, **common)
