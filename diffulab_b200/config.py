"""Hydra-subset configuration loader: enough of Hydra/OmegaConf to drive this package from DiffuLab's own YAML
configs (reference configs/*.yaml; features actually used are listed in SURVEY.md Appendix C), because neither
hydra nor omegaconf is installed on the GPU image.

Supported: a `defaults:` list of `- group: option` entries and `- _self_`; group files `<dir>/<group>/<option>.yaml`
mounted at key `<group>`; deep merge in defaults order; `key.sub=value` command-line overrides; `_target_`
instantiation (recursive) with extra keyword arguments; the `hydra:` block is ignored. Reference targets
(`diffulab.networks.MMDiT`, ...) are transparently mapped to this package's drop-in classes.
"""

from __future__ import annotations

import copy
import importlib
import os
import re
from typing import Any

import yaml

# reference dotted path -> drop-in implementation in this package
TARGET_MAP = {
    "diffulab.networks.MMDiT": "diffulab_b200.MMDiT",
    "diffulab.networks.denoisers.MMDiT": "diffulab_b200.MMDiT",
    "diffulab.networks.SprintDiT": "diffulab_b200.SprintDiT",
    "diffulab.networks.DDT": "diffulab_b200.DDT",
    "diffulab.networks.PrecomputedEmbedder": "diffulab_b200.PrecomputedEmbedder",
    "diffulab.diffuse.Diffuser": "diffulab_b200.Diffuser",
    "diffulab.training.losses.RepaLoss": "diffulab_b200.RepaLoss",
    "torch.optim.AdamW": "diffulab_b200.training.FusedAdamW",
}


def deep_merge(base: dict, over: dict) -> dict:
    out = copy.deepcopy(base)
    for k, v in over.items():
        if isinstance(v, dict) and isinstance(out.get(k), dict):
            out[k] = deep_merge(out[k], v)
        else:
            out[k] = copy.deepcopy(v)
    return out


class _Loader(yaml.SafeLoader):
    """SafeLoader that also reads `1e-4` as a float (OmegaConf does; YAML 1.1 wants `1.0e-4`)."""


_Loader.add_implicit_resolver(
    "tag:yaml.org,2002:float",
    re.compile(r"^[-+]?(?:[0-9][0-9_]*)(?:\.[0-9_]*)?[eE][-+]?[0-9]+$"),
    list("-+0123456789"),
)


def _parse(text: str) -> Any:
    return yaml.load(text, Loader=_Loader)


def _load_yaml(path: str) -> dict:
    with open(path) as f:
        data = _parse(f.read())
    return data or {}


def _set_path(cfg: dict, dotted: str, value: Any) -> None:
    keys = dotted.split(".")
    node = cfg
    for k in keys[:-1]:
        node = node.setdefault(k, {})
        if not isinstance(node, dict):
            raise ValueError(f"override {dotted}: {k} is not a mapping")
    node[keys[-1]] = value


def load_config(path: str, overrides: list[str] | None = None) -> dict:
    """Compose `path` the way `@hydra.main(config_path=..., config_name=...)` would (subset, see module docstring)."""
    root = os.path.dirname(os.path.abspath(path))
    top = _load_yaml(path)
    defaults = top.pop("defaults", [])
    top.pop("hydra", None)
    cfg: dict = {}
    self_done = False
    for entry in defaults:
        if entry == "_self_":
            cfg = deep_merge(cfg, top)
            self_done = True
            continue
        if not isinstance(entry, dict) or len(entry) != 1:
            raise ValueError(f"unsupported defaults entry: {entry!r}")
        (group, option), = entry.items()
        if option is None:
            continue
        gpath = os.path.join(root, group, f"{option}.yaml")
        if not os.path.exists(gpath):
            raise FileNotFoundError(f"config group file not found: {gpath}")
        cfg = deep_merge(cfg, {group: _load_yaml(gpath)})
    if not self_done:
        cfg = deep_merge(cfg, top)
    for ov in overrides or []:
        if "=" not in ov:
            raise ValueError(f"override must look like key=value, got {ov!r}")
        k, v = ov.split("=", 1)
        _set_path(cfg, k.lstrip("+"), _parse(v))
    return cfg


def _resolve(target: str):
    target = TARGET_MAP.get(target, target)
    mod, _, attr = target.rpartition(".")
    return getattr(importlib.import_module(mod), attr)


def instantiate(node: Any, **extra: Any) -> Any:
    """hydra.utils.instantiate subset: build `_target_(**kwargs, **extra)`, recursing into nested `_target_` nodes."""
    if isinstance(node, dict) and "_target_" in node:
        kwargs = {k: instantiate(v) for k, v in node.items() if k != "_target_"}
        kwargs.update(extra)
        return _resolve(node["_target_"])(**kwargs)
    if isinstance(node, dict):
        return {k: instantiate(v) for k, v in node.items()}
    if isinstance(node, list):
        return [instantiate(v) for v in node]
    return node
