"""TEST INFRASTRUCTURE ONLY (never imported by the product path).

Import shim that makes the UNMODIFIED reference (`/root/reference/src/diffulab`) importable in this container,
where nine of its third-party dependencies are absent. Only those absent top-level packages are replaced by
inert stubs; every arithmetic module of the reference (torch, einops) is the real thing. Used by
`oracle/make_golden.py` to generate the committed fixtures under `tests/golden/` and by
`tests/test_oracle_vs_reference.py` (skipped where /root/reference does not exist, i.e. on the GPU box).
"""

from __future__ import annotations

import importlib.abc
import importlib.machinery
import os
import sys
import types
from unittest.mock import MagicMock

REFERENCE_SRC = os.environ.get("DIFFULAB_REFERENCE_SRC", "/root/reference/src")
_MISSING = ("hydra", "omegaconf", "accelerate", "ema_pytorch", "diffusers", "timm", "streaming", "blobfile", "qwen_vl_utils")


class _StubModule(types.ModuleType):
    def __getattr__(self, name: str):
        if name.startswith("__"):
            raise AttributeError(name)
        m = MagicMock(name=f"{self.__name__}.{name}")
        setattr(self, name, m)
        return m


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] in _MISSING:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _StubModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_SRC, "diffulab"))


def import_reference():
    """Returns the reference's `diffulab` package (raises if /root/reference is not mounted)."""
    if not reference_available():
        raise ImportError(f"reference sources not found under {REFERENCE_SRC}")
    if "diffulab" in sys.modules:
        return sys.modules["diffulab"]
    import transformers  # noqa: F401  real package first: its availability probes must not see the stubs

    missing = []
    for name in _MISSING:
        try:
            __import__(name)
        except Exception:
            missing.append(name)
    if missing:
        sys.meta_path.append(_StubFinder())
    if REFERENCE_SRC not in sys.path:
        sys.path.insert(0, REFERENCE_SRC)
    import diffulab

    return diffulab
