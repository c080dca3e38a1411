"""ctypes binding of libdiffulab_b200.so (the C-ABI boundary declared in include/diffulab_b200.h).

There is no CPU or PyTorch fallback: if the shared library is missing or a call fails, this module raises.
"""

from __future__ import annotations

import ctypes as C
from pathlib import Path

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "libdiffulab_b200.so"

_lib: C.CDLL | None = None

p = C.c_void_p
i64 = C.c_int64
i32 = C.c_int
f32 = C.c_float

# name -> argtypes (restype is always int unless listed in _RESTYPES). Mirrors include/diffulab_b200.h.
_SIGNATURES: dict[str, list] = {
    "dlb_version": [],
    "dlb_last_error": [],
    "dlb_launch_count": [],
    "dlb_reset_launch_count": [],
    "dlb_device_check": [],
    # C[M,N] (+)= A*B^T (+bias): A, B, C, bias, M, N, K, lda, ldb, ldc, a_mn, b_mn, out_mode, split_k, tile_n, stream
    "dlb_gemm_bf16": [p, p, p, p, i64, i64, i64, i64, i64, i64, i32, i32, i32, i32, i32, p],
}
_RESTYPES = {"dlb_last_error": C.c_char_p, "dlb_launch_count": C.c_longlong, "dlb_reset_launch_count": None}


def register(name: str, argtypes: list) -> None:
    _SIGNATURES[name] = argtypes


def exported_symbols() -> list[str]:
    return sorted(_SIGNATURES)


def load() -> C.CDLL:
    """Load the library once; raise (never fall back) when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m diffulab_b200.build` "
            "(diffulab_b200 has no CPU or PyTorch fallback path)"
        )
    lib = C.CDLL(str(LIB_PATH))
    for name, argtypes in _SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the header and the library disagree
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, C.c_int)
    _lib = lib
    return lib


class DlbError(RuntimeError):
    pass


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().dlb_last_error()
        raise DlbError(f"{what} failed (rc={rc}): {msg.decode() if msg else ''}")


def launch_count() -> int:
    return int(load().dlb_launch_count())


def reset_launch_count() -> None:
    load().dlb_reset_launch_count()
