"""Thin tensor-level wrappers over the C ABI (raw device pointers + current CUDA stream).

Every function here launches hand-written sm_100a kernels from libdiffulab_b200.so; nothing in this module
computes with PyTorch ops. Callers own all memory (outputs are `torch.empty` tensors passed by pointer).
"""

from __future__ import annotations

import torch
from torch import Tensor

from . import _lib

BF16 = torch.bfloat16
F32 = torch.float32


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Tensor | None) -> int | None:
    return None if t is None else t.data_ptr()


def _check_2d(t: Tensor, name: str) -> None:
    if t.dim() != 2 or t.stride(1) != 1:
        raise ValueError(f"{name}: expected a 2-D tensor with unit inner stride, got shape {tuple(t.shape)} strides {t.stride()}")
    if not t.is_cuda:
        raise ValueError(f"{name}: expected a CUDA tensor (diffulab_b200 has no CPU path)")


def gemm(
    a: Tensor,
    b: Tensor,
    *,
    bias: Tensor | None = None,
    out: Tensor | None = None,
    a_mn: bool = False,
    b_mn: bool = False,
    out_dtype: torch.dtype = BF16,
    accumulate: bool = False,
    split_k: int = 1,
    tile_n: int = 0,
) -> Tensor:
    """C[M,N] (+)= A @ B^T (+ bias) with bf16 operands and fp32 accumulation on tcgen05.

    a: [M,K] (or [K,M] when a_mn), b: [N,K] (or [K,N] when b_mn). Row strides may exceed the row length
    (views into packed buffers) as long as they are multiples of 8 elements.
    accumulate=True adds into an fp32 `out` (TMA reduce-add), optionally split along K.
    """
    _check_2d(a, "a")
    _check_2d(b, "b")
    if a.dtype != BF16 or b.dtype != BF16:
        raise ValueError("gemm operands must be bfloat16")
    M, K = (a.shape[1], a.shape[0]) if a_mn else (a.shape[0], a.shape[1])
    N, Kb = (b.shape[1], b.shape[0]) if b_mn else (b.shape[0], b.shape[1])
    if K != Kb:
        raise ValueError(f"gemm: inner dimensions differ ({K} vs {Kb})")
    if accumulate:
        if out is None or out.dtype != F32:
            raise ValueError("accumulate=True needs an fp32 `out`")
        mode = 2
    else:
        mode = 0 if out_dtype == BF16 else 1
        if out is None:
            out = torch.empty((M, N), device=a.device, dtype=out_dtype)
    _check_2d(out, "out")
    if tuple(out.shape) != (M, N):
        raise ValueError(f"gemm: out has shape {tuple(out.shape)}, expected {(M, N)}")
    if bias is not None and (bias.dtype != F32 or bias.numel() != N or not bias.is_contiguous()):
        raise ValueError("gemm: bias must be a contiguous fp32 vector of length N")
    rc = _lib.load().dlb_gemm_bf16(
        a.data_ptr(), b.data_ptr(), out.data_ptr(), _ptr(bias), M, N, K,
        a.stride(0), b.stride(0), out.stride(0), int(a_mn), int(b_mn), mode, split_k, tile_n, _stream(),
    )
    _lib.check(rc, "dlb_gemm_bf16")
    return out
