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


# Optional per-call device timing (CUDA events on the launching stream). `profile_start()` turns it on; every C-ABI
# call then records (name, work, start_event, end_event); `profile_stop()` synchronises and returns
# {name: {"calls", "ms", "work"}} where work = FLOPs for the GEMM and 0 otherwise. Used by bench.py's roofline.
_prof: list | None = None


def profile_start() -> None:
    global _prof
    _prof = []


def profile_stop() -> dict[str, dict[str, float]]:
    global _prof
    rec, _prof = _prof or [], None
    torch.cuda.synchronize()
    out: dict[str, dict[str, float]] = {}
    for name, work, e0, e1 in rec:
        d = out.setdefault(name, {"calls": 0, "ms": 0.0, "work": 0.0})
        d["calls"] += 1
        d["ms"] += e0.elapsed_time(e1)
        d["work"] += work
    return out


class _Timed:
    __slots__ = ("name", "work", "e0")

    def __init__(self, name: str, work: float = 0.0):
        self.name, self.work = name, work

    def __enter__(self):
        if _prof is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if _prof is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            _prof.append((self.name, self.work, self.e0, e1))
        return False


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
    split_k: int | None = None,
    tile_n: int = 0,
    pair: bool = False,
) -> Tensor:
    """C[M,N] (+)= A @ B^T (+ bias) with bf16 operands and fp32 accumulation on tcgen05.

    a: [M,K] (or [K,M] when a_mn), b: [N,K] (or [K,N] when b_mn). Row strides may exceed the row length
    (views into packed buffers) as long as they are multiples of 8 elements.
    accumulate=True adds into an fp32 `out` (TMA reduce-add), optionally split along K
    (split_k None/0 = chosen by the library's cost model).
    """
    if split_k is None:
        split_k = 0 if accumulate else 1
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
    kind = "gemm_wgrad" if (a_mn and b_mn) else ("gemm_dgrad" if b_mn else "gemm_fwd")
    with _Timed(f"{kind} {M}x{N}x{K}" if _prof is not None else kind, 2.0 * M * N * K):
        fn = _lib.load().dlb_gemm2_bf16 if pair else _lib.load().dlb_gemm_bf16  # pair: force the CTA-pair kernel (tests / benches)
        rc = fn(
            a.data_ptr(), b.data_ptr(), out.data_ptr(), _ptr(bias), M, N, K,
            a.stride(0), b.stride(0), out.stride(0), int(a_mn), int(b_mn), mode, max(split_k, 1) if pair else split_k,
            (tile_n or 256) if pair else tile_n, _stream(),
        )
    _lib.check(rc, "dlb_gemm_bf16")
    return out


# ---------------------------------------------------------------------------------------------------------
# helpers
# ---------------------------------------------------------------------------------------------------------
def _lib_call(name: str, *args) -> None:
    if _prof is not None:
        with _Timed(name[4:]):
            rc = getattr(_lib.load(), name)(*args)
    else:
        rc = getattr(_lib.load(), name)(*args)
    _lib.check(rc, name)


def _rows(x: Tensor) -> int:
    return x.numel() // x.shape[-1]


def _req(t: Tensor, dtype: torch.dtype, name: str, contiguous: bool = True) -> None:
    if t.dtype != dtype:
        raise ValueError(f"{name}: expected {dtype}, got {t.dtype}")
    if not t.is_cuda:
        raise ValueError(f"{name}: expected a CUDA tensor (diffulab_b200 has no CPU path)")
    if contiguous and not t.is_contiguous():
        raise ValueError(f"{name}: expected a contiguous tensor")


def _mod_view(m: Tensor, d: int, name: str) -> tuple[int, int, int]:
    """Modulation chunk [G, d] (a column slice of the [G, k*d] adaLN output). Returns (ptr, ld, groups)."""
    m2 = m.reshape(-1, m.shape[-1]) if m.is_contiguous() else m
    if m2.dim() == 3:  # [B, 1, d] or [B, N, d] views keep their strides
        if m2.shape[1] == 1:
            m2 = m2[:, 0, :]
        else:
            if m2.stride(0) != m2.shape[1] * m2.stride(1):
                raise ValueError(f"{name}: unsupported per-token modulation strides {m.stride()}")
            m2 = m2.as_strided((m2.shape[0] * m2.shape[1], m2.shape[2]), (m2.stride(1), 1), m2.storage_offset())
    if m2.dim() != 2 or m2.shape[1] != d or m2.stride(1) != 1 or m2.dtype != BF16:
        raise ValueError(f"{name}: expected a bf16 [G, {d}] view with unit inner stride, got {tuple(m.shape)} {m.stride()} {m.dtype}")
    return m2.data_ptr(), m2.stride(0) if m2.shape[0] > 1 else d, m2.shape[0]


# ---------------------------------------------------------------------------------------------------------
# LayerNorm + modulate
# ---------------------------------------------------------------------------------------------------------
def ln_modulate_fwd(x: Tensor, w: Tensor | None, b: Tensor | None, scale: Tensor, shift: Tensor, eps: float):
    _req(x, BF16, "x")
    d = x.shape[-1]
    R = _rows(x)
    sp, sld, G = _mod_view(scale, d, "scale")
    hp, hld, G2 = _mod_view(shift, d, "shift")
    if G != G2 or sld != hld or R % G != 0:
        raise ValueError("ln_modulate: scale/shift layout mismatch")
    y = torch.empty_like(x)
    mean = torch.empty(R, device=x.device, dtype=F32)
    rstd = torch.empty(R, device=x.device, dtype=F32)
    _lib_call("dlb_ln_modulate_fwd", x.data_ptr(), _ptr(w), _ptr(b), sp, hp, sld, R // G, y.data_ptr(),
              mean.data_ptr(), rstd.data_ptr(), R, d, eps, _stream())
    return y, mean, rstd


def ln_modulate_bwd(dy: Tensor, x: Tensor, mean: Tensor, rstd: Tensor, w: Tensor | None, b: Tensor | None,
                    scale: Tensor, dres: Tensor | None, dscale: Tensor, dshift: Tensor,
                    dw: Tensor | None, db: Tensor | None) -> Tensor:
    """Returns dx (+dres). Accumulates into dscale/dshift (fp32 [G,d] views for per-sample modulation; bf16 [R,d]
    views written (not accumulated) for per-token modulation) and into dw/db (fp32 [d])."""
    _req(dy, BF16, "dy")
    _req(x, BF16, "x")
    d = x.shape[-1]
    R = _rows(x)
    sp, sld, G = _mod_view(scale, d, "scale")
    per_token = G == R and R > 1 and dscale.dtype == BF16
    dx = torch.empty_like(x)
    if per_token:
        ds2 = dscale.reshape(-1, d) if dscale.is_contiguous() else dscale.as_strided((R, d), (dscale.stride(-2), 1), dscale.storage_offset())
        dh2 = dshift.reshape(-1, d) if dshift.is_contiguous() else dshift.as_strided((R, d), (dshift.stride(-2), 1), dshift.storage_offset())
        _lib_call("dlb_ln_modulate_bwd", dy.data_ptr(), x.data_ptr(), mean.data_ptr(), rstd.data_ptr(), _ptr(w), _ptr(b),
                  sp, sld, 1, R, 1, _ptr(dres), dx.data_ptr(), None, None, 0, ds2.data_ptr(), dh2.data_ptr(),
                  ds2.stride(0), _ptr(dw), _ptr(db), d, _stream())
    else:
        if dscale.dtype != F32 or dshift.dtype != F32 or dscale.stride(-1) != 1:
            raise ValueError("ln_modulate_bwd: dscale/dshift must be fp32 views")
        ds2 = dscale.reshape(-1, d) if dscale.dim() != 2 else dscale
        dh2 = dshift.reshape(-1, d) if dshift.dim() != 2 else dshift
        ld = ds2.stride(0) if ds2.shape[0] > 1 else d
        _lib_call("dlb_ln_modulate_bwd", dy.data_ptr(), x.data_ptr(), mean.data_ptr(), rstd.data_ptr(), _ptr(w), _ptr(b),
                  sp, sld, G, R // G, 0, _ptr(dres), dx.data_ptr(), ds2.data_ptr(), dh2.data_ptr(), ld, None, None, 0,
                  _ptr(dw), _ptr(db), d, _stream())
    return dx


# ---------------------------------------------------------------------------------------------------------
# gated residual
# ---------------------------------------------------------------------------------------------------------
def gate_residual_fwd(x: Tensor, a1: Tensor, a2: Tensor | None, gate: Tensor) -> Tensor:
    _req(x, BF16, "x")
    _req(a1, BF16, "a1")
    d = x.shape[-1]
    R = _rows(x)
    gp, gld, G = _mod_view(gate, d, "gate")
    out = torch.empty_like(x)
    _lib_call("dlb_gate_residual_fwd", x.data_ptr(), a1.data_ptr(), _ptr(a2), gp, gld, R // G, out.data_ptr(), R, d, _stream())
    return out


def gate_residual_bwd(dout: Tensor, a1: Tensor, a2: Tensor | None, gate: Tensor, dgate: Tensor) -> Tensor:
    """Returns da (= dout * gate). dgate: fp32 [G,d] view (accumulated) or bf16 [R,d] view (per-token, written)."""
    _req(dout, BF16, "dout")
    d = dout.shape[-1]
    R = _rows(dout)
    gp, gld, G = _mod_view(gate, d, "gate")
    da = torch.empty_like(dout)
    per_token = G == R and R > 1 and dgate.dtype == BF16
    if per_token:
        dg2 = dgate.reshape(-1, d) if dgate.is_contiguous() else dgate.as_strided((R, d), (dgate.stride(-2), 1), dgate.storage_offset())
        _lib_call("dlb_gate_residual_bwd", dout.data_ptr(), a1.data_ptr(), _ptr(a2), gp, gld, 1, R, 1, da.data_ptr(), None, 0,
                  dg2.data_ptr(), dg2.stride(0), d, _stream())
    else:
        dg2 = dgate.reshape(-1, d) if dgate.dim() != 2 else dgate
        ld = dg2.stride(0) if dg2.shape[0] > 1 else d
        _lib_call("dlb_gate_residual_bwd", dout.data_ptr(), a1.data_ptr(), _ptr(a2), gp, gld, G, R // G, 0, da.data_ptr(),
                  dg2.data_ptr(), ld, None, 0, d, _stream())
    return da


# ---------------------------------------------------------------------------------------------------------
# SwiGLU
# ---------------------------------------------------------------------------------------------------------
def gemm_swiglu(a: Tensor, w: Tensor, bias: Tensor | None = None) -> tuple[Tensor, Tensor]:
    """(H, ACT) = (a @ w^T (+ bias), silu(H[:, :F]) * H[:, F:]) with the activation fused into the GEMM epilogue.
    a [M,K], w [2F,K] bf16; F % 128 == 0 (callers fall back to gemm + swiglu_fwd otherwise)."""
    _check_2d(a, "a")
    _check_2d(w, "w")
    if a.dtype != BF16 or w.dtype != BF16:
        raise ValueError("gemm_swiglu operands must be bfloat16")
    M, K = a.shape
    F = w.shape[0] // 2
    if w.shape[1] != K or w.shape[0] != 2 * F:
        raise ValueError(f"gemm_swiglu: weight shape {tuple(w.shape)} does not match K={K}")
    h = torch.empty((M, 2 * F), device=a.device, dtype=BF16)
    act = torch.empty((M, F), device=a.device, dtype=BF16)
    with _Timed(f"gemm_swiglu_fwd {M}x{2 * F}x{K}" if _prof is not None else "gemm_swiglu_fwd", 2.0 * M * 2 * F * K):
        rc = _lib.load().dlb_gemm_swiglu_bf16(a.data_ptr(), w.data_ptr(), _ptr(bias), h.data_ptr(), act.data_ptr(), M, F, K,
                                              a.stride(0), w.stride(0), h.stride(0), act.stride(0), _stream())
    _lib.check(rc, "dlb_gemm_swiglu_bf16")
    return h, act


def gemm_swiglu_bwd(dy: Tensor, w2: Tensor, h: Tensor) -> Tensor:
    """dH = SwiGLU-backward(h, dy @ w2) with d(act) kept on chip. dy [M,D], w2 [D,F] (the down-projection weight), h [M,2F]."""
    _check_2d(dy, "dy")
    _check_2d(w2, "w2")
    _check_2d(h, "h")
    if dy.dtype != BF16 or w2.dtype != BF16 or h.dtype != BF16:
        raise ValueError("gemm_swiglu_bwd operands must be bfloat16")
    M, D = dy.shape
    F = w2.shape[1]
    if w2.shape[0] != D or tuple(h.shape) != (M, 2 * F):
        raise ValueError(f"gemm_swiglu_bwd: shapes dy {tuple(dy.shape)}, w2 {tuple(w2.shape)}, h {tuple(h.shape)} do not match")
    dh = torch.empty_like(h)
    with _Timed(f"gemm_swiglu_bwd {M}x{F}x{D}" if _prof is not None else "gemm_swiglu_bwd", 2.0 * M * F * D):
        rc = _lib.load().dlb_gemm_swiglu_bwd_bf16(dy.data_ptr(), w2.data_ptr(), h.data_ptr(), dh.data_ptr(), M, F, D, dy.stride(0),
                                                  w2.stride(0), h.stride(0), dh.stride(0), _stream())
    _lib.check(rc, "dlb_gemm_swiglu_bwd_bf16")
    return dh


def swiglu_fwd(h: Tensor) -> Tensor:
    _req(h, BF16, "h")
    F = h.shape[-1] // 2
    out = torch.empty(*h.shape[:-1], F, device=h.device, dtype=BF16)
    _lib_call("dlb_swiglu_fwd", h.data_ptr(), out.data_ptr(), _rows(h), F, _stream())
    return out


def swiglu_bwd(dout: Tensor, h: Tensor) -> Tensor:
    _req(dout, BF16, "dout")
    _req(h, BF16, "h")
    dh = torch.empty_like(h)
    _lib_call("dlb_swiglu_bwd", dout.data_ptr(), h.data_ptr(), dh.data_ptr(), _rows(h), h.shape[-1] // 2, _stream())
    return dh


# ---------------------------------------------------------------------------------------------------------
# RoPE table, QK-norm + RoPE
# ---------------------------------------------------------------------------------------------------------
class RopeTable:
    """cos / sin fp32 [P, R/2] (get_cos_sin_ndim_grid, nn.py:262-307) + the packed bf16x2 table the kernels read."""

    def __init__(self, cos: Tensor, sin: Tensor, cs: Tensor):
        self.cos, self.sin, self.cs = cos, sin, cs

    def rows(self, idx: Tensor) -> "RopeTable":
        return RopeTable(self.cos[idx].contiguous(), self.sin[idx].contiguous(), self.cs[idx].contiguous())


def rope_table(pos_ids: Tensor, axes_dim: list[int], base: float) -> RopeTable:
    """pos_ids: int32 [P, n_axes] on device."""
    _req(pos_ids, torch.int32, "pos_ids")
    P, n_axes = pos_ids.shape
    axis_of, local_of = [], []
    for a, dim in enumerate(axes_dim):
        for j in range(dim // 2):
            axis_of.append(a)
            local_of.append(j)
    rot_half = len(axis_of)
    dev = pos_ids.device
    ax = torch.tensor(axis_of, dtype=torch.int32, device=dev)
    lo = torch.tensor(local_of, dtype=torch.int32, device=dev)
    ad = torch.tensor(axes_dim, dtype=torch.int32, device=dev)
    cos = torch.empty(P, rot_half, device=dev, dtype=F32)
    sin = torch.empty(P, rot_half, device=dev, dtype=F32)
    cs = torch.empty(P, rot_half, device=dev, dtype=torch.int32)
    _lib_call("dlb_rope_table", pos_ids.data_ptr(), n_axes, ax.data_ptr(), lo.data_ptr(), ad.data_ptr(), float(base),
              cos.data_ptr(), sin.data_ptr(), cs.data_ptr(), P, rot_half, _stream())
    return RopeTable(cos, sin, cs)


def qknorm_rope_fwd(qkv: Tensor, sq: Tensor, sk: Tensor, rope: RopeTable, hd: int, *, tokens_per_sample: int,
                    pos_offset: int = 0, pos_idx: Tensor | None = None, eps: float = 1e-6) -> tuple[Tensor, Tensor]:
    """qkv: [R, 3d] packed projection -> ([R, 2d] normalised + rotated (q | k), rrms fp32 [R, 2] for the backward)."""
    _req(qkv, BF16, "qkv")
    R = _rows(qkv)
    d = qkv.shape[-1] // 3
    out = torch.empty(R, 2 * d, device=qkv.device, dtype=BF16)
    rrms = torch.empty(R, 2, device=qkv.device, dtype=F32)
    _lib_call("dlb_qknorm_rope_fwd", qkv.data_ptr(), 3 * d, sq.data_ptr(), sk.data_ptr(), rope.cs.data_ptr(),
              rope.cs.shape[-1], _ptr(pos_idx), pos_offset, tokens_per_sample, hd, out.data_ptr(), 2 * d, rrms.data_ptr(), R, d, eps,
              _stream())
    return out, rrms


def qknorm_rope_bwd(dqk: Tensor, qkv: Tensor, rrms: Tensor, sq: Tensor, sk: Tensor, rope: RopeTable, hd: int,
                    dqkv: Tensor, dsq: Tensor | None, dsk: Tensor | None, *, tokens_per_sample: int, pos_offset: int = 0,
                    pos_idx: Tensor | None = None) -> None:
    """Writes dq, dk into dqkv[:, :2d] (dv at [:, 2d:] is produced by attention bwd); accumulates dsq/dsk (fp32)."""
    R = _rows(qkv)
    d = qkv.shape[-1] // 3
    _lib_call("dlb_qknorm_rope_bwd", dqk.data_ptr(), 2 * d, qkv.data_ptr(), 3 * d, sq.data_ptr(), sk.data_ptr(),
              rope.cs.data_ptr(), rope.cs.shape[-1], _ptr(pos_idx), pos_offset, tokens_per_sample, hd,
              rrms.data_ptr(), dqkv.data_ptr(), 3 * d, _ptr(dsq), _ptr(dsk), R, d, _stream())


# ---------------------------------------------------------------------------------------------------------
# attention over one or two packed segments
# ---------------------------------------------------------------------------------------------------------
class AttnSegSpec:
    """One sequence segment: qk [B*len, 2d] (rotated q|k), qkv [B*len, 3d] (v read in place), out [B*len, d]."""

    def __init__(self, qk: Tensor, qkv: Tensor, length: int):
        self.qk, self.qkv, self.len = qk, qkv, length
        self.d = qk.shape[-1] // 2


def _seg_array(specs, outs, douts=None, dqks=None, dqkvs=None):
    arr = (_lib.AttnSeg * len(specs))()
    for i, s in enumerate(specs):
        d = s.d
        e = arr[i]
        if isinstance(s, AttnSegViews):
            e.q, e.k, e.v = s.q.data_ptr(), s.k.data_ptr(), s.v.data_ptr()
            e.ldq, e.ldk, e.ldv = s.q.stride(0), s.k.stride(0), s.v.stride(0)
            e.o, e.ldo, e.len = outs[i].data_ptr(), outs[i].stride(0), s.len
            if douts is not None:
                dq, dk, dv = dqks[i]
                e.dout, e.lddo = douts[i].data_ptr(), douts[i].stride(0)
                e.dq, e.dk, e.dv = dq.data_ptr(), dk.data_ptr(), dv.data_ptr()
                e.lddq, e.lddk, e.lddv = dq.stride(0), dk.stride(0), dv.stride(0)
            continue
        e.q = s.qk.data_ptr()
        e.k = s.qk.data_ptr() + d * 2
        e.v = s.qkv.data_ptr() + 2 * d * 2
        e.ldq = e.ldk = 2 * d
        e.ldv = 3 * d
        e.o = outs[i].data_ptr()
        e.ldo = d
        e.len = s.len
        if douts is not None:
            e.dout = douts[i].data_ptr()
            e.lddo = d
            e.dq = dqks[i].data_ptr()
            e.dk = dqks[i].data_ptr() + d * 2
            e.lddq = e.lddk = 2 * d
            e.dv = dqkvs[i].data_ptr() + 2 * d * 2
            e.lddv = 3 * d
    return arr


class AttnSegViews:
    """One sequence segment given as explicit 2-D bf16 views q, k, v [B*len, d] (unit inner stride, any row stride): the general
    form of AttnSegSpec (cross-attention layouts such as the PerceiverResampler's)."""

    def __init__(self, q: Tensor, k: Tensor, v: Tensor, length: int):
        for name, t in (("q", q), ("k", k), ("v", v)):
            _check_2d(t, name)
        self.q, self.k, self.v, self.len = q, k, v, length
        self.d = q.shape[-1]


ATTN_FWD_IMPL = "dlb_attn_fwd_tc"  # tcgen05 forward; "dlb_attn_fwd" is the mma.sync kernel (kept for comparison tests)


def attn_fwd(specs: list[AttnSegSpec], B: int, H: int, hd: int, scale: float, kmask: Tensor | None = None, impl: str | None = None):
    S = sum(s.len for s in specs)
    dev = (specs[0].q if isinstance(specs[0], AttnSegViews) else specs[0].qk).device
    outs = [torch.empty(B * s.len, s.d, device=dev, dtype=BF16) for s in specs]
    lse = torch.empty(B, H, S, device=dev, dtype=F32)
    arr = _seg_array(specs, outs)
    mask_len = 0
    if kmask is not None:
        _req(kmask, torch.uint8, "kmask")
        mask_len = kmask.shape[1]
    import ctypes as C
    _lib_call(impl or ATTN_FWD_IMPL, C.cast(arr, C.c_void_p), len(specs), lse.data_ptr(), _ptr(kmask), mask_len, B, H, hd, scale, _stream())
    return outs, lse


ATTN_BWD_IMPL = "dlb_attn_bwd_tc"  # tcgen05 backward; "dlb_attn_bwd" is the mma.sync kernel pair (comparison tests)


def attn_bwd(specs: list[AttnSegSpec], outs: list[Tensor], douts: list[Tensor], lse: Tensor, B: int, H: int, hd: int,
             scale: float, dqkvs: list[Tensor], kmask: Tensor | None = None, impl: str | None = None) -> list[Tensor]:
    """Returns dqk (grad wrt rotated q|k) per segment; writes dv into dqkvs[i][:, 2d:]."""
    dev = specs[0].qk.device
    dqks = [torch.empty(B * s.len, 2 * s.d, device=dev, dtype=BF16) for s in specs]
    dsum = torch.empty_like(lse)
    arr = _seg_array(specs, outs, douts, dqks, dqkvs)
    mask_len = kmask.shape[1] if kmask is not None else 0
    import ctypes as C
    _lib_call(impl or ATTN_BWD_IMPL, C.cast(arr, C.c_void_p), len(specs), lse.data_ptr(), dsum.data_ptr(), _ptr(kmask), mask_len,
              B, H, hd, scale, _stream())
    return dqks


def attn_bwd_views(specs: list[AttnSegViews], outs: list[Tensor], douts: list[Tensor], lse: Tensor, B: int, H: int, hd: int, scale: float,
                   kmask: Tensor | None = None) -> list[tuple[Tensor, Tensor, Tensor]]:
    """Backward of attn_fwd over AttnSegViews segments -> per segment (dq, dk, dv), each [B*len, d] bf16."""
    dev = specs[0].q.device
    grads = [tuple(torch.empty(B * s.len, s.d, device=dev, dtype=BF16) for _ in range(3)) for s in specs]
    dsum = torch.empty_like(lse)
    arr = _seg_array(specs, outs, douts, grads, None)
    mask_len = kmask.shape[1] if kmask is not None else 0
    import ctypes as C
    _lib_call(ATTN_BWD_IMPL, C.cast(arr, C.c_void_p), len(specs), lse.data_ptr(), dsum.data_ptr(), _ptr(kmask), mask_len, B, H, hd, scale, _stream())
    return grads


# ---------------------------------------------------------------------------------------------------------
# glue
# ---------------------------------------------------------------------------------------------------------
def cast_bf16(x: Tensor, ld_out: int | None = None) -> Tensor:
    """fp32 [rows, cols] (or any shape, treated flat) -> bf16; ld_out > cols zero-pads each row."""
    _req(x, F32, "x")
    if ld_out is None or ld_out == x.shape[-1]:
        out = torch.empty_like(x, dtype=BF16)
        _lib_call("dlb_cast_f32_bf16", x.data_ptr(), out.data_ptr(), 1, x.numel(), x.numel(), _stream())
        return out
    rows, cols = _rows(x), x.shape[-1]
    out = torch.empty(rows, ld_out, device=x.device, dtype=BF16)
    _lib_call("dlb_cast_f32_bf16", x.data_ptr(), out.data_ptr(), rows, cols, ld_out, _stream())
    return out


def cast_f32(x: Tensor) -> Tensor:
    _req(x, BF16, "x")
    out = torch.empty_like(x, dtype=F32)
    _lib_call("dlb_cast_bf16_f32", x.data_ptr(), out.data_ptr(), x.numel(), _stream())
    return out


def _dt(t: Tensor) -> int:
    if t.dtype == BF16:
        return 0
    if t.dtype == F32:
        return 1
    raise ValueError(f"unsupported dtype {t.dtype}")


def add(a: Tensor, b: Tensor) -> Tensor:
    _req(a, BF16, "a")
    _req(b, BF16, "b")
    out = torch.empty_like(a)
    _lib_call("dlb_add_bf16", a.data_ptr(), b.data_ptr(), out.data_ptr(), a.numel(), _stream())
    return out


def gelu_fwd(x: Tensor) -> Tensor:
    _req(x, BF16, "x")
    y = torch.empty_like(x)
    _lib_call("dlb_gelu_fwd", x.data_ptr(), y.data_ptr(), x.numel(), _stream())
    return y


def gelu_bwd(dy: Tensor, x: Tensor) -> Tensor:
    _req(dy, BF16, "dy")
    _req(x, BF16, "x")
    dx = torch.empty_like(x)
    _lib_call("dlb_gelu_bwd", dy.data_ptr(), x.data_ptr(), dx.data_ptr(), x.numel(), _stream())
    return dx


def rope_apply(x: Tensor, rope: "RopeTable", hd: int, *, tokens_per_sample: int, pos_offset: int = 0, pos_idx: Tensor | None = None,
               inverse: bool = False, out: Tensor | None = None) -> Tensor:
    """RoPE (no norm / scale) on the heads of a 2-D bf16 [R, d] view (row stride allowed); inverse = transposed rotation."""
    _check_2d(x, "x")
    if x.dtype != BF16:
        raise ValueError("rope_apply: expected bfloat16")
    R, d = x.shape
    if out is None:
        out = torch.empty(R, d, device=x.device, dtype=BF16)
    _lib_call("dlb_rope_apply", x.data_ptr(), x.stride(0), out.data_ptr(), out.stride(0), rope.cs.data_ptr(), rope.cs.shape[-1], _ptr(pos_idx),
              pos_offset, tokens_per_sample, hd, d, R, int(inverse), _stream())
    return out


def bias_silu_fwd(x: Tensor, v: Tensor) -> Tensor:
    """silu(x[b,n,:] + v[b,:]) for x [B,N,d], v [B,d] (bf16)."""
    _req(x, BF16, "x")
    _req(v, BF16, "v")
    B, N, d = x.shape
    out = torch.empty_like(x)
    _lib_call("dlb_bias_silu_fwd", x.data_ptr(), v.data_ptr(), out.data_ptr(), B, N, d, _stream())
    return out


def bias_silu_bwd(dy: Tensor, x: Tensor, v: Tensor) -> tuple[Tensor, Tensor]:
    B, N, d = x.shape
    dx = torch.empty_like(x)
    dv = torch.zeros(B, d, device=x.device, dtype=F32)
    _lib_call("dlb_bias_silu_bwd", dy.data_ptr(), x.data_ptr(), v.data_ptr(), dx.data_ptr(), dv.data_ptr(), B, N, d, _stream())
    return dx, dv


def silu_fwd(x: Tensor) -> Tensor:
    _req(x, x.dtype, "x")
    y = torch.empty_like(x, dtype=BF16)
    _lib_call("dlb_silu_fwd", x.data_ptr(), _dt(x), y.data_ptr(), x.numel(), _stream())
    return y


def silu_bwd(dy: Tensor, x: Tensor, out_dtype: torch.dtype) -> Tensor:
    _req(dy, dy.dtype, "dy")
    dx = torch.empty_like(x, dtype=out_dtype)
    _lib_call("dlb_silu_bwd", dy.data_ptr(), _dt(dy), x.data_ptr(), _dt(x), dx.data_ptr(), _dt(dx), x.numel(), _stream())
    return dx


def timestep_embed(t: Tensor, dim: int, max_period: float = 10000.0) -> Tensor:
    _req(t, F32, "t")
    out = torch.empty(t.shape[0], dim, device=t.device, dtype=BF16)
    _lib_call("dlb_timestep_embed", t.data_ptr(), out.data_ptr(), t.shape[0], dim, max_period, _stream())
    return out


def cond_combine(te: Tensor, table: Tensor | None, labels: Tensor | None) -> tuple[Tensor, Tensor]:
    _req(te, BF16, "te")
    B, E = te.shape
    emb = torch.empty(B, E, device=te.device, dtype=F32)
    emb_silu = torch.empty(B, E, device=te.device, dtype=BF16)
    _lib_call("dlb_cond_combine", te.data_ptr(), _ptr(table), _ptr(labels), emb.data_ptr(), emb_silu.data_ptr(), B, E, _stream())
    return emb, emb_silu


def embedding_bwd(g: Tensor, labels: Tensor, dtable: Tensor) -> None:
    _req(g, F32, "g")
    _lib_call("dlb_embedding_bwd", g.data_ptr(), labels.data_ptr(), dtable.data_ptr(), g.shape[0], g.shape[1], _stream())


def patchify(x: Tensor, p: int) -> Tensor:
    _req(x, F32, "x")
    B, C, H, W = x.shape
    Kp = (C * p * p + 7) // 8 * 8
    out = torch.empty(B * (H // p) * (W // p), Kp, device=x.device, dtype=BF16)
    _lib_call("dlb_patchify", x.data_ptr(), out.data_ptr(), B, C, H, W, p, Kp, _stream())
    return out


def unpatchify(tok: Tensor, B: int, C: int, H: int, W: int, p: int, out_dtype: torch.dtype = BF16) -> Tensor:
    _req(tok, BF16, "tok", contiguous=False)
    img = torch.empty(B, C, H, W, device=tok.device, dtype=out_dtype)
    _lib_call("dlb_unpatchify", tok.data_ptr(), tok.stride(-2), img.data_ptr(), 0 if out_dtype == BF16 else 1, B, C, H, W, p, _stream())
    return img


def patchify_grad(img: Tensor, p: int, ld: int | None = None) -> Tensor:
    B, C, H, W = img.shape
    ppc = p * p * C
    ld = ld or (ppc + 7) // 8 * 8
    if ld != ppc:
        tok = torch.zeros(B * (H // p) * (W // p), ld, device=img.device, dtype=BF16)
    else:
        tok = torch.empty(B * (H // p) * (W // p), ld, device=img.device, dtype=BF16)
    _lib_call("dlb_patchify_grad", img.data_ptr(), _dt(img), tok.data_ptr(), ld, B, C, H, W, p, _stream())
    return tok


def colsum_(x: Tensor, out: Tensor) -> None:
    """out[c] += sum_r x[r, c]; x bf16/fp32 2-D (row stride allowed), out fp32 [C]."""
    _lib_call("dlb_colsum", x.data_ptr(), _dt(x), x.stride(0), out.data_ptr(), x.shape[0], x.shape[1], _stream())


# ---------------------------------------------------------------------------------------------------------
# formalisation-side kernels
# ---------------------------------------------------------------------------------------------------------
def interp(x0: Tensor, eps: Tensor, a: Tensor, b: Tensor) -> Tensor:
    _req(x0, F32, "x0")
    _req(eps, F32, "eps")
    xt = torch.empty_like(x0)
    B = x0.shape[0]
    _lib_call("dlb_interp", x0.data_ptr(), eps.data_ptr(), a.data_ptr(), b.data_ptr(), xt.data_ptr(), B, x0.numel() // B, _stream())
    return xt


def mse_fwd(pred: Tensor, x0: Tensor | None, eps: Tensor, xt: Tensor | None = None, t: Tensor | None = None) -> Tensor:
    _req(pred, pred.dtype, "pred")
    B = pred.shape[0]
    loss = torch.zeros((), device=pred.device, dtype=F32)
    _lib_call("dlb_mse_fwd", pred.data_ptr(), _dt(pred), _ptr(x0), eps.data_ptr(), _ptr(xt), _ptr(t), B, pred.numel() // B,
              loss.data_ptr(), _stream())
    return loss


def mse_bwd(pred: Tensor, x0: Tensor | None, eps: Tensor, gout: Tensor | None, xt: Tensor | None = None,
            t: Tensor | None = None) -> Tensor:
    B = pred.shape[0]
    dpred = torch.empty_like(pred)
    _lib_call("dlb_mse_bwd", pred.data_ptr(), _dt(pred), _ptr(x0), eps.data_ptr(), _ptr(xt), _ptr(t), B, pred.numel() // B,
              _ptr(gout), dpred.data_ptr(), _stream())
    return dpred


def euler_maruyama_step(x: Tensor, v: Tensor, noise: Tensor | None, x_prev_in: Tensor | None, c: float, one_minus_t: float, dt: float,
                        t_curr: float, std: float) -> tuple[Tensor, Tensor, Tensor, Tensor]:
    """-> (x_prev, x_prev_mean, estimated_x0, logprob); exactly one of noise / x_prev_in."""
    _req(x, F32, "x")
    v = v.contiguous()
    for name, t in (("noise", noise), ("x_prev", x_prev_in)):
        if t is not None:
            _req(t, F32, name)
    x_prev, mean, x0, logprob = (torch.empty_like(x) for _ in range(4))
    _lib_call("dlb_euler_maruyama_step", x.data_ptr(), v.data_ptr(), _dt(v), _ptr(noise), _ptr(x_prev_in), float(c), float(one_minus_t),
              float(dt), float(t_curr), float(std), x_prev.data_ptr(), mean.data_ptr(), x0.data_ptr(), logprob.data_ptr(), x.numel(), _stream())
    return x_prev, mean, x0, logprob


def gaussian_step(pred: Tensor, xt: Tensor, noise: Tensor, table: Tensor, t: Tensor, sampler: int, mean_type: int, clamp: bool,
                  eta: float, want_logprob: bool, var_mode: int = 0) -> tuple[Tensor, Tensor, Tensor, Tensor | None, Tensor | None]:
    """One fused DDPM (sampler 0) / DDIM (sampler 1) reverse step -> (x_prev, x0, mean, logprob, std).
    var_mode 1 / 2 (learned / learned_range): `pred` holds 2C channels [mean prediction | variance head] and `std` is the
    per-element standard deviation (DDPM); var_mode 0: std is None (it comes from the schedule table)."""
    _req(xt, F32, "xt")
    _req(noise, F32, "noise")
    _req(table, F32, "table")
    if t.dtype != torch.int32 or not t.is_contiguous() or not t.is_cuda:
        raise ValueError("gaussian_step: timesteps must be a contiguous int32 CUDA tensor")
    pred = pred.contiguous()
    B = xt.shape[0]
    if pred.numel() != (2 if var_mode else 1) * xt.numel():
        raise ValueError(f"gaussian_step: prediction has {pred.numel()} elements, expected {(2 if var_mode else 1) * xt.numel()}")
    x_prev, x0, mean = torch.empty_like(xt), torch.empty_like(xt), torch.empty_like(xt)
    logprob = torch.empty_like(xt) if want_logprob else None
    std = torch.empty_like(xt) if (var_mode and sampler == 0) else None
    _lib_call("dlb_gaussian_step", pred.data_ptr(), _dt(pred), xt.data_ptr(), noise.data_ptr(), table.data_ptr(), t.data_ptr(),
              int(sampler), int(mean_type), int(var_mode), int(bool(clamp)), float(eta), B, xt.numel() // B, x_prev.data_ptr(), x0.data_ptr(),
              mean.data_ptr(), _ptr(logprob), _ptr(std), _stream())
    return x_prev, x0, mean, logprob, std


def repa_cos_fwd(s: Tensor, z: Tensor, coeff: float) -> Tensor:
    _req(s, BF16, "s")
    _req(z, F32, "z")
    loss = torch.zeros((), device=s.device, dtype=F32)
    _lib_call("dlb_repa_cos_fwd", s.data_ptr(), z.data_ptr(), _rows(s), s.shape[-1], coeff, loss.data_ptr(), _stream())
    return loss


def repa_cos_bwd(s: Tensor, z: Tensor, coeff: float, gout: Tensor | None) -> Tensor:
    ds = torch.empty_like(s)
    _lib_call("dlb_repa_cos_bwd", s.data_ptr(), z.data_ptr(), _rows(s), s.shape[-1], coeff, _ptr(gout), ds.data_ptr(), _stream())
    return ds


def sprint_select(scores: Tensor, k: int) -> tuple[Tensor, Tensor, Tensor]:
    """scores fp32 [B,S] -> kept int64 [B,k] ascending, kept32 int32 [B,k], inv int32 [B,S] (slot or -1)."""
    _req(scores, F32, "scores")
    B, S = scores.shape
    kept = torch.empty(B, k, device=scores.device, dtype=torch.int64)
    kept32 = torch.empty(B, k, device=scores.device, dtype=torch.int32)
    inv = torch.empty(B, S, device=scores.device, dtype=torch.int32)
    _lib_call("dlb_sprint_select", scores.data_ptr(), B, S, k, kept.data_ptr(), kept32.data_ptr(), inv.data_ptr(), _stream())
    return kept, kept32, inv


def gather_rows(x: Tensor, idx: Tensor) -> Tensor:
    _req(x, BF16, "x")
    B, S, d = x.shape
    k = idx.shape[1]
    out = torch.empty(B, k, d, device=x.device, dtype=BF16)
    _lib_call("dlb_gather_rows", x.data_ptr(), idx.data_ptr(), out.data_ptr(), B, S, k, d, _stream())
    return out


def restore_rows(xk: Tensor, inv: Tensor, fill: Tensor | None, drop: Tensor | None) -> Tensor:
    _req(xk, BF16, "xk")
    B, k, d = xk.shape
    S = inv.shape[1]
    out = torch.empty(B, S, d, device=xk.device, dtype=BF16)
    _lib_call("dlb_restore_rows", xk.data_ptr(), inv.data_ptr(), _ptr(fill), _ptr(drop), out.data_ptr(), B, S, k, d, _stream())
    return out


def restore_rows_bwd(dy: Tensor, idx: Tensor, inv: Tensor, drop: Tensor | None, dfill: Tensor | None) -> Tensor:
    _req(dy, BF16, "dy")
    B, S, d = dy.shape
    k = idx.shape[1]
    dxk = torch.empty(B, k, d, device=dy.device, dtype=BF16)
    _lib_call("dlb_restore_rows_bwd", dy.data_ptr(), idx.data_ptr(), inv.data_ptr(), _ptr(drop), dxk.data_ptr(), _ptr(dfill),
              B, S, k, d, _stream())
    return dxk


def euler_step(x: Tensor, vc: Tensor, vu: Tensor | None, guidance: float, t_curr: float, t_prev: float,
               want_x0: bool = True, want_v: bool = False):
    """x_prev = x - v dt (dt = t_curr - t_prev), estimated_x0 = x - v t_curr with v = vc, or vu + guidance (vc - vu)
    when vu is given, in one launch. Returns (x_prev, x0 | None) and, with want_v, also the combined v (fp32)."""
    _req(x, F32, "x")
    _req(vc, vc.dtype, "vc")
    x_prev = torch.empty_like(x)
    x0 = torch.empty_like(x) if want_x0 else None
    v_out = torch.empty_like(x) if want_v else None
    _lib_call("dlb_euler_step", x.data_ptr(), vc.data_ptr(), _ptr(vu), _dt(vc), guidance, t_curr, t_prev, x_prev.data_ptr(),
              _ptr(x0), _ptr(v_out), x.numel(), _stream())
    if want_v:
        return x_prev, x0, v_out
    return x_prev, x0


def adamw_step(p: Tensor, g: Tensor, m: Tensor, v: Tensor, shadow: Tensor | None, *, lr: float, beta1: float, beta2: float,
               eps: float, weight_decay: float, step: int, grad_scale: float = 1.0, ema: Tensor | None = None,
               ema_decay: float = 0.0, chunk_active: Tensor | None = None) -> None:
    """chunk_active: uint8 [ceil(n / 64)], 0 = leave that 64-element chunk untouched (torch skips grad-less parameters)."""
    _lib_call("dlb_adamw_step", p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), _ptr(shadow), _ptr(ema), ema_decay,
              _ptr(chunk_active), p.numel(), float(lr), float(beta1), float(beta2), float(eps), float(weight_decay), step, grad_scale,
              _stream())


def reduce_pieces_(own: Tensor, staged: Tensor, stage_stride: int, world: int, rank: int, scale: float, max_ctas: int = 0) -> None:
    """own = scale * (own + the world-1 peer pieces staged by the copy engines), summed in rank order (dlb_reduce_pieces)."""
    _req(own, F32, "own")
    _req(staged, F32, "staged")
    _lib_call("dlb_reduce_pieces", own.data_ptr(), staged.data_ptr(), int(stage_stride), int(world), int(rank), own.numel(), float(scale),
              int(max_ctas), _stream())


def multimem_allreduce_(mc_ptr: int, n: int, scale: float, ctas: int) -> None:
    """In-switch (NVLS) sum of n floats at multicast address mc_ptr over all ranks, scaled, written back to every rank."""
    _lib_call("dlb_multimem_allreduce", int(mc_ptr), int(n), float(scale), int(ctas), _stream())


def ema_lerp_(ema: Tensor, p: Tensor, decay: float) -> None:
    """ema = ema * decay + p * (1 - decay), flat fp32 buffers."""
    _req(ema, F32, "ema")
    _req(p, F32, "p")
    _lib_call("dlb_ema_lerp", ema.data_ptr(), p.data_ptr(), float(decay), ema.numel(), _stream())
