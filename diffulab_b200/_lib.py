"""ctypes binding of libdiffulab_b200.so (the C-ABI boundary declared in include/diffulab_b200.h).

There is no CPU or PyTorch fallback: if the shared library is missing or a call fails, this module raises.
"""

from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_HERE = Path(__file__).resolve().parent
# DIFFULAB_B200_LIB: an alternative BUILD of the same library (same exported symbols), for same-box A/B timing in scripts/
LIB_PATH = Path(os.environ["DIFFULAB_B200_LIB"]) if os.environ.get("DIFFULAB_B200_LIB") else _HERE / "libdiffulab_b200.so"

_lib: C.CDLL | None = None

p = C.c_void_p
i64 = C.c_int64
i32 = C.c_int
f32 = C.c_float
f64 = C.c_double

# name -> argtypes (restype is always int unless listed in _RESTYPES). Mirrors include/diffulab_b200.h.
_SIGNATURES: dict[str, list] = {
    "dlb_version": [],
    "dlb_last_error": [],
    "dlb_launch_count": [],
    "dlb_reset_launch_count": [],
    "dlb_device_check": [],
    "dlb_set_sm_budget": [i32],
    "dlb_reduce_pieces": [p, p, i64, i32, i32, i64, f32, i32, p],
    "dlb_multimem_allreduce": [p, i64, f32, i32, p],
    # C[M,N] (+)= A*B^T (+bias): A, B, C, bias, M, N, K, lda, ldb, ldc, a_mn, b_mn, out_mode, split_k, tile_n, stream
    "dlb_gemm_bf16": [p, p, p, p, i64, i64, i64, i64, i64, i64, i32, i32, i32, i32, i32, p],
    # x, w, b, scale, shift, mod_ld, rows_per_mod, y, mean, rstd, R, d, eps, stream
    "dlb_ln_modulate_fwd": [p, p, p, p, p, i64, i64, p, p, p, i64, i32, f32, p],
    # dy, x, mean, rstd, w, b, scale, mod_ld, groups, rows_per_group, per_token, dres, dx, dscale, dshift, dmod_ld,
    # dscale_tok, dshift_tok, dtok_ld, dw, db, d, stream
    "dlb_ln_modulate_bwd": [p, p, p, p, p, p, p, i64, i64, i64, i32, p, p, p, p, i64, p, p, i64, p, p, i32, p],
    # x, a1, a2, gate, gate_ld, rows_per_mod, out, R, d, stream
    "dlb_gate_residual_fwd": [p, p, p, p, i64, i64, p, i64, i32, p],
    # dout, a1, a2, gate, gate_ld, groups, rows_per_group, per_token, da, dgate, dgate_ld, dgate_tok, dtok_ld, d, stream
    "dlb_gate_residual_bwd": [p, p, p, p, i64, i64, i64, i32, p, p, i64, p, i64, i32, p],
    "dlb_swiglu_fwd": [p, p, i64, i32, p],
    "dlb_swiglu_bwd": [p, p, p, i64, i32, p],
    # qkv, ld_in, sq, sk, cs, rot_half, pos_idx, pos_offset, tokens_per_sample, hd, out, ld_out, rrms, R, d, eps, stream
    "dlb_qknorm_rope_fwd": [p, i64, p, p, p, i32, p, i32, i32, i32, p, i64, p, i64, i32, f32, p],
    # dqk, ld_dqk, qkv, ld_in, sq, sk, cs, rot_half, pos_idx, pos_offset, tps, hd, rrms, dqkv, ld_out, dsq, dsk, R, d, stream
    "dlb_qknorm_rope_bwd": [p, i64, p, i64, p, p, p, i32, p, i32, i32, i32, p, p, i64, p, p, i64, i32, p],
    # pos, n_axes, axis_of_pair, local_of_pair, axis_dim, base, cos, sin, cs, P, rot_half, stream
    "dlb_rope_table": [p, i32, p, p, p, C.c_double, p, p, p, i64, i32, p],
    # segs, nseg, lse, kmask, mask_len, B, H, hd, scale, stream
    "dlb_attn_fwd": [p, i32, p, p, i32, i32, i32, i32, f32, p],
    "dlb_attn_fwd_tc": [p, i32, p, p, i32, i32, i32, i32, f32, p],
    # segs, nseg, lse, dsum, kmask, mask_len, B, H, hd, scale, stream
    "dlb_attn_bwd": [p, i32, p, p, p, i32, i32, i32, i32, f32, p],
    "dlb_attn_bwd_tc": [p, i32, p, p, p, i32, i32, i32, i32, f32, p],
    "dlb_cast_f32_bf16": [p, p, i64, i64, i64, p],
    "dlb_cast_bf16_f32": [p, p, i64, p],
    "dlb_attn_set_trace": [p],
    "dlb_gemm2_bf16": [p, p, p, p, i64, i64, i64, i64, i64, i64, i32, i32, i32, i32, i32, p],
    # dY, W2, H, dH, M, F, D, lddy, ldw2, ldh, lddh, stream
    "dlb_gemm_swiglu_bwd_bf16": [p, p, p, p, i64, i64, i64, i64, i64, i64, i64, p],
    # A, W, bias, H, ACT, M, F, K, lda, ldw, ldh, ldact, stream
    "dlb_gemm_swiglu_bf16": [p, p, p, p, p, i64, i64, i64, i64, i64, i64, i64, p],
    # x, v, v_dtype, noise, x_prev_in, c, one_minus_t, dt, t_curr, std, x_prev, mean, x0_est, logprob, n, stream
    "dlb_euler_maruyama_step": [p, p, i32, p, p, f32, f32, f32, f32, f32, p, p, p, p, i64, p],
    # pred, pred_dtype, xt, noise, table, t, sampler, mean_type, clamp, eta, B, per_sample, x_prev, x0, mean, logprob, stream
    "dlb_gaussian_step": [p, i32, p, p, p, p, i32, i32, i32, i32, f32, i64, i64, p, p, p, p, p, p],
    "dlb_add_bf16": [p, p, p, i64, p],
    "dlb_gelu_fwd": [p, p, i64, p],
    "dlb_gelu_bwd": [p, p, p, i64, p],
    # x, ld_in, y, ld_out, cs, rot_half, pos_idx, pos_offset, tokens_per_sample, hd, d, R, inverse, stream
    "dlb_rope_apply": [p, i64, p, i64, p, i32, p, i32, i32, i32, i32, i64, i32, p],
    "dlb_bias_silu_fwd": [p, p, p, i64, i64, i32, p],
    "dlb_bias_silu_bwd": [p, p, p, p, p, i64, i64, i32, p],
    "dlb_silu_fwd": [p, i32, p, i64, p],
    "dlb_silu_bwd": [p, i32, p, i32, p, i32, i64, p],
    "dlb_timestep_embed": [p, p, i32, i32, f32, p],
    "dlb_cond_combine": [p, p, p, p, p, i32, i32, p],
    "dlb_embedding_bwd": [p, p, p, i32, i32, p],
    "dlb_patchify": [p, p, i32, i32, i32, i32, i32, i32, p],
    "dlb_unpatchify": [p, i64, p, i32, i32, i32, i32, i32, i32, p],
    "dlb_patchify_grad": [p, i32, p, i64, i32, i32, i32, i32, i32, p],
    "dlb_colsum": [p, i32, i64, p, i64, i32, p],
    "dlb_interp": [p, p, p, p, p, i64, i64, p],
    # pred, pred_dtype, x0, eps, xt, t, B, per_sample, loss, stream
    "dlb_mse_fwd": [p, i32, p, p, p, p, i64, i64, p, p],
    "dlb_mse_bwd": [p, i32, p, p, p, p, i64, i64, p, p, p],
    "dlb_repa_cos_fwd": [p, p, i64, i32, f32, p, p],
    "dlb_repa_cos_bwd": [p, p, i64, i32, f32, p, p, p],
    "dlb_sprint_select": [p, i32, i32, i32, p, p, p, p],
    "dlb_gather_rows": [p, p, p, i32, i32, i32, i32, p],
    "dlb_restore_rows": [p, p, p, p, p, i32, i32, i32, i32, p],
    "dlb_restore_rows_bwd": [p, p, p, p, p, p, i32, i32, i32, i32, p],
    # x, vc, vu, v_dtype, guidance, t_curr, t_prev, x_prev, x0_est, v_out, n, stream
    "dlb_euler_step": [p, p, p, i32, f32, f32, f32, p, p, p, i64, p],
    # p, g, m, v, shadow, ema, ema_decay, chunk_active, n, lr, b1, b2, eps, wd, step, grad_scale, stream
    "dlb_adamw_step": [p, p, p, p, p, p, f32, p, i64, f64, f64, f64, f64, f64, i64, f32, p],
    # ema, p, decay, n, stream
    "dlb_ema_lerp": [p, p, f32, i64, p],
}


# development probes (csrc/probes/, libdiffulab_b200_probes.so; include/diffulab_b200_probes.h): tests / scripts only
_PROBE_SIGNATURES: dict[str, list] = {
    "dlb_umma_probe": [p, p, p, i32, i32, i32, i32, i32, p],
    # base, rows, ld, H, hd, grid, tiles_per_cta, dump, stream
    "dlb_tma_gather_probe": [p, i64, i64, i32, i32, i32, i32, p, p],
    # X, Y, P, D1, D2, rows, ld, H, hd, h, xrow, yrow, dump, stream
    "dlb_attn_sw_probe": [p, p, p, p, p, i64, i64, i32, i32, i32, i32, i32, p, p],
    # X, rows, ld, H, hd, grid, tiles_per_cta, stream
    "dlb_attn_sw_stream_probe": [p, i64, i64, i32, i32, i32, i32, p],
}


class AttnSeg(C.Structure):
    """Mirror of `dlb_attn_seg` (include/diffulab_b200.h)."""

    _fields_ = (
        [(n, C.c_void_p) for n in ("q", "k", "v", "o", "dout", "dq", "dk", "dv")]
        + [(n, C.c_int64) for n in ("ldq", "ldk", "ldv", "ldo", "lddo", "lddq", "lddk", "lddv")]
        + [("len", C.c_int32)]
    )
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


_probes: C.CDLL | None = None


def load_probes() -> C.CDLL:
    """The separate development-probe library (hardware layout probes used by tests/ and scripts/ only)."""
    global _probes
    if _probes is None:
        path = _HERE / "libdiffulab_b200_probes.so"
        if not path.exists():
            raise RuntimeError(f"{path} is missing: build it with `python -m diffulab_b200.build`")
        lib = C.CDLL(str(path))
        for name, argtypes in _PROBE_SIGNATURES.items():
            fn = getattr(lib, name)
            fn.argtypes = argtypes
            fn.restype = C.c_int
        lib.dlb_last_error.restype = C.c_char_p
        _probes = lib
    return _probes


def check_probe(rc: int, what: str) -> None:
    if rc != 0:
        msg = load_probes().dlb_last_error()
        raise DlbError(f"{what} failed (rc={rc}): {msg.decode() if msg else ''}")


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


def set_sm_budget(sms: int) -> int:
    """SMs the persistent kernels use (0 = all); returns the previous budget (see include/diffulab_b200.h)."""
    return int(load().dlb_set_sm_budget(int(sms)))
