"""CPU: the C-ABI library builds / loads and exports every symbol include/diffulab_b200.h declares (no compute calls),
the ctypes table mirrors the header, and the header compiles as plain C."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "diffulab_b200.h")


def header_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dlb_[a-z0-9_]+)\s*\(", text)))


def test_header_is_plain_c(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "diffulab_b200.h"\nint main(void){ dlb_attn_seg s; (void)s; return DLB_OK; }\n')
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", str(src), "-o", str(tmp_path / "t.o")], check=True)


def test_library_exports_every_declared_symbol():
    from diffulab_b200 import _lib
    from diffulab_b200.build import build

    build()
    lib = _lib.load()  # raises if missing: there is no fallback
    syms = header_symbols()
    assert len(syms) >= 40
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in the header but not exported"
    assert set(_lib.exported_symbols()) == set(syms), set(_lib.exported_symbols()) ^ set(syms)
    assert lib.dlb_version() == 100
    assert lib.dlb_last_error() is not None


def test_probe_library_is_separate_and_matches_its_header():
    """Development probes live in libdiffulab_b200_probes.so with their own header; the product ABI carries none."""
    from diffulab_b200 import _lib

    text = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "diffulab_b200_probes.h")).read(), flags=re.S)
    declared = sorted(set(re.findall(r"\b(dlb_[a-z0-9_]+)\s*\(", text)))
    lib = _lib.load_probes()
    assert declared == sorted(_lib._PROBE_SIGNATURES)
    for s in declared:
        assert hasattr(lib, s)
    assert not any("probe" in s for s in header_symbols())


def test_attn_seg_struct_layout_matches_header():
    import ctypes as C

    from diffulab_b200 import _lib

    assert C.sizeof(_lib.AttnSeg) == 8 * 8 + 8 * 8 + 8  # 8 pointers, 8 int64, int32 + padding
    names = [f[0] for f in _lib.AttnSeg._fields_]
    assert names == ["q", "k", "v", "o", "dout", "dq", "dk", "dv", "ldq", "ldk", "ldv", "ldo", "lddo", "lddq", "lddk", "lddv", "len"]


def test_product_path_never_imports_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU arms may touch oracle/."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "diffulab_b200")):
        for f in files:
            if f.endswith(".py"):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("# oracle", ""), f"{f} references the oracle"


def test_ops_refuse_cpu_tensors():
    import pytest
    import torch

    from diffulab_b200 import ops

    a = torch.zeros(8, 8, dtype=torch.bfloat16)
    with pytest.raises(ValueError):
        ops.gemm(a, a)
    with pytest.raises(ValueError):
        ops.swiglu_fwd(a)
