#!/usr/bin/env python
"""SASS evidence for the hot kernels of libdiffulab_b200.so (runs without a GPU): per-kernel mnemonic histogram of the
Blackwell-specific instructions (UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG/UTMAREDG = TMA tensor
load / store / reduce, UBLKCP = cp.async.bulk, SYNCS = mbarrier, LDGSTS = cp.async, HMMA = legacy mma.sync) plus the
instruction window around the first tcgen05.mma of each kernel.   usage: python scripts/sass_extract.py > profiles/sass_r2.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "diffulab_b200", "libdiffulab_b200.so")
WATCH = ("UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UBLKCP", "UBLKRED", "SYNCS", "LDGSTS", "HMMA", "MUFU", "BAR", "ELECT",
         "UTCBAR", "UTCCP", "R2UR", "REDG", "ATOMG", "STG", "LDG", "LDS", "STS")
PATTERNS = sys.argv[1:] or ["gemm2_tcgen05_kernel", "gemm_tcgen05_kernel", "attn_fwd", "attn_bwd", "ln_modulate", "qknorm_rope", "gate_residual",
                            "adamw_kernel", "mse_", "repa_cos", "euler_step", "sprint", "restore_rows", "gather_rows"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kernels = re.split(r"\n\s*Function : ", out)[1:]
    print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)} — {len(kernels)} kernels; instruction counts per kernel (static SASS)")
    totals = collections.Counter()
    for k in kernels:
        name, _, body = k.partition("\n")
        name = name.strip()
        if not any(p in name for p in PATTERNS):
            continue
        demangled = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        ins = re.findall(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", body)
        hist = collections.Counter()
        for i in ins:
            base = i.split(".")[0]
            if base in WATCH:
                hist[base] += 1
        totals.update(hist)
        print(f"\n## {demangled[:150]}\n   instructions {len(ins)}; " + ", ".join(f"{k}={v}" for k, v in sorted(hist.items())))
        lines = body.split("\n")
        idx = next((i for i, l in enumerate(lines) if "UTCHMMA" in l), None)
        if idx is not None:
            code = [l for l in lines[max(0, idx - 12): idx + 14] if "/*" in l and not re.match(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", l)]
            print("   first tcgen05.mma issue site:")
            for l in code:
                print("     " + re.sub(r"\s+/\* 0x[0-9a-f]+ \*/\s*$", "", l.rstrip())[:140])
    print("\n# totals over the listed kernels: " + ", ".join(f"{k}={v}" for k, v in sorted(totals.items())))


if __name__ == "__main__":
    main()
