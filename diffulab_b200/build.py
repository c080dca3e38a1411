"""Build libdiffulab_b200.so (hand-written CUDA for sm_100a) in-tree with nvcc.

`python -m diffulab_b200.build` compiles every `csrc/*.cu` to an object (in parallel, incremental on
mtime) and links one C-ABI shared library next to this file. nvcc cross-compiles without a GPU.
The hardware-layout probes under `csrc/probes/` (used only by tests/ and scripts/) go into a SEPARATE
library, libdiffulab_b200_probes.so, so that the product ABI (include/diffulab_b200.h) carries no development aids.
"""

from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
OBJ = HERE / "csrc" / "build"
LIB = HERE / "libdiffulab_b200.so"
PROBES = HERE / "csrc" / "probes"
PROBES_LIB = HERE / "libdiffulab_b200_probes.so"

NVCC_FLAGS = [
    "-gencode",
    "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-O3",
    "-std=c++17",
    "-Xcompiler",
    "-fPIC,-fvisibility=hidden",
    "-I",
    str(HERE.parent / "include"),
]


def _nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(nvcc).exists():
        raise RuntimeError("nvcc not found; libdiffulab_b200 cannot be built")
    return nvcc


def _newest_header_mtime() -> float:
    hdrs = list(CSRC.glob("*.cuh")) + list((HERE.parent / "include").glob("*.h"))
    return max((h.stat().st_mtime for h in hdrs), default=0.0)


def _compile(src: Path, force: bool, verbose: bool) -> Path:
    obj = OBJ / (("probe_" if src.parent == PROBES else "") + src.stem + ".o")
    stale = force or not obj.exists() or obj.stat().st_mtime < max(src.stat().st_mtime, _newest_header_mtime())
    if stale:
        cmd = [_nvcc(), *NVCC_FLAGS, "-c", str(src), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
    return obj


def build(force: bool = False, verbose: bool = False) -> Path:
    OBJ.mkdir(parents=True, exist_ok=True)
    srcs = sorted(CSRC.glob("*.cu"))
    if not srcs:
        raise RuntimeError(f"no CUDA sources under {CSRC}")
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(lambda s: _compile(s, force, verbose), srcs))
    psrcs = sorted(PROBES.glob("*.cu"))
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        pobjs = list(ex.map(lambda s: _compile(s, force, verbose), psrcs))
    common = [o for o in objs if o.stem == "common"]
    for lib, members in ((LIB, objs), (PROBES_LIB, pobjs + common)):
        if not members:
            continue
        if force or not lib.exists() or lib.stat().st_mtime < max(o.stat().st_mtime for o in members):
            cmd = [_nvcc(), "-shared", "-o", str(lib), *map(str, members), "-gencode", "arch=compute_100a,code=sm_100a"]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(p)
