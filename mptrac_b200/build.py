"""Build the sm_100a shared library (and, when the reference tree is present, the drop-in shim).

Artefacts are written IN-TREE under ``mptrac_b200/_lib`` (git-ignored ``*.so``; they travel to the
GPU box with the repo snapshot).  nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import os
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
LIBDIR = PKG / "_lib"
LIB = LIBDIR / "libmptrac_b200.so"
LIB_STRICT = LIBDIR / "libmptrac_b200_strict.so"
SHIM = LIBDIR / "libmptrac_b200_shim.so"
REFERENCE = Path(os.environ.get("MPTRAC_REFERENCE", "/root/reference"))

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-fopenmp", "-shared",
]


def _stale(target: Path, sources) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(s).stat().st_mtime > t for s in sources)


def _run(cmd, **kw):
    print("+", " ".join(str(c) for c in cmd), flush=True)
    subprocess.run([str(c) for c in cmd], check=True, **kw)


def build_lib(force: bool = False, verbose_ptxas: bool = False) -> Path:
    """nvcc -> mptrac_b200/_lib/libmptrac_b200.so (+ the -fmad=false 'strict' flavour used by parity tests)."""
    LIBDIR.mkdir(exist_ok=True)
    srcs = [CSRC / "engine.cu", CSRC / "physics.cuh", CSRC / "met_tables.hpp", ROOT / "include" / "mptrac_b200.h"]
    extra = ["-Xptxas", "-v"] if verbose_ptxas else []
    if force or _stale(LIB, srcs):
        _run(["nvcc", *NVCC_FLAGS, *extra, CSRC / "engine.cu", "-o", LIB])
    if force or _stale(LIB_STRICT, srcs):
        # (the strict flavour also carries the TMA bulk-copy staging of the parcel stream, so that both staging paths
        # run in the GPU test suite: csrc/engine.cu MPB_TMA_STAGE)
        _run(["nvcc", *NVCC_FLAGS, "-fmad=false", "-DMPB_STRICT=1", "-DMPB_TMA_STAGE=1", CSRC / "engine.cu", "-o", LIB_STRICT])
    return LIB


def build_shim(force: bool = False):
    """The reference-facing shim needs the reference's own header (never copied): only built where it exists."""
    src = CSRC / "shim" / "mptrac_shim.c"
    hdr = REFERENCE / "src" / "mptrac.h"
    deps_inc = ROOT / "oracle" / "_ref" / "deps" / "include"
    if not (src.exists() and hdr.exists() and deps_inc.exists()):
        return None
    build_lib()
    # the shim reads and writes the driver's atm_t / cache_t / met_t in place: it must see the dimensions (-DNP, -DNQ, -DEX,
    # -DEY, -DEP) the reference library was compiled with (its build recipe takes them from the same variable); the shim
    # also checks the layout at run time against the symbol mptrac_ref_layout when the library in front of it exports one
    defines = os.environ.get("MPTRAC_DEFINES", "").split()
    stamp = LIBDIR / ".shim_defines"
    same_defines = stamp.exists() and stamp.read_text() == " ".join(defines)
    if force or not same_defines or _stale(SHIM, [src, hdr, ROOT / "include" / "mptrac_b200.h"]):
        _run(["gcc", "-O2", "-g", "-fPIC", "-shared", "-fshort-enums", "-fopenmp", "-DHAVE_INLINE", *defines,
              f"-I{REFERENCE / 'src'}", f"-I{deps_inc}", f"-I{ROOT / 'include'}", src,
              f"-L{LIBDIR}", "-lmptrac_b200", "-Wl,-rpath,$ORIGIN", "-ldl", "-o", SHIM])
        stamp.write_text(" ".join(defines))
    return SHIM


if __name__ == "__main__":
    build_lib(force="--force" in sys.argv, verbose_ptxas="-v" in sys.argv)
    build_shim(force="--force" in sys.argv)
