"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/mptrac_b200.h
declares, and refuses loudly to work without a CUDA device (no CPU fallback)."""
import ctypes as C
import re

import pytest

from conftest import ROOT, has_gpu


def header_symbols():
    txt = (ROOT / "include" / "mptrac_b200.h").read_text()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(mpb_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from mptrac_b200 import load_library
    lib = load_library()
    syms = header_symbols()
    assert len(syms) >= 35
    for s in syms:
        assert hasattr(lib, s), f"{s} is declared in include/mptrac_b200.h but not exported"
    assert lib.mpb_abi_version() == 3


def test_strict_flavour_exports_the_same_abi():
    from mptrac_b200 import load_library
    lib = load_library(strict=True)
    for s in header_symbols():
        assert hasattr(lib, s)


def test_python_mirror_binds_every_symbol():
    from mptrac_b200 import load_library
    lib = load_library()
    assert set(lib._mpb_symbols) == set(header_symbols())


def test_ctl_struct_layout_matches_header():
    # 16 int32 + 23 int32 + pad = 40 int32 = 160 bytes, then 25 doubles, then 16 int32 (ABI version 2)
    from mptrac_b200.host import _CtlStruct, _GridStruct, _MetViewStruct
    assert C.sizeof(_CtlStruct) == 160 + 25 * 8 + 16 * 4 + 2 * 4
    assert _CtlStruct.t_start.offset == 160
    assert _CtlStruct.met_dt_out.offset == 160 + 24 * 8
    assert C.sizeof(_MetViewStruct) == 8 + 16 + 9 * 8 + 3 * 8 + 8 + 6 * 8 + 2 * 8
    assert C.sizeof(_GridStruct) == 16 + 8 * 8


@pytest.mark.skipif(has_gpu(), reason="only meaningful on a machine without a GPU")
def test_no_cpu_fallback():
    from mptrac_b200 import Engine, MpbError
    with pytest.raises(MpbError, match="no CUDA device"):
        Engine(100)


def test_product_does_not_import_oracle():
    for f in (ROOT / "mptrac_b200").rglob("*"):
        if f.suffix in (".py", ".cu", ".cuh", ".c", ".h") and f.is_file():
            txt = f.read_text()
            assert "oracle" not in txt.replace("oracle/_ref", "").replace("oracle\" / \"_ref", "") or f.name == "build.py", f
