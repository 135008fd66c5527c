"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/mptrac_b200.h
declares, and refuses loudly to work without a CUDA device (no CPU fallback)."""
import ctypes as C
import re
from pathlib import Path

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
    assert lib.mpb_abi_version() == 4


def test_strict_flavour_exports_the_same_abi():
    from mptrac_b200 import load_library
    lib = load_library(strict=True)
    for s in header_symbols():
        assert hasattr(lib, s)


def test_python_mirror_binds_every_symbol():
    from mptrac_b200 import load_library
    lib = load_library()
    assert set(lib._mpb_symbols) == set(header_symbols())


def _c_layout(struct, members):
    """sizeof and offsetof of every member, from a C program compiled against include/mptrac_b200.h"""
    import shutil
    import subprocess
    import tempfile
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    body = "".join(f'  printf("{m} %zu\\n", offsetof({struct}, {m}));\n' for m in members)
    src = ('#include <stdio.h>\n#include <stddef.h>\n#include "mptrac_b200.h"\nint main(void) {\n'
           f'  printf("sizeof %zu\\n", sizeof({struct}));\n{body}  return 0;\n}}\n')
    with tempfile.TemporaryDirectory() as d:
        (Path(d) / "l.c").write_text(src)
        subprocess.run(["gcc", "-I", str(ROOT / "include"), str(Path(d) / "l.c"), "-o", str(Path(d) / "l")], check=True)
        out = subprocess.run([str(Path(d) / "l")], check=True, capture_output=True, text=True).stdout
    return {k: int(v) for k, v in (ln.split() for ln in out.strip().splitlines())}


@pytest.mark.parametrize("which", ["ctl", "met_view", "grid"])
def test_struct_layouts_match_the_header(which):
    """every member of the ctypes mirrors (mptrac_b200/host.py) sits where the C compiler puts it (ABI version 4); the
    oracle's own control structure shares the layout of mpb_ctl_t by construction, which is checked too"""
    from mptrac_b200.host import _CtlStruct, _GridStruct, _MetViewStruct
    from oracle.oracle import OrcCtl
    cls, cname = {"ctl": (_CtlStruct, "mpb_ctl_t"), "met_view": (_MetViewStruct, "mpb_met_view_t"), "grid": (_GridStruct, "mpb_grid_t")}[which]
    members = [n for n, _ in cls._fields_]
    lay = _c_layout(cname, members)
    assert lay.pop("sizeof") == C.sizeof(cls)
    for m in members:
        assert lay[m] == getattr(cls, m).offset, m
    if which == "ctl":
        assert C.sizeof(OrcCtl) == C.sizeof(_CtlStruct)
        for n, _ in OrcCtl._fields_:
            assert getattr(OrcCtl, n).offset == getattr(_CtlStruct, n).offset, n


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


def test_shim_defines_exactly_the_reference_symbols_it_replaces():
    """libmptrac_b200_shim.so (built where the reference's mptrac.h is available) interposes four symbols of the reference
     and nothing else, and resolves the C ABI from libmptrac_b200.so next to it"""
    import shutil
    import subprocess
    shim = ROOT / "mptrac_b200" / "_lib" / "libmptrac_b200_shim.so"
    if not shim.exists() or shutil.which("nm") is None:
        pytest.skip("shim not built on this machine (needs the reference's mptrac.h)")
    out = subprocess.run(["nm", "-D", str(shim)], check=True, capture_output=True, text=True).stdout
    defined = {ln.split()[-1] for ln in out.splitlines() if " T " in ln}
    assert defined == {"mptrac_free", "mptrac_run_timestep", "mptrac_update_device", "mptrac_update_host"}
    undefined = {ln.split()[-1] for ln in out.splitlines() if " U " in ln}
    used = {s for s in undefined if s.startswith("mpb_")}
    assert used and used <= set(header_symbols())
    # the reference modules it delegates to in hybrid mode come from the reference's own library at run time
    assert {"module_diff_pbl", "module_convection", "module_isosurf", "module_meteo", "module_bound_cond", "module_decay"} <= undefined
