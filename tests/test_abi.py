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
    assert lib.mpb_abi_version() == 5


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


def _trac_env():
    import os
    ref = ROOT / "oracle" / "_ref"
    trac, data = ref / "bin" / "trac_shared", ref / "data"
    if not (trac.exists() and (data / "ei_2011_06_05_00.nc").exists() and (ref / "deps" / "include").exists()):
        pytest.skip("oracle/_ref (the built reference) is not present on this machine")
    return trac, data, dict(os.environ, OMP_NUM_THREADS="2", LANG="C", LC_ALL="C")


def _dt_test_dir(tmp_path, data):
    d = tmp_path / "data"
    d.mkdir()
    (d / "trac.ctl").write_text(f"NQ = 0\nMETBASE = {data}/ei\nDT_MOD = 10.0\nDT_MET = 86400.0\nT_STOP = 360547210\nMET_DT_OUT = 0\n")
    (d / "atm_in.tab").write_bytes((data / "dt_test.ref" / "atm_split.tab").read_bytes())
    (tmp_path / "dirlist").write_text(str(d) + "\n")
    return d


def test_shim_refuses_a_library_built_with_other_dimensions(tmp_path):
    """The shim reads and writes atm_t / cache_t / met_t in place, so it must be compiled with the -DNP/-DNQ/-DEX/-DEY/-DEP
    of the libmptrac.so behind it.  The oracle build exports its dimensions (oracle/ref_layout.c); a shim compiled with a
    different NP must stop with the reference's ERRMSG convention instead of touching the structs (no GPU needed: the
    check comes before the device context)."""
    import subprocess
    from mptrac_b200 import build as b
    trac, data, env = _trac_env()
    if not (b.REFERENCE / "src" / "mptrac.h").exists():
        pytest.skip("the reference's header is needed to compile a shim")
    bad = tmp_path / "libshim_np.so"
    subprocess.run(["gcc", "-O1", "-fPIC", "-shared", "-fshort-enums", "-fopenmp", "-DHAVE_INLINE", "-DNP=123456",
                    f"-I{b.REFERENCE / 'src'}", f"-I{ROOT / 'oracle' / '_ref' / 'deps' / 'include'}", f"-I{ROOT / 'include'}",
                    str(b.CSRC / "shim" / "mptrac_shim.c"), f"-L{b.LIBDIR}", "-lmptrac_b200", f"-Wl,-rpath,{b.LIBDIR}", "-ldl", "-o", str(bad)],
                   check=True)
    _dt_test_dir(tmp_path, data)
    r = subprocess.run([str(trac), str(tmp_path / "dirlist"), "trac.ctl", "atm_in.tab"], env=dict(env, LD_PRELOAD=str(bad)),
                       cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert r.returncode != 0
    assert "shim compiled with NP = 123456" in r.stdout and "MPTRAC_DEFINES" in r.stdout, r.stdout[-2000:]


def test_drop_in_fails_loudly_without_a_gpu(tmp_path):
    """no CPU fallback: on a machine without a CUDA device the unmodified `trac` with the shim pre-loaded stops with an
    error from mpb_create instead of silently running the reference's CPU path"""
    import subprocess
    from mptrac_b200 import build as b
    if has_gpu():
        pytest.skip("this machine has a GPU")
    trac, data, env = _trac_env()
    shim = b.LIBDIR / "libmptrac_b200_shim.so"
    if not shim.exists():
        pytest.skip("shim not built")
    _dt_test_dir(tmp_path, data)
    r = subprocess.run([str(trac), str(tmp_path / "dirlist"), "trac.ctl", "atm_in.tab"], env=dict(env, LD_PRELOAD=str(shim)),
                       cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert r.returncode != 0 and "mptrac_b200:" in r.stdout, r.stdout[-2000:]
