"""End-to-end drop-in test: the reference's UNMODIFIED `trac` driver (oracle/_ref/bin/trac_shared, linked against the
reference's own libmptrac.so) with libmptrac_b200_shim.so pre-loaded, on the reference's own tests/dt_test and
tests/coord_test inputs, compared with the reference's shipped goldens (data.ref/*.tab).  Needs a GPU and oracle/_ref
(built in the development container and shipped with the repository snapshot)."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, abserr, relerr

pytestmark = pytest.mark.gpu

REFDIR = ROOT / "oracle" / "_ref"
TRAC = REFDIR / "bin" / "trac_shared"
SHIM = ROOT / "mptrac_b200" / "_lib" / "libmptrac_b200_shim.so"
DATA = REFDIR / "data"


def _need():
    for f in (TRAC, SHIM, DATA / "ei_2011_06_05_00.nc"):
        if not f.exists():
            pytest.skip(f"{f} not present (built only where /root/reference exists)")


def _run_trac(tmp_path, ctl_text, atm_in, extra, preload=True, env_extra=None, atm_name="atm_in.tab"):
    d = tmp_path / "data"
    d.mkdir()
    (d / "trac.ctl").write_text(ctl_text)
    (d / atm_name).write_bytes(atm_in.read_bytes())
    (tmp_path / "dirlist").write_text(str(d) + "\n")
    env = dict(os.environ, OMP_NUM_THREADS="4", LANG="C", LC_ALL="C", MPTRAC_B200_VERBOSE="1")
    if preload:
        env["LD_PRELOAD"] = str(SHIM)
    env.update(env_extra or {})
    r = subprocess.run([str(TRAC), str(tmp_path / "dirlist"), "trac.ctl", atm_name, *extra], env=env, cwd=tmp_path,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    return d, r.stdout


def _tab(f):
    return np.loadtxt(f, comments="#", ndmin=2)


DT_CTL = """NQ = 4
QNT_NAME[0] = t
QNT_NAME[1] = u
QNT_NAME[2] = v
QNT_NAME[3] = w
METBASE = {met}/ei
ATM_DT_OUT = 10.0
DT_MOD = 10.0
DIFFUSION = 1
DT_MET = 86400.0
T_STOP = 360547260
"""


@pytest.mark.parametrize("mode", ["as_shipped", "hybrid", "device_only"])
def test_trac_dt_test_through_the_shim(tmp_path, mode):
    """as_shipped: the control file of tests/dt_test unchanged -- module_meteo (quantities t, u, v, w) runs on the
    device after every step -> all 8 columns of the shipped goldens without any per-step host round trip.
    hybrid: the same with module_meteo forced onto the reference's CPU code between device steps (the path every
    module that is not on the device takes).  device_only: MET_DT_OUT 0 -> every step is one fused launch; columns 1-4."""
    _need()
    extra = ["ATM_BASENAME", "atm_pl"] + (["MET_DT_OUT", "0"] if mode == "device_only" else [])
    d, out = _run_trac(tmp_path, DT_CTL.format(met=DATA), DATA / "dt_test.ref" / "atm_split.tab", extra,
                       env_extra={"MPTRAC_B200_HOST_METEO": "1"} if mode == "hybrid" else None)
    assert "mptrac_b200:" in out and "kernel launches" in out, "the shim was not in the call path"
    gold = sorted((DATA / "dt_test.ref").glob("atm_pl_*.tab"))
    assert len(gold) == 7
    for g in gold:
        a, b = _tab(d / g.name), _tab(g)
        assert a.shape == b.shape
        ncol = 4 if mode == "device_only" else 8
        assert abserr(a[:, 0], b[:, 0]) < 0.006
        for c in range(1, ncol):
            # %g text: 6 significant digits; a last-digit flip is 1e-5 relative at worst
            assert relerr(a[:, c], b[:, c]) < 2e-5 or abserr(a[:, c], b[:, c]) < 1e-9, (g.name, c)


def test_trac_coord_test_through_the_shim(tmp_path):
    _need()
    ctl = """NQ = 4
QNT_NAME[0] = t
QNT_NAME[1] = u
QNT_NAME[2] = v
QNT_NAME[3] = w
METBASE = {met}/era5_utm32
TRACER_CHEM = 0
DIFFUSION = 1
DT_MET = 3600.0
T_STOP = 799380000
""".format(met=DATA)
    extra = ["ATM_BASENAME", "atm", "MET_CAPE", "0", "DT_MOD", "600", "ATM_DT_OUT", "600", "MET_COORD_TYPE", "1",
             "MET_UTM_REF_LON", "11.5692782", "MET_UTM_REF_LAT", "48.1507476"]
    # the shipped t0 snapshot is the (rounded) initial state; run the CPU reference from the same file for a tight check
    d, out = _run_trac(tmp_path, ctl, DATA / "coord_test.ref" / "atm_2025_05_01_00_00_00.tab", extra)
    assert "kernel launches" in out
    cpu = tmp_path / "cpu"
    cpu.mkdir()
    dc, _ = _run_trac(cpu, ctl, DATA / "coord_test.ref" / "atm_2025_05_01_00_00_00.tab", extra, preload=False)
    files = sorted(d.glob("atm_2025_05_01_*.tab"))
    assert len(files) == 13
    for f in files:
        a, b, g = _tab(f), _tab(dc / f.name), _tab(DATA / "coord_test.ref" / f.name)
        assert abserr(a[:, 2], b[:, 2]) < 0.011 and abserr(a[:, 3], b[:, 3]) < 0.011 and relerr(a[:, 1], b[:, 1]) < 2e-5
        assert abserr(a[:, 2], g[:, 2]) < 0.03 and abserr(a[:, 3], g[:, 3]) < 0.03


TRAC_TEST_CTL = """NQ = 13
QNT_NAME[0] = t
QNT_NAME[1] = u
QNT_NAME[2] = v
QNT_NAME[3] = w
QNT_NAME[4] = zg
QNT_NAME[5] = pv
QNT_NAME[6] = ps
QNT_NAME[7] = pt
QNT_NAME[8] = m
QNT_NAME[9] = stat
QNT_NAME[10] = ens
QNT_NAME[11] = Cccl3f
QNT_NAME[12] = Cx
METBASE = {met}/ei
MET_DT_OUT = 86400.0
SPECIES = SO2
BOUND_LAT0 = -90
BOUND_LAT1 = 90
BOUND_P0 = 1e10
BOUND_P1 = -1e10
BOUND_DPS = 100.0
BOUND_MASS = 0.0
CONV_CAPE = 0.0
H2O2_CHEM_REACTION = 1
TRACER_CHEM = 1
CHEMGRID_NX = 72
CHEMGRID_NY = 36
CHEMGRID_NZ = 30
DIFFUSION = 1
TDEC_TROP = 259200.0
TDEC_STRAT = 259200.0
DRY_DEPO_VDEP = 0.15
DRY_DEPO_DP = 300
MIXING_TROP = 1e-3
MIXING_STRAT = 1e-6
DT_MET = 86400.0
DT_MOD = 300.0
T_STOP = 360806400
"""


@pytest.mark.timeout(900)
@pytest.mark.parametrize("levels", ["pl", "ml"])
def test_trac_trac_test_through_the_shim(tmp_path, levels):
    """The reference's own end-to-end test, tests/trac_test: 10 000 parcels, 3 days at 300 s, with diffusion, convection,
    H2O2 + tracer chemistry, decay, dry deposition, mixing and boundary conditions all on -- the modules of the path on
    the GPU, the other twelve through the reference's CPU code in between (hybrid mode, one random number stream across
    both).  `pl` is the pressure-level run, `ml` the run on ERA5 model levels (MET_VERT_COORD 1, ADVECT_VERT_COORD 2:
    omega on model levels, tests/trac_test/run.sh:98-108).  Compared with the goldens the reference ships
    (data.ref/atm_pl_*.tab, atm_ml_*.tab: %g text)."""
    _need()
    gold = sorted((DATA / "trac_test.ref").glob(f"atm_{levels}_2011_*.tab"))
    met = "ei" if levels == "pl" else "era5ml"
    if len(gold) != 4 or not (DATA / f"{met}_2011_06_08_00.nc").exists() or not (DATA / "clim" / "cams_H2O2.nc").exists():
        pytest.skip("trac_test data not shipped (oracle/build_ref.sh)")
    # the chemistry modules read their climatologies from ../../data relative to the working directory, like the
    # reference's own test does from tests/trac_test
    (tmp_path / "data").symlink_to(DATA / "clim")
    base = tmp_path / "tests" / "trac_test"
    base.mkdir(parents=True)
    extra = ["ATM_BASENAME", f"atm_{levels}", "STAT_BASENAME", f"station_{levels}", "STAT_LON", "-22", "STAT_LAT", "-40"]
    if levels == "ml":
        extra += ["METBASE", f"{DATA}/era5ml", "MET_PRESS_LEVEL_DEF", "6", "MET_VERT_COORD", "1", "ADVECT_VERT_COORD", "2"]
    d, out = _run_trac(base, TRAC_TEST_CTL.format(met=DATA), DATA / "trac_test.ref" / "atm_init.tab", extra)
    assert "kernel launches" in out
    report = {}
    for g in gold:
        a, b = _tab(d / g.name), _tab(g)
        assert a.shape == b.shape == (10000, 17)
        assert abserr(a[:, 0], b[:, 0]) < 0.006
        # position: z [km], lon, lat; a parcel "agrees" when all three match the 6-digit text to 1e-4 relative (+ small abs)
        pos_ok = np.all(np.abs(a[:, 1:4] - b[:, 1:4]) <= 1e-4 * np.abs(b[:, 1:4]) + 1e-4, axis=1)
        qnt_ok = np.all(np.abs(a[:, 4:] - b[:, 4:]) <= 1e-3 * np.abs(b[:, 4:]) + 1e-3 * np.max(np.abs(b[:, 4:]), axis=0), axis=1)
        report[g.name] = (float(pos_ok.mean()), float((pos_ok & qnt_ok).mean()))
    print("fraction of parcels agreeing with the shipped goldens (position, position + quantities):", report)
    out_dir = ROOT / "gpurun_out"
    out_dir.mkdir(exist_ok=True)
    (out_dir / f"trac_test_{levels}_shim.json").write_text(__import__("json").dumps(report, indent=1))
    # convection, mixing and the boundary conditions are discontinuous in the parcel position (a parcel on the edge of a
    # CAPE column or a mixing box switches sides with a last-digit difference), so after 864 steps a small fraction differs
    assert report[gold[0].name][1] == 1.0                      # t0: initial state + module_meteo + boundary conditions
    # measured on B200: positions 1.0 / 1.0 / 1.0 / 0.9999, positions + quantities 1.0 / 0.9998 / 0.9994 / 0.999
    assert all(v[0] >= 0.999 and v[1] >= 0.99 for v in report.values()), report


INTEROPER_CTL = """MET_CONVENTION = 1
MET_PRESS_LEVEL_DEF = 5
ATM_TYPE = 3
ATM_TYPE_OUT = 0
ADVECT = 2
MET_CLAMS = 1
ADVECT_VERT_COORD = 1
MET_VERT_COORD = 1
NQ = 7
QNT_NAME[0] = theta
QNT_NAME[1] = pv
QNT_NAME[2] = m
QNT_NAME[3] = zeta
QNT_NAME[4] = zeta_d
QNT_NAME[5] = ps
QNT_NAME[6] = p
METBASE = {met}/erai_vlr
DIRECTION = 1
MET_TROPO = 3
TDEC_TROP = 259200
TDEC_STRAT = 259200
DT_MOD = 180
DT_MET = 21600
T_START = 520646400
T_STOP = 520668000
CHUNKSZHINT = 163840000
ATM_DT_OUT = 21600
"""


@pytest.mark.timeout(900)
def test_trac_interoper_test_zeta_through_the_shim(tmp_path):
    """The reference's tests/interoper_test, part 2 (run.sh:22-66): diabatic transport in the zeta coordinate
    (ADVECT_VERT_COORD 1: module_advect_init + module_advect's zeta branch) on CLaMS-convention ERA-Interim data, parcels
    read from a CLaMS netCDF position file, 6 hours at 180 s with the midpoint scheme; module_meteo (pv needs the host) and
    module_decay run through the reference's CPU code between the device segments.  Compared with the shipped goldens."""
    _need()
    ref = DATA / "interoper_test.ref"
    gold = sorted(ref.glob("atm_2016_07_01_*.tab"))
    if len(gold) != 2 or not (DATA / "erai_vlr_16070106.nc").exists():
        pytest.skip("interoper_test data not shipped (oracle/build_ref.sh)")
    (tmp_path / "data").symlink_to(DATA / "clim")
    base = tmp_path / "tests" / "interoper_test"
    base.mkdir(parents=True)
    d, out = _run_trac(base, INTEROPER_CTL.format(met=DATA), ref / "pos_glo_16070100.nc", ["ATM_BASENAME", "atm", "GRID_BASENAME", "grid"],
                       atm_name="pos_glo_16070100.nc")
    assert "kernel launches" in out
    report = {}
    for g in gold:
        a, b = _tab(d / g.name), _tab(g)
        assert a.shape == b.shape
        ok = np.all(np.abs(a - b) <= 1e-4 * np.abs(b) + 1e-4 * np.max(np.abs(b), axis=0), axis=1)
        report[g.name] = float(ok.mean())
    print("fraction of parcels agreeing with the shipped goldens in every column:", report)
    out_dir = ROOT / "gpurun_out"
    out_dir.mkdir(exist_ok=True)
    (out_dir / "interoper_test_shim.json").write_text(__import__("json").dumps(report, indent=1))
    assert all(v >= 0.999 for v in report.values()), report


def test_trac_dirlist_continues_the_random_stream(tmp_path):
    """Two directories in ONE dirlist (ensemble members, src/trac.c:98-185): the reference's file-static Squares counter
    (src/mptrac.c:35) keeps counting from the first directory into the second, so the two members get different random
    numbers.  The shim carries the device counter across mptrac_free / the next context: the second member must match the
    reference's CPU run of the same dirlist (and differ from the first member)."""
    _need()
    ctl = DT_CTL.format(met=DATA).replace("T_STOP = 360547260", "T_STOP = 360547230")

    def run(root, preload):
        dirs = []
        for k in range(2):
            d = root / f"member{k}"
            d.mkdir(parents=True)
            (d / "trac.ctl").write_text(ctl)
            (d / "atm_in.tab").write_bytes((DATA / "dt_test.ref" / "atm_split.tab").read_bytes())
            dirs.append(d)
        (root / "dirlist").write_text("".join(f"{d}\n" for d in dirs))
        env = dict(os.environ, OMP_NUM_THREADS="4", LANG="C", LC_ALL="C", MPTRAC_B200_VERBOSE="1")
        if preload:
            env["LD_PRELOAD"] = str(SHIM)
        r = subprocess.run([str(TRAC), str(root / "dirlist"), "trac.ctl", "atm_in.tab", "ATM_BASENAME", "atm_pl", "MET_DT_OUT", "0"],
                           env=env, cwd=root, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
        return dirs, r.stdout

    gpu_dirs, out = run(tmp_path / "gpu", True)
    assert out.count("kernel launches") == 2
    cpu_dirs, _ = run(tmp_path / "cpu", False)
    last = "atm_pl_2011_06_05_00_00_30.tab"
    g0, g1, c1 = _tab(gpu_dirs[0] / last), _tab(gpu_dirs[1] / last), _tab(cpu_dirs[1] / last)
    assert abserr(g0[:, 2], g1[:, 2]) > 1e-4, "both members drew the same random numbers"
    for c in range(1, 4):
        assert relerr(g1[:, c], c1[:, c]) < 2e-5 or abserr(g1[:, c], c1[:, c]) < 1e-9, c


@pytest.mark.parametrize("variant", ["two_devices", "three_contexts_module_timers"])
def test_trac_dt_test_on_a_team_of_devices_is_bit_identical(tmp_path, variant):
    """MPTRAC_B200_DEVICES: the unmodified `trac` on several GPUs behind its one host thread (contiguous parcel ranges, met
    packed once and copied device to device).  The atm files must equal the single-device run byte for byte -- turbulent and
    mesoscale diffusion included, since random numbers are addressed by global parcel index.  On a one-GPU box the team's
    contexts all sit on device 0 (same code path, no NVLink)."""
    _need()
    from mptrac_b200 import load_library
    ndev = load_library().mpb_device_count()
    members = 2 if variant == "two_devices" else 3
    devices = ",".join(str(i % ndev) for i in range(members))
    env = {"MPTRAC_B200_DEVICES": devices}
    if variant != "two_devices":
        env["MPTRAC_B200_MODULE_TIMERS"] = "1"
    extra = ["ATM_BASENAME", "atm_pl"]
    one = tmp_path / "one"
    one.mkdir()
    d1, _ = _run_trac(one, DT_CTL.format(met=DATA), DATA / "dt_test.ref" / "atm_split.tab", extra)
    many = tmp_path / "many"
    many.mkdir()
    d2, out = _run_trac(many, DT_CTL.format(met=DATA), DATA / "dt_test.ref" / "atm_split.tab", extra, env_extra=env)
    assert f"on GPU {devices}" in out
    if variant != "two_devices":
        assert "TIMER_MODULE_B200_STEP" in out and "TIMER_MODULE_METEO" in out
    files = sorted(d1.glob("atm_pl_*.tab"))
    assert len(files) == 7
    for f in files:
        assert (d2 / f.name).read_bytes() == f.read_bytes(), f.name
