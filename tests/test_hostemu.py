"""The device physics source (mptrac_b200/csrc/physics.cuh), compiled for the HOST by tests/_hostemu, against the
oracle.  This only exists because the development container has no GPU: it lets the per-parcel arithmetic of the
kernels be debugged on the CPU.  It is test scaffolding -- not a product path, not a parity claim for the GPU build
(those are the `-m gpu` tests)."""
import ctypes as C
import shutil
import subprocess

import numpy as np
import pytest

from conftest import ROOT

EMU_DIR = ROOT / "tests" / "_hostemu"


class EmuCtl(C.Structure):
    _fields_ = ([(n, C.c_double) for n in "t t_start t_stop dt_met utm_ref_lat dx_pbl dx_trop dx_strat dz_pbl dz_trop dz_strat mesox mesoz pbl_trans".split()]
                + [("ctr_turb", C.c_uint64), ("ctr_meso", C.c_uint64)]
                + [(n, C.c_int) for n in "direction pbl_scheme advect phys".split()] + [("modules", C.c_uint)])


@pytest.fixture(scope="module", params=[(0, 0), (1, 0), (0, 1)], ids=["cube_f32", "cube_f64", "level_cache"])
def emu(request):
    """Every build flavour of the device physics must state the same arithmetic: both cubes of the step kernel (build flag
    MPB_CUBE_F64) and the model-level advection with and without its record cache (MPB_LEVEL_CACHE, on in the shipped library)."""
    if shutil.which("nvcc") is None:
        pytest.skip("nvcc not available")
    cube, cache = request.param
    so = EMU_DIR / (f"hostemu_{cube}.so" if not cache else f"hostemu_{cube}_cache.so")
    src = [EMU_DIR / "hostemu.cu", ROOT / "mptrac_b200" / "csrc" / "physics.cuh", ROOT / "mptrac_b200" / "csrc" / "met_tables.hpp"]
    if not so.exists() or any(s.stat().st_mtime > so.stat().st_mtime for s in src):
        subprocess.run(["nvcc", "-O2", "-std=c++17", "-Wno-deprecated-gpu-targets", f"-DMPB_CUBE_F64={cube}", f"-DMPB_LEVEL_CACHE={cache}",
                        "-Xcompiler", "-fPIC,-fopenmp,-ffp-contract=off", "-shared", str(src[0]), "-o", str(so)], check=True)
    return C.CDLL(str(so))


def emu_timestep(emu, ctl, clim, m0, m1, a, t, ctr):
    from oracle.oracle import met_struct
    e = EmuCtl()
    e.t, e.t_start, e.t_stop, e.dt_met, e.utm_ref_lat = t, ctl.t_start, ctl.t_stop, ctl.dt_met, ctl.met_utm_ref_lat
    e.dx_pbl, e.dx_trop, e.dx_strat = ctl.turb_dx_pbl, ctl.turb_dx_trop, ctl.turb_dx_strat
    e.dz_pbl, e.dz_trop, e.dz_strat = ctl.turb_dz_pbl, ctl.turb_dz_trop, ctl.turb_dz_strat
    e.mesox, e.mesoz, e.pbl_trans = ctl.turb_mesox, ctl.turb_mesoz, ctl.turb_pbl_trans
    e.direction, e.pbl_scheme, e.advect = ctl.direction, ctl.turb_pbl_scheme, ctl.advect
    n, phys = a.np, 0
    turb = ctl.diffusion and any(x > 0 for x in (ctl.turb_dx_pbl, ctl.turb_dz_pbl, ctl.turb_dx_trop, ctl.turb_dz_trop, ctl.turb_dx_strat, ctl.turb_dz_strat))
    meso = ctl.diffusion and (ctl.turb_mesox > 0 or ctl.turb_mesoz > 0)
    if turb:
        phys |= 1; e.ctr_turb = ctr; ctr += 3 * n + 1
    if meso:
        phys |= 2; e.ctr_meso = ctr; ctr += 3 * n + 1
    if ctl.qnt_rp >= 0 and ctl.qnt_rhop >= 0:
        phys |= 4
    e.phys, e.modules = phys, 1 | 2 | 64
    s0, s1 = met_struct(m0), met_struct(m1)
    rp = a.q[ctl.qnt_rp] if phys & 4 else a.time
    rhop = a.q[ctl.qnt_rhop] if phys & 4 else a.time
    vp = lambda x: C.c_void_p(x.ctypes.data)  # noqa: E731
    rc = emu.emu_step(C.byref(s0), C.byref(s1), C.byref(e), clim[0].size, clim[1].size, vp(clim[0]), vp(clim[1]), vp(clim[2]),
                      C.c_longlong(n), C.c_longlong(0), vp(a.time), vp(a.lon), vp(a.lat), vp(a.p), vp(a.dt), vp(a.uvwp), vp(rp), vp(rhop))
    assert rc == 0
    return ctr


@pytest.mark.parametrize("advect", [1, 2, 4])
@pytest.mark.parametrize("diffusion", [0, 1])
@pytest.mark.parametrize("lat_desc", [False, True])
@pytest.mark.parametrize("direction", [1, -1])
def test_device_source_on_host_is_bit_exact(emu, oracle, advect, diffusion, lat_desc, direction):
    from mptrac_b200 import Ctl, synth
    from oracle.oracle import Parcels
    m0, m1 = synth.make_met_pair(48, 25, 24, t0=0.0, dt_met=21600.0, lat_descending=lat_desc)
    n = 5000
    _, p, lon, lat = synth.make_parcels(n, t0=0.0, zmin=0.05, zmax=45.0)
    clim = synth.make_clim_tropo()
    q = np.stack([np.full(n, 2.0), np.full(n, 1500.0)])
    t_start = 0.0 if direction == 1 else 21600.0
    ctl = Ctl(nq=2, qnt_rp=0, qnt_rhop=1, advect=advect, diffusion=diffusion, direction=direction, t_start=t_start,
              t_stop=t_start + direction * 86400.0, dt_mod=300.0, dt_met=21600.0, turb_dz_trop=0.5, turb_dz_pbl=1.0,
              turb_dx_strat=20.0, turb_pbl_trans=0.3)
    tm = np.full(n, t_start)
    a, b = Parcels(tm, p, lon, lat, q), Parcels(tm, p, lon, lat, q)
    oracle.ctr = ctr = 0
    for s in range(6):
        ctr = emu_timestep(emu, ctl, clim, m0, m1, a, t_start + s * direction * 300.0, ctr)
    oracle.run("timestep", ctl, clim, m0, m1, b, t=t_start, nsteps=6)
    assert ctr == oracle.ctr
    assert np.max(np.abs(b.lat - lat)) > 1e-3
    for k in ("time", "lon", "lat", "p", "uvwp"):
        assert np.array_equal(getattr(a, k), getattr(b, k)), k


def test_surface_gaps_on_host_are_bit_exact(emu, oracle):
    """non-finite ps / pbl nodes through the device source (bilerp_guarded, time_blend_guarded, the reference's MAX / MIN
    with a NaN): same bits and same NaN positions as the oracle, which is pinned against the reference for this case"""
    from mptrac_b200 import Ctl, synth
    from oracle.oracle import Parcels
    from test_oracle_vs_reference import _with_surface_gaps
    rng = np.random.default_rng(13)
    m0, m1 = synth.make_met_pair(48, 25, 24, t0=0.0, dt_met=21600.0)
    m0, m1 = _with_surface_gaps(m0, rng, 0.10, 0.15), _with_surface_gaps(m1, rng, 0.05, 0.0)
    n = 20000
    tm, p, lon, lat = synth.make_parcels(n, t0=0.0, zmin=0.05, zmax=30.0, seed=8)
    clim = synth.make_clim_tropo()
    ctl = Ctl(advect=2, diffusion=1, turb_mesox=0.0, turb_mesoz=0.0, turb_dz_trop=0.5, turb_dz_pbl=1.0, turb_dx_pbl=30.0,
              t_start=0.0, t_stop=1e6, dt_mod=300.0, dt_met=21600.0)
    a, b = Parcels(tm, p, lon, lat), Parcels(tm, p, lon, lat)
    oracle.ctr = ctr = 0
    for s in range(3):
        ctr = emu_timestep(emu, ctl, clim, m0, m1, a, s * 300.0, ctr)
    oracle.run("timestep", ctl, clim, m0, m1, b, t=0.0, nsteps=3)
    for k in ("time", "lon", "lat", "p"):
        assert np.array_equal(getattr(a, k), getattr(b, k), equal_nan=True), k
    assert 0 < np.mean(~np.isfinite(b.p)) < 0.5


def test_device_sort_key_on_host(emu, oracle):
    from mptrac_b200 import synth
    from oracle.oracle import Parcels, met_struct
    m0, _ = synth.make_met_pair(36, 19, 20, lat_descending=True)
    tm, p, lon, lat = synth.make_parcels(20000, seed=3, zmin=0.0, zmax=70.0)
    keys = np.zeros(tm.size, np.int32)
    s0 = met_struct(m0)
    emu.emu_sort_keys(C.byref(s0), C.c_longlong(tm.size), *[C.c_void_p(x.ctypes.data) for x in (lon, lat, p, keys)])
    assert np.array_equal(keys, oracle.sort_keys(m0, Parcels(tm, p, lon, lat)))


@pytest.mark.parametrize("vert_coord", [1, 2, 3])
@pytest.mark.parametrize("advect", [1, 2, 4])
def test_model_level_advection_on_host_is_bit_exact(emu, oracle, vert_coord, advect):
    """advect_on_levels / pressure_of_zeta of the device source against the oracle's model-level module_advect and
    module_advect_init."""
    from mptrac_b200 import Ctl, synth
    from oracle.oracle import Parcels, met_struct
    m0, m1 = synth.make_met_pair(48, 25, 24, t0=0.0, dt_met=21600.0)
    m0, m1 = synth.add_model_levels(m0, npl=30), synth.add_model_levels(m1, npl=30)
    n = 4000
    tm, p, lon, lat = synth.make_parcels(n, t0=0.0, zmin=1.0, zmax=45.0, seed=5)
    clim = synth.make_clim_tropo()
    q = np.random.default_rng(2).uniform(300.0, 1500.0, (1, n))
    ctl = Ctl(nq=1, advect=advect, advect_vert_coord=vert_coord, t_start=0.0, t_stop=1e6, dt_mod=300.0, dt_met=21600.0,
              qnt_zeta=0 if vert_coord == 1 else -1, qnt_eta=0 if vert_coord == 3 else -1)
    a = Parcels(tm, p, lon, lat, q)
    oracle.run("timesteps", ctl, clim, m0, m1, a, t=300.0)
    b = a.copy()
    s0, s1 = met_struct(m0), met_struct(m1)
    vp = lambda x: C.c_void_p(x.ctypes.data)  # noqa: E731
    if vert_coord == 1:
        oracle.run("advect_init", ctl, clim, m0, m1, b, t=0.0)
        assert emu.emu_advect_init(C.byref(s0), C.byref(s1), C.c_longlong(n), vp(a.time), vp(a.lon), vp(a.lat), vp(a.p), vp(a.q[0])) == 0
        assert np.array_equal(a.p, b.p) and np.max(np.abs(a.p - p)) > 1.0
    for _ in range(3):
        oracle.run("advect", ctl, clim, m0, m1, b, t=0.0)
        zq = vp(a.q[0]) if vert_coord != 2 else None
        assert emu.emu_advect_levels(C.byref(s0), C.byref(s1), vert_coord, advect, C.c_longlong(n), vp(a.time), vp(a.lon),
                                     vp(a.lat), vp(a.p), vp(a.dt), zq) == 0
    assert np.max(np.abs(b.lat - lat)) > 1e-3
    for k in ("time", "lon", "lat", "p", "q"):
        assert np.array_equal(getattr(a, k), getattr(b, k)), k


@pytest.mark.parametrize("lat_desc", [False, True])
def test_module_meteo_all_fields_on_host_is_bit_exact(emu, oracle, lat_desc):
    """meteo_at, field2_at, field3_at and moist_at of the device source against the oracle's module_meteo: all 53
    quantities, 2-D fields with NaN gaps."""
    from mptrac_b200 import Ctl, synth
    from oracle.oracle import METEO_QNT, Parcels, met_struct
    m0, m1 = synth.make_met_pair(48, 25, 24, t0=0.0, dt_met=21600.0, lat_descending=lat_desc)
    m0, m1 = synth.add_meteo_fields(m0), synth.add_meteo_fields(m1)
    n = 5000
    tm, p, lon, lat = synth.make_parcels(n, t0=0.0, zmin=0.05, zmax=45.0, seed=4)
    tm = tm + np.random.default_rng(3).uniform(0.0, 21600.0, n)
    qm = {name: i for i, name in enumerate(METEO_QNT)}
    ctl = Ctl(nq=len(qm), t_start=0.0, t_stop=1e6, dt_mod=300.0, dt_met=21600.0, met_dt_out=300.0, qnt_meteo=qm)
    a = Parcels(tm, p, lon, lat, np.zeros((len(qm), n)))
    b = a.copy()
    oracle.run("meteo", ctl, synth.make_clim_tropo(), m0, m1, b, t=300.0)
    s0, s1 = met_struct(m0), met_struct(m1)
    qnt = (C.c_int * 64)(*([qm[nm] for nm in METEO_QNT] + [-1] * (64 - len(METEO_QNT))))
    vp = lambda x: C.c_void_p(x.ctypes.data)  # noqa: E731
    assert emu.emu_meteo(C.byref(s0), C.byref(s1), C.c_longlong(n), vp(a.time), vp(a.lon), vp(a.lat), vp(a.p), vp(a.q),
                         C.c_longlong(n), qnt) == 0
    for name, i in qm.items():
        assert np.array_equal(a.q[i], b.q[i], equal_nan=True), name
        assert np.any(b.q[i] != 0), name


@pytest.mark.parametrize("mix_pbl,cape,cin", [(1, -999.0, -999.0), (0, 100.0, -999.0), (1, 50.0, 10.0)])
def test_module_convection_on_host_is_bit_exact(emu, oracle, mix_pbl, cape, cin):
    from mptrac_b200 import Ctl, synth
    from oracle.oracle import Parcels, met_struct
    m0, m1 = synth.make_met_pair(48, 25, 24, t0=0.0, dt_met=21600.0)
    m0, m1 = synth.add_meteo_fields(m0), synth.add_meteo_fields(m1)
    n = 5000
    tm, p, lon, lat = synth.make_parcels(n, t0=0.0, zmin=0.0, zmax=14.0, seed=6)
    ctl = Ctl(advect=2, t_start=0.0, t_stop=1e6, dt_mod=300.0, dt_met=21600.0, conv_mix_pbl=mix_pbl, conv_cape=cape, conv_cin=cin,
              conv_pbl_trans=0.2 if mix_pbl else 0.0)
    clim = synth.make_clim_tropo()
    a = Parcels(tm, p, lon, lat)
    oracle.run("timesteps", ctl, clim, m0, m1, a, t=300.0)
    b = a.copy()
    oracle.ctr = 17
    oracle.run("convection", ctl, clim, m0, m1, b, t=300.0)
    r = b.rs[:n].copy()        # the uniform random numbers the module drew (cache->rs)
    s0, s1 = met_struct(m0), met_struct(m1)
    vp = lambda x: C.c_void_p(x.ctypes.data)  # noqa: E731
    assert emu.emu_convection(C.byref(s0), C.byref(s1), C.c_double(cape), C.c_double(cin), C.c_double(ctl.conv_pbl_trans),
                              mix_pbl, C.c_longlong(n), vp(a.time), vp(a.lon), vp(a.lat), vp(a.p), vp(a.dt), vp(r)) == 0
    assert np.array_equal(a.p, b.p)
    assert 0.02 < np.mean(b.p != p) < 0.98


def test_module_decay_on_host_is_bit_exact(emu, oracle):
    from mptrac_b200 import Ctl, synth
    from oracle.oracle import Parcels
    m0, m1 = synth.make_met_pair(36, 19, 20, t0=0.0, dt_met=21600.0)
    n = 4000
    tm, p, lon, lat = synth.make_parcels(n, t0=0.0, zmin=0.1, zmax=40.0, seed=9)
    clim = synth.make_clim_tropo()
    q = np.ones((1, n))
    ctl = Ctl(nq=1, advect=4, t_start=0.0, t_stop=1e6, dt_mod=300.0, dt_met=21600.0, tdec_trop=86400.0, tdec_strat=10 * 86400.0, qnt_m=0)
    a = Parcels(tm, p, lon, lat, q)
    oracle.run("timesteps", ctl, clim, m0, m1, a, t=300.0)
    b = a.copy()
    oracle.run("decay", ctl, clim, m0, m1, b, t=300.0)
    aux, tdec = np.zeros(n), np.zeros(n)
    vp = lambda x: C.c_void_p(x.ctypes.data)  # noqa: E731
    assert emu.emu_decay(0, C.c_double(0.0), C.c_double(86400.0), C.c_double(864000.0), clim[0].size, clim[1].size, vp(clim[0]),
                         vp(clim[1]), vp(clim[2]), C.c_longlong(n), vp(a.time), vp(a.lat), vp(a.p), vp(a.dt), vp(aux), vp(tdec)) == 0
    assert np.array_equal(a.q[0] * aux, b.q[0]) and np.all(aux < 1) and np.ptp(tdec) > 0


@pytest.mark.parametrize("isosurf", [1, 2, 3, 4])
def test_module_isosurf_on_host_is_bit_exact(emu, oracle, isosurf):
    from mptrac_b200 import Ctl, synth
    from oracle.oracle import Parcels, met_struct
    m0, m1 = synth.make_met_pair(48, 25, 24, t0=0.0, dt_met=21600.0)
    n = 3000
    tm, p, lon, lat = synth.make_parcels(n, t0=0.0, zmin=1.0, zmax=30.0, seed=12)
    clim = synth.make_clim_tropo()
    ctl = Ctl(advect=4, t_start=0.0, t_stop=1e6, dt_mod=300.0, dt_met=21600.0, isosurf=isosurf)
    ts, ps = np.array([0.0, 400.0, 900.0, 1300.0]), np.array([500.0, 420.0, 380.0, 300.0])
    a = Parcels(tm, p, lon, lat)
    a.balloon = (ts, ps)
    b = a.copy()
    s0, s1 = met_struct(m0), met_struct(m1)
    vp = lambda x: C.c_void_p(x.ctypes.data)  # noqa: E731
    args = lambda init: (C.byref(s0), C.byref(s1), isosurf, init, C.c_longlong(n), vp(a.time), vp(a.lon), vp(a.lat), vp(a.p),  # noqa: E731
                         vp(a.iso_var), vp(ts), vp(ps), 4)
    if isosurf != 4:
        oracle.run("isosurf_init", ctl, clim, m0, m1, b)
        assert emu.emu_isosurf(*args(1)) == 0
        assert np.array_equal(a.iso_var, b.iso_var) and np.any(a.iso_var != 0)
    for x in (a, b):                       # move the parcels: the restored pressure then differs from the present one
        x.time[:] = np.linspace(-100.0, 1500.0, n)
        x.lon[:] = (x.lon + 3.0 + 180.0) % 360.0 - 180.0
        x.p[:] = x.p * 1.02
    oracle.run("isosurf", ctl, clim, m0, m1, b)
    assert emu.emu_isosurf(*args(0)) == 0
    assert np.array_equal(a.p, b.p)
    if isosurf != 1:
        assert np.max(np.abs(a.p - p * 1.02)) > 1e-6


def test_module_diff_pbl_on_host_is_bit_exact(emu, oracle):
    from mptrac_b200 import Ctl, synth
    from oracle.oracle import Parcels, met_struct
    m0, m1 = synth.make_met_pair(48, 25, 24, t0=0.0, dt_met=21600.0)
    m0, m1 = synth.add_meteo_fields(m0, with_gaps=False), synth.add_meteo_fields(m1, with_gaps=False)
    n = 6000
    tm, p, lon, lat = synth.make_parcels(n, t0=0.0, zmin=0.0, zmax=4.0, seed=21)
    clim = synth.make_clim_tropo()
    ctl = Ctl(advect=2, diffusion=1, turb_pbl_scheme=1, t_start=0.0, t_stop=1e6, dt_mod=300.0, dt_met=21600.0)
    uvwp = np.random.default_rng(2).standard_normal((n, 3)).astype(np.float32)
    a = Parcels(tm, p, lon, lat, None, uvwp)
    oracle.run("timesteps", ctl, clim, m0, m1, a, t=300.0)
    b = a.copy()
    oracle.ctr = 23
    oracle.run("diff_pbl", ctl, clim, m0, m1, b, t=300.0)
    s0, s1 = met_struct(m0), met_struct(m1)
    vp = lambda x: C.c_void_p(x.ctypes.data)  # noqa: E731
    assert emu.emu_diff_pbl(C.byref(s0), C.byref(s1), C.c_ulonglong(23), C.c_longlong(n), vp(a.time), vp(a.lon), vp(a.lat), vp(a.p),
                            vp(a.dt), vp(a.uvwp)) == 0
    for k in ("lon", "lat", "p", "uvwp"):
        assert np.array_equal(getattr(a, k), getattr(b, k)), k
    assert 0.1 < np.mean(b.p != p) < 0.95


@pytest.mark.parametrize("layer", ["none", "dps", "dzs_pbl", "zetas"])
def test_module_bound_cond_on_host_is_bit_exact(emu, oracle, layer):
    from mptrac_b200 import Ctl, synth
    from oracle.oracle import Parcels, met_struct
    m0, m1 = synth.make_met_pair(36, 19, 20, t0=0.0, dt_met=21600.0)
    n = 4000
    tm, p, lon, lat = synth.make_parcels(n, t0=0.0, zmin=0.0, zmax=12.0, seed=31)
    clim = synth.make_clim_tropo()
    series = {"Cccl3f": (np.array([-1e4, 500.0, 2000.0, 1e5]), np.array([2e-10, 2.2e-10, 2.1e-10, 1.9e-10]))}
    oracle.set_cts(series)
    kw = dict(none={}, dps=dict(bound_dps=150.0), dzs_pbl=dict(bound_dzs=1.5, bound_pbl=1), zetas=dict(bound_zetas=320.0))[layer]
    ctl = Ctl(nq=2, advect=4, t_start=0.0, t_stop=1e6, dt_mod=300.0, dt_met=21600.0, qnt_aoa=0, qnt_cts=(-1, 1, -1, -1, -1), cts_on=2,
              bound_lat0=-60.0, bound_lat1=70.0, bound_p0=1e10, bound_p1=300.0, **kw)
    a = Parcels(tm + np.linspace(0.0, 3000.0, n), p, lon, lat, np.full((2, n), -1.0))
    a.dt[:] = 300.0
    oracle.run("bound_cond", ctl, clim, m0, m1, a, t=300.0)
    s0, s1 = met_struct(m0), met_struct(m1)
    k7 = np.array([ctl.bound_lat0, ctl.bound_lat1, ctl.bound_p0, ctl.bound_p1, ctl.bound_dps, ctl.bound_dzs, ctl.bound_zetas])
    hit = np.zeros(n, np.int32)
    vp = lambda x: C.c_void_p(x.ctypes.data)  # noqa: E731
    assert emu.emu_bound_applies(C.byref(s0), C.byref(s1), vp(k7), ctl.bound_pbl, C.c_longlong(n), vp(a.time), vp(a.lon), vp(a.lat),
                                 vp(a.p), vp(hit)) == 0
    assert np.array_equal(hit == 1, a.q[0] == a.time) and 0.05 < hit.mean() < 0.95
    emu.emu_series_at.restype = C.c_double
    emu.emu_series_at.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_double]
    ts, v = series["Cccl3f"]
    got = np.array([emu.emu_series_at(ts.ctypes.data, v.ctypes.data, ts.size, float(t)) for t in a.time[hit == 1]])
    assert np.array_equal(got, a.q[1][hit == 1])


@pytest.mark.parametrize("nens", [0, 3])
def test_module_chem_grid_on_host_is_bit_exact(emu, oracle, nens):
    from mptrac_b200 import Ctl, synth
    from oracle.oracle import Parcels, met_struct
    m0, m1 = synth.make_met_pair(36, 19, 20, t0=0.0, dt_met=21600.0)
    n = 20000
    tm, p, lon, lat = synth.make_parcels(n, t0=0.0, zmin=0.0, zmax=30.0, seed=41)
    tm = tm + np.random.default_rng(2).choice([0.0, 0.0, 0.0, 400.0], n)
    rng = np.random.default_rng(3)
    q = np.zeros((3, n))
    q[0], q[2] = rng.uniform(0.5, 2.0, n), rng.integers(0, 3, n)
    ctl = Ctl(nq=3, advect=4, t_start=0.0, t_stop=1e6, dt_mod=300.0, dt_met=21600.0, qnt_m=0, qnt_Cx=1, qnt_ens=2, nens=nens, molmass=64.07,
              chemgrid_nx=36, chemgrid_ny=18, chemgrid_nz=15, chemgrid_z0=0.0, chemgrid_z1=30.0, chemgrid=1)
    b = Parcels(tm, p, lon, lat, q)
    oracle.run("chem_grid", ctl, synth.make_clim_tropo(), m0, m1, b, t=0.0)
    dlon, dlat, dz = 360.0 / 36, 180.0 / 18, 30.0 / 15
    k13 = np.array([-180.0, 180.0, -90.0, 90.0, 0.0, 30.0, dlon, dlat, dz, -150.0, 150.0, 0.0, 64.07])
    cx = np.zeros(n)
    s0, s1 = met_struct(m0), met_struct(m1)
    vp = lambda x: C.c_void_p(x.ctypes.data)  # noqa: E731
    m, ens = np.ascontiguousarray(q[0]), np.ascontiguousarray(q[2])
    assert emu.emu_chem_grid(C.byref(s0), C.byref(s1), vp(k13), 36, 18, 15, nens, C.c_longlong(n), vp(tm), vp(lon), vp(lat), vp(p),
                             vp(m), vp(ens), vp(cx)) == 0
    assert np.array_equal(cx, b.q[1]) and 0.5 < np.mean(cx > 0) < 0.9


def test_model_level_advection_fuzz_on_host(emu, oracle):
    """many small cases (level counts, grids, latitude order, coordinates, integrators) with parcels exactly on level
    pressures and beyond both ends of the columns: the verified-hint level search must give the bisection's answer"""
    import itertools
    from mptrac_b200 import Ctl, synth
    from oracle.oracle import Parcels, met_struct
    vp = lambda x: C.c_void_p(x.ctypes.data)  # noqa: E731
    clim = synth.make_clim_tropo()
    for seed, npl, vc, adv, desc in itertools.product(range(2), (8, 61), (1, 2, 3), (2, 4), (False, True)):
        m0, m1 = synth.make_met_pair(24, 13, 10, t0=0.0, dt_met=21600.0, lat_descending=desc)
        m0, m1 = synth.add_model_levels(m0, npl=npl, seed=seed), synth.add_model_levels(m1, npl=npl, seed=seed + 100)
        n = 800
        tm, p, lon, lat = synth.make_parcels(n, t0=0.0, zmin=0.0, zmax=62.0, seed=seed)
        p[:50] = m0.pl[3, 4, np.random.default_rng(seed).integers(0, npl, 50)]
        p[50:60], p[60:70] = 1200.0, 0.05
        q = np.random.default_rng(seed).uniform(200.0, 2200.0, (1, n))
        ctl = Ctl(nq=1, advect=adv, advect_vert_coord=vc, t_start=0.0, t_stop=1e6, dt_mod=600.0, dt_met=21600.0,
                  qnt_zeta=0 if vc == 1 else -1, qnt_eta=0 if vc == 3 else -1)
        a = Parcels(tm, p, lon, lat, q)
        oracle.run("timesteps", ctl, clim, m0, m1, a, t=600.0)
        b = a.copy()
        s0, s1 = met_struct(m0), met_struct(m1)
        for _ in range(4):
            oracle.run("advect", ctl, clim, m0, m1, b, t=0.0)
            zq = vp(a.q[0]) if vc != 2 else None
            assert emu.emu_advect_levels(C.byref(s0), C.byref(s1), vc, adv, C.c_longlong(n), vp(a.time), vp(a.lon), vp(a.lat), vp(a.p),
                                         vp(a.dt), zq) == 0
            a.time[:] = b.time[:] = 0.0       # stay inside the met interval
        for k in ("lon", "lat", "p", "q"):
            assert np.array_equal(getattr(a, k), getattr(b, k), equal_nan=True), (seed, npl, vc, adv, desc, k)
