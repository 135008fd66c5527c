"""GPU parity tests proper: the CUDA engine, called through the C ABI, against the CPU oracle on the same seeded
inputs, against the committed golden fixtures generated from the unmodified reference, and -- at BASELINE.json's
full sizes -- through size-independent properties.

Tolerances (fp64 state, fp32 met; stated per test): the device build contracts a*b+c into FMAs and uses CUDA's
libm (cos/log/exp/pow/sinf/cosf are not bit-identical to glibc's), so agreement is ~1e-13 relative per step instead
of bit-exact; index / integer work (sort keys, box indices, counts, RNG integers) must be exact.
"""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, abserr, clim_from_npz, met_from_npz, relerr

pytestmark = pytest.mark.gpu

# Tolerances.  lon/lat in degrees (absolute), pressure relative.
#  - diffusion off: only FMA contraction and CUDA's cos() separate the device from the x86-64 oracle
#    (measured on B200: ~3e-14 deg, ~1e-15 relative after 5 RK4 steps)
#  - diffusion on: Box-Muller takes sinf/cosf of a float angle; CUDA's are 2-ulp, glibc's ~0.5-ulp, so single normals
#    differ by ~1e-7 relative, i.e. ~1e-9 deg of displacement per step (north_star asks for moments only here)
TOL_POS_DEG = 1e-11
TOL_P_REL = 1e-12
TOL_POS_DEG_DIFF = 1e-7
TOL_P_REL_DIFF = 1e-7

REPORT = {}


def _report(name, **kw):
    REPORT[name] = {k: float(v) for k, v in kw.items()}
    out = ROOT / "gpurun_out"
    try:
        out.mkdir(exist_ok=True)
        (out / "parity_report.json").write_text(json.dumps(REPORT, indent=1, sort_keys=True))
    except OSError:
        pass


def _engine(n, nq=0, strict=False):
    from mptrac_b200 import Engine
    return Engine(n, nq=nq, device=0, strict=strict)


def _setup(eng, ctl, clim, m0, m1, tm, p, lon, lat, q=None):
    eng.set_ctl(ctl)
    eng.set_clim_tropo(*clim)
    eng.set_met(0, m0)
    eng.set_met(1, m1)
    eng.set_atm(tm, p, lon, lat, q)


def _case(n=6000, grid=(48, 25, 24), lat_desc=False, seed=5, zmax=45.0):
    from mptrac_b200 import synth
    m0, m1 = synth.make_met_pair(*grid, t0=0.0, dt_met=21600.0, lat_descending=lat_desc)
    tm, p, lon, lat = synth.make_parcels(n, t0=0.0, zmin=0.05, zmax=zmax, seed=seed)
    return m0, m1, tm, p, lon, lat, synth.make_clim_tropo()


def _compare(tag, out, ref, tol_pos=TOL_POS_DEG, tol_p=TOL_P_REL):
    # longitude differences are measured along the parallel (x cos(lat)): near the poles a metre is many degrees
    dlon = (np.asarray(out["lon"]) - ref.lon + 180.0) % 360.0 - 180.0
    e_lon = float(np.max(np.abs(dlon) * np.maximum(np.cos(np.deg2rad(ref.lat)), 1e-6))) if dlon.size else 0.0
    e_lat, e_p, e_t = abserr(out["lat"], ref.lat), relerr(out["p"], ref.p), abserr(out["time"], ref.time)
    _report(tag, lon_abs=e_lon, lat_abs=e_lat, p_rel=e_p, time_abs=e_t)
    assert e_t == 0.0, f"{tag}: time differs"
    assert e_lon < tol_pos and e_lat < tol_pos, f"{tag}: lon {e_lon:.3e} lat {e_lat:.3e}"
    assert e_p < tol_p, f"{tag}: p rel {e_p:.3e}"


@pytest.mark.parametrize("advect", [1, 2, 4])
@pytest.mark.parametrize("diffusion", [0, 1])
@pytest.mark.parametrize("lat_desc", [False, True])
@pytest.mark.parametrize("direction", [1, -1])
def test_timestep_vs_oracle(oracle, advect, diffusion, lat_desc, direction):
    from mptrac_b200 import Ctl
    from oracle.oracle import Parcels
    m0, m1, tm, p, lon, lat, clim = _case(lat_desc=lat_desc)
    n = tm.size
    t_start = 0.0 if direction == 1 else 21600.0
    tm = np.full(n, t_start)
    q = np.stack([np.full(n, 2.0), np.full(n, 1500.0)])
    ctl = Ctl(nq=2, qnt_rp=0, qnt_rhop=1, advect=advect, diffusion=diffusion, direction=direction, t_start=t_start,
              t_stop=t_start + direction * 86400.0, dt_mod=300.0, dt_met=21600.0, turb_dz_trop=0.5, turb_dz_pbl=1.0,
              turb_dx_strat=20.0, turb_pbl_trans=0.3)
    nsteps = 6
    with _engine(n, 2) as eng:
        _setup(eng, ctl, clim, m0, m1, tm, p, lon, lat, q)
        for s in range(nsteps):
            eng.run_timestep(t_start + s * direction * 300.0)
        out = eng.get_atm()
        uv = eng.get_uvwp()
        ctr = eng.rng_ctr
    ref = Parcels(tm, p, lon, lat, q)
    oracle.ctr = 0
    oracle.run("timestep", ctl, clim, m0, m1, ref, t=t_start, nsteps=nsteps)
    assert ctr == oracle.ctr
    assert abserr(ref.lat, lat) > 1e-3, "nothing moved"
    _compare(f"timestep[advect={advect},diff={diffusion},latdesc={int(lat_desc)},dir={direction}]", out, ref,
             TOL_POS_DEG_DIFF if diffusion else TOL_POS_DEG, TOL_P_REL_DIFF if diffusion else TOL_P_REL)
    if diffusion:
        assert relerr(uv + 1e-30, ref.uvwp + 1e-30) < 1e-4 or abserr(uv, ref.uvwp) < 1e-5


def test_strict_build_is_tighter(oracle):
    """-fmad=false flavour: remaining differences come from libm only."""
    from mptrac_b200 import Ctl
    from oracle.oracle import Parcels
    m0, m1, tm, p, lon, lat, clim = _case()
    n = tm.size
    ctl = Ctl(advect=4, t_start=0.0, t_stop=1e6, dt_mod=300.0, dt_met=21600.0)
    with _engine(n, 0, strict=True) as eng:
        _setup(eng, ctl, clim, m0, m1, tm, p, lon, lat)
        for s in range(6):
            eng.run_timestep(s * 300.0)
        out = eng.get_atm()
    ref = Parcels(tm, p, lon, lat)
    oracle.run("timestep", ctl, clim, m0, m1, ref, t=0.0, nsteps=6)
    _compare("strict_rk4", out, ref, tol_pos=1e-12, tol_p=1e-13)
    _report("strict_rk4_bitexact_fraction", lon=np.mean(out["lon"] == ref.lon), lat=np.mean(out["lat"] == ref.lat), p=np.mean(out["p"] == ref.p))


@pytest.mark.parametrize("module", ["position", "advect", "diff_turb", "diff_meso", "sedi"])
def test_single_modules_vs_oracle(oracle, module):
    """Each exported module_* entry point on its own (cache->dt supplied by module_timesteps)."""
    from mptrac_b200 import Ctl
    from oracle.oracle import Parcels
    m0, m1, tm, p, lon, lat, clim = _case(seed=9)
    n = tm.size
    rng = np.random.default_rng(1)
    # push some parcels out of range so that module_position has something to do
    lon = lon + rng.choice([0.0, 360.0, -360.0], n)
    lat = np.where(rng.uniform(size=n) < 0.05, lat + 100.0, lat)
    p = np.where(rng.uniform(size=n) < 0.05, p * 1.5, p)
    q = np.stack([rng.uniform(0.1, 10.0, n), rng.uniform(500.0, 2500.0, n)])
    uvwp = rng.standard_normal((n, 3)).astype(np.float32)
    ctl = Ctl(nq=2, qnt_rp=0, qnt_rhop=1, advect=4, diffusion=1, t_start=0.0, t_stop=1e6, dt_mod=300.0, dt_met=21600.0,
              turb_dz_trop=0.5, turb_dz_pbl=1.0, turb_dx_strat=20.0, turb_pbl_trans=0.3)
    with _engine(n, 2) as eng:
        _setup(eng, ctl, clim, m0, m1, tm, p, lon, lat, q)
        eng.set_uvwp(uvwp)
        eng.rng_ctr = 777
        eng.module_timesteps(300.0)
        dt = eng.get_dt()
        getattr(eng, f"module_{module}")()
        out = eng.get_atm()
        uv = eng.get_uvwp()
        ctr = eng.rng_ctr
    ref = Parcels(tm, p, lon, lat, q, uvwp)
    oracle.ctr = 777
    oracle.run("timesteps", ctl, clim, m0, m1, ref, t=300.0)
    assert np.array_equal(dt, ref.dt)
    oracle.run(module, ctl, clim, m0, m1, ref, t=300.0)
    assert ctr == oracle.ctr
    loose = module in ("diff_turb", "diff_meso")
    _compare(f"module_{module}", out, ref, TOL_POS_DEG_DIFF if loose else TOL_POS_DEG, TOL_P_REL_DIFF if loose else TOL_P_REL)
    assert abserr(uv, ref.uvwp) < 1e-5


@pytest.mark.parametrize("lat_desc", [False, True])
def test_module_meteo_vs_oracle(oracle, lat_desc):
    """module_meteo on the device (all 14 quantities that derive from the resident fields) at scattered parcel times;
    1e-13 relative: the device contracts FMAs and its pow / exp / sin differ from glibc's in the last bits"""
    from mptrac_b200 import Ctl
    from mptrac_b200.host import METEO_QNT
    from oracle.oracle import Parcels
    m0, m1, tm, p, lon, lat, clim = _case(lat_desc=lat_desc)
    n = tm.size
    tm = tm + np.random.default_rng(1).uniform(0, 20000, n)
    qm = {name: i for i, name in enumerate(METEO_QNT[:14])}    # ps ... zeta_d: the quantities of the resident fields
    ctl = Ctl(nq=len(qm), advect=4, t_start=0.0, t_stop=1e6, dt_mod=300.0, dt_met=21600.0, met_dt_out=0.1, qnt_meteo=qm)
    q0 = np.zeros((len(qm), n))
    with _engine(n, len(qm)) as eng:
        _setup(eng, ctl, clim, m0, m1, tm, p, lon, lat, q0)
        eng.module_meteo()
        out = eng.get_atm()
    ref = Parcels(tm, p, lon, lat, q0)
    oracle.run("meteo", ctl, clim, m0, m1, ref)
    for name, i in qm.items():
        signed = name in ("u", "v", "w", "vz")      # these pass through zero: error relative to the field's scale
        e = abserr(out["q"][i], ref.q[i]) / np.max(np.abs(ref.q[i])) if signed else relerr(out["q"][i], ref.q[i])
        _report(f"module_meteo[{name},latdesc={int(lat_desc)}]", rel=e)
        assert e < 1e-12, name
    assert np.array_equal(out["lon"], lon) and np.array_equal(out["p"], p)


@pytest.mark.parametrize("vert_coord", [1, 2, 3])
@pytest.mark.parametrize("advect", [1, 2, 4])
@pytest.mark.parametrize("diffusion", [0, 1])
def test_model_level_advection_vs_oracle(oracle, vert_coord, advect, diffusion):
    """ADVECT_VERT_COORD 1 / 2 / 3 (zeta, omega on model levels, eta): module_advect_init at t_start, the model-level
    advection launch between the timesteps/position segment and the diffusion/sedimentation/position segment, the zeta /
    eta quantity written back; met levels uploaded in swapped order and exchanged with mpb_swap_met."""
    from mptrac_b200 import Ctl, synth
    from oracle.oracle import Parcels
    m0, m1, tm, p, lon, lat, clim = _case(n=6000, grid=(48, 25, 24))
    m0, m1 = synth.add_model_levels(m0, npl=30), synth.add_model_levels(m1, npl=30)
    n = tm.size
    rng = np.random.default_rng(2)
    q = np.stack([rng.uniform(0.1, 10, n), rng.uniform(500, 2500, n), rng.uniform(300.0, 1500.0, n)])
    ctl = Ctl(nq=3, qnt_rp=0, qnt_rhop=1, advect=advect, advect_vert_coord=vert_coord, diffusion=diffusion, t_start=0.0,
              t_stop=1e6, dt_mod=300.0, dt_met=21600.0, turb_dz_trop=0.5, turb_dx_strat=20.0, turb_mesox=0.16, turb_mesoz=0.16,
              qnt_zeta=2 if vert_coord == 1 else -1, qnt_eta=2 if vert_coord == 3 else -1, sort_dt=600.0)
    with _engine(n, 3) as eng:
        eng.set_ctl(ctl)
        eng.set_clim_tropo(*clim)
        eng.set_met(0, m1)
        eng.set_met(1, m0)
        eng.swap_met()
        eng.set_atm(tm, p, lon, lat, q)
        for s in range(5):
            eng.run_timestep(300.0 * s)
        out = eng.get_atm()
    ref = Parcels(tm, p, lon, lat, q)
    oracle.ctr = 0
    oracle.run("timestep", ctl, clim, m0, m1, ref, t=0.0, nsteps=5)
    assert abserr(ref.lat, lat) > 1e-3
    # the cell sort permutes both sides alike (stable, same keys): compare in place
    tol = (TOL_POS_DEG_DIFF, TOL_P_REL_DIFF) if diffusion else (TOL_POS_DEG, TOL_P_REL)
    _compare(f"levels_vc{vert_coord}_adv{advect}_diff{diffusion}", out, ref, *tol)
    if vert_coord != 2:
        assert relerr(out["q"][2], ref.q[2]) < tol[1]
    assert np.array_equal(out["q"][:2], ref.q[:2])


def test_model_level_advection_needs_the_fields():
    from mptrac_b200 import Ctl
    m0, m1, tm, p, lon, lat, clim = _case(n=100)
    ctl = Ctl(advect=4, advect_vert_coord=2, t_start=0.0, t_stop=1e6, dt_mod=300.0, dt_met=21600.0)
    with _engine(100) as eng:
        _setup(eng, ctl, clim, m0, m1, tm, p, lon, lat)
        with pytest.raises(RuntimeError, match="model-level"):
            eng.run_timestep(300.0)


def test_rng_stream(oracle):
    """Squares counters are integers: uniforms must be bit-exact; Box-Muller normals agree to float-trig accuracy."""
    with _engine(10) as eng:
        eng.rng_ctr = 12345
        u = eng.module_rng(3001, 0)
        g = eng.module_rng(3001, 1)
        end = eng.rng_ctr
    oracle.ctr = 12345
    u0 = oracle.module_rng(3001, 0)
    g0 = oracle.module_rng(3001, 1)
    assert end == oracle.ctr
    assert np.array_equal(u, u0)
    assert abserr(g[:3001], g0[:3001]) < 5e-6
    assert abs(np.mean(g[:3000])) < 0.1 and abs(np.std(g[:3000]) - 1) < 0.1


def _golden(name):
    f = GOLDEN / name
    if not f.exists():
        pytest.skip(f"{f} missing")
    return np.load(f)


def test_golden_dt_test():
    """tests/dt_test of the reference (ERA-Interim, midpoint + turbulent + mesoscale diffusion, 7 outputs)."""
    from mptrac_b200 import Ctl
    z = _golden("dt_test.npz")
    m0, m1, clim = met_from_npz(z, "m0"), met_from_npz(z, "m1"), clim_from_npz(z)
    t0, n_total, k = float(z["t_start"]), int(z["np_total"]), z["time"].size
    # all 8 columns of the reference's goldens: positions, and the quantities t, u, v, w that module_meteo writes after
    # every step (the test's control file leaves MET_DT_OUT at its default 0.1)
    ctl = Ctl(nq=4, advect=2, diffusion=1, dt_mod=10.0, dt_met=86400.0, t_start=t0, t_stop=t0 + 60.0, met_dt_out=0.1,
              qnt_meteo=dict(t=0, u=1, v=2, w=3))
    with _engine(k, 4) as eng:
        _setup(eng, ctl, clim, m0, m1, z["time"], z["p"], z["lon"], z["lat"], np.zeros((4, k)))
        eng.set_shard(0, n_total)   # the fixture holds the first k of n_total parcels: counters advance as for all
        for s in range(7):
            eng.run_timestep(t0 + 10.0 * s)
            out = eng.get_atm()
            rb, txt = z["ref_binary"][s], z["ref_shipped_text"][s]
            e = dict(lon_abs=abserr(out["lon"], rb[2]), lat_abs=abserr(out["lat"], rb[3]), p_rel=relerr(out["p"], rb[1]))
            _report(f"golden_dt_test_step{s}", **e)
            assert e["lon_abs"] < 1e-7 and e["lat_abs"] < 1e-7 and e["p_rel"] < 1e-7 and np.array_equal(out["time"], rb[0])
            # the reference's own shipped text goldens (%g: 6 significant digits)
            zkm = 7.0 * np.log(1013.25 / out["p"])
            assert relerr(zkm, txt[:, 1]) < 1e-5 and relerr(out["lon"], txt[:, 2]) < 1e-5 and relerr(out["lat"], txt[:, 3]) < 1e-5
            rq = z["ref_binary_q"][s]
            eq = float(np.max(np.abs(out["q"] - rq) / np.max(np.abs(rq), axis=1, keepdims=True)))
            _report(f"golden_dt_test_step{s}_quantities", q_rel_to_scale=eq)
            assert eq < 1e-7      # the positions they are evaluated at differ by ~1e-10 deg (diffusion: sinf / cosf)
            tq = z["ref_shipped_q"][s]
            assert np.all(np.abs(out["q"].T - tq) <= 1e-5 * np.maximum(np.abs(tq), 1e-3 * np.max(np.abs(tq), axis=0)))


def test_c1_trac_test_parcels_midpoint_vs_reference(reference):
    """BASELINE configs[0]: the 10 000 parcels of the reference's tests/trac_test (data.ref/atm_init.tab) on its
    ERA-Interim test data, ADVECT 2 (midpoint), diffusion and every other module off, one day at DT_MOD 300 -- against
    the UNMODIFIED reference run right here through oracle/_ref (met read + pre-processed by the reference itself).
    north_star's bound is 1e-6 relative on lon / lat / p; measured ~1e-13."""
    from mptrac_b200 import Ctl, Met
    from oracle.oracle import Parcels
    data = ROOT / "oracle" / "_ref" / "data"
    files = [data / "ei_2011_06_05_00.nc", data / "ei_2011_06_06_00.nc", data / "trac_test.ref" / "atm_init.tab"]
    if not all(f.exists() for f in files):
        pytest.skip("reference test data not shipped (oracle/build_ref.sh)")
    reference.read_ctl([], "")
    m0, m1 = reference.read_met(files[0], 0, Met), reference.read_met(files[1], 1, Met)
    tm, p, lon, lat = reference.read_atm(files[2])
    n = tm.size
    assert n == 10000 and m1.time - m0.time == 86400.0
    t0 = m0.time
    ctl = Ctl(advect=2, diffusion=0, t_start=t0, t_stop=t0 + 86400.0, dt_mod=300.0, dt_met=86400.0)
    nsteps = 289          # t0, t0 + 300, ..., t0 + 86400 (the first call has dt = 0)
    ref = Parcels(tm, p, lon, lat)
    reference.ctr = 0
    reference.run("timestep", ctl, ref, t=t0, nsteps=nsteps)
    with _engine(n) as eng:
        _setup(eng, ctl, reference.clim_tropo(), m0, m1, tm, p, lon, lat)
        for s in range(nsteps):
            eng.run_timestep(t0 + 300.0 * s)
        out = eng.get_atm()
    assert abserr(ref.lat, lat) > 1.0, "nothing moved"
    # parcels that brushed a pole (DX2DEG switches the zonal displacement off within 0.001 deg of it, a discontinuity) are
    # compared on the sphere like all others: the metric below weighs dlon with cos(lat)
    _compare("c1_trac_test_midpoint_1day", out, ref, tol_pos=1e-6, tol_p=1e-6)
    dlon = (out["lon"] - ref.lon + 180.0) % 360.0 - 180.0
    med = float(np.median(np.hypot(dlon * np.cos(np.deg2rad(ref.lat)), out["lat"] - ref.lat)))
    _report("c1_trac_test_midpoint_1day_median", pos_deg=med)
    assert med < 1e-12


def test_golden_coord_test():
    """tests/coord_test of the reference: Cartesian (UTM) met, three hourly levels (met swap), 13 outputs."""
    from mptrac_b200 import Ctl
    z = _golden("coord_test.npz")
    mets = [met_from_npz(z, f"m{i}") for i in range(3)]
    clim = clim_from_npz(z)
    t0, n = float(z["t_start"]), z["time"].size
    ctl = Ctl(nq=0, advect=2, diffusion=1, dt_mod=600.0, dt_met=3600.0, t_start=t0, t_stop=t0 + 7200.0, met_coord_type=1,
              met_utm_ref_lat=48.1507476)
    with _engine(n) as eng:
        _setup(eng, ctl, clim, mets[0], mets[1], z["time"], z["p"], z["lon"], z["lat"])
        for s in range(13):
            t = t0 + 600.0 * s
            if s == 7:   # first step with t > met1->time: the swap of mptrac_get_met, then one new level
                eng.swap_met()
                eng.set_met(1, mets[2])
            eng.run_timestep(t)
            out = eng.get_atm()
            rb = z["ref_binary"][s]
            e = dict(x_abs_m=abserr(out["lon"], rb[2]), y_abs_m=abserr(out["lat"], rb[3]), p_rel=relerr(out["p"], rb[1]))
            _report(f"golden_coord_test_step{s}", **e)
            assert e["x_abs_m"] < 1e-3 and e["y_abs_m"] < 1e-3 and e["p_rel"] < 1e-7
            txt = z["ref_shipped_text"][s]
            assert abserr(out["lon"], txt[:, 2]) < 0.02 and abserr(out["lat"], txt[:, 3]) < 0.02   # metres; text has 0.01 m resolution


def test_golden_synth_full():
    """RK4 + diff_turb + diff_meso + sedi + mixing on a north->south grid; reference binary results."""
    from mptrac_b200 import Ctl, synth
    z = _golden("synth_full.npz")
    m0, m1 = synth.make_met_pair(24, 13, 20, t0=0.0, dt_met=21600.0, lat_descending=True)
    clim = clim_from_npz(z)
    n = z["time"].size
    ctl = Ctl(nq=3, qnt_rp=0, qnt_rhop=1, advect=4, diffusion=1, t_start=0.0, t_stop=86400.0, dt_mod=600.0, dt_met=21600.0,
              turb_dz_trop=0.5, turb_dz_pbl=1.0, turb_dx_strat=20.0, turb_pbl_trans=0.2, mixing_trop=0.3, mixing_strat=0.1,
              mixing_dt=1200.0, mix_qnt=[2], mixing_nx=18, mixing_ny=9, mixing_nz=12)
    with _engine(n, 3) as eng:
        _setup(eng, ctl, clim, m0, m1, z["time"], z["p"], z["lon"], z["lat"], z["q"])
        for s in range(8):
            eng.run_timestep(600.0 * s)
            out = eng.get_atm()
            rb = z["ref_binary"][s]
            e = dict(lon_abs=abserr(out["lon"], rb[2]), lat_abs=abserr(out["lat"], rb[3]), p_rel=relerr(out["p"], rb[1]),
                     q_rel=relerr(out["q"][2], z["ref_q"][s][2]))
            _report(f"golden_synth_full_step{s}", **e)
            assert e["lon_abs"] < 1e-6 and e["lat_abs"] < 1e-6 and e["p_rel"] < 1e-6 and e["q_rel"] < 1e-6
        assert eng.rng_ctr == int(z["ref_ctr"])


def test_sort_matches_stable_oracle(oracle):
    from mptrac_b200 import Ctl
    from oracle.oracle import Parcels
    m0, m1, tm, p, lon, lat, clim = _case(n=50000, seed=2)
    n = tm.size
    q = np.arange(n, dtype=np.float64)[None, :]          # original index rides along as a quantity
    ctl = Ctl(nq=1, advect=4, t_start=0.0, t_stop=1e6, dt_mod=300.0, dt_met=21600.0)
    with _engine(n, 1) as eng:
        _setup(eng, ctl, clim, m0, m1, tm, p, lon, lat, q)
        eng.module_sort()
        out = eng.get_atm()
    ref = Parcels(tm, p, lon, lat, q)
    keys_before = oracle.sort_keys(m0, ref)
    oracle.run("sort", ctl, clim, m0, m1, ref)
    # index work: bit-exact, including the order inside a cell (CUB radix sort is stable)
    assert np.array_equal(out["q"][0], ref.q[0])
    for k in ("lon", "lat", "p", "time"):
        assert np.array_equal(out[k], getattr(ref, k))
    keys_after = oracle.sort_keys(m0, Parcels(out["time"], out["p"], out["lon"], out["lat"]))
    assert np.all(np.diff(keys_after) >= 0)
    assert np.array_equal(np.sort(keys_before), keys_after)


def test_sort_edge_cases(oracle):
    """empty, single parcel, all parcels in one cell"""
    from mptrac_b200 import Ctl
    m0, m1, tm, p, lon, lat, clim = _case(n=100)
    ctl = Ctl(nq=0, advect=2, t_start=0.0, t_stop=1e6, dt_mod=300.0, dt_met=21600.0, sort_dt=300.0)
    for n in (0, 1, 100):
        with _engine(100) as eng:
            same = n == 100
            lo, la, pp = (np.full(n, 10.0), np.full(n, 20.0), np.full(n, 500.0)) if same else (lon[:n], lat[:n], p[:n])
            _setup(eng, ctl, clim, m0, m1, tm[:n], pp, lo, la)
            eng.run_timestep(0.0)
            eng.run_timestep(300.0)
            out = eng.get_atm()
            assert out["lon"].size == n
            if same:
                assert np.all(out["lon"] == out["lon"][0]) and out["lon"][0] != 10.0


def test_mixing_and_grid_vs_oracle(oracle):
    from mptrac_b200 import Ctl
    from oracle.oracle import Parcels
    m0, m1, tm, p, lon, lat, clim = _case(n=40000, seed=4)
    n = tm.size
    rng = np.random.default_rng(0)
    q = np.stack([rng.uniform(0, 1, n), rng.uniform(10, 20, n), rng.integers(0, 3, n).astype(float)])
    ctl = Ctl(nq=3, advect=0, t_start=0.0, t_stop=1e6, dt_mod=300.0, dt_met=21600.0, mixing_trop=0.4, mixing_strat=0.05,
              mixing_dt=300.0, mix_qnt=[0, 1], mixing_nx=36, mixing_ny=18, mixing_nz=15, nens=3, qnt_ens=2)
    with _engine(n, 3) as eng:
        _setup(eng, ctl, clim, m0, m1, tm, p, lon, lat, q)
        eng.module_mixing(0.0)
        out = eng.get_atm()
        eng.grid_accumulate(36, 18, 4, -180, 180, -90, 90, 0, 40, -150.0, 150.0)
        cnt, s, sq = eng.grid_fetch()
    ref = Parcels(tm, p, lon, lat, q)
    oracle.run("mixing", ctl, clim, m0, m1, ref, t=0.0)
    assert abserr(ref.q[0], q[0]) > 1e-3
    # sums run in a different order (atomics): ~1e-14 relative
    e = dict(q0_rel=relerr(out["q"][0], ref.q[0]), q1_rel=relerr(out["q"][1], ref.q[1]))
    _report("mixing", **e)
    assert e["q0_rel"] < 1e-11 and e["q1_rel"] < 1e-11 and np.array_equal(out["q"][2], ref.q[2])
    c0, s0, sq0 = oracle.grid_bin(Parcels(out["time"], out["p"], out["lon"], out["lat"], out["q"]), 36, 18, 4, -180, 180, -90, 90, 0, 40, -150.0, 150.0)
    assert np.array_equal(cnt, c0) and cnt.sum() > 0.5 * n
    assert relerr(s + 1e-300, s0 + 1e-300) < 1e-11 and relerr(sq + 1e-300, sq0 + 1e-300) < 1e-11


def test_binning_at_the_cell_faces(oracle):
    """the production library decides the level of a parcel from a single-precision altitude wherever that is decisive and
    from the reference's double-precision sequence elsewhere (physics.cuh box_level); longitudes and latitudes go through
    quotients by grid constants (div_for_index).  Parcels ON every face of the output grid and 1e-15 .. 1e-2 beside it must
    fall into the boxes the reference puts them in -- counts exact, in sorted-looking and in shuffled order.  Vertically
    the claim starts 1e-11 km from a face: closer than that the device's `log` (1 ulp) and glibc's may differ in the last bit
    of Z(p), which is a property of the two math libraries, not of the kernel; those parcels must only be counted somewhere."""
    from oracle.oracle import Parcels
    from mptrac_b200 import Ctl
    m0, m1, tm, p, lon, lat, clim = _case(n=1000, seed=9)
    nx, ny, nz, z0, z1 = 36, 18, 40, -2.0, 38.0
    mag = np.concatenate([[0.0], 10.0 ** np.arange(-15.0, -1.5, 0.5)])
    eps = np.concatenate([-mag[::-1], mag])
    faces_z = z0 + (z1 - z0) / nz * np.arange(nz + 1)
    far = eps[np.abs(eps) >= 1e-11]
    near = eps[np.abs(eps) < 1e-11]
    pz_far = (1013.25 * np.exp(-(faces_z[:, None] + far[None, :]) / 7.0)).ravel()
    pz_near = (1013.25 * np.exp(-(faces_z[1:-1, None] + near[None, :]) / 7.0)).ravel()      # (inner faces: inside either way)
    xf = ((-180.0 + 360.0 / nx * np.arange(nx + 1))[:, None] + 10.0 * eps[None, :]).ravel()
    yf = ((-90.0 + 180.0 / ny * np.arange(ny + 1))[:, None] + 10.0 * eps[None, :]).ravel()
    rng = np.random.default_rng(5)
    m = 40000
    ctl = Ctl(nq=1, advect=0, t_start=0.0, t_stop=1e6, dt_mod=300.0, dt_met=21600.0)

    def binned(P, X, Y, order):
        n = P.size
        q = rng.uniform(0.0, 1.0, (1, n))
        t_, p_, x_, y_, q_ = np.zeros(n), P[order], X[order], Y[order], np.ascontiguousarray(q[:, order])
        with _engine(n, 1) as eng:
            _setup(eng, ctl, clim, m0, m1, t_, p_, x_, y_, q_)
            eng.grid_accumulate(nx, ny, nz, -180, 180, -90, 90, z0, z1, -150.0, 150.0)
            got = eng.grid_fetch()
        want = oracle.grid_bin(Parcels(t_, p_, x_, y_, q_), nx, ny, nz, -180, 180, -90, 90, z0, z1, -150.0, 150.0)
        return got, want

    sets = {
        "levels": (rng.choice(pz_far, m), rng.uniform(-180, 180, m), rng.uniform(-90, 90, m)),
        "columns": (rng.uniform(1.0, 1100.0, m), rng.choice(xf, m), rng.choice(yf, m)),
        "both": (rng.choice(pz_far, m), rng.choice(xf, m), rng.choice(yf, m)),
    }
    for name, (P, X, Y) in sets.items():
        for order in (np.lexsort((P, Y, X)), rng.permutation(m)):
            (cnt, s, sq), (c0, s0, sq0) = binned(P, X, Y, order)
            bad = np.flatnonzero(cnt != c0)
            assert bad.size == 0, f"{name}: {bad.size} boxes differ, first {bad[:5]}: {cnt[bad[:5]]} against {c0[bad[:5]]}"
            assert 0.3 * m < cnt.sum() <= m
            assert relerr(s + 1e-300, s0 + 1e-300) < 1e-11 and relerr(sq + 1e-300, sq0 + 1e-300) < 1e-11
    # on the faces themselves (vertically): every parcel is counted once, in one of the two boxes that share the face
    P, X, Y = rng.choice(pz_near, m), rng.uniform(-179, 179, m), rng.uniform(-89, 89, m)
    (cnt, s, sq), (c0, s0, sq0) = binned(P, X, Y, rng.permutation(m))
    assert cnt.sum() == m == c0.sum()
    col = lambda c: c.reshape(nx, ny, nz).sum(axis=2)
    assert np.array_equal(col(cnt), col(c0))


@pytest.mark.parametrize("levels", [False, True], ids=["pressure_levels", "model_levels"])
def test_the_engines_own_parcel_order_is_invisible(oracle, monkeypatch, levels):
    """Between two cell sorts the engine lays the parcels out by the cell their lookups fall into, not by the reference's sort
    key (engine.cu do_sort): slots, random numbers per slot, the slot-bound uvwp / dt handed to the slot's new parcel,
    the reference's order restored whenever parcels are addressed from outside.  With that order forced on at every sort
    (MPTRAC_B200_PRIVATE_ORDER=2) and switched off (=0) every array the API returns must be the same bit for bit --
    diffusion, mesoscale memory and sedimentation on, sorts every second step, a read-back in the middle of the run,
    np < np_max; and both must agree with the oracle like every other run."""
    from mptrac_b200 import Ctl, synth
    from oracle.oracle import Parcels
    m0, m1, tm, p, lon, lat, clim = _case(n=20011, grid=(48, 25, 24), seed=13)
    if levels:
        m0, m1 = synth.add_model_levels(m0, npl=24), synth.add_model_levels(m1, npl=24)
    n = tm.size
    q = np.stack([np.full(n, 2.0), np.full(n, 1500.0)])
    kw = dict(nq=2, qnt_rp=0, qnt_rhop=1, advect=4, diffusion=1, t_start=0.0, t_stop=1e6, dt_mod=300.0, dt_met=21600.0,
              turb_dz_trop=0.5, turb_dx_strat=20.0, sort_dt=600.0)
    if levels:
        kw.update(advect_vert_coord=2)
    ctl = Ctl(**kw)
    runs = {}
    for mode in ("2", "0"):
        monkeypatch.setenv("MPTRAC_B200_PRIVATE_ORDER", mode)
        with _engine(n + 77, 2) as eng:
            _setup(eng, ctl, clim, m0, m1, tm, p, lon, lat, q)
            for s in range(5):
                eng.run_timestep(300.0 * s)
            mid = eng.get_atm()                               # restores the reference's order; the run goes on from there
            for s in range(5, 11):
                eng.run_timestep(300.0 * s)
            runs[mode] = dict(mid=mid, end=eng.get_atm(), uvwp=eng.get_uvwp(), dt=eng.get_dt())
    a, b = runs["2"], runs["0"]
    for when in ("mid", "end"):
        for k in ("time", "lon", "lat", "p", "q"):
            assert np.array_equal(a[when][k], b[when][k]), f"{when}: {k} depends on the order in memory"
    assert np.array_equal(a["uvwp"], b["uvwp"]) and np.array_equal(a["dt"], b["dt"])
    assert np.abs(a["uvwp"]).max() > 0
    ref = Parcels(tm, p, lon, lat, q)
    oracle.ctr = 0
    oracle.run("timestep", ctl, clim, m0, m1, ref, t=0.0, nsteps=11)
    _compare("own_order", a["end"], ref, TOL_POS_DEG_DIFF, TOL_P_REL_DIFF)


def test_inactive_and_ragged_parcels(oracle):
    """parcels that start later / are already past t_stop keep dt = 0 and must not be touched; np < np_max"""
    from mptrac_b200 import Ctl
    from oracle.oracle import Parcels
    m0, m1, tm, p, lon, lat, clim = _case(n=3000)
    n = tm.size
    tm = np.where(np.arange(n) % 3 == 0, 900.0, 0.0)      # a third is released at t = 900 s
    ctl = Ctl(advect=2, diffusion=1, t_start=0.0, t_stop=1200.0, dt_mod=300.0, dt_met=21600.0)
    with _engine(5000) as eng:
        _setup(eng, ctl, clim, m0, m1, tm, p, lon, lat)
        for s in range(6):                                  # runs past t_stop
            eng.run_timestep(300.0 * s)
        out = eng.get_atm()
    ref = Parcels(tm, p, lon, lat)
    oracle.ctr = 0
    oracle.run("timestep", ctl, clim, m0, m1, ref, t=0.0, nsteps=6)
    _compare("ragged", out, ref, TOL_POS_DEG_DIFF, TOL_P_REL_DIFF)
    # t_stop is inclusive (direction * (time - t_stop) <= 0, src/mptrac.c:6019-6024): a parcel AT t_stop still takes
    # the step to t = 1500 s, then nothing moves any more
    assert np.all(out["time"] == 1500.0)


def test_tma_staging_strict_library(oracle):
    """the strict library stages the parcel stream with bulk copies (TMA: one cp.async.bulk per 256-byte tile, one
    mbarrier per warp and slot) where the production library uses per-lane cp.async: ragged size (partial last tile),
    cell sort, diffusion and sedimentation, against the oracle and against the production library"""
    from mptrac_b200 import Ctl
    from oracle.oracle import Parcels
    m0, m1, tm, p, lon, lat, clim = _case(n=100_003, grid=(72, 37, 30))
    n = tm.size
    q = np.stack([np.full(n, 2.0), np.full(n, 1500.0)])
    ctl = Ctl(nq=2, qnt_rp=0, qnt_rhop=1, advect=4, diffusion=1, t_start=0.0, t_stop=1e6, dt_mod=300.0, dt_met=21600.0,
              turb_dz_trop=0.5, turb_dx_strat=20.0)
    outs = []
    for strict in (True, False):
        with _engine(n, 2, strict=strict) as eng:
            _setup(eng, ctl, clim, m0, m1, tm, p, lon, lat, q)
            for s in range(5):
                eng.run_timestep(300.0 * s)
            outs.append(eng.get_atm())
    ref = Parcels(tm, p, lon, lat, q)
    oracle.ctr = 0
    oracle.run("timestep", ctl, clim, m0, m1, ref, t=0.0, nsteps=5)
    _compare("tma_staging_strict", outs[0], ref, TOL_POS_DEG_DIFF, TOL_P_REL_DIFF)
    assert abserr(outs[0]["lat"], outs[1]["lat"]) < TOL_POS_DEG_DIFF and relerr(outs[0]["p"], outs[1]["p"]) < TOL_P_REL_DIFF
    assert abserr(outs[0]["lat"], lat) > 1e-3


@pytest.mark.parametrize("layout", ["separate_arrays", "atm_t_layout", "pinned", "pinned:staggered", "pinned:explicit_time", "pinned:dma_in",
                                    "pinned:dma_out", "pinned:copy"])
def test_host_resident_step_equals_three_calls(layout, monkeypatch):
    """mpb_run_timestep_host (chunked upload / step / download pipeline) == set_atm + run_timestep + get_atm, bit for bit,
    including steps that fall back because a cell sort is due, with diffusion (random numbers addressed per chunk) and
    sedimentation (rp / rhop uploaded per chunk)."""
    from mptrac_b200 import Ctl
    staggered = False
    if ":" in layout:   # how pinned arrays cross the host link (MPTRAC_B200_HOST_MODE); the default is zerocopy
        layout, mode = layout.split(":")
        if mode == "staggered":            # parcels released at different times: time[] has to travel (default zero-copy path)
            staggered = True
        elif mode == "explicit_time":      # one common time, but the array is moved anyway
            monkeypatch.setenv("MPTRAC_B200_HOST_TIME", "explicit")
        else:
            monkeypatch.setenv("MPTRAC_B200_HOST_MODE", mode)
    m0, m1, tm, p, lon, lat, clim = _case(n=300_000, grid=(72, 37, 30))
    n = tm.size
    if staggered:
        tm = np.where(np.arange(n) % 5 == 0, 6000.0, 0.0)    # (still waiting at the last step: the times never become one)
    q = np.stack([np.full(n, 2.0), np.full(n, 1500.0)])
    ctl = Ctl(nq=2, qnt_rp=0, qnt_rhop=1, advect=4, diffusion=1, t_start=0.0, t_stop=1e6, dt_mod=300.0, dt_met=21600.0,
              turb_dz_trop=0.5, turb_dx_strat=20.0, sort_dt=900.0)
    with _engine(n, 2) as a, _engine(n, 2) as b:
        _setup(a, ctl, clim, m0, m1, tm, p, lon, lat, q)
        _setup(b, ctl, clim, m0, m1, tm, p, lon, lat, q)
        a.run_timestep(0.0)
        b.run_timestep(0.0)                          # fmod(0, SORT_DT) == 0: the very first step already sorts
        h = b.get_atm()                              # the host copy a driver would hold from here on
        hq = h["q"]
        if layout != "separate_arrays":              # time, p, lon, lat as consecutive rows of one block (atm_t): 2-D copies
            blk = np.stack([h[k] for k in ("time", "p", "lon", "lat")])
            if layout == "pinned":                   # page-locked: the kernel reads / writes the host arrays itself
                import torch
                keep = (torch.from_numpy(blk).pin_memory(), torch.from_numpy(np.ascontiguousarray(hq)).pin_memory())
                blk, hq = keep[0].numpy(), keep[1].numpy()
            h = {k: blk[i] for i, k in enumerate(("time", "p", "lon", "lat"))}
        for s in range(1, 5):                        # t = 900 triggers the sort -> fallback path
            a.run_timestep(300.0 * s)
            b.run_timestep_host(300.0 * s, h["time"], h["p"], h["lon"], h["lat"], hq)
        ref = a.get_atm()
        assert a.rng_ctr == b.rng_ctr
        uva, uvb = a.get_uvwp(), b.get_uvwp()
        if layout == "pinned" and "MPTRAC_B200_HOST_MODE" not in os.environ:
            # the last step was a plain zero-copy step: time[] stays at home when (and only when) all parcels share one time
            moved = b.host_step_bytes
            uniform = not staggered and "MPTRAC_B200_HOST_TIME" not in os.environ
            assert moved == ((24 + 16) * n, 24 * n) if uniform else moved == ((32 + 16) * n, 32 * n), moved
    for k in ("time", "p", "lon", "lat"):
        assert np.array_equal(ref[k], h[k]), k
    assert np.array_equal(uva, uvb)
    assert abserr(h["lat"], lat) > 1e-3


# ------------------------------------------------------------------------------------------------------------------
# full-size (BASELINE configs[1]) properties
# ------------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def c2():
    from mptrac_b200 import synth
    m0, m1 = synth.make_met_pair(360, 181, 60, t0=0.0, dt_met=21600.0)
    tm, p, lon, lat = synth.make_parcels(1_000_000, t0=0.0, seed=123)
    return m0, m1, tm, p, lon, lat, synth.make_clim_tropo()


def test_fullsize_subsample_vs_oracle(oracle, c2):
    """1 M parcels, 1 deg x 60 levels, RK4, 12 steps on the GPU; parcels are independent without diffusion, so a random
    subset re-run through the oracle must land on the same positions."""
    from mptrac_b200 import Ctl
    from oracle.oracle import Parcels
    m0, m1, tm, p, lon, lat, clim = c2
    n = tm.size
    ctl = Ctl(advect=4, t_start=0.0, t_stop=1e6, dt_mod=300.0, dt_met=21600.0)
    with _engine(n) as eng:
        _setup(eng, ctl, clim, m0, m1, tm, p, lon, lat)
        for s in range(13):
            eng.run_timestep(300.0 * s)
        out = eng.get_atm()
    idx = np.random.default_rng(0).choice(n, 20000, replace=False)
    ref = Parcels(tm[idx], p[idx], lon[idx], lat[idx])
    oracle.run("timestep", ctl, clim, m0, m1, ref, t=0.0, nsteps=13)
    _compare("fullsize_c2_subsample", {k: out[k][idx] for k in ("time", "lon", "lat", "p")}, ref)


def test_fullsize_sharded_equals_single(c2):
    """two contexts that each own half of the parcels (set_shard) reproduce the single-context run bit for bit,
    diffusion included: random numbers are addressed by global parcel index"""
    from mptrac_b200 import Ctl
    m0, m1, tm, p, lon, lat, clim = c2
    n = 400_000
    ctl = Ctl(advect=2, diffusion=1, t_start=0.0, t_stop=1e6, dt_mod=300.0, dt_met=21600.0)
    outs = []
    for parts in ([(0, n)], [(0, n // 2), (n // 2, n)]):
        res = {k: np.empty(n) for k in ("lon", "lat", "p")}
        for a, b in parts:
            with _engine(b - a) as eng:
                _setup(eng, ctl, clim, m0, m1, tm[a:b], p[a:b], lon[a:b], lat[a:b])
                eng.set_shard(a, n)
                for s in range(4):
                    eng.run_timestep(300.0 * s)
                o = eng.get_atm()
                for k in res:
                    res[k][a:b] = o[k]
        outs.append(res)
    for k in ("lon", "lat", "p"):
        assert np.array_equal(outs[0][k], outs[1][k])


def test_fullsize_sort_is_a_permutation_and_keeps_results(oracle, c2):
    """sorting by met cell (SORT_DT) must not change where (diffusion-free) parcels end up, only their order"""
    from mptrac_b200 import Ctl
    m0, m1, tm, p, lon, lat, clim = c2
    n = tm.size
    ident = np.arange(n, dtype=np.float64)[None, :]
    res = []
    for sort_dt in (-999.0, 600.0):
        ctl = Ctl(nq=1, advect=4, t_start=0.0, t_stop=1e6, dt_mod=300.0, dt_met=21600.0, sort_dt=sort_dt)
        with _engine(n, 1) as eng:
            _setup(eng, ctl, clim, m0, m1, tm, p, lon, lat, ident)
            for s in range(5):
                eng.run_timestep(300.0 * s)
            res.append(eng.get_atm())
    a, b = res
    order = b["q"][0].astype(np.int64)
    assert np.array_equal(np.sort(order), np.arange(n))
    for k in ("lon", "lat", "p", "time"):
        assert np.array_equal(a[k][order], b[k])


def test_fullsize_diffusion_moments(c2):
    """statistical check at full size (north_star: moments when diffusion is on): the mean turbulent displacement over
    1 M parcels is zero and its spread matches sqrt(2 K dt)"""
    from mptrac_b200 import Ctl
    m0, m1, tm, p, lon, lat, clim = c2
    n = tm.size
    lat0 = np.clip(lat, -60, 60)
    ctl = Ctl(advect=0, diffusion=1, t_start=0.0, t_stop=1e6, dt_mod=300.0, dt_met=21600.0, turb_mesox=0.0, turb_mesoz=0.0,
              turb_dx_pbl=50.0, turb_dx_trop=50.0, turb_dx_strat=50.0, turb_dz_strat=0.0)
    with _engine(n) as eng:
        _setup(eng, ctl, clim, m0, m1, tm, p, lon, lat0)
        eng.run_timestep(0.0)
        eng.run_timestep(300.0)
        out = eng.get_atm()
    dy_m = (out["lat"] - lat0) * np.pi / 180.0 * 6367.421e3
    sigma = np.sqrt(2 * 50.0 * 300.0)
    _report("diffusion_moments", mean_over_sigma=np.mean(dy_m) / sigma, std_over_sigma=np.std(dy_m) / sigma)
    assert abs(np.mean(dy_m)) < 5 * sigma / np.sqrt(n)
    assert abs(np.std(dy_m) / sigma - 1) < 5e-3


def test_smoke_entry():
    import __graft_entry__ as g
    g.smoke()
