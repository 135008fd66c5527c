"""GPU parity of the widened rows: the remaining module_meteo fields, module_convection, module_decay,
module_isosurf, module_diff_pbl, module_bound_cond, module_chem_grid -- each inside the dispatcher against the oracle --
and the reference's own trac_test / interoper_test through the shim with these modules routed to the device.
First run on a B200: round 2 (profiles/r02_gpu_first_call.txt)."""
import numpy as np
import pytest

from conftest import abserr, has_gpu, relerr

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not has_gpu(), reason="needs a CUDA device")]


@pytest.mark.parametrize("lat_desc", [False, True])
def test_module_meteo_all_fields_vs_oracle(oracle, lat_desc):
    """all 53 module_meteo quantities on the device (meteo_kernel + meteo_fields_kernel), met levels uploaded in swapped
    order and exchanged with mpb_swap_met; 1e-12 relative like test_module_meteo_vs_oracle"""
    from mptrac_b200 import Ctl, Engine, synth
    from mptrac_b200.host import METEO_QNT
    from oracle.oracle import Parcels
    m0, m1 = synth.make_met_pair(48, 25, 24, t0=0.0, dt_met=21600.0, lat_descending=lat_desc)
    m0, m1 = synth.add_meteo_fields(m0), synth.add_meteo_fields(m1)
    n = 6000
    tm, p, lon, lat = synth.make_parcels(n, t0=0.0, zmin=0.05, zmax=45.0, seed=4)
    tm = tm + np.random.default_rng(3).uniform(0.0, 21600.0, n)
    qm = {name: i for i, name in enumerate(METEO_QNT)}
    ctl = Ctl(nq=len(qm), t_start=0.0, t_stop=1e6, dt_mod=300.0, dt_met=21600.0, met_dt_out=300.0, qnt_meteo=qm)
    q0 = np.zeros((len(qm), n))
    with Engine(n, nq=len(qm), device=0) as eng:
        eng.set_ctl(ctl)
        eng.set_clim_tropo(*synth.make_clim_tropo())
        eng.set_met(0, m1)
        eng.set_met(1, m0)
        eng.swap_met()
        eng.set_atm(tm, p, lon, lat, q0)
        eng.module_meteo()
        out = eng.get_atm()
    ref = Parcels(tm, p, lon, lat, q0)
    oracle.run("meteo", ctl, synth.make_clim_tropo(), m0, m1, ref, t=300.0)
    for name, i in qm.items():
        a, b = out["q"][i], ref.q[i]
        assert np.array_equal(np.isnan(a), np.isnan(b)), name
        ok = ~np.isnan(b)
        scale = np.max(np.abs(b[ok]))
        assert scale > 0, name
        assert abserr(a[ok], b[ok]) / scale < 1e-12, name


def test_module_meteo_field_quantity_without_its_field_fails_loudly():
    from mptrac_b200 import Ctl, Engine, synth
    m0, m1 = synth.make_met_pair(24, 13, 10, t0=0.0, dt_met=21600.0)
    tm, p, lon, lat = synth.make_parcels(100, t0=0.0)
    ctl = Ctl(nq=1, t_start=0.0, t_stop=1e6, dt_mod=300.0, dt_met=21600.0, met_dt_out=300.0, qnt_meteo={"o3": 0})
    with Engine(100, nq=1, device=0) as eng:
        eng.set_ctl(ctl)
        eng.set_met(0, m0)
        eng.set_met(1, m1)
        eng.set_atm(tm, p, lon, lat, np.zeros((1, 100)))
        with pytest.raises(RuntimeError, match="met field"):
            eng.module_meteo()


@pytest.mark.timeout(900)
@pytest.mark.parametrize("levels", ["pl", "ml"])
def test_trac_trac_test_hybrid_host_modules(tmp_path, monkeypatch, levels):
    """tests/trac_test through the shim in full hybrid mode (MPTRAC_B200_DEVICE_MODULES=0, _DEVICE_METEO_FIELDS=0: the
    shim's defaults route module_convection, zg / pv / pt ... to the device, test_shim_trac.py covers that): convection,
    the further module_meteo quantities, chemistry and deposition all run through the reference's CPU code between device
    segments"""
    import test_shim_trac as T
    monkeypatch.setenv("MPTRAC_B200_DEVICE_MODULES", "0")
    monkeypatch.setenv("MPTRAC_B200_DEVICE_METEO_FIELDS", "0")
    T.test_trac_trac_test_through_the_shim(tmp_path, levels)


@pytest.mark.timeout(900)
def test_trac_interoper_test_hybrid_host_modules(tmp_path, monkeypatch):
    """tests/interoper_test (zeta coordinate) in full hybrid mode: module_decay and module_meteo (pv) on the host"""
    import test_shim_trac as T
    monkeypatch.setenv("MPTRAC_B200_DEVICE_MODULES", "0")
    monkeypatch.setenv("MPTRAC_B200_DEVICE_METEO_FIELDS", "0")
    T.test_trac_interoper_test_zeta_through_the_shim(tmp_path)


@pytest.mark.timeout(900)
def test_trac_trac_test_pl_with_device_meteo_fields(tmp_path, monkeypatch):
    """tests/trac_test (pressure-level run) through the shim with the modules on the host but zg, pv and pt from the
    device's module_meteo"""
    import test_shim_trac as T
    monkeypatch.setenv("MPTRAC_B200_DEVICE_MODULES", "0")
    T.test_trac_trac_test_through_the_shim(tmp_path, "pl")


@pytest.mark.parametrize("mix_pbl,cape,cin", [(1, -999.0, -999.0), (0, 100.0, -999.0), (1, 50.0, 10.0)])
@pytest.mark.parametrize("vert_coord", [0, 2])
def test_convection_in_the_step_vs_oracle(oracle, mix_pbl, cape, cin, vert_coord):
    """module_convection between diff_meso and sedi (the fused step splits around it; with model-level advection five
    launches), its uniform random numbers taken from the shared counter stream"""
    from mptrac_b200 import Ctl, Engine, synth
    from oracle.oracle import Parcels
    m0, m1 = synth.make_met_pair(48, 25, 24, t0=0.0, dt_met=21600.0)
    m0, m1 = synth.add_meteo_fields(m0), synth.add_meteo_fields(m1)
    if vert_coord:
        m0, m1 = synth.add_model_levels(m0, npl=30), synth.add_model_levels(m1, npl=30)
    n = 6000
    tm, p, lon, lat = synth.make_parcels(n, t0=0.0, zmin=0.0, zmax=14.0, seed=6)
    q = np.stack([np.full(n, 2.0), np.full(n, 1500.0)])
    ctl = Ctl(nq=2, qnt_rp=0, qnt_rhop=1, advect=2, advect_vert_coord=vert_coord, diffusion=1, t_start=0.0, t_stop=1e6, dt_mod=300.0,
              dt_met=21600.0, turb_dz_trop=0.5, turb_mesox=0.16, turb_mesoz=0.16, conv_mix_pbl=mix_pbl, conv_cape=cape, conv_cin=cin,
              conv_pbl_trans=0.2 if mix_pbl else 0.0)
    clim = synth.make_clim_tropo()
    with Engine(n, nq=2, device=0) as eng:
        eng.set_ctl(ctl)
        eng.set_clim_tropo(*clim)
        eng.set_met(0, m0)
        eng.set_met(1, m1)
        eng.set_atm(tm, p, lon, lat, q)
        for s in range(5):
            eng.run_timestep(300.0 * s)
        out = eng.get_atm()
        ctr = eng.rng_ctr
    ref = Parcels(tm, p, lon, lat, q)
    oracle.ctr = 0
    oracle.run("timestep", ctl, clim, m0, m1, ref, t=0.0, nsteps=5)
    assert ctr == oracle.ctr
    # diffusion on: normals differ by ~1e-7 relative between CUDA's and glibc's sinf / cosf (see test_gpu_parity.py)
    assert abserr(out["lat"], ref.lat) < 1e-7 and abserr(out["time"], ref.time) == 0
    # a parcel on the edge of the mixing range may fall on the other side: allow a handful of them
    bad = np.abs(out["p"] - ref.p) > 1e-6 * ref.p
    assert bad.mean() < 2e-3, bad.mean()
    assert np.mean(ref.p != p) > 0.02


def test_decay_in_the_step_vs_oracle(oracle):
    from mptrac_b200 import Ctl, Engine, synth
    from oracle.oracle import Parcels
    m0, m1 = synth.make_met_pair(36, 19, 20, t0=0.0, dt_met=21600.0)
    n = 4000
    tm, p, lon, lat = synth.make_parcels(n, t0=0.0, zmin=0.1, zmax=40.0, seed=9)
    q = np.random.default_rng(4).uniform(0.5, 2.0, (4, n))
    ctl = Ctl(nq=4, advect=4, t_start=0.0, t_stop=1e6, dt_mod=300.0, dt_met=21600.0, tdec_trop=86400.0, tdec_strat=10 * 86400.0,
              qnt_m=0, qnt_vmr=1, qnt_mloss_decay=2, qnt_loss_rate=3)
    clim = synth.make_clim_tropo()
    with Engine(n, nq=4, device=0) as eng:
        eng.set_ctl(ctl)
        eng.set_clim_tropo(*clim)
        eng.set_met(0, m0)
        eng.set_met(1, m1)
        eng.set_atm(tm, p, lon, lat, q)
        for s in range(4):
            eng.run_timestep(300.0 * s)
        out = eng.get_atm()
    ref = Parcels(tm, p, lon, lat, q)
    oracle.run("timestep", ctl, clim, m0, m1, ref, t=0.0, nsteps=4)
    assert abserr(out["lat"], ref.lat) < 1e-11
    for i in range(4):
        assert relerr(out["q"][i], ref.q[i]) < 1e-12, i
    assert np.all(ref.q[0] < q[0]) and np.all(ref.q[3] > 0)


@pytest.mark.parametrize("isosurf", [1, 2, 3, 4])
def test_isosurf_in_the_step_vs_oracle(oracle, isosurf):
    """module_isosurf_init at t_start and module_isosurf between sedi and the final position check (which then runs as
    its own launch); ISOSURF 4 with a balloon series uploaded through mpb_set_balloon"""
    from mptrac_b200 import Ctl, Engine, synth
    from oracle.oracle import Parcels
    m0, m1 = synth.make_met_pair(48, 25, 24, t0=0.0, dt_met=21600.0)
    n = 3000
    tm, p, lon, lat = synth.make_parcels(n, t0=0.0, zmin=1.0, zmax=30.0, seed=12)
    clim = synth.make_clim_tropo()
    ctl = Ctl(advect=4, t_start=0.0, t_stop=1e6, dt_mod=300.0, dt_met=21600.0, isosurf=isosurf, sort_dt=600.0)
    balloon = (np.array([0.0, 400.0, 900.0, 1300.0]), np.array([500.0, 420.0, 380.0, 300.0]))
    with Engine(n, nq=0, device=0) as eng:
        eng.set_ctl(ctl)
        eng.set_clim_tropo(*clim)
        eng.set_met(0, m0)
        eng.set_met(1, m1)
        if isosurf == 4:
            eng.set_balloon(*balloon)
        eng.set_atm(tm, p, lon, lat)
        for s in range(6):
            eng.run_timestep(300.0 * s)
        out = eng.get_atm()
        iso = eng.get_iso_var()
    ref = Parcels(tm, p, lon, lat)
    ref.balloon = balloon
    oracle.run("timestep", ctl, clim, m0, m1, ref, t=0.0, nsteps=6)
    assert abserr(out["lat"], ref.lat) < 1e-11 and abserr(out["time"], ref.time) == 0
    assert relerr(out["p"], ref.p) < 1e-12
    if isosurf != 4:
        assert relerr(iso, ref.iso_var) < 1e-13


def test_diff_pbl_in_the_step_vs_oracle(oracle):
    """module_diff_pbl (TURB_PBL_SCHEME 1) between diff_turb and diff_meso: the fused step splits around it; its normals
    come from the shared counter stream between those of the two other diffusion modules"""
    from mptrac_b200 import Ctl, Engine, synth
    from oracle.oracle import Parcels
    m0, m1 = synth.make_met_pair(48, 25, 24, t0=0.0, dt_met=21600.0)
    m0, m1 = synth.add_meteo_fields(m0, with_gaps=False), synth.add_meteo_fields(m1, with_gaps=False)
    n = 6000
    tm, p, lon, lat = synth.make_parcels(n, t0=0.0, zmin=0.0, zmax=4.0, seed=21)
    clim = synth.make_clim_tropo()
    ctl = Ctl(advect=2, diffusion=1, turb_pbl_scheme=1, turb_dz_trop=0.5, turb_dx_pbl=30.0, turb_dz_pbl=1.0, turb_mesox=0.16,
              turb_mesoz=0.16, t_start=0.0, t_stop=1e6, dt_mod=300.0, dt_met=21600.0)
    with Engine(n, nq=0, device=0) as eng:
        eng.set_ctl(ctl)
        eng.set_clim_tropo(*clim)
        eng.set_met(0, m0)
        eng.set_met(1, m1)
        eng.set_atm(tm, p, lon, lat)
        for s in range(4):
            eng.run_timestep(300.0 * s)
        out = eng.get_atm()
        ctr = eng.rng_ctr
        uv = eng.get_uvwp()
    ref = Parcels(tm, p, lon, lat)
    oracle.ctr = 0
    oracle.run("timestep", ctl, clim, m0, m1, ref, t=0.0, nsteps=4)
    assert ctr == oracle.ctr
    assert abserr(out["lat"], ref.lat) < 1e-7 and abserr(out["time"], ref.time) == 0
    # the closure switches regime and the reflection switches side at thresholds: a few parcels may fall on the other side
    # (diff_pbl keeps a velocity in m/s in the uvwp slot that diff_meso then moves the pressure with as hPa/s -- the
    # reference shares cache->uvwp between the two modules likewise -- so some pressures of this configuration are negative)
    bad = np.abs(out["p"] - ref.p) > 1e-6 * np.abs(ref.p)
    assert bad.mean() < 2e-3, bad.mean()
    assert np.mean(np.abs(uv - ref.uvwp) > 1e-4 * (1 + np.abs(ref.uvwp))) < 2e-3


@pytest.mark.parametrize("layer", ["none", "dps", "dzs_pbl", "zetas"])
def test_bound_cond_in_the_step_vs_oracle(oracle, layer):
    """module_bound_cond after module_meteo and again at the end of the step, with mass, volume mixing ratio, two trace-gas
    time series (mpb_set_clim_ts) and age of air"""
    from mptrac_b200 import Ctl, Engine, synth
    from oracle.oracle import Parcels
    m0, m1 = synth.make_met_pair(36, 19, 20, t0=0.0, dt_met=21600.0)
    n = 4000
    tm, p, lon, lat = synth.make_parcels(n, t0=0.0, zmin=0.0, zmax=12.0, seed=31)
    clim = synth.make_clim_tropo()
    series = {"Cccl3f": (np.array([-1e4, 500.0, 2000.0, 1e5]), np.array([2e-10, 2.2e-10, 2.1e-10, 1.9e-10])),
              "Csf6": (np.array([0.0, 1000.0]), np.array([1e-11, 1.2e-11]))}
    oracle.set_cts(series)
    kw = dict(none={}, dps=dict(bound_dps=150.0), dzs_pbl=dict(bound_dzs=1.5, bound_pbl=1), zetas=dict(bound_zetas=320.0))[layer]
    ctl = Ctl(nq=5, advect=4, t_start=0.0, t_stop=1e6, dt_mod=300.0, dt_met=21600.0, qnt_m=0, qnt_vmr=1, qnt_aoa=2,
              qnt_cts=(-1, 3, -1, -1, 4), cts_on=0b10010, bound_mass=3.0, bound_mass_trend=1e-3, bound_vmr=1e-9, bound_lat0=-60.0,
              bound_lat1=70.0, bound_p0=1e10, bound_p1=300.0, **kw)
    q = np.random.default_rng(1).uniform(0.5, 2.0, (5, n))
    with Engine(n, nq=5, device=0) as eng:
        eng.set_ctl(ctl)
        eng.set_clim_tropo(*clim)
        eng.set_met(0, m0)
        eng.set_met(1, m1)
        eng.set_clim_ts(1, *series["Cccl3f"])
        eng.set_clim_ts(4, *series["Csf6"])
        eng.set_atm(tm, p, lon, lat, q)
        for s in range(4):
            eng.run_timestep(300.0 * s)
        out = eng.get_atm()
    ref = Parcels(tm, p, lon, lat, q)
    oracle.run("timestep", ctl, clim, m0, m1, ref, t=0.0, nsteps=4)
    assert abserr(out["lat"], ref.lat) < 1e-11
    # a parcel on the edge of the window or of the surface layer may fall on the other side
    bad = np.any(np.abs(out["q"] - ref.q) > 1e-12 * np.abs(ref.q), axis=0)
    assert bad.mean() < 2e-3, bad.mean()
    assert 0.05 < np.mean(ref.q[2] == ref.time) < 0.95


@pytest.mark.parametrize("nens", [0, 3])
def test_chem_grid_in_the_step_vs_oracle(oracle, nens):
    """module_chem_grid after the step: box masses by atomics (the oracle sums in parcel order: agreement to rounding)"""
    from mptrac_b200 import Ctl, Engine, synth
    from oracle.oracle import Parcels
    m0, m1 = synth.make_met_pair(36, 19, 20, t0=0.0, dt_met=21600.0)
    n = 20000
    tm, p, lon, lat = synth.make_parcels(n, t0=0.0, zmin=0.0, zmax=30.0, seed=41)
    rng = np.random.default_rng(3)
    q = np.zeros((3, n))
    q[0], q[2] = rng.uniform(0.5, 2.0, n), rng.integers(0, 3, n)
    ctl = Ctl(nq=3, advect=4, t_start=0.0, t_stop=1e6, dt_mod=300.0, dt_met=21600.0, qnt_m=0, qnt_Cx=1, qnt_ens=2, nens=nens, molmass=64.07,
              chemgrid_nx=36, chemgrid_ny=18, chemgrid_nz=15, chemgrid_z0=0.0, chemgrid_z1=30.0, chemgrid=1)
    clim = synth.make_clim_tropo()
    with Engine(n, nq=3, device=0) as eng:
        eng.set_ctl(ctl)
        eng.set_clim_tropo(*clim)
        eng.set_met(0, m0)
        eng.set_met(1, m1)
        eng.set_atm(tm, p, lon, lat, q)
        for s in range(3):
            eng.run_timestep(300.0 * s)
        out = eng.get_atm()
    ref = Parcels(tm, p, lon, lat, q)
    oracle.run("timestep", ctl, clim, m0, m1, ref, t=0.0, nsteps=3)
    assert abserr(out["lat"], ref.lat) < 1e-11
    # a parcel on a box edge may fall into the neighbouring box: its own value and that of the two boxes' other parcels differ
    bad = np.abs(out["q"][1] - ref.q[1]) > 1e-9 * np.max(ref.q[1])
    assert bad.mean() < 5e-3, bad.mean()
    assert np.mean(ref.q[1] > 0) > 0.5
