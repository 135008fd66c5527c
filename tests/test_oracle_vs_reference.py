"""Pin the oracle: the C restatement against the UNMODIFIED reference compiled into oracle/_ref (development
container only -- skipped where oracle/_ref does not exist).  Everything must be bit-exact."""
import numpy as np
import pytest

from conftest import ROOT, abserr


def _case(n, grid, lat_desc, seed=3):
    from mptrac_b200 import synth
    m0, m1 = synth.make_met_pair(*grid, t0=0.0, dt_met=21600.0, lat_descending=lat_desc)
    tm, p, lon, lat = synth.make_parcels(n, t0=0.0, zmin=0.05, zmax=45.0, seed=seed)
    return m0, m1, tm, p, lon, lat


def _same(a, b):
    return all(np.array_equal(getattr(a, k), getattr(b, k)) for k in ("time", "p", "lon", "lat", "q", "uvwp", "dt"))


@pytest.mark.parametrize("advect", [1, 2, 4])
@pytest.mark.parametrize("diffusion", [0, 1])
@pytest.mark.parametrize("lat_desc", [False, True])
@pytest.mark.parametrize("direction", [1, -1])
def test_run_timestep_bit_exact(oracle, reference, advect, diffusion, lat_desc, direction):
    from mptrac_b200 import Ctl
    from oracle.oracle import Parcels
    m0, m1, tm, p, lon, lat = _case(5000, (48, 25, 24), lat_desc)
    n = tm.size
    t_start = 0.0 if direction == 1 else 21600.0
    tm = np.full(n, t_start)
    rng = np.random.default_rng(0)
    q = np.stack([rng.uniform(0.1, 10, n), rng.uniform(500, 2500, n), rng.uniform(0, 1, n)])
    assert reference.read_ctl(["rp", "rhop", "m"]) == 3
    clim = reference.clim_tropo()
    reference.set_met(m0, m1)
    ctl = Ctl(nq=3, qnt_rp=0, qnt_rhop=1, advect=advect, diffusion=diffusion, direction=direction, t_start=t_start,
              t_stop=t_start + direction * 86400.0, dt_mod=300.0, dt_met=21600.0, turb_dz_trop=0.5, turb_dz_pbl=1.0,
              turb_dx_strat=20.0, turb_pbl_trans=0.3, mixing_trop=0.3, mixing_strat=0.1, mixing_dt=600.0, mix_qnt=[2],
              mixing_nx=36, mixing_ny=18, mixing_nz=20)
    a, b = Parcels(tm, p, lon, lat, q), Parcels(tm, p, lon, lat, q)
    reference.ctr = oracle.ctr = 5
    reference.run("timestep", ctl, a, t=t_start, nsteps=8)
    oracle.run("timestep", ctl, clim, m0, m1, b, t=t_start, nsteps=8)
    assert reference.ctr == oracle.ctr
    assert abserr(a.lat, lat) > 1e-3
    assert _same(a, b)


@pytest.mark.parametrize("module", ["timesteps", "position", "advect", "diff_turb", "diff_meso", "sedi", "mixing"])
def test_single_modules_bit_exact(oracle, reference, module):
    from mptrac_b200 import Ctl
    from oracle.oracle import Parcels
    m0, m1, tm, p, lon, lat = _case(4000, (36, 19, 20), True, seed=8)
    n = tm.size
    rng = np.random.default_rng(1)
    lon = lon + rng.choice([0.0, 360.0, -360.0], n)
    lat = np.where(rng.uniform(size=n) < 0.05, lat + 100.0, lat)
    p = np.where(rng.uniform(size=n) < 0.05, p * 1.5, p)
    tm = np.where(rng.uniform(size=n) < 0.1, 900.0, 0.0)
    q = np.stack([rng.uniform(0.1, 10, n), rng.uniform(500, 2500, n), rng.uniform(0, 1, n)])
    uvwp = rng.standard_normal((n, 3)).astype(np.float32)
    reference.read_ctl(["rp", "rhop", "m"])
    clim = reference.clim_tropo()
    reference.set_met(m0, m1)
    ctl = Ctl(nq=3, qnt_rp=0, qnt_rhop=1, advect=4, diffusion=1, t_start=0.0, t_stop=1e6, dt_mod=300.0, dt_met=21600.0,
              turb_dz_trop=0.5, turb_dz_pbl=1.0, turb_dx_strat=20.0, turb_pbl_trans=0.3, mixing_trop=0.3, mixing_strat=0.1,
              mix_qnt=[2], mixing_nx=36, mixing_ny=18, mixing_nz=20)
    a, b = Parcels(tm, p, lon, lat, q, uvwp), Parcels(tm, p, lon, lat, q, uvwp)
    reference.ctr = oracle.ctr = 99
    for x, run in ((a, lambda w, t: reference.run(w, ctl, a, t=t)), (b, lambda w, t: oracle.run(w, ctl, clim, m0, m1, b, t=t))):
        run("timesteps", 300.0)
        if module != "timesteps":
            run(module, 300.0)
    assert reference.ctr == oracle.ctr
    assert _same(a, b)


def test_rng_stream_bit_exact(oracle, reference):
    for method in (0, 1):
        for n in (10, 3001):
            reference.ctr = oracle.ctr = 424242
            assert np.array_equal(reference.module_rng(n, method), oracle.module_rng(n, method))
            assert reference.ctr == oracle.ctr == 424242 + n + 1


def test_sort_same_cells(oracle, reference):
    """module_sort: the reference (gsl_sort_index, a heapsort) leaves the order inside a cell unspecified, so the
    comparison is by cell key and by the multiset of parcels per key."""
    from mptrac_b200 import Ctl
    from oracle.oracle import Parcels
    m0, m1, tm, p, lon, lat = _case(20000, (36, 19, 20), False)
    n = tm.size
    q = np.arange(n, dtype=np.float64)[None, :]
    reference.read_ctl(["idx"])
    reference.set_met(m0, m1)
    ctl = Ctl(nq=1, advect=2, t_start=0.0, t_stop=1e6, dt_mod=300.0, dt_met=21600.0)
    a, b = Parcels(tm, p, lon, lat, q), Parcels(tm, p, lon, lat, q)
    reference.run("sort", ctl, a)
    oracle.run("sort", ctl, None, m0, m1, b)
    ka, kb = oracle.sort_keys(m0, a), oracle.sort_keys(m0, b)
    assert np.array_equal(ka, kb) and np.all(np.diff(kb) >= 0)
    # same parcels in every cell
    oa = np.lexsort((a.q[0], ka))
    ob = np.lexsort((b.q[0], kb))
    for k in ("lon", "lat", "p"):
        assert np.array_equal(getattr(a, k)[oa], getattr(b, k)[ob])


def test_sedi_bit_exact(oracle, reference):
    rng = np.random.default_rng(0)
    for _ in range(200):
        args = (rng.uniform(1, 1000), rng.uniform(180, 320), 10 ** rng.uniform(-2, 2), rng.uniform(500, 3000))
        assert reference.sedi(*args) == oracle.sedi(*args)


def test_clim_tropo_table_is_what_the_goldens_hold(reference):
    import numpy as np
    from conftest import GOLDEN
    t, la, tr = reference.clim_tropo()
    z = np.load(GOLDEN / "dt_test.npz")
    assert np.array_equal(tr, z["tropo"]) and np.array_equal(t, z["tropo_time"]) and np.array_equal(la, z["tropo_lat"])


@pytest.mark.parametrize("lat_desc", [False, True])
def test_module_meteo_bit_exact(oracle, reference, lat_desc):
    """module_meteo restricted to the quantities of the path (src/mptrac.c:5062-5165), all 14 of them, parcels at
    scattered times (check_dt = 0: every parcel is visited)"""
    from mptrac_b200 import Ctl, synth
    from mptrac_b200.host import METEO_QNT
    from oracle.oracle import Parcels
    m0, m1 = synth.make_met_pair(48, 25, 24, t0=0.0, dt_met=21600.0, lat_descending=lat_desc)
    n = 4000
    tm, p, lon, lat = synth.make_parcels(n, t0=0.0, zmin=0.05, zmax=45.0, seed=5)
    tm = tm + np.random.default_rng(1).uniform(0, 20000, n)
    names = list(METEO_QNT[:14])
    nq = reference.read_ctl(names, "")
    assert nq == len(names) and set(reference.qnt_meteo) == set(names)
    reference.set_met(m0, m1)
    ctl = Ctl(nq=nq, advect=4, t_start=0.0, t_stop=1e6, dt_mod=300.0, dt_met=21600.0, met_dt_out=0.1, qnt_meteo=reference.qnt_meteo)
    a = Parcels(tm, p, lon, lat, np.zeros((nq, n)))
    b = a.copy()
    reference.run("meteo", ctl, a)
    oracle.run("meteo", ctl, synth.make_clim_tropo(), m0, m1, b)
    for name, k in reference.qnt_meteo.items():
        assert np.array_equal(a.q[k], b.q[k]), name
    assert np.all(a.q[reference.qnt_meteo["t"]] > 150)


def test_run_timestep_with_meteo_bit_exact(oracle, reference):
    """the dispatcher with MET_DT_OUT on: module_meteo after the final position check, before mixing (src/mptrac.c:7927)"""
    from mptrac_b200 import Ctl, synth
    from oracle.oracle import Parcels
    m0, m1 = synth.make_met_pair(48, 25, 24, t0=0.0, dt_met=21600.0)
    n = 3000
    tm, p, lon, lat = synth.make_parcels(n, t0=0.0, zmin=0.05, zmax=45.0, seed=7)
    nq = reference.read_ctl(["t", "u", "v", "w", "theta"], "")
    reference.set_met(m0, m1)
    ctl = Ctl(nq=nq, advect=2, diffusion=1, t_start=0.0, t_stop=1e6, dt_mod=300.0, dt_met=21600.0, met_dt_out=0.1,
              qnt_meteo=reference.qnt_meteo)
    a = Parcels(tm, p, lon, lat, np.zeros((nq, n)))
    b = a.copy()
    reference.ctr = oracle.ctr = 0
    reference.run("timestep", ctl, a, t=0.0, nsteps=4)
    oracle.run("timestep", ctl, reference.clim_tropo(), m0, m1, b, t=0.0, nsteps=4)
    for k in ("lon", "lat", "p"):
        assert np.array_equal(getattr(a, k), getattr(b, k)), k
    assert np.array_equal(a.q, b.q) and np.all(a.q[0] > 150)


def test_c1_trac_test_parcels_midpoint_bit_exact(oracle, reference):
    """BASELINE configs[0]: the reference's tests/trac_test parcels on its ERA-Interim data, ADVECT 2, everything else
    off, one day at DT_MOD 300: the restatement follows the reference bit for bit over all 289 steps"""
    from mptrac_b200 import Ctl, Met
    from oracle.oracle import Parcels
    data = ROOT / "oracle" / "_ref" / "data"
    files = [data / "ei_2011_06_05_00.nc", data / "ei_2011_06_06_00.nc", data / "trac_test.ref" / "atm_init.tab"]
    if not all(f.exists() for f in files):
        pytest.skip("reference test data not present")
    reference.read_ctl([], "")
    m0, m1 = reference.read_met(files[0], 0, Met), reference.read_met(files[1], 1, Met)
    tm, p, lon, lat = reference.read_atm(files[2])
    t0 = m0.time
    ctl = Ctl(advect=2, diffusion=0, t_start=t0, t_stop=t0 + 86400.0, dt_mod=300.0, dt_met=86400.0)
    a = Parcels(tm, p, lon, lat)
    b = a.copy()
    reference.run("timestep", ctl, a, t=t0, nsteps=289)
    oracle.run("timestep", ctl, reference.clim_tropo(), m0, m1, b, t=t0, nsteps=289)
    for k in ("time", "lon", "lat", "p"):
        assert np.array_equal(getattr(a, k), getattr(b, k)), k
    assert np.max(np.abs(a.lat - lat)) > 1.0


@pytest.mark.parametrize("vert_coord", [1, 2, 3])
@pytest.mark.parametrize("advect", [1, 2, 4])
def test_model_level_advection_bit_exact(oracle, reference, vert_coord, advect):
    """ADVECT_VERT_COORD 1 / 2 / 3: intpol_met_4d_zeta, module_advect's model-level branches and module_advect_init
    (src/mptrac.c:2808-2981, 3646-3657, 3680-3785), with diffusion and sedimentation running after them."""
    from mptrac_b200 import Ctl, synth
    from oracle.oracle import Parcels
    m0, m1 = synth.make_met_pair(48, 25, 24, t0=0.0, dt_met=21600.0)
    m0, m1 = synth.add_model_levels(m0, npl=30), synth.add_model_levels(m1, npl=30)
    n = 4000
    tm, p, lon, lat = synth.make_parcels(n, t0=0.0, zmin=1.0, zmax=45.0, seed=5)
    names = ["rp", "rhop"] + (["zeta"] if vert_coord == 1 else ["eta"] if vert_coord == 3 else [])
    nq = reference.read_ctl(names, "")
    reference.set_met(m0, m1)
    clim = reference.clim_tropo()
    rng = np.random.default_rng(2)
    q = np.zeros((nq, n))
    q[0], q[1] = rng.uniform(0.1, 10, n), rng.uniform(500, 2500, n)
    if vert_coord == 1:
        q[reference.qnt["zeta"]] = rng.uniform(300.0, 1500.0, n)   # module_advect_init derives the pressure from it
    ctl = Ctl(nq=nq, qnt_rp=0, qnt_rhop=1, advect=advect, advect_vert_coord=vert_coord, diffusion=1, t_start=0.0, t_stop=1e6,
              dt_mod=300.0, dt_met=21600.0, turb_dz_trop=0.5, turb_dx_strat=20.0, turb_mesox=0.16, turb_mesoz=0.16,
              qnt_zeta=reference.qnt["zeta"], qnt_eta=reference.qnt["eta"])
    a = Parcels(tm, p, lon, lat, q)
    b = a.copy()
    reference.ctr = oracle.ctr = 0
    reference.run("timestep", ctl, a, t=0.0, nsteps=5)
    oracle.run("timestep", ctl, clim, m0, m1, b, t=0.0, nsteps=5)
    assert abserr(a.lat, lat) > 1e-3
    if vert_coord == 1:
        assert abserr(a.p, p) > 1.0
    assert _same(a, b)


@pytest.mark.parametrize("lat_desc", [False, True])
def test_module_meteo_all_fields_bit_exact(oracle, reference, lat_desc):
    """module_meteo with every quantity the met fields give (INTPOL_TIME_ALL: 13 3-D and 24 2-D fields, src/mptrac.h:1278)
    plus the ones derived from t and h2o -- 53 quantities; 2-D fields with NaN gaps (nearest-neighbour rule, 3084-3107)."""
    from mptrac_b200 import Ctl, synth
    from oracle.oracle import METEO_QNT, Parcels
    m0, m1 = synth.make_met_pair(48, 25, 24, t0=0.0, dt_met=21600.0, lat_descending=lat_desc)
    m0, m1 = synth.add_meteo_fields(m0), synth.add_meteo_fields(m1)
    n = 5000
    tm, p, lon, lat = synth.make_parcels(n, t0=0.0, zmin=0.05, zmax=45.0, seed=4)
    tm = tm + np.random.default_rng(3).uniform(0.0, 21600.0, n)
    assert len(METEO_QNT) == 53
    clim = reference.clim_tropo()
    seen = set()
    for k in range(0, len(METEO_QNT), 13):      # the reference build holds NQ = 15 quantities at most
        names = list(METEO_QNT[k:k + 13])
        nq = reference.read_ctl(names)
        assert nq == len(names) == len(reference.qnt_meteo)
        reference.set_met(m0, m1)
        ctl = Ctl(nq=nq, t_start=0.0, t_stop=1e6, dt_mod=300.0, dt_met=21600.0, met_dt_out=300.0, qnt_meteo=reference.qnt_meteo)
        a = Parcels(tm, p, lon, lat, np.zeros((nq, n)))
        b = a.copy()
        reference.run("meteo", ctl, a, t=300.0)
        oracle.run("meteo", ctl, clim, m0, m1, b, t=300.0)
        for name, i in reference.qnt_meteo.items():
            assert np.array_equal(a.q[i], b.q[i], equal_nan=True), name
            assert np.any(a.q[i] != 0), name
            seen.add(name)
    assert seen == set(METEO_QNT)


@pytest.mark.parametrize("mix_pbl,cape,cin", [(1, -999.0, -999.0), (0, 100.0, -999.0), (1, 50.0, 10.0)])
def test_module_convection_bit_exact(oracle, reference, mix_pbl, cape, cin):
    """module_convection (src/mptrac.c:4102-4171): PBL mixing, CAPE / CIN thresholds with the equilibrium level, NaN gaps in
    the equilibrium-level field; alone and inside the dispatcher (between diff_meso and sedi, its own random numbers)"""
    from mptrac_b200 import Ctl, synth
    from oracle.oracle import Parcels
    m0, m1 = synth.make_met_pair(48, 25, 24, t0=0.0, dt_met=21600.0)
    m0, m1 = synth.add_meteo_fields(m0), synth.add_meteo_fields(m1)
    n = 5000
    tm, p, lon, lat = synth.make_parcels(n, t0=0.0, zmin=0.0, zmax=14.0, seed=6)
    nq = reference.read_ctl(["rp", "rhop"])
    reference.set_met(m0, m1)
    clim = reference.clim_tropo()
    q = np.stack([np.full(n, 2.0), np.full(n, 1500.0)])
    ctl = Ctl(nq=nq, qnt_rp=0, qnt_rhop=1, advect=2, diffusion=1, t_start=0.0, t_stop=1e6, dt_mod=300.0, dt_met=21600.0,
              turb_dz_trop=0.5, turb_mesox=0.16, turb_mesoz=0.16, conv_mix_pbl=mix_pbl, conv_cape=cape, conv_cin=cin,
              conv_pbl_trans=0.2 if mix_pbl else 0.0)
    a = Parcels(tm, p, lon, lat, q)
    reference.ctr = oracle.ctr = 3
    reference.run("timesteps", ctl, a, t=300.0)
    b = a.copy()
    reference.run("convection", ctl, a, t=300.0)
    oracle.run("convection", ctl, clim, m0, m1, b, t=300.0)
    assert reference.ctr == oracle.ctr == 3 + n + 1
    assert _same(a, b)
    moved = np.mean(a.p != p)
    assert 0.02 < moved < 0.98, moved
    a = Parcels(tm, p, lon, lat, q)
    b = a.copy()
    reference.ctr = oracle.ctr = 0
    reference.run("timestep", ctl, a, t=0.0, nsteps=5)
    oracle.run("timestep", ctl, clim, m0, m1, b, t=0.0, nsteps=5)
    assert reference.ctr == oracle.ctr
    assert _same(a, b)


def test_module_decay_bit_exact(oracle, reference):
    """module_decay with mass, volume mixing ratio, the decay-loss and total-loss-rate quantities (src/mptrac.c:4227-4263,
    7931-7940); alone and inside the dispatcher"""
    from mptrac_b200 import Ctl, synth
    from oracle.oracle import Parcels
    m0, m1 = synth.make_met_pair(36, 19, 20, t0=0.0, dt_met=21600.0)
    n = 4000
    tm, p, lon, lat = synth.make_parcels(n, t0=0.0, zmin=0.1, zmax=40.0, seed=9)
    nq = reference.read_ctl(["m", "vmr", "mloss_decay", "loss_rate"])
    qi = reference.qnt
    reference.set_met(m0, m1)
    clim = reference.clim_tropo()
    rng = np.random.default_rng(4)
    q = rng.uniform(0.5, 2.0, (nq, n))
    ctl = Ctl(nq=nq, advect=4, t_start=0.0, t_stop=1e6, dt_mod=300.0, dt_met=21600.0, tdec_trop=86400.0, tdec_strat=10 * 86400.0,
              qnt_m=qi["m"], qnt_vmr=qi["vmr"], qnt_mloss_decay=qi["mloss_decay"], qnt_loss_rate=qi["loss_rate"])
    a = Parcels(tm, p, lon, lat, q)
    reference.run("timesteps", ctl, a, t=300.0)
    b = a.copy()
    reference.run("decay", ctl, a, t=300.0)
    oracle.run("decay", ctl, clim, m0, m1, b, t=300.0)
    assert _same(a, b) and np.all(a.q[qi["m"]] < q[qi["m"]])
    a = Parcels(tm, p, lon, lat, q)
    b = a.copy()
    reference.run("timestep", ctl, a, t=0.0, nsteps=4)
    oracle.run("timestep", ctl, clim, m0, m1, b, t=0.0, nsteps=4)
    assert _same(a, b)
    assert np.allclose(a.q[qi["loss_rate"]], 1.0 / 86400.0, rtol=0.9) and np.all(a.q[qi["loss_rate"]] > 0)


@pytest.mark.parametrize("isosurf", [1, 2, 3, 4])
def test_module_isosurf_bit_exact(oracle, reference, isosurf, tmp_path):
    """module_isosurf_init + module_isosurf (src/mptrac.c:4886-5004): parcels kept on pressure / density / potential
    temperature surfaces or on a balloon's pressure time series, inside the dispatcher (between sedi and the final
    position check, every parcel every step)"""
    from mptrac_b200 import Ctl, synth
    from oracle.oracle import Parcels
    m0, m1 = synth.make_met_pair(48, 25, 24, t0=0.0, dt_met=21600.0)
    n = 3000
    tm, p, lon, lat = synth.make_parcels(n, t0=0.0, zmin=1.0, zmax=30.0, seed=12)
    balloon = (np.array([0.0, 400.0, 900.0, 1300.0]), np.array([500.0, 420.0, 380.0, 300.0]))
    bfile = tmp_path / "balloon.tab"
    bfile.write_text("# t p\n" + "".join(f"{t:.1f} {q:.1f}\n" for t, q in zip(*balloon)))
    nq = reference.read_ctl([], f"BALLOON {bfile}" if isosurf == 4 else "")
    reference.set_met(m0, m1)
    clim = reference.clim_tropo()
    ctl = Ctl(nq=nq, advect=4, diffusion=1, turb_dz_trop=0.5, turb_mesox=0.16, turb_mesoz=0.16, t_start=0.0, t_stop=1e6, dt_mod=300.0,
              dt_met=21600.0, isosurf=isosurf)
    a = Parcels(tm, p, lon, lat)
    if isosurf == 4:
        a.balloon = balloon
    b = a.copy()
    reference.ctr = oracle.ctr = 0
    reference.run("timestep", ctl, a, t=0.0, nsteps=6)
    oracle.run("timestep", ctl, clim, m0, m1, b, t=0.0, nsteps=6)
    assert _same(a, b) and np.array_equal(a.iso_var, b.iso_var)
    assert abserr(a.lat, lat) > 1e-3
    if isosurf == 1:
        assert np.array_equal(a.p, p)
    if isosurf == 4:
        assert np.all(a.p == 300.0)


def test_module_diff_pbl_bit_exact(oracle, reference):
    """module_diff_pbl (src/mptrac.c:4343-4584, TURB_PBL_SCHEME 1): neutral, unstable and stable closures, reflection at the
    ground and the PBL top; alone and inside the dispatcher (between diff_turb, which then skips the PBL, and diff_meso)"""
    from mptrac_b200 import Ctl, synth
    from oracle.oracle import Parcels
    m0, m1 = synth.make_met_pair(48, 25, 24, t0=0.0, dt_met=21600.0)
    m0, m1 = synth.add_meteo_fields(m0, with_gaps=False), synth.add_meteo_fields(m1, with_gaps=False)
    n = 6000
    tm, p, lon, lat = synth.make_parcels(n, t0=0.0, zmin=0.0, zmax=4.0, seed=21)
    nq = reference.read_ctl([])
    reference.set_met(m0, m1)
    clim = reference.clim_tropo()
    ctl = Ctl(nq=nq, advect=2, diffusion=1, turb_pbl_scheme=1, turb_dz_trop=0.5, turb_dx_pbl=30.0, turb_dz_pbl=1.0, turb_mesox=0.16,
              turb_mesoz=0.16, t_start=0.0, t_stop=1e6, dt_mod=300.0, dt_met=21600.0)
    rng = np.random.default_rng(2)
    uvwp = rng.standard_normal((n, 3)).astype(np.float32)
    a = Parcels(tm, p, lon, lat, None, uvwp)
    reference.ctr = oracle.ctr = 11
    reference.run("timesteps", ctl, a, t=300.0)
    b = a.copy()
    reference.run("diff_pbl", ctl, a, t=300.0)
    oracle.run("diff_pbl", ctl, clim, m0, m1, b, t=300.0)
    assert reference.ctr == oracle.ctr == 11 + 3 * n + 1
    assert _same(a, b)
    inside = a.p != p
    assert 0.1 < inside.mean() < 0.95, inside.mean()
    a = Parcels(tm, p, lon, lat, None, uvwp)
    b = a.copy()
    reference.ctr = oracle.ctr = 0
    reference.run("timestep", ctl, a, t=0.0, nsteps=6)
    oracle.run("timestep", ctl, clim, m0, m1, b, t=0.0, nsteps=6)
    assert reference.ctr == oracle.ctr and _same(a, b)


@pytest.mark.parametrize("layer", ["none", "dps", "dzs_pbl", "zetas"])
def test_module_bound_cond_bit_exact(oracle, reference, layer):
    """module_bound_cond (src/mptrac.c:3789-3881) with mass, volume mixing ratio, two trace-gas time series and age of air,
    the latitude / pressure window and the four surface-layer tests; alone and inside the dispatcher (twice per step)"""
    from mptrac_b200 import Ctl, synth
    from oracle.oracle import Parcels
    m0, m1 = synth.make_met_pair(36, 19, 20, t0=0.0, dt_met=21600.0)
    n = 4000
    tm, p, lon, lat = synth.make_parcels(n, t0=0.0, zmin=0.0, zmax=12.0, seed=31)
    nq = reference.read_ctl(["m", "vmr", "aoa", "Cccl3f", "Csf6"])
    qi = reference.qnt
    reference.set_met(m0, m1)
    clim = reference.clim_tropo()
    series = {"Cccl3f": (np.array([-1e4, 500.0, 2000.0, 1e5]), np.array([2e-10, 2.2e-10, 2.1e-10, 1.9e-10])),
              "Csf6": (np.array([0.0, 1000.0]), np.array([1e-11, 1.2e-11]))}
    reference.set_cts(series)
    oracle.set_cts(series)
    kw = dict(none={}, dps=dict(bound_dps=150.0), dzs_pbl=dict(bound_dzs=1.5, bound_pbl=1), zetas=dict(bound_zetas=320.0))[layer]
    ctl = Ctl(nq=nq, advect=4, t_start=0.0, t_stop=1e6, dt_mod=300.0, dt_met=21600.0, qnt_m=qi["m"], qnt_vmr=qi["vmr"],
              qnt_aoa=qi["aoa"], qnt_cts=qi["cts"], cts_on=0b10010, bound_mass=3.0, bound_mass_trend=1e-3, bound_vmr=1e-9,
              bound_lat0=-60.0, bound_lat1=70.0, bound_p0=1e10, bound_p1=300.0, **kw)
    q = np.random.default_rng(1).uniform(0.5, 2.0, (nq, n))
    a = Parcels(tm, p, lon, lat, q)
    reference.run("timesteps", ctl, a, t=300.0)
    b = a.copy()
    reference.run("bound_cond", ctl, a, t=300.0)
    oracle.run("bound_cond", ctl, clim, m0, m1, b, t=300.0)
    assert _same(a, b)
    hit = a.q[qi["aoa"]] == a.time
    assert 0.05 < hit.mean() < 0.95, hit.mean()
    a = Parcels(tm, p, lon, lat, q)
    b = a.copy()
    reference.run("timestep", ctl, a, t=0.0, nsteps=4)
    oracle.run("timestep", ctl, clim, m0, m1, b, t=0.0, nsteps=4)
    assert _same(a, b)


@pytest.mark.parametrize("nens", [0, 3])
def test_module_chem_grid_bit_exact(oracle, reference, nens):
    """module_chem_grid (src/mptrac.c:3885-4054): box masses summed in parcel order, volume mixing ratio at the temperature of
    the box centre, with and without ensembles"""
    from mptrac_b200 import Ctl, synth
    from oracle.oracle import Parcels
    m0, m1 = synth.make_met_pair(36, 19, 20, t0=0.0, dt_met=21600.0)
    n = 20000
    tm, p, lon, lat = synth.make_parcels(n, t0=0.0, zmin=0.0, zmax=30.0, seed=41)
    tm = tm + np.random.default_rng(2).choice([0.0, 0.0, 0.0, 400.0], n)     # some parcels outside the time window
    nq = reference.read_ctl(["m", "Cx", "ens"], "MOLMASS 64.07")
    qi = reference.qnt
    reference.set_met(m0, m1)
    clim = reference.clim_tropo()
    rng = np.random.default_rng(3)
    q = np.zeros((nq, n))
    q[qi["m"]] = rng.uniform(0.5, 2.0, n)
    q[qi["ens"]] = rng.integers(0, 3, n)
    ctl = Ctl(nq=nq, advect=4, t_start=0.0, t_stop=1e6, dt_mod=300.0, dt_met=21600.0, qnt_m=qi["m"], qnt_Cx=qi["Cx"], qnt_ens=qi["ens"],
              nens=nens, molmass=64.07, chemgrid_nx=36, chemgrid_ny=18, chemgrid_nz=15, chemgrid_z0=0.0, chemgrid_z1=30.0, chemgrid=1)
    a = Parcels(tm, p, lon, lat, q)
    b = a.copy()
    reference.run("chem_grid", ctl, a, t=0.0)
    oracle.run("chem_grid", ctl, clim, m0, m1, b, t=0.0)
    assert _same(a, b)
    assert 0.5 < np.mean(a.q[qi["Cx"]] > 0) < 0.9


def _with_surface_gaps(m, rng, frac_ps, frac_pbl):
    from dataclasses import replace
    ps, pbl = m.ps.copy(), m.pbl.copy()
    ps[rng.uniform(size=ps.shape) < frac_ps] = np.nan
    pbl[rng.uniform(size=pbl.shape) < frac_pbl] = np.inf
    ps[1, 1] = m.ps[1, 1]
    ps[-1], pbl[-1] = ps[0], pbl[0]
    return replace(m, ps=ps, pbl=pbl)


def test_surface_gaps_bit_exact(oracle, reference):
    """non-finite nodes in ps / pbl: the nearest-neighbour rule of intpol_met_space_2d, the nearer-level rule of
    intpol_met_time_2d (src/mptrac.c:3084-3107, 3163-3169) and what MAX / MIN / the comparisons of module_diff_turb do with a
    NaN -- bit for bit (NaN positions included) in module_meteo's ps / pbl and in whole steps with turbulent diffusion"""
    from mptrac_b200 import Ctl, synth
    from oracle.oracle import Parcels
    rng = np.random.default_rng(13)
    m0, m1 = synth.make_met_pair(48, 25, 24, t0=0.0, dt_met=21600.0)
    m0, m1 = _with_surface_gaps(m0, rng, 0.10, 0.15), _with_surface_gaps(m1, rng, 0.05, 0.0)
    n = 20000
    tm, p, lon, lat = synth.make_parcels(n, t0=0.0, zmin=0.05, zmax=30.0, seed=8)
    nq = reference.read_ctl(["ps", "pbl"], "")
    reference.set_met(m0, m1)
    ctl = Ctl(nq=nq, t_start=0.0, t_stop=1e6, dt_mod=300.0, dt_met=21600.0, met_dt_out=300.0, qnt_meteo=reference.qnt_meteo)
    tmm = tm + rng.uniform(0.0, 21600.0, n)
    a = Parcels(tmm, p, lon, lat, np.zeros((nq, n)))
    b = a.copy()
    reference.run("meteo", ctl, a)
    oracle.run("meteo", ctl, reference.clim_tropo(), m0, m1, b, t=300.0)
    assert np.array_equal(a.q, b.q, equal_nan=True)
    assert 0 < np.mean(~np.isfinite(a.q)) < 0.5
    reference.read_ctl([], "")
    ctl = Ctl(advect=2, diffusion=1, turb_mesox=0.0, turb_mesoz=0.0, turb_dz_trop=0.5, turb_dz_pbl=1.0, turb_dx_pbl=30.0,
              t_start=0.0, t_stop=1e6, dt_mod=300.0, dt_met=21600.0)
    a, b = Parcels(tm, p, lon, lat), Parcels(tm, p, lon, lat)
    reference.ctr = oracle.ctr = 0
    reference.run("timestep", ctl, a, t=0.0, nsteps=3)
    oracle.run("timestep", ctl, reference.clim_tropo(), m0, m1, b, t=0.0, nsteps=3)
    for k in ("lon", "lat", "p"):
        assert np.array_equal(getattr(a, k), getattr(b, k), equal_nan=True), k
    assert np.mean(np.isfinite(a.p)) > 0.5


def test_binary_met_file_as_the_reference_reads_it(reference, tmp_path):
    """synth.write_met_bin writes the reference's uncompressed binary met format (MET_TYPE 1): the reference's own
    read_met_bin (src/mptrac.c:8887-9181, through mptrac_read_met) must recover exactly the fields that were written --
    with its bounds applied (negative temperatures become 0)"""
    from mptrac_b200 import Met, synth
    m = synth.add_meteo_fields(synth.make_met(48, 25, 24, time=360547200.0), with_gaps=False)
    m.t[3, 4, 5] = -7.0            # read_met_bin_3d clamps T to [0, 1e34]
    f = tmp_path / "met_2011_06_05_00.bin"
    synth.write_met_bin(f, m)
    reference.read_ctl([], "MET_TYPE 1")
    got = reference.read_met(f, 0, Met)
    assert got.time == m.time and got.u.shape == m.u.shape
    for k in ("lon", "lat", "p", "u", "v", "w", "ps", "pbl"):
        assert np.array_equal(getattr(got, k), getattr(m, k)), k
    want_t = m.t.copy()
    want_t[3, 4, 5] = 0.0
    assert np.array_equal(got.t, want_t)
    reference.read_ctl([], "")


def test_grid_binning_is_what_the_reference_tool_writes(oracle, tmp_path):
    """write_grid's binning (src/mptrac.c:13840-13918) has no entry point of its own in the harness: the oracle's `orc_grid_bin`
    is pinned here against the reference's own `atm2grid` tool run on a binary particle file (raw doubles in, counts and
    17-digit means / standard deviations out) -- parcels on and beside the faces of the grid included."""
    import os
    import subprocess
    from oracle.oracle import Parcels
    tool = ROOT / "oracle" / "_ref" / "bin" / "atm2grid"
    if not tool.exists():
        pytest.skip("oracle/_ref is not built here")
    rng = np.random.default_rng(21)
    nx, ny, nz, z0, z1 = 24, 12, 10, -2.0, 38.0
    n = 60000
    mag = np.concatenate([[0.0], 10.0 ** np.arange(-15.0, -2.0, 1.0)])
    eps = np.concatenate([-mag[::-1], mag])
    xf = ((-180.0 + 360.0 / nx * np.arange(nx + 1))[:, None] + eps[None, :]).ravel()
    yf = ((-90.0 + 180.0 / ny * np.arange(ny + 1))[:, None] + eps[None, :]).ravel()
    lon = np.concatenate([rng.uniform(-180, 180, n // 2), rng.choice(xf, n // 2)])
    lat = np.concatenate([rng.uniform(-90, 90, n // 2), rng.choice(yf, n // 2)])
    p = 1013.25 * np.exp(-rng.uniform(z0 - 2, z1 + 2, n) / 7.0)
    tm = np.where(rng.uniform(size=n) < 0.9, 0.0, rng.choice([-151.0, -150.0, 150.0, 151.0], n))   # the window is [t - dt/2, t + dt/2]
    q = rng.uniform(0.5, 2.0, (1, n))
    with open(tmp_path / "atm_2000_01_01_00_00_00.bin", "wb") as f:
        np.array([100, n], np.int32).tofile(f)                       # read_atm_bin: version, np, arrays, final flag (:8422-8470)
        for a in (tm, p, lon, lat, q[0]):
            a.astype(np.float64).tofile(f)
        np.array([999], np.int32).tofile(f)
    env = dict(os.environ, OMP_NUM_THREADS="2")
    r = subprocess.run([str(tool), "-", "atm_2000_01_01_00_00_00.bin", "ATM_TYPE", "1", "NQ", "1", "QNT_NAME[0]", "zeta",
                        "QNT_FORMAT[0]", "%.17g", "GRID_BASENAME", "grid", "GRID_STDDEV", "1", "DT_MOD", "300",
                        "GRID_NX", str(nx), "GRID_NY", str(ny), "GRID_NZ", str(nz), "GRID_Z0", str(z0), "GRID_Z1", str(z1),
                        "GRID_LON0", "-180", "GRID_LON1", "180", "GRID_LAT0", "-90", "GRID_LAT1", "90"],
                       cwd=tmp_path, env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    tab = np.loadtxt(tmp_path / "grid_2000_01_01_00_00_00.tab")
    assert tab.shape == (nx * ny * nz, 11)                                # rows in box order: ix, then iy, then iz (:14026-14052)
    cnt_ref, mean_ref, sig_ref = tab[:, 8].astype(np.int64), tab[:, 9], tab[:, 10]
    cnt, s, sq = oracle.grid_bin(Parcels(tm, p, lon, lat, q), nx, ny, nz, -180, 180, -90, 90, z0, z1, -150.0, 150.0)
    assert np.array_equal(cnt, cnt_ref) and 0.5 * n < cnt.sum() < n
    full = cnt > 0
    mean = s[0][full] / cnt[full]
    var = sq[0][full] / cnt[full] - mean ** 2
    assert np.array_equal(mean, mean_ref[full])                            # same sums in the same order, same division
    assert np.array_equal(np.where(var > 0, np.sqrt(np.where(var > 0, var, 0.0)), 0.0), sig_ref[full])
    assert np.all(np.isnan(mean_ref[~full]))
