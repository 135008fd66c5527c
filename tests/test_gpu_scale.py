"""GPU parity at the grids and densities of BASELINE configs[2], [3] and [4] (the configurations bench.py measures as c3,
c4 / c4g and c5), through the C ABI against the oracle on the same seeded inputs -- so that every bench line has a
parity case of its own shape beside it:

  configs[2]  721 x 361 x 137 grid, RK4 + mesoscale diffusion + sedimentation     (1 M of its 10 M parcels)
  configs[3]  361 x 181 x 60 grid at the density of 100 M parcels (6 per cell), RK4 + turbulent + mesoscale
              diffusion, cell sort, gridded output on 360 x 180 x 1 boxes          (a regional 2 M-parcel cloud)
  configs[4]  361 x 181 x 60 grid, RK4, cell sort + inter-parcel mixing on the reference's default 360 x 180 x 90
              boxes every step                                                     (one GPU's 1.25 M parcels)

Tolerances as in test_gpu_parity.py: positions 1e-7 with diffusion on (CUDA's sinf / cosf), 1e-11 deg / 1e-12 without;
box counts exact; box means 1e-12 (sums run in a different order)."""
import numpy as np
import pytest

from conftest import abserr, relerr
from test_gpu_parity import TOL_P_REL, TOL_P_REL_DIFF, TOL_POS_DEG, TOL_POS_DEG_DIFF, _compare, _engine, _report, _setup

pytestmark = pytest.mark.gpu


def test_c3_grid_rk4_meso_sedi_vs_oracle(oracle):
    """BASELINE configs[2] at its own grid: 0.5 deg x 137 levels (1.14 GB of packed met), RK4 + mesoscale diffusion +
    sedimentation, the sort cadence of the bench workload; 1 M parcels, 3 model steps, every parcel against the oracle"""
    from mptrac_b200 import Ctl, synth
    from oracle.oracle import Parcels
    m0, m1 = synth.make_met_pair(720, 361, 137, t0=0.0, dt_met=21600.0)
    n = 1_000_000
    tm, p, lon, lat = synth.make_parcels(n, t0=0.0, seed=321)
    q = np.stack([np.full(n, 1.0), np.full(n, 1500.0)])
    clim = synth.make_clim_tropo()
    ctl = Ctl(nq=2, qnt_rp=0, qnt_rhop=1, advect=4, diffusion=1, turb_dx_pbl=0, turb_dx_trop=0, turb_dz_strat=0,
              t_start=0.0, t_stop=1e9, dt_mod=300.0, dt_met=21600.0)
    with _engine(n, 2) as eng:
        _setup(eng, ctl, clim, m0, m1, tm, p, lon, lat, q)
        for s in range(4):
            eng.run_timestep(300.0 * s)
        out = eng.get_atm()
        uv = eng.get_uvwp()
        ctr = eng.rng_ctr
    ref = Parcels(tm, p, lon, lat, q)
    oracle.ctr = 0
    oracle.run("timestep", ctl, clim, m0, m1, ref, t=0.0, nsteps=4)
    assert ctr == oracle.ctr
    assert abserr(ref.lat, lat) > 1e-3
    _compare("c3_grid_rk4_meso_sedi", out, ref, TOL_POS_DEG_DIFF, TOL_P_REL_DIFF)
    assert np.mean(np.abs(uv - ref.uvwp) > 1e-4 * (1 + np.abs(ref.uvwp))) < 1e-4


def _dense_cloud(n, seed):
    """parcels at the density of configs[3] (100 M on the 1 deg x 60 level grid = 6 per cell where the cloud is): a regional
    box of 70 x 56 columns x the levels between 1 and 30 km"""
    rng = np.random.default_rng(seed)
    lon = rng.uniform(-35.0, 35.0, n)
    lat = rng.uniform(20.0, 76.0, n)
    z = rng.uniform(1.0, 30.0, n)
    return np.zeros(n), 1013.25 * np.exp(-z / 7.0), lon, lat


def test_c4_dense_share_turb_meso_sorted_vs_oracle(oracle):
    """the dense workload of configs[3]: > 6 parcels per met cell, RK4 + turbulent + mesoscale diffusion.  Unsorted run
    against the oracle parcel by parcel; then the same with the cell sort every step -- a parcel carries its identity as a
    quantity -- whose permutation must match the oracle's sorted run exactly (slot-attached random numbers and uvwp
    included)"""
    from mptrac_b200 import Ctl, synth
    from oracle.oracle import Parcels
    m0, m1 = synth.make_met_pair(360, 181, 60, t0=0.0, dt_met=21600.0)
    n = 2_000_000
    tm, p, lon, lat = _dense_cloud(n, 77)
    ident = np.arange(n, dtype=np.float64)[None, :]
    clim = synth.make_clim_tropo()
    for sort_dt in (-999.0, 300.0):
        ctl = Ctl(nq=1, advect=4, diffusion=1, sort_dt=sort_dt, t_start=0.0, t_stop=1e9, dt_mod=300.0, dt_met=21600.0)
        with _engine(n, 1) as eng:
            _setup(eng, ctl, clim, m0, m1, tm, p, lon, lat, ident)
            for s in range(3):
                eng.run_timestep(300.0 * s)
            out = eng.get_atm()
            ctr = eng.rng_ctr
        ref = Parcels(tm, p, lon, lat, ident)
        oracle.ctr = 0
        oracle.run("timestep", ctl, clim, m0, m1, ref, t=0.0, nsteps=3)
        assert ctr == oracle.ctr
        if sort_dt > 0:
            assert np.array_equal(out["q"][0], ref.q[0]), "the cell sort's permutation differs from the oracle's"
            assert np.any(np.diff(out["q"][0]) < 0)
        _compare(f"c4_dense_turb_meso[sort_dt={sort_dt:g}]", out, ref, TOL_POS_DEG_DIFF, TOL_P_REL_DIFF)


def test_c4g_gridded_output_default_grid_vs_oracle(oracle):
    """configs[3]'s per-step gridded output on the reference's default output grid (360 x 180 x 1 boxes, ~30 parcels per
    occupied box): counts exact, sums and sums of squares to rounding (the device sums runs of equal boxes in a warp)"""
    from mptrac_b200 import Ctl, synth
    from oracle.oracle import Parcels
    m0, m1 = synth.make_met_pair(360, 181, 60, t0=0.0, dt_met=21600.0)
    n = 2_000_000
    tm, p, lon, lat = _dense_cloud(n, 78)
    q = np.random.default_rng(5).uniform(0.5, 2.0, (1, n))
    clim = synth.make_clim_tropo()
    ctl = Ctl(nq=1, advect=4, diffusion=0, sort_dt=300.0, t_start=0.0, t_stop=1e9, dt_mod=300.0, dt_met=21600.0)
    grid = (360, 180, 1, -180.0, 180.0, -90.0, 90.0, -5.0, 85.0)
    with _engine(n, 1) as eng:
        _setup(eng, ctl, clim, m0, m1, tm, p, lon, lat, q)
        for s in range(2):
            eng.run_timestep(300.0 * s)
        eng.grid_accumulate(*grid, 300.0 - 150.0, 300.0 + 150.0)
        cnt, s1, s2 = eng.grid_fetch()
        out = eng.get_atm()
    c0, a0, b0 = oracle.grid_bin(Parcels(out["time"], out["p"], out["lon"], out["lat"], out["q"]), *grid, 150.0, 450.0)
    assert np.array_equal(cnt, c0) and cnt.sum() == n and cnt.max() > 100
    occ = c0 > 0
    e1, e2 = relerr(s1[0][occ], a0[0][occ]), relerr(s2[0][occ], b0[0][occ])
    _report("c4g_grid_output", sum_rel=e1, sq_rel=e2)
    assert e1 < 1e-12 and e2 < 1e-12


def test_c5_mixing_default_grid_vs_oracle(oracle):
    """configs[4]: one GPU's 1.25 M parcels, RK4, cell sort and inter-parcel mixing on the reference's default mixing grid
    (360 x 180 x 90 = 5.8 M boxes, src/mptrac.c:5273-5335) every step, two mixed quantities; the permutation and the box
    counts are exact, the relaxed quantities agree to 1e-12 (means of sums taken in a different order)"""
    from mptrac_b200 import Ctl, synth
    from oracle.oracle import Parcels
    m0, m1 = synth.make_met_pair(360, 181, 60, t0=0.0, dt_met=21600.0)
    n = 1_250_000
    tm, p, lon, lat = synth.make_parcels(n, t0=0.0, seed=123)
    rng = np.random.default_rng(7)
    q = np.stack([rng.uniform(0.0, 1.0, n), rng.uniform(10.0, 20.0, n), np.arange(n, dtype=np.float64)])
    clim = synth.make_clim_tropo()
    ctl = Ctl(nq=3, advect=4, diffusion=0, sort_dt=300.0, mixing_trop=0.3, mixing_strat=0.05, mixing_dt=300.0, mix_qnt=[0, 1],
              t_start=0.0, t_stop=1e9, dt_mod=300.0, dt_met=21600.0)
    with _engine(n, 3) as eng:
        _setup(eng, ctl, clim, m0, m1, tm, p, lon, lat, q)
        for s in range(3):
            eng.run_timestep(300.0 * s)
        out = eng.get_atm()
    ref = Parcels(tm, p, lon, lat, q)
    oracle.run("timestep", ctl, clim, m0, m1, ref, t=0.0, nsteps=3)
    assert np.array_equal(out["q"][2], ref.q[2]), "permutation"
    _compare("c5_mixing_default_grid", out, ref, TOL_POS_DEG, TOL_P_REL)
    moved = np.abs(ref.q[0] - q[0][ref.q[2].astype(np.int64)])
    assert np.mean(moved > 1e-6) > 0.01, "mixing did nothing: no box held two parcels"
    e0, e1 = relerr(out["q"][0], ref.q[0]), relerr(out["q"][1], ref.q[1])
    _report("c5_mixing_default_grid_q", q0_rel=e0, q1_rel=e1)
    assert e0 < 1e-12 and e1 < 1e-12


def test_nan_gaps_in_surface_fields_vs_oracle(oracle):
    """ps / pbl with non-finite nodes (missing data): intpol_met_space_2d falls back to the nearest neighbour and
    intpol_met_time_2d to the nearer time level (src/mptrac.c:3084-3107, 3163-3169).  Exercised where the path reads
    them: module_diff_turb (pbl, ps at the parcel) and module_meteo (quantities ps, pbl)"""
    from dataclasses import replace
    from mptrac_b200 import Ctl, synth
    from oracle.oracle import Parcels
    m0, m1 = synth.make_met_pair(48, 25, 24, t0=0.0, dt_met=21600.0)
    rng = np.random.default_rng(13)

    def gaps(m, frac_ps, frac_pbl):
        ps, pbl = m.ps.copy(), m.pbl.copy()
        ps[rng.uniform(size=ps.shape) < frac_ps] = np.nan
        pbl[rng.uniform(size=pbl.shape) < frac_pbl] = np.inf
        ps[1, 1] = m.ps[1, 1]          # module_position's node (src/mptrac.c:5483) stays finite
        ps[-1], pbl[-1] = ps[0], pbl[0]
        return replace(m, ps=ps, pbl=pbl)
    m0, m1 = gaps(m0, 0.10, 0.15), gaps(m1, 0.05, 0.0)
    n = 20000
    tm, p, lon, lat = synth.make_parcels(n, t0=0.0, zmin=0.05, zmax=30.0, seed=8)
    clim = synth.make_clim_tropo()
    # module_meteo alone: ps and pbl of every parcel, NaN / inf where the reference yields them
    ctl = Ctl(nq=2, t_start=0.0, t_stop=1e6, dt_mod=300.0, dt_met=21600.0, met_dt_out=300.0, qnt_meteo={"ps": 0, "pbl": 1})
    tmm = tm + rng.uniform(0.0, 21600.0, n)
    with _engine(n, 2) as eng:
        _setup(eng, ctl, clim, m0, m1, tmm, p, lon, lat, np.zeros((2, n)))
        eng.module_meteo()
        out = eng.get_atm()
    ref = Parcels(tmm, p, lon, lat, np.zeros((2, n)))
    oracle.run("meteo", ctl, clim, m0, m1, ref, t=300.0)
    for i in range(2):
        a, b = out["q"][i], ref.q[i]
        assert np.array_equal(np.isfinite(a), np.isfinite(b)) and np.array_equal(np.isnan(a), np.isnan(b))
        ok = np.isfinite(b)
        assert 0.5 < ok.mean() < 1.0 or i == 0
        assert relerr(a[ok], b[ok]) < 1e-12
    # the turbulent diffusion reads both at the parcel position (parcels with non-finite pbl / ps follow the reference's
    # comparisons with NaN: weights 0 or 1)
    ctl = Ctl(advect=2, diffusion=1, turb_mesox=0.0, turb_mesoz=0.0, turb_dz_trop=0.5, turb_dz_pbl=1.0, turb_dx_pbl=30.0,
              t_start=0.0, t_stop=1e6, dt_mod=300.0, dt_met=21600.0)
    with _engine(n) as eng:
        _setup(eng, ctl, clim, m0, m1, tm, p, lon, lat)
        for s in range(3):
            eng.run_timestep(300.0 * s)
        out = eng.get_atm()
    ref = Parcels(tm, p, lon, lat)
    oracle.ctr = 0
    oracle.run("timestep", ctl, clim, m0, m1, ref, t=0.0, nsteps=3)
    fin = np.isfinite(ref.p) & np.isfinite(ref.lon) & np.isfinite(ref.lat)
    assert np.array_equal(fin, np.isfinite(out["p"]) & np.isfinite(out["lon"]) & np.isfinite(out["lat"]))
    assert fin.mean() > 0.5
    sub = {k: out[k][fin] for k in ("time", "lon", "lat", "p")}
    refs = Parcels(ref.time[fin], ref.p[fin], ref.lon[fin], ref.lat[fin])
    _compare("nan_gaps_diff_turb", sub, refs, TOL_POS_DEG_DIFF, TOL_P_REL_DIFF)
