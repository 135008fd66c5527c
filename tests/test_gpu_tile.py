"""The TMA-tile form of the step kernel (MPTRAC_B200_STEP=tile: a window of the met grid staged in shared memory by one bulk
tensor copy per block, engine.cu tile_step_kernel) against the default kernel.  Both run the same per-parcel code -- only
the source of the cube's nodes differs -- so the strict builds (-fmad=false) must agree bit for bit; in the production
build the compiler contracts multiply-add pairs differently in the two kernels, so they agree to the last FMA."""
import numpy as np
import pytest

from test_gpu_parity import _case, _setup

pytestmark = pytest.mark.gpu


def _run(monkeypatch, form, n, ctl, clim, m0, m1, tm, p, lon, lat, q=None, nsteps=6, strict=False, window=None):
    from mptrac_b200 import Engine
    monkeypatch.setenv("MPTRAC_B200_STEP", form)
    if window:
        monkeypatch.setenv("MPTRAC_B200_TILE", window)
    else:
        monkeypatch.delenv("MPTRAC_B200_TILE", raising=False)
    with Engine(n, nq=ctl.nq, device=0, strict=strict) as eng:
        _setup(eng, ctl, clim, m0, m1, tm, p, lon, lat, q)
        for s in range(nsteps):
            eng.run_timestep(s * ctl.dt_mod)
        out = eng.get_atm()
        out["uvwp"] = eng.get_uvwp()
    return out


def _same(a, b, exact=True):
    for k in ("time", "q"):
        assert np.array_equal(a[k], b[k], equal_nan=True), k
    if exact:
        for k in ("lon", "lat", "p", "uvwp"):
            assert np.array_equal(a[k], b[k], equal_nan=True), k
    else:
        dlon = (a["lon"] - b["lon"] + 180.0) % 360.0 - 180.0
        assert np.max(np.abs(dlon) * np.maximum(np.cos(np.deg2rad(b["lat"])), 1e-6)) < 1e-10
        assert np.max(np.abs(a["lat"] - b["lat"])) < 1e-10 and np.max(np.abs(a["p"] - b["p"]) / np.abs(b["p"])) < 1e-10
        assert np.max(np.abs(a["uvwp"] - b["uvwp"])) < 1e-5


@pytest.mark.parametrize("strict", [False, True])
@pytest.mark.parametrize("sort_dt,window", [(300.0, None), (-999.0, None), (600.0, "2,3,8"), (300.0, "4,6,24")])
def test_tile_equals_default_kernel(monkeypatch, strict, sort_dt, window):
    """RK4 + turbulent + mesoscale diffusion + sedimentation; sorted every step (parcels inside the window), never sorted
    (nearly every cell outside it: global path), a tiny window, a window taller than the grid; ragged size, late starters"""
    from mptrac_b200 import Ctl
    m0, m1, tm, p, lon, lat, clim = _case(n=40_037, grid=(72, 37, 24))
    n = tm.size
    tm = np.where(np.arange(n) % 11 == 0, 900.0, 0.0)
    q = np.stack([np.full(n, 2.0), np.full(n, 1500.0), np.arange(n, dtype=np.float64)])
    ctl = Ctl(nq=3, qnt_rp=0, qnt_rhop=1, advect=4, diffusion=1, sort_dt=sort_dt, t_start=0.0, t_stop=1e6, dt_mod=300.0, dt_met=21600.0,
              turb_dz_trop=0.5, turb_dz_pbl=1.0, turb_dx_strat=20.0)
    a = _run(monkeypatch, "tile", n, ctl, clim, m0, m1, tm, p, lon, lat, q, strict=strict, window=window)
    b = _run(monkeypatch, "classic", n, ctl, clim, m0, m1, tm, p, lon, lat, q, strict=strict)
    assert np.max(np.abs(a["lat"] - lat[a["q"][2].astype(np.int64)])) > 1e-3
    _same(a, b, exact=strict)


@pytest.mark.parametrize("advect", [1, 2])
def test_tile_other_integrators_and_dense_cloud(monkeypatch, advect):
    """Euler / midpoint, advection only, a dense regional cloud (many parcels per cell: every block fits its window)"""
    from mptrac_b200 import Ctl
    m0, m1, tm, p, lon, lat, clim = _case(n=60_000, grid=(72, 37, 24))
    n = tm.size
    rng = np.random.default_rng(9)
    lon, lat = rng.uniform(10.0, 40.0, n), rng.uniform(-20.0, 10.0, n)
    ctl = Ctl(advect=advect, sort_dt=300.0, t_start=0.0, t_stop=1e6, dt_mod=300.0, dt_met=21600.0)
    a = _run(monkeypatch, "tile", n, ctl, clim, m0, m1, tm, p, lon, lat, strict=True)
    b = _run(monkeypatch, "classic", n, ctl, clim, m0, m1, tm, p, lon, lat, strict=True)
    _same(a, b)
