"""The lane-per-coordinate step kernel (csrc/quad.cuh, an opt-in variant: MPTRAC_B200_STEP=quad) against the default
one-thread-per-parcel kernel: the two state the same arithmetic operation for operation.  The strict builds (-fmad=false)
must agree BIT FOR BIT; the production builds contract different multiply-add pairs into FMAs, so they agree like either
agrees with the oracle (1e-11 deg, 1e-12 relative in p) -- every integrator, forward and backward, both latitude orders,
poles and date line included, ragged sizes, host-resident stepping."""
import numpy as np
import pytest

from test_gpu_parity import _case, _setup

pytestmark = pytest.mark.gpu


def _run(monkeypatch, form, n, ctl, clim, m0, m1, tm, p, lon, lat, q=None, nsteps=6, strict=False, t0=0.0, direction=1, env=None):
    from mptrac_b200 import Engine
    monkeypatch.setenv("MPTRAC_B200_STEP", form)
    for k, v in (env or {}).items():
        monkeypatch.setenv(k, v)
    with Engine(n, nq=ctl.nq, device=0, strict=strict) as eng:
        _setup(eng, ctl, clim, m0, m1, tm, p, lon, lat, q)
        for s in range(nsteps):
            eng.run_timestep(t0 + s * direction * ctl.dt_mod)
        out = eng.get_atm()
        out["uvwp"] = eng.get_uvwp()
        out["dt"] = eng.get_dt()
        out["launches"] = eng.launch_count
    return out


def _same(a, b, exact=True, dt=True):
    for k in ("time", "uvwp") + (("dt",) if dt else ()):
        assert np.array_equal(a[k], b[k]), k
    if a["q"].size:
        assert np.array_equal(a["q"], b["q"])
    if exact:
        for k in ("lon", "lat", "p"):
            assert np.array_equal(a[k], b[k]), k
    else:
        dlon = (a["lon"] - b["lon"] + 180.0) % 360.0 - 180.0
        assert np.max(np.abs(dlon) * np.maximum(np.cos(np.deg2rad(b["lat"])), 1e-6)) < 1e-11
        # (the parcels placed above the grid top are reflected there, p -> ptop^2 / p, which amplifies a last-bit difference)
        assert np.max(np.abs(a["lat"] - b["lat"])) < 1e-11 and np.max(np.abs(a["p"] - b["p"]) / b["p"]) < 1e-10


@pytest.mark.parametrize("advect", [1, 2, 4])
@pytest.mark.parametrize("lat_desc", [False, True])
@pytest.mark.parametrize("direction", [1, -1])
@pytest.mark.parametrize("strict", [False, True])
def test_quad_equals_classic(monkeypatch, advect, lat_desc, direction, strict):
    from mptrac_b200 import Ctl
    m0, m1, tm, p, lon, lat, clim = _case(n=20011, lat_desc=lat_desc, zmax=50.0)
    n = tm.size
    rng = np.random.default_rng(3)
    # parcels on the poles, on the date line, beyond the grid top and below the surface, and some that start later
    lat[:50] = rng.choice([-90.0, 90.0, 89.9995, -89.9995], 50)
    lon[50:100] = rng.choice([-180.0, 179.999999, 0.0, 359.5], 50)
    p[100:150] = rng.choice([0.005, 1200.0, 1013.25], 50)
    t_start = 0.0 if direction == 1 else 21600.0
    tm = np.where(np.arange(n) % 7 == 0, t_start + direction * 600.0, t_start)
    ctl = Ctl(advect=advect, direction=direction, t_start=t_start, t_stop=t_start + direction * 86400.0, dt_mod=300.0, dt_met=21600.0)
    outs = [_run(monkeypatch, form, n, ctl, clim, m0, m1, tm, p, lon, lat, strict=strict, t0=t_start, direction=direction)
            for form in ("quad", "classic")]
    assert np.max(np.abs(outs[0]["lat"] - lat)) > 1e-3
    _same(*outs, exact=strict)


def test_quad_with_sort_and_quantities(monkeypatch):
    """cell sort in between (the kernel then reads dt from memory), quantities ride along"""
    from mptrac_b200 import Ctl
    m0, m1, tm, p, lon, lat, clim = _case(n=50000, grid=(72, 37, 30))
    n = tm.size
    q = np.stack([np.arange(n, dtype=np.float64), np.full(n, 3.0)])
    ctl = Ctl(nq=2, advect=4, sort_dt=600.0, t_start=0.0, t_stop=1e6, dt_mod=300.0, dt_met=21600.0)
    outs = [_run(monkeypatch, form, n, ctl, clim, m0, m1, tm, p, lon, lat, q, nsteps=7, strict=True) for form in ("quad", "classic")]
    _same(*outs)
    assert np.any(np.diff(outs[0]["q"][0]) < 0)


def test_quad_split_with_diffusion_and_sedimentation(monkeypatch):
    """MPTRAC_B200_QUAD_SPLIT=1: advection in the lane-per-coordinate kernel, diffusion and sedimentation in a second launch"""
    from mptrac_b200 import Ctl
    m0, m1, tm, p, lon, lat, clim = _case(n=30000)
    n = tm.size
    q = np.stack([np.full(n, 2.0), np.full(n, 1500.0)])
    ctl = Ctl(nq=2, qnt_rp=0, qnt_rhop=1, advect=4, diffusion=1, t_start=0.0, t_stop=1e6, dt_mod=300.0, dt_met=21600.0,
              turb_dz_trop=0.5, turb_dz_pbl=1.0, turb_dx_strat=20.0)
    a = _run(monkeypatch, "quad", n, ctl, clim, m0, m1, tm, p, lon, lat, q, strict=True, env={"MPTRAC_B200_QUAD_SPLIT": "1"})
    b = _run(monkeypatch, "classic", n, ctl, clim, m0, m1, tm, p, lon, lat, q, strict=True, env={"MPTRAC_B200_QUAD_SPLIT": "0"})
    _same(a, b, dt=False)      # (the split form leaves cache_t::dt in memory for its second launch; the fused step does not)
    assert a["launches"] > b["launches"]


def test_quad_host_resident_step(monkeypatch):
    """zero-copy stepping of pinned host arrays goes through the same kernel"""
    import torch
    from mptrac_b200 import Ctl, Engine
    m0, m1, tm, p, lon, lat, clim = _case(n=100_003, grid=(72, 37, 30))
    n = tm.size
    ctl = Ctl(advect=4, t_start=0.0, t_stop=1e6, dt_mod=300.0, dt_met=21600.0)
    res = []
    for form in ("quad", "classic"):
        monkeypatch.setenv("MPTRAC_B200_STEP", form)
        blk = torch.from_numpy(np.stack([tm, p, lon, lat])).pin_memory()
        h = {k: blk[i].numpy() for i, k in enumerate(("time", "p", "lon", "lat"))}
        with Engine(n, nq=0, device=0, strict=True) as eng:
            _setup(eng, ctl, clim, m0, m1, tm, p, lon, lat)
            for s in range(4):
                eng.run_timestep_host(300.0 * s, h["time"], h["p"], h["lon"], h["lat"])
            dev = eng.get_atm()
        for k in h:
            assert np.array_equal(h[k], dev[k]), k
        res.append({k: v.copy() for k, v in h.items()})
    for k in res[0]:
        assert np.array_equal(res[0][k], res[1][k]), k
