"""The CPU oracle against the committed golden fixtures (tests/golden/*.npz, generated FROM THE UNMODIFIED
REFERENCE by tests/golden/make_golden.py): the reference's own known answers and its binary results.
Runs anywhere (no GPU, no reference tree)."""
import numpy as np
import pytest

from conftest import GOLDEN, abserr, clim_from_npz, met_from_npz, relerr


def _golden(name):
    f = GOLDEN / name
    if not f.exists():
        pytest.skip(f"{f} missing")
    return np.load(f)


def test_sedi_known_answers(oracle):
    """tests/tools_test/data.ref/sedi.tab: 144 values of the settling velocity"""
    z = _golden("sedi_kat.npz")
    v = np.array([oracle.sedi(*a) for a in zip(z["p"], z["T"], z["rp"], z["rhop"])])
    assert v.size == 144
    assert relerr(v, z["vs_text"]) < 5e-6          # the shipped text has 6 significant digits
    assert np.array_equal(v, z["vs_exact"])        # bit-exact against the reference's sedi() run here


def _prefix_steps(oracle, ctl, clim, m0, m1, a, t, n_total):
    """one model step for a PREFIX of n_total parcels: module_rng would draw 3*n_total+1 counters per call"""
    base = oracle.ctr
    oracle.run("timesteps", ctl, clim, m0, m1, a, t=t)
    oracle.run("position", ctl, clim, m0, m1, a)
    oracle.run("advect", ctl, clim, m0, m1, a)
    oracle.ctr = base
    oracle.run("diff_turb", ctl, clim, m0, m1, a)
    oracle.ctr = base + 3 * n_total + 1
    oracle.run("diff_meso", ctl, clim, m0, m1, a)
    oracle.ctr = base + 2 * (3 * n_total + 1)
    oracle.run("position", ctl, clim, m0, m1, a)


def test_dt_test_goldens(oracle):
    """the reference's tests/dt_test: bit-exact against its binary state, and against its shipped %g text goldens"""
    from mptrac_b200 import Ctl
    from oracle.oracle import Parcels
    z = _golden("dt_test.npz")
    m0, m1, clim = met_from_npz(z, "m0"), met_from_npz(z, "m1"), clim_from_npz(z)
    t0, n_total = float(z["t_start"]), int(z["np_total"])
    # quantities t, u, v, w: module_meteo after every step (the test's control file leaves MET_DT_OUT at 0.1)
    ctl = Ctl(nq=4, advect=2, diffusion=1, dt_mod=10.0, dt_met=86400.0, t_start=t0, t_stop=t0 + 60.0, met_dt_out=0.1,
              qnt_meteo=dict(t=0, u=1, v=2, w=3))
    a = Parcels(z["time"], z["p"], z["lon"], z["lat"], np.zeros((4, z["time"].size)))
    oracle.ctr = 0
    for s in range(7):
        _prefix_steps(oracle, ctl, clim, m0, m1, a, t0 + 10.0 * s, n_total)
        oracle.run("meteo", ctl, clim, m0, m1, a)
        rb, txt = z["ref_binary"][s], z["ref_shipped_text"][s]
        assert np.array_equal(np.stack([a.time, a.p, a.lon, a.lat]), rb), f"step {s}"
        assert np.array_equal(a.q, z["ref_binary_q"][s]), f"step {s}: quantities"
        assert relerr(a.q.T, z["ref_shipped_q"][s]) < 1e-5            # columns 5-8 of the shipped goldens
        zkm = 7.0 * np.log(1013.25 / a.p)
        assert abserr(a.time, txt[:, 0]) < 0.006
        assert relerr(zkm, txt[:, 1]) < 1e-5 and relerr(a.lon, txt[:, 2]) < 1e-5 and relerr(a.lat, txt[:, 3]) < 1e-5
    assert abserr(a.lon, z["lon"]) > 1e-3


def test_coord_test_goldens(oracle):
    """the reference's tests/coord_test (Cartesian met, met-level swap after one hour)"""
    from mptrac_b200 import Ctl
    from oracle.oracle import Parcels
    z = _golden("coord_test.npz")
    mets = [met_from_npz(z, f"m{i}") for i in range(3)]
    clim = clim_from_npz(z)
    t0 = float(z["t_start"])
    ctl = Ctl(nq=0, advect=2, diffusion=1, dt_mod=600.0, dt_met=3600.0, t_start=t0, t_stop=t0 + 7200.0, met_coord_type=1,
              met_utm_ref_lat=48.1507476)
    a = Parcels(z["time"], z["p"], z["lon"], z["lat"])
    oracle.ctr = 0
    for s in range(13):
        lvl = 0 if s <= 6 else 1
        oracle.run("timestep", ctl, clim, mets[lvl], mets[lvl + 1], a, t=t0 + 600.0 * s)
        assert np.array_equal(np.stack([a.time, a.p, a.lon, a.lat]), z["ref_binary"][s]), f"step {s}"
        txt = z["ref_shipped_text"][s]
        assert abserr(a.lon, txt[:, 2]) < 0.02 and abserr(a.lat, txt[:, 3]) < 0.02


def test_synth_full_goldens(oracle):
    from mptrac_b200 import Ctl, synth
    from oracle.oracle import Parcels
    z = _golden("synth_full.npz")
    m0, m1 = synth.make_met_pair(24, 13, 20, t0=0.0, dt_met=21600.0, lat_descending=True)
    clim = clim_from_npz(z)
    ctl = Ctl(nq=3, qnt_rp=0, qnt_rhop=1, advect=4, diffusion=1, t_start=0.0, t_stop=86400.0, dt_mod=600.0, dt_met=21600.0,
              turb_dz_trop=0.5, turb_dz_pbl=1.0, turb_dx_strat=20.0, turb_pbl_trans=0.2, mixing_trop=0.3, mixing_strat=0.1,
              mixing_dt=1200.0, mix_qnt=[2], mixing_nx=18, mixing_ny=9, mixing_nz=12)
    a = Parcels(z["time"], z["p"], z["lon"], z["lat"], z["q"])
    oracle.ctr = 0
    for s in range(8):
        oracle.run("timestep", ctl, clim, m0, m1, a, t=600.0 * s)
        assert np.array_equal(np.stack([a.time, a.p, a.lon, a.lat]), z["ref_binary"][s]), f"step {s}"
        assert np.array_equal(a.q, z["ref_q"][s]) and np.array_equal(a.uvwp, z["ref_uvwp"][s])
    assert oracle.ctr == int(z["ref_ctr"])
