"""GPU tests that have NOT run on a B200 yet (the round's GPU budget was spent when the code they cover was written).
The device source they exercise is bit-exact against the oracle when compiled for the host (tests/test_hostemu.py) and
the oracle is bit-exact against the reference (tests/test_oracle_vs_reference.py); what is missing is the run on the
device.  They carry their own marker so that neither `-m gpu` nor `-m "not gpu"` (on a CPU box: skipped) depends on
them: run `pytest -m gpu_pending` on a GPU box, then move them into tests/test_gpu_parity.py / test_shim_trac.py."""
import numpy as np
import pytest

from conftest import abserr, has_gpu, relerr

pytestmark = [pytest.mark.gpu_pending, pytest.mark.skipif(not has_gpu(), reason="needs a CUDA device")]


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
def test_trac_trac_test_pl_with_device_meteo_fields(tmp_path, monkeypatch):
    """tests/trac_test (pressure-level run) through the shim with MPTRAC_B200_DEVICE_METEO_FIELDS=1: zg, pv and pt of its
    control file come from the device's module_meteo instead of the reference's CPU code"""
    import test_shim_trac as T
    monkeypatch.setenv("MPTRAC_B200_DEVICE_METEO_FIELDS", "1")
    T.test_trac_trac_test_through_the_shim(tmp_path, "pl")
