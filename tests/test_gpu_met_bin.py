"""mpb_set_met_bin: a met level straight from one of the reference's uncompressed binary met files (MET_TYPE 1) into the
device layout.  The file format is pinned against the reference's own reader on the CPU
(tests/test_oracle_vs_reference.py::test_binary_met_file_as_the_reference_reads_it); here the device state it produces must
equal, bit for bit, what mpb_set_met produces from the same fields -- transport, diffusion, sedimentation and all 53
module_meteo quantities included."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_met_from_binary_file_equals_met_from_arrays(tmp_path):
    from dataclasses import replace
    from mptrac_b200 import Ctl, Engine, MpbError, synth
    from mptrac_b200.host import METEO_QNT
    m0, m1 = synth.make_met_pair(48, 25, 24, t0=360547200.0, dt_met=21600.0)
    m0, m1 = synth.add_meteo_fields(m0), synth.add_meteo_fields(m1)
    m0.t[2, 3, 4] = -5.0                              # the reader clamps temperatures to [0, 1e34] like read_met_bin_3d
    files = [tmp_path / "met0.bin", tmp_path / "met1.bin"]
    synth.write_met_bin(files[0], m0)
    synth.write_met_bin(files[1], m1)
    t_clamped = m0.t.copy()
    t_clamped[2, 3, 4] = 0.0
    m0 = replace(m0, t=t_clamped)
    n = 20000
    tm, p, lon, lat = synth.make_parcels(n, t0=360547200.0, zmin=0.05, zmax=40.0, seed=17)
    qm = {name: 2 + i for i, name in enumerate(METEO_QNT)}
    nq = 2 + len(qm)
    q = np.zeros((nq, n))
    q[0], q[1] = 2.0, 1500.0
    ctl = Ctl(nq=nq, qnt_rp=0, qnt_rhop=1, advect=4, diffusion=1, t_start=360547200.0, t_stop=360547200.0 + 86400.0, dt_mod=300.0,
              dt_met=21600.0, turb_dz_trop=0.5, turb_dx_strat=20.0, met_dt_out=300.0, qnt_meteo=qm)
    clim = synth.make_clim_tropo()
    outs = []
    for from_file in (False, True):
        with Engine(n, nq=nq, device=0) as eng:
            eng.set_ctl(ctl)
            eng.set_clim_tropo(*clim)
            if from_file:
                eng.set_met_bin(0, files[0], all_fields=True)
                eng.set_met_bin(1, files[1], all_fields=True)
            else:
                eng.set_met(0, m0)
                eng.set_met(1, m1)
            eng.set_atm(tm, p, lon, lat, q)
            for s in range(4):
                eng.run_timestep(360547200.0 + 300.0 * s)
            o = eng.get_atm()
            o["uvwp"] = eng.get_uvwp()
            outs.append(o)
    a, b = outs
    assert np.max(np.abs(a["lat"] - lat)) > 1e-3
    for k in ("time", "lon", "lat", "p", "uvwp"):
        assert np.array_equal(a[k], b[k]), k
    assert np.array_equal(a["q"], b["q"], equal_nan=True)
    assert np.all(np.isfinite(a["q"][qm["t"]])) and np.max(a["q"][qm["zg"]]) > 1.0
    # without the further fields a quantity that needs one fails loudly, and a truncated file is an error, not garbage
    with Engine(n, nq=nq, device=0) as eng:
        eng.set_ctl(ctl)
        eng.set_clim_tropo(*clim)
        eng.set_met_bin(0, files[0])
        eng.set_met_bin(1, files[1])
        eng.set_atm(tm, p, lon, lat, q)
        with pytest.raises(MpbError, match="met field"):
            eng.module_meteo()
        short = tmp_path / "short.bin"
        short.write_bytes(files[0].read_bytes()[:-100000])
        with pytest.raises(MpbError, match="ends early"):
            eng.set_met_bin(0, short)
