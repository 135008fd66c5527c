import sys; sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
from mptrac_b200 import Ctl, Engine, synth
def case(diff, sedi, sort_dt, n=300_000):
    m0, m1 = synth.make_met_pair(72, 37, 30, t0=0.0, dt_met=21600.0)
    tm, p, lon, lat = synth.make_parcels(n, t0=0.0, zmin=0.05, zmax=45.0, seed=5)
    clim = synth.make_clim_tropo()
    q = np.stack([np.full(n, 2.0), np.full(n, 1500.0)])
    kw = dict(nq=2, advect=4, diffusion=diff, t_start=0.0, t_stop=1e6, dt_mod=300.0, dt_met=21600.0, turb_dz_trop=0.5, turb_dx_strat=20.0, sort_dt=sort_dt)
    if sedi: kw.update(qnt_rp=0, qnt_rhop=1)
    ctl = Ctl(**kw)
    engs = [Engine(n, 2), Engine(n, 2)]
    for e in engs:
        e.set_ctl(ctl); e.set_clim_tropo(*clim); e.set_met(0, m0); e.set_met(1, m1); e.set_atm(tm, p, lon, lat, q)
    a, b = engs
    a.run_timestep(0.0); b.run_timestep(0.0)
    h = b.get_atm(); hq = h["q"]
    for s in range(1, 5):
        a.run_timestep(300.0 * s); b.run_timestep_host(300.0 * s, h["time"], h["p"], h["lon"], h["lat"], hq)
        ref = a.get_atm()
        msg = []
        for k in ("time", "p", "lon", "lat"):
            bad = np.nonzero(ref[k] != h[k])[0]
            msg.append(f"{k}:{bad.size}" + (f"[{bad[0]}..{bad[-1]}] max|d|={np.max(np.abs(ref[k][bad]-h[k][bad])):.3e}" if bad.size else ""))
        print(f"diff={diff} sedi={sedi} sort_dt={sort_dt} step {s}: " + " ".join(msg), flush=True)
    for e in engs: e.close()
for diff, sedi, sort_dt in [(0, 0, -999.0), (0, 1, -999.0), (1, 0, -999.0), (1, 1, -999.0), (1, 1, 900.0)]:
    case(diff, sedi, sort_dt)
