"""GPU diagnosis of module_diff_pbl vs the oracle: error distribution per step for the production and the strict library."""
import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
from mptrac_b200 import Ctl, Engine, synth
from oracle.oracle import Oracle, Parcels

m0, m1 = synth.make_met_pair(48, 25, 24, t0=0.0, dt_met=21600.0)
m0, m1 = synth.add_meteo_fields(m0, with_gaps=False), synth.add_meteo_fields(m1, with_gaps=False)
n = 6000
tm, p, lon, lat = synth.make_parcels(n, t0=0.0, zmin=0.0, zmax=4.0, seed=21)
clim = synth.make_clim_tropo()
orc = Oracle()
for strict in (False, True):
    for nsteps in (2, 3, 4):
        for mode in ("all", "pbl_only"):
            kw = dict(turb_mesox=0.16, turb_mesoz=0.16) if mode == "all" else dict(turb_mesox=0.0, turb_mesoz=0.0)
            ctl = Ctl(advect=2 if mode == "all" else 0, diffusion=1, turb_pbl_scheme=1, turb_dz_trop=0.5, turb_dx_pbl=30.0, turb_dz_pbl=1.0,
                      t_start=0.0, t_stop=1e6, dt_mod=300.0, dt_met=21600.0, **kw)
            with Engine(n, nq=0, device=0, strict=strict) as eng:
                eng.set_ctl(ctl); eng.set_clim_tropo(*clim); eng.set_met(0, m0); eng.set_met(1, m1)
                eng.set_atm(tm, p, lon, lat)
                for s in range(nsteps):
                    eng.run_timestep(300.0 * s)
                out = eng.get_atm(); uv = eng.get_uvwp()
            ref = Parcels(tm, p, lon, lat)
            orc.ctr = 0
            orc.run("timestep", ctl, clim, m0, m1, ref, t=0.0, nsteps=nsteps)
            rel = np.abs(out["p"] - ref.p) / ref.p
            h = [int(np.sum(rel > x)) for x in (1e-12, 1e-10, 1e-8, 1e-6, 1e-4, 1e-2)]
            dlat = np.abs(out["lat"] - ref.lat)
            print(f"strict={strict} steps={nsteps} {mode}: p rel > [1e-12,1e-10,1e-8,1e-6,1e-4,1e-2] = {h}; max {rel.max():.2e}; lat max {dlat.max():.2e}; "
                  f"uvwp maxdiff {np.abs(uv - ref.uvwp).max() if hasattr(ref, 'uvwp') and ref.uvwp is not None else 'n/a'}")
            if not strict and nsteps == 4 and mode == "all":
                bad = np.where(rel > 1e-6)[0][:12]
                for i in bad:
                    print("   ", i, f"p0={p[i]:.4f} out={out['p'][i]:.6f} ref={ref.p[i]:.6f} rel={rel[i]:.2e} lat={lat[i]:.2f}")
