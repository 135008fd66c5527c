import os, sys; sys.path.insert(0, ".")
os.environ["MPTRAC_B200_TRACE"] = "1"
import numpy as np, torch, time
from mptrac_b200 import Ctl, Engine, synth
n = 1_000_000
m0, m1 = synth.make_met_pair(360, 181, 60, t0=0.0, dt_met=21600.0)
tm, p, lon, lat = synth.make_parcels(n, t0=0.0)
host = torch.from_numpy(np.stack([tm, p, lon, lat])).pin_memory()
h = [host[i].numpy() for i in range(4)]
ctl = Ctl(advect=4, t_start=0.0, t_stop=1e9, dt_mod=300.0, dt_met=21600.0)
with Engine(n) as eng:
    eng.set_ctl(ctl); eng.set_clim_tropo(*synth.make_clim_tropo()); eng.set_met(0, m0); eng.set_met(1, m1)
    for s in range(1, 6):
        t0 = time.perf_counter()
        eng.run_timestep_host(300.0 * s, *h)
        print(f"step {s}: wall {1e6*(time.perf_counter()-t0):.0f} us", file=sys.stderr)
