#!/usr/bin/env python
"""Tuning aid (GPU box): time mpb_run_timestep_host in each host-link mode (MPTRAC_B200_HOST_MODE) on the c2 workload
and check that every mode returns the same parcels.  usage: python scripts/e2e_modes.py [steps]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

import bench  # noqa: E402
from mptrac_b200 import Engine, synth  # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 30
wl = dict(bench.WORKLOADS["c2"])
wl["ctl"] = dict(wl["ctl"], sort_dt=-999.0)
ctl, m0, m1, (tm, p, lon, lat, q) = bench.build_inputs(wl, 0, 1)
n = wl["np"]
cases = [("zerocopy", None), ("dma_in", None), ("dma_in", 65536), ("dma_in", 32768), ("dma_out", None), ("dma_out", 65536),
         ("copy", None), ("copy", 65536)]
results = {}
for mode, chunk in cases:
    os.environ["MPTRAC_B200_HOST_MODE"] = mode
    if chunk:
        os.environ["MPTRAC_B200_HOST_CHUNK"] = str(chunk)
    else:
        os.environ.pop("MPTRAC_B200_HOST_CHUNK", None)
    host = torch.from_numpy(np.stack([tm, p, lon, lat])).pin_memory()
    hn = [host[i].numpy() for i in range(4)]
    with Engine(n, nq=0, device=0) as eng:
        eng.set_ctl(ctl)
        eng.set_clim_tropo(*synth.make_clim_tropo())
        eng.set_met(0, m0)
        eng.set_met(1, m1)
        t = 0.0
        for _ in range(5):
            t += bench.DT_MOD
            eng.run_timestep_host(t, *hn, None)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(K):
            t += bench.DT_MOD
            eng.run_timestep_host(t, *hn, None)
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) / K * 1e3
    results[(mode, chunk)] = (ms, host.clone())
    print(f"{mode:9s} chunk {chunk or 'auto':>6}: {ms:.3f} ms/step  {n / ms / 1e6:.3f} G parcel-steps/s", flush=True)
ref = results[("zerocopy", None)][1]
for k, (ms, h) in results.items():
    assert torch.equal(h, ref), f"mode {k} returns different parcels"
print("all modes return identical parcels")
