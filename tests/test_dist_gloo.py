"""world_size-2 test of the multi-GPU host logic on CPU (gloo): contiguous sharding, global-index random numbers
and the box-array reduction of inter-parcel mixing.  The per-rank compute stand-in is the oracle (this is a test of
the host plumbing, not of the kernels)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT  # noqa: F401


def test_shard_bounds():
    from mptrac_b200.dist import rng_draws_per_module, shard_bounds
    assert shard_bounds(10, 3) == [(0, 4), (4, 7), (7, 10)]
    assert shard_bounds(0, 2) == [(0, 0), (0, 0)]
    b = shard_bounds(100_000_000, 8)
    assert b[0] == (0, 12_500_000) and b[-1][1] == 100_000_000
    assert all(b[i][1] == b[i + 1][0] for i in range(7))
    assert rng_draws_per_module(10) == 31


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from mptrac_b200 import Ctl, synth
        from mptrac_b200.dist import gather_parcels, reduce_records, shard_bounds
        from oracle.oracle import Oracle, Parcels
        n = 20001
        m0, m1 = synth.make_met_pair(36, 19, 20, t0=0.0, dt_met=21600.0)
        tm, p, lon, lat = synth.make_parcels(n, t0=0.0, zmin=0.1, zmax=40.0, seed=11)
        q = np.random.default_rng(5).uniform(0, 1, (1, n))
        clim = synth.make_clim_tropo()
        ctl = Ctl(nq=1, advect=2, diffusion=1, t_start=0.0, t_stop=1e6, dt_mod=300.0, dt_met=21600.0, mixing_trop=0.4,
                  mixing_strat=0.1, mixing_dt=300.0, mix_qnt=[0], mixing_nx=36, mixing_ny=18, mixing_nz=15)
        lo, hi = shard_bounds(n, world)[rank]
        orc = Oracle()
        a = Parcels(tm[lo:hi], p[lo:hi], lon[lo:hi], lat[lo:hi], q[:, lo:hi])
        nl = hi - lo
        # ---- transport for the shard: counters are global => draw the full stream, use the shard's slice ----
        ctr = 0
        for s in range(3):
            t = 300.0 * s
            orc.run("timesteps", ctl, clim, m0, m1, a, t=t)
            orc.run("position", ctl, clim, m0, m1, a)
            orc.run("advect", ctl, clim, m0, m1, a)
            for mod in ("diff_turb", "diff_meso"):
                orc.ctr = ctr + 3 * lo          # parcel ig uses counters ctr + 3*ig .. : shift the local stream
                if (3 * lo) % 2 == 1:           # pairs are aligned on even GLOBAL indices
                    # odd offset: run one parcel earlier so the pair alignment is global, then drop it
                    b = Parcels(np.r_[tm[lo - 1], a.time], np.r_[p[lo - 1], a.p], np.r_[lon[lo - 1], a.lon],
                                np.r_[lat[lo - 1], a.lat], np.c_[q[:, lo - 1], a.q], np.r_[np.zeros((1, 3), np.float32), a.uvwp],
                                np.r_[0.0, a.dt])
                    orc.ctr = ctr + 3 * (lo - 1)
                    orc.run(mod, ctl, clim, m0, m1, b)
                    a.lon[:], a.lat[:], a.p[:], a.uvwp[:] = b.lon[1:], b.lat[1:], b.p[1:], b.uvwp[1:]
                else:
                    orc.run(mod, ctl, clim, m0, m1, a)
                ctr += 3 * n + 1
            orc.run("position", ctl, clim, m0, m1, a)
            # ---- mixing: local box records -> reduce_records (one all-reduce) -> local relaxation ----
            ngrid = ctl.mixing_nx * ctl.mixing_ny * ctl.mixing_nz
            z = 7.0 * np.log(1013.25 / a.p)
            ok = (np.abs(a.time - t) <= 150.0) & (a.lon >= -180) & (a.lon < 180) & (a.lat >= -90) & (a.lat < 90) & (z >= -5) & (z < 85)
            ix = ((a.lon + 180.0) / (360.0 / 36)).astype(int); iy = ((a.lat + 90.0) / (180.0 / 18)).astype(int); iz = ((z + 5.0) / (90.0 / 15)).astype(int)
            box = np.where(ok, (ix * 18 + iy) * 15 + iz, -1)
            rec = np.zeros((ngrid, 2))        # the engine's box records: {count, sum of the mixed quantity}
            np.add.at(rec[:, 0], box[box >= 0], 1.0); np.add.at(rec[:, 1], box[box >= 0], a.q[0][box >= 0])
            reduce_records(torch.from_numpy(rec))
            scnt, ssum = rec[:, 0].astype(np.int64), rec[:, 1]
            mean = np.where(scnt > 0, ssum / np.maximum(scnt, 1), 0.0)
            w = np.array([_tropo_w(orc, clim, a.time[i], a.lat[i], a.p[i]) for i in range(nl)])
            mix = w * ctl.mixing_trop + (1 - w) * ctl.mixing_strat
            sel = box >= 0
            a.q[0][sel] += (mean[box[sel]] - a.q[0][sel]) * mix[sel]
        full = gather_parcels({"lon": a.lon, "lat": a.lat, "p": a.p, "q": a.q[0]}, n)
        if rank == 0:
            one = Parcels(tm, p, lon, lat, q)
            orc.ctr = 0
            orc.run("timestep", ctl, clim, m0, m1, one, t=0.0, nsteps=3)
            ret["pos"] = all(np.array_equal(full[k], getattr(one, k)) for k in ("lon", "lat", "p"))
            ret["q"] = float(np.max(np.abs(full["q"] - one.q[0])))
            ret["moved"] = float(np.max(np.abs(one.q[0] - q[0])))
    finally:
        dist.destroy_process_group()


def _tropo_w(orc, clim, t, lat, p):
    pt = orc.clim_tropo(clim, t, lat)
    p1, p0 = pt * 0.866877899, pt / 0.866877899
    if p > p0:
        return 1.0
    if p < p1:
        return 0.0
    return 1.0 + (0.0 - 1.0) / (p1 - p0) * (p - p0)


@pytest.mark.timeout(600)
def test_two_rank_sharded_run_matches_single():
    port = _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
        assert ret["pos"], "sharded transport (global-index RNG) differs from the single-rank run"
        assert ret["moved"] > 1e-3 and ret["q"] < 1e-12
