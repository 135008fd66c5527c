"""Two ranks, two GPUs, NCCL: the sharded run (transport without any collective, inter-parcel mixing and gridded output
through ONE sum-reduction of the box arrays) against a single-GPU run of all parcels.  Skipped on a one-GPU box; the
host logic itself is covered on CPU by tests/test_dist_gloo.py."""
import os
import socket

import numpy as np
import pytest

from conftest import ROOT  # noqa: F401

pytestmark = pytest.mark.gpu

GRID = dict(nx=36, ny=18, nz=4, lon0=-180.0, lon1=180.0, lat0=-90.0, lat1=90.0, z0=0.0, z1=40.0)


def _inputs():
    from mptrac_b200 import Ctl, synth
    n = 200_001
    m0, m1 = synth.make_met_pair(72, 37, 30, t0=0.0, dt_met=21600.0)
    tm, p, lon, lat = synth.make_parcels(n, t0=0.0, zmin=0.1, zmax=40.0, seed=11)
    q = np.random.default_rng(5).uniform(0, 1, (1, n))
    ctl = Ctl(nq=1, advect=4, diffusion=1, t_start=0.0, t_stop=1e6, dt_mod=300.0, dt_met=21600.0, mixing_trop=0.4,
              mixing_strat=0.1, mixing_dt=300.0, mix_qnt=[0], mixing_nx=36, mixing_ny=18, mixing_nz=15)
    return n, m0, m1, tm, p, lon, lat, q, synth.make_clim_tropo(), ctl


def _run(eng, ctl, device, nsteps, group=None):
    """transport on the device, mixing through dist.mixing_step (all-reduce between accumulate and apply)"""
    from mptrac_b200 import dist as mdist
    from mptrac_b200.host import MOD_ALL, MOD_MIXING
    for s in range(nsteps):
        t = 300.0 * s
        eng.run_modules(t, MOD_ALL & ~MOD_MIXING)
        mdist.mixing_step(eng, t, device, group)
    return mdist.grid_output(eng, dict(GRID, t0=300.0 * (nsteps - 1) - 150.0, t1=300.0 * (nsteps - 1) + 150.0), device, group)


def _worker(rank, world, port, ret):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from mptrac_b200 import Engine
        from mptrac_b200.dist import gather_parcels, shard_bounds
        n, m0, m1, tm, p, lon, lat, q, clim, ctl = _inputs()
        lo, hi = shard_bounds(n, world)[rank]
        dev = torch.device("cuda", rank)
        stream = torch.cuda.Stream()
        torch.cuda.set_stream(stream)
        with Engine(hi - lo, nq=1, device=rank) as eng:
            eng.set_stream(stream.cuda_stream)      # NCCL runs on torch's current stream: the engine shares it
            eng.set_ctl(ctl); eng.set_clim_tropo(*clim); eng.set_met(0, m0); eng.set_met(1, m1)
            eng.set_atm(tm[lo:hi], p[lo:hi], lon[lo:hi], lat[lo:hi], np.ascontiguousarray(q[:, lo:hi]))
            eng.set_shard(lo, n)
            grid = _run(eng, ctl, dev, 4)
            out = eng.get_atm()
        # gather for the comparison goes over a gloo side group (numpy arrays on the host)
        side = dist.new_group(backend="gloo")
        full = gather_parcels({"lon": out["lon"], "lat": out["lat"], "p": out["p"], "q": out["q"][0]}, n, group=side)
        if rank == 0:
            ret["full"] = {k: np.array(v) for k, v in full.items()}
            ret["grid"] = [np.array(g) for g in grid]
    finally:
        dist.destroy_process_group()


def test_two_gpu_sharded_run_with_mixing_and_grid_matches_single():
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
        full, grid = ret["full"], ret["grid"]
    from mptrac_b200 import Engine
    n, m0, m1, tm, p, lon, lat, q, clim, ctl = _inputs()
    dev = torch.device("cuda", 0)
    with Engine(n, nq=1, device=0) as eng:
        eng.set_ctl(ctl); eng.set_clim_tropo(*clim); eng.set_met(0, m0); eng.set_met(1, m1)
        eng.set_atm(tm, p, lon, lat, q)
        one_grid = _run(eng, ctl, dev, 4)
        one = eng.get_atm()
    # transport is bit-identical (random numbers are addressed by global parcel index)
    for k in ("lon", "lat", "p"):
        assert np.array_equal(full[k], one[k]), k
    # box means are sums of doubles accumulated in a different order: agreement to rounding, counts exact
    assert np.max(np.abs(full["q"] - one["q"][0])) < 1e-12
    assert np.max(np.abs(one["q"][0] - q[0])) > 1e-3
    assert np.array_equal(grid[0], one_grid[0])
    assert np.allclose(grid[1], one_grid[1], rtol=1e-12, atol=1e-12) and np.allclose(grid[2], one_grid[2], rtol=1e-12, atol=1e-12)
