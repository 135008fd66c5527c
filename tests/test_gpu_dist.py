"""Several ranks on the GPU(s): the sharded run -- transport without any collective, inter-parcel mixing and gridded output
with ONE exchange step each -- against a single-context run of all parcels.  Three transports of the exchange step:

  team   several contexts behind one host thread (mpb_team_*), box records exchanged through peer memory, stream events;
  peers  one process per rank attached through CUDA IPC (mpb_peer_*): NVLink atomics + flag barriers in stream order;
  allreduce  one process per rank, ONE all-reduce of the dense box records: NCCL with one GPU per rank, gloo (which moves CUDA
         tensors through the host) when both ranks have to share device 0.

All three also run on a single GPU (two contexts / two processes on device 0), so the exchange logic is exercised by every
run of the suite; with two or more GPUs the same tests use distinct devices (real peer traffic, NCCL)."""
import os
import socket

import numpy as np
import pytest

from conftest import ROOT  # noqa: F401

pytestmark = pytest.mark.gpu

GRID = dict(nx=36, ny=18, nz=4, lon0=-180.0, lon1=180.0, lat0=-90.0, lat1=90.0, z0=0.0, z1=40.0)
NSTEPS = 4


def _inputs(nq=2):
    from mptrac_b200 import Ctl, synth
    n = 200_001
    m0, m1 = synth.make_met_pair(72, 37, 30, t0=0.0, dt_met=21600.0)
    tm, p, lon, lat = synth.make_parcels(n, t0=0.0, zmin=0.1, zmax=40.0, seed=11)
    q = np.random.default_rng(5).uniform(0, 1, (nq, n))
    ctl = Ctl(nq=nq, advect=4, diffusion=1, t_start=0.0, t_stop=1e6, dt_mod=300.0, dt_met=21600.0, mixing_trop=0.4,
              mixing_strat=0.1, mixing_dt=300.0, mix_qnt=list(range(nq)), mixing_nx=36, mixing_ny=18, mixing_nz=15)
    return n, m0, m1, tm, p, lon, lat, q, synth.make_clim_tropo(), ctl


def _grid_args(nsteps=NSTEPS):
    return dict(GRID, t0=300.0 * (nsteps - 1) - 150.0, t1=300.0 * (nsteps - 1) + 150.0)


def _single():
    """all parcels in one context: the run every sharded variant must reproduce"""
    from mptrac_b200 import Engine
    n, m0, m1, tm, p, lon, lat, q, clim, ctl = _inputs()
    with Engine(n, nq=ctl.nq, device=0) as eng:
        eng.set_ctl(ctl); eng.set_clim_tropo(*clim); eng.set_met(0, m0); eng.set_met(1, m1)
        eng.set_atm(tm, p, lon, lat, q)
        for s in range(NSTEPS):
            eng.run_timestep(300.0 * s)
        eng.grid_accumulate(**_grid_args())
        grid = eng.grid_fetch()
        one = eng.get_atm()
    return one, grid, q


def _check(full, grid, one, one_grid, q0):
    # transport is bit-identical (random numbers are addressed by global parcel index)
    for k in ("lon", "lat", "p"):
        assert np.array_equal(full[k], one[k]), k
    # box means are sums of doubles accumulated in a different order: agreement to rounding, counts exact
    assert np.max(np.abs(full["q"] - one["q"])) < 1e-12
    assert np.max(np.abs(one["q"] - q0)) > 1e-3
    assert np.array_equal(grid[0], one_grid[0]) and grid[0].sum() > 0
    assert np.allclose(grid[1], one_grid[1], rtol=1e-12, atol=1e-12) and np.allclose(grid[2], one_grid[2], rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("members", [2, 3])
def test_team_matches_single(members):
    """mpb_team_*: `members` contexts behind one host thread -- on distinct GPUs when the box has them, else all on device 0"""
    from mptrac_b200 import Team, load_library
    ndev = load_library().mpb_device_count()
    devices = [i % ndev for i in range(members)]
    n, m0, m1, tm, p, lon, lat, q, clim, ctl = _inputs()
    with Team(devices, n, nq=ctl.nq) as team:
        team.set_ctl(ctl); team.set_clim_tropo(*clim); team.set_met(0, m1); team.set_met(1, m0); team.swap_met()
        team.set_atm(tm, p, lon, lat, q)
        for s in range(NSTEPS):
            team.run_timestep(300.0 * s)
        team.grid_accumulate(**_grid_args())
        grid = team.grid_fetch()
        full = team.get_atm()
        assert team.launch_count > 0
    one, one_grid, q0 = _single()
    _check(full, grid, one, one_grid, q0)


def _worker(rank, world, port, ret, transport, ndev):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    device = rank % ndev
    torch.cuda.set_device(device)
    if transport == "allreduce" and ndev >= world:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", device))
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from mptrac_b200 import Engine
        from mptrac_b200 import dist as mdist
        from mptrac_b200.host import MOD_ALL, MOD_MIXING
        n, m0, m1, tm, p, lon, lat, q, clim, ctl = _inputs()
        lo, hi = mdist.shard_bounds(n, world)[rank]
        dev = torch.device("cuda", device)
        stream = torch.cuda.Stream()
        torch.cuda.set_stream(stream)
        with Engine(hi - lo, nq=ctl.nq, device=device) as eng:
            eng.set_stream(stream.cuda_stream)      # the collective runs on torch's current stream: the engine shares it
            eng.set_ctl(ctl); eng.set_clim_tropo(*clim); eng.set_met(0, m0); eng.set_met(1, m1)
            eng.set_atm(tm[lo:hi], p[lo:hi], lon[lo:hi], lat[lo:hi], np.ascontiguousarray(q[:, lo:hi]))
            eng.set_shard(lo, n)
            attached = transport == "peers" and mdist.attach_peers(eng, ctl, GRID["nx"] * GRID["ny"] * GRID["nz"])
            for s in range(NSTEPS):
                t = 300.0 * s
                if attached:
                    eng.run_timestep(t)             # the exchange happens inside (module_mixing)
                else:
                    eng.run_modules(t, MOD_ALL & ~MOD_MIXING)
                    mdist.mixing_step(eng, t, dev)
            grid = mdist.grid_output(eng, _grid_args(), dev, attached=attached)
            eng.sync()
            out = eng.get_atm()
            dist.barrier()                          # nobody frees its exchange area while a peer may still read it
        side = dist.new_group(backend="gloo") if dist.get_backend() == "nccl" else None
        full = mdist.gather_parcels({"lon": out["lon"], "lat": out["lat"], "p": out["p"], "q0": out["q"][0], "q1": out["q"][1]},
                                    n, group=side)
        if rank == 0:
            ret["full"] = {k: np.array(v) for k, v in full.items()}
            ret["grid"] = [np.array(g) for g in grid]
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(900)
@pytest.mark.parametrize("transport", ["peers", "allreduce"])
def test_two_processes_match_single(transport):
    import torch
    import torch.multiprocessing as mp
    ndev = torch.cuda.device_count()
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(2, port, ret, transport, ndev), nprocs=2, join=True)
        full, grid = ret["full"], ret["grid"]
    full["q"] = np.stack([full.pop("q0"), full.pop("q1")])
    one, one_grid, q0 = _single()
    _check(full, grid, one, one_grid, q0)


def test_team_with_cell_sort_moves_every_parcel_like_the_single_run():
    """SORT_DT > 0 on a team: every member sorts ITS index range (no global order is needed), so the array order
    differs from the single-context run -- but without diffusion each parcel's trajectory does not depend on its slot, so
    matched by an identity quantity the positions must be bit-identical.  (With diffusion on, random numbers and uvwp are
    attached to the slot: a sharded run with SORT_DT > 0 then agrees in distribution only -- DESIGN.md 4.)"""
    from mptrac_b200 import Ctl, Engine, Team, load_library, synth
    ndev = load_library().mpb_device_count()
    n = 150_001
    m0, m1 = synth.make_met_pair(72, 37, 30, t0=0.0, dt_met=21600.0)
    tm, p, lon, lat = synth.make_parcels(n, t0=0.0, zmin=0.1, zmax=40.0, seed=3)
    ident = np.arange(n, dtype=np.float64)[None, :]
    ctl = Ctl(nq=1, advect=4, diffusion=0, sort_dt=600.0, t_start=0.0, t_stop=1e6, dt_mod=300.0, dt_met=21600.0)
    clim = synth.make_clim_tropo()
    outs = []
    for make in (lambda: Team([i % ndev for i in range(3)], n, nq=1), lambda: Engine(n, nq=1, device=0)):
        with make() as eng:
            eng.set_ctl(ctl); eng.set_clim_tropo(*clim); eng.set_met(0, m0); eng.set_met(1, m1)
            eng.set_atm(tm, p, lon, lat, ident)
            for s in range(6):
                eng.run_timestep(300.0 * s)
            outs.append(eng.get_atm())
    a, b = outs
    ia, ib = np.argsort(a["q"][0]), np.argsort(b["q"][0])
    assert np.array_equal(a["q"][0][ia], np.arange(n)) and np.array_equal(b["q"][0][ib], np.arange(n))
    assert not np.array_equal(a["q"][0], b["q"][0])          # the two runs do order the parcels differently
    for k in ("time", "lon", "lat", "p"):
        assert np.array_equal(a[k][ia], b[k][ib]), k
