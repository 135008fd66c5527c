"""Multi-GPU host logic: one process per GPU, parcels sharded contiguously by index, no data-path collective for
transport; ONE exchange step per model step for inter-parcel mixing and one for gridded output.

Two transports for that step: the engines' own kernels over peer memory (``attach_peers``: NVLink atomics into the owner's
slice of the box records, flag barriers; the default of bench.py) or one ``torch.distributed`` all-reduce of the dense box
records (``mixing_step``; NCCL on the GPUs, gloo on CPU tensors in tests/test_dist_gloo.py).  ``torch.distributed`` is the
plumbing either way (rendezvous, handle exchange).  Nothing here computes physics.
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np


def shard_bounds(n_total: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous index ranges [lo, hi) per rank; sizes differ by at most one parcel."""
    base, rem = divmod(int(n_total), int(world))
    out, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < rem else 0)
        out.append((lo, hi))
        lo = hi
    return out


def rng_draws_per_module(n_total: int) -> int:
    """Counters one module_rng(3*np, normal) call consumes (src/mptrac.c:5793-5812): the same on every rank,
    because random numbers are addressed by GLOBAL parcel index."""
    return 3 * int(n_total) + 1


def device_tensor(ptr: int, n: int, dtype: str, device):
    """Alias ``n`` elements of engine-owned device memory as a torch tensor (no copy)."""
    import torch
    typestr = {"f8": "<f8", "i4": "<i4"}[dtype]

    class _Wrap:
        __cuda_array_interface__ = {"shape": (int(n),), "typestr": typestr, "data": (int(ptr), False), "version": 2}

    return torch.as_tensor(_Wrap(), device=device)


def attach_peers(engine, ctl, grid_boxes: int = 0, group=None) -> bool:
    """Connect the ranks' engines through peer memory (one process per GPU on one node): every rank allocates its exchange
    area, the CUDA IPC handles travel over ``torch.distributed``, every rank opens the others'.  Afterwards
    ``engine.run_timestep`` / ``module_mixing`` / ``grid_reduce`` do their exchange steps themselves (contributions routed to
    the owner of each box as coalesced stores over NVLink, answers routed back, flag barriers in stream order; no collective
    library on the data path).  Returns False (and leaves the
    engine alone) when there is only one rank."""
    import torch.distributed as dist
    from .host import exchange_area_bytes
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return False
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    # (every rank must size its area for the LARGEST shard: the capacity of an inbox is derived from the area size)
    sizes = [None] * world
    dist.all_gather_object(sizes, int(engine.np_max), group=group)
    mix_bytes, grid_bytes = exchange_area_bytes(ctl, world, engine.nq, max(sizes), grid_boxes)
    err = None
    try:
        handle = engine.peer_init(rank, world, mix_bytes, grid_bytes)
    except Exception as exc:           # (e.g. out of memory) -- decided collectively below
        handle, err = None, repr(exc)
    handles = [None] * world
    dist.all_gather_object(handles, handle, group=group)
    if err is None and all(h is not None for h in handles):
        try:
            engine.peer_attach(handles)
        except Exception as exc:       # CUDA IPC not permitted between these processes
            err = repr(exc)
    errs = [None] * world
    dist.all_gather_object(errs, err, group=group)   # also the barrier: every rank has zeroed and mapped its area
    if any(e is not None for e in errs) or any(h is None for h in handles):
        engine.peer_init(0, 1, 0, 0)   # back to a single, unattached rank on every rank: the caller uses the all-reduce path
        return False
    return True


def reduce_records(rec, group=None):
    """Sum the per-rank box records {count, sum of every mixed quantity} over all ranks, in place: ONE all-reduce per model
    step whatever the number of mixed quantities (counts travel as doubles; integers below 2^53 stay exact)."""
    import torch.distributed as dist
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(rec, op=dist.ReduceOp.SUM, group=group)


def mixing_step(engine, t: float, device, group=None):
    """module_mixing (src/mptrac.c:5169) across ranks that are NOT attached through peer memory: local box records of all
    mixed quantities -> ONE all-reduce -> local relaxation.  (Attached ranks just call ``engine.module_mixing``.)"""
    import torch.distributed as dist
    engine.mixing_accumulate_all(t)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        # (same stream as the engine: the caller made torch's current stream the engine's)
        reduce_records(device_tensor(engine.device_ptr("mix_rec"), engine.mixing_rec_len, "f8", device), group)
    engine.mixing_apply_all()


def grid_output(engine, grid: dict, device, group=None, dst: int = 0, attached: bool = False):
    """write_grid binning (src/mptrac.c:13840-13872) across ranks; returns (count, sum, sumsq) on rank ``dst``.
    ``attached``: the ranks are connected through peer memory -- rank 0 adds up the partial arrays itself."""
    import torch.distributed as dist
    engine.grid_accumulate(**grid)
    nbox = grid["nx"] * grid["ny"] * grid["nz"]
    nq = max(engine.nq, 1)
    multi = dist.is_initialized() and dist.get_world_size(group) > 1
    if multi and attached:
        assert dst == 0
        engine.grid_reduce()
        if dist.get_rank(group) != 0:
            return None
    elif multi:
        parts = [device_tensor(engine.device_ptr("grid_cnt"), nbox, "i4", device),
                 device_tensor(engine.device_ptr("grid_sum"), nbox * nq, "f8", device),
                 device_tensor(engine.device_ptr("grid_sq"), nbox * nq, "f8", device)]
        for x in parts:      # (an all-reduce: 1.3 MB, and unlike reduce it is available for device tensors on every backend)
            dist.all_reduce(x, op=dist.ReduceOp.SUM, group=group)
        if dist.get_rank(group) != dst:
            return None
    return engine.grid_fetch()


def gather_parcels(local: dict, n_total: int, group=None):
    """All ranks' parcel arrays -> full arrays on every rank (used for output / tests; not on the step path)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local
    bounds = shard_bounds(n_total, world)
    out = {}
    for k, a in local.items():
        a = np.ascontiguousarray(a)
        mx = max(hi - lo for lo, hi in bounds)
        pad = np.zeros(mx, a.dtype)
        pad[: a.size] = a
        bufs = [torch.zeros(mx, dtype=torch.from_numpy(pad).dtype) for _ in range(world)]
        dist.all_gather(bufs, torch.from_numpy(pad), group=group)
        out[k] = np.concatenate([b.numpy()[: hi - lo] for b, (lo, hi) in zip(bufs, bounds)])
    return out
