"""Multi-GPU host logic: one process per GPU, parcels sharded contiguously by index, no data-path collective for
transport; one sum-reduction over the box arrays for inter-parcel mixing and gridded output.

``torch.distributed`` is the plumbing (NCCL on the GPUs; the same functions run on CPU tensors over gloo in
tests/test_dist_gloo.py).  Nothing here computes physics.
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


def reduce_boxes(box_sum, box_cnt, group=None):
    """Sum the per-rank partial box arrays over all ranks, in place (the one exchange step of the path)."""
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    dist.all_reduce(box_sum, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(box_cnt, op=dist.ReduceOp.SUM, group=group)


def mixing_step(engine, t: float, device, group=None):
    """module_mixing (src/mptrac.c:5169) across ranks: local accumulate -> all-reduce -> local relaxation."""
    ctl = engine.ctl
    engine.mixing_begin(t)
    nbox = engine.mixing_nbox
    s = device_tensor(engine.device_ptr("mix_sum"), nbox, "f8", device)
    c = device_tensor(engine.device_ptr("mix_cnt"), nbox, "i4", device)
    for iq in ctl.mix_qnt:
        if iq < 0:
            continue
        engine.mixing_accumulate(iq)
        reduce_boxes(s, c, group)      # same stream as the engine (the caller passed torch's current stream)
        engine.mixing_apply(iq)


def grid_output(engine, grid: dict, device, group=None, dst: int = 0):
    """write_grid binning (src/mptrac.c:13840-13872) across ranks; returns (count, sum, sumsq) on rank ``dst``."""
    import torch.distributed as dist
    engine.grid_accumulate(**grid)
    nbox = grid["nx"] * grid["ny"] * grid["nz"]
    nq = max(engine.nq, 1)
    parts = [device_tensor(engine.device_ptr("grid_cnt"), nbox, "i4", device),
             device_tensor(engine.device_ptr("grid_sum"), nbox * nq, "f8", device),
             device_tensor(engine.device_ptr("grid_sq"), nbox * nq, "f8", device)]
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        for x in parts:
            dist.reduce(x, dst=dst, op=dist.ReduceOp.SUM, group=group)
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
