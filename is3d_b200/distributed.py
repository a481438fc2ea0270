"""One process per GPU: shard the freeze-out surface by cell, run the spectra kernel on the local shard, combine the
small per-species spectra array with ONE all-reduce (NCCL over NVLink on the GPU box, gloo in the CPU tests).

The path shards naturally: cells are independent and the only coupling is the final sum, so there is no
data-path collective inside the kernel.  Everything global -- species list, momentum tables, delta-f tables, the Jonah
lambda/z tables built at the *surface-average* temperature -- is computed once before sharding and replicated.
"""
import numpy as np


def shard_bounds(n_cells, rank, world_size):
    """Contiguous cell range [lo, hi) of `rank`: ceil(N / G) cells per rank, the last ranks may get fewer (or none)."""
    per = -(-int(n_cells) // int(world_size)) if n_cells > 0 else 0
    lo = min(rank * per, n_cells)
    hi = min(lo + per, n_cells)
    return lo, hi


def shard_cells(cells, rank, world_size):
    n = len(cells["tau"])
    lo, hi = shard_bounds(n, rank, world_size)
    return {k: (v[lo:hi] if v is not None else None) for k, v in cells.items()}


def smooth_spectra_sharded(flags, cells, species, grid, df_tables=None, laguerre=None, memory="host", kernel=None,
                           group=None, already_sharded=False, **kw):
    """Run the local shard through `kernel` (default: the CUDA C ABI) and all-reduce the spectra.

    Returns (dN summed over ranks, local stats).  With memory='device' the all-reduce runs on the CUDA tensor (NCCL);
    with memory='host' the numpy result is reduced through a CPU tensor (gloo).
    """
    import torch
    import torch.distributed as dist
    if kernel is None:
        from . import api
        kernel = api.smooth_spectra
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    local = cells if already_sharded else shard_cells(cells, rank, world)
    dN, stats = kernel(flags, local, species, grid, df_tables, laguerre, memory=memory, **kw)
    if world > 1:
        if memory == "device":
            dist.all_reduce(dN, op=dist.ReduceOp.SUM, group=group)
        else:
            t = torch.from_numpy(np.ascontiguousarray(dN))
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
            dN = t.numpy()
    return dN, stats


SPACETIME_KEYS = ("dN_tau", "dN_r", "dN_taur", "dN_dydeta", "dN_dy")


def spacetime_distributions_sharded(flags, cells, species, grid, df_tables, laguerre, bins, memory="host", kernel=None,
                                    group=None, already_sharded=False, **kw):
    """operation = 0 on the local shard + ONE all-reduce of the concatenated histograms (they are sums over cells, so the
    shards simply add).  Returns (dict of reduced raw sums, local stats)."""
    import torch
    import torch.distributed as dist
    if kernel is None:
        from . import api
        kernel = api.spacetime_distributions
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    local = cells if already_sharded else shard_cells(cells, rank, world)
    res, stats = kernel(flags, local, species, grid, df_tables, laguerre, bins, memory=memory, **kw)
    if world > 1:
        shapes = [np.asarray(res[k]).shape for k in SPACETIME_KEYS]
        flat = torch.from_numpy(np.concatenate([np.asarray(res[k], dtype=np.float64).ravel() for k in SPACETIME_KEYS]))
        if memory == "device":
            flat = flat.cuda()
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat = flat.cpu().numpy()
        off = 0
        for k, shp in zip(SPACETIME_KEYS, shapes):
            n = int(np.prod(shp)); res[k] = flat[off:off + n].reshape(shp); off += n
    return res, stats
