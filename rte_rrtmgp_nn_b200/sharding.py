"""Column sharding across ranks (one process per GPU).

Every column is independent through gas optics and both solvers (the reference's only parallelism is OpenMP over
column blocks, examples/rfmip-clear-sky/rrtmgp_rfmip_lw.F90:364-368), so the column dimension is cut into contiguous
shards, spectral tables and weights are replicated, and the ONLY collective is the final gather of the broadband
fluxes (flux_up/dn [, dir], (nlay+1) floats per column each).  Works with NCCL (GPUs over NVLink/NVSwitch) and with
gloo (CPU tests).
"""
import numpy as np


def shard_bounds(ncol, rank, world):
    """Contiguous shard [c0, c1) of rank; sizes differ by at most one column."""
    c0 = (ncol * rank) // world
    c1 = (ncol * (rank + 1)) // world
    return c0, c1


def max_shard(ncol, world):
    return max(shard_bounds(ncol, r, world)[1] - shard_bounds(ncol, r, world)[0] for r in range(world))


def gather_fluxes(local, ncol_total, group=None):
    """All-gather a list of per-rank flux tensors [(ncol_local, nlev), ...] into [(ncol_total, nlev), ...] on every rank.
    One collective for all arrays: they are packed into equal-size slots (largest shard) and unpacked by shard bounds."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    nloc, nlev = local[0].shape
    nmax = max_shard(ncol_total, world)
    send = torch.zeros((len(local), nmax, nlev), dtype=local[0].dtype, device=local[0].device)
    for i, t in enumerate(local):
        send[i, :nloc].copy_(t)
    # output = the ranks' buffers concatenated along dim 0 (the form both NCCL and gloo accept)
    flat = torch.empty((world * len(local), nmax, nlev), dtype=send.dtype, device=send.device)
    dist.all_gather_into_tensor(flat, send, group=group)
    recv = flat.view(world, len(local), nmax, nlev)
    out = []
    for i in range(len(local)):
        parts = []
        for r in range(world):
            c0, c1 = shard_bounds(ncol_total, r, world)
            parts.append(recv[r, i, :c1 - c0])
        out.append(torch.cat(parts, dim=0))
    return out
