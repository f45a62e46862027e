"""Trajectory sharding over ranks (one process per GPU) — SURVEY.md §8e.

The path has no exchange step: trajectory i belongs to rank i mod G (round-robin, so parameter sweeps
such as the Van der Pol mu-sweep stay balanced), every rank integrates its shard independently, and
the only collectives are at the very end: all-gather of the final states and all-reduce of the
statistics.  Backend-agnostic (`nccl` on the GPU box, `gloo` in the CPU tests).
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_indices(n_global, rank, world):
    """Global trajectory indices owned by `rank`: rank, rank+world, ..."""
    return np.arange(rank, n_global, world, dtype=np.int64)


def shard_size(n_global, rank, world):
    return (n_global - rank + world - 1) // world if rank < n_global else 0


def gather_final_states(y_end_local, n_global, world=None):
    """All-gather the (dim, n_local) final states of every rank and interleave them back into global
    trajectory order: returns (dim, n_global) on every rank.  Shards may differ in size by one."""
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        return y_end_local
    dim, n_loc = y_end_local.shape
    n_max = (n_global + world - 1) // world
    pad = y_end_local
    if n_loc < n_max:
        pad = torch.cat([y_end_local, y_end_local.new_zeros(dim, n_max - n_loc)], dim=1)
    flat = y_end_local.new_empty((world * dim, n_max))  # ranks concatenated along dim 0 (what gloo and nccl both accept)
    dist.all_gather_into_tensor(flat, pad.contiguous())
    buf = flat.view(world, dim, n_max)
    out = y_end_local.new_empty((dim, n_global))
    for r in range(world):
        m = shard_size(n_global, r, world)
        out[:, r::world] = buf[r, :, :m]
    return out


def reduce_stats_device(n_accept, n_reject, n_rhs, status):
    """[sum n_accept, sum n_reject, sum n_rhs, #failed] over all ranks, as a float64 tensor that stays on
    the device (no host synchronisation: usable inside a timed region)."""
    sums = torch.stack([n_accept.sum(dtype=torch.float64), n_reject.sum(dtype=torch.float64),
                        n_rhs.sum(dtype=torch.float64), (status != 0).sum(dtype=torch.float64)])
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    return sums


def reduce_stats(n_accept, n_reject, n_rhs, status, kernel_ms=0.0):
    """Global totals (sum) and the slowest rank's kernel time (max). Returns a dict of Python numbers."""
    dev = n_accept.device
    sums = reduce_stats_device(n_accept, n_reject, n_rhs, status)
    tmax = torch.tensor([float(kernel_ms)], dtype=torch.float64, device=dev)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    s = sums.cpu().tolist()
    return {"n_accept": s[0], "n_reject": s[1], "n_rhs": s[2], "n_failed": s[3], "kernel_ms_max": float(tmax.cpu()[0])}
