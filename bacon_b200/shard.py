"""Trajectory sharding over ranks (one process per GPU) — SURVEY.md §8e.

The path has no exchange step: trajectory i belongs to rank i mod G (round-robin, so parameter sweeps
such as the Van der Pol mu-sweep stay balanced), every rank integrates its shard independently, and
the only collectives are at the very end: all-gather of the per-trajectory records (final state, end
time, status, counters) and all-reduce of the statistics.  Backend-agnostic (`nccl` on the GPU box, `gloo` in the CPU tests).
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_indices(n_global, rank, world):
    """Global trajectory indices owned by `rank`: rank, rank+world, ..."""
    return np.arange(rank, n_global, world, dtype=np.int64)


def shard_size(n_global, rank, world):
    return (n_global - rank + world - 1) // world if rank < n_global else 0


def _interleave(buf, n_global):
    """(world, F, n_max) gathered shards -> (F, n_global) in global trajectory order: trajectory k of rank r is global
    index k * world + r, so the answer is ONE permuted copy (the padding of the shorter shards lands at the very end)."""
    world, f, n_max = buf.shape
    return buf.permute(1, 2, 0).reshape(f, n_max * world)[:, :n_global]


def _pad(x, n_max):
    if x.shape[-1] == n_max:
        return x.contiguous()
    return torch.cat([x, x.new_zeros(*x.shape[:-1], n_max - x.shape[-1])], dim=-1)


def gather_final_states(y_end_local, n_global, world=None):
    """All-gather the (dim, n_local) final states of every rank and interleave them back into global
    trajectory order: returns (dim, n_global) on every rank.  Shards may differ in size by one."""
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        return y_end_local
    dim, _ = y_end_local.shape
    n_max = (n_global + world - 1) // world
    flat = y_end_local.new_empty((world * dim, n_max))  # ranks concatenated along dim 0 (what gloo and nccl both accept)
    dist.all_gather_into_tensor(flat, _pad(y_end_local, n_max))
    return _interleave(flat.view(world, dim, n_max), n_global)


def gather_records(out, n_global, world=None):
    """The whole per-trajectory record of SURVEY.md section 8e on every rank, in global trajectory order:
    {y_end (dim, n), t_end (n), status, n_accept, n_reject, n_rhs (n)} from each rank's local result (a dict of tensors
    as returned by solve_ivp_ensemble_device, or anything with those keys).  Two collectives (one per element width:
    the float64 rows and the 32-bit rows) and two permuted copies, whatever the number of ranks."""
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    ints = ("status", "n_accept", "n_reject", "n_rhs")
    if world == 1:
        return {k: out[k] for k in ("y_end", "t_end") + ints}
    y = out["y_end"]
    dim = y.shape[0]
    n_max = (n_global + world - 1) // world
    f64 = torch.cat([y, out["t_end"].reshape(1, -1)], dim=0)                      # (dim + 1, n_local)
    i32 = torch.stack([out[k].to(torch.int32) for k in ints], dim=0)              # (4, n_local)
    g64 = f64.new_empty((world * (dim + 1), n_max))
    g32 = i32.new_empty((world * 4, n_max))
    dist.all_gather_into_tensor(g64, _pad(f64, n_max))
    dist.all_gather_into_tensor(g32, _pad(i32, n_max))
    a = _interleave(g64.view(world, dim + 1, n_max), n_global)
    b = _interleave(g32.view(world, 4, n_max), n_global)
    res = {"y_end": a[:dim], "t_end": a[dim]}
    for j, k in enumerate(ints):
        res[k] = b[j]
    return res


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
