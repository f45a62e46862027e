"""world_size-2 gloo tests (CPU) of the N>1 host logic: round-robin sharding, gather of final states
back into global order, reduction of statistics.  The per-rank solver here is the CPU oracle (test
infrastructure) standing in for the CUDA kernel, so the collective plumbing can be checked without a GPU."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from bacon_b200 import _abi, ensembles as E
from bacon_b200.shard import gather_final_states, gather_records, reduce_stats, shard_indices, shard_size


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_global, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O
    idx = shard_indices(n_global, rank, world)
    y0 = E.lorenz_y0(idx)
    r = O.solve_ensemble(_abi.RK45, "lorenz", y0, np.array(E.LORENZ["params"]), shared_params=True, dt_min=1e-9,
                         dt_max=0.1, tol=1e-8, t_start=0.0, t_end=0.3, n_threads=1)
    y_glob = gather_final_states(torch.from_numpy(r["y_end"]), n_global)
    stats = reduce_stats(torch.from_numpy(r["n_accept"].astype(np.int64)), torch.from_numpy(r["n_reject"].astype(np.int64)),
                         torch.from_numpy(r["n_rhs"].astype(np.int64)), torch.from_numpy(r["status"]), kernel_ms=10.0 + rank)
    rec = gather_records({k: torch.from_numpy(np.ascontiguousarray(r[k])) for k in ("y_end", "t_end", "status", "n_accept", "n_reject", "n_rhs")},
                         n_global)
    np.save(os.path.join(out_dir, f"y_{rank}.npy"), y_glob.numpy())
    np.savez(os.path.join(out_dir, f"rec_{rank}.npz"), **{k: v.numpy() for k, v in rec.items()})
    np.save(os.path.join(out_dir, f"s_{rank}.npy"), np.array([stats["n_accept"], stats["n_reject"], stats["n_rhs"],
                                                               stats["n_failed"], stats["kernel_ms_max"]]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_global", [257, 64])
def test_sharded_solve_matches_single_process(oracle, tmp_path, n_global):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), n_global, str(tmp_path)), nprocs=world, join=True)
    full = oracle.solve_ensemble(_abi.RK45, "lorenz", E.lorenz_y0(np.arange(n_global)), np.array(E.LORENZ["params"]),
                                 shared_params=True, dt_min=1e-9, dt_max=0.1, tol=1e-8, t_start=0.0, t_end=0.3)
    for rank in range(world):
        y = np.load(tmp_path / f"y_{rank}.npy")
        assert np.array_equal(y, full["y_end"])  # every rank holds the whole ensemble, in global order, bit for bit
        s = np.load(tmp_path / f"s_{rank}.npy")
        assert s[0] == full["n_accept"].sum() and s[1] == full["n_reject"].sum() and s[2] == full["n_rhs"].sum()
        assert s[3] == 0 and s[4] == 11.0  # max over ranks of the per-rank kernel time
        rec = np.load(tmp_path / f"rec_{rank}.npz")  # the whole section-8e record, every trajectory, in global order
        for k in ("y_end", "t_end", "status", "n_accept", "n_reject", "n_rhs"):
            assert np.array_equal(rec[k], full[k]), k


def test_shard_arithmetic():
    for n in (0, 1, 7, 8, 9, 1 << 20):
        for world in (1, 2, 3, 8):
            sizes = [shard_size(n, r, world) for r in range(world)]
            assert sum(sizes) == n and max(sizes) - min(sizes) <= 1
            allidx = np.concatenate([shard_indices(n, r, world) for r in range(world)]) if n else np.array([])
            assert sorted(allidx.tolist()) == list(range(n))
            for r in range(world):
                assert len(shard_indices(n, r, world)) == sizes[r]
