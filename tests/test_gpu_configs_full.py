"""Full-horizon parity of BASELINE.json's configs 3 and 5 against the CPU oracle, through the C ABI, plus the N > 1
collective path under NCCL.  (The other GPU tests use shortened horizons so that the oracle finishes in a second;
these run the horizons and tolerances the configs name, on ensembles spread over the whole parameter range.)

  config 3  Van der Pol mu-sweep, "the second adaptive RK" (RK23, src/ivp/rk.rs:563-621), tol 1e-10, t in [0, 0.25]
  config 5  Robertson kinetics, BDF6 (src/ivp/bdf.rs:495-634), tol 1e-6, t in [0, 0.5]: batched Newton + in-register LU
            and the reference's own Broyden iteration
"""
import os
import socket

import numpy as np
import pytest

from bacon_b200 import _abi, ensembles as E
from parity import band, make_solver, rel_err, run_both

pytestmark = pytest.mark.gpu


@pytest.mark.timeout(600, method="thread")
def test_config3_vdp_mu_sweep_full_horizon(cuda, engine, oracle):
    n = 4096
    w = E.VDP
    y0, mu = E.vdp_problem(np.arange(n) * (w["n"] // n), w["n"])  # mu from 0.1 to 5 over the ensemble, as in config 3
    cfg = dict(dt_min=w["dt_min"], dt_max=w["dt_max"], tol=w["tol"], t_start=w["t_start"], t_end=w["t_end"])
    assert cfg["t_end"] == 0.25 and cfg["tol"] == 1e-10
    gpu, ref = run_both(engine, oracle, "RK23", "vdp", y0, mu, **cfg)
    assert (gpu.status == _abi.OK).all() and (ref["status"] == _abi.OK).all()
    err = rel_err(gpu.y_end, ref["y_end"])
    assert err.max() <= band(cfg["tol"]), f"worst relative error {err.max():.3e} > {band(cfg['tol']):.1e}"
    np.testing.assert_array_equal(gpu.t_end, ref["t_end"])
    # accepted / rejected step counts side by side (north_star): a handful of borderline decisions per trajectory
    d_acc = np.abs(gpu.n_accept.astype(np.int64) - ref["n_accept"].astype(np.int64))
    d_rej = np.abs(gpu.n_reject.astype(np.int64) - ref["n_reject"].astype(np.int64))
    assert d_acc.max() <= 8 and d_rej.max() <= 8, (d_acc.max(), d_rej.max())
    assert abs(int(gpu.n_accept.sum()) - int(ref["n_accept"].sum())) <= 1e-5 * ref["n_accept"].sum()
    assert gpu.n_accept.max() > 6 * gpu.n_accept.min()  # the 8x spread of step counts the config is there for
    print(f"config 3, {n} trajectories: worst rel err {err.max():.2e} (band {band(cfg['tol']):.0e}), accepted "
          f"{int(gpu.n_accept.sum())} vs {int(ref['n_accept'].sum())}, rejected {int(gpu.n_reject.sum())} vs {int(ref['n_reject'].sum())}")


@pytest.mark.timeout(900, method="thread")
def test_config5_robertson_full_horizon_newton_and_broyden(cuda, engine, oracle):
    n = 4096
    w = E.ROBERTSON
    y0, k = E.robertson_problem(np.arange(n) * (w["n"] // n))  # k perturbed +-10 % per trajectory, as in config 5
    cfg = dict(dt_min=w["dt_min"], dt_max=w["dt_max"], tol=w["tol"], t_start=w["t_start"], t_end=w["t_end"])
    assert cfg["t_end"] == 0.5 and cfg["tol"] == 1e-6
    # north-star item 4: batched Newton with the in-register 3x3 LU, against the oracle's Newton statement
    s = make_solver(engine, "BDF6", 3, rhs="robertson", flags=_abi.FLAG_BDF_NEWTON, **cfg)
    nw = s.solve_ivp_ensemble(y0, k)
    ref_nw = oracle.solve_ensemble(_abi.BDF6, "robertson", y0, k, bdf_newton=True, **cfg)
    assert (nw.status == _abi.OK).all() and (ref_nw["status"] == _abi.OK).all()
    e_nw = rel_err(nw.y_end, ref_nw["y_end"])
    assert e_nw.max() <= band(cfg["tol"]), e_nw.max()
    np.testing.assert_array_equal(nw.n_accept, ref_nw["n_accept"])
    np.testing.assert_array_equal(nw.n_reject, ref_nw["n_reject"])
    # the reference's own iteration (Broyden, bdf.rs:414-475), fast build, against the oracle's statement of it
    br, ref_br = run_both(engine, oracle, "BDF6", "robertson", y0, k, **cfg)
    assert (br.status == _abi.OK).all() and (ref_br["status"] == _abi.OK).all()
    e_br = rel_err(br.y_end, ref_br["y_end"])
    assert e_br.max() <= band(cfg["tol"]), e_br.max()
    assert np.abs(br.n_accept.astype(int) - ref_br["n_accept"].astype(int)).max() <= 16  # one restart block = 7 + 1 points
    # the two iterations agree with each other inside the band, and the kinetics conserve mass
    assert rel_err(nw.y_end, ref_br["y_end"]).max() <= band(cfg["tol"])
    assert np.abs(nw.y_end.sum(0) - 1.0).max() < 1e-9 and np.abs(br.y_end.sum(0) - 1.0).max() < 1e-9
    print(f"config 5, {n} trajectories: Newton worst rel err {e_nw.max():.2e}, Broyden {e_br.max():.2e} (band "
          f"{band(cfg['tol']):.0e}); accepted {int(nw.n_accept.sum())} (Newton) / {int(br.n_accept.sum())} (Broyden)")


# ------------------------------------------------------------------ the N > 1 path under NCCL
def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _nccl_worker(rank, world, port, n_global, out_dir):
    import torch
    import torch.distributed as dist

    import bacon_b200 as B
    from bacon_b200.shard import gather_final_states, gather_records, reduce_stats, shard_indices
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    w = E.LORENZ
    idx = shard_indices(n_global, rank, world)  # trajectory i -> rank i mod N
    y0 = torch.from_numpy(E.lorenz_y0(idx)).to(dev)
    p = torch.tensor(w["params"], dtype=torch.float64, device=dev)
    s = (B.RK45.new(3).with_dt_min(w["dt_min"]).with_dt_max(w["dt_max"]).with_tolerance(w["tol"]).with_start(0.0)
         .with_end(0.5).with_derivative("lorenz"))
    out = s.solve_ivp_ensemble_device(y0, p, shared_params=True)
    rec = gather_records(out, n_global, world)
    y_only = gather_final_states(out["y_end"], n_global, world)
    stats = reduce_stats(out["n_accept"], out["n_reject"], out["n_rhs"], out["status"], kernel_ms=1.0 + rank)
    torch.cuda.synchronize()
    assert torch.equal(y_only, rec["y_end"])
    np.savez(os.path.join(out_dir, f"rec_{rank}.npz"), **{k: v.cpu().numpy() for k, v in rec.items()},
             stats=np.array([stats["n_accept"], stats["n_reject"], stats["n_rhs"], stats["n_failed"], stats["kernel_ms_max"]]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_global", [100_001, 4096])
def test_nccl_gather_records_matches_single_gpu(cuda, engine, tmp_path, n_global):
    """Two ranks, one GPU each: every rank integrates its shard (i mod 2) and the collectives of SURVEY.md section 8e —
    all-gather of {y_end, t_end, status, n_accept, n_reject, n_rhs} into global order, all-reduce of the totals — must
    reproduce the single-GPU solve of the whole ensemble bit for bit (ragged shards included)."""
    if cuda.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    world = 2
    mp.spawn(_nccl_worker, args=(world, _free_port(), n_global, str(tmp_path)), nprocs=world, join=True)
    w = E.LORENZ
    s = make_solver(engine, "RK45", 3, rhs="lorenz", dt_min=w["dt_min"], dt_max=w["dt_max"], tol=w["tol"], t_start=0.0, t_end=0.5)
    full = s.solve_ivp_ensemble(E.lorenz_y0(np.arange(n_global)), np.array(w["params"]), shared_params=True)
    for rank in range(world):
        rec = np.load(tmp_path / f"rec_{rank}.npz")
        for k in ("y_end", "t_end"):
            assert np.array_equal(rec[k].view(np.uint64), getattr(full, k).view(np.uint64)), k
        for k in ("status", "n_accept", "n_reject", "n_rhs"):
            np.testing.assert_array_equal(rec[k], getattr(full, k), err_msg=k)
        st = rec["stats"]
        assert st[0] == full.n_accept.sum() and st[1] == full.n_reject.sum() and st[2] == full.n_rhs.sum()
        assert st[3] == 0 and st[4] == 2.0
