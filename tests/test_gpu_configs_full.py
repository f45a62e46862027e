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


# ---------------------------------------------------------------- BASELINE.json's full sizes: size-independent properties
def _device_solve(torch, solver, y0, par, **kw):
    dev = "cuda:0"
    d_y0 = torch.from_numpy(np.ascontiguousarray(y0)).to(dev)
    d_par = torch.from_numpy(np.ascontiguousarray(par)).to(dev)
    out = solver.solve_ivp_ensemble_device(d_y0, d_par, **kw)
    torch.cuda.synchronize()
    return d_y0, d_par, out


@pytest.mark.timeout(900, method="thread")
def test_config3_full_size_properties(cuda, engine, oracle):
    """Config 3 at its full size (2^22 trajectories): every trajectory reaches t_end, a second run gives the same bits
    (idempotence: the result does not depend on which lane ran what), and a sample spread over the mu-sweep is inside
    the band against the oracle."""
    torch = cuda
    w = E.VDP
    n = w["n"]
    assert n == 1 << 22
    y0, mu = E.vdp_problem(np.arange(n), n)
    s = make_solver(engine, "RK23", 2, rhs="vdp", dt_min=w["dt_min"], dt_max=w["dt_max"], tol=w["tol"],
                    t_start=w["t_start"], t_end=w["t_end"])
    d_y0, d_mu, a = _device_solve(torch, s, y0, mu)
    assert int((a["status"] != 0).sum().item()) == 0
    assert bool((a["t_end"] == w["t_end"]).all().item())
    ya = a["y_end"].clone()
    acc = a["n_accept"].clone()
    b = s.solve_ivp_ensemble_device(d_y0, d_mu)
    torch.cuda.synchronize()
    assert torch.equal(ya.view(torch.int64), b["y_end"].view(torch.int64)) and torch.equal(acc, b["n_accept"])
    sel = np.linspace(0, n - 1, 2048).astype(np.int64)
    ref = oracle.solve_ensemble(_abi.RK23, "vdp", y0[:, sel], mu[:, sel], dt_min=w["dt_min"], dt_max=w["dt_max"],
                                tol=w["tol"], t_start=w["t_start"], t_end=w["t_end"])
    g = ya[:, torch.from_numpy(sel).to(ya.device)].cpu().numpy()
    assert rel_err(g, ref["y_end"]).max() <= band(w["tol"])
    a64 = acc.to(torch.int64)
    assert int(a64.max().item()) > 6 * int(a64.min().item())


@pytest.mark.timeout(900, method="thread")
def test_config5_full_size_properties(cuda, engine, oracle):
    """Config 5 at its full size (2^20 Robertson trajectories, BDF6 with the batched Newton LU): all reach t_end, mass is
    conserved (y1 + y2 + y3 = 1 is an invariant of the kinetics and of every linear multistep method up to rounding),
    bitwise idempotent, sample inside the band."""
    torch = cuda
    w = E.ROBERTSON
    n = w["n"]
    assert n == 1 << 20
    y0, k = E.robertson_problem(np.arange(n))
    s = make_solver(engine, "BDF6", 3, rhs="robertson", flags=_abi.FLAG_BDF_NEWTON, dt_min=w["dt_min"], dt_max=w["dt_max"],
                    tol=w["tol"], t_start=w["t_start"], t_end=w["t_end"])
    d_y0, d_k, a = _device_solve(torch, s, y0, k)
    assert int((a["status"] != 0).sum().item()) == 0
    assert float((a["t_end"] - w["t_end"]).abs().max().item()) <= 1e-12
    mass = a["y_end"].sum(dim=0)
    assert float((mass - 1.0).abs().max().item()) < 1e-9
    ya = a["y_end"].clone()
    b = s.solve_ivp_ensemble_device(d_y0, d_k)
    torch.cuda.synchronize()
    assert torch.equal(ya.view(torch.int64), b["y_end"].view(torch.int64))
    sel = np.linspace(0, n - 1, 1024).astype(np.int64)
    ref = oracle.solve_ensemble(_abi.BDF6, "robertson", y0[:, sel], k[:, sel], bdf_newton=True, dt_min=w["dt_min"],
                                dt_max=w["dt_max"], tol=w["tol"], t_start=w["t_start"], t_end=w["t_end"])
    g = ya[:, torch.from_numpy(sel).to(ya.device)].cpu().numpy()
    assert rel_err(g, ref["y_end"]).max() <= band(w["tol"])


@pytest.mark.timeout(900, method="thread")
def test_config4_full_size_properties(cuda, engine, oracle):
    """Config 4 at its full size (2^18 trajectories, 2 GiB of per-trajectory matrices, dense output of capacity 256): all
    reach t_end, every stored path is as long as its step count and strictly increasing in time up to t_end, the end of
    the path is the final state, a sample agrees with the matrix exponential, and the events query over the 8.5 GB of
    264-byte records (TMA-staged kernel + deferred location) agrees with the one-lane-per-record kernel bit for bit."""
    from scipy.linalg import expm
    torch = cuda
    w = E.LINEAR32
    n = w["n"]
    assert n == 1 << 18
    chunks = [E.linear32_problem(np.arange(c, min(c + (1 << 15), n))) for c in range(0, n, 1 << 15)]
    y0 = np.concatenate([c[0] for c in chunks], axis=1)
    A = np.concatenate([c[1] for c in chunks], axis=0).reshape(n, 1024)
    del chunks
    cap = w["history_capacity"]
    s = make_solver(engine, "RK45", 32, rhs="linear32", dt_min=w["dt_min"], dt_max=w["dt_max"], tol=w["tol"],
                    t_start=w["t_start"], t_end=w["t_end"], history=cap)
    d_y0, d_A, a = _device_solve(torch, s, y0, A, params_aos=True)
    assert int((a["status"] != 0).sum().item()) == 0 and bool((a["t_end"] == w["t_end"]).all().item())
    acc = a["n_accept"].to(torch.int64)
    assert torch.equal(a["hist_len"].to(torch.int64), acc) and int(acc.max().item()) <= cap
    ht = a["hist_t"]
    idx = torch.arange(cap, device=ht.device)[None, :]
    valid = idx < acc[:, None]
    inc = (ht[:, 1:] > ht[:, :-1]) | ~valid[:, 1:]
    assert bool(inc.all().item())
    last = torch.gather(a["hist"], 1, (acc - 1).clamp(min=0)[:, None, None].expand(-1, 1, 33))[:, 0]
    assert bool((last[:, 0] == w["t_end"]).all().item())
    assert torch.equal(last[:, 1:].t().contiguous().view(torch.int64), a["y_end"].view(torch.int64))
    sel = np.linspace(0, n - 1, 48).astype(np.int64)
    ye = a["y_end"][:, torch.from_numpy(sel).to(ht.device)].cpu().numpy()
    for j, i in enumerate(sel):
        want = expm(A[i].reshape(32, 32) * (w["t_end"] - w["t_start"])) @ y0[:, i]
        assert np.linalg.norm(ye[:, j] - want) <= 1e-6 * max(np.linalg.norm(want), 1e-3), (i, np.linalg.norm(ye[:, j] - want))
    wv = np.zeros(32)
    wv[0] = 1.0
    ev1, c1 = s.locate_events_device(d_y0, d_A, a, wv, 0.0, 0, 4, params_aos=True)
    torch.cuda.synchronize()
    os.environ["BACON_EV_NO_TMA"] = "1"  # the round-1 kernel (one lane per record, located inside the stream)
    try:
        ev2, c2 = s.locate_events_device(d_y0, d_A, a, wv, 0.0, 0, 4, params_aos=True)
        torch.cuda.synchronize()
    finally:
        del os.environ["BACON_EV_NO_TMA"]
    assert torch.equal(c1, c2) and torch.equal(ev1.view(torch.int64), ev2.view(torch.int64))
    assert int(c1.sum().item()) > n // 2
    k = torch.minimum(c1.to(torch.int64), torch.tensor(4, device=c1.device))
    t_ev = ev1[:, :, 0]
    ok = (torch.arange(4, device=c1.device)[None, :] >= k[:, None]) | ((t_ev > w["t_start"]) & (t_ev <= w["t_end"]))
    assert bool(ok.all().item())
    g_ev = ev1[:, :, 1]  # y[0] at the event: on the surface
    on = (torch.arange(4, device=c1.device)[None, :] >= k[:, None]) | (g_ev.abs() < 1e-9)
    assert bool(on.all().item())
