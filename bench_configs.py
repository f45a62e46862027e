#!/usr/bin/env python
"""bench_configs.py — the other BASELINE.json configs (parity-test cases, not the headline bench line):

  2  (dense-output variant of the headline) Lorenz-63 RK45, 2^18 trajectories, EVERY accepted point to HBM
  3  Van der Pol mu-sweep, 2^22 trajectories, RK23 ("the second adaptive RK"), tol 1e-10
  4  32-dim linear ODE y' = A y, 2^18 trajectories with per-trajectory A, RK45, dense output to HBM (capacity 256)
  5  Robertson kinetics, 2^20 trajectories, BDF6 with the batched in-register 3x3 Newton LU (and Broyden), tol 1e-6

  python bench_configs.py --config 3 [--scale 0.25] [--steps 3]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
      bench_configs.py --config 5 --gpus 8        # the config's ensemble sharded i mod 8 over 8 B200 (strong scaling)

One JSON line per run: accepted trajectory-steps/s (device-resident, CUDA events on the launch stream), algorithmic
FLOP/s against the FP64 peak measured in the same run, dense-output GB/s against MEASURED_PEAKS.json's HBM figure,
and the CPU oracle on a bounded sub-sample of the same seeded ensemble.

--gpus N (under torchrun, one rank per GPU): the config's GLOBAL ensemble is dealt trajectory i -> rank i mod N
(bacon_b200/shard.py; the mu-sweep of config 3 stays balanced that way), every rank integrates its shard, and a pass ends
with the NCCL all-gather of the whole per-trajectory record into global order plus the all-reduce of the counters — both
inside the timed region.  Time = max over ranks (CUDA events), value = all ranks' accepted steps / that time.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def hbm_peak():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"], "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, required=True, choices=[2, 3, 4, 5])
    ap.add_argument("--scale", type=float, default=1.0, help="fraction of the config's trajectory count")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--broyden", action="store_true", help="config 5 with the reference's Broyden iteration instead of Newton+LU")
    ap.add_argument("--cpu-sample", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-hist", action="store_true", help="configs 2 and 4 without dense output (what the history costs)")
    ap.add_argument("--n", type=int, default=0, help="override the trajectory count")
    ap.add_argument("--paths", action="store_true", help="configs 2 and 4: also time the path queries (events, sampling) on the stored history")
    ap.add_argument("--gpus", type=int, default=1, help="ranks (under torchrun): the config's ensemble sharded i mod N")
    args = ap.parse_args()

    import torch
    import torch.distributed as dist

    import bacon_b200 as B
    from bacon_b200 import _abi, ensembles as E, shard

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus:
        sys.exit(f"--gpus {args.gpus} needs {args.gpus} ranks (torchrun --nproc-per-node {args.gpus}); WORLD_SIZE is {world}")
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    kw_dev = {}
    if args.config == 2:
        w = dict(E.LORENZ, n=args.n or (1 << 18))
        n_glob = max(1024, int(w["n"] * args.scale))
        mine = shard.shard_indices(n_glob, rank, world)
        n = mine.size
        y0 = E.lorenz_y0(mine)
        par = np.tile(np.array(w["params"])[:, None], (1, n))
        make = B.RungeKutta45
        flags, hist = 0, 4608
    elif args.config == 3:
        w = dict(E.VDP)
        n_glob = max(1024, int(w["n"] * args.scale))
        mine = shard.shard_indices(n_glob, rank, world)
        n = mine.size
        y0, par = E.vdp_problem(mine * (w["n"] // n_glob), w["n"])  # same mu range at any scale
        make = B.RungeKutta23
        flags, hist = 0, 0
    elif args.config == 4:
        w = dict(E.LINEAR32)
        n_glob = max(256, int(w["n"] * args.scale))
        mine = shard.shard_indices(n_glob, rank, world)
        n = mine.size
        y0, par = E.linear32_problem(mine)
        par = par.reshape(n, 1024)
        make = B.RungeKutta45
        flags, hist = 0, w["history_capacity"]
        kw_dev["params_aos"] = True
    else:
        w = dict(E.ROBERTSON)
        n_glob = max(1024, int(w["n"] * args.scale))
        mine = shard.shard_indices(n_glob, rank, world)
        n = mine.size
        y0, par = E.robertson_problem(mine)
        make = B.BDF6
        flags, hist = (0 if args.broyden else _abi.FLAG_BDF_NEWTON), 0
    if args.no_hist:
        hist = 0
    dim = w["dim"]
    s = (make.new(dim).with_minimum_dt(w["dt_min"]).with_maximum_dt(w["dt_max"]).with_tolerance(w["tol"])
         .with_initial_time(w["t_start"]).with_ending_time(w["t_end"]).with_derivative(w["rhs"]).with_flags(flags)
         .with_history(hist))

    d_y0 = torch.from_numpy(np.ascontiguousarray(y0)).to(dev)
    d_par = torch.from_numpy(np.ascontiguousarray(par)).to(dev)
    out = None
    for _ in range(max(args.warmup, 1)):
        out = s.solve_ivp_ensemble_device(d_y0, d_par, out=out, **kw_dev)
    torch.cuda.synchronize()
    peak = B.fp64_peak_tflops(1 << 15)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    gathered, totals = None, None
    if world > 1:  # (first collectives outside the timed region: NCCL sets its rings up on first use)
        gathered = shard.gather_records(out, n_glob, world)
        totals = shard.reduce_stats_device(out["n_accept"], out["n_reject"], out["n_rhs"], out["status"])
        torch.cuda.synchronize()
        dist.barrier()
    from bench import ClockSampler
    sampler = ClockSampler(dev.index or 0)
    sampler.start()
    for k in range(args.steps):
        flush.fill_(k)
        if world > 1:
            torch.cuda.synchronize()
            dist.barrier()
        ev[k][0].record()
        kev[k][0].record()
        out = s.solve_ivp_ensemble_device(d_y0, d_par, out=out, **kw_dev)
        kev[k][1].record()
        if world > 1:  # the pass ends with the gather of the whole record and the reduction of the counters
            gathered = shard.gather_records(out, n_glob, world)
            totals = shard.reduce_stats_device(out["n_accept"], out["n_reject"], out["n_rhs"], out["status"])
        ev[k][1].record()
    torch.cuda.synchronize()
    sampler.stop_flag = True
    sampler.join(timeout=1.0)
    ms = sum(a.elapsed_time(b) for a, b in ev) / args.steps
    kms = sum(a.elapsed_time(b) for a, b in kev) / args.steps
    if world > 1:
        t = torch.tensor([ms, kms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)  # max over ranks
        ms, kms = float(t[0]), float(t[1])
    st = out["status"].cpu().numpy()
    acc = out["n_accept"].cpu().numpy().astype(np.int64)
    rej = out["n_reject"].cpu().numpy().astype(np.int64)
    nrhs = out["n_rhs"].cpu().numpy().astype(np.int64)
    ok_status = (0, _abi.E_HISTORY_OVERFLOW) if hist else (0,)
    n_bad = int((~np.isin(st, ok_status)).sum())
    launch = B.last_launch()
    peak_all = peak
    if world > 1:  # whole-job totals; every rank's DFMA peak (they agree to a fraction of a percent)
        acc_all, rej_all, nrhs_all, bad_all = (float(v) for v in totals.cpu().tolist())
        g_acc = gathered["n_accept"].to(torch.int64)
        assert int(g_acc.sum().item()) == int(acc_all) and gathered["y_end"].shape == (w["dim"], n_glob)
        ok = torch.isin(gathered["status"], torch.tensor(list((0, _abi.E_HISTORY_OVERFLOW) if hist else (0,)), device=dev))
        n_bad_all = int((~ok).sum().item())
        pk = torch.tensor([peak], dtype=torch.float64, device=dev)
        dist.all_reduce(pk, op=dist.ReduceOp.SUM)
        peak_all = float(pk[0])
        pts_t = torch.tensor([float(np.minimum(acc, hist).sum()) if hist else 0.0], dtype=torch.float64, device=dev)
        dist.all_reduce(pts_t, op=dist.ReduceOp.SUM)
        pts_all = float(pts_t[0])
    if rank != 0:
        dist.destroy_process_group()
        return

    line = {"config": args.config, "workload": f"{w['rhs']} x {n_glob} trajectories, {w['method']}, tol {w['tol']}, "
            f"t in [{w['t_start']},{w['t_end']}], dt in [{w['dt_min']},{w['dt_max']}]"
            + (f", dense output capacity {hist}" if hist else "") + (", Newton+LU" if flags & _abi.FLAG_BDF_NEWTON else ""),
            "metric": "accepted f64 trajectory-steps/sec", "unit": "trajectory-steps/s", "n_gpus": world,
            "scaling": "strong", "trajectories_per_gpu": n, "n": n_glob, "dtype": "f64", "data": "synthetic",
            "grid": launch["grid"], "block": launch["block"], "regs_per_thread": launch["regs_per_thread"],
            "clocks": sampler.summary()}
    if world > 1:
        tot_acc, tot_rej, tot_rhs = acc_all, rej_all, nrhs_all
        line.update({"value": acc_all / (ms * 1e-3), "ms_per_pass": ms, "kernel_ms_max_over_ranks": kms,
                     "accepted": int(acc_all), "rejected": int(rej_all), "n_rhs": int(nrhs_all), "failed": n_bad_all,
                     "accept_min_max": [int(g_acc.min().item()), int(g_acc.max().item())],
                     "parallelism": f"trajectory-sharded x{world} (i mod N); NCCL all-gather of (y_end, t_end, status, counters) "
                                    "into global order + all-reduce of the totals inside every timed pass"})
    else:
        tot_acc, tot_rej, tot_rhs = float(acc.sum()), float(rej.sum()), float(nrhs.sum())
        line.update({"value": float(acc.sum()) / (ms * 1e-3), "ms_per_pass": ms, "accepted": int(acc.sum()),
                     "rejected": int(rej.sum()), "n_rhs": int(nrhs.sum()), "failed": n_bad,
                     "accept_min_max": [int(acc.min()), int(acc.max())]})
    if args.config in (2, 3, 4):
        fl = E.rk_flops(w["method"], dim, E.F_RHS[w["rhs"]], tot_acc + tot_rej, tot_acc)
    else:  # BDF: event-counted: RHS evaluations dominate; 15*D per g-evaluation on top
        fl = tot_rhs * (E.F_RHS["robertson"] + 15 * dim)
    tf = fl / (ms * 1e-3) / 1e12  # whole job, against the sum of the ranks' measured peaks
    line["roofline"] = {"bound": "fp64", "achieved": tf, "peak": peak_all, "unit": "TFLOP/s", "frac": tf / peak_all,
                        "peak_source": "DFMA loop measured in this run" + (" on every rank (summed)" if world > 1 else "")}
    if hist:
        pts = pts_all if world > 1 else np.minimum(acc, hist).sum()
        gb = float(pts) * 8 * (1 + dim) / 1e9
        hp, src = hbm_peak()
        hp *= world
        line["dense_output"] = {"bytes_per_accepted_step": 8 * (1 + dim), "GB_written": gb, "achieved_GBs": gb / (ms * 1e-3),
                                "peak_GBs": hp, "frac": gb / (ms * 1e-3) / hp, "peak_source": src}
    if args.paths and hist and args.config in (2, 4):
        # Path queries on the history just written (path_query.cuh): both HBM-bound.  Algorithmic bytes:
        # events = every record once, 8(1 + D) per accepted step; sampling = two records in + D doubles out per sample.
        hp, src = hbm_peak()
        pts = float(np.minimum(acc, hist).sum())
        n_times = 64 if args.config == 2 else 16
        ev_w, ev_c = ([0.0, 0.0, 1.0], 27.0) if args.config == 2 else ([1.0] + [0.0] * (dim - 1), 0.0)
        d_times = torch.linspace(w["t_start"], w["t_end"], n_times, dtype=torch.float64, device=dev)
        samples = torch.empty((n, n_times, dim), dtype=torch.float64, device=dev)

        def timed(fn):
            fn()
            torch.cuda.synchronize()
            tt = []
            for k in range(max(args.steps, 3)):
                flush.fill_(k)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                r = fn()
                b.record()
                torch.cuda.synchronize()
                tt.append(a.elapsed_time(b))
            return sum(tt) / len(tt), r, B.last_launch()

        # "call_ms": CUDA events around the whole Python call (incl. zeroing the event buffers and the host-side call path,
        # during which the GPU idles); "ms": the library's own event pair around the kernel (bacon_ivp_last_launch)
        call_ev, (d_ev, d_cnt), ev_launch = timed(lambda: s.locate_events_device(d_y0, d_par, out, ev_w, ev_c, 0, 4, **kw_dev))
        call_sm, _, sm_launch = timed(lambda: s.sample_paths_device(d_y0, d_par, out, d_times, samples=samples, **kw_dev))
        ms_ev, ms_sm = ev_launch["kernel_ms"], sm_launch["kernel_ms"]
        ev_bytes = pts * 8 * (1 + dim)
        sm_bytes = float(n) * n_times * (2 * 8 * (1 + dim) + 8 * dim)
        line["path_queries"] = {
            "events": {"surface": ("z = 27 (Poincare section)" if args.config == 2 else "y[0] = 0") + ", both directions, capacity 4", "ms": ms_ev, "call_ms": call_ev,
                       "events_found": int(d_cnt.sum().item()), "algorithmic_GB": ev_bytes / 1e9,
                       "achieved_GBs": ev_bytes / 1e9 / (ms_ev * 1e-3), "peak_GBs": hp,
                       "frac": ev_bytes / 1e9 / (ms_ev * 1e-3) / hp, "regs_per_thread": ev_launch["regs_per_thread"],
                       "traffic_note": "config 2, ncu: dram bytes read = 30.08 GB = the algorithmic bytes (every record once)"},
            "sampling": {"n_times": n_times, "ms": ms_sm, "call_ms": call_sm,
                         "traffic_note": "config 2, ncu: 5.1 GB of DRAM reads at 5.2 TB/s (bisection in round 1: 9.9 GB): every probe of a knot's time "
                                         "moves a 128-byte line for 8 bytes; DRAM-bound on its actual traffic (profiles/r04_sampling_search.md)", "samples_per_s": float(n) * n_times / (ms_sm * 1e-3),
                         "algorithmic_GB": sm_bytes / 1e9, "achieved_GBs": sm_bytes / 1e9 / (ms_sm * 1e-3), "peak_GBs": hp,
                         "frac": sm_bytes / 1e9 / (ms_sm * 1e-3) / hp, "regs_per_thread": sm_launch["regs_per_thread"]},
            "peak_source": src}
    if not args.no_cpu_baseline:
        from oracle import oracle as O
        O.build()
        cores = os.cpu_count() or 1
        m = args.cpu_sample or {2: 8192 * cores, 3: 1024 * cores, 4: 4096 * cores, 5: 8192 * cores}[args.config]  # ~10 s of CPU work
        m = min(m, n)
        sel = np.linspace(0, n - 1, m).astype(np.int64)  # spread over the ensemble (the mu-sweep is ordered)
        yy = np.ascontiguousarray(y0[:, sel])
        pp = np.ascontiguousarray(par[sel]) if args.config == 4 else np.ascontiguousarray(par[:, sel])
        t0 = time.perf_counter()
        r = O.solve_ensemble({"RK23": _abi.RK23, "RK45": _abi.RK45, "BDF6": _abi.BDF6}[w["method"]], w["rhs"], yy, pp,
                             params_aos=(args.config == 4), dt_min=w["dt_min"], dt_max=w["dt_max"], tol=w["tol"],
                             t_start=w["t_start"], t_end=w["t_end"], bdf_newton=bool(flags & _abi.FLAG_BDF_NEWTON))
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": float(r["n_accept"].sum()) / dt, "unit": "trajectory-steps/s", "cores": cores,
                                "kind": "port", "sample": f"{m} trajectories spread over {'rank 0 shard of ' if world > 1 else ''}the same seeded ensemble ({dt:.1f} s)"}
        # parity of the sample, side by side
        g = out["y_end"][:, torch.from_numpy(sel).to(dev)].cpu().numpy()
        err = np.sqrt(((g - r["y_end"]) ** 2).sum(0)) / np.maximum(np.sqrt((r["y_end"] ** 2).sum(0)), 1e-300)
        line["parity_sample"] = {"worst_rel_err": float(err.max()), "band": max(10 * w["tol"], 1e-12),
                                 "accepted_gpu": int(acc[sel].sum()), "accepted_cpu": int(r["n_accept"].sum()),
                                 "rejected_gpu": int(rej[sel].sum()), "rejected_cpu": int(r["n_reject"].sum())}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
