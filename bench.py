#!/usr/bin/env python
"""bench.py — headline benchmark: accepted f64 trajectory-steps/s of the RK45 Lorenz-63 ensemble
(BASELINE.json configs[1]: 2^20 trajectories, tol 1e-8, random initial conditions; T=5).

  python bench.py --gpus N --steps K --warmup W            # our arm (CUDA, sm_100a)
  python bench.py --impl reference --gpus N ...            # CPU arm: the oracle port on the host cores

A "step" is one pass of the hot path over the whole ensemble.  N>1: one rank per GPU (torchrun), rank r
owns trajectories r, r+N, ... — of the 2^20-trajectory ensemble the metric names (strong scaling, the
headline), and, measured in the same run, of an ensemble of N*2^20 (weak scaling).  No data-path
collective; NCCL gathers the per-trajectory records and sums the statistics at the end of each step.
Prints ONE JSON line.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "accepted f64 trajectory-steps/sec (1M Lorenz RK45)"
UNIT = "trajectory-steps/s"
FP64_NOMINAL_TFLOPS = 148 * 64 * 2 * 1.965e9 / 1e12  # 37.2: 64 DFMA/clk/SM at clocks.max.sm


N_GLOBAL = 1 << 20  # the 1M-trajectory ensemble the metric is quoted on


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="strong (default): the 2^20-trajectory ensemble is dealt i mod N over the N GPUs (what the metric "
                         "names); weak: 2^20 trajectories PER GPU.  The line carries both; this picks the headline `value`.")
    ap.add_argument("--n", type=int, default=0, help="global trajectories (default 2^20)")
    ap.add_argument("--t-end", type=float, default=0.0)
    ap.add_argument("--cpu-sample", type=int, default=0, help="trajectories in the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-scaling", action="store_true", help="skip the second (non-headline) scaling measurement")
    ap.add_argument("--e2e-mode", default="auto", choices=["auto", "staged", "zero_copy"],
                    help="host-buffer C-ABI call: staged = async H2D, kernel, async D2H; zero_copy = the kernel reads/"
                         "writes the pinned host buffers itself; auto = zero_copy on one GPU, staged when several ranks "
                         "share the host (8 kernels issuing per-trajectory PCIe writes at once: 45 ms against 31.5)")
    return ap.parse_args()


def workload(args):
    from bacon_b200 import ensembles as E
    w = dict(E.LORENZ)
    w["n"] = args.n or N_GLOBAL
    if args.t_end:
        w["t_end"] = args.t_end
    return w


def config_dict(w, n_gpus, scaling):
    """The same dict from both arms (the driver compares them): what is computed, not how."""
    per_gpu = w["n"] // n_gpus if scaling == "strong" else w["n"]
    return {"workload": f"BASELINE configs[1]: RK45 ensemble, {w['n']} Lorenz-63 trajectories "
                        f"(sigma=10, rho=28, beta=8/3), y0~U([-15,15]x[-20,20]x[5,40]) SplitMix64 seed 0x5EED0001, "
                        f"t in [0,{w['t_end']}], tol {w['tol']}, dt in [{w['dt_min']},{w['dt_max']}], final state only",
            "scaling": scaling, "trajectories_per_gpu": per_gpu, "global_trajectories": per_gpu * n_gpus, "method": "RK45",
            "rhs": "lorenz", "semantics": "REF_CORRECTED", "parallelism": f"trajectory-sharded x{n_gpus} (i mod N)"}


# --------------------------------------------------------------------------- CPU arm
def cpu_run(w, n_sample, threads=0):
    """The oracle port (oracle/liboracle.so, OpenMP over trajectories = the reference's rayon arm) on the first
    n_sample trajectories of the same seeded ensemble.  Returns (accepted steps, seconds, cores)."""
    from bacon_b200 import _abi, ensembles as E
    from oracle import oracle as O
    O.build()
    y0 = E.lorenz_y0(np.arange(n_sample))
    p = np.array(w["params"])
    cores = O.max_threads() if threads <= 0 else threads
    t0 = time.perf_counter()
    r = O.solve_ensemble(_abi.RK45, "lorenz", y0, p, shared_params=True, dt_min=w["dt_min"], dt_max=w["dt_max"],
                         tol=w["tol"], t_start=w["t_start"], t_end=w["t_end"], n_threads=threads)
    dt = time.perf_counter() - t0
    assert (r["status"] == 0).all()
    return int(r["n_accept"].sum()), dt, cores


def host_cores():
    """Cores this process may run on (affinity / cgroup aware), not OMP_NUM_THREADS: torchrun sets that to 1."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return os.cpu_count() or 1


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    w = workload(args)
    cores = host_cores()
    n_sample = args.cpu_sample or max(1024, min(w["n"], 4096 * cores))
    # (threads named explicitly: torchrun exports OMP_NUM_THREADS=1 to its workers, and this arm is meant to use every
    # host core, like the reference's rayon pool)
    for _ in range(min(args.warmup, 1)):
        cpu_run(w, max(256, n_sample // 16), threads=cores)
    steps_total, secs = 0, 0.0
    for _ in range(args.steps):
        s, dt, cores = cpu_run(w, n_sample, threads=cores)
        steps_total += s
        secs += dt
    v = steps_total / secs
    sample = f"first {n_sample} trajectories of the seeded ensemble per step, {args.steps} step(s)"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(w, args.gpus, args.scaling),
        "arm": {"note": "CPU arm: C++ restatement of src/ivp/rk.rs (oracle port; the Rust reference cannot be built in "
                        "this image), OpenMP over trajectories; a step = the sample below, not the whole ensemble"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))
    return 0


# --------------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        self.power = []
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.05)  # (an NVML query now and then costs the running kernel ~1.8 ms: tools/outlier.sh)

    def summary(self):
        med = float(np.median(self.samples)) if self.samples else None
        out = {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}
        if self.power:
            out["power_w_median"] = float(np.median(self.power))
        return out


# --------------------------------------------------------------------------- our arm
def ours(args):
    import torch
    import torch.distributed as dist

    import bacon_b200 as B
    from bacon_b200 import ensembles as E
    from bacon_b200.shard import gather_records, reduce_stats_device

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    w = workload(args)
    solver = (B.RK45.new(3).with_dt_min(w["dt_min"]).with_dt_max(w["dt_max"]).with_tolerance(w["tol"])
              .with_start(w["t_start"]).with_end(w["t_end"]).with_derivative("lorenz"))
    p_host = torch.tensor(w["params"], dtype=torch.float64).pin_memory()
    p = p_host.to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    e2e_mode = args.e2e_mode if args.e2e_mode != "auto" else ("zero_copy" if world == 1 else "staged")

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def measure(scaling, sampler=None):
        """K timed passes over this rank's shard of the ensemble: strong = N_GLOBAL dealt i mod N, weak = N_GLOBAL per
        GPU.  One pass = the hot path + the only collectives of the path (N > 1): all-gather of the per-trajectory
        records into global trajectory order, all-reduce of the counters."""
        n = w["n"] // world if scaling == "strong" else w["n"]
        n_glob = n * world
        idx = np.arange(n, dtype=np.uint64) * np.uint64(world) + np.uint64(rank)  # trajectory i -> rank i mod N
        y0_host = torch.from_numpy(E.lorenz_y0(idx)).pin_memory()
        y0 = y0_host.to(dev)
        # Two result buffers: the collectives of pass k (side stream) read buffer k & 1 while pass k + 1 (main stream)
        # writes the other.  The persistent kernel fills every SM, so the side stream's small kernels (packing, NCCL)
        # get their turn as the kernel's warps retire at the end of a pass: the hand-over of one pass overlaps the
        # thinning end of the next.  Timing: window k opens (event) before pass k is launched and closes after pass k
        # AND the collectives of pass k - 1 are done; a last window holds the collectives of the last pass.  The L2
        # flushes sit between the windows.  value = K passes / sum of the K + 1 windows.
        outs = [None, None]
        side = torch.cuda.Stream(device=dev)
        main = torch.cuda.current_stream(dev)
        state = {}

        def collect(out):
            state["stats"] = reduce_stats_device(out["n_accept"], out["n_reject"], out["n_rhs"], out["status"])
            state["rec"] = gather_records(out, n_glob, world)

        def run_passes(k_passes, timed):
            ev, kev = [], []
            done = [None, None]
            for k in range(k_passes + 1):
                if timed:
                    flush.fill_(k & 0xFF)  # L2 flush between timed windows (outside the event pairs)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(main)
                if k > 0:  # the collectives of pass k - 1, on the side stream, inside window k
                    side.wait_event(e0)
                    with torch.cuda.stream(side):
                        collect(outs[(k - 1) & 1])
                        done[(k - 1) & 1] = torch.cuda.Event()
                        done[(k - 1) & 1].record(side)
                if k < k_passes:
                    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    k0.record(main)
                    outs[k & 1] = solver.solve_ivp_ensemble_device(y0, p, shared_params=True, out=outs[k & 1])
                    k1.record(main)
                    kev.append((k0, k1))
                if k > 0:
                    main.wait_event(done[(k - 1) & 1])
                e1.record(main)
                ev.append((e0, e1))
            return ev, kev

        run_passes(max(args.warmup, 3), False)
        barrier()
        if sampler:
            sampler.start()
        barrier()
        t_wall0 = time.perf_counter()
        ev, kev = run_passes(args.steps, True)
        barrier()
        t_wall = time.perf_counter() - t_wall0
        if sampler:
            sampler.stop_flag = True
            sampler.join()
        dev_ms = sum(a.elapsed_time(b) for a, b in ev)    # all windows: K passes + their stats and collectives, this rank
        ker_ms = sum(a.elapsed_time(b) for a, b in kev)   # the ensemble kernel alone
        tmax = torch.tensor([dev_ms, ker_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)   # max over ranks
        dev_ms, ker_ms = float(tmax[0]), float(tmax[1])
        acc_total, rej_total, _, bad = (float(x) for x in state["stats"].cpu())  # global sums of the last pass (all passes identical)
        assert bad == 0, f"{bad} trajectories did not finish with status Ok"
        rec = state["rec"]
        assert rec["y_end"].shape == (3, n_glob) and int((rec["status"] != 0).sum()) == 0  # every rank holds every record
        launch = B.last_launch()

        # ---- end to end through the host-buffer C-ABI call: pinned host in, host out, copies inside the timed region
        y0_np, p_np = y0_host.numpy(), p_host.numpy()
        zc = e2e_mode == "zero_copy"
        for _ in range(2):
            r = solver.solve_ivp_ensemble(y0_np, p_np, shared_params=True, zero_copy=zc)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            r = solver.solve_ivp_ensemble(y0_np, p_np, shared_params=True, zero_copy=zc)
        torch.cuda.synchronize()
        t_e2e = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
        e2e_s = float(t_e2e[0])
        flops_step = E.rk_flops("RK45", 3, E.F_RHS["lorenz"], (acc_total + rej_total) / world, acc_total / world)
        return {
            "scaling": scaling, "trajectories_per_gpu": n, "global_trajectories": n_glob,
            "value": acc_total * args.steps / (dev_ms * 1e-3), "ms_per_step": dev_ms / args.steps,
            "kernel_ms_per_launch": ker_ms / args.steps, "kernel_ms_each_rank0": [round(a.elapsed_time(b), 3) for a, b in kev],
            "tflops_per_gpu": flops_step * args.steps / (ker_ms * 1e-3) / 1e12,  # kernel time = max over ranks
            "accepted": acc_total, "rejected": rej_total, "wall_s": t_wall, "launch": launch,
            "windows_ms_rank0": [round(a.elapsed_time(b), 3) for a, b in ev],
            "e2e": {"value": float(r.n_accept.sum()) * world * args.steps / e2e_s, "unit": UNIT,
                    "h2d_bytes_per_step": y0_np.nbytes + p_np.nbytes,
                    "d2h_bytes_per_step": sum(getattr(r, k).nbytes for k in ("y_end", "t_end", "dt_end", "status", "n_accept",
                                                                             "n_reject", "n_rhs")),
                    "ms_per_step": 1e3 * e2e_s / args.steps, "mode": e2e_mode,
                    "how": "bacon_ivp_solve_ensemble (C ABI, host buffers): pinned y0/params in, pinned result arrays "
                           "out, wall clock around the blocking calls"}}

    # warm up the device and measure the roofline denominator on this GPU
    solver.solve_ivp_ensemble_device(torch.from_numpy(E.lorenz_y0(np.arange(1 << 16))).to(dev), p, shared_params=True)
    barrier()
    fp64_peak = B.fp64_peak_tflops(1 << 15) if rank == 0 else None
    barrier()

    sampler = ClockSampler(local)
    if os.environ.get("BENCH_NO_CLOCKS"):  # (diagnosis: does the NVML sampling thread perturb the timed region?)
        sampler.nv = None
    head = measure(args.scaling, sampler)
    other_name = "weak" if args.scaling == "strong" else "strong"
    if world == 1:
        other = None  # one GPU: both scalings are the same run
    elif args.no_other_scaling:
        other = "skipped"
    else:
        other = measure(other_name)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant (only) kernel: FP64 vector pipe
    peak = fp64_peak if fp64_peak and fp64_peak > 0 else FP64_NOMINAL_TFLOPS
    n = head["trajectories_per_gpu"]
    launch = head["launch"]
    traffic, traffic_src = None, None
    try:  # DRAM bytes per launch from the committed `ncu --set full` capture of this command (profiles/)
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tj = json.load(f)
        if tj.get("trajectories") == n:
            traffic, traffic_src = tj["dram_bytes_per_launch"], tj["source"]
    except Exception:
        pass
    achieved = head["tflops_per_gpu"]
    roofline = {"bound": "fp64", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src,
                "algorithmic_hbm_bytes_per_launch": n * 80,  # 24 B in + 56 B out per trajectory: HBM is not the bound
                "peak_source": "measured on this GPU in this run: register-resident DFMA loop "
                "(bacon_fp64_peak_tflops); MEASURED_PEAKS.json has no FP64 entry",
                "nominal_peak": FP64_NOMINAL_TFLOPS, "frac_of_nominal": achieved / FP64_NOMINAL_TFLOPS,
                "flops_per_accepted_step": 230, "flops_per_attempt": 205,
                "kernel": "ensemble_kernel<RkFastStepper<RhsLorenz,TabRKF45>> (one launch per pass: refills, "
                          "end-of-ensemble regrouping and retirement inside the kernel)",
                "kernels_per_launch": launch["n_kernels"], "kernel_ms_per_launch": head["kernel_ms_per_launch"],
                "kernel_ms_each_rank0": head["kernel_ms_each_rank0"]}

    cpu_baseline = None
    if not args.no_cpu_baseline and world == 1:
        cores = host_cores()
        n_sample = args.cpu_sample or max(1024, min(n, 8192 * cores))
        s, dt, cores = cpu_run(w, n_sample, threads=cores)
        cpu_baseline = {"value": s / dt, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": f"first {n_sample} trajectories of the same seeded ensemble, one pass ({dt:.1f} s)"}

    def brief(m):
        if not isinstance(m, dict):
            return m
        return {k: m[k] for k in ("scaling", "trajectories_per_gpu", "global_trajectories", "value", "ms_per_step",
                                  "kernel_ms_per_launch", "tflops_per_gpu")} | {
            "frac_of_fp64_peak": m["tflops_per_gpu"] / peak, "e2e_value": m["e2e"]["value"], "e2e_ms_per_step": m["e2e"]["ms_per_step"],
            "grid": m["launch"]["grid"], "block": m["launch"]["block"]}

    line = {
        "metric": METRIC, "value": head["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": config_dict(w, world, args.scaling),
        "arm": {"l2": "256 MB buffer written between timed iterations", "grid": launch["grid"], "block": launch["block"],
                "regs_per_thread": launch["regs_per_thread"]},
        "accepted_steps_per_step": head["accepted"], "rejected_steps_per_step": head["rejected"], "wall_s": head["wall_s"],
        "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": head["e2e"],
        args.scaling: brief(head), other_name: brief(other) if other is not None else brief(head) | {"scaling": other_name, "note": "one GPU: the same run"},
        "gpu_launches": args.steps * world * launch["n_kernels"], "clocks": sampler.summary(),
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    a = parse()
    sys.exit(reference_arm(a) if a.impl == "reference" else ours(a))
