#!/usr/bin/env python
"""tools/multi_entry_bench.py — the in-library multi-device entry point (bacon_ivp_solve_ensemble_multi: ONE process, host
buffers dealt i mod G over G devices, one stream per device, pinned staging packed / scattered on host threads) on the
headline ensemble: wall clock around the blocking call, per G.  One JSON line."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import bacon_b200 as B  # noqa: E402
from bacon_b200 import ensembles as E  # noqa: E402
import torch  # noqa: E402


def main():
    n = 1 << 20
    w = E.LORENZ
    y0 = B.pinned_empty((3, n))
    y0[:] = E.lorenz_y0(np.arange(n))
    p = np.array(w["params"])
    s = (B.RK45.new(3).with_dt_min(w["dt_min"]).with_dt_max(w["dt_max"]).with_tolerance(w["tol"])
         .with_start(w["t_start"]).with_end(w["t_end"]).with_derivative("lorenz"))
    out = {"workload": f"2^20 Lorenz RK45 through bacon_ivp_solve_ensemble_multi, host buffers (pinned)", "gpus_visible": torch.cuda.device_count(), "runs": []}
    ref = None
    for g in (1, 2, 4, 8):
        if g > torch.cuda.device_count():
            break
        for _ in range(2):  # warm-up: contexts, staging buffers
            r = s.solve_ivp_ensemble(y0, p, shared_params=True, n_gpus=g, zero_copy=False)
        ts = []
        for _ in range(5):
            t0 = time.perf_counter()
            r = s.solve_ivp_ensemble(y0, p, shared_params=True, n_gpus=g, zero_copy=False)
            ts.append(time.perf_counter() - t0)
        if ref is None:
            ref = r.y_end.copy()
        same = bool(np.array_equal(ref, r.y_end))
        acc = float(r.n_accept.sum())
        ms = 1e3 * float(np.median(ts))
        out["runs"].append({"n_gpus": g, "ms_per_call_median": ms, "steps_per_s": acc / (ms * 1e-3), "kernel_ms_max": r.launch["kernel_ms"],
                            "h2d_ms": r.launch["h2d_ms"], "d2h_ms": r.launch["d2h_ms"], "bit_identical_with_1_gpu": same,
                            "all_ok": bool((r.status == 0).all())})
    base = out["runs"][0]["ms_per_call_median"]
    for run in out["runs"]:
        run["speedup_vs_1"] = base / run["ms_per_call_median"]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
