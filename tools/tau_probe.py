"""tools/tau_probe.py — time per attempt of one warp as a function of the warps resident on its SM sub-partition.
Every trajectory is cut by the attempt cap after exactly CAP attempts (status MaxAttempts), the ensemble holds m warps
per sub-partition (n = 148 SMs x 4 x 32 x m, static first deal, no refills): kernel time / CAP = tau(m)."""
import sys
import numpy as np
import torch
import bacon_b200 as B
from bacon_b200 import ensembles as E

CAP = 3000
w = E.LORENZ
sm = torch.cuda.get_device_properties(0).multi_processor_count
p = torch.tensor(w["params"], dtype=torch.float64).cuda()
for m in (1, 2, 3, 4, 5, 6, 7):
    n = sm * 128 * m
    y0 = torch.from_numpy(E.lorenz_y0(np.arange(n))).cuda()
    s = (B.RK45.new(3).with_dt_min(w["dt_min"]).with_dt_max(w["dt_max"]).with_tolerance(w["tol"]).with_start(0.0)
         .with_end(1e9).with_derivative("lorenz").with_max_attempts(CAP))
    best = 1e9
    for _ in range(4):
        out = s.solve_ivp_ensemble_device(y0, p, shared_params=True)
        torch.cuda.synchronize()
        best = min(best, B.last_launch()["kernel_ms"])
    st = out["status"].cpu().numpy()
    print(f"m={m} warps/SMSP  n={n:7d}  kernel {best:8.3f} ms  tau = {1e3 * best / CAP:6.3f} us/attempt/warp  "
          f"-> {m * 1e-3 * CAP / best / (1 / 0.1313):5.1%} of the FP64 pipe (129 FP64 instr x 2 clk)   status {np.unique(st)}  grid {B.last_launch()['grid']}x{B.last_launch()['block']}")
