"""tools/small_n_probe.py — the headline workload (Lorenz RK45, T = 5) on ensembles that fill m warps per sub-partition
from the start (n = SMs x 128 x m): the regime of the END of every launch."""
import numpy as np
import torch
import bacon_b200 as B
from bacon_b200 import ensembles as E
w = E.LORENZ
sm = torch.cuda.get_device_properties(0).multi_processor_count
p = torch.tensor(w["params"], dtype=torch.float64).cuda()
for m in (1, 2, 3, 4, 6):
    n = sm * 128 * m
    y0 = torch.from_numpy(E.lorenz_y0(np.arange(n))).cuda()
    s = (B.RK45.new(3).with_dt_min(w["dt_min"]).with_dt_max(w["dt_max"]).with_tolerance(w["tol"]).with_start(0.0)
         .with_end(w["t_end"]).with_derivative("lorenz"))
    best = 1e9
    for _ in range(4):
        out = s.solve_ivp_ensemble_device(y0, p, shared_params=True)
        torch.cuda.synchronize()
        best = min(best, B.last_launch()["kernel_ms"])
    att = (out["n_accept"] + out["n_reject"]).cpu().numpy()
    print(f"m={m} n={n:7d} kernel {best:7.3f} ms   attempts mean {att.mean():7.1f} max {att.max()}   max x tau(m) = "
          f"{att.max() * {1: .272, 2: .388, 3: .474, 4: .603, 6: .887}[m] / 1000:6.3f} ms")
