import sys, time, numpy as np
sys.path.insert(0, '.')
import bacon_b200 as B
from bacon_b200 import ensembles as E
import torch
w = dict(E.LORENZ); n = w["n"]
y0 = torch.from_numpy(E.lorenz_y0(np.arange(n, dtype=np.uint64))).pin_memory().numpy()
p = torch.tensor(w["params"], dtype=torch.float64).pin_memory().numpy()
s = (B.RK45.new(3).with_dt_min(w["dt_min"]).with_dt_max(w["dt_max"]).with_tolerance(w["tol"]).with_start(0.0).with_end(w["t_end"]).with_derivative("lorenz"))
for zc in (True, False):
    for _ in range(2): r = s.solve_ivp_ensemble(y0, p, shared_params=True, zero_copy=zc)
    ts = []
    for _ in range(5):
        t0 = time.perf_counter(); r = s.solve_ivp_ensemble(y0, p, shared_params=True, zero_copy=zc); ts.append((time.perf_counter() - t0) * 1e3)
    print("zero_copy" if zc else "staged", "wall ms", np.round(ts, 3), "launch", B.last_launch())
