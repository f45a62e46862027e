"""tools/sorted_probe.py N — what perfect packing would buy: the same N Lorenz trajectories, (a) in seeded order,
(b) sorted by their step count (known from run a), so that the 32 lanes of a warp end together and nothing needs to
be regrouped.  Run with the production library and with a -DBACON_NO_REGROUP build."""
import sys
import numpy as np
import torch
import bacon_b200 as B
from bacon_b200 import ensembles as E
n = int(sys.argv[1])
w = E.LORENZ
p = torch.tensor(w["params"], dtype=torch.float64).cuda()
s = (B.RK45.new(3).with_dt_min(w["dt_min"]).with_dt_max(w["dt_max"]).with_tolerance(w["tol"]).with_start(0.0)
     .with_end(w["t_end"]).with_derivative("lorenz"))
def run(y0):
    d = torch.from_numpy(np.ascontiguousarray(y0)).cuda()
    best = 1e9
    for _ in range(4):
        out = s.solve_ivp_ensemble_device(d, p, shared_params=True)
        torch.cuda.synchronize()
        best = min(best, B.last_launch()["kernel_ms"])
    return best, (out["n_accept"] + out["n_reject"]).cpu().numpy()
y0 = E.lorenz_y0(np.arange(n))
t_seed, att = run(y0)
order = np.argsort(-att, kind="stable")
t_sorted, _ = run(y0[:, order])
print(f"n={n} grid {B.last_launch()['grid']}x{B.last_launch()['block']}: seeded order {t_seed:.3f} ms, sorted by step count {t_sorted:.3f} ms")
