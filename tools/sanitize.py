"""tools/sanitize.py — small ensembles through every kernel family on a tiny grid (BACON_IVP_GRID), so that lanes refill,
the work counter runs dry, warps suspend and the tail kernel runs; meant to be run under compute-sanitizer:
    BACON_IVP_GRID=6 compute-sanitizer --tool memcheck  python tools/sanitize.py
    BACON_IVP_GRID=6 compute-sanitizer --tool racecheck python tools/sanitize.py
Checks results against the oracle as it goes (the sanitizer only sees what actually ran)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import bacon_b200 as B  # noqa: E402
from bacon_b200 import _abi, ensembles as E  # noqa: E402
from oracle import oracle as O  # noqa: E402
from parity import band, rel_err, run_both  # noqa: E402

O.build()
P = np.array(E.LORENZ["params"])
LOR = dict(dt_min=1e-9, dt_max=0.1, tol=1e-8, t_start=0.0)
n = 3000
y0 = E.lorenz_y0(np.arange(n))
# fast RK45: final state only (tail kernel), with history (work-queue blocks + tail), strict with history
g, r = run_both(B, O, "RK45", "lorenz", y0, P, shared_params=True, t_end=0.2, **LOR)
assert B.last_launch()["n_kernels"] == 2 and rel_err(g.y_end, r["y_end"]).max() <= band(1e-8)
g, r = run_both(B, O, "RK45", "lorenz", y0, P, shared_params=True, t_end=0.2, history=256, **LOR)
assert B.last_launch()["n_kernels"] == 2 and (g.hist_len == r["hist_len"]).all()
g, r = run_both(B, O, "RK45", "lorenz", y0, P, shared_params=True, strict=True, t_end=0.1, history=100, **LOR)
assert np.array_equal(g.y_end.view(np.uint64), r["y_end"].view(np.uint64))
# RK23 with per-trajectory parameters
yv, mu = E.vdp_problem(np.arange(n) * (E.VDP["n"] // n), E.VDP["n"])
g, r = run_both(B, O, "RK23", "vdp", yv, mu, dt_min=1e-12, dt_max=0.1, tol=1e-8, t_start=0.0, t_end=0.01)
assert rel_err(g.y_end, r["y_end"]).max() <= band(1e-8)
# BDF6 Newton + Broyden, Adams5, Euler on a few hundred trajectories
yr, kr = E.robertson_problem(np.arange(512))
for flags in (_abi.FLAG_BDF_NEWTON, 0):
    g, r = run_both(B, O, "BDF6", "robertson", yr, kr, extra_flags=flags, dt_min=1e-10, dt_max=1e-4, tol=1e-6, t_start=0.0,
                    t_end=0.002)
    assert (g.status == _abi.OK).all() and rel_err(g.y_end, r["y_end"]).max() <= band(1e-6)
g, r = run_both(B, O, "Adams5", "harmonic", np.vstack([np.ones((1, 512)), np.zeros((1, 512))]), np.full((1, 512), 2.0),
                dt_min=1e-6, dt_max=0.1, tol=1e-6, t_start=0.0, t_end=1.0)
assert (g.status == r["status"]).all()
g, r = run_both(B, O, "Euler", "decay", np.ones((1, 512)), dt_min=1e-3, dt_max=1e-3, tol=1e-3, t_start=0.0, t_end=0.5)
assert (g.status == r["status"]).all()
# warp-per-trajectory linear32 with history
y32, A32 = E.linear32_problem(np.arange(64))
s = (B.RungeKutta45.new(32).with_minimum_dt(1e-9).with_maximum_dt(0.1).with_tolerance(1e-8).with_initial_time(0.0)
     .with_ending_time(0.5).with_derivative("linear32").with_history(64))
res = s.solve_ivp_ensemble(y32, A32.reshape(64, 1024), params_aos=True)
assert (res.status == _abi.OK).all()
print("sanitize.py: all families ran")
