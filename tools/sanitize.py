"""tools/sanitize.py — small ensembles through every kernel family on a tiny grid (BACON_IVP_GRID), so that lanes refill,
the work counter runs dry and the CTAs regroup (drive.cuh); meant to be run under compute-sanitizer:
    BACON_IVP_GRID=2 compute-sanitizer --tool memcheck  python tools/sanitize.py
    BACON_IVP_GRID=2 compute-sanitizer --tool racecheck python tools/sanitize.py
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
# fast RK45: final state only (regrouping), with history (work-queue blocks + regrouping), strict with history
g, r = run_both(B, O, "RK45", "lorenz", y0, P, shared_params=True, t_end=0.2, **LOR)
assert B.last_launch()["block"] == 768 and rel_err(g.y_end, r["y_end"]).max() <= band(1e-8)
g, r = run_both(B, O, "RK45", "lorenz", y0, P, shared_params=True, t_end=0.2, history=256, **LOR)
assert B.last_launch()["block"] == 640 and (g.hist_len == r["hist_len"]).all()
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
# path queries (path_query.cuh, path_query_warp.cuh): sampling + events incl. a queue that overflows (> 32 crossings
# between flushes) and a capacity below the count, against the oracle on the same paths
def _sv(r):
    return dict(hist=r.hist, hist_len=r.hist_len, t_end=r.t_end, y_end=r.y_end)


times = np.linspace(-0.1, 0.6, 23)
got = res.sample(times)
ref = O.sample_paths("linear32", y32, A32.reshape(64, 1024), _sv(res), times, t_start=0.0, params_aos=True)
assert np.allclose(got, ref, rtol=1e-12, atol=1e-13, equal_nan=True)
w32 = np.linspace(-1.0, 1.0, 32)
ev, cnt = res.locate_events(w32, 0.02, 0, 2)
rev, rcnt = O.locate_events("linear32", y32, A32.reshape(64, 1024), _sv(res), w32, 0.02, 0, 2, t_start=0.0, params_aos=True)
assert (cnt == rcnt).all() and np.allclose(ev, rev, rtol=1e-11, atol=1e-12)
g, r = run_both(B, O, "RK45", "lorenz", y0[:, :600], P, shared_params=True, t_end=1.5, history=2000, **LOR)
sv = _sv(g)
times = np.linspace(0.0, 1.5, 37)
assert np.allclose(g.sample(times), O.sample_paths("lorenz", y0[:, :600], P, sv, times, t_start=0.0, shared_params=True),
                   rtol=1e-12, atol=1e-12)
for cap in (1, 16):
    ev, cnt = g.locate_events([0.0, 0.0, 1.0], 27.0, 0, cap)
    rev, rcnt = O.locate_events("lorenz", y0[:, :600], P, sv, [0.0, 0.0, 1.0], 27.0, 0, cap, t_start=0.0, shared_params=True)
    assert (cnt == rcnt).all() and np.allclose(ev, rev, rtol=1e-12, atol=1e-12)
wv = np.array([[45.0, 3.0, 60.0]])
sh = (B.RungeKutta45.new(2).with_minimum_dt(1e-9).with_maximum_dt(0.1).with_tolerance(1e-5).with_initial_time(0.0)
      .with_ending_time(5.0).with_derivative("harmonic").with_history(8192))
h = sh.solve_ivp_ensemble(np.array([[1.0, 1.0, 1.0], [0.0, 0.0, 0.0]]), wv)
ev, cnt = h.locate_events([1.0, 0.0], 0.0, 0, 200)
assert cnt.max() > 64 and (np.diff(ev[2, :cnt[2], 0]) > 0).all()
print("sanitize.py: all families ran")
