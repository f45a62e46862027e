"""Generates tests/golden/anchors.json: independent anchors for the oracle (and through it the
CUDA path).  The reference ships no golden vectors and cannot be executed here
(pure Rust, no toolchain), so the anchors are closed forms and SciPy's DOP853 / Radau at
rtol 1e-13 on the BASELINE problems.  Run from the repo root:  python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np
from scipy.integrate import solve_ivp

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from bacon_b200 import ensembles as E  # noqa: E402


def lorenz(t, y, s, r, b):
    return [s * (y[1] - y[0]), y[0] * (r - y[2]) - y[1], y[0] * y[1] - b * y[2]]


def vdp(t, y, mu):
    return [y[1], mu * (1 - y[0] ** 2) * y[1] - y[0]]


def robertson(t, y, k1, k2, k3):
    return [-k1 * y[0] + k3 * y[1] * y[2], k1 * y[0] - k3 * y[1] * y[2] - k2 * y[1] ** 2, k2 * y[1] ** 2]


def rob_jac(t, y, k1, k2, k3):
    return [[-k1, k3 * y[2], k3 * y[1]], [k1, -k3 * y[2] - 2 * k2 * y[1], -k3 * y[1]], [0, 2 * k2 * y[1], 0]]


out = {}
# Lorenz, y0 = (1,1,1), T = 1 (SURVEY §8c predicted-answer table)
s = solve_ivp(lorenz, (0, 1), [1, 1, 1], method="DOP853", rtol=1e-13, atol=1e-13, args=(10.0, 28.0, 8.0 / 3.0))
out["lorenz_111_T1"] = s.y[:, -1].tolist()
# first 16 trajectories of the seeded config-2 ensemble, T = 2
y0 = E.lorenz_y0(np.arange(16))
ys = []
for i in range(16):
    s = solve_ivp(lorenz, (0, 2), y0[:, i], method="DOP853", rtol=1e-13, atol=1e-13, args=(10.0, 28.0, 8.0 / 3.0))
    ys.append(s.y[:, -1].tolist())
out["lorenz_seeded16_T2"] = ys
# Van der Pol, y0 = (2, 0), T = 0.25, mu in {0.1, 1, 5}
out["vdp_T025"] = {}
for mu in (0.1, 1.0, 5.0):
    s = solve_ivp(vdp, (0, 0.25), [2.0, 0.0], method="DOP853", rtol=1e-13, atol=1e-14, args=(mu,))
    out["vdp_T025"][str(mu)] = s.y[:, -1].tolist()
# Robertson, T = 0.5, Radau
s = solve_ivp(robertson, (0, 0.5), [1.0, 0.0, 0.0], method="Radau", rtol=1e-12, atol=1e-15, jac=rob_jac, args=(0.04, 3e7, 1e4))
out["robertson_T05"] = s.y[:, -1].tolist()
s = solve_ivp(robertson, (0, 0.02), [1.0, 0.0, 0.0], method="Radau", rtol=1e-12, atol=1e-15, jac=rob_jac, args=(0.04, 3e7, 1e4))
out["robertson_T002"] = s.y[:, -1].tolist()
# linear 32: first 4 trajectories of config 4 at T = 4, y(T) = expm(A T) y0
from scipy.linalg import expm  # noqa: E402
y0, A = E.linear32_problem(np.arange(4))
out["linear32_seeded4_T4"] = [(expm(A[i] * 4.0) @ y0[:, i]).tolist() for i in range(4)]

# path queries: an independent continuous solution and independent event times — SciPy's DOP853 dense
# output and its own event root-finder on the first 4 seeded Lorenz trajectories, T = 2
times = np.linspace(0.0, 2.0, 21)


def z27(t, y, s, r, b):
    return y[2] - 27.0


y0 = E.lorenz_y0(np.arange(4))
pq = {"times": times.tolist(), "states": [], "z27_events": []}
for i in range(4):
    s = solve_ivp(lorenz, (0, 2), y0[:, i], method="DOP853", rtol=1e-13, atol=1e-13, args=(10.0, 28.0, 8.0 / 3.0),
                  dense_output=True, events=z27)
    pq["states"].append(s.sol(times).T.tolist())
    pq["z27_events"].append(s.t_events[0].tolist())
out["lorenz_seeded4_T2_paths"] = pq

with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "anchors.json"), "w") as f:
    json.dump(out, f, indent=1)
print("wrote anchors.json:", {k: (len(v) if hasattr(v, "__len__") else v) for k, v in out.items()})
