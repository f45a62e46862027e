#!/usr/bin/env python
"""tools/knot_search_model.py — CPU model of the sampling kernels' knot search (PathView::first_knot_at_or_after):
distinct memory units (32-byte sectors, 64-byte bursts, 128-byte lines) and dependent probes per sample on the oracle's
own Lorenz paths (config 2, 64 trajectories, 64 sample times, a warp = 32 neighbouring times of one trajectory).
Test/measurement infrastructure: it runs the ORACLE, nothing here is on the product path.

    python tools/knot_search_model.py            # table in profiles/r04_sampling_search.md
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def paths(n=64):
    from bacon_b200 import _abi, ensembles as E
    from oracle import oracle as O
    O.build()
    r = O.solve_ensemble(_abi.RK45, "lorenz", E.lorenz_y0(np.arange(n)), np.array(E.LORENZ["params"]), shared_params=True,
                         history_capacity=4608, pow_mode=1, t_start=0.0, t_end=5.0, dt_min=1e-9, dt_max=0.1, tol=1e-8)
    return r["hist_t"], r["hist_len"]


def search(T, K, tau, probes, below, cap=8):
    """The kernel's probe sequence: bisection while the bracket is wider than `below`, then interpolation (float32)."""
    lo, hi, tl, th, ni = 1, K, T[0], T[K], 0
    while lo < hi:
        width = hi - lo
        mid = (lo + hi) >> 1
        if width <= below and width > 1 and ni < cap and th > tl:
            ni += 1
            x = np.float32(tau - tl) / np.float32(th - tl) * np.float32(width + 1)
            mid = min(max((lo - 1) + int(np.ceil(x)), lo), hi - 1)
        probes.append(mid)
        if T[mid] >= tau:
            hi, th = mid, T[mid]
        else:
            lo, tl = mid + 1, T[mid]
    return lo


def model(ht, hl, below, gran):
    times = np.linspace(0.0, 5.0, 64)
    units = n_probe = n_samp = 0
    worst = []
    for i in range(ht.shape[0]):
        m = int(hl[i])
        T = np.concatenate([[0.0], ht[i, :m]])
        for w in range(2):
            seen, mx = set(), 0
            for tau in times[32 * w:32 * w + 32]:
                if not (T[0] <= tau <= T[m]):
                    continue
                probes = []
                lo = search(T, m, tau, probes, below)
                assert lo == int(np.searchsorted(T[1:m + 1], tau, side="left")) + 1
                for k in probes + [lo - 1, lo]:  # records 1..m sit at (k - 1) * 32 bytes
                    if k >= 1:
                        seen.add((k - 1) * 32 // gran)
                n_probe += len(probes)
                mx = max(mx, len(probes))
                n_samp += 1
            units += len(seen)
            worst.append(mx)
    return units / n_samp, n_probe / n_samp, float(np.mean(worst)), int(np.max(worst))


if __name__ == "__main__":
    ht, hl = paths()
    print("| bracket below which the search interpolates | 32-B sectors / sample | 64-B bursts | 128-B lines | x 128 B | probes / sample | slowest lane of a warp (mean, max) |")
    print("|---|---:|---:|---:|---:|---:|---|")
    for below, name in ((0, "never (bisection, round 1)"), (32, "32"), (64, "64 (shipped)"), (128, "128"), (4096, "whole path")):
        s = [model(ht, hl, below, g) for g in (32, 64, 128)]
        print(f"| {name} | {s[0][0]:.2f} | {s[1][0]:.2f} | {s[2][0]:.2f} | {s[2][0] * 128:.0f} B | {s[0][1]:.1f} | {s[0][2]:.1f}, {s[0][3]} |")
