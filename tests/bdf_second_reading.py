"""A SECOND, independent reading of the reference's BDF path for ONE-dimensional problems, in plain Python floats —
test infrastructure (tests/test_oracle.py::test_bdf_second_reading_...), never on the product path.

Written from the Rust source alone, statement by statement, WITHOUT looking at oracle/bacon_oracle.hpp:
  BDF::solve                   src/ivp/bdf.rs:257-343   (dt = (dt_max + dt_min) * 1/2, empty history)
  BDFSolver::runge_kutta       src/ivp/bdf.rs:346-387
  BDFSolver::jac_finite_diff   src/ivp/bdf.rs:390-411
  BDFSolver::secant            src/ivp/bdf.rs:414-475   (Broyden with Sherman-Morrison updates)
  BDFSolver::step              src/ivp/bdf.rs:495-634
  IVPIterator::next            src/ivp.rs:220-238
over the coefficient lists parsed out of the reference's source text (tests/golden/reference_coefficients.json).
Dimension 1 on purpose: `jac.lu().try_inverse()` lives in nalgebra, which is not in the reference tree, and for a 1 x 1
matrix every inversion algorithm is 1 / x — so nothing here leans on a restatement of nalgebra.  What the reading covers
is everything that is bacon's own: the start-up blocks, the speculative first implicit step and its rollback, the
yield bookkeeping, both implicit functions, the Broyden iteration, halving / doubling.

as_written=True : the source as it stands (SURVEY.md D4 `above + below`, D5 the lower function walks the HIGHER
                  coefficients, D6 `time -= dt - order`, D7 g evaluated at t_n)            = REF_LITERAL
as_written=False: D4-D7 repaired as SURVEY.md section 8c states them (central difference, lower coefficients,
                  `time -= dt * order`, g at t_n + dt), nothing else                         = REF_CORRECTED
"""
import json
import math
import os

_REF = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_coefficients.json")))


class Failure(Exception):
    pass


def solve(name, f, y0, *, dt_min, dt_max, tol, t_start, t_end, as_written, max_points=100000, max_calls=10**7):
    """f(t, y) -> y' for scalars.  Returns (path [(t, y)], status, (time, dt, state)) — status "Done" or an IVPError name."""
    higher = list(_REF[name]["higher_coefficients"]["values"])
    lower = list(_REF[name]["lower_coefficients"]["values"])
    O = len(higher)
    two = 2.0
    half = 1.0 / two
    one_sixth = 1.0 / 6.0
    one_tenth = 1.0 / 10.0
    order = float(O)
    S = dict(time=t_start, dt=(dt_max + dt_min) * half, state=y0, prev=[], save=0.0, yield_memory=0)
    end = t_end

    def runge_kutta(iterations):
        for i in range(iterations):
            k1 = f(S["time"], S["state"]) * S["dt"]
            inter = S["state"] + k1 * half
            k2 = f(S["time"] + half * S["dt"], inter) * S["dt"]
            inter = S["state"] + k2 * half
            k3 = f(S["time"] + half * S["dt"], inter) * S["dt"]
            inter = S["state"] + k3
            k4 = f(S["time"] + S["dt"], inter) * S["dt"]
            if i != 0:
                S["prev"].append((S["time"], S["state"]))
            S["state"] = S["state"] + (k1 + k2 * two + k3 * two + k4) * one_sixth
            S["time"] = S["time"] + S["dt"]
        S["prev"].append((S["time"], S["state"]))

    def implicit(coefs0, tail):
        def g(t, y):
            scratch = -f(t, y) * S["dt"] * coefs0
            for ind in range(1, O):
                scratch = scratch + S["prev"][O - ind][1] * tail[ind]
            return scratch + y
        return g

    def jac_finite_diff(x, g, t):
        denom = 1.0 / (two * S["dt"])
        x = x + S["dt"]
        above = g(t, x)
        x = x - two * S["dt"]
        below = g(t, x)
        x = x + S["dt"]
        col = (above + below) * denom if as_written else (above - below) * denom
        return col, x  # (the caller's guess comes back through the += / -= of the source, rounding included)

    def secant(g):
        t = S["time"] if as_written else S["time"] + S["dt"]
        n = 2
        guess = S["state"]
        derivative = g(t, guess)
        jac, guess = jac_finite_diff(guess, g, t)
        if jac == 0.0 or jac != jac:
            raise Failure("SingularMatrix")
        jac_inv = 1.0 / jac
        shift = -jac_inv * derivative
        guess = guess + shift
        while n < 1000:
            derivative_last = derivative
            derivative = g(t, guess)
            difference = derivative - derivative_last
            adjustment = -jac_inv * difference
            p = -shift * adjustment
            u = shift * jac_inv
            jac_inv = jac_inv + (shift + adjustment) * u / p
            shift = -jac_inv * derivative
            guess = guess + shift
            if math.sqrt(shift * shift) <= tol:
                return guess
            n += 1
        raise Failure("MaximumIterationsExceeded")

    def step():
        """'ok', (t, y) | 'redo' | 'done'; raises Failure"""
        if 0 < S["yield_memory"] <= O:
            get_item = O - S["yield_memory"]
            S["yield_memory"] -= 1
            if S["yield_memory"] == 0:
                S["yield_memory"] = O + 2
            return "ok", S["prev"][get_item]
        if S["yield_memory"] == O + 2:
            S["yield_memory"] = 0
            S["prev"].append((S["time"], S["state"]))
            S["prev"].pop(0)
            return "ok", (S["time"], S["state"])
        if S["time"] >= end:
            return "done", None
        if S["time"] + S["dt"] >= end:
            S["dt"] = end - S["time"]
            runge_kutta(1)
            return "ok", (S["time"], S["prev"][-1][1])
        if not S["prev"]:
            S["save"] = S["state"]
            if S["time"] + S["dt"] * order >= end:
                S["dt"] = (end - S["time"]) / order
            runge_kutta(O)
            S["yield_memory"] = O + 1
            return "redo", None
        higher_step = secant(implicit(higher[0], higher))
        lower_step = secant(implicit(lower[0], higher if as_written else lower))
        difference = higher_step - lower_step
        error = math.sqrt(difference * difference)
        if error <= tol:
            S["state"] = higher_step
            S["time"] = S["time"] + S["dt"]
            if S["yield_memory"] == O + 1:
                S["yield_memory"] -= 1
                return "redo", None
            S["prev"].append((S["time"], S["state"]))
            S["prev"].pop(0)
            if error < one_tenth * tol:
                S["dt"] = S["dt"] * two
                if S["dt"] > dt_max:
                    S["dt"] = dt_max
                S["prev"].clear()
            return "ok", (S["time"], S["state"])
        if S["yield_memory"] == O + 1:
            S["time"] = S["time"] - (S["dt"] - order) if as_written else S["time"] - S["dt"] * order
            S["state"] = S["save"]
        S["dt"] = S["dt"] * half
        if S["dt"] < dt_min:
            raise Failure("MinimumTimeDeltaExceeded")
        S["prev"].clear()
        return "redo", None

    path, calls = [], 0
    while True:
        calls += 1
        if calls > max_calls:
            return path, "Truncated", (S["time"], S["dt"], S["state"])
        try:
            what, item = step()
        except Failure as e:
            return path, str(e), (S["time"], S["dt"], S["state"])
        if what == "done":
            return path, "Done", (S["time"], S["dt"], S["state"])
        if what == "ok":
            path.append(item)
            if len(path) >= max_points:
                return path, "Truncated", (S["time"], S["dt"], S["state"])
