"""A SECOND, independent reading of the reference's adaptive Runge-Kutta path, in plain Python floats — test
infrastructure (tests/test_oracle.py::test_rk_second_reading_...), never on the product path.

Written from the Rust source alone, statement by statement, WITHOUT looking at oracle/bacon_oracle.hpp:
  RungeKutta::solve            src/ivp/rk.rs:249-343   (dt = (dt_max + dt_min) * 1/2, half_steps = 0)
  RungeKuttaSolver::step       src/ivp/rk.rs:361-423
  IVPIterator::next            src/ivp.rs:220-238      (Ok -> yield, Redo -> again, Done -> end, Failure -> stop)
and fed with the coefficient lists parsed out of the reference's source text (tests/golden/reference_coefficients.json,
SOURCE order).  The reference cannot be executed here, so this does not make the oracle "the reference"; it makes a
transcription slip in the oracle's step or controller logic on a y-dependent right-hand side show up as a bit
difference between two readings made apart.  Python floats are IEEE doubles, `a * b + c` is never contracted, `**` is
libm's pow (the oracle's PowMode::LibmPow).

as_written=True : the tables exactly as the source builds them — BSMatrix::from_vec fills column by column, so
                  M[r][c] = listed[c * O + r] and row_iter() walks the rows of THAT matrix, 1859/4014
                  (D2), a safety factor of 100/100 (D3)  = the product's REF_LITERAL.
as_written=False: the same statements over the tables as their "Row i" comments label them, Fehlberg's 1859/4104 and
                  84/100                                = the product's REF_CORRECTED.
"""
import json
import math
import os

_REF = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_coefficients.json")))


def tables(name, as_written):
    t = _REF[name]
    o = len(t["t_coefficients"]["values"])
    listed = list(t["k_coefficients"]["values"])
    if as_written:
        m = [[listed[c * o + r] for c in range(o)] for r in range(o)]  # from_vec: column-major
        safety = _REF["RK_safety"]["values"][0]
    else:
        m = [[listed[r * o + c] for c in range(o)] for r in range(o)]
        if name == "RK45":
            assert m[5][3] == 1859.0 / 4014.0
            m[5][3] = 1859.0 / 4104.0
        safety = 84.0 / 100.0
    return dict(o=o, t=list(t["t_coefficients"]["values"]), k=m, avg=list(t["avg_coefficients"]["values"]),
                err=list(t["error_coefficients"]["values"]), safety=safety)


def lorenz(_t, y, p):
    return [p[0] * (y[1] - y[0]), y[0] * (p[1] - y[2]) - y[1], y[0] * y[1] - p[2] * y[2]]


def vdp(_t, y, p):
    return [y[1], (p[0] * (1.0 - y[0] * y[0])) * y[1] - y[0]]


def solve(name, f, y0, params, *, dt_min, dt_max, tol, t_start, t_end, as_written, max_points=100000):
    """Returns (path [(t, state)], status) with status in {"Done", "MinimumTimeDeltaExceeded"}, plus the counters
    (accepted, rejected) and the final (time, dt, state)."""
    T = tables(name, as_written)
    o, dim = T["o"], len(y0)
    # rk.rs:258-268
    half = 1.0 / 2.0
    one_tenth = 1.0 / 10.0
    four = 4.0
    one_fourth = 1.0 / four
    # rk.rs:311-339
    time, end, dt = t_start, t_end, (dt_max + dt_min) * half
    state = list(y0)
    half_steps = [[0.0] * dim for _ in range(o)]  # column j of the D x O matrix
    path, n_acc, n_rej = [], 0, 0
    while True:
        # ---- step(), rk.rs:361-423
        if time >= end:
            return path, "Done", (n_acc, n_rej), (time, dt, state)
        if time + dt >= end:
            dt = end - time
        for i in range(o):
            scratch = list(state)
            for j in range(o):
                kc = T["k"][i][j]
                scratch = [scratch[d] + half_steps[j][d] * kc for d in range(dim)]
            step_time = time + T["t"][i] * dt
            fy = f(step_time, scratch, params)
            half_steps[i] = [fy[d] * dt for d in range(dim)]
        scratch = [half_steps[0][d] * T["err"][0] for d in range(dim)]
        for ind in range(1, o):
            ec = T["err"][ind]
            scratch = [scratch[d] + half_steps[ind][d] * ec for d in range(dim)]
        ss = 0.0
        for d in range(dim):
            ss = ss + scratch[d] * scratch[d]
        error = math.sqrt(ss) / dt
        accepted = error <= tol
        if accepted:
            time += dt
            for ind in range(o):
                ac = T["avg"][ind]
                state = [state[d] + half_steps[ind][d] * ac for d in range(dim)]
        ratio = tol / error if error != 0.0 else math.inf
        delta = T["safety"] * (ratio ** one_fourth)
        if delta <= one_tenth:
            dt *= one_tenth
        elif delta >= four:
            dt *= four
        else:
            dt *= delta
        if dt > dt_max:
            dt = dt_max
        if dt < dt_min and time < end:
            if not accepted:
                n_rej += 1
            return path, "MinimumTimeDeltaExceeded", (n_acc, n_rej), (time, dt, state)
        # ---- IVPIterator::next, ivp.rs:220-238
        if accepted:
            n_acc += 1
            path.append((time, list(state)))
            if len(path) >= max_points:
                return path, "Truncated", (n_acc, n_rej), (time, dt, state)
        else:
            n_rej += 1
