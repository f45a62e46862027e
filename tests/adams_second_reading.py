"""A SECOND, independent reading of the reference's Adams predictor-corrector path, in plain Python floats — test
infrastructure (tests/test_oracle.py::test_adams_second_reading_...), never on the product path.

Written from the Rust source alone, statement by statement, WITHOUT looking at oracle/bacon_oracle.hpp:
  Adams::solve                 src/ivp/adams.rs:249-336   (dt = (dt_max + dt_min) * 1/2, empty histories)
  AdamsSolver::runge_kutta     src/ivp/adams.rs:344-394
  AdamsSolver::step            src/ivp/adams.rs:411-561
  IVPIterator::next            src/ivp.rs:220-238
over the coefficient lists parsed out of the reference's source text (tests/golden/reference_coefficients.json).
Any dimension (there is no linear algebra on this path).

as_written=True : the source as it stands                                                      = REF_LITERAL
as_written=False: D10 of DESIGN.md section 2 repaired — the speculative first predictor-corrector step after a start-up
                  block, when accepted, also records its derivative (push_back + pop_front, as every regular step
                  does at adams.rs:503-509), so that the derivative history the later steps read "already has the
                  derivatives for this step" as the comment at adams.rs:429-433 says                = REF_CORRECTED
"""
import json
import math
import os

_REF = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_coefficients.json")))


def _axpy(a, x, y):  # y + x * a, elementwise (one multiplication, one addition: nothing is contracted)
    return [y[d] + x[d] * a for d in range(len(x))]


def solve(name, f, y0, params, *, dt_min, dt_max, tol, t_start, t_end, as_written, max_points=100000, max_calls=10**7):
    pc = list(_REF[name]["predictor_coefficients"]["values"])
    cc = list(_REF[name]["corrector_coefficients"]["values"])
    ec = _REF[name]["error_coefficient"]["values"][0]
    O = len(pc)
    dim = len(y0)
    two = 2.0
    half = 1.0 / two
    one_sixth = 1.0 / 6.0
    one_tenth = 1.0 / 10.0
    four = two * two
    order = float(O)
    S = dict(time=t_start, dt=(dt_max + dt_min) * half, state=list(y0), pv=[], pd=[], save=None, ym=0, impl=None)
    end = t_end

    def F(t, y):
        return f(t, y, params)

    def runge_kutta(iterations):
        for i in range(iterations):
            dt, st, t = S["dt"], S["state"], S["time"]
            k1 = [v * dt for v in F(t, st)]
            inter = [st[d] + k1[d] * half for d in range(dim)]
            k2 = [v * dt for v in F(t + half * dt, inter)]
            inter = [st[d] + k2[d] * half for d in range(dim)]
            k3 = [v * dt for v in F(t + half * dt, inter)]
            inter = [st[d] + k3[d] for d in range(dim)]
            k4 = [v * dt for v in F(t + dt, inter)]
            if i != 0:
                S["pd"].append(F(t, st))
                S["pv"].append((t, list(st)))
            S["state"] = [st[d] + (k1[d] + k2[d] * two + k3[d] * two + k4[d]) * one_sixth for d in range(dim)]
            S["time"] = t + dt
        S["pd"].append(F(S["time"], S["state"]))
        S["pv"].append((S["time"], list(S["state"])))

    def step():
        if 0 < S["ym"] < O:
            get_item = O - S["ym"] - 1
            S["ym"] -= 1
            if S["ym"] == 0:
                S["ym"] = O + 1
            return "ok", S["pv"][get_item]
        if S["ym"] == O + 1:
            S["ym"] = 0
            S["pv"].append((S["time"], list(S["state"])))
            S["pv"].pop(0)
            return "ok", (S["time"], list(S["state"]))
        if S["time"] >= end:
            return "done", None
        if S["time"] + S["dt"] >= end:
            S["dt"] = end - S["time"]
            runge_kutta(1)
            return "ok", (S["time"], list(S["pv"][-1][1]))
        if not S["pv"]:
            S["save"] = list(S["state"])
            if S["time"] + S["dt"] * (order - 1.0) >= end:
                S["dt"] = (end - S["time"]) / (order - 1.0)
            runge_kutta(O - 1)
            S["ym"] = O
            return "redo", None
        scratch = [v * pc[O - 2] for v in S["pd"][0]]
        for i in range(1, O - 1):
            scratch = _axpy(pc[O - i - 2], S["pd"][i], scratch)
        predictor = [S["state"][d] + scratch[d] * S["dt"] for d in range(dim)]
        S["impl"] = F(S["time"] + S["dt"], predictor)
        scratch = [v * cc[0] for v in S["impl"]]
        for i in range(0, O - 1):
            scratch = _axpy(cc[O - i - 1], S["pd"][i], scratch)
        corrector = [S["state"][d] + scratch[d] * S["dt"] for d in range(dim)]
        ss = 0.0
        for d in range(dim):
            diff = corrector[d] - predictor[d]
            ss = ss + diff * diff
        error = ec / S["dt"] * math.sqrt(ss)
        if error <= tol:
            S["state"] = corrector
            S["time"] = S["time"] + S["dt"]
            if S["ym"] == O:
                S["ym"] -= 1
                if not as_written:  # D10
                    S["pd"].append(list(S["impl"]))
                    S["pd"].pop(0)
                return "redo", None
            S["pd"].append(list(S["impl"]))
            S["pv"].append((S["time"], list(S["state"])))
            S["pv"].pop(0)
            S["pd"].pop(0)
            if error < one_tenth * tol:
                q = (tol / (two * error)) ** (1.0 / order) if error != 0.0 else math.inf
                if q > four:
                    S["dt"] = S["dt"] * four
                else:
                    S["dt"] = S["dt"] * q
                if S["dt"] > dt_max:
                    S["dt"] = dt_max
                S["pv"].clear()
                S["pd"].clear()
            return "ok", (S["time"], list(S["state"]))
        if S["ym"] == O:
            S["time"] = S["time"] - S["dt"] * (order - 1.0)
            S["state"] = list(S["save"])
        q = (tol / (two * error)) ** (1.0 / order)
        if q < one_tenth:
            S["dt"] = S["dt"] * one_tenth
        else:
            S["dt"] = S["dt"] * q
        if S["dt"] < dt_min:
            return "fail", "MinimumTimeDeltaExceeded"
        S["pv"].clear()
        S["pd"].clear()
        return "redo", None

    path, calls = [], 0
    while True:
        calls += 1
        if calls > max_calls:
            return path, "Truncated", (S["time"], S["dt"], S["state"])
        what, item = step()
        if what == "done":
            return path, "Done", (S["time"], S["dt"], S["state"])
        if what == "fail":
            return path, item, (S["time"], S["dt"], S["state"])
        if what == "ok":
            path.append(item)
            if len(path) >= max_points:
                return path, "Truncated", (S["time"], S["dt"], S["state"])
