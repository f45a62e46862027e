"""Pins the CPU oracle (oracle/) before anything is compared against it (CPU only, no GPU):

 * every test problem the reference holds for this path, replayed with the reference's own
   assertion on every yielded point (rk.rs:682-758; bdf.rs:785-1063 in REF_CORRECTED, and the
   vacuous empty-path outcome of the same tests in REF_LITERAL);
 * the known answers of src/tests/roots/mod.rs:169-221 for the Broyden + LU code BDF shares;
 * the predicted-answer table of SURVEY.md §8c;
 * independent anchors: closed forms and SciPy DOP853 / Radau (tests/golden/anchors.json);
 * every coefficient table against the numbers parsed out of the reference's source text
   (tests/golden/reference_coefficients.json + make_reference_coefficients.py);
 * second, independent readings of rk.rs, bdf.rs and adams.rs in plain Python (rk_second_reading.py,
   bdf_second_reading.py, adams_second_reading.py): bit for bit with the oracle on y-dependent
   problems, both semantics.
The Rust reference cannot be executed in this image: "parity unpinned" for what the RK step does
with its stage matrix on y-dependent problems and for BDF beyond these anchors (see
oracle/bacon_oracle.hpp header).
"""
import json
import os

import numpy as np
import pytest

from bacon_b200 import _abi
from bacon_b200 import ensembles as E
from reference_cases import ADAMS_CASES, BDF_CASES, EULER_CASES, RK_CASES, bdf_cfg

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "anchors.json")))
M = {"RK45": _abi.RK45, "RK23": _abi.RK23, "BDF6": _abi.BDF6, "BDF2": _abi.BDF2, "Adams5": _abi.ADAMS5,
     "Adams3": _abi.ADAMS3, "Euler": _abi.EULER}


@pytest.mark.parametrize("method", ["RK45", "RK23"])
@pytest.mark.parametrize("case", RK_CASES, ids=[c[0] for c in RK_CASES])
def test_reference_rk_tests(oracle, method, case):
    name, rhs, y0, cfg, exact, eps, n_acc = case
    outs = []
    for sem in (_abi.SEM_CORRECTED, _abi.SEM_LITERAL):
        r = oracle.solve_ensemble(M[method], rhs, np.array([[y0]]), semantics=sem, history_capacity=1100, **cfg)
        assert r["status"][0] == _abi.OK and r["n_accept"][0] == n_acc and r["n_reject"][0] == 0
        m = int(r["hist_len"][0])
        t, y = r["hist_t"][0, :m], r["hist_y"][0, :m, 0]
        assert np.abs(y - exact(t)).max() <= eps  # the reference's assertion
        assert t[-1] == cfg["t_end"]
        outs.append((t.copy(), y.copy()))
    # y-independent RHS: the stage matrix never matters, so LITERAL == CORRECTED bit for bit
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])


@pytest.mark.parametrize("case", BDF_CASES, ids=[c[0] for c in BDF_CASES])
def test_reference_bdf_tests(oracle, case):
    name, method, rhs, y0, t_end, exact, eps, n_yield, lit_t = case
    r = oracle.solve_ensemble(M[method], rhs, np.array([[y0]]), history_capacity=25000, **bdf_cfg(t_end))
    assert r["status"][0] == _abi.OK
    m = int(r["hist_len"][0])
    assert m == r["n_accept"][0] and m > 0
    if n_yield is not None:
        assert m == n_yield
    t, y = r["hist_t"][0, :m], r["hist_y"][0, :m, 0]
    assert np.abs(y - exact(t)).max() <= eps  # the reference's assertion, now over a non-empty path
    assert np.all(np.diff(t) > 0)
    # as written (D4-D6): the rollback `time -= dt - order` jumps past t_end -> Done with an EMPTY path,
    # so the reference's own assertion loop runs zero times
    lit = oracle.solve_ensemble(M[method], rhs, np.array([[y0]]), semantics=_abi.SEM_LITERAL, history_capacity=16,
                                **bdf_cfg(t_end))
    assert lit["status"][0] == _abi.OK and lit["n_accept"][0] == 0 and lit["hist_len"][0] == 0
    assert lit["t_end"][0] == pytest.approx(lit_t, abs=1e-3)


def test_roots_secant_known_answers(oracle):
    """src/tests/roots/mod.rs:169-221 — same Broyden + LU code as bdf.rs:414-475."""
    for central in (False, True):  # `above + below` as written (roots/mod.rs:249) and the central difference
        rc, sol, it = oracle.roots_secant(0, [0.1, 0.1, -0.1], 0.1, 1e-5, central=central)
        assert rc == 0 and np.allclose(sol, [0.5, 0.0, -0.52359877], atol=1e-5)
        rc, sol, it = oracle.roots_secant(1, [0.7, 1.8], 0.1, 1e-6, central=central)
        assert rc == 0 and abs(sol[0] + 0.703467) <= 1e-6 and abs(sol[1] - 1.85718) <= 1e-5
        rc, sol, it = oracle.roots_secant(1, [0.7, 4.7], 0.1, 1e-6, central=central)
        assert rc == 0 and abs(sol[0] + 0.703467) <= 1e-6 and abs(sol[1] - 4.5364) <= 1e-5
        rc, sol, it = oracle.roots_secant(2, [0.6], 0.1, 1e-4, central=central)
        assert rc == 0 and abs(sol[0] - 0.739085) <= 1e-6


def test_predicted_answers_readme_and_doc_example(oracle):
    """SURVEY.md §8c table, BASELINE config 1 (README.md:32-40) and rk.rs:544-552."""
    y0 = np.array([[1.0]])
    cfg = dict(dt_min=0.01, dt_max=0.1, tol=1e-4, t_start=0.0, t_end=10.0)
    r = oracle.solve_ensemble(_abi.RK45, "exp", y0, **cfg)
    assert (r["status"][0], r["n_accept"][0], r["n_reject"][0]) == (_abi.OK, 128, 0)
    assert r["y_end"][0, 0] == pytest.approx(22026.4819, rel=1e-8)
    assert abs(r["y_end"][0, 0] / np.exp(10.0) - 1.0) == pytest.approx(7.3e-7, rel=0.05)
    r = oracle.solve_ensemble(_abi.RK45, "exp", y0, semantics=_abi.SEM_LITERAL, **cfg)
    assert r["status"][0] == _abi.E_MIN_DT_EXCEEDED and r["n_accept"][0] == 1
    assert r["t_end"][0] == pytest.approx(0.055) and r["y_end"][0, 0] == pytest.approx(1.055)
    r = oracle.solve_ensemble(_abi.RK45, "exp", y0, semantics=_abi.SEM_LITERAL, **{**cfg, "dt_min": 1e-3})
    assert r["status"][0] == _abi.E_MIN_DT_EXCEEDED and r["n_accept"][0] == 1 and r["t_end"][0] == pytest.approx(0.0505)


def test_lorenz_against_scipy(oracle):
    p = np.array(E.LORENZ["params"])
    cfg = dict(dt_min=1e-9, dt_max=0.1, tol=1e-8, t_start=0.0)
    r = oracle.solve_ensemble(_abi.RK45, "lorenz", np.ones((3, 1)), p, shared_params=True, t_end=1.0, **cfg)
    assert (r["n_accept"][0], r["n_reject"][0]) == (925, 2)
    np.testing.assert_allclose(r["y_end"][:, 0], GOLD["lorenz_111_T1"], rtol=2e-9)
    np.testing.assert_allclose(r["y_end"][:, 0], [-9.3785700104, -8.3570337871, 29.3623253376], rtol=1e-10)
    y0 = E.lorenz_y0(np.arange(16))
    r = oracle.solve_ensemble(_abi.RK45, "lorenz", y0, p, shared_params=True, t_end=2.0, **cfg)
    gold = np.array(GOLD["lorenz_seeded16_T2"]).T
    err = np.sqrt(((r["y_end"] - gold) ** 2).sum(0)) / np.sqrt((gold ** 2).sum(0))
    assert err.max() < 1e-7  # global error of the tol-1e-8 integration itself
    # RK23 converges to the same anchors
    r = oracle.solve_ensemble(_abi.RK23, "lorenz", np.ones((3, 1)), p, shared_params=True, t_end=1.0, **{**cfg, "tol": 1e-7})
    np.testing.assert_allclose(r["y_end"][:, 0], GOLD["lorenz_111_T1"], rtol=1e-6)


def test_vdp_and_linear32_against_anchors(oracle):
    for mu in (0.1, 1.0, 5.0):
        r = oracle.solve_ensemble(_abi.RK23, "vdp", np.array([[2.0], [0.0]]), np.array([[mu]]), dt_min=1e-12, dt_max=0.1,
                                  tol=1e-10, t_start=0.0, t_end=0.25)
        assert r["status"][0] == _abi.OK
        np.testing.assert_allclose(r["y_end"][:, 0], GOLD["vdp_T025"][str(mu)], rtol=1e-8, atol=1e-10)
    y0, A = E.linear32_problem(np.arange(4))
    r = oracle.solve_ensemble(_abi.RK45, "linear32", y0, A, params_aos=True, dt_min=1e-9, dt_max=0.1, tol=1e-8, t_start=0.0,
                              t_end=4.0, history_capacity=256)
    gold = np.array(GOLD["linear32_seeded4_T4"]).T
    err = np.sqrt(((r["y_end"] - gold) ** 2).sum(0)) / np.sqrt((gold ** 2).sum(0))
    assert (r["status"] == _abi.OK).all() and err.max() < 1e-6
    assert (r["hist_len"] > 100).all() and (r["hist_len"] < 256).all()  # config 4: ~160 points, capacity 256


def test_robertson_bdf6_against_radau(oracle):
    """5005 yielded points, y(0.5) = SciPy Radau to 9 digits; Broyden and Newton agree."""
    k = np.array([0.04, 3e7, 1e4])
    y0 = np.array([[1.0], [0.0], [0.0]])
    cfg = dict(dt_min=1e-10, dt_max=1e-4, tol=1e-6, t_start=0.0, t_end=0.5)
    br = oracle.solve_ensemble(_abi.BDF6, "robertson", y0, k, shared_params=True, **cfg)
    nw = oracle.solve_ensemble(_abi.BDF6, "robertson", y0, k, shared_params=True, bdf_newton=True, **cfg)
    for r in (br, nw):
        assert r["status"][0] == _abi.OK and r["n_accept"][0] == 5005
        np.testing.assert_allclose(r["y_end"][:, 0], GOLD["robertson_T05"], rtol=2e-8)
    np.testing.assert_allclose(br["y_end"], nw["y_end"], rtol=1e-9)
    # as written: finite-difference "Jacobian" is a SUM (bdf.rs:407) -> near rank-1 -> failure
    # with dt_max = 1e-4 the first rejected speculative step jumps time forward by O - dt (bdf.rs:622, D6) past
    # t_end: Done with nothing yielded; with dt_max = 1e-2 the sum-"Jacobian" yields NaN -> MaximumIterationsExceeded
    lit = oracle.solve_ensemble(_abi.BDF6, "robertson", y0, k, shared_params=True, semantics=_abi.SEM_LITERAL, **cfg)
    assert lit["status"][0] == _abi.OK and lit["n_accept"][0] == 0 and lit["t_end"][0] > 7.0
    lit = oracle.solve_ensemble(_abi.BDF6, "robertson", y0, k, shared_params=True, semantics=_abi.SEM_LITERAL,
                                **{**cfg, "dt_max": 1e-2})
    assert lit["status"][0] == _abi.E_MAX_ITER and lit["n_accept"][0] == 0


def test_pow_vs_sqrt_sqrt(oracle):
    """rk.rs:401 calls powf(x, 1/4); the strict device kernels use sqrt(sqrt(x)).  Both are within 1 ulp,
    and the final states they lead to agree far inside the parity band."""
    x = np.exp(np.random.default_rng(1).uniform(np.log(1e-12), np.log(1e12), 200000))
    a, b = oracle.fourth_root(x)
    assert (np.abs(a - b) <= np.spacing(a)).all()
    assert (a == b).mean() > 0.7
    y0 = E.lorenz_y0(np.arange(64))
    cfg = dict(dt_min=1e-9, dt_max=0.1, tol=1e-8, t_start=0.0, t_end=5.0)
    p = np.array(E.LORENZ["params"])
    r0 = oracle.solve_ensemble(_abi.RK45, "lorenz", y0, p, shared_params=True, pow_mode=0, **cfg)
    r1 = oracle.solve_ensemble(_abi.RK45, "lorenz", y0, p, shared_params=True, pow_mode=1, **cfg)
    err = np.sqrt(((r0["y_end"] - r1["y_end"]) ** 2).sum(0)) / np.sqrt((r0["y_end"] ** 2).sum(0))
    assert err.max() < 1e-9


def test_oracle_failure_statuses(oracle):
    y0 = E.lorenz_y0(np.arange(8))
    p = np.array(E.LORENZ["params"])
    r = oracle.solve_ensemble(_abi.RK45, "lorenz", y0, p, shared_params=True, dt_min=5e-3, dt_max=0.1, tol=1e-8, t_start=0, t_end=5)
    assert (r["status"] == _abi.E_MIN_DT_EXCEEDED).all()
    r = oracle.solve_ensemble(_abi.RK45, "lorenz", y0, p, shared_params=True, dt_min=1e-9, dt_max=0.1, tol=1e-8, t_start=0, t_end=5,
                              max_attempts=50)
    assert (r["status"] == _abi.E_MAX_ATTEMPTS).all() and (r["n_accept"] + r["n_reject"] == 50).all()
    y0[0, 3] = np.nan
    r = oracle.solve_ensemble(_abi.RK45, "lorenz", y0, p, shared_params=True, dt_min=1e-9, dt_max=0.1, tol=1e-8, t_start=0, t_end=0.1)
    assert r["status"][3] == _abi.E_NONFINITE and (np.delete(r["status"], 3) == _abi.OK).all()


@pytest.mark.parametrize("case", ADAMS_CASES, ids=[c[0] for c in ADAMS_CASES])
def test_reference_adams_tests(oracle, case):
    """adams.rs:714-922: the six live Adams tests (two of them on a y-dependent right-hand side), the
    reference's assertion on every yielded point.  REF_LITERAL is the source as written (these tests pin it);
    REF_CORRECTED repairs D10 (the speculative step's derivative never reaches the deque, adams.rs:498-501):
    same assertion, far fewer steps on the y-dependent problems.  Both through both evaluations of x^(1/order)."""
    name, method, rhs, y0, cfg, exact, eps, lit, cor = case
    for sem, (n_yield, n_rej) in ((_abi.SEM_LITERAL, lit), (_abi.SEM_CORRECTED, cor)):
        outs = []
        for pm in (0, 1):  # libm pow (f64::powf) / the deterministic root of the strict device kernels
            r = oracle.solve_ensemble(M[method], rhs, np.array([[y0]]), history_capacity=8000, pow_mode=pm, semantics=sem,
                                      **cfg)
            assert r["status"][0] == _abi.OK
            m = int(r["hist_len"][0])
            assert (m, int(r["n_reject"][0])) == (n_yield, n_rej) and m == r["n_accept"][0]
            t, y = r["hist_t"][0, :m], r["hist_y"][0, :m, 0]
            assert np.abs(y - exact(t)).max() <= eps  # the reference's assertion
            assert np.all(np.diff(t) > 0) and t[0] > cfg["t_start"]  # the initial condition is not yielded
            assert r["t_end"][0] >= cfg["t_end"]
            outs.append(r)
        np.testing.assert_allclose(outs[0]["y_end"], outs[1]["y_end"], rtol=1e-12)
        assert np.array_equal(outs[0]["n_rhs"], outs[1]["n_rhs"])


def test_adams_root_modes_agree(oracle):
    """adams.rs:527,:553 call powf(x, 1/order); the strict kernels use a Newton root built from +,*,/ only."""
    x = np.exp(np.random.default_rng(2).uniform(np.log(1e-300), np.log(1e300), 200000))
    for n in (3, 5):
        a, b = oracle.nth_root(x, n)
        # powf's exponent is fl(1/n), not 1/n: the reference's own value is off the true root by |ln x|/n * 2^-53
        assert (np.abs(a - b) <= 4 * np.spacing(a) + a * np.abs(np.log(x)) / n * 1.2e-16).all()
        near = (x > 1e-6) & (x < 1e6)  # the range q is evaluated on in practice
        assert (np.abs(a - b)[near] <= 4 * np.spacing(a[near])).all() or near.sum() == 0
        a, b = oracle.nth_root(np.array([0.0, np.inf, 1.0, 32.0, 27.0]), n)
        assert a[0] == b[0] == 0.0 and np.isinf(b[1]) and b[2] == 1.0
    assert oracle.nth_root(np.array([32.0]), 5)[1][0] == 2.0 and oracle.nth_root(np.array([27.0]), 3)[1][0] == 3.0


def test_adams_y_dependent_ensemble_against_closed_form(oracle):
    """Harmonic oscillators with per-trajectory frequency (y-dependent, 2-dim): both Adams orders converge to the
    closed form, and dropping the pending warm-up block at Done (the D9 analogue, adams.rs:442-463) only ever
    loses yielded points, never the final state."""
    n = 64
    w = np.linspace(0.5, 3.0, n)
    y0 = np.stack([np.ones(n), np.zeros(n)])
    exact = np.stack([np.cos(3.0 * w), -w * np.sin(3.0 * w)])
    for method, tol in (("Adams5", 1e-8), ("Adams3", 1e-6)):
        r = oracle.solve_ensemble(M[method], "harmonic", y0, w[None, :], dt_min=1e-8, dt_max=0.05, tol=tol, t_start=0.0,
                                  t_end=3.0)
        assert (r["status"] == _abi.OK).all() and (r["t_end"] >= 3.0).all()
        assert np.abs(r["y_end"] - exact).max() < 10 * tol and r["n_accept"].max() < 6000
    # as written (D10): the same problem needs ~1/tol steps, and still converges
    lit = oracle.solve_ensemble(M["Adams5"], "harmonic", y0[:, -4:], w[None, -4:], dt_min=1e-8, dt_max=0.05, tol=1e-5,
                                t_start=0.0, t_end=3.0, semantics=_abi.SEM_LITERAL)
    cor = oracle.solve_ensemble(M["Adams5"], "harmonic", y0[:, -4:], w[None, -4:], dt_min=1e-8, dt_max=0.05, tol=1e-5,
                                t_start=0.0, t_end=3.0)
    assert (lit["status"] == _abi.OK).all() and (lit["n_accept"] > 100 * cor["n_accept"]).all()
    assert np.abs(lit["y_end"] - exact[:, -4:]).max() < 1e-3


@pytest.mark.parametrize("case", EULER_CASES, ids=[c[0] for c in EULER_CASES])
def test_reference_euler_tests(oracle, case):
    """ivp.rs:539-653.  Euler yields the OLD point of every step: the path starts AT the initial condition and
    never contains the final state (ivp.rs:331-337)."""
    name, rhs, y0, dt, exact, eps = case
    y0 = np.array(y0).reshape(-1, 1)
    par = np.ones((1, 1)) if rhs == "harmonic" else None
    r = oracle.solve_ensemble(_abi.EULER, rhs, y0, par, dt_min=dt, dt_max=dt, tol=1.0, t_start=0.0, t_end=1.0,
                              history_capacity=400)
    m = int(r["hist_len"][0])
    assert r["status"][0] == _abi.OK and m == round(1.0 / dt) == r["n_accept"][0] == r["n_rhs"][0]
    t, y = r["hist_t"][0, :m], r["hist_y"][0, :m, 0]
    assert np.abs(y - exact(t)).max() <= eps  # the reference's assertion
    assert t[0] == 0.0 and y[0] == y0[0, 0] and t[-1] < 1.0 and r["t_end"][0] >= 1.0


# ---- path queries: the oracle's own statement against closed forms ------------------------------
def test_path_sampling_against_closed_forms(oracle):
    """Cubic Hermite between accepted points: the sampled harmonic oscillator stays within a small multiple of the
    error at the knots themselves; exact at the knots; NaN outside the path."""
    wv = np.array([1.0, 2.0, 3.5])
    y0 = np.stack([np.ones(3), np.zeros(3)])
    s = oracle.solve_ensemble(_abi.RK45, "harmonic", y0, wv.reshape(1, 3), dt_min=1e-9, dt_max=0.1, tol=1e-10,
                              t_start=0.0, t_end=5.0, history_capacity=4096)
    assert (s["status"] == _abi.OK).all()
    times = np.linspace(0.0, 5.0, 1001)
    got = oracle.sample_paths("harmonic", y0, wv.reshape(1, 3), s, times, t_start=0.0)
    for i, w in enumerate(wv):
        m = int(s["hist_len"][i])
        tk = s["hist"][i, :m, 0]
        knot_err = np.abs(s["hist"][i, :m, 1] - np.cos(w * tk)).max()
        assert np.abs(got[i, :, 0] - np.cos(w * times)).max() < 3 * knot_err + 1e-12
        at_knots = oracle.sample_paths("harmonic", y0, wv.reshape(1, 3), s, tk, t_start=0.0)[i]
        assert np.array_equal(at_knots, s["hist"][i, :m, 1:])
    edge = oracle.sample_paths("harmonic", y0, wv.reshape(1, 3), s, [-1e-9, 0.0, 5.0, 5.0 + 1e-9, np.nan], t_start=0.0)
    assert np.isnan(edge[:, [0, 3, 4]]).all()
    assert np.array_equal(edge[:, 1], y0.T) and np.array_equal(edge[:, 2], s["y_end"].T)


def test_path_events_against_closed_forms(oracle):
    """Zeros of cos(wt) and of its derivative, by direction; capacity overflow is counted, not stored."""
    wv = np.array([1.0, 2.0, 3.5])
    y0 = np.stack([np.ones(3), np.zeros(3)])
    p = wv.reshape(1, 3)
    s = oracle.solve_ensemble(_abi.RK45, "harmonic", y0, p, dt_min=1e-9, dt_max=0.1, tol=1e-10, t_start=0.0, t_end=5.0,
                              history_capacity=4096)
    ev, cnt = oracle.locate_events("harmonic", y0, p, s, [1.0, 0.0], 0.0, 0, 8, t_start=0.0)
    ev_f, cnt_f = oracle.locate_events("harmonic", y0, p, s, [1.0, 0.0], 0.0, -1, 8, t_start=0.0)
    ev_r, cnt_r = oracle.locate_events("harmonic", y0, p, s, [1.0, 0.0], 0.0, 1, 8, t_start=0.0)
    for i, w in enumerate(wv):
        exact = (2 * np.arange(64) + 1) * np.pi / (2 * w)
        exact = exact[exact < 5.0]
        assert cnt[i] == exact.size and cnt_f[i] + cnt_r[i] == cnt[i]
        assert np.abs(ev[i, :exact.size, 0] - exact).max() < 1e-8
        assert np.abs(ev[i, :exact.size, 1]).max() < 1e-12  # on the surface y = 0
        assert np.abs(ev_f[i, :cnt_f[i], 0] - exact[0::2]).max() < 1e-8  # cos falls through zero first
        assert np.abs(ev_r[i, :cnt_r[i], 0] - exact[1::2]).max(initial=0.0) < 1e-8
    ev1, cnt1 = oracle.locate_events("harmonic", y0, p, s, [1.0, 0.0], 0.0, 0, 1, t_start=0.0)
    np.testing.assert_array_equal(cnt1, cnt)
    assert np.array_equal(ev1[:, 0], ev[:, 0])
    # an exact zero at a knot belongs to the interval it closes, once: y' = -2t... use y = 1 - t^2 sampled through t = 1
    q = oracle.solve_ensemble(_abi.EULER, "quadratic", np.array([[1.0]]), None, dt_min=0.25, dt_max=0.25, tol=1e-3,
                              t_start=0.0, t_end=2.0, history_capacity=16)
    tk = q["hist"][0, :q["hist_len"][0], 0]
    yk = q["hist"][0, :q["hist_len"][0], 1]
    assert (yk == 0.25).any()  # Euler on y' = -2t with dt = 1/4 is exact in binary: knots at y = 1 - t(t - 1/4)
    ev, cnt = oracle.locate_events("quadratic", np.array([[1.0]]), None, q, [1.0], 0.25, 0, 4, t_start=0.0)
    assert cnt[0] == 1 and ev[0, 0, 0] == tk[yk == 0.25][0] and ev[0, 0, 1] == 0.25


def test_path_queries_against_scipy_dense_output_and_events(oracle):
    """Independent anchors (tests/golden/anchors.json, made by make_golden.py): SciPy DOP853's own dense output and its own
    event root-finder on seeded Lorenz trajectories.  The oracle's sampled states and located z = 27 crossings agree to
    the accuracy of the underlying RK45 solve (tol 1e-10; Lorenz amplifies errors by ~e^{0.9 t})."""
    anchors = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "anchors.json")))["lorenz_seeded4_T2_paths"]
    times = np.array(anchors["times"])
    y0 = E.lorenz_y0(np.arange(4))
    P = np.array(E.LORENZ["params"])
    s = oracle.solve_ensemble(_abi.RK45, "lorenz", y0, P, shared_params=True, dt_min=1e-9, dt_max=0.1, tol=1e-10,
                              t_start=0.0, t_end=2.0, history_capacity=8192)
    assert (s["status"] == _abi.OK).all()
    got = oracle.sample_paths("lorenz", y0, P, s, times, t_start=0.0, shared_params=True)
    ref = np.array(anchors["states"])
    assert np.abs(got - ref).max() < 1e-9 * np.abs(ref).max()  # (measured 3e-12)
    ev, cnt = oracle.locate_events("lorenz", y0, P, s, [0.0, 0.0, 1.0], 27.0, 0, 16, t_start=0.0, shared_params=True)
    for i in range(4):
        want = np.array(anchors["z27_events"][i])
        assert cnt[i] == want.size
        assert np.abs(ev[i, :want.size, 0] - want).max() < 1e-9  # (measured 9e-12)
        assert np.abs(ev[i, :want.size, 3] - 27.0).max() < 1e-10


def test_oracle_is_clean_under_sanitizers():
    """the reference is safe Rust; its C++ restatement is at least clean under AddressSanitizer and
    UndefinedBehaviorSanitizer on every stepper family, both semantics, dense output below and above the capacity, the
    optional inputs and the path queries (oracle/selftest.cpp)."""
    import os
    import subprocess
    here = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle")
    subprocess.run(["make", "-C", here, "_selftest"], check=True, capture_output=True)
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=1:abort_on_error=0", OMP_NUM_THREADS="2")
    r = subprocess.run([os.path.join(here, "_selftest")], capture_output=True, text=True, env=env)
    assert r.returncode == 0 and "selftest ok" in r.stdout, (r.stdout[-2000:], r.stderr[-4000:])


def test_coefficient_tables_are_the_reference_sources(oracle):
    """Every coefficient table of the oracle against the numbers PARSED OUT OF THE REFERENCE'S SOURCE TEXT
    (tests/golden/reference_coefficients.json, made in the build container by make_reference_coefficients.py from
    /root/reference/src/ivp/{rk,bdf,adams}.rs: the Rust table expressions evaluated in IEEE double, lists in source
    order) — bit for bit.  This pins the part of the restatement where a transcription slip would be silent: the RK
    stage matrix, weights and error weights, the BDF and Adams coefficients, and the three defects of the tables that
    REF_LITERAL keeps and REF_CORRECTED repairs (SURVEY.md D1 column-major fill of the row-listed matrix, D2
    1859/4014 for Fehlberg's 1859/4104, D3 a safety factor built from 100/100)."""
    import json
    import os
    from bacon_b200 import _abi
    ref = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_coefficients.json")))

    def bits(x):
        return np.asarray(x, dtype=np.float64).view(np.uint64)

    for name, method, o in (("RK45", _abi.RK45, 6), ("RK23", _abi.RK23, 4)):
        listed = np.array(ref[name]["k_coefficients"]["values"]).reshape(o, o)  # listed[i][j] = entry i*O + j of the source
        for literal in (True, False):
            t = oracle.coefficients(method, literal)
            assert np.array_equal(bits(t["c"]), bits(ref[name]["t_coefficients"]["values"]))
            assert np.array_equal(bits(t["b"]), bits(ref[name]["avg_coefficients"]["values"]))
            assert np.array_equal(bits(t["e"]), bits(ref[name]["error_coefficients"]["values"]))
            if literal:
                # BSMatrix::from_vec fills column by column: M(r, c) = source[c*O + r] (D1), and the stepper reads rows of M
                assert np.array_equal(bits(t["A"]), bits(listed.T))
                assert t["safety"] == ref["RK_safety"]["values"][0] == 1.0  # D3
            else:
                want = listed.copy()  # the rows as the source's "Row i" comments label them
                if name == "RK45":
                    assert want[5, 3] == 1859.0 / 4014.0  # D2 as written ...
                    want[5, 3] = 1859.0 / 4104.0           # ... and Fehlberg's coefficient
                assert np.array_equal(bits(t["A"]), bits(want))
                assert t["safety"] == 84.0 / 100.0
        # consistency of the CORRECTED tableau (what D1/D2 break): row sums = nodes, weights sum to one, error weights to zero
        t = oracle.coefficients(method, False)
        np.testing.assert_allclose(t["A"].sum(axis=1), t["c"], rtol=0, atol=4e-16)
        assert abs(t["b"].sum() - 1.0) <= 4e-16 and abs(t["e"].sum()) <= 4e-16
    for name, method in (("BDF6", _abi.BDF6), ("BDF2", _abi.BDF2)):
        t = oracle.coefficients(method)
        assert np.array_equal(bits(t["higher"]), bits(ref[name]["higher_coefficients"]["values"]))
        assert np.array_equal(bits(t["lower"]), bits(ref[name]["lower_coefficients"]["values"]))
    for name, method in (("Adams5", _abi.ADAMS5), ("Adams3", _abi.ADAMS3)):
        t = oracle.coefficients(method)
        assert np.array_equal(bits(t["predictor"]), bits(ref[name]["predictor_coefficients"]["values"]))
        assert np.array_equal(bits(t["corrector"]), bits(ref[name]["corrector_coefficients"]["values"]))
        assert t["error"] == ref[name]["error_coefficient"]["values"][0] == 19.0 / 270.0


@pytest.mark.parametrize("as_written", [True, False])
def test_rk_second_reading_agrees_bit_for_bit_on_y_dependent_problems(oracle, as_written):
    """tests/rk_second_reading.py — rk.rs:249-423 and ivp.rs:220-238 read a second time, in plain Python over the
    coefficient lists parsed out of the reference's source — against the oracle on the problems no reference test
    covers: y-DEPENDENT right-hand sides (Lorenz, Van der Pol with per-trajectory mu), both pairs, both semantics.
    Every yielded (time, state), the step counts and the exit dt, bit for bit."""
    import rk_second_reading as R2
    sem = _abi.SEM_LITERAL if as_written else _abi.SEM_CORRECTED
    cases = []
    y0 = E.lorenz_y0(np.arange(6))
    P = np.array(E.LORENZ["params"])
    cases.append(("RK45", _abi.RK45, "lorenz", R2.lorenz, y0, np.tile(P[:, None], (1, 6)), dict(dt_min=1e-9, dt_max=0.1, tol=1e-8, t_start=0.0, t_end=0.4)))
    cases.append(("RK23", _abi.RK23, "lorenz", R2.lorenz, y0, np.tile(P[:, None], (1, 6)),
                  dict(dt_min=1e-9, dt_max=0.1, tol=1e-5, t_start=0.0, t_end=0.05)))
    yv, mu = E.vdp_problem(np.arange(6) * (E.VDP["n"] // 6), E.VDP["n"])
    cases.append(("RK23", _abi.RK23, "vdp", R2.vdp, yv, mu, dict(dt_min=1e-12, dt_max=0.1, tol=1e-8, t_start=0.0, t_end=0.05)))
    cases.append(("RK45", _abi.RK45, "vdp", R2.vdp, yv, mu, dict(dt_min=1e-12, dt_max=0.1, tol=1e-9, t_start=0.0, t_end=0.3)))
    for name, method, rhs, f, y, p, cfg in cases:
        cap = 4096
        # (as written, the tables can make a path crawl at dt ~ 1e-9 for 1e8 attempts: both readings are cut short there —
        # the oracle by its attempt cap, the second reading after 600 points — and the common prefix is compared)
        ref = oracle.solve_ensemble(method, rhs, y, p, semantics=sem, history_capacity=cap, pow_mode=0,
                                    max_attempts=4000 if as_written else 0, **cfg)
        for i in range(y.shape[1]):
            path, status, (n_acc, n_rej), (t_fin, dt_fin, y_fin) = R2.solve(name, f, list(y[:, i]), list(p[:, i]), as_written=as_written,
                                                                           max_points=600 if as_written else cap, **cfg)
            m = min(int(ref["hist_len"][i]), len(path))
            assert m > (0 if as_written else 10), (name, rhs, i)
            t2 = np.array([q[0] for q in path[:m]])
            y2 = np.array([q[1] for q in path[:m]], dtype=np.float64)
            assert np.array_equal(t2.view(np.uint64), ref["hist_t"][i, :m].view(np.uint64)), (name, rhs, i)
            assert np.array_equal(y2.view(np.uint64), np.ascontiguousarray(ref["hist_y"][i, :m]).view(np.uint64)), (name, rhs, i)
            if as_written and (status == "Truncated" or ref["status"][i] == _abi.E_MAX_ATTEMPTS):
                continue
            assert status != "Truncated", (name, rhs, i, len(path))
            assert status == ("Done" if ref["status"][i] == _abi.OK else "MinimumTimeDeltaExceeded"), (name, rhs, i)
            assert n_acc == ref["n_accept"][i] and n_rej == ref["n_reject"][i], (name, rhs, i, n_acc, n_rej)
            assert len(path) == int(ref["hist_len"][i])
            assert np.array_equal(np.array(y_fin, dtype=np.float64).view(np.uint64), np.ascontiguousarray(ref["y_end"][:, i]).view(np.uint64))
            assert np.float64(dt_fin).view(np.uint64) == ref["dt_end"][i].view(np.uint64) and t_fin == ref["t_end"][i]


@pytest.mark.parametrize("as_written", [True, False])
def test_bdf_second_reading_agrees_bit_for_bit_in_one_dimension(oracle, as_written):
    """tests/bdf_second_reading.py — bdf.rs:257-634 and ivp.rs:220-238 read a second time, in plain Python over the parsed
    coefficient lists, for one-dimensional problems (where the matrix inverse nalgebra provides is 1 / x for every
    algorithm) — against the oracle: start-up blocks, the speculative implicit step and its rollback, the yield
    bookkeeping, both implicit functions, Broyden, halving and doubling.  The reference's own BDF test problems
    (bdf.rs:769-783: y' = y, -y, -2t, cos t — two of them non-autonomous, which is where D7 shows) with BDF6 and BDF2,
    settings that make the steppers halve, double and roll back.  Every yielded (time, state), the exit time, state
    and dt, the status: bit for bit, as written (where the paths are empty and the exit times are the 7.3 / 14.45 ... of
    SURVEY.md section 4) and corrected."""
    import math
    import bdf_second_reading as B2
    sem = _abi.SEM_LITERAL if as_written else _abi.SEM_CORRECTED
    fs = {"decay": lambda t, y: -y, "exp": lambda t, y: y, "quadratic": lambda t, y: -2.0 * t, "cos": lambda t, y: math.cos(t)}
    names = {0: "Done", _abi.E_MIN_DT_EXCEEDED: "MinimumTimeDeltaExceeded", _abi.E_MAX_ITER: "MaximumIterationsExceeded",
             _abi.E_SINGULAR: "SingularMatrix"}
    y0 = np.array([[1.0, 0.7, 1.3, -0.4]])
    compared = rolled_back = 0
    for cfg in (dict(dt_min=1e-7, dt_max=0.01, tol=1e-6, t_start=0.0, t_end=0.5),
                dict(dt_min=1e-9, dt_max=0.1, tol=1e-9, t_start=0.0, t_end=1.0),
                dict(dt_min=1e-4, dt_max=0.05, tol=1e-10, t_start=0.25, t_end=0.9)):
        for method, name in ((_abi.BDF6, "BDF6"), (_abi.BDF2, "BDF2")):
            for rhs, f in fs.items():
                ref = oracle.solve_ensemble(method, rhs, y0, None, semantics=sem, history_capacity=1 << 15, max_attempts=400000, **cfg)
                for i in range(y0.shape[1]):
                    path, status, (t_fin, dt_fin, y_fin) = B2.solve(name, f, float(y0[0, i]), as_written=as_written, max_points=1 << 15,
                                                                    max_calls=2000000, **cfg)
                    key = (name, rhs, i, cfg["tol"])
                    m = min(int(ref["hist_len"][i]), len(path))
                    t2 = np.array([q[0] for q in path[:m]], dtype=np.float64)
                    y2 = np.array([q[1] for q in path[:m]], dtype=np.float64)
                    assert np.array_equal(t2.view(np.uint64), ref["hist_t"][i, :m].view(np.uint64)), key
                    assert np.array_equal(y2.view(np.uint64), np.ascontiguousarray(ref["hist_y"][i, :m, 0]).view(np.uint64)), key
                    compared += m
                    if status == "Truncated" or int(ref["status"][i]) == _abi.E_HISTORY_OVERFLOW:  # (BDF2 at a tight tolerance: > 32768 points;
                        assert m == 1 << 15                                                         #  the first 32768 are compared)
                        continue
                    assert status == names.get(int(ref["status"][i]), "?"), key + (status, int(ref["status"][i]))
                    assert len(path) == int(ref["hist_len"][i]) == int(ref["n_accept"][i]), key
                    assert np.float64(t_fin).view(np.uint64) == ref["t_end"][i].view(np.uint64), key
                    assert np.float64(y_fin).view(np.uint64) == ref["y_end"][0, i].view(np.uint64), key
                    assert np.float64(dt_fin).view(np.uint64) == ref["dt_end"][i].view(np.uint64), key
                    rolled_back += int(dt_fin < 0.25 * (cfg["dt_max"] + cfg["dt_min"]))
    if as_written:
        assert compared == 0 or compared > 0  # (as written every BDF path of these problems is empty or short: the exit records carry the check)
    else:
        assert compared > 20000 and rolled_back > 0


@pytest.mark.parametrize("as_written", [True, False])
def test_adams_second_reading_agrees_bit_for_bit(oracle, as_written):
    """tests/adams_second_reading.py — adams.rs:249-561 and ivp.rs:220-238 read a second time, in plain Python over the
    parsed coefficient lists — against the oracle on y-dependent problems in two and three dimensions (the harmonic
    oscillator with per-trajectory frequency, Lorenz), Adams5 and Adams3: every yielded (time, state), the exit time,
    state and dt, the status, bit for bit; as written (where D10 makes every start-up block end in a reject and the
    paths grow like 1/tol: the first 4096 points are compared) and with D10 repaired."""
    import adams_second_reading as A2
    import rk_second_reading as R2
    sem = _abi.SEM_LITERAL if as_written else _abi.SEM_CORRECTED
    cap = 4096

    def harmonic(_t, y, p):
        return [y[1], -(p[0] * p[0]) * y[0]]

    cases = [("harmonic", harmonic, np.array([[1.0, 0.5, -0.8], [0.0, 0.3, 1.1]]), np.array([[2.0, 3.0, 0.7]]),
              dict(dt_min=1e-6, dt_max=0.1, tol=1e-6, t_start=0.0, t_end=1.0)),
             ("lorenz", R2.lorenz, E.lorenz_y0(np.arange(3)), np.tile(np.array(E.LORENZ["params"])[:, None], (1, 3)),
              dict(dt_min=1e-9, dt_max=0.01, tol=1e-4, t_start=0.0, t_end=0.2))]
    compared = 0
    for method, name in ((_abi.ADAMS5, "Adams5"), (_abi.ADAMS3, "Adams3")):
        for rhs, f, y0, p, cfg in cases:
            ref = oracle.solve_ensemble(method, rhs, y0, p, semantics=sem, history_capacity=cap, pow_mode=0,
                                        max_attempts=60000 if as_written else 0, **cfg)
            for i in range(y0.shape[1]):
                path, status, (t_fin, dt_fin, y_fin) = A2.solve(name, f, list(y0[:, i]), list(p[:, i]), as_written=as_written,
                                                                max_points=cap, **cfg)
                key = (name, rhs, i)
                m = min(len(path), int(ref["hist_len"][i]))
                assert m > 10, key
                t2 = np.array([q[0] for q in path[:m]], dtype=np.float64)
                y2 = np.array([q[1] for q in path[:m]], dtype=np.float64)
                assert np.array_equal(t2.view(np.uint64), ref["hist_t"][i, :m].view(np.uint64)), key
                assert np.array_equal(y2.view(np.uint64), np.ascontiguousarray(ref["hist_y"][i, :m]).view(np.uint64)), key
                compared += m
                if status == "Truncated" or int(ref["status"][i]) in (_abi.E_HISTORY_OVERFLOW, _abi.E_MAX_ATTEMPTS):
                    assert as_written, key  # (only the as-written paths are that long)
                    continue
                assert status == "Done" and int(ref["status"][i]) == _abi.OK, key
                assert len(path) == int(ref["hist_len"][i]) == int(ref["n_accept"][i]), key
                assert np.float64(t_fin).view(np.uint64) == ref["t_end"][i].view(np.uint64), key
                assert np.array_equal(np.array(y_fin, dtype=np.float64).view(np.uint64), np.ascontiguousarray(ref["y_end"][:, i]).view(np.uint64)), key
                assert np.float64(dt_fin).view(np.uint64) == ref["dt_end"][i].view(np.uint64), key
    assert compared > 5000
