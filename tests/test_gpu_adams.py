"""GPU parity of the Adams predictor-corrector and Euler paths and of the README's
`solve_ivp` fallback chain against the CPU oracle, through the C ABI.

strict kernels : bit-exact with the oracle (same operation order, the shared deterministic x^(1/order))
fast kernels   : FMA contraction + SFU root -> final state inside the band, step counts side by side
"""
import numpy as np
import pytest

from bacon_b200 import _abi, ensembles as E
from parity import band, make_solver, rel_err, run_both
from reference_cases import ADAMS_CASES, EULER_CASES

pytestmark = pytest.mark.gpu


def _bit_exact(gpu, ref, keys=("y_end", "t_end", "dt_end")):
    for k in ("status", "n_accept", "n_reject", "n_rhs"):
        np.testing.assert_array_equal(getattr(gpu, k), ref[k], err_msg=k)
    for k in keys:
        a, b = getattr(gpu, k), ref[k]
        assert np.array_equal(a.view(np.uint64), b.view(np.uint64)), f"{k}: max |d| = {np.abs(a - b).max()}"


def _hist_equal(gpu, ref, exact_bits=True):
    np.testing.assert_array_equal(gpu.hist_len, ref["hist_len"])
    cap = gpu.hist_t.shape[1]
    mask = np.arange(cap)[None, :] < gpu.hist_len[:, None]
    if exact_bits:
        assert np.array_equal(gpu.hist_t[mask], ref["hist_t"][mask]) and np.array_equal(gpu.hist_y[mask], ref["hist_y"][mask])
    else:
        np.testing.assert_allclose(gpu.hist_t[mask], ref["hist_t"][mask], rtol=1e-13, atol=1e-15)
        np.testing.assert_allclose(gpu.hist_y[mask], ref["hist_y"][mask], rtol=1e-12, atol=1e-14)


@pytest.mark.parametrize("case", ADAMS_CASES, ids=[c[0] for c in ADAMS_CASES])
def test_reference_adams_tests_on_gpu(cuda, engine, oracle, case):
    """adams.rs:714-922 in both semantics: strict kernel bit-exact with the oracle (whole path), every yielded
    point inside the reference's epsilon; the fast kernel meets the same assertion."""
    name, method, rhs, y0, cfg, exact, eps, lit, cor = case
    y0 = np.array([[y0]])
    for sem, (n_yield, n_rej) in ((_abi.SEM_LITERAL, lit), (_abi.SEM_CORRECTED, cor)):
        gpu, ref = run_both(engine, oracle, method, rhs, y0, strict=True, semantics=sem, history=8000, **cfg)
        m = int(gpu.hist_len[0])
        assert gpu.status[0] == _abi.OK and (m, int(gpu.n_reject[0])) == (n_yield, n_rej)
        if rhs == "cos":  # device cos() and glibc's differ in the last ulp
            for k in ("status", "n_accept", "n_reject", "n_rhs"):
                np.testing.assert_array_equal(getattr(gpu, k), ref[k], err_msg=k)
            _hist_equal(gpu, ref, exact_bits=False)
            np.testing.assert_allclose(gpu.y_end, ref["y_end"], rtol=1e-12, atol=1e-14)
        else:
            _bit_exact(gpu, ref)
            _hist_equal(gpu, ref)
        t, y = gpu.hist_t[0, :m], gpu.hist_y[0, :m, 0]
        assert np.abs(y - exact(t)).max() <= eps  # the reference's assertion
    # fast kernel, REF_CORRECTED
    s = make_solver(engine, method, 1, rhs=rhs, history=8000, **cfg)
    r = s.solve_ivp_ensemble(y0)
    m = int(r.hist_len[0])
    assert r.status[0] == _abi.OK and m > 0 and abs(m - cor[0]) <= 8
    assert np.abs(r.hist_y[0, :m, 0] - exact(r.hist_t[0, :m])).max() <= eps and r.t_end[0] >= cfg["t_end"]


@pytest.mark.parametrize("method,tol", [("Adams5", 1e-8), ("Adams3", 1e-6)])
def test_adams_ensembles_strict_bit_exact_and_fast_in_band(cuda, engine, oracle, method, tol):
    """y-dependent systems (D = 2 and 3, per-trajectory parameters): harmonic oscillators, Van der Pol, Lorenz."""
    n = 2048
    rng = np.random.default_rng(5)
    w = rng.uniform(0.5, 3.0, (1, n))
    cases = [("harmonic", np.stack([np.ones(n), np.zeros(n)]), w, False, dict(dt_min=1e-9, dt_max=0.05, t_end=3.0)),
             ("vdp", np.stack([np.full(n, 2.0), np.zeros(n)]), rng.uniform(0.1, 5.0, (1, n)), False,
              dict(dt_min=1e-10, dt_max=0.05, t_end=0.5)),
             ("lorenz", E.lorenz_y0(np.arange(n)), np.array(E.LORENZ["params"]), True, dict(dt_min=1e-10, dt_max=0.05, t_end=0.5))]
    for rhs, y0, p, shared, c in cases:
        cfg = dict(tol=tol, t_start=0.0, **c)
        gpu, ref = run_both(engine, oracle, method, rhs, y0, p, shared_params=shared, strict=True, **cfg)
        assert (gpu.status == _abi.OK).all(), rhs
        _bit_exact(gpu, ref)
        fast, _ = run_both(engine, oracle, method, rhs, y0, p, shared_params=shared, **cfg)
        assert (fast.status == _abi.OK).all()
        assert rel_err(fast.y_end, ref["y_end"]).max() <= band(tol), rhs
        # step counts side by side: a flipped accept/reject at a threshold costs at most a few warm-up blocks
        d_acc = np.abs(fast.n_accept.astype(int) - ref["n_accept"].astype(int))
        assert np.median(d_acc) <= 1 and d_acc.max() <= 0.05 * ref["n_accept"].max() + 16, (rhs, d_acc.max())
    # closed form for the oscillators
    exact = np.stack([np.cos(3.0 * w[0]), -w[0] * np.sin(3.0 * w[0])])
    s = make_solver(engine, method, 2, rhs="harmonic", tol=tol, t_start=0.0, dt_min=1e-9, dt_max=0.05, t_end=3.0)
    r = s.solve_ivp_ensemble(cases[0][1], w)
    assert np.abs(r.y_end - exact).max() < 20 * tol


def test_adams_literal_d10_on_gpu(cuda, engine, oracle):
    """REF_LITERAL (the source as written): the first regular step after every warm-up uses stale derivatives
    (D10) -> ~1/tol steps.  Same bits as the oracle, same answer, two orders of magnitude more work."""
    n = 64
    w = np.linspace(2.0, 3.0, n)[None, :]
    y0 = np.stack([np.ones(n), np.zeros(n)])
    cfg = dict(dt_min=1e-8, dt_max=0.05, tol=1e-5, t_start=0.0, t_end=1.0)
    lit, ref = run_both(engine, oracle, "Adams5", "harmonic", y0, w, semantics=_abi.SEM_LITERAL, **cfg)
    _bit_exact(lit, ref)
    cor, _ = run_both(engine, oracle, "Adams5", "harmonic", y0, w, **cfg)
    assert (lit.status == _abi.OK).all() and (lit.n_accept > 50 * cor.n_accept).all()
    assert rel_err(lit.y_end, cor.y_end).max() < 1e-3


def test_adams_dense_output_failures_and_edges(cuda, engine, oracle):
    n = 777  # ragged: not a multiple of the warp or CTA size
    rng = np.random.default_rng(9)
    w = rng.uniform(0.5, 3.0, (1, n))
    y0 = np.stack([np.ones(n), np.zeros(n)])
    cfg = dict(dt_min=1e-9, dt_max=0.05, tol=1e-7, t_start=0.0, t_end=2.0)
    gpu, ref = run_both(engine, oracle, "Adams5", "harmonic", y0, w, strict=True, history=64, **cfg)
    _bit_exact(gpu, ref)
    _hist_equal(gpu, ref)
    assert (gpu.status == _abi.E_HISTORY_OVERFLOW).any() and (gpu.status == _abi.OK).any()  # capacity 64 is tight on purpose
    # dt_min too large -> MinimumTimeDeltaExceeded on the stiffer oscillators, per trajectory, same as the oracle
    gpu, ref = run_both(engine, oracle, "Adams3", "harmonic", y0, w * 40.0, strict=True, dt_min=2e-3, dt_max=0.05, tol=1e-6,
                        t_start=0.0, t_end=2.0)
    _bit_exact(gpu, ref)
    assert (gpu.status == _abi.E_MIN_DT_EXCEEDED).any()
    # attempt cap, NaN input, n = 1, n = 0
    gpu, ref = run_both(engine, oracle, "Adams5", "harmonic", y0[:, :33], w[:, :33], strict=True, max_attempts=40, **cfg)
    _bit_exact(gpu, ref)
    assert (gpu.status == _abi.E_MAX_ATTEMPTS).all()
    bad = y0[:, :5].copy()
    bad[0, 2] = np.nan
    gpu, ref = run_both(engine, oracle, "Adams5", "harmonic", bad, w[:, :5], strict=True, **cfg)
    np.testing.assert_array_equal(gpu.status, ref["status"])
    assert gpu.status[2] == _abi.E_NONFINITE and (np.delete(gpu.status, 2) == _abi.OK).all()
    s = make_solver(engine, "Adams5", 2, rhs="harmonic", **cfg)
    assert s.solve_ivp_ensemble(np.zeros((2, 0)), np.zeros((1, 0))).status.shape == (0,)
    one = s.solve_ivp_ensemble(y0[:, :1], w[:, :1])
    assert one.status[0] == _abi.OK


@pytest.mark.parametrize("case", EULER_CASES, ids=[c[0] for c in EULER_CASES])
def test_reference_euler_tests_on_gpu(cuda, engine, oracle, case):
    """ivp.rs:539-653: the path starts at the initial condition and never holds the final state."""
    name, rhs, y0, dt, exact, eps = case
    y0 = np.array(y0).reshape(-1, 1)
    par = np.ones((1, 1)) if rhs == "harmonic" else None
    dim = y0.shape[0]
    ref = oracle.solve_ensemble(_abi.EULER, rhs, y0, par, dt_min=dt, dt_max=dt, tol=1.0, t_start=0.0, t_end=1.0,
                                history_capacity=400)
    for strict in (True, False):
        # the builder the reference's helper uses (ivp.rs:497-512): only with_maximum_dt, no tolerance, no dt_min
        s = (engine.Euler.new(dim).with_initial_time(0.0).with_ending_time(1.0).with_maximum_dt(dt).with_derivative(rhs)
             .with_flags(_abi.FLAG_STRICT_FP if strict else 0).with_history(400))
        r = s.solve_ivp_ensemble(y0, par)
        m = int(r.hist_len[0])
        assert r.status[0] == _abi.OK and m == round(1.0 / dt) == r.n_accept[0] == r.n_rhs[0] == ref["n_accept"][0]
        t, y = r.hist_t[0, :m], r.hist_y[0, :m, 0]
        assert np.abs(y - exact(t)).max() <= eps  # the reference's assertion
        assert t[0] == 0.0 and y[0] == y0[0, 0] and t[-1] < 1.0 and r.t_end[0] >= 1.0
        if strict and rhs != "cos":
            assert np.array_equal(r.hist_t[0, :m], ref["hist_t"][0, :m]) and np.array_equal(r.hist_y[0, :m], ref["hist_y"][0, :m])
            assert np.array_equal(r.y_end, ref["y_end"])
        else:
            np.testing.assert_allclose(r.hist_y[0, :m], ref["hist_y"][0, :m], rtol=1e-12, atol=1e-14)


def test_euler_builder_and_ensemble(cuda, engine, oracle):
    # dt = average of the bounds given (ivp.rs:396-421); with_tolerance is a no-op, even for a bad value
    s = engine.Euler.new(3).with_tolerance(-1.0).with_maximum_dt(0.004).with_minimum_dt(0.002)
    with pytest.raises(engine.IVPError) as e:
        s.with_derivative("lorenz").solve_ivp_ensemble(np.ones((3, 1)), np.array(E.LORENZ["params"]), shared_params=True)
    assert e.value.variant == "MissingParameters"  # no initial / ending time yet (ivp.rs:451-459)
    with pytest.raises(engine.IVPError) as e:
        engine.Euler.new(1).with_maximum_dt(0.0)
    assert e.value.variant == "TimeDeltaOOB"
    s = s.with_initial_time(0.0).with_ending_time(0.5)
    n = 3000
    y0 = E.lorenz_y0(np.arange(n))
    p = np.array(E.LORENZ["params"])
    r = s.with_flags(_abi.FLAG_STRICT_FP).solve_ivp_ensemble(y0, p, shared_params=True)
    ref = oracle.solve_ensemble(_abi.EULER, "lorenz", y0, p, shared_params=True, dt_min=0.003, dt_max=0.003, tol=1.0,
                                t_start=0.0, t_end=0.5)
    assert (r.status == _abi.OK).all() and (r.n_accept == 167).all()
    _bit_exact(r, ref)
    fast = s.with_flags(0).solve_ivp_ensemble(y0, p, shared_params=True)
    assert rel_err(fast.y_end, ref["y_end"]).max() < 1e-11


def test_solve_ivp_fallback_chain(cuda, engine, oracle):
    """README.md:45-47: Adams5, then RK45, then BDF6, per trajectory.  Oscillators whose frequency spans
    three decades: dt_min is too large for Adams5 on the fast ones and for RK45 on the fastest."""
    n = 96
    w = np.logspace(0.0, 3.3, n)[None, :]
    y0 = np.stack([np.ones(n), np.zeros(n)])
    kw = dict(dt_min=2e-4, dt_max=0.05, tol=1e-6, t_start=0.0, t_end=0.25)
    res = engine.solve_ivp("harmonic", y0, w, t_span=(0.0, 0.25), dt_min=kw["dt_min"], dt_max=kw["dt_max"], tolerance=kw["tol"])
    # the same chain through the oracle
    status = np.full(n, -1)
    y_end = np.zeros((2, n))
    method = np.zeros(n, dtype=int)
    todo = np.arange(n)
    for k, m in enumerate((_abi.ADAMS5, _abi.RK45, _abi.BDF6)):
        if todo.size == 0:
            break
        r = oracle.solve_ensemble(m, "harmonic", y0[:, todo], w[:, todo], **kw)
        status[todo], y_end[:, todo], method[todo] = r["status"], r["y_end"], k
        todo = todo[r["status"] != _abi.OK]
    np.testing.assert_array_equal(res.status, status)
    np.testing.assert_array_equal(res.method, method)
    assert set(np.unique(res.method)) == {0, 1, 2}, np.bincount(res.method)
    ok = res.status == _abi.OK
    assert ok.sum() > n // 2
    assert rel_err(res.y_end[:, ok], y_end[:, ok]).max() <= band(1e-6)
    exact = np.stack([np.cos(0.25 * w[0]), -w[0] * np.sin(0.25 * w[0])])
    assert rel_err(res.y_end[:, ok], exact[:, ok]).max() < 1e-3
