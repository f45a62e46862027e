"""GPU parity of the BDF path (K3) against the CPU oracle, through the C ABI.

strict Broyden : bit-exact with the oracle (the reference's own iteration, same operation order)
fast Broyden   : same algorithm with FMA contraction -> final state inside the band
Newton + LU    : north-star item 4; against the oracle's Newton twin (tight) and its Broyden (band)
"""
import json
import os

import numpy as np
import pytest

from bacon_b200 import _abi, ensembles as E
from parity import METHODS, band, make_solver, rel_err, run_both
from reference_cases import BDF_CASES, bdf_cfg

pytestmark = pytest.mark.gpu
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "anchors.json")))
ROB = dict(dt_min=1e-10, dt_max=1e-4, tol=1e-6, t_start=0.0)


def _bit_exact(gpu, ref, keys=("y_end", "t_end", "dt_end")):
    for k in ("status", "n_accept", "n_reject", "n_rhs"):
        np.testing.assert_array_equal(getattr(gpu, k), ref[k], err_msg=k)
    for k in keys:
        a, b = getattr(gpu, k), ref[k]
        assert np.array_equal(a.view(np.uint64), b.view(np.uint64)), f"{k}: max |d| = {np.abs(a - b).max()}"


@pytest.mark.parametrize("case", BDF_CASES, ids=[c[0] for c in BDF_CASES])
def test_reference_bdf_tests_on_gpu(cuda, engine, oracle, case):
    """bdf.rs:785-1063 in REF_CORRECTED: strict kernel bit-exact with the oracle, every yielded point
    inside the reference's epsilon; fast and Newton variants meet the same assertion."""
    name, method, rhs, y0, t_end, exact, eps, n_yield, _ = case
    cap = 25000
    y0 = np.array([[y0]])
    gpu, ref = run_both(engine, oracle, method, rhs, y0, strict=True, history=cap, **bdf_cfg(t_end))
    m = int(gpu.hist_len[0])
    assert m == ref["hist_len"][0] and (n_yield is None or m == n_yield)
    if rhs == "cos":
        # the device cos() and glibc's differ in the last ulp, so this one problem is compared to 1e-12, not bitwise
        for k in ("status", "n_accept", "n_reject", "n_rhs"):
            np.testing.assert_array_equal(getattr(gpu, k), ref[k], err_msg=k)
        np.testing.assert_array_equal(gpu.hist_t[0, :m], ref["hist_t"][0, :m])
        np.testing.assert_allclose(gpu.hist_y[0, :m], ref["hist_y"][0, :m], rtol=1e-12, atol=1e-14)
        np.testing.assert_allclose(gpu.y_end, ref["y_end"], rtol=1e-12, atol=1e-14)
    else:
        _bit_exact(gpu, ref)
        assert np.array_equal(gpu.hist_t[0, :m], ref["hist_t"][0, :m]) and np.array_equal(gpu.hist_y[0, :m], ref["hist_y"][0, :m])
    for flags in (0, _abi.FLAG_BDF_NEWTON):
        s = make_solver(engine, method, 1, rhs=rhs, flags=flags, history=cap, **bdf_cfg(t_end))
        r = s.solve_ivp_ensemble(y0)
        assert r.status[0] == _abi.OK
        m = int(r.hist_len[0])
        t, y = r.hist_t[0, :m], r.hist_y[0, :m, 0]
        # the stepper reaches t_end exactly; the last YIELDED point may be earlier (warm-up block pending at Done)
        assert m > 0 and np.abs(y - exact(t)).max() <= eps and abs(r.t_end[0] - t_end) <= 1e-12 * t_end


def test_robertson_strict_bit_exact(cuda, engine, oracle):
    n = 600
    y0, k = E.robertson_problem(np.arange(n))
    gpu, ref = run_both(engine, oracle, "BDF6", "robertson", y0, k, strict=True, t_end=0.01, **ROB)
    assert (gpu.status == _abi.OK).all()
    _bit_exact(gpu, ref)
    gpu, ref = run_both(engine, oracle, "BDF2", "robertson", y0[:, :64], k[:, :64], strict=True, t_end=0.002, **ROB)
    _bit_exact(gpu, ref)


def test_robertson_fast_broyden_and_newton_within_band(cuda, engine, oracle):
    """Config 5 shape (k perturbed +-10 % per trajectory), T = 0.02."""
    n = 4096
    y0, k = E.robertson_problem(np.arange(n))
    cfg = dict(t_end=0.02, **ROB)
    gpu, ref = run_both(engine, oracle, "BDF6", "robertson", y0, k, **cfg)
    assert (gpu.status == _abi.OK).all() and (ref["status"] == _abi.OK).all()
    assert rel_err(gpu.y_end, ref["y_end"]).max() <= band(1e-6)
    np.testing.assert_array_equal(gpu.t_end, ref["t_end"])
    assert np.abs(gpu.n_accept.astype(int) - ref["n_accept"].astype(int)).max() <= 16  # one restart block = 7+1 points
    # Newton + in-register LU: against the oracle's Newton twin and against the reference-style Broyden oracle
    s = make_solver(engine, "BDF6", 3, rhs="robertson", flags=_abi.FLAG_BDF_NEWTON, **cfg)
    nw = s.solve_ivp_ensemble(y0, k)
    ref_nw = oracle.solve_ensemble(_abi.BDF6, "robertson", y0, k, bdf_newton=True, **cfg)
    assert (nw.status == _abi.OK).all()
    assert rel_err(nw.y_end, ref_nw["y_end"]).max() <= band(1e-6)
    assert rel_err(nw.y_end, ref["y_end"]).max() <= band(1e-6)
    assert (nw.n_rhs < gpu.n_rhs).all()  # no 2D finite-difference evaluations per solve
    # mass conservation of the kinetics (sum y = 1) survives the implicit solves
    assert np.abs(nw.y_end.sum(0) - 1.0).max() < 1e-9 and np.abs(gpu.y_end.sum(0) - 1.0).max() < 1e-9


def test_robertson_T05_against_radau_anchor(cuda, engine):
    """y(0.5) = (0.981791774, 3.32809109e-5, 0.0181749452), 5005 yielded points."""
    y0 = np.array([[1.0], [0.0], [0.0]])
    k = np.array([0.04, 3e7, 1e4])
    for flags in (0, _abi.FLAG_STRICT_FP, _abi.FLAG_BDF_NEWTON):
        s = make_solver(engine, "BDF6", 3, rhs="robertson", flags=flags, t_end=0.5, **ROB)
        r = s.solve_ivp_ensemble(y0, k, shared_params=True)
        assert r.status[0] == _abi.OK and r.n_accept[0] == 5005
        np.testing.assert_allclose(r.y_end[:, 0], GOLD["robertson_T05"], rtol=2e-8)


def test_bdf_dense_output_and_failures(cuda, engine, oracle):
    n = 500
    y0, k = E.robertson_problem(np.arange(n))
    gpu, ref = run_both(engine, oracle, "BDF6", "robertson", y0, k, strict=True, history=128, t_end=0.004, **ROB)
    _bit_exact(gpu, ref)
    np.testing.assert_array_equal(gpu.hist_len, ref["hist_len"])
    mask = np.arange(128)[None, :] < gpu.hist_len[:, None]
    assert np.array_equal(gpu.hist_t[mask], ref["hist_t"][mask]) and np.array_equal(gpu.hist_y[mask], ref["hist_y"][mask])
    # dt_min above what the stiff problem needs -> MinimumTimeDeltaExceeded, same as the oracle
    gpu, ref = run_both(engine, oracle, "BDF6", "robertson", y0[:, :64], k[:, :64], strict=True, dt_min=1e-3, dt_max=1e-2,
                        tol=1e-8, t_start=0.0, t_end=1.0)
    np.testing.assert_array_equal(gpu.status, ref["status"])
    assert (gpu.status != _abi.OK).all()
    # REF_LITERAL BDF (the source as written) runs in the strict Broyden build: the oracle's bits, Robertson included
    # (tests/test_gpu_options.py replays the reference's own BDF tests in this mode); the Newton variant has no LITERAL form
    gpu, ref = run_both(engine, oracle, "BDF6", "robertson", y0[:, :64], k[:, :64], semantics=_abi.SEM_LITERAL,
                        max_attempts=3000, t_end=0.004, **ROB)
    _bit_exact(gpu, ref)
    s = make_solver(engine, "BDF6", 3, rhs="robertson", semantics=_abi.SEM_LITERAL, flags=_abi.FLAG_BDF_NEWTON, t_end=0.01, **ROB)
    with pytest.raises(engine.IVPError) as e:
        s.solve_ivp_ensemble(y0, k)
    assert e.value.variant == "Unsupported"


def test_bdf_on_nonstiff_systems(cuda, engine, oracle):
    """BDF6/BDF2 on Lorenz, Van der Pol and a 4x4 linear system (D = 2, 3, 4 Jacobians)."""
    rng = np.random.default_rng(11)
    n = 256
    cases = [("lorenz", E.lorenz_y0(np.arange(n)), np.tile(np.array(E.LORENZ["params"])[:, None], (1, n)), 0.05),
             ("vdp", np.stack([np.full(n, 2.0), np.zeros(n)]), rng.uniform(0.1, 5.0, (1, n)), 0.1),
             ("linear4", rng.normal(size=(4, n)), rng.normal(size=(16, n)) * 0.5, 0.2)]
    for rhs, y0, p, t_end in cases:
        cfg = dict(dt_min=1e-9, dt_max=1e-3, tol=1e-7, t_start=0.0, t_end=t_end)
        gpu, ref = run_both(engine, oracle, "BDF6", rhs, y0, p, strict=True, **cfg)
        _bit_exact(gpu, ref)
        s = make_solver(engine, "BDF6", y0.shape[0], rhs=rhs, flags=_abi.FLAG_BDF_NEWTON, **cfg)
        nw = s.solve_ivp_ensemble(y0, p)
        assert (nw.status == _abi.OK).all()
        assert rel_err(nw.y_end, ref["y_end"]).max() <= band(1e-7)
