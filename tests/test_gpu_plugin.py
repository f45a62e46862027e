"""User RHS plug-in and the C++ façade on the GPU."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from bacon_b200 import _abi
from parity import make_solver

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def user_lib(engine):
    so = os.path.join(ROOT, "examples", "libuser_rhs.so")
    if not os.path.exists(so):
        subprocess.run(["make", "-C", os.path.join(ROOT, "examples")], check=True)
    from bacon_b200._lib import lib
    lib()
    return C.CDLL(so, mode=C.RTLD_GLOBAL)


def test_user_rhs_against_scipy(cuda, engine, user_lib):
    from scipy.integrate import solve_ivp
    n = 64
    rng = np.random.default_rng(5)
    y0 = rng.uniform(0.5, 2.0, (2, n))
    ab = np.stack([rng.uniform(0.8, 1.2, n), rng.uniform(1.5, 3.0, n)])
    cfg = dict(dt_min=1e-10, dt_max=0.05, tol=1e-9, t_start=0.0, t_end=3.0)
    res = {}
    for method, flags in (("RK45", 0), ("RK23", 0), ("BDF6", 0), ("BDF6", _abi.FLAG_BDF_NEWTON)):
        if method == "BDF6":
            c = dict(cfg, dt_max=2e-3, tol=1e-8, t_end=1.0)
        else:
            c = cfg
        r = make_solver(engine, method, 2, rhs="brusselator", flags=flags, **c).solve_ivp_ensemble(y0, ab)
        assert (r.status == _abi.OK).all(), (method, flags, np.unique(r.status))
        res[(method, flags)] = (r, c)
    def f(t, y, a, b):
        return [a + y[0] ** 2 * y[1] - (b + 1) * y[0], b * y[0] - y[0] ** 2 * y[1]]
    for (method, flags), (r, c) in res.items():
        for i in range(0, n, 8):
            s = solve_ivp(f, (0, c["t_end"]), y0[:, i], method="DOP853", rtol=1e-12, atol=1e-13, args=tuple(ab[:, i]))
            np.testing.assert_allclose(r.y_end[:, i], s.y[:, -1], rtol=2e-6 if method != "RK45" else 1e-7, atol=1e-9)
    # functor without jac(): Newton falls back to finite differences
    th = np.stack([rng.uniform(-1, 1, n), np.zeros(n)])
    gl = rng.uniform(5, 15, (1, n))
    c = dict(dt_min=1e-10, dt_max=2e-3, tol=1e-8, t_start=0.0, t_end=0.5)
    a = make_solver(engine, "BDF6", 2, rhs="pendulum", flags=_abi.FLAG_BDF_NEWTON, **c).solve_ivp_ensemble(th, gl)
    b = make_solver(engine, "RK45", 2, rhs="pendulum", **dict(c, dt_max=0.05, tol=1e-10)).solve_ivp_ensemble(th, gl)
    assert (a.status == _abi.OK).all() and (b.status == _abi.OK).all()
    np.testing.assert_allclose(a.y_end, b.y_end, rtol=1e-5, atol=1e-7)
    # a user RHS has no strict build unless it ships one: loud error, no fallback
    with pytest.raises(engine.IVPError) as e:
        make_solver(engine, "RK45", 2, rhs="pendulum", flags=_abi.FLAG_STRICT_FP, **c).solve_ivp_ensemble(th, gl)
    assert e.value.variant == "Unsupported"


def test_cpp_facade_readme_example(cuda):
    exe = "/tmp/bacon_cpp_facade_test"
    subprocess.run(["g++", "-std=c++17", "-O1", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp_facade_test.cpp"),
                    "-o", exe, "-L" + os.path.join(ROOT, "bacon_b200"), "-lbacon_ivp", "-Wl,-rpath," + os.path.join(ROOT, "bacon_b200")],
                   check=True)
    out = subprocess.run([exe, "solve"], check=True, capture_output=True, text=True).stdout
    assert "solve ok: 128 points" in out and out.strip().endswith("ok")


def test_single_trajectory_solve_is_the_reference_call(cuda, engine):
    """README.md:32-40 through the Python mirror: builder -> solve -> Path."""
    s = (engine.RK45.new(1).with_dt_min(0.01).with_dt_max(0.1).with_tolerance(1e-4).with_initial_conditions([1.0])
         .with_start(0.0).with_end(10.0).build())
    path = s.solve_ivp("exp")
    assert len(path) == 128 and path[-1][0] == 10.0
    assert abs(path[-1][1][0] / np.exp(10.0) - 1.0) < 1e-6
    # failure: the error is raised after the points yielded before it (ivp.rs:232-235)
    s = (engine.RK45.new(1).with_dt_min(0.01).with_dt_max(0.1).with_tolerance(1e-4).with_initial_conditions([1.0])
         .with_start(0.0).with_end(10.0).with_semantics(_abi.SEM_LITERAL))
    with pytest.raises(engine.IVPError) as e:
        s.solve_ivp("exp")
    assert e.value.variant == "MinimumTimeDeltaExceeded" and len(e.value.path) == 1
