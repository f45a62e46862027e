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
    # always through make: the plug-in is compiled against the library's headers and must follow them
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


def test_user_rhs_path_queries(cuda, engine, user_lib):
    """A functor registered with BACON_REGISTER_RHS gets the path-query kernels like a built-in: the pendulum's
    sampled energy stays constant to the solver's accuracy, and theta = 0 crossings alternate in direction."""
    n = 32
    rng = np.random.default_rng(9)
    th = np.stack([rng.uniform(0.3, 1.0, n), np.zeros(n)])
    gl = rng.uniform(5, 15, (1, n))
    s = make_solver(engine, "RK45", 2, rhs="pendulum", dt_min=1e-10, dt_max=0.05, tol=1e-10, t_start=0.0, t_end=3.0,
                    history=2048)
    r = s.solve_ivp_ensemble(th, gl)
    assert (r.status == _abi.OK).all()
    times = np.linspace(0.0, 3.0, 301)
    y = r.sample(times)
    energy = 0.5 * y[:, :, 1] ** 2 - gl.T * np.cos(y[:, :, 0])
    assert np.abs(energy - energy[:, :1]).max() < 1e-7
    ev, cnt = r.locate_events([1.0, 0.0], 0.0, 0, 16)
    up = r.locate_events([1.0, 0.0], 0.0, 1, 16)[1]
    down = r.locate_events([1.0, 0.0], 0.0, -1, 16)[1]
    assert (cnt >= 2).all() and (up + down == cnt).all() and (np.abs(down.astype(int) - up.astype(int)) <= 1).all()
    for i in range(n):  # half-periods are equal
        t = ev[i, :min(int(cnt[i]), 16), 0]
        assert np.ptp(np.diff(t)) < 1e-6


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


LORENZ_SRC = """
struct LorenzRtc {  // p = (sigma, rho, beta): the same expression trees as the built-in
    static constexpr int DIM = 3, NPARAM = 3;
    __device__ void operator()(double, const double (&y)[3], const double* p, double (&dy)[3]) const {
        dy[0] = p[0] * (y[1] - y[0]);
        dy[1] = y[0] * (p[1] - y[2]) - y[1];
        dy[2] = y[0] * y[1] - p[2] * y[2];
    }
    __device__ void scaled(double h, double, const double (&y)[3], const double* p, double (&k)[3]) const {
        const double hs = h * p[0];
        k[0] = hs * (y[1] - y[0]);
        k[1] = h * (y[0] * (p[1] - y[2]) - y[1]);
        k[2] = h * (y[0] * y[1] - p[2] * y[2]);
    }
    __device__ void jac(double, const double (&y)[3], const double* p, double (&J)[3][3]) const {
        J[0][0] = -p[0];       J[0][1] = p[0];  J[0][2] = 0.0;
        J[1][0] = p[1] - y[2]; J[1][1] = -1.0;  J[1][2] = -y[0];
        J[2][0] = y[1];        J[2][1] = y[0];  J[2][2] = -p[2];
    }
};
"""


def test_runtime_compiled_rhs_equals_the_built_in(cuda, engine, oracle, monkeypatch):
    """bacon_rhs_register_source: the Lorenz functor handed over as source text and compiled by NVRTC with the library's
    own kernel headers must behave like the built-in compiled by nvcc — bit for bit with the oracle in the strict
    kernels (history included), inside the parity band of the built-in in the fast ones, through every stepper family, with regrouping, and a source error must surface as UserError."""
    from bacon_b200 import ensembles as E
    from parity import run_both
    rid = engine.register_rhs_source("lorenz_rtc", "LorenzRtc", LORENZ_SRC, 3, 3)
    assert rid >= 0
    P = np.array(E.LORENZ["params"])
    # path queries: the runtime-compiled program (path_query.cuh through NVRTC) against the oracle on the same paths —
    # strict bit for bit, fast within 1e-12
    for strict in (True, False):
        q = make_solver(engine, "RK45", 3, rhs="lorenz_rtc", history=1400, t_end=0.5, dt_min=1e-9, dt_max=0.1, tol=1e-8,
                        t_start=0.0, flags=_abi.FLAG_STRICT_FP if strict else 0)
        yq = E.lorenz_y0(np.arange(300))
        r = q.solve_ivp_ensemble(yq, P, shared_params=True)
        assert (r.status == _abi.OK).all()
        sv = dict(hist=r.hist, hist_len=r.hist_len, t_end=r.t_end, y_end=r.y_end)
        times = np.linspace(-0.05, 0.55, 41)
        got = r.sample(times)
        ref = oracle.sample_paths("lorenz", yq, P, sv, times, t_start=0.0, shared_params=True)
        ev, cnt = r.locate_events([0.0, 0.0, 1.0], 27.0, 0, 3)
        rev, rcnt = oracle.locate_events("lorenz", yq, P, sv, [0.0, 0.0, 1.0], 27.0, 0, 3, t_start=0.0, shared_params=True)
        assert (cnt == rcnt).all() and cnt.sum() > 0
        if strict:
            assert np.array_equal(got.view(np.uint64), ref.view(np.uint64)) and np.array_equal(ev.view(np.uint64), rev.view(np.uint64))
        else:
            np.testing.assert_allclose(got, ref, rtol=1e-12, atol=1e-12, equal_nan=True)
            np.testing.assert_allclose(ev, rev, rtol=1e-12, atol=1e-12)
    LOR = dict(dt_min=1e-9, dt_max=0.1, tol=1e-8, t_start=0.0)
    n = 4000
    y0 = E.lorenz_y0(np.arange(n))
    # strict: oracle parity (the oracle knows the problem as "lorenz")
    for method in ("RK45", "RK23"):
        s = make_solver(engine, method, 3, rhs="lorenz_rtc", flags=_abi.FLAG_STRICT_FP, history=96, t_end=0.1, **LOR)
        g = s.solve_ivp_ensemble(y0, P, shared_params=True)
        r = oracle.solve_ensemble(getattr(_abi, method), "lorenz", y0, P, shared_params=True, history_capacity=96, pow_mode=1,
                                  t_end=0.1, **LOR)
        assert np.array_equal(g.y_end.view(np.uint64), r["y_end"].view(np.uint64)), method
        np.testing.assert_array_equal(g.n_accept, r["n_accept"])
        mask = np.arange(96)[None, :] < g.hist_len[:, None]
        assert np.array_equal(g.hist_y[mask], r["hist_y"][mask])
    # fast, on a tiny grid so that lanes refill and the CTAs regroup: the same bits as the built-in
    monkeypatch.setenv("BACON_IVP_GRID", "2")
    for method, extra in (("RK45", {}), ("RK23", {}), ("BDF6", dict(flags=_abi.FLAG_BDF_NEWTON)), ("BDF2", {}), ("Adams5", {}),
                          ("Euler", {})):
        cfg = dict(LOR, t_end=0.05)
        if method.startswith("BDF") or method == "Euler":
            cfg.update(dt_max=1e-3, tol=1e-6)
        a = make_solver(engine, method, 3, rhs="lorenz_rtc", **extra, **cfg).solve_ivp_ensemble(y0, P, shared_params=True)
        launch = engine.last_launch()
        b = make_solver(engine, method, 3, rhs="lorenz", **extra, **cfg).solve_ivp_ensemble(y0, P, shared_params=True)
        assert launch["n_kernels"] == engine.last_launch()["n_kernels"] == 1, method
        assert launch["block"] == engine.last_launch()["block"] == (768 if method.startswith("RK") else 128), method
        np.testing.assert_array_equal(a.status, b.status, err_msg=method)
        # (the process may hold an NVRTC of another minor version than the nvcc that built the library: FMA contraction
        # can then differ in the last bit, so the fast kernels are compared inside the parity band, counts side by side)
        assert np.abs(a.n_accept.astype(np.int64) - b.n_accept.astype(np.int64)).max() <= 2, method
        num = np.sqrt(((a.y_end - b.y_end) ** 2).sum(0))
        assert (num / np.sqrt((b.y_end ** 2).sum(0))).max() <= 10 * cfg["tol"], method
    # a functor that does not compile
    with pytest.raises(engine.IVPError) as e:
        engine.register_rhs_source("broken_rtc", "LorenzRtc", LORENZ_SRC.replace("p[2] * y[2];\n    }\n    __device__ void scaled", "p[2] * z;\n    }\n    __device__ void scaled"), 3, 3)
    assert e.value.variant == "UserError" and "z" in str(e.value)
