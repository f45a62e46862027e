"""GPU parity of the path queries (sampling the continuous extension, locating events) against the
CPU oracle, through the C ABI.  Both sides are given the SAME stored paths (the GPU solve's own history), so what is
compared is the query kernels (bacon_b200/csrc/path_query.cuh) alone:

strict build : bit-exact with the oracle (same operation order, no FMA contraction)
fast build   : within 1e-12 relative of the oracle; event counts identical
and, independently of the oracle, against closed forms.
"""
import numpy as np
import pytest

from bacon_b200 import _abi, ensembles as E
from parity import make_solver

pytestmark = pytest.mark.gpu

LOR_P = np.array(E.LORENZ["params"])


def _solved(res):
    return dict(hist=res.hist, hist_len=res.hist_len, t_end=res.t_end, y_end=res.y_end)


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint64)


def _lorenz(engine, n, *, strict, t_end=1.0, history=1200, method="RK45", tol=1e-8):
    y0 = E.lorenz_y0(np.arange(n))
    s = make_solver(engine, method, 3, rhs="lorenz", dt_min=1e-9, dt_max=0.1, tol=tol, t_start=0.0, t_end=t_end,
                    flags=_abi.FLAG_STRICT_FP if strict else 0, history=history)
    res = s.solve_ivp_ensemble(y0, LOR_P, shared_params=True)
    assert (res.status == _abi.OK).all()
    return y0, res


@pytest.mark.parametrize("strict", [True, False])
def test_sample_lorenz_matches_oracle(cuda, engine, oracle, strict):
    n = 777
    y0, res = _lorenz(engine, n, strict=strict)
    rng = np.random.default_rng(11)
    times = np.concatenate([[0.0, 1.0, -0.5, 1.5, np.nan], rng.uniform(0.0, 1.0, 120), res.hist_t[5, :7]])
    got = res.sample(times)
    ref = oracle.sample_paths("lorenz", y0, LOR_P, _solved(res), times, t_start=0.0, shared_params=True)
    assert got.shape == (n, times.size, 3)
    assert np.isnan(got[:, 2:5]).all() and np.isfinite(got[:, :2]).all() and np.isfinite(got[:, 5:]).all()
    if strict:
        assert np.array_equal(_bits(got), _bits(ref))
    else:
        np.testing.assert_allclose(got, ref, rtol=1e-12, atol=1e-12, equal_nan=True)
    # exact at the knots: t_start gives y0, t_end gives the final state, a record's time gives the record
    assert np.array_equal(got[:, 0], y0.T)
    assert np.array_equal(got[:, 1], res.y_end.T)
    assert np.array_equal(got[5, 125:132], res.hist_y[5, :7])


@pytest.mark.parametrize("strict", [True, False])
def test_knot_search_on_geometric_paths(cuda, engine, oracle, strict):
    """The sampling kernels look a sample's interval up by bisection down to 64 knots and then by interpolation on the
    bracket's end times (PathView::first_knot_at_or_after), which assumes a step size that varies slowly.  Here it does
    not: y' = -y from a first step of 1e-9 with dt allowed up to 50 — the steps grow fourfold per attempt, level off
    where the tolerance bites and grow again as y decays — so the interpolated guesses stagnate at one end of the
    bracket and the search must fall back to bisection after PATH_INTERP_MAX tries.  Whatever the probe sequence, the
    knot found is the oracle's (a plain bisection): strict samples bit for bit."""
    n = 96
    rng = np.random.default_rng(23)
    y0 = rng.uniform(0.5, 1.5, size=(1, n)) * 10.0 ** rng.integers(-3, 4, size=(1, n))
    s = make_solver(engine, "RK45", 1, rhs="decay", dt_min=1e-12, dt_max=50.0, tol=1e-6, t_start=0.0, t_end=200.0,
                    flags=_abi.FLAG_STRICT_FP if strict else 0, history=2048).with_initial_dt(1e-9)
    res = s.solve_ivp_ensemble(y0, None)
    assert (res.status == _abi.OK).all()
    steps = np.diff(res.hist_t[0, :res.hist_len[0]])
    assert res.hist_len.min() > 40 and steps.max() / steps.min() > 1e6  # the path really is geometric
    times = np.concatenate([[0.0, 200.0], np.minimum(np.logspace(-10, np.log10(200.0), 150), 200.0), rng.uniform(0.0, 200.0, 60),
                            res.hist_t[3, :20], np.nextafter(res.hist_t[7, :20], np.inf)])
    got = res.sample(times)
    ref = oracle.sample_paths("decay", y0, None, _solved(res), times, t_start=0.0)
    assert np.isfinite(got).all()
    if strict:
        assert np.array_equal(_bits(got), _bits(ref))
    else:
        np.testing.assert_allclose(got, ref, rtol=1e-12, atol=1e-300)
    exact = y0.T[:, None, :] * np.exp(-times)[None, :, None]
    assert np.abs(got - exact).max() <= 1e-4  # (the controller's tolerance is absolute: 1e-6 per unit of time, steps up to 5.8)


@pytest.mark.parametrize("strict", [True, False])
@pytest.mark.parametrize("direction", [0, 1, -1])
def test_events_lorenz_match_oracle(cuda, engine, oracle, strict, direction):
    """x = 0 crossings of Lorenz trajectories (lobe switches) and the Poincare section z = rho - 1."""
    n = 513
    y0, res = _lorenz(engine, n, strict=strict, t_end=3.0, history=3200)
    for w, c, cap in (([1.0, 0.0, 0.0], 0.0, 6), ([0.0, 0.0, 1.0], 27.0, 3)):
        ev, cnt = res.locate_events(w, c, direction, cap)
        rev, rcnt = oracle.locate_events("lorenz", y0, LOR_P, _solved(res), w, c, direction, cap, t_start=0.0,
                                         shared_params=True)
        np.testing.assert_array_equal(cnt, rcnt)
        assert cnt.sum() > n // 2
        if c != 0.0:
            assert cnt.max() > cap  # more events than capacity on some trajectory: counted, not stored
        if strict:
            assert np.array_equal(_bits(ev), _bits(rev))
        else:
            np.testing.assert_allclose(ev, rev, rtol=1e-12, atol=1e-12)
        for i in (0, 100, n - 1):  # events sit on the surface, in time order, and slots past the count stay zero
            k = min(int(cnt[i]), cap)
            g = ev[i, :k, 1:] @ np.array(w) - c
            assert np.abs(g).max(initial=0.0) < 1e-9
            assert (np.diff(ev[i, :k, 0]) > 0).all()
            assert (ev[i, k:] == 0).all()
    if direction == 0:  # rising + falling = both
        w = [1.0, 0.0, 0.0]
        up = res.locate_events(w, 0.0, 1, 1)[1]
        down = res.locate_events(w, 0.0, -1, 1)[1]
        np.testing.assert_array_equal(up + down, res.locate_events(w, 0.0, 0, 1)[1])


def test_harmonic_closed_form(cuda, engine):
    """y'' = -w^2 y from (1, 0): y = cos wt.  Samples within the interpolation error of the exact solution; zeros of
    y at (2k + 1) pi / (2w), of y' at k pi / w."""
    wv = np.array([1.0, 2.0, 3.5])
    n = wv.size
    y0 = np.stack([np.ones(n), np.zeros(n)])
    s = make_solver(engine, "RK45", 2, rhs="harmonic", dt_min=1e-9, dt_max=0.1, tol=1e-10, t_start=0.0, t_end=5.0,
                    history=4096)
    res = s.solve_ivp_ensemble(y0, wv.reshape(1, n))
    assert (res.status == _abi.OK).all()
    times = np.linspace(0.0, 5.0, 1001)
    got = res.sample(times)
    for i, w in enumerate(wv):
        m = int(res.hist_len[i])
        knot_err = np.abs(res.hist_y[i, :m, 0] - np.cos(w * res.hist_t[i, :m])).max()
        assert np.abs(got[i, :, 0] - np.cos(w * times)).max() < 3 * knot_err + 1e-12
        assert np.abs(got[i, :, 1] + w * np.sin(w * times)).max() < 3 * w * knot_err + 1e-12
    ev, cnt = res.locate_events([1.0, 0.0], 0.0, 0, 8)
    for i, w in enumerate(wv):
        exact = (2 * np.arange(64) + 1) * np.pi / (2 * w)
        exact = exact[exact < 5.0]
        assert cnt[i] == exact.size
        assert np.abs(ev[i, :exact.size, 0] - exact[:8]).max() < 1e-8
    ev, cnt = res.locate_events([0.0, 1.0], 0.0, 1, 8)  # y' rising through zero: minima of y, t = (2k + 1) pi / w
    for i, w in enumerate(wv):
        exact = (2 * np.arange(64) + 1) * np.pi / w
        exact = exact[exact < 5.0]
        assert cnt[i] == exact.size
        assert np.abs(ev[i, :exact.size, 0] - exact).max(initial=0.0) < 1e-8


def test_against_scipy_dense_output_and_events(cuda, engine):
    """Independent of the oracle: SciPy DOP853's own dense output and event root-finder on seeded Lorenz trajectories
    (tests/golden/anchors.json, generated by tests/golden/make_golden.py).  The CPU oracle meets these anchors to 3e-12 /
    9e-12 at this tolerance (tests/test_oracle.py); the fast kernels' solve differs from it inside the parity band."""
    import json
    import os
    anchors = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "anchors.json")))["lorenz_seeded4_T2_paths"]
    times = np.array(anchors["times"])
    y0 = E.lorenz_y0(np.arange(4))
    for strict in (False, True):
        s = make_solver(engine, "RK45", 3, rhs="lorenz", dt_min=1e-9, dt_max=0.1, tol=1e-10, t_start=0.0, t_end=2.0,
                        flags=_abi.FLAG_STRICT_FP if strict else 0, history=8192)
        res = s.solve_ivp_ensemble(y0, LOR_P, shared_params=True)
        assert (res.status == _abi.OK).all()
        got = res.sample(times)
        ref = np.array(anchors["states"])
        assert np.abs(got - ref).max() < 1e-8 * np.abs(ref).max()
        ev, cnt = res.locate_events([0.0, 0.0, 1.0], 27.0, 0, 16)
        for i in range(4):
            want = np.array(anchors["z27_events"][i])
            assert cnt[i] == want.size
            assert np.abs(ev[i, :want.size, 0] - want).max() < 1e-8
            assert np.abs(ev[i, :want.size, 3] - 27.0).max() < 1e-9


def test_many_events_per_trajectory(cuda, engine, oracle):
    """More crossings than the warp's queue holds (32) between two flushes, and capacity below / above the count."""
    wv = np.array([40.0, 55.0, 3.0, 20.0, 0.5])
    n = wv.size
    y0 = np.stack([np.ones(n), np.zeros(n)])
    p = wv.reshape(1, n)
    s = make_solver(engine, "RK45", 2, rhs="harmonic", dt_min=1e-9, dt_max=0.1, tol=1e-5, t_start=0.0, t_end=5.0,
                    history=8192)
    res = s.solve_ivp_ensemble(y0, p)
    assert (res.status == _abi.OK).all()
    for cap in (256, 40):
        ev, cnt = res.locate_events([1.0, 0.0], 0.0, 0, cap)
        rev, rcnt = oracle.locate_events("harmonic", y0, p, _solved(res), [1.0, 0.0], 0.0, 0, cap, t_start=0.0)
        np.testing.assert_array_equal(cnt, rcnt)
        np.testing.assert_allclose(ev, rev, rtol=1e-12, atol=1e-13)
        for i, w in enumerate(wv):
            exact = (2 * np.arange(1024) + 1) * np.pi / (2 * w)
            exact = exact[exact < 5.0]
            assert cnt[i] == exact.size
            k = min(exact.size, cap)
            assert np.abs(ev[i, :k, 0] - exact[:k]).max() < 1e-4
            assert (ev[i, k:] == 0).all()
    assert cnt.max() > 80


@pytest.mark.parametrize("method,rhs,cfg", [
    ("BDF6", "robertson", dict(dt_min=1e-10, dt_max=1e-4, tol=1e-6, t_end=0.02)),
    ("Adams5", "exp", dict(dt_min=1e-5, dt_max=0.1, tol=1e-5, t_end=2.0)),
    ("Euler", "decay", dict(dt_min=0.01, dt_max=0.01, tol=1e-5, t_end=1.0)),
    ("RK23", "vdp", dict(dt_min=1e-9, dt_max=0.1, tol=1e-5, t_end=2.0)),
])
def test_every_method_family(cuda, engine, oracle, method, rhs, cfg):
    """The queries serve every stepper: BDF (whose last block of points may go unyielded: the exit state closes the
    path), Adams, Euler (records are the OLD points: knot 1 repeats the initial condition), the second RK pair."""
    n = 64
    rng = np.random.default_rng(5)
    dim = {"robertson": 3, "exp": 1, "decay": 1, "vdp": 2}[rhs]
    if rhs == "robertson":
        y0 = np.tile(np.array([[1.0], [0.0], [0.0]]), (1, n))
        params = np.array([[0.04], [3e7], [1e4]]) * (1.0 + 0.1 * rng.uniform(-1, 1, size=(3, n)))
    elif rhs == "vdp":
        y0 = np.tile(np.array([[2.0], [0.0]]), (1, n))
        params = rng.uniform(0.1, 3.0, size=(1, n))
    else:
        y0 = rng.uniform(0.5, 1.5, size=(dim, n))
        params = None
    s = make_solver(engine, method, dim, rhs=rhs, t_start=0.0, history=4096, **cfg)
    res = s.solve_ivp_ensemble(y0, params)
    assert (res.status == _abi.OK).all()
    t_end = float(res.t_end.min())  # (Euler and Adams overshoot the ending time by an ulp or a step; the path goes that far)
    assert t_end >= cfg["t_end"]
    times = np.concatenate([[0.0, t_end], rng.uniform(0.0, t_end, 40)])
    got = res.sample(times)
    ref = oracle.sample_paths(rhs, y0, params, _solved(res), times, t_start=0.0)
    assert np.isfinite(got).all()
    np.testing.assert_allclose(got, ref, rtol=1e-12, atol=1e-14)
    assert np.array_equal(got[:, 0], y0.T)
    # the path reaches the stepper's exit time for every method (closing knot), and ends in the final state
    at_end = res.t_end == t_end
    assert at_end.any() and np.array_equal(got[at_end, 1], res.y_end.T[at_end])
    w = np.zeros(dim)
    w[0] = 1.0
    c = float(np.median(got[:, 2:, 0]))
    ev, cnt = res.locate_events(w, c, 0, 4)
    rev, rcnt = oracle.locate_events(rhs, y0, params, _solved(res), w, c, 0, 4, t_start=0.0)
    np.testing.assert_array_equal(cnt, rcnt)
    np.testing.assert_allclose(ev, rev, rtol=1e-12, atol=1e-14)
    assert cnt.sum() > 0


@pytest.mark.parametrize("strict", [True, False])
@pytest.mark.parametrize("aos", [True, False])
def test_linear32_warp_queries(cuda, engine, oracle, strict, aos):
    """BASELINE config 4's right-hand side (D = 32, per-trajectory matrix): the warp-cooperative query kernels
    (path_query_warp.cuh) against the oracle on the same paths — strict bit-exact, fast within 1e-12 — and against
    the matrix exponential."""
    from scipy.linalg import expm
    n = 37
    y0, A = E.linear32_problem(np.arange(n))          # (32, n), (n, 32, 32)
    par = A.reshape(n, 1024) if aos else np.ascontiguousarray(A.reshape(n, 1024).T)
    s = make_solver(engine, "RK45", 32, rhs="linear32", dt_min=1e-9, dt_max=0.1, tol=1e-8, t_start=0.0, t_end=2.0,
                    flags=_abi.FLAG_STRICT_FP if strict else 0, history=256)
    res = s.solve_ivp_ensemble(y0, par, params_aos=aos)
    assert (res.status == _abi.OK).all()
    rng = np.random.default_rng(3)
    times = np.concatenate([[0.0, 2.0, 2.5], rng.uniform(0.0, 2.0, 29)])
    got = res.sample(times)
    ref = oracle.sample_paths("linear32", y0, par, _solved(res), times, t_start=0.0, params_aos=aos)
    assert np.isnan(got[:, 2]).all() and np.array_equal(got[:, 0], y0.T) and np.array_equal(got[:, 1], res.y_end.T)
    if strict:
        assert np.array_equal(_bits(got), _bits(ref))
    else:
        np.testing.assert_allclose(got, ref, rtol=1e-12, atol=1e-13, equal_nan=True)
    for i in (0, n - 1):
        for j in (3, 17):
            exact = expm(A[i] * times[j]) @ y0[:, i]
            assert np.abs(got[i, j] - exact).max() < 1e-6
    w = rng.normal(size=32)
    for direction, cap in ((0, 8), (1, 2)):
        ev, cnt = res.locate_events(w, 0.05, direction, cap)
        rev, rcnt = oracle.locate_events("linear32", y0, par, _solved(res), w, 0.05, direction, cap, t_start=0.0,
                                         params_aos=aos)
        np.testing.assert_array_equal(cnt, rcnt)
        assert cnt.sum() > 0
        if strict:
            assert np.array_equal(_bits(ev), _bits(rev))
        else:
            np.testing.assert_allclose(ev, rev, rtol=1e-11, atol=1e-12)
        for i in range(n):
            k = min(int(cnt[i]), cap)
            assert np.abs(ev[i, :k, 1:] @ w - 0.05).max(initial=0.0) < 1e-9


def test_failed_trajectory_path_ends_early(cuda, engine, oracle):
    """A trajectory that fails (here: the attempt cap) has a path up to its failure time: NaN beyond."""
    y0 = np.array([[1.0, 1.0]])
    s = make_solver(engine, "RK45", 1, rhs="exp", dt_min=1e-3, dt_max=0.1, tol=1e-6, t_start=0.0, t_end=10.0, history=64,
                    max_attempts=10)
    res = s.solve_ivp_ensemble(y0)
    assert (res.status == _abi.E_MAX_ATTEMPTS).all() and (res.n_accept == 10).all()
    t_fail = float(res.t_end[0])
    assert 0.5 < t_fail < 1.5
    times = np.array([0.0, 0.5 * t_fail, t_fail, t_fail + 1e-3, 10.0])
    got = res.sample(times)
    ref = oracle.sample_paths("exp", y0, None, _solved(res), times, t_start=0.0)
    np.testing.assert_allclose(got, ref, rtol=1e-12, equal_nan=True)
    assert np.isnan(got[:, 3:]).all() and np.isfinite(got[:, :3]).all()


def test_device_entry_points_match_host(cuda, engine):
    torch = cuda
    n = 300
    y0 = E.lorenz_y0(np.arange(n))
    s = make_solver(engine, "RK45", 3, rhs="lorenz", dt_min=1e-9, dt_max=0.1, tol=1e-8, t_start=0.0, t_end=1.0,
                    history=1200)
    host = s.solve_ivp_ensemble(y0, LOR_P, shared_params=True)
    d_y0 = torch.from_numpy(y0).cuda()
    d_p = torch.from_numpy(LOR_P).cuda()
    out = s.solve_ivp_ensemble_device(d_y0, d_p, shared_params=True)
    times = np.linspace(0.0, 1.0, 33)
    d_s = s.sample_paths_device(d_y0, d_p, out, torch.from_numpy(times).cuda(), shared_params=True)
    d_ev, d_cnt = s.locate_events_device(d_y0, d_p, out, [1.0, 0.0, 0.0], 0.0, 0, 4, shared_params=True)
    torch.cuda.synchronize()
    assert np.array_equal(_bits(d_s.cpu().numpy()), _bits(host.sample(times)))
    ev, cnt = host.locate_events([1.0, 0.0, 0.0], 0.0, 0, 4)
    assert np.array_equal(d_cnt.cpu().numpy().astype(np.uint32), cnt)
    assert np.array_equal(_bits(d_ev.cpu().numpy()), _bits(ev))
    info = engine.ivp.last_launch()
    assert info["n_kernels"] == 1 and info["block"] == 128


def test_params_layouts(cuda, engine, oracle):
    """Per-trajectory parameters SoA and AoS give the same samples as the oracle."""
    n = 130
    rng = np.random.default_rng(2)
    y0 = E.lorenz_y0(np.arange(n))
    p = LOR_P.reshape(3, 1) * (1.0 + 0.05 * rng.uniform(-1, 1, size=(3, n)))
    times = np.linspace(0.0, 0.5, 21)
    outs = []
    for aos in (False, True):
        s = make_solver(engine, "RK45", 3, rhs="lorenz", dt_min=1e-9, dt_max=0.1, tol=1e-8, t_start=0.0, t_end=0.5,
                        history=700)
        pp = np.ascontiguousarray(p.T) if aos else p
        res = s.solve_ivp_ensemble(y0, pp, params_aos=aos)
        got = res.sample(times)
        ref = oracle.sample_paths("lorenz", y0, pp, _solved(res), times, t_start=0.0, params_aos=aos)
        np.testing.assert_allclose(got, ref, rtol=1e-12, atol=1e-12)
        outs.append(got)
    assert np.array_equal(_bits(outs[0]), _bits(outs[1]))


def test_unsupported_and_bad_arguments(cuda, engine):
    y0 = E.lorenz_y0(np.arange(4))
    s = make_solver(engine, "RK45", 3, rhs="lorenz", dt_min=1e-9, dt_max=0.1, tol=1e-8, t_start=0.0, t_end=0.1)
    res = s.solve_ivp_ensemble(y0, LOR_P, shared_params=True)  # no history
    with pytest.raises(engine.IVPError) as e:
        res.sample([0.05])
    assert e.value.code == _abi.E_BAD_ARGUMENT
    s = make_solver(engine, "RK45", 3, rhs="lorenz", dt_min=1e-9, dt_max=0.1, tol=1e-8, t_start=0.0, t_end=0.1, history=64)
    res = s.solve_ivp_ensemble(y0, LOR_P, shared_params=True)
    with pytest.raises(engine.IVPError) as e:
        res.locate_events([1.0, 0.0, 0.0], 0.0, 2, 4)  # direction out of range
    assert e.value.code == _abi.E_BAD_ARGUMENT
    assert res.sample([]).shape == (4, 0, 3)
    ev, cnt = res.locate_events([1.0, 0.0, 0.0], 1e9, 0, 0)  # capacity 0: counts only
    assert ev.shape == (4, 0, 4) and (cnt == 0).all()
