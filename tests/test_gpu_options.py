"""GPU parity of what ABI v6 added, through the C ABI, against the CPU oracle's statement of the same rules:

* REF_LITERAL BDF on the device (bdf.rs:407 `above + below`, :568 lower formula with the higher coefficients, :622
  `time -= dt - order`, g at t_n) — the source as written;
* `with_initial_dt` and the restart record (t_end, y_end, dt_end) -> (t_start_each, y0, dt_start_each): the C-ABI form
  of the reference's in-memory resumable iterator (src/ivp.rs:220-238);
* terminal events (stop at the first zero of w . y - c): NOT in the reference; pinned on closed forms, on the oracle's
  statement (strict build bit-exact) and on the events query of the stored full path;
* a path cut short by its capacity ends at its last record (no closing knot across the unrecorded span);
* strict / LITERAL solves in flight on two streams keep their own Butcher tableau;
* `solve()` of one trajectory grows its path like collect_vec (ivp.rs:209-211).
"""
import numpy as np
import pytest

from bacon_b200 import _abi, ensembles as E
from parity import METHODS, band, make_solver, rel_err
from reference_cases import BDF_CASES, bdf_cfg

pytestmark = pytest.mark.gpu
LOR = dict(dt_min=1e-9, dt_max=0.1, tol=1e-8, t_start=0.0)
P_LOR = np.array(E.LORENZ["params"])


def _same_bits(a, b, what):
    a, b = np.ascontiguousarray(a, dtype=np.float64), np.ascontiguousarray(b, dtype=np.float64)
    assert np.array_equal(a.view(np.uint64), b.view(np.uint64)), f"{what}: max |d| = {np.nanmax(np.abs(a - b))}"


def _bit_exact(gpu, ref, keys=("y_end", "t_end", "dt_end")):
    for k in ("status", "n_accept", "n_reject", "n_rhs"):
        np.testing.assert_array_equal(getattr(gpu, k), ref[k], err_msg=k)
    for k in keys:
        _same_bits(getattr(gpu, k), ref[k], k)


# ---------------------------------------------------------------- REF_LITERAL BDF on the device
@pytest.mark.parametrize("case", BDF_CASES, ids=[c[0] for c in BDF_CASES])
def test_literal_bdf_is_the_source_as_written(cuda, engine, oracle, case):
    """bdf.rs:785-1063 in REF_LITERAL: the strict Broyden kernel gives the oracle's bits — status, counters, the
    stepper's internal final time (7.3, 14.45, ... : SURVEY.md §4) and the (empty) path."""
    name, method, rhs, y0, t_end, exact, eps, n_yield, _ = case
    y0 = np.array([[y0]])
    cfg = bdf_cfg(t_end)
    s = make_solver(engine, method, 1, rhs=rhs, semantics=_abi.SEM_LITERAL, history=64, max_attempts=200000, **cfg)
    gpu = s.solve_ivp_ensemble(y0)
    ref = oracle.solve_ensemble(METHODS[method], rhs, y0, semantics=_abi.SEM_LITERAL, history_capacity=64,
                                max_attempts=200000, pow_mode=1, **cfg)
    for k in ("status", "n_accept", "n_reject", "n_rhs", "hist_len"):
        np.testing.assert_array_equal(getattr(gpu, k), ref[k], err_msg=k)
    if rhs == "cos":  # device cos() and glibc's differ in the last ulp
        np.testing.assert_allclose(gpu.y_end, ref["y_end"], rtol=1e-12, atol=1e-14)
        np.testing.assert_allclose(gpu.t_end, ref["t_end"], rtol=1e-12)
    else:
        _bit_exact(gpu, ref)
    m = int(gpu.hist_len[0])
    _same_bits(gpu.hist_t[0, :m], ref["hist_t"][0, :m], "hist_t")


def test_literal_bdf_newton_or_fast_is_unsupported(cuda, engine):
    s = make_solver(engine, "BDF6", 1, rhs="exp", semantics=_abi.SEM_LITERAL, flags=_abi.FLAG_BDF_NEWTON, **bdf_cfg(1.0))
    with pytest.raises(engine.IVPError) as e:
        s.solve_ivp_ensemble(np.array([[1.0]]))
    assert e.value.variant == "Unsupported"


# ---------------------------------------------------------------- first dt and the restart record
@pytest.mark.parametrize("method,rhs", [("RK45", "lorenz"), ("RK23", "lorenz"), ("BDF6", "lorenz"), ("Adams5", "lorenz")])
def test_initial_dt_strict_bit_exact(cuda, engine, oracle, method, rhs):
    n = 96
    y0 = E.lorenz_y0(np.arange(n))
    cfg = dict(LOR, t_end=0.4, tol=1e-6, dt_min=1e-7)
    for dt_init in (3e-3, 5.0, 1e-12):  # inside the bounds, above dt_max and below dt_min (both clamped)
        s = make_solver(engine, method, 3, rhs=rhs, flags=_abi.FLAG_STRICT_FP, **cfg).with_initial_dt(dt_init)
        gpu = s.solve_ivp_ensemble(y0, P_LOR, shared_params=True)
        ref = oracle.solve_ensemble(METHODS[method], rhs, y0, P_LOR, shared_params=True, pow_mode=1, dt_init=dt_init, **cfg)
        _bit_exact(gpu, ref)
    # and it is not a no-op: the default first step (dt_max + dt_min)/2 gives another step sequence
    base = make_solver(engine, method, 3, rhs=rhs, flags=_abi.FLAG_STRICT_FP, **cfg).solve_ivp_ensemble(y0, P_LOR, shared_params=True)
    assert not np.array_equal(base.n_rhs, gpu.n_rhs) or not np.array_equal(base.y_end, gpu.y_end)


@pytest.mark.parametrize("method", ["RK45", "RK23", "BDF6", "Adams5", "Euler"])
def test_restart_record_resumes_where_the_solve_stopped(cuda, engine, oracle, method):
    """Leg 1 to T1, leg 2 from leg 1's (t_end, y_end, dt_end) to T2: the strict kernels give the oracle's bits for the
    same two legs, per-trajectory start times and first steps included (trajectories cut by the attempt cap in leg 1
    start leg 2 from wherever they were)."""
    n = 200
    y0 = E.lorenz_y0(np.arange(n))
    cfg = dict(LOR, tol=1e-6 if method != "RK23" else 1e-4, dt_min=1e-7)  # (RK23 at 1e-6 takes > 20000 steps per leg)
    if method == "Euler":
        cfg = dict(dt_min=2e-3, dt_max=2e-3, tol=1.0, t_start=0.0)
    cap1 = 40 if method != "Euler" else 0
    s1 = make_solver(engine, method, 3, rhs="lorenz", flags=_abi.FLAG_STRICT_FP, t_end=0.5, max_attempts=cap1, **cfg)
    g1 = s1.solve_ivp_ensemble(y0, P_LOR, shared_params=True)
    r1 = oracle.solve_ensemble(METHODS[method], "lorenz", y0, P_LOR, shared_params=True, pow_mode=1, t_end=0.5,
                               max_attempts=cap1, **cfg)
    _bit_exact(g1, r1)
    if cap1:
        assert (g1.status == _abi.E_MAX_ATTEMPTS).any() and len(np.unique(g1.t_end)) > 10
    y1, t1, dt1 = g1.restart_record()
    cap2 = 600 if method == "RK45" else 20000
    s2 = make_solver(engine, method, 3, rhs="lorenz", flags=_abi.FLAG_STRICT_FP, t_end=1.0, history=cap2, **cfg)
    g2 = s2.solve_ivp_ensemble(np.ascontiguousarray(y1), P_LOR, shared_params=True, restart=(t1, dt1))
    r2 = oracle.solve_ensemble(METHODS[method], "lorenz", y1, P_LOR, shared_params=True, pow_mode=1, t_end=1.0,
                               history_capacity=cap2, t_start_each=t1, dt_start_each=dt1, **cfg)
    assert (g2.status == _abi.OK).all(), np.unique(g2.status)
    _bit_exact(g2, r2)
    np.testing.assert_array_equal(g2.hist_len, r2["hist_len"])
    for i in (0, n // 2, n - 1):
        m = int(g2.hist_len[i])
        _same_bits(g2.hist[i, :m], r2["hist"][i, :m], f"path {i}")
        assert m == 0 or g2.hist_t[i, 0] >= t1[i]  # the leg starts at its own start time (Euler's first record IS the start)
    assert np.abs(g2.t_end - 1.0).max() < 1e-12
    # path queries on the resumed leg take the per-trajectory start times as knot 0
    times = np.array([0.75, 0.99, 1.0])
    smp = g2.sample(times)
    ref_s = oracle.sample_paths("lorenz", y1, P_LOR, dict(r2, t_start=t1), times, t_start=0.0, shared_params=True)
    _same_bits(smp, ref_s, "samples on the resumed leg")
    assert np.isfinite(smp).all()
    # the resumed solution is the same trajectory: it agrees with one uninterrupted solve at the tolerance's level
    if method in ("RK45", "RK23"):
        one = make_solver(engine, method, 3, rhs="lorenz", flags=_abi.FLAG_STRICT_FP, t_end=1.0, **cfg) \
            .solve_ivp_ensemble(y0, P_LOR, shared_params=True)
        assert rel_err(g2.y_end, one.y_end).max() < 1e-3


def test_restart_record_fast_kernels_and_device_entry(cuda, engine, oracle):
    torch = cuda
    n = 5000
    y0 = E.lorenz_y0(np.arange(n))
    s1 = make_solver(engine, "RK45", 3, rhs="lorenz", t_end=1.0, **LOR)
    g1 = s1.solve_ivp_ensemble(y0, P_LOR, shared_params=True)
    y1, t1, dt1 = g1.restart_record()
    s2 = make_solver(engine, "RK45", 3, rhs="lorenz", t_end=2.0, **LOR)
    g2 = s2.solve_ivp_ensemble(np.ascontiguousarray(y1), P_LOR, shared_params=True, restart=(t1, dt1))
    r2 = oracle.solve_ensemble(_abi.RK45, "lorenz", y1, P_LOR, shared_params=True, t_end=2.0, t_start_each=t1,
                               dt_start_each=dt1, **LOR)
    assert (g2.status == _abi.OK).all() and rel_err(g2.y_end, r2["y_end"]).max() <= band(1e-8)
    assert np.abs(g2.n_accept.astype(int) - r2["n_accept"].astype(int)).max() <= 2
    # device buffers: same bits as the host entry point
    dev = "cuda:0"
    d = s2.solve_ivp_ensemble_device(torch.from_numpy(np.ascontiguousarray(y1)).to(dev), torch.from_numpy(P_LOR).to(dev),
                                     shared_params=True, restart=(torch.from_numpy(t1.copy()).to(dev), torch.from_numpy(dt1.copy()).to(dev)))
    torch.cuda.synchronize()
    assert np.array_equal(d["y_end"].cpu().numpy(), g2.y_end) and np.array_equal(d["t_end"].cpu().numpy(), g2.t_end)
    # linear32 (warp-per-trajectory kernel) takes the record too
    m = 64
    z0, A = E.linear32_problem(np.arange(m))
    A = A.reshape(m, 1024)
    lin = dict(dt_min=1e-6, dt_max=0.05, tol=1e-8, t_start=0.0)
    a1 = make_solver(engine, "RK45", 32, rhs="linear32", flags=_abi.FLAG_STRICT_FP, t_end=0.3, **lin).solve_ivp_ensemble(z0, A, params_aos=True)
    a2 = make_solver(engine, "RK45", 32, rhs="linear32", flags=_abi.FLAG_STRICT_FP, t_end=0.6, **lin) \
        .solve_ivp_ensemble(np.ascontiguousarray(a1.y_end), A, params_aos=True, restart=(a1.t_end, a1.dt_end))
    b2 = oracle.solve_ensemble(_abi.RK45, "linear32", a1.y_end, A, params_aos=True, pow_mode=1, t_end=0.6,
                               t_start_each=a1.t_end, dt_start_each=a1.dt_end, **lin)
    _bit_exact(a2, b2)


# ---------------------------------------------------------------- terminal events
@pytest.mark.parametrize("method", ["RK45", "RK23", "BDF6", "BDF2", "Adams5", "Adams3", "Euler"])
def test_terminal_event_strict_bit_exact(cuda, engine, oracle, method):
    """Every stepper family: stop at the first crossing of z = 25 (either direction), strict build against the
    oracle's statement — status, event time and state, counters, and the path up to the event."""
    n = 256
    y0 = E.lorenz_y0(np.arange(n))
    cfg = dict(LOR, tol=1e-6, dt_min=1e-7, t_end=1.5)
    if method == "Euler":
        cfg = dict(dt_min=1e-3, dt_max=1e-3, tol=1.0, t_start=0.0, t_end=1.5)
    if method in ("BDF6", "BDF2"):
        cfg["t_end"] = 0.4
    w, c = [0.0, 0.0, 1.0], 25.0
    cap = 2000 if method in ("RK45", "BDF6", "Adams5", "Euler") else 30000  # (the lower orders take many more steps)
    for direction in (0, 1, -1):
        s = make_solver(engine, method, 3, rhs="lorenz", flags=_abi.FLAG_STRICT_FP, history=cap, **cfg) \
            .with_terminal_event(w, c, direction)
        gpu = s.solve_ivp_ensemble(y0, P_LOR, shared_params=True)
        ref = oracle.solve_ensemble(METHODS[method], "lorenz", y0, P_LOR, shared_params=True, pow_mode=1,
                                    history_capacity=cap, event=(w, c, direction), **cfg)
        _bit_exact(gpu, ref)
        np.testing.assert_array_equal(gpu.hist_len, ref["hist_len"])
        stopped = gpu.status == _abi.STOPPED_AT_EVENT
        assert stopped.sum() > n // 4, (method, direction, int(stopped.sum()))
        assert set(np.unique(gpu.status)) <= {_abi.OK, _abi.STOPPED_AT_EVENT, _abi.E_HISTORY_OVERFLOW}
        # on the surface, inside the horizon, and the path holds the points before the event only
        assert np.abs(gpu.y_end[2, stopped] - c).max() < 1e-9
        assert (gpu.t_end[stopped] > 0).all() and (gpu.t_end[stopped] <= cfg["t_end"]).all()
        for i in np.flatnonzero(stopped)[:8]:
            m = int(gpu.hist_len[i])
            _same_bits(gpu.hist[i, :m], ref["hist"][i, :m], f"path {i}")
            assert m == gpu.n_accept[i] and (m == 0 or gpu.hist_t[i, m - 1] <= gpu.t_end[i])
        assert np.abs(gpu.t_end[~stopped] - cfg["t_end"]).max() < 1e-12


def test_terminal_event_is_the_first_event_of_the_full_path(cuda, engine):
    """Stop-at-event and locate-events-on-the-stored-path share the crossing rule and the root finder: the event a
    terminal solve stops at is, bit for bit (strict build), the first event the query finds on the uninterrupted path."""
    n = 512
    y0 = E.lorenz_y0(np.arange(n))
    cfg = dict(LOR, tol=1e-7, t_end=2.0)
    w, c = [1.0, -1.0, 0.0], 0.5  # x - y = 0.5
    for direction in (1, -1, 0):
        full = make_solver(engine, "RK45", 3, rhs="lorenz", flags=_abi.FLAG_STRICT_FP, history=3000, **cfg) \
            .solve_ivp_ensemble(y0, P_LOR, shared_params=True)
        ev, cnt = full.locate_events(w, c, direction, capacity=1)
        term = make_solver(engine, "RK45", 3, rhs="lorenz", flags=_abi.FLAG_STRICT_FP, **cfg) \
            .with_terminal_event(w, c, direction).solve_ivp_ensemble(y0, P_LOR, shared_params=True)
        hit = cnt > 0
        assert hit.sum() > n // 2
        np.testing.assert_array_equal(term.status == _abi.STOPPED_AT_EVENT, hit)
        _same_bits(term.t_end[hit], ev[hit, 0, 0], "event time")
        _same_bits(term.y_end[:, hit].T, ev[hit, 0, 1:], "event state")
        _same_bits(term.y_end[:, ~hit], full.y_end[:, ~hit], "trajectories without an event")


def test_terminal_event_closed_form_and_fast_kernels(cuda, engine, oracle):
    """y = cos(w t) falls through zero at t = pi / 2w; fast kernels, every family; per-trajectory w."""
    n = 1000
    om = np.linspace(1.0, 4.0, n)
    y0 = np.vstack([np.ones(n), np.zeros(n)])
    cfg = dict(dt_min=1e-8, dt_max=0.05, tol=1e-9, t_start=0.0, t_end=3.0)
    for method, tol_t in (("RK45", 2e-7), ("RK23", 2e-6), ("Adams5", 2e-6), ("BDF6", 5e-4)):
        c2 = dict(cfg)
        if method == "BDF6":
            c2.update(tol=1e-7, dt_max=1e-3)
        s = make_solver(engine, method, 2, rhs="harmonic", **c2).with_terminal_event([1.0, 0.0], 0.0, -1)
        r = s.solve_ivp_ensemble(y0, om[None, :])
        assert (r.status == _abi.STOPPED_AT_EVENT).all(), (method, np.unique(r.status))
        assert np.abs(r.t_end - np.pi / (2 * om)).max() < tol_t, (method, np.abs(r.t_end - np.pi / (2 * om)).max())
        assert np.abs(r.y_end[0]).max() < 1e-12 and np.abs(r.y_end[1] + om).max() < 1e-3 * om.max()
    # rising only: cos starts at 1 and first RISES through zero at 3 pi / 2w
    s = make_solver(engine, "RK45", 2, rhs="harmonic", **cfg).with_terminal_event([1.0, 0.0], 0.0, +1)
    r = s.solve_ivp_ensemble(y0, om[None, :])
    want = 3 * np.pi / (2 * om)
    inside = want < 3.0 - 1e-6
    assert (r.status[inside] == _abi.STOPPED_AT_EVENT).all() and (r.status[~inside & (want > 3.0 + 1e-6)] == _abi.OK).all()
    assert np.abs(r.t_end[inside] - want[inside]).max() < 2e-7
    # fast RK45 on config 2's ensemble against the oracle's event (the north-star band on the event state)
    n = 20000
    z0 = E.lorenz_y0(np.arange(n))
    s = make_solver(engine, "RK45", 3, rhs="lorenz", t_end=1.0, **LOR).with_terminal_event([0.0, 0.0, 1.0], 30.0, 1)
    g = s.solve_ivp_ensemble(z0, P_LOR, shared_params=True)
    ref = oracle.solve_ensemble(_abi.RK45, "lorenz", z0, P_LOR, shared_params=True, t_end=1.0, event=([0.0, 0.0, 1.0], 30.0, 1), **LOR)
    np.testing.assert_array_equal(g.status, ref["status"])
    assert rel_err(g.y_end, ref["y_end"]).max() <= band(1e-8) and np.abs(g.t_end - ref["t_end"]).max() < 1e-7
    assert (g.status == _abi.STOPPED_AT_EVENT).sum() > n // 3


def test_terminal_event_linear32_warp_kernels(cuda, engine, oracle):
    """The warp-per-trajectory kernels of config 4 watch the event too: strict build bit-exact with the oracle's
    statement (event point, counters, the path before it), the fast build inside the band; y[0] = 0, every direction."""
    m = 96
    z0, A = E.linear32_problem(np.arange(m))
    A = A.reshape(m, 1024)
    lin = dict(dt_min=1e-7, dt_max=0.05, tol=1e-8, t_start=0.0, t_end=4.0)
    w = np.zeros(32)
    w[0], w[5] = 1.0, -0.5
    for direction in (0, 1, -1):
        s = make_solver(engine, "RK45", 32, rhs="linear32", flags=_abi.FLAG_STRICT_FP, history=300, **lin) \
            .with_terminal_event(w, 0.01, direction)
        gpu = s.solve_ivp_ensemble(z0, A, params_aos=True)
        ref = oracle.solve_ensemble(_abi.RK45, "linear32", z0, A, params_aos=True, pow_mode=1, history_capacity=300,
                                    event=(w, 0.01, direction), **lin)
        _bit_exact(gpu, ref)
        np.testing.assert_array_equal(gpu.hist_len, ref["hist_len"])
        stopped = gpu.status == _abi.STOPPED_AT_EVENT
        assert stopped.sum() > m // 3 and set(np.unique(gpu.status)) <= {_abi.OK, _abi.STOPPED_AT_EVENT}
        assert np.abs(w @ gpu.y_end[:, stopped] - 0.01).max() < 1e-10
        for i in np.flatnonzero(stopped)[:4]:
            k = int(gpu.hist_len[i])
            _same_bits(gpu.hist[i, :k], ref["hist"][i, :k], f"path {i}")
        fast = make_solver(engine, "RK45", 32, rhs="linear32", **lin).with_terminal_event(w, 0.01, direction) \
            .solve_ivp_ensemble(z0, A, params_aos=True)
        ref_f = oracle.solve_ensemble(_abi.RK45, "linear32", z0, A, params_aos=True, event=(w, 0.01, direction), **lin)
        np.testing.assert_array_equal(fast.status, ref_f["status"])
        assert rel_err(fast.y_end, ref_f["y_end"]).max() <= band(1e-8) and np.abs(fast.t_end - ref_f["t_end"]).max() < 1e-7


def test_terminal_event_bad_direction(cuda, engine):
    with pytest.raises(engine.IVPError):
        make_solver(engine, "RK45", 3, rhs="lorenz", t_end=1.0, **LOR).with_terminal_event([0, 0, 1.0], 0.0, 2)


@pytest.mark.skipif("__import__('torch').cuda.device_count() < 2")
def test_options_through_the_multi_gpu_entry(cuda, engine):
    n = 3001
    y0 = E.lorenz_y0(np.arange(n))
    s1 = make_solver(engine, "RK45", 3, rhs="lorenz", t_end=0.7, **LOR)
    g1 = s1.solve_ivp_ensemble(y0, P_LOR, shared_params=True)
    y1, t1, dt1 = g1.restart_record()
    s2 = make_solver(engine, "RK45", 3, rhs="lorenz", t_end=1.4, history=64, **LOR).with_terminal_event([0, 0, 1.0], 28.0, 0)
    one = s2.solve_ivp_ensemble(np.ascontiguousarray(y1), P_LOR, shared_params=True, restart=(t1, dt1))
    two = s2.solve_ivp_ensemble(np.ascontiguousarray(y1), P_LOR, shared_params=True, restart=(t1, dt1), n_gpus=2)
    for k in ("y_end", "t_end", "dt_end", "status", "n_accept", "n_reject", "n_rhs", "hist", "hist_len"):
        assert np.array_equal(getattr(one, k), getattr(two, k)), k


# ---------------------------------------------------------------- paths cut short by their capacity
def test_overflowed_path_ends_at_its_last_record(cuda, engine, oracle):
    """A trajectory with more accepted points than capacity: the stored path covers [t_start, t_cap]; times past its last
    record give NaN and no crossing is invented on the unrecorded span (no closing knot to (t_end, y_end))."""
    n = 64
    y0 = E.lorenz_y0(np.arange(n))
    cap = 40
    cfg = dict(LOR, tol=1e-8, t_end=2.0)
    s = make_solver(engine, "RK45", 3, rhs="lorenz", flags=_abi.FLAG_STRICT_FP, history=cap, **cfg)
    r = s.solve_ivp_ensemble(y0, P_LOR, shared_params=True)
    assert (r.status == _abi.E_HISTORY_OVERFLOW).all() and (r.n_accept > cap).all() and (r.hist_len == cap).all()
    t_last = r.hist_t[:, cap - 1]
    times = np.array([0.0, float(t_last.min()) * 0.5, float(t_last.max()) + 1e-3, 1.5, 2.0])
    smp = r.sample(times)
    assert np.isfinite(smp[:, :2]).all() and np.isnan(smp[:, 2:]).all()
    ref = dict(hist=r.hist, hist_len=r.hist_len, t_end=r.t_end, y_end=r.y_end, n_accept=r.n_accept, status=r.status)
    _same_bits(np.nan_to_num(smp, nan=-7.0), np.nan_to_num(oracle.sample_paths("lorenz", y0, P_LOR, ref, times, t_start=0.0, shared_params=True), nan=-7.0), "samples")
    ev, cnt = r.locate_events([0.0, 0.0, 1.0], 25.0, 0, capacity=16)
    ev_o, cnt_o = oracle.locate_events("lorenz", y0, P_LOR, ref, [0.0, 0.0, 1.0], 25.0, 0, capacity=16, t_start=0.0, shared_params=True)
    np.testing.assert_array_equal(cnt, cnt_o)
    for i in range(n):
        k = min(int(cnt[i]), 16)
        assert (ev[i, :k, 0] <= t_last[i]).all()
    # without n_accept the status array says the same thing
    ref2 = dict(hist=r.hist, hist_len=r.hist_len, t_end=r.t_end, y_end=r.y_end, status=r.status)
    np.testing.assert_array_equal(cnt, oracle.locate_events("lorenz", y0, P_LOR, ref2, [0.0, 0.0, 1.0], 25.0, 0, capacity=16, t_start=0.0, shared_params=True)[1])
    # exactly full is not cut: capacity == n_accept keeps the closing knot rule
    capn = int(r.n_accept[0])
    r1 = make_solver(engine, "RK45", 3, rhs="lorenz", flags=_abi.FLAG_STRICT_FP, history=capn, **cfg) \
        .solve_ivp_ensemble(y0[:, :1], P_LOR, shared_params=True)
    assert r1.status[0] == _abi.OK and np.isfinite(r1.sample(np.array([2.0]))).all()


# ---------------------------------------------------------------- concurrent strict solves keep their own tableau
def test_strict_solves_on_two_streams_keep_their_tableaux(cuda, engine, oracle):
    """RK45 LITERAL, RK45 CORRECTED and RK23 CORRECTED strict launches enqueued back to back on different streams (the
    device entry point is asynchronous): each reads its own constant tableau (one shared __constant__ buffer
    could be overwritten under a running kernel)."""
    torch = cuda
    n = 4096
    y0 = E.lorenz_y0(np.arange(n))
    cfg = dict(LOR, tol=1e-6, t_end=0.5)
    dev = "cuda:0"
    dy0, dp = torch.from_numpy(y0).to(dev), torch.from_numpy(P_LOR).to(dev)
    jobs = [("RK45", _abi.SEM_LITERAL), ("RK45", _abi.SEM_CORRECTED), ("RK23", _abi.SEM_CORRECTED), ("RK23", _abi.SEM_LITERAL)]
    outs = []
    streams = [torch.cuda.Stream(device=dev) for _ in jobs]
    torch.cuda.synchronize()
    for rep in range(3):
        for (m, sem), st in zip(jobs, streams):
            s = make_solver(engine, m, 3, rhs="lorenz", flags=_abi.FLAG_STRICT_FP, semantics=sem, max_attempts=4000, **cfg)
            outs.append((m, sem, s.solve_ivp_ensemble_device(dy0, dp, shared_params=True, stream=st)))
    torch.cuda.synchronize()
    refs = {}
    for m, sem, o in outs:
        if (m, sem) not in refs:
            refs[(m, sem)] = oracle.solve_ensemble(METHODS[m], "lorenz", y0, P_LOR, shared_params=True, semantics=sem,
                                                   pow_mode=1, max_attempts=4000, **cfg)
        r = refs[(m, sem)]
        np.testing.assert_array_equal(o["status"].cpu().numpy(), r["status"])
        _same_bits(o["y_end"].cpu().numpy(), r["y_end"], f"{m} sem {sem}")


# ---------------------------------------------------------------- solve(): the path grows like collect_vec
def test_single_trajectory_path_longer_than_the_default_capacity(cuda, engine):
    """Euler with dt = 1e-5 on [0, 1] yields 100 000 points (ivp.rs:539-560 at a smaller step): the reference's Vec
    grows; here the solve is repeated once with the capacity the first pass reported."""
    s = (engine.Euler.new(1).with_maximum_dt(1e-5).with_initial_time(0.0).with_ending_time(1.0)
         .with_initial_conditions([1.0]).with_derivative("exp"))
    path = s.solve()
    assert len(path) in (100000, 100001) and path[0][0] == 0.0  # (1e5 additions of 1e-5 need not land on 1.0 exactly)
    assert abs(path[-1][1][0] - np.exp(path[-1][0])) < 1e-4
