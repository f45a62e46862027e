"""GPU parity of the RK path (K1/K2) against the CPU oracle, through the C ABI.

strict kernels  : bit-exact with the oracle (same operation order, no FMA contraction)
fast kernels    : final state within max(10*tol, 1e-12) relative, step counts side by side
"""
import numpy as np
import pytest

from bacon_b200 import _abi, ensembles as E
from parity import METHODS, band, make_solver, rel_err, run_both

pytestmark = pytest.mark.gpu

LOR = dict(dt_min=1e-9, dt_max=0.1, tol=1e-8, t_start=0.0)
LOR_P = np.array(E.LORENZ["params"])


def _assert_bit_exact(gpu, ref):
    for k in ("status", "n_accept", "n_reject", "n_rhs"):
        np.testing.assert_array_equal(getattr(gpu, k), ref[k], err_msg=k)
    for k in ("y_end", "t_end", "dt_end"):
        a, b = getattr(gpu, k), ref[k]
        assert np.array_equal(a.view(np.uint64), b.view(np.uint64)), f"{k} differs bitwise: max |d|={np.abs(a - b).max()}"


def test_readme_example_both_semantics(cuda, engine, oracle):
    """BASELINE config 1: y'=y, t in [0,10], tol 1e-4, dt in [0.01,0.1] (README.md:32-40)."""
    y0 = np.array([[1.0]])
    cfg = dict(dt_min=0.01, dt_max=0.1, tol=1e-4, t_start=0.0, t_end=10.0)
    gpu, ref = run_both(engine, oracle, "RK45", "exp", y0, **cfg)
    assert gpu.status[0] == _abi.OK and gpu.n_accept[0] == 128 and gpu.n_reject[0] == 0
    assert rel_err(gpu.y_end, ref["y_end"])[0] <= band(1e-4)
    assert abs(gpu.y_end[0, 0] / np.exp(10.0) - 1.0) < 1e-6
    # strict + corrected: bit-exact
    gpu, ref = run_both(engine, oracle, "RK45", "exp", y0, strict=True, **cfg)
    _assert_bit_exact(gpu, ref)
    # the source exactly as written: 1 accepted point then MinimumTimeDeltaExceeded
    gpu, ref = run_both(engine, oracle, "RK45", "exp", y0, semantics=_abi.SEM_LITERAL, **cfg)
    _assert_bit_exact(gpu, ref)
    assert gpu.status[0] == _abi.E_MIN_DT_EXCEEDED and gpu.n_accept[0] == 1
    assert gpu.t_end[0] == pytest.approx(0.055) and gpu.y_end[0, 0] == pytest.approx(1.055)


@pytest.mark.parametrize("method", ["RK45", "RK23"])
def test_strict_bit_exact_lorenz(cuda, engine, oracle, method):
    n = 4096
    y0 = E.lorenz_y0(np.arange(n))
    gpu, ref = run_both(engine, oracle, method, "lorenz", y0, LOR_P, shared_params=True, strict=True, t_end=0.5,
                        **{**LOR, "tol": 1e-8 if method == "RK45" else 1e-6})
    assert (gpu.status == _abi.OK).all()
    _assert_bit_exact(gpu, ref)


@pytest.mark.parametrize("method,rhs", [("RK45", "vdp"), ("RK23", "vdp"), ("RK45", "robertson"), ("RK45", "harmonic"),
                                        ("RK23", "linear4"), ("RK45", "decay"), ("RK23", "quadratic")])
def test_strict_bit_exact_other_rhs(cuda, engine, oracle, method, rhs):
    n = 1000
    rng = np.random.default_rng(7)
    dim = {"vdp": 2, "robertson": 3, "harmonic": 2, "linear4": 4, "decay": 1, "quadratic": 1}[rhs]
    npar = {"vdp": 1, "robertson": 3, "harmonic": 1, "linear4": 16, "decay": 0, "quadratic": 0}[rhs]
    y0 = rng.uniform(0.1, 1.0, size=(dim, n))
    params = rng.uniform(0.5, 2.0, size=(npar, n)) if npar else None
    if rhs == "linear4":
        params = rng.normal(size=(npar, n)) * 0.5
    t_end = 0.01 if rhs == "robertson" else 2.0
    gpu, ref = run_both(engine, oracle, method, rhs, y0, params, strict=True, dt_min=1e-9, dt_max=0.05, tol=1e-7,
                        t_start=0.0, t_end=t_end)
    _assert_bit_exact(gpu, ref)


def test_literal_semantics_bit_exact(cuda, engine, oracle):
    """REF_LITERAL on a y-dependent RHS: transposed stage matrix with stale k's (rk.rs:459/370)."""
    n = 512
    rng = np.random.default_rng(3)
    y0 = rng.uniform(0.5, 1.5, size=(2, n))
    w = rng.uniform(0.5, 2.0, size=(1, n))
    for method in ("RK45", "RK23"):
        gpu, ref = run_both(engine, oracle, method, "harmonic", y0, w, semantics=_abi.SEM_LITERAL, dt_min=1e-6,
                            dt_max=0.01, tol=1e-3, t_start=0.0, t_end=0.2, max_attempts=20000)
        _assert_bit_exact(gpu, ref)


@pytest.mark.parametrize("method", ["RK45", "RK23"])
def test_reference_rk_tests_literal_equals_corrected(cuda, engine, oracle, method):
    """rk.rs:682-758: y'=-2t and y'=cos t ignore y, so LITERAL == CORRECTED == GPU and every
    yielded point meets the reference's own assertion."""
    for rhs, cfg, exact, eps, n_acc in [
        ("quadratic", dict(dt_min=1e-4, dt_max=0.1, tol=1e-5, t_start=0.0, t_end=10.0), lambda t: 1.0 - t * t, 1e-4, 101),
        ("cos", dict(dt_min=1e-3, dt_max=1e-2, tol=1e-4, t_start=0.0, t_end=10.0), lambda t: np.sin(t), 1e-2, 1001),
    ]:
        y0 = np.array([[1.0 if rhs == "quadratic" else 0.0]])
        outs = []
        for sem, strict in ((0, False), (0, True), (1, True)):
            gpu, ref = run_both(engine, oracle, method, rhs, y0, semantics=sem, strict=strict, history=1100, **cfg)
            assert gpu.status[0] == _abi.OK and gpu.n_accept[0] == n_acc and gpu.n_reject[0] == 0
            assert ref["n_accept"][0] == n_acc
            m = int(gpu.hist_len[0])
            assert m == n_acc
            t, y = gpu.hist_t[0, :m], gpu.hist_y[0, :m, 0]
            assert np.abs(y - exact(t)).max() <= eps  # the reference's assertion, every yielded point
            np.testing.assert_allclose(t, ref["hist_t"][0, :m], rtol=1e-12, atol=1e-13)
            np.testing.assert_allclose(y, ref["hist_y"][0, :m, 0], rtol=1e-9, atol=1e-12)
            outs.append(gpu.y_end[0, 0])
        assert max(outs) - min(outs) <= 1e-9 * max(1.0, abs(outs[0]))


@pytest.mark.parametrize("method,tol,t_end", [("RK45", 1e-8, 5.0), ("RK23", 1e-6, 2.0)])
def test_fast_lorenz_within_band(cuda, engine, oracle, method, tol, t_end):
    """The headline kernel against the oracle (libm pow, no FMA) on the seeded config-2 ensemble."""
    n = 16384
    y0 = E.lorenz_y0(np.arange(n))
    gpu, ref = run_both(engine, oracle, method, "lorenz", y0, LOR_P, shared_params=True, t_end=t_end,
                        **{**LOR, "tol": tol})
    assert (gpu.status == _abi.OK).all() and (ref["status"] == _abi.OK).all()
    err = rel_err(gpu.y_end, ref["y_end"])
    assert err.max() <= band(tol), f"worst relative error {err.max():.3e} > {band(tol):.1e}"
    np.testing.assert_array_equal(gpu.t_end, ref["t_end"])
    # step counts side by side: identical up to a handful of borderline accept/reject decisions
    d_acc = np.abs(gpu.n_accept.astype(np.int64) - ref["n_accept"].astype(np.int64))
    d_rej = np.abs(gpu.n_reject.astype(np.int64) - ref["n_reject"].astype(np.int64))
    assert d_acc.max() <= 3 and d_rej.max() <= 3, (d_acc.max(), d_rej.max())
    assert abs(int(gpu.n_accept.sum()) - int(ref["n_accept"].sum())) <= 1e-4 * ref["n_accept"].sum()
    np.testing.assert_array_equal(gpu.n_rhs, (gpu.n_accept + gpu.n_reject) * (6 if method == "RK45" else 4))


def test_fast_vdp_mu_sweep_within_band(cuda, engine, oracle):
    """Config 3 shape: per-trajectory mu, strongly divergent step counts -> work-queue refill path."""
    n = 8192
    y0, mu = E.vdp_problem(np.arange(n) * (E.VDP["n"] // n), E.VDP["n"])
    cfg = dict(dt_min=1e-12, dt_max=0.1, tol=1e-10, t_start=0.0, t_end=0.02)
    gpu, ref = run_both(engine, oracle, "RK23", "vdp", y0, mu, **cfg)
    assert (gpu.status == _abi.OK).all()
    assert rel_err(gpu.y_end, ref["y_end"]).max() <= band(1e-10)
    assert gpu.n_accept.max() > 3 * gpu.n_accept.min()  # the spread that motivates the refill
    d_acc = np.abs(gpu.n_accept.astype(np.int64) - ref["n_accept"].astype(np.int64))
    assert d_acc.max() <= 3


def test_dense_output_matches_oracle_path(cuda, engine, oracle):
    """K4: every accepted (t, y) of every trajectory, one record per accepted point written straight to the
    trajectory-major history (one 256-bit store for D = 3; no staging: profiles/r01i_dense_output.md)."""
    n = 3000  # not a multiple of the block size; refill + partial flushes
    y0 = E.lorenz_y0(np.arange(n))
    for strict in (True, False):
        for cap in (600, 37):  # 37: odd capacity (scalar store path) and overflow
            gpu, ref = run_both(engine, oracle, "RK45", "lorenz", y0, LOR_P, shared_params=True, strict=strict,
                                history=cap, t_end=0.5, **LOR)
            np.testing.assert_array_equal(gpu.hist_len, ref["hist_len"])
            np.testing.assert_array_equal(gpu.status, ref["status"])
            if cap == 37:
                assert (gpu.status == _abi.E_HISTORY_OVERFLOW).all()
            mask = np.arange(cap)[None, :] < gpu.hist_len[:, None]
            if strict:
                assert np.array_equal(gpu.hist_t[mask], ref["hist_t"][mask])
                assert np.array_equal(gpu.hist_y[mask], ref["hist_y"][mask])
                assert (gpu.hist_t[~mask] == 0).all()  # nothing written beyond hist_len
            else:
                if not np.array_equal(gpu.n_accept, ref["n_accept"]):
                    continue  # a borderline accept/reject flip shifts indices; covered by the final-state band
                # The embedded error estimate is a cancelling sum (sum_j e_j = 0): ~1e-6 relative rounding
                # noise, different with and without FMA, so dt (prop. to err^-1/4) jitters by ~1e-7 relative
                # and the two time grids drift apart by up to ~1e-7.  Compare y after removing the
                # first-order effect of that grid offset: y_gpu(t_gpu) ~ y_ref(t_ref) + f(y_ref) (t_gpu - t_ref).
                dt_grid = gpu.hist_t - ref["hist_t"]
                assert np.abs(dt_grid[mask]).max() <= 1e-6
                yr = ref["hist_y"]
                f = np.stack([LOR_P[0] * (yr[..., 1] - yr[..., 0]), yr[..., 0] * (LOR_P[1] - yr[..., 2]) - yr[..., 1],
                              yr[..., 0] * yr[..., 1] - LOR_P[2] * yr[..., 2]], axis=-1)
                resid = gpu.hist_y - (yr + f * dt_grid[..., None])
                scale = np.sqrt((yr ** 2).sum(-1))
                assert (np.sqrt((resid ** 2).sum(-1))[mask] / scale[mask]).max() <= band(LOR["tol"])
            # the last history point of a completed trajectory is the final state
            ok = (gpu.status == _abi.OK)
            last = gpu.hist_len.astype(np.int64) - 1
            np.testing.assert_array_equal(gpu.hist_y[np.arange(n)[ok], last[ok]], gpu.y_end[:, ok].T)


def test_regrouping_and_dense_output_do_not_change_a_trajectory(cuda, engine):
    """Scheduling must be invisible: a trajectory's numbers depend on its inputs only.  150 000 trajectories take more
    than one per lane, so the work counter runs dry and the CTAs regroup: warps report in at their checkpoints, live
    trajectories are suspended, compacted through shared memory and resumed on other lanes, freed warps exit
    (drive.cuh) — each such trajectory's history is written partly by one lane and partly by another (hist_stage.cuh);
    lanes take their trajectories from per-warp blocks (WarpQueue).  The same trajectories solved in three smaller
    launches (one trajectory per lane from the static first deal, regrouped at other moments) must give the same bits:
    final states, counters and every history record."""
    n, cap = 150_000, 512
    y0 = E.lorenz_y0(np.arange(n))
    s = make_solver(engine, "RK45", 3, rhs="lorenz", t_end=0.15, history=cap, **LOR)
    a = s.solve_ivp_ensemble(y0, LOR_P, shared_params=True)
    assert engine.last_launch()["n_kernels"] == 1 and engine.last_launch()["block"] == 640  # one wide CTA per SM
    assert (a.status == _abi.OK).all() and len(set(a.hist_len % 4)) == 4
    for lo in range(0, n, 50_000):
        b = s.solve_ivp_ensemble(np.ascontiguousarray(y0[:, lo:lo + 50_000]), LOR_P, shared_params=True)
        assert engine.last_launch()["n_kernels"] == 1
        sl = slice(lo, lo + 50_000)
        for k in ("y_end", "t_end", "dt_end"):
            assert np.array_equal(getattr(a, k)[..., sl].view(np.uint64), getattr(b, k).view(np.uint64)), k
        for k in ("n_accept", "n_reject", "n_rhs", "hist_len"):
            np.testing.assert_array_equal(getattr(a, k)[sl], getattr(b, k), err_msg=k)
        assert np.array_equal(a.hist_t[sl].view(np.uint64), b.hist_t.view(np.uint64))
        assert np.array_equal(a.hist_y[sl].view(np.uint64), b.hist_y.view(np.uint64))
    mask = np.arange(cap)[None, :] < a.hist_len[:, None]
    assert (a.hist_t[~mask] == 0).all() and (np.diff(a.hist_t, axis=1)[mask[:, 1:]] > 0).all()
    last = a.hist_len.astype(np.int64) - 1
    np.testing.assert_array_equal(a.hist_y[np.arange(n), last], a.y_end.T)
    # final state only, and a history overflow that ends after a regrouping
    s0 = make_solver(engine, "RK45", 3, rhs="lorenz", t_end=0.15, **LOR)
    c = s0.solve_ivp_ensemble(y0, LOR_P, shared_params=True)
    assert np.array_equal(c.y_end.view(np.uint64), a.y_end.view(np.uint64))
    s1 = make_solver(engine, "RK45", 3, rhs="lorenz", t_end=0.15, history=64, **LOR)
    d = s1.solve_ivp_ensemble(y0, LOR_P, shared_params=True)
    over = a.n_accept > 64
    assert over.sum() > n // 2 and (d.status[over] == _abi.E_HISTORY_OVERFLOW).all() and (d.status[~over] == _abi.OK).all()
    np.testing.assert_array_equal(d.hist_len, np.minimum(a.n_accept, 64))
    assert np.array_equal(d.hist_y.view(np.uint64), a.hist_y[:, :64].view(np.uint64))


def test_work_queue_blocks_strict_history_bit_exact(cuda, engine, oracle):
    """More trajectories than lanes with dense output: lanes refill from their warp's block of consecutive indices
    (WarpQueue, drive.cuh), blocks shrink towards the end of the ensemble, and a ragged n leaves partial blocks.  The
    strict kernels (they do not migrate: no regrouping) must still match the oracle bit for bit, every record."""
    for n in (130_001, 114_000):  # just above the 113 664 resident lanes (6 CTAs of 128 per SM): most lanes refill once, from small blocks
        y0 = E.lorenz_y0(np.arange(n))
        gpu, ref = run_both(engine, oracle, "RK45", "lorenz", y0, LOR_P, shared_params=True, strict=True, history=48,
                            t_end=0.03, **LOR)
        _assert_bit_exact(gpu, ref)
        np.testing.assert_array_equal(gpu.hist_len, ref["hist_len"])
        mask = np.arange(48)[None, :] < gpu.hist_len[:, None]
        assert np.array_equal(gpu.hist_t[mask], ref["hist_t"][mask]) and np.array_equal(gpu.hist_y[mask], ref["hist_y"][mask])
        assert (gpu.hist_t[~mask] == 0).all()


def test_tiny_grid_refill_dry_regroup(cuda, engine, oracle, monkeypatch):
    """BACON_IVP_GRID caps the grid: 2 CTAs for 3000 trajectories make every lane refill several times, the counter run
    dry and the CTAs regroup until their last warp — on sizes where the oracle checks every record."""
    n = 3000
    y0 = E.lorenz_y0(np.arange(n))
    ref_fast = run_both(engine, oracle, "RK45", "lorenz", y0, LOR_P, shared_params=True, history=160, t_end=0.15, **LOR)[0]
    monkeypatch.setenv("BACON_IVP_GRID", "2")
    gpu, ref = run_both(engine, oracle, "RK45", "lorenz", y0, LOR_P, shared_params=True, strict=True, history=160,
                        t_end=0.15, **LOR)
    assert engine.last_launch()["grid"] == 2 and engine.last_launch()["block"] == 128
    _assert_bit_exact(gpu, ref)
    mask = np.arange(160)[None, :] < gpu.hist_len[:, None]
    assert np.array_equal(gpu.hist_t[mask], ref["hist_t"][mask]) and np.array_equal(gpu.hist_y[mask], ref["hist_y"][mask])
    fast = run_both(engine, oracle, "RK45", "lorenz", y0, LOR_P, shared_params=True, history=160, t_end=0.15, **LOR)[0]
    assert engine.last_launch()["grid"] == 2 and engine.last_launch()["block"] == 640
    for k in ("y_end", "t_end", "dt_end"):  # the same bits as on the full grid, where no lane ever refilled
        assert np.array_equal(getattr(fast, k).view(np.uint64), getattr(ref_fast, k).view(np.uint64)), k
    np.testing.assert_array_equal(fast.n_accept, ref_fast.n_accept)
    assert np.array_equal(fast.hist_y.view(np.uint64), ref_fast.hist_y.view(np.uint64))


def test_failure_statuses_never_abort_the_batch(cuda, engine, oracle):
    """Per-trajectory failures land in status[i] (the reference aborts one trajectory, ivp.rs:232-235)."""
    n = 256
    y0 = E.lorenz_y0(np.arange(n))
    # (a) dt_min too large -> MinimumTimeDeltaExceeded for every trajectory, same as the oracle
    for strict in (True, False):
        gpu, ref = run_both(engine, oracle, "RK45", "lorenz", y0, LOR_P, shared_params=True, strict=strict,
                            dt_min=5e-3, dt_max=0.1, tol=1e-8, t_start=0.0, t_end=5.0)
        np.testing.assert_array_equal(gpu.status, ref["status"])
        assert (gpu.status == _abi.E_MIN_DT_EXCEEDED).all()
        np.testing.assert_array_equal(gpu.n_accept, ref["n_accept"])
    # (b) attempt cap
    gpu, ref = run_both(engine, oracle, "RK45", "lorenz", y0, LOR_P, shared_params=True, t_end=5.0, max_attempts=100, **LOR)
    np.testing.assert_array_equal(gpu.status, ref["status"])
    assert (gpu.status == _abi.E_MAX_ATTEMPTS).all() and (gpu.n_accept + gpu.n_reject == 100).all()
    # (c) NaN initial condition in ONE trajectory: NonFinite there (the reference would spin forever), rest fine
    y0b = y0.copy()
    y0b[1, 17] = np.nan
    gpu, ref = run_both(engine, oracle, "RK45", "lorenz", y0b, LOR_P, shared_params=True, t_end=0.2, **LOR)
    assert gpu.status[17] == _abi.E_NONFINITE and ref["status"][17] == _abi.E_NONFINITE
    ok = np.arange(n) != 17
    assert (gpu.status[ok] == _abi.OK).all()
    assert rel_err(gpu.y_end[:, ok], ref["y_end"][:, ok]).max() <= band(1e-8)


def test_fast_controller_corner_cases(cuda, engine, oracle):
    """The fast stepper decides "common case" from high words and evaluates the step-size factor on floats re-biased
    into a per-ensemble exponent window (rk_fast.cuh).  Everything outside that window must still take the
    reference's decisions: compare step COUNTS with the oracle on problems built to leave it."""
    def counts_match(gpu, ref, slack=1):
        np.testing.assert_array_equal(gpu.status, ref["status"])
        d_acc = np.abs(gpu.n_accept.astype(np.int64) - ref["n_accept"].astype(np.int64))
        d_rej = np.abs(gpu.n_reject.astype(np.int64) - ref["n_reject"].astype(np.int64))
        assert d_acc.max() <= slack and d_rej.max() <= slack, (d_acc.max(), d_rej.max())

    one = np.ones((1, 64))
    # (a) error estimate exactly zero (RK45 integrates y' = -2t exactly): factor 4 until dt_max, then pinned there
    for method in ("RK45", "RK23"):
        gpu, ref = run_both(engine, oracle, method, "quadratic", one, dt_min=1e-6, dt_max=0.1, tol=1e-6, t_start=0.0,
                            t_end=3.0)
        counts_match(gpu, ref, slack=0 if method == "RK45" else 1)
        assert rel_err(gpu.y_end, ref["y_end"]).max() <= band(1e-6)
    # (b) the solution decays to 1e-130: the squared error estimate (~1e-280) is far below the window -> factor 4
    gpu, ref = run_both(engine, oracle, "RK45", "decay", one, dt_min=1e-6, dt_max=0.5, tol=1e-8, t_start=0.0, t_end=300.0)
    counts_match(gpu, ref)
    # (the error test is absolute, rk.rs:386-392: at 5e-131 the two solutions agree to ~1e-4 relative, 1e-134 absolute)
    assert (gpu.status == _abi.OK).all() and np.abs(gpu.y_end - ref["y_end"]).max() <= 1e-8
    # (c) no window: (tol dt_max)^2 >= 1, and (tol dt_min)^2 underflows -> every attempt takes the exact path
    y0 = E.lorenz_y0(np.arange(64))
    gpu, ref = run_both(engine, oracle, "RK45", "harmonic", np.vstack([one, 0 * one]), np.full((1, 64), 2.0),
                        dt_min=1e-6, dt_max=4.0, tol=0.5, t_start=0.0, t_end=50.0)
    counts_match(gpu, ref)
    gpu, ref = run_both(engine, oracle, "RK45", "lorenz", y0, LOR_P, shared_params=True, dt_min=1e-200, dt_max=0.1,
                        tol=1e-8, t_start=0.0, t_end=1.0)
    counts_match(gpu, ref, slack=3)
    assert rel_err(gpu.y_end, ref["y_end"]).max() <= band(1e-8)
    # (d) a narrow dt range that pins dt at dt_min's edge and at dt_max in turn (both clamps live on the exact path)
    gpu, ref = run_both(engine, oracle, "RK45", "lorenz", y0, LOR_P, shared_params=True, dt_min=1e-3, dt_max=2e-3,
                        tol=1e-6, t_start=0.0, t_end=1.0)
    counts_match(gpu, ref, slack=3)
    # (e) an infinity in the state: inf - inf inside the RHS GENERATES a NaN (sign bit set on this hardware) -> NonFinite
    y0b = y0.copy()
    y0b[0, 5] = np.inf
    gpu, ref = run_both(engine, oracle, "RK45", "lorenz", y0b, LOR_P, shared_params=True, t_end=0.2, **LOR)
    assert gpu.status[5] == _abi.E_NONFINITE and ref["status"][5] == _abi.E_NONFINITE
    assert (np.delete(gpu.status, 5) == _abi.OK).all()
    # (f) a time axis that ends exactly where a step lands, negative start time
    gpu, ref = run_both(engine, oracle, "RK45", "lorenz", y0, LOR_P, shared_params=True, dt_min=1e-9, dt_max=0.1, tol=1e-8,
                        t_start=-2.0, t_end=-1.0)
    counts_match(gpu, ref, slack=3)
    np.testing.assert_array_equal(gpu.t_end, ref["t_end"])
    assert rel_err(gpu.y_end, ref["y_end"]).max() <= band(1e-8)


def test_edge_sizes(cuda, engine, oracle):
    """n = 1, ragged n around warp/block multiples, and n = 0."""
    for n in (1, 31, 33, 127, 129, 1025):
        y0 = E.lorenz_y0(np.arange(n))
        gpu, ref = run_both(engine, oracle, "RK45", "lorenz", y0, LOR_P, shared_params=True, strict=True, t_end=0.1, **LOR)
        assert np.array_equal(gpu.y_end.view(np.uint64), ref["y_end"].view(np.uint64))
    from parity import make_solver
    s = make_solver(engine, "RK45", 3, rhs="lorenz", t_end=0.1, **LOR)
    r = s.solve_ivp_ensemble(np.zeros((3, 0)), LOR_P, shared_params=True)
    assert r.y_end.shape == (3, 0)


def test_full_size_properties(cuda, engine):
    """BASELINE config 2 at full size (2^20 trajectories, T=5): size-independent properties —
    every trajectory reaches t_end exactly, the run is deterministic (bitwise idempotent), and the
    ensemble mean of z stays on the Lorenz attractor's range."""
    from parity import make_solver
    n = E.LORENZ["n"]
    y0 = E.lorenz_y0(np.arange(n))
    s = make_solver(engine, "RK45", 3, rhs="lorenz", t_end=E.LORENZ["t_end"], **LOR)
    a = s.solve_ivp_ensemble(y0, LOR_P, shared_params=True)
    b = s.solve_ivp_ensemble(y0, LOR_P, shared_params=True)
    assert (a.status == _abi.OK).all() and (a.t_end == E.LORENZ["t_end"]).all()
    assert np.array_equal(a.y_end.view(np.uint64), b.y_end.view(np.uint64))
    np.testing.assert_array_equal(a.n_accept, b.n_accept)
    assert 2500 < a.n_accept.mean() < 4500 and a.n_reject.sum() < 0.01 * a.n_accept.sum()
    assert np.isfinite(a.y_end).all() and 15.0 < a.y_end[2].mean() < 32.0


# ---------------------------------------------------------------- K5: D = 32, warp per trajectory
def _linear32(n):
    y0, A = E.linear32_problem(np.arange(n))
    return y0, A


def test_linear32_strict_bit_exact_and_history(cuda, engine, oracle):
    n = 300
    y0, A = _linear32(n)
    cfg = dict(dt_min=1e-9, dt_max=0.1, tol=1e-8, t_start=0.0, t_end=1.0)
    for method in ("RK45", "RK23"):
        s = make_solver(engine, method, 32, rhs="linear32", flags=_abi.FLAG_STRICT_FP, history=64, **cfg)
        gpu = s.solve_ivp_ensemble(y0, A, params_aos=True)
        ref = oracle.solve_ensemble(METHODS[method], "linear32", y0, A, params_aos=True, history_capacity=64, pow_mode=1, **cfg)
        _assert_bit_exact(gpu, ref)
        np.testing.assert_array_equal(gpu.hist_len, ref["hist_len"])
        mask = np.arange(64)[None, :] < gpu.hist_len[:, None]
        assert np.array_equal(gpu.hist_t[mask], ref["hist_t"][mask])
        assert np.array_equal(gpu.hist_y[mask], ref["hist_y"][mask])
    # SoA parameter layout gives the same bits as AoS
    s = make_solver(engine, "RK45", 32, rhs="linear32", flags=_abi.FLAG_STRICT_FP, **cfg)
    a = s.solve_ivp_ensemble(y0, A, params_aos=True)
    b = s.solve_ivp_ensemble(y0, np.ascontiguousarray(A.reshape(n, 1024).T))
    assert np.array_equal(a.y_end.view(np.uint64), b.y_end.view(np.uint64))


def test_linear32_fast_within_band_and_matrix_exponential(cuda, engine, oracle):
    """Config 4 shape (T = 4, dense output, capacity 256) against the oracle and against expm(A T) y0."""
    import json
    import os
    n = 512
    y0, A = _linear32(n)
    cfg = dict(dt_min=1e-9, dt_max=0.1, tol=1e-8, t_start=0.0, t_end=4.0)
    s = make_solver(engine, "RK45", 32, rhs="linear32", history=256, **cfg)
    gpu = s.solve_ivp_ensemble(y0, A, params_aos=True)
    ref = oracle.solve_ensemble(_abi.RK45, "linear32", y0, A, params_aos=True, history_capacity=256, **cfg)
    assert (gpu.status == _abi.OK).all() and (ref["status"] == _abi.OK).all()
    assert rel_err(gpu.y_end, ref["y_end"]).max() <= band(1e-8)
    assert np.abs(gpu.n_accept.astype(int) - ref["n_accept"].astype(int)).max() <= 2
    same = gpu.n_accept == ref["n_accept"]
    mask = (np.arange(256)[None, :] < gpu.hist_len[:, None]) & same[:, None]
    np.testing.assert_allclose(gpu.hist_t[mask], ref["hist_t"][mask], rtol=1e-6)
    np.testing.assert_allclose(gpu.hist_y[mask], ref["hist_y"][mask], rtol=0, atol=1e-6)
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "anchors.json")))["linear32_seeded4_T4"]
    gold = np.array(gold).T  # (32, 4)
    assert rel_err(gpu.y_end[:, :4], gold).max() <= 1e-6  # global error of RKF45 at tol 1e-8 over T=4


def test_pinned_and_zero_copy_host_paths_agree(cuda, engine):
    """bacon_host_alloc buffers: staged (async H2D / kernel / D2H) and zero-copy (the kernel reads and writes
    the pinned host buffers itself) give the same bits as pageable numpy inputs; the result arrays are pinned."""
    from bacon_b200.ivp import pinned_empty
    n = 5000
    y0 = E.lorenz_y0(np.arange(n))
    s = make_solver(engine, "RK45", 3, rhs="lorenz", t_end=0.6, **LOR)
    a = s.solve_ivp_ensemble(y0, LOR_P, shared_params=True, zero_copy=False)
    y0p = pinned_empty(y0.shape)
    y0p[...] = y0
    pp = pinned_empty(LOR_P.shape)
    pp[...] = LOR_P
    b = s.solve_ivp_ensemble(y0p, pp, shared_params=True, zero_copy=False)
    c = s.solve_ivp_ensemble(y0p, pp, shared_params=True, zero_copy=True)
    assert b.launch["d2h_ms"] > 0
    assert (a.status == _abi.OK).all()
    for k in ("y_end", "t_end", "dt_end", "status", "n_accept", "n_reject", "n_rhs"):
        np.testing.assert_array_equal(getattr(a, k), getattr(b, k), err_msg=k)
        np.testing.assert_array_equal(getattr(a, k), getattr(c, k), err_msg=k)
    assert c.launch["h2d_ms"] == 0 and c.launch["d2h_ms"] == 0 and c.launch["kernel_ms"] > 0
    # zero_copy with a pageable input silently takes the staged path
    d = s.solve_ivp_ensemble(y0, LOR_P, shared_params=True, zero_copy=True)
    np.testing.assert_array_equal(a.y_end, d.y_end)
    # arrays outlive the result object (they own their pinned block)
    ye = c.y_end
    del a, b, c, d
    assert np.isfinite(ye).all()


def test_multi_gpu_host_entry_matches_single_gpu(cuda, engine):
    """bacon_ivp_solve_ensemble_multi: trajectories dealt i mod G over the GPUs of this process,
    bit-identical with the single-GPU solve, ragged n, per-trajectory parameters and dense output included."""
    g = cuda.cuda.device_count()
    if g < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    g = min(g, 4)
    n = 10007
    y0 = E.lorenz_y0(np.arange(n))
    s = make_solver(engine, "RK45", 3, rhs="lorenz", t_end=0.5, **LOR)
    one = s.solve_ivp_ensemble(y0, LOR_P, shared_params=True)
    many = s.solve_ivp_ensemble(y0, LOR_P, shared_params=True, n_gpus=g)
    for k in ("y_end", "t_end", "dt_end", "status", "n_accept", "n_reject", "n_rhs"):
        np.testing.assert_array_equal(getattr(one, k), getattr(many, k), err_msg=k)
    assert many.launch["n_kernels"] == g
    idx = np.arange(1001)
    y0v, mu = E.vdp_problem(idx * 4000, 1 << 22)
    sv = make_solver(engine, "RK23", 2, rhs="vdp", dt_min=1e-12, dt_max=0.1, tol=1e-8, t_start=0.0, t_end=0.05, history=64)
    a = sv.solve_ivp_ensemble(y0v, mu)
    b = sv.solve_ivp_ensemble(y0v, mu, n_gpus=g)
    for k in ("y_end", "status", "n_accept", "hist_len", "hist_t", "hist_y"):
        np.testing.assert_array_equal(getattr(a, k), getattr(b, k), err_msg=k)


def test_concurrent_solves_on_two_streams_lose_nothing(cuda, engine):
    """Two solves in flight on different streams share the SMs (each launch has its own work counter and its own
    CTAs: nothing in the regrouping depends on where a CTA or a warp landed).  Status arrays start at -1: every
    trajectory of both ensembles must have been finished, with the bits of the same solve run alone."""
    import torch
    n = 200_000  # more than one trajectory per lane: refills, dry counter, regrouping — in both launches at once
    s = make_solver(engine, "RK45", 3, rhs="lorenz", t_end=0.1, **LOR)
    p = torch.from_numpy(LOR_P).cuda()
    ya = torch.from_numpy(E.lorenz_y0(np.arange(n))).cuda()
    yb = torch.from_numpy(E.lorenz_y0(np.arange(n, 2 * n))).cuda()
    alone_a = {k: v.clone() for k, v in s.solve_ivp_ensemble_device(ya, p, shared_params=True).items()}
    alone_b = {k: v.clone() for k, v in s.solve_ivp_ensemble_device(yb, p, shared_params=True).items()}
    torch.cuda.synchronize()
    st_a, st_b = torch.cuda.Stream(), torch.cuda.Stream()
    for _ in range(3):
        oa = s.solve_ivp_ensemble_device(ya, p, shared_params=True, stream=st_a)
        ob = s.solve_ivp_ensemble_device(yb, p, shared_params=True, stream=st_b)
        torch.cuda.synchronize()
        for got, ref in ((oa, alone_a), (ob, alone_b)):
            assert (got["status"] == _abi.OK).all()
            for k in ("y_end", "t_end", "dt_end", "n_accept", "n_reject"):
                assert torch.equal(got[k], ref[k]), k


@pytest.mark.parametrize("n", [700, 769, 1000, 1536, 1537, 2000, 2305, 5000])
def test_regrouping_edges_on_one_cta(cuda, engine, monkeypatch, n):
    """One wide CTA (BACON_IVP_GRID=1) and ensembles around the sizes where the driver changes regime (drive.cuh): fewer
    trajectories than lanes (ragged last bundle, regrouping only), one more than the lanes (a single refill), around
    two per lane, several per lane.  Every trajectory must be finished exactly once, with the bits of the same solve
    on the full grid (one trajectory per lane)."""
    y0 = E.lorenz_y0(np.arange(n))
    for hist in (0, 40):
        s = make_solver(engine, "RK45", 3, rhs="lorenz", t_end=0.12, history=hist, **LOR)
        monkeypatch.delenv("BACON_IVP_GRID", raising=False)
        ref = s.solve_ivp_ensemble(y0, LOR_P, shared_params=True)
        monkeypatch.setenv("BACON_IVP_GRID", "1")
        got = s.solve_ivp_ensemble(y0, LOR_P, shared_params=True)
        assert engine.last_launch()["grid"] == 1
        assert (got.status == ref.status).all() and (got.status != -1).all()
        for k in ("y_end", "t_end", "dt_end"):
            assert np.array_equal(getattr(got, k).view(np.uint64), getattr(ref, k).view(np.uint64)), (k, hist)
        for k in ("n_accept", "n_reject", "n_rhs"):
            np.testing.assert_array_equal(getattr(got, k), getattr(ref, k), err_msg=k)
        if hist:
            np.testing.assert_array_equal(got.hist_len, ref.hist_len)
            assert np.array_equal(got.hist_y.view(np.uint64), ref.hist_y.view(np.uint64))
