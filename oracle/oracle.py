"""Python binding of the CPU ORACLE (oracle/liboracle.so).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module; nothing under bacon_b200/ does.

The oracle restates src/ivp/rk.rs:361-423, src/ivp/bdf.rs:346-634 and
src/ivp.rs:220-238 of aftix/bacon (see oracle/bacon_oracle.hpp for the pinning
statement).
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))
from bacon_b200 import _abi  # noqa: E402  (struct layouts only)

_LIB = None


def build(force=False):
    """Compile oracle/liboracle.so with the committed Makefile."""
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("oracle_capi.cpp", "bacon_oracle.hpp", "Makefile")]
    srcs.append(os.path.join(os.path.dirname(_HERE), "include", "bacon_ivp.h"))
    stale = force or not os.path.exists(so) or any(
        os.path.getmtime(s) > os.path.getmtime(so) for s in srcs)
    if stale:
        subprocess.run(["make", "-C", _HERE, "-B", "liboracle.so"], check=True,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(so):
            build()
        L = C.CDLL(so)
        L.oracle_rhs_lookup.argtypes = [C.c_char_p]
        L.oracle_rhs_lookup.restype = C.c_int
        L.oracle_rhs_info.argtypes = [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.oracle_ivp_solve_ensemble.argtypes = [
            C.POINTER(_abi.Config), C.c_int, C.c_size_t, C.c_void_p, C.c_void_p,
            C.POINTER(_abi.Result), C.c_int, C.c_int]
        L.oracle_ivp_solve_ensemble.restype = C.c_int
        L.oracle_ivp_solve_ensemble_ex.argtypes = [
            C.POINTER(_abi.Config), C.c_int, C.c_size_t, C.c_void_p, C.c_void_p, C.POINTER(_abi.Options),
            C.POINTER(_abi.Result), C.c_int, C.c_int]
        L.oracle_ivp_solve_ensemble_ex.restype = C.c_int
        L.oracle_roots_secant.argtypes = [C.c_int, C.c_void_p, C.c_double, C.c_double, C.c_int,
                                          C.c_int, C.c_void_p, C.POINTER(C.c_ulonglong)]
        L.oracle_roots_secant.restype = C.c_int
        L.oracle_fourth_root.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
        L.oracle_fourth_root.restype = None
        L.oracle_nth_root.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_void_p]
        L.oracle_nth_root.restype = None
        L.oracle_max_threads.restype = C.c_int
        L.oracle_coefficients.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.oracle_coefficients.restype = C.c_int
        L.oracle_sample_paths.argtypes = [C.POINTER(_abi.Config), C.c_int, C.c_size_t, C.c_void_p, C.c_void_p,
                                          C.POINTER(_abi.Result), C.c_size_t, C.c_void_p, C.c_void_p]
        L.oracle_sample_paths.restype = C.c_int
        L.oracle_locate_events.argtypes = [C.POINTER(_abi.Config), C.c_int, C.c_size_t, C.c_void_p, C.c_void_p,
                                           C.POINTER(_abi.Result), C.c_void_p, C.c_double, C.c_int, C.c_int,
                                           C.c_void_p, C.c_void_p]
        L.oracle_locate_events.restype = C.c_int
        _LIB = L
    return _LIB


def max_threads():
    return lib().oracle_max_threads()


def rhs_info(name):
    L = lib()
    rid = L.oracle_rhs_lookup(name.encode())
    if rid < 0:
        raise KeyError(name)
    d, p = C.c_int(), C.c_int()
    L.oracle_rhs_info(rid, C.byref(d), C.byref(p))
    return rid, d.value, p.value


def solve_ensemble(method, rhs, y0, params=None, *, dt_min, dt_max, tol, t_start, t_end,
                   semantics=_abi.SEM_CORRECTED, shared_params=False, params_aos=False, history_capacity=0,
                   max_attempts=0, pow_mode=0, n_threads=0, bdf_newton=False, dt_init=0.0, t_start_each=None,
                   dt_start_each=None, event=None):
    """Run the oracle on an ensemble.  y0: (dim, n) float64; params: (n_params, n) or (n_params,).
    Optional inputs as in bacon_ivp_options: dt_init, per-trajectory t_start_each / dt_start_each (restart record),
    event = (w, c, direction) terminal event.

    Returns a dict of numpy arrays with the same names/layouts as bacon_ivp_result.
    """
    L = lib()
    rid, dim, npar = rhs_info(rhs)
    y0 = np.ascontiguousarray(y0, dtype=np.float64)
    assert y0.ndim == 2 and y0.shape[0] == dim, (y0.shape, dim)
    n = y0.shape[1]
    flags = _abi.FLAG_BDF_NEWTON if bdf_newton else 0
    pptr = None
    if npar > 0:
        params = np.ascontiguousarray(params, dtype=np.float64)
        if shared_params:
            assert params.shape == (npar,)
            flags |= _abi.FLAG_SHARED_PARAMS
        elif params_aos:
            params = params.reshape(n, npar)
            flags |= _abi.FLAG_PARAMS_AOS
        else:
            assert params.shape == (npar, n), (params.shape, npar, n)
        pptr = params.ctypes.data
    cfg = _abi.Config(method=method, dim=dim, n_params=npar, semantics=semantics, flags=flags,
                      history_capacity=history_capacity, dt_min=dt_min, dt_max=dt_max, tol=tol,
                      t_start=t_start, t_end=t_end, max_attempts=max_attempts, dt_init=dt_init)
    opts = _abi.Options()
    keep = []
    if t_start_each is not None:
        keep.append(np.ascontiguousarray(t_start_each, dtype=np.float64))
        assert keep[-1].shape == (n,)
        opts.t_start_each = keep[-1].ctypes.data
    if dt_start_each is not None:
        keep.append(np.ascontiguousarray(dt_start_each, dtype=np.float64))
        assert keep[-1].shape == (n,)
        opts.dt_start_each = keep[-1].ctypes.data
    if event is not None:
        w, c, direction = event
        keep.append(np.ascontiguousarray(w, dtype=np.float64).reshape(-1))
        assert keep[-1].size == dim
        opts.event_w = keep[-1].ctypes.data
        opts.event_c = float(c)
        opts.event_direction = int(direction)
    out = {
        "y_end": np.zeros((dim, n)), "t_end": np.zeros(n), "dt_end": np.zeros(n),
        "status": np.zeros(n, dtype=np.int32), "n_accept": np.zeros(n, dtype=np.uint32),
        "n_reject": np.zeros(n, dtype=np.uint32), "n_rhs": np.zeros(n, dtype=np.uint32),
    }
    if history_capacity > 0:
        out["hist"] = np.zeros((n, history_capacity, 1 + dim))
        out["hist_len"] = np.zeros(n, dtype=np.uint32)
    res = _abi.Result(**{k: v.ctypes.data for k, v in out.items()})
    if history_capacity > 0:  # the two columns of the record array, as views
        out["hist_t"] = out["hist"][:, :, 0]
        out["hist_y"] = out["hist"][:, :, 1:]
    rc = L.oracle_ivp_solve_ensemble_ex(C.byref(cfg), rid, n, y0.ctypes.data, pptr, C.byref(opts), C.byref(res),
                                        pow_mode, n_threads)
    if rc != 0:
        raise RuntimeError(f"oracle_ivp_solve_ensemble rc={rc}")
    return out


def _path_query_args(rhs, y0, params, solved, t_start, shared_params, params_aos):
    """solved: dict with hist, hist_len and optionally t_end + y_end (closing knot), n_accept / status (a path cut
    short by its capacity has no closing knot), t_start (per-trajectory start times of a resumed leg)."""
    rid, dim, npar = rhs_info(rhs)
    y0 = np.ascontiguousarray(y0, dtype=np.float64)
    hist = np.ascontiguousarray(solved["hist"], dtype=np.float64)
    n, cap = hist.shape[0], hist.shape[1]
    assert y0.shape == (dim, n) and hist.shape[2] == 1 + dim
    flags = (_abi.FLAG_SHARED_PARAMS if shared_params else 0) | (_abi.FLAG_PARAMS_AOS if params_aos else 0)
    keep = [y0, hist, np.ascontiguousarray(solved["hist_len"], dtype=np.uint32)]
    res = _abi.Result(hist=hist.ctypes.data, hist_len=keep[2].ctypes.data)
    if solved.get("t_end") is not None and solved.get("y_end") is not None:
        keep += [np.ascontiguousarray(solved["t_end"], dtype=np.float64), np.ascontiguousarray(solved["y_end"], dtype=np.float64)]
        res.t_end, res.y_end = keep[3].ctypes.data, keep[4].ctypes.data
    if solved.get("n_accept") is not None:
        keep.append(np.ascontiguousarray(solved["n_accept"], dtype=np.uint32))
        res.n_accept = keep[-1].ctypes.data
    if solved.get("status") is not None:
        keep.append(np.ascontiguousarray(solved["status"], dtype=np.int32))
        res.status = keep[-1].ctypes.data
    if solved.get("t_start") is not None:
        keep.append(np.ascontiguousarray(solved["t_start"], dtype=np.float64))
        res.t_start = keep[-1].ctypes.data
    pptr = None
    if npar > 0:
        keep.append(np.ascontiguousarray(params, dtype=np.float64))
        pptr = keep[-1].ctypes.data
    cfg = _abi.Config(method=0, dim=dim, n_params=npar, flags=flags, history_capacity=cap, t_start=t_start)
    return cfg, rid, n, dim, y0.ctypes.data, pptr, res, keep


def sample_paths(rhs, y0, params, solved, times, *, t_start, shared_params=False, params_aos=False):
    """CPU statement of bacon_ivp_sample_paths.  solved: dict with hist, hist_len (and t_end, y_end)."""
    cfg, rid, n, dim, yptr, pptr, res, keep = _path_query_args(rhs, y0, params, solved, t_start, shared_params, params_aos)
    times = np.ascontiguousarray(times, dtype=np.float64).reshape(-1)
    out = np.empty((n, times.size, dim))
    rc = lib().oracle_sample_paths(C.byref(cfg), rid, n, yptr, pptr, C.byref(res), times.size, times.ctypes.data,
                                   out.ctypes.data)
    if rc != 0:
        raise RuntimeError(f"oracle_sample_paths rc={rc}")
    return out


def locate_events(rhs, y0, params, solved, w, c=0.0, direction=0, capacity=8, *, t_start, shared_params=False,
                  params_aos=False):
    """CPU statement of bacon_ivp_locate_events."""
    cfg, rid, n, dim, yptr, pptr, res, keep = _path_query_args(rhs, y0, params, solved, t_start, shared_params, params_aos)
    w = np.ascontiguousarray(w, dtype=np.float64).reshape(-1)
    assert w.size == dim
    events = np.zeros((n, capacity, 1 + dim))
    counts = np.zeros(n, dtype=np.uint32)
    rc = lib().oracle_locate_events(C.byref(cfg), rid, n, yptr, pptr, C.byref(res), w.ctypes.data, float(c),
                                    int(direction), int(capacity), events.ctypes.data, counts.ctypes.data)
    if rc != 0:
        raise RuntimeError(f"oracle_locate_events rc={rc}")
    return events, counts


def roots_secant(which, start, h, tol, n_max=1000, central=False):
    """roots::secant (src/roots/mod.rs:289-337) on the reference's test functions."""
    L = lib()
    start = np.ascontiguousarray(start, dtype=np.float64)
    sol = np.zeros(3)
    it = C.c_ulonglong(0)
    rc = L.oracle_roots_secant(which, start.ctypes.data, h, tol, n_max, int(central),
                               sol.ctypes.data, C.byref(it))
    return rc, sol[:len(start)], it.value


def fourth_root(x):
    x = np.ascontiguousarray(x, dtype=np.float64)
    a, b = np.empty_like(x), np.empty_like(x)
    lib().oracle_fourth_root(x.ctypes.data, x.size, a.ctypes.data, b.ctypes.data)
    return a, b


def nth_root(x, n):
    """x^(1/n) by libm pow and by the deterministic Newton root of the strict Adams kernels (n = 3 or 5)."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    a, b = np.empty_like(x), np.empty_like(x)
    lib().oracle_nth_root(x.ctypes.data, x.size, int(n), a.ctypes.data, b.ctypes.data)
    return a, b


def coefficients(method, literal=False):
    """The method's coefficient tables as the oracle's steppers use them (oracle_coefficients in oracle_capi.cpp)."""
    n = lib().oracle_coefficients(int(method), int(bool(literal)), None, 0)
    if n < 0:
        raise ValueError(f"no coefficient tables for method {method}")
    v = np.empty(n, dtype=np.float64)
    lib().oracle_coefficients(int(method), int(bool(literal)), v.ctypes.data, n)
    if method in (_abi.RK45, _abi.RK23):
        o = 6 if method == _abi.RK45 else 4
        return dict(c=v[:o], A=v[o:o + o * o].reshape(o, o), b=v[o + o * o:2 * o + o * o], e=v[2 * o + o * o:3 * o + o * o],
                    safety=float(v[-1]))
    if method in (_abi.BDF6, _abi.BDF2):
        o = n // 2
        return dict(higher=v[:o], lower=v[o:])
    o = (n - 1) // 2
    return dict(predictor=v[:o], corrector=v[o:2 * o], error=float(v[-1]))
