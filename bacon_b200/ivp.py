"""Host-side mirror of `bacon_sci::ivp`'s solver front end over the C ABI.

Names, argument meaning and error behaviour follow the reference's `IVPSolver`
builder trait (src/ivp.rs:134-190) as implemented by `RungeKutta`
(src/ivp/rk.rs:118-343) and `BDF` (src/ivp/bdf.rs:124-332); the README's older
names (README.md:24-40: RK45, with_dt_min, with_dt_max, with_start, with_end,
build, solve_ivp) are kept as aliases.  New: `solve_ivp_ensemble`.

Every setter returns `self` or raises `IVPError` (the reference returns
`Result<Self, IVPError>`).  Validation itself runs inside the C library
(bacon_solver_with_*), so any other binding gets identical behaviour.

The right-hand side is a CUDA device functor registered in the library
("lorenz", "vdp", "robertson", ... or a user RHS); the
per-trajectory parameter block plays the role of the reference's `UserData`.
"""
import ctypes as C
import os
import weakref

import numpy as np

from . import _abi
from ._lib import lib, last_error


_SENTINEL = os.environ.get("BACON_IVP_SENTINEL", "0") not in ("", "0")


class IVPError(Exception):
    """Mirror of `IVPError` (src/ivp.rs:50-76) plus the engine's own call-level codes."""

    def __init__(self, code, message=""):
        self.code = int(code)
        self.variant = _abi.STATUS_NAMES.get(self.code, f"Unknown({code})")
        super().__init__(f"{self.variant}: {message}" if message else self.variant)


def _check(rc):
    if rc != 0:
        raise IVPError(rc, last_error())


_RESULT_FIELDS = {name for name, _ in _abi.Result._fields_}


class PinnedBlock:
    """One page-locked host block from `bacon_host_alloc`, carved into numpy arrays.  The arrays keep the
    block alive (ndarray.base -> ctypes view -> this object); it returns to the library's cache when the
    last of them is dropped."""

    def __init__(self, nbytes):
        self.nbytes = max(int(nbytes), 1)
        self.ptr = lib().bacon_host_alloc(self.nbytes)
        if not self.ptr:
            raise IVPError(_abi.E_CUDA, last_error())
        self._off = 0
        self._fin = weakref.finalize(self, lib().bacon_host_free, self.ptr)

    def take(self, shape, dtype):
        dtype = np.dtype(dtype)
        count = int(np.prod(shape)) if len(shape) else 1
        nb = count * dtype.itemsize
        off = (self._off + 255) // 256 * 256
        if off + nb > self.nbytes:
            raise IVPError(_abi.E_BAD_ARGUMENT, "pinned block too small")
        self._off = off + nb
        if nb == 0:
            return np.zeros(shape, dtype=dtype)
        view = (C.c_char * nb).from_address(self.ptr + off)
        view._owner = self
        return np.frombuffer(view, dtype=dtype).reshape(shape)

    @staticmethod
    def size_for(specs):
        return sum((int(np.prod(sh)) * np.dtype(dt).itemsize + 255) // 256 * 256 + 256 for sh, dt in specs)


def pinned_empty(shape, dtype=np.float64):
    """A numpy array in page-locked memory (uninitialised): inputs built here reach the GPU by async DMA."""
    shape = tuple(np.atleast_1d(shape).tolist()) if not isinstance(shape, tuple) else shape
    blk = PinnedBlock(PinnedBlock.size_for([(shape, dtype)]))
    return blk.take(shape, dtype)


class EnsembleResult:
    """Per-trajectory records of one ensemble solve (layouts of bacon_ivp_result)."""

    def __init__(self, arrays, launch):
        self.y_end = arrays["y_end"]          # (dim, n)
        self.t_end = arrays["t_end"]          # (n,)
        self.dt_end = arrays["dt_end"]
        self.status = arrays["status"]        # bacon_status per trajectory
        self.n_accept = arrays["n_accept"]
        self.n_reject = arrays["n_reject"]
        self.n_rhs = arrays["n_rhs"]
        self.hist = arrays.get("hist")        # (n, cap, 1 + dim): (t, y) records, one Path per trajectory
        self.hist_t = None if self.hist is None else self.hist[:, :, 0]    # views of the record array
        self.hist_y = None if self.hist is None else self.hist[:, :, 1:]
        self.hist_len = arrays.get("hist_len")
        self.launch = launch                  # dict: kernel_ms, h2d_ms, d2h_ms, grid, block, regs_per_thread
        self._query = None                    # (cfg, rhs id, y0, params) of the solve, for sample() / locate_events()

    def path(self, i):
        """The `Path` of trajectory i (src/ivp.rs:203): [(t, y)] of accepted points."""
        m = int(self.hist_len[i])
        return [(float(self.hist_t[i, k]), np.array(self.hist_y[i, k])) for k in range(m)]

    # ---- queries on the stored paths (not in the reference: include/bacon_ivp.h)
    def _solved(self):
        if self.hist is None or self._query is None:
            raise IVPError(_abi.E_BAD_ARGUMENT, "path queries need a dense-output solve (with_history(capacity))")
        cfg, rid, y0, params = self._query[:4]
        res = _abi.Result(hist=self.hist.ctypes.data, hist_len=self.hist_len.ctypes.data,
                          t_end=self.t_end.ctypes.data, y_end=self.y_end.ctypes.data,
                          n_accept=self.n_accept.ctypes.data, status=self.status.ctypes.data)
        if len(self._query) > 4 and self._query[4] is not None:  # a resumed leg: per-trajectory start times
            res.t_start = self._query[4].ctypes.data
        return cfg, rid, y0, (None if params is None else params.ctypes.data), res

    def restart_record(self):
        """(y_end, t_end, dt_end): what `solve_ivp_ensemble(..., restart=...)` takes to go on where this solve stopped —
        the C-ABI form of the reference's in-memory resumable iterator (src/ivp.rs:220-238)."""
        return self.y_end, self.t_end, self.dt_end

    def sample(self, times):
        """The state of every trajectory at `times`: (n, len(times), dim); NaN where a time lies outside a
        trajectory's path.  Cubic Hermite between accepted points with the right-hand side's slopes."""
        cfg, rid, y0, pptr, res = self._solved()
        times = np.ascontiguousarray(times, dtype=np.float64).reshape(-1)
        n = y0.shape[1]
        out = np.empty((n, times.size, cfg.dim))
        _check(lib().bacon_ivp_sample_paths(C.byref(cfg), rid, n, y0.ctypes.data, pptr, C.byref(res), times.size,
                                            times.ctypes.data, out.ctypes.data))
        return out

    def locate_events(self, w, c=0.0, direction=0, capacity=8):
        """Zeros of g(y) = w . y - c along every path, in order.  Returns (events, n_events): events is
        (n, capacity, 1 + dim) records (t*, y(t*)), n_events the number found per trajectory (may exceed capacity).
        direction +1: rising only, -1: falling only, 0: both."""
        cfg, rid, y0, pptr, res = self._solved()
        w = np.ascontiguousarray(w, dtype=np.float64).reshape(-1)
        if w.size != cfg.dim:
            raise IVPError(_abi.E_BAD_ARGUMENT, f"w must have {cfg.dim} entries")
        n = y0.shape[1]
        events = np.zeros((n, int(capacity), 1 + cfg.dim))
        counts = np.zeros(n, dtype=np.uint32)
        _check(lib().bacon_ivp_locate_events(C.byref(cfg), rid, n, y0.ctypes.data, pptr, C.byref(res), w.ctypes.data,
                                             float(c), int(direction), int(capacity), events.ctypes.data,
                                             counts.ctypes.data))
        return events, counts


class _Solver:
    METHOD = None
    _ORDER = None

    def __init__(self, dim, *, dyn=False, dim_type=None):
        """`dim_type` is the reference's type parameter D: a static dimension C >= 1 (`Const<C>`) or `DYN` (`Dyn`).
        new(dim) builds Const<dim>; new_dyn(size) builds Dyn with `size`; the mixed-up calls fail like the
        reference's `Dimension` (src/lib.rs:53-76): see new / new_dyn."""
        h = C.c_void_p()
        if dyn:
            _check(lib().bacon_solver_new_dyn(self.METHOD, _abi.DIM_DYN if dim_type is None else int(dim_type), int(dim),
                                              C.byref(h)))
        else:
            _check(lib().bacon_solver_new_static(self.METHOD, int(dim) if dim_type is None else int(dim_type), C.byref(h)))
        self._h = h.value
        self._dim = int(dim) if (dyn or dim_type is None) else int(dim_type)
        self._y0 = None
        self._rhs = None
        self._event = None

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                lib().bacon_solver_free(h)
            except Exception:
                pass

    # ---- IVPSolver::new / new_dyn (ivp.rs:159-163) with the reference's Dimension check (lib.rs:53-76).  Python has
    # no type parameters, so the solver's `D` is an argument: new(3) is `RK45::<U3>::new()`, new_dyn(3) is
    # `RK45::<Dyn>::new_dyn(3)`; new(dim_type=DYN) raises StaticOnDynamic and new_dyn(3, dim_type=3) DynamicOnStatic.
    DYN = _abi.DIM_DYN

    @classmethod
    def new(cls, dim=1, *, dim_type=None):
        return cls(dim, dim_type=dim_type)

    @classmethod
    def new_dyn(cls, size, *, dim_type=None):
        return cls(size, dyn=True, dim_type=dim_type)

    def dim(self):
        return self._dim

    # ---- setters (rk.rs:168-247)
    def with_tolerance(self, tol):
        _check(lib().bacon_solver_with_tolerance(self._h, float(tol)))
        return self

    def with_maximum_dt(self, max_dt):
        _check(lib().bacon_solver_with_maximum_dt(self._h, float(max_dt)))
        return self

    def with_minimum_dt(self, min_dt):
        _check(lib().bacon_solver_with_minimum_dt(self._h, float(min_dt)))
        return self

    def with_initial_time(self, initial):
        _check(lib().bacon_solver_with_initial_time(self._h, float(initial)))
        return self

    def with_ending_time(self, ending):
        _check(lib().bacon_solver_with_ending_time(self._h, float(ending)))
        return self

    def with_initial_conditions_slice(self, start):
        start = np.asarray(start, dtype=np.float64)
        if start.shape != (self._dim,):
            # nalgebra panics on a length mismatch (ivp.rs:177-180); here it is an error value
            raise IVPError(_abi.E_BAD_ARGUMENT, f"initial conditions have shape {start.shape}, dim is {self._dim}")
        self._y0 = start.copy()
        return self

    def with_initial_conditions(self, start):
        return self.with_initial_conditions_slice(start)

    def with_derivative(self, rhs_name):
        rid = lib().bacon_rhs_lookup(str(rhs_name).encode())
        if rid < 0:
            raise IVPError(_abi.E_BAD_ARGUMENT, f"no right-hand side named {rhs_name!r} is registered")
        self._rhs = (rid, str(rhs_name))
        return self

    # ---- README aliases (README.md:33-39)
    with_dt_max = with_maximum_dt
    with_dt_min = with_minimum_dt
    with_start = with_initial_time
    with_end = with_ending_time

    def build(self):
        return self

    # ---- engine knobs (not in the reference)
    def with_semantics(self, semantics):
        _check(lib().bacon_solver_with_semantics(self._h, int(semantics)))
        return self

    def with_flags(self, flags):
        _check(lib().bacon_solver_with_flags(self._h, int(flags)))
        return self

    def with_history(self, capacity):
        _check(lib().bacon_solver_with_history(self._h, int(capacity)))
        return self

    def with_max_attempts(self, cap):
        _check(lib().bacon_solver_with_max_attempts(self._h, int(cap)))
        return self

    def with_initial_dt(self, dt):
        """First step size instead of the reference's (dt_max + dt_min)/2 (rk.rs:315); clamped into [dt_min, dt_max]."""
        _check(lib().bacon_solver_with_initial_dt(self._h, float(dt)))
        return self

    def with_terminal_event(self, w, c=0.0, direction=0):
        """Stop every trajectory at the first zero of g(y) = w . y - c (direction +1 rising, -1 falling, 0 both): status
        STOPPED_AT_EVENT, t_end / y_end = the event point.  `None` removes it.  Not in the reference."""
        if w is None:
            self._event = None
            return self
        w = np.ascontiguousarray(w, dtype=np.float64).reshape(-1)
        if w.size != self._dim:
            raise IVPError(_abi.E_BAD_ARGUMENT, f"w must have {self._dim} entries")
        if int(direction) not in (-1, 0, 1):
            raise IVPError(_abi.E_BAD_ARGUMENT, "direction must be -1, 0 or +1")
        self._event = (w, float(c), int(direction))
        return self

    def _options(self, n, restart, keep):
        """bacon_ivp_options for this call (host arrays) or None.  restart = (t_start_each, dt_start_each), either may
        be None."""
        if self._event is None and restart is None:
            return None
        o = _abi.Options()
        if self._event is not None:
            o.event_w = self._event[0].ctypes.data
            o.event_c = self._event[1]
            o.event_direction = self._event[2]
        if restart is not None:
            t0, dt0 = restart
            for name, arr in (("t_start_each", t0), ("dt_start_each", dt0)):
                if arr is None:
                    continue
                arr = np.ascontiguousarray(arr, dtype=np.float64)
                if arr.shape != (n,):
                    raise IVPError(_abi.E_BAD_ARGUMENT, f"restart {name} must have shape ({n},)")
                keep.append(arr)
                setattr(o, name, arr.ctypes.data)
        return o

    # ---- config
    def _config(self, n_params, extra_flags=0):
        cfg = _abi.Config()
        _check(lib().bacon_solver_config(self._h, C.byref(cfg)))
        cfg.n_params = n_params
        cfg.flags |= extra_flags
        return cfg

    def _rhs_info(self):
        if self._rhs is None:
            raise IVPError(_abi.E_MISSING_PARAMETERS, "with_derivative was not called")
        rid = self._rhs[0]
        d, p = C.c_int(), C.c_int()
        _check(lib().bacon_rhs_info(rid, None, C.byref(d), C.byref(p)))
        if d.value != self._dim:
            raise IVPError(_abi.E_BAD_ARGUMENT, f"rhs {self._rhs[1]!r} has dimension {d.value}, solver has {self._dim}")
        return rid, d.value, p.value

    # ---- solve(data) + collect_vec (rk.rs:249-343, ivp.rs:209-211): one trajectory
    def solve(self, data=None, capacity=1 << 16):
        """Integrate the single trajectory set by with_initial_conditions and return its
        `Path`: a list of (t, y) accepted points.  A per-trajectory failure raises
        IVPError after the points yielded before it (the reference yields Err once,
        ivp.rs:232-235); the partial path is attached as `.path`."""
        if self._y0 is None:
            raise IVPError(_abi.E_MISSING_PARAMETERS, "with_initial_conditions was not called")
        params = None if data is None else np.asarray(data, dtype=np.float64).reshape(-1, 1)
        # collect_vec grows its Vec (ivp.rs:209-211); here the path's storage is sized up front, so a path longer
        # than `capacity` is integrated once more with the capacity it reported (n_accept)
        for _ in range(2):
            self.with_history(capacity)
            try:
                res = self.solve_ivp_ensemble(self._y0.reshape(self._dim, 1), params)
            finally:
                self.with_history(0)
            if int(res.n_accept[0]) <= capacity:
                break
            capacity = int(res.n_accept[0])
        path = res.path(0)
        st = int(res.status[0])
        if st not in (_abi.OK, _abi.STOPPED_AT_EVENT):
            err = IVPError(st, "trajectory 0")
            err.path = path
            err.result = res
            raise err
        return path

    def solve_ivp(self, rhs_name, data=None, **kw):  # README.md:40
        return self.with_derivative(rhs_name).solve(data, **kw)

    # ---- the new entry point: N initial conditions x N parameter sets
    def solve_ivp_ensemble(self, y0, params=None, *, n_gpus=1, shared_params=False, params_aos=False, rhs=None,
                           zero_copy=True, restart=None):
        """y0: (dim, n) float64 host array; params: (n_params, n), or (n, ...) = one contiguous block
        per trajectory with params_aos=True, or (n_params,) with shared_params=True.
        Host buffers in, host buffers out (H2D, kernel, D2H).  The result arrays live in page-locked memory
        from the library (`bacon_host_alloc`), so the D2H leg is an asynchronous DMA; inputs made with
        `pinned_empty` get the same treatment.  zero_copy (default; applies on 1 GPU, final state only, when
        every buffer is pinned and the per-trajectory input is small): the kernel reads and writes the host
        buffers itself, no staging copies; otherwise the staged path runs.
        restart = (t_start_each, dt_start_each): per-trajectory start time and first dt, e.g. the (t_end, dt_end) a
        previous solve returned together with its y_end as y0 (EnsembleResult.restart_record)."""
        if rhs is not None:
            self.with_derivative(rhs)
        rid, dim, npar = self._rhs_info()
        y0 = np.ascontiguousarray(y0, dtype=np.float64)
        if y0.ndim != 2 or y0.shape[0] != dim:
            raise IVPError(_abi.E_BAD_ARGUMENT, f"y0 must have shape ({dim}, n), got {y0.shape}")
        n = y0.shape[1]
        flags = 0
        pptr = None
        if npar > 0:
            if params is None:
                raise IVPError(_abi.E_MISSING_PARAMETERS, f"rhs needs {npar} parameter(s) per trajectory")
            params = np.ascontiguousarray(params, dtype=np.float64)
            if shared_params:
                if params.shape != (npar,):
                    raise IVPError(_abi.E_BAD_ARGUMENT, f"shared params must have shape ({npar},)")
                flags |= _abi.FLAG_SHARED_PARAMS
            elif params_aos:
                if params.size != npar * n or params.shape[0] != n:
                    raise IVPError(_abi.E_BAD_ARGUMENT, f"AoS params must have shape ({n}, {npar}), got {params.shape}")
                flags |= _abi.FLAG_PARAMS_AOS
            elif params.shape != (npar, n):
                raise IVPError(_abi.E_BAD_ARGUMENT, f"params must have shape ({npar}, {n}), got {params.shape}")
            pptr = params.ctypes.data
        if zero_copy:
            flags |= _abi.FLAG_ZERO_COPY
        cfg = self._config(npar, flags)
        cap = cfg.history_capacity
        L = lib()
        _check(L.bacon_ivp_validate(C.byref(cfg)))  # argument errors before anything touches the GPU
        specs = {"y_end": ((dim, n), np.float64), "t_end": ((n,), np.float64), "dt_end": ((n,), np.float64),
                 "status": ((n,), np.int32), "n_accept": ((n,), np.uint32), "n_reject": ((n,), np.uint32),
                 "n_rhs": ((n,), np.uint32)}
        if cap > 0:
            specs.update({"hist": ((n, cap, 1 + dim), np.float64), "hist_len": ((n,), np.uint32)})
        if n == 0:
            arrays = {k: np.zeros(sh, dtype=dt) for k, (sh, dt) in specs.items()}
        else:
            block = PinnedBlock(PinnedBlock.size_for(specs.values()))
            arrays = {k: block.take(sh, dt) for k, (sh, dt) in specs.items()}
            # Every trajectory is handed out exactly once and stores its status when it retires, so the result arrays
            # need no initialisation; BACON_IVP_SENTINEL=1 (set by tests/conftest.py) pre-fills status with -1 so that
            # a trajectory the kernels lost would show (0.5 ms of CPU time per 2^20 trajectories: 1.5 % of a launch).
            if _SENTINEL:
                arrays["status"].fill(-1)
            if cap > 0:
                arrays["hist_len"].fill(0)
        res = _abi.Result(**{k: v.ctypes.data for k, v in arrays.items()})
        keep = []
        opts = self._options(n, restart, keep)
        if opts is not None:
            _check(L.bacon_ivp_solve_ensemble_ex(C.byref(cfg), rid, n, y0.ctypes.data, pptr, C.byref(opts), C.byref(res),
                                                 int(n_gpus)))
        elif n_gpus == 1:
            _check(L.bacon_ivp_solve_ensemble(C.byref(cfg), rid, n, y0.ctypes.data, pptr, C.byref(res)))
        else:
            _check(L.bacon_ivp_solve_ensemble_multi(C.byref(cfg), rid, n, y0.ctypes.data, pptr, C.byref(res), int(n_gpus)))
        out = EnsembleResult(arrays, last_launch())
        t0_each = None if restart is None or restart[0] is None else np.ascontiguousarray(restart[0], dtype=np.float64)
        out._query = (cfg, rid, y0, params if npar > 0 else None, t0_each)
        return out

    # ---- queries on paths resident in HBM (torch CUDA tensors; `out` = what solve_ivp_ensemble_device returned,
    # the solver still configured as for that solve)
    def _device_query(self, y0, params, out, shared_params, params_aos):
        rid, dim, npar = self._rhs_info()
        flags = (_abi.FLAG_SHARED_PARAMS if shared_params else 0) | (_abi.FLAG_PARAMS_AOS if params_aos else 0)
        cfg = self._config(npar, flags)
        if cfg.history_capacity <= 0 or "hist" not in out:
            raise IVPError(_abi.E_BAD_ARGUMENT, "path queries need a dense-output solve (with_history(capacity))")
        res = _abi.Result(hist=out["hist"].data_ptr(), hist_len=out["hist_len"].data_ptr(),
                          t_end=out["t_end"].data_ptr(), y_end=out["y_end"].data_ptr(),
                          n_accept=out["n_accept"].data_ptr(), status=out["status"].data_ptr())
        if out.get("t_start") is not None:  # a resumed leg
            res.t_start = out["t_start"].data_ptr()
        return cfg, rid, dim, (params.data_ptr() if npar > 0 else None), res

    def sample_paths_device(self, y0, params, out, times, *, shared_params=False, params_aos=False, samples=None,
                            stream=None):
        import torch
        cfg, rid, dim, pptr, res = self._device_query(y0, params, out, shared_params, params_aos)
        n, m = y0.shape[1], times.numel()
        if samples is None:
            samples = torch.empty((n, m, dim), dtype=torch.float64, device=y0.device)
        with torch.cuda.device(y0.device):
            s = torch.cuda.current_stream(y0.device) if stream is None else stream
            _check(lib().bacon_ivp_sample_paths_device(C.byref(cfg), rid, n, y0.data_ptr(), pptr, C.byref(res), m,
                                                       times.data_ptr(), samples.data_ptr(), C.c_void_p(s.cuda_stream)))
        return samples

    def locate_events_device(self, y0, params, out, w, c=0.0, direction=0, capacity=8, *, shared_params=False,
                             params_aos=False, stream=None):
        import torch
        cfg, rid, dim, pptr, res = self._device_query(y0, params, out, shared_params, params_aos)
        n = y0.shape[1]
        w = np.ascontiguousarray(w, dtype=np.float64).reshape(-1)
        if w.size != dim:
            raise IVPError(_abi.E_BAD_ARGUMENT, f"w must have {dim} entries")
        events = torch.zeros((n, int(capacity), 1 + dim), dtype=torch.float64, device=y0.device)
        counts = torch.zeros(n, dtype=torch.int32, device=y0.device)
        with torch.cuda.device(y0.device):
            s = torch.cuda.current_stream(y0.device) if stream is None else stream
            _check(lib().bacon_ivp_locate_events_device(C.byref(cfg), rid, n, y0.data_ptr(), pptr, C.byref(res),
                                                        w.ctypes.data, float(c), int(direction), int(capacity),
                                                        events.data_ptr(), counts.data_ptr(), C.c_void_p(s.cuda_stream)))
        return events, counts

    # ---- device-resident variant: torch CUDA tensors in, torch CUDA tensors out, no copies
    def solve_ivp_ensemble_device(self, y0, params=None, *, shared_params=False, params_aos=False, out=None,
                                  stream=None, rhs=None, restart=None):
        import torch
        if rhs is not None:
            self.with_derivative(rhs)
        rid, dim, npar = self._rhs_info()
        if not (y0.is_cuda and y0.dtype == torch.float64 and y0.is_contiguous() and y0.shape[0] == dim):
            raise IVPError(_abi.E_BAD_ARGUMENT, f"y0 must be a contiguous CUDA float64 tensor of shape ({dim}, n)")
        n = y0.shape[1]
        flags = 0
        pptr = None
        if npar > 0:
            if params is None:
                raise IVPError(_abi.E_MISSING_PARAMETERS, f"rhs needs {npar} parameter(s) per trajectory")
            want = (npar,) if shared_params else ((n, npar) if params_aos else (npar, n))
            if not (params.is_cuda and params.dtype == torch.float64 and params.is_contiguous()
                    and params.numel() == want[0] * (want[1] if len(want) > 1 else 1) and params.shape[0] == want[0]):
                raise IVPError(_abi.E_BAD_ARGUMENT, f"params must be a contiguous CUDA float64 tensor of shape {want}")
            if shared_params:
                flags |= _abi.FLAG_SHARED_PARAMS
            elif params_aos:
                flags |= _abi.FLAG_PARAMS_AOS
            pptr = params.data_ptr()
        cfg = self._config(npar, flags)
        cap = cfg.history_capacity
        dev = y0.device
        if out is None:
            out = {
                "y_end": torch.empty((dim, n), dtype=torch.float64, device=dev),
                "t_end": torch.empty(n, dtype=torch.float64, device=dev),
                "dt_end": torch.empty(n, dtype=torch.float64, device=dev),
                "status": torch.full((n,), -1, dtype=torch.int32, device=dev),
                "n_accept": torch.zeros(n, dtype=torch.int32, device=dev),
                "n_reject": torch.zeros(n, dtype=torch.int32, device=dev),
                "n_rhs": torch.zeros(n, dtype=torch.int32, device=dev),
            }
            if cap > 0:
                out["hist"] = torch.zeros((n, cap, 1 + dim), dtype=torch.float64, device=dev)
                out["hist_len"] = torch.zeros(n, dtype=torch.int32, device=dev)
        res = _abi.Result(**{k: v.data_ptr() for k, v in out.items() if k in _RESULT_FIELDS})
        if cap > 0:  # views of the record array
            out["hist_t"] = out["hist"][:, :, 0]
            out["hist_y"] = out["hist"][:, :, 1:]
        opts = None
        if self._event is not None or restart is not None:  # restart = (t_start_each, dt_start_each) CUDA tensors
            opts = self._options(n, None, [])
            if opts is None:
                opts = _abi.Options()
            if restart is not None:
                for name, ten in (("t_start_each", restart[0]), ("dt_start_each", restart[1])):
                    if ten is None:
                        continue
                    if not (ten.is_cuda and ten.dtype == torch.float64 and ten.is_contiguous() and ten.numel() == n):
                        raise IVPError(_abi.E_BAD_ARGUMENT, f"restart {name} must be a contiguous CUDA float64 tensor of {n}")
                    setattr(opts, name, ten.data_ptr())
                if restart[0] is not None:
                    out["t_start"] = restart[0]
        with torch.cuda.device(dev):
            s = torch.cuda.current_stream(dev) if stream is None else stream
            if opts is not None:
                _check(lib().bacon_ivp_solve_ensemble_device_ex(C.byref(cfg), rid, n, y0.data_ptr(), pptr, C.byref(opts),
                                                                C.byref(res), C.c_void_p(s.cuda_stream)))
            else:
                _check(lib().bacon_ivp_solve_ensemble_device(C.byref(cfg), rid, n, y0.data_ptr(), pptr, C.byref(res),
                                                             C.c_void_p(s.cuda_stream)))
        return out


def register_rhs_source(name, type_name, source, dim, n_params=0):
    """A user right-hand side as CUDA C++ source text (a functor `type_name` with DIM, NPARAM and operator(), optionally
    jac / scaled: include/bacon_ivp_rhs.cuh), compiled by the library with NVRTC and inlined into the kernels like a
    built-in.  The device form of `with_derivative(closure)` (src/ivp.rs:186) for callers without nvcc.  Returns the
    rhs id; afterwards `with_derivative(name)` works.  A source that does not compile raises IVPError(UserError) with
    the compiler log."""
    rid = lib().bacon_rhs_register_source(str(name).encode(), str(type_name).encode(), str(source).encode(), int(dim),
                                          int(n_params))
    if rid < 0:
        _check(-rid)
    return rid


def last_launch():
    info = _abi.LaunchInfo()
    _check(lib().bacon_ivp_last_launch(C.byref(info)))
    return {k: getattr(info, k) for k, _ in _abi.LaunchInfo._fields_}


def fp64_peak_tflops(iters=1 << 15):
    return float(lib().bacon_fp64_peak_tflops(int(iters), None))


class RungeKutta45(_Solver):
    """Runge-Kutta-Fehlberg 4(5) (src/ivp/rk.rs:561)."""
    METHOD = _abi.RK45


class RungeKutta23(_Solver):
    """Bogacki-Shampine 3(2), "the second adaptive RK" (src/ivp/rk.rs:656)."""
    METHOD = _abi.RK23


class BDF6(_Solver):
    """Backwards differentiation formula, order 6 with order-5 error estimate (src/ivp/bdf.rs:706)."""
    METHOD = _abi.BDF6


class BDF2(_Solver):
    """Backwards differentiation formula, order 2 (src/ivp/bdf.rs:762)."""
    METHOD = _abi.BDF2


class Adams5(_Solver):
    """Adams-Bashforth-Moulton predictor-corrector, order 5 (src/ivp/adams.rs:633)."""
    METHOD = _abi.ADAMS5


class Adams3(_Solver):
    """Adams-Bashforth-Moulton predictor-corrector, order 3 (src/ivp/adams.rs:693)."""
    METHOD = _abi.ADAMS3


class Euler(_Solver):
    """Explicit Euler, fixed step (src/ivp.rs:269-485).  `with_tolerance` is a no-op; each of
    `with_maximum_dt` / `with_minimum_dt` sets the step, or averages it with the one already set
    (ivp.rs:389-421).  The path starts with the initial condition and never holds the final state
    (every step yields the OLD point, ivp.rs:331-337); `y_end` is the final state."""
    METHOD = _abi.EULER


RK45 = RungeKutta45  # README.md:24
RK23 = RungeKutta23


def solve_ivp(rhs, y0, params=None, *, t_span, dt_min, dt_max, tolerance, shared_params=False, params_aos=False,
              n_gpus=1, chain=(Adams5, RungeKutta45, BDF6)):
    """The README's free function `ivp::solve_ivp` (README.md:45-47: "tries a fifth-order predictor-corrector
    followed by the Runge-Kutta-Fehlberg method followed by BDF6"), per trajectory of an ensemble: every
    trajectory a solver failed on (status != Ok) is handed to the next solver of the chain, restarted from its
    initial condition.  Returns the merged EnsembleResult with `.method` = index into `chain` of the solver
    that produced each trajectory (the last one tried if all failed)."""
    y0 = np.ascontiguousarray(y0, dtype=np.float64)
    dim, n = y0.shape
    todo = np.arange(n)
    merged = None
    method = np.zeros(n, dtype=np.int32)
    for k, cls in enumerate(chain):
        if todo.size == 0:
            break
        s = (cls.new(dim).with_minimum_dt(dt_min).with_maximum_dt(dt_max).with_tolerance(tolerance)
             .with_initial_time(t_span[0]).with_ending_time(t_span[1]).with_derivative(rhs))
        sub_p = params
        if params is not None and not shared_params:
            pa = np.asarray(params, dtype=np.float64)
            sub_p = pa[todo] if params_aos else pa[:, todo]
        res = s.solve_ivp_ensemble(np.ascontiguousarray(y0[:, todo]), sub_p, shared_params=shared_params,
                                   params_aos=params_aos, n_gpus=n_gpus)
        if merged is None:
            merged = res
        else:
            merged.y_end[:, todo] = res.y_end
            for name in ("t_end", "dt_end", "status", "n_accept", "n_reject", "n_rhs"):
                getattr(merged, name)[todo] = getattr(res, name)
        method[todo] = k
        todo = todo[res.status != _abi.OK]
    if merged is None:  # n == 0
        merged = chain[0].new(dim).with_minimum_dt(dt_min).with_maximum_dt(dt_max).with_tolerance(tolerance) \
            .with_initial_time(t_span[0]).with_ending_time(t_span[1]).with_derivative(rhs) \
            .solve_ivp_ensemble(y0, params, shared_params=shared_params, params_aos=params_aos)
    merged.method = method
    return merged
