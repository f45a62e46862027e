//! `bacon_sci::ivp`'s solver front end (src/ivp.rs:134-190, src/ivp/rk.rs:118-343, src/ivp/bdf.rs:124-332 of
//! aftix/bacon 0.16.2) over the B200 ensemble engine.  UNVERIFIED: no Rust toolchain in the build image; the C++ and
//! Python mirrors of this file are the tested ones.
//!
//! ```ignore
//! use bacon_ivp::RK45;
//! let solver = RK45::new(1)?.with_dt_min(0.01)?.with_dt_max(0.1)?.with_tolerance(1e-4)?
//!     .with_initial_conditions(&[1.0])?.with_start(0.0)?.with_end(10.0)?.build();
//! let path = solver.solve_ivp("exp", &[])?;                       // README.md:32-40
//! let ens = RK45::new(3)?.with_dt_min(1e-9)?.with_dt_max(0.1)?.with_tolerance(1e-8)?
//!     .with_start(0.0)?.with_end(5.0)?.with_derivative("lorenz")?
//!     .solve_ivp_ensemble(&y0 /* [3][n] */, &params /* [3][n] */, false)?;
//! ```
use bacon_ivp_sys as sys;
use std::ffi::{CStr, CString};

/// Mirror of `IVPError` (src/ivp.rs:50-76); discriminants are `bacon_status`.
#[derive(thiserror::Error, Debug, Clone, Copy, PartialEq, Eq)]
#[repr(i32)]
pub enum IVPError {
    #[error("the solver does not have all required parameters set")] MissingParameters = 1,
    #[error("user error in the derivative")] UserError = 2,
    #[error("the given tolerance was out of bounds")] ToleranceOOB = 3,
    #[error("the given time delta was out of bounds")] TimeDeltaOOB = 4,
    #[error("the given ending time was out of bounds")] TimeEndOOB = 5,
    #[error("the given starting time was out of bounds")] TimeStartOOB = 6,
    #[error("a conversion from a necessary primitive failed")] FromPrimitiveFailure = 7,
    #[error("the time step fell below the parameter minimum allowed value")] MinimumTimeDeltaExceeded = 8,
    #[error("the number of iterations exceeded the maximum allowable")] MaximumIterationsExceeded = 9,
    #[error("a matrix was unable to be inverted")] SingularMatrix = 10,
    #[error("attempted to build a dynamic solver with static dimension")] DynamicOnStatic = 11,
    #[error("attempted to build a static solver with dynamic dimension")] StaticOnDynamic = 12,
    #[error("non-finite error estimate")] NonFinite = 13,
    #[error("attempt cap reached")] MaxAttempts = 14,
    #[error("more accepted points than the history capacity")] HistoryOverflow = 15,
    #[error("CUDA error")] Cuda = 16,
    #[error("bad argument")] BadArgument = 17,
    #[error("combination not built")] Unsupported = 18,
}

/// Per-trajectory status of a trajectory that stopped at a terminal event (not an error).
pub const STOPPED_AT_EVENT: i32 = sys::BACON_STOPPED_AT_EVENT;

fn check(rc: i32) -> Result<(), IVPError> {
    if rc == 0 { Ok(()) } else if (1..=18).contains(&rc) { Err(unsafe { std::mem::transmute::<i32, IVPError>(rc) }) } else { Err(IVPError::Cuda) }
}

pub fn last_error() -> String {
    unsafe { CStr::from_ptr(sys::bacon_last_error()).to_string_lossy().into_owned() }
}

/// One trajectory's accepted points (src/ivp.rs:203).
pub type Path = Vec<(f64, Vec<f64>)>;

pub struct EnsembleResult {
    pub n: usize,
    pub dim: usize,
    pub y_end: Vec<f64>,   // [dim][n]
    pub t_end: Vec<f64>,
    pub dt_end: Vec<f64>,
    pub status: Vec<i32>,
    pub n_accept: Vec<u32>,
    pub n_reject: Vec<u32>,
    pub n_rhs: Vec<u32>,
    pub hist: Vec<f64>,    // [n][cap][1 + dim]: (t, y) records, the image of Vec<(f64, SVector<f64, D>)>
    pub hist_len: Vec<u32>,
    pub capacity: usize,
    // what the path queries need of the solve (kept when capacity > 0)
    cfg: sys::bacon_ivp_config,
    rhs: i32,
    y0: Vec<f64>,
    params: Vec<f64>,
    t_start: Vec<f64>,     // per-trajectory start times of a resumed leg (empty otherwise)
}

/// The restart record of a solve: per-trajectory start times and first step sizes (`t_end`, `dt_end` of the previous
/// leg, whose `y_end` is the next leg's `y0`) — the C-ABI form of the reference's resumable iterator (ivp.rs:220-238).
#[derive(Clone, Copy, Default)]
pub struct Restart<'a> {
    pub t_start_each: Option<&'a [f64]>,
    pub dt_start_each: Option<&'a [f64]>,
}

impl EnsembleResult {
    pub fn path(&self, i: usize) -> Path {
        (0..self.hist_len[i] as usize)
            .map(|k| {
                let rec = (i * self.capacity + k) * (1 + self.dim);
                (self.hist[rec], self.hist[rec + 1..rec + 1 + self.dim].to_vec())
            })
            .collect()
    }

    fn solved(&self) -> sys::bacon_ivp_result {
        sys::bacon_ivp_result {
            y_end: self.y_end.as_ptr() as *mut f64, t_end: self.t_end.as_ptr() as *mut f64, dt_end: std::ptr::null_mut(),
            // (a path cut short by its capacity has no closing knot: the queries read n_accept / status)
            status: self.status.as_ptr() as *mut i32, n_accept: self.n_accept.as_ptr() as *mut u32, n_reject: std::ptr::null_mut(),
            n_rhs: std::ptr::null_mut(), hist: self.hist.as_ptr() as *mut f64, hist_len: self.hist_len.as_ptr() as *mut u32,
            t_start: if self.t_start.is_empty() { std::ptr::null() } else { self.t_start.as_ptr() },
        }
    }

    /// The state of every trajectory at `times` ([n][times.len()][dim], NaN outside a trajectory's path): the cubic
    /// Hermite interpolant between accepted points (`bacon_ivp_sample_paths`; not in bacon 0.16.2, whose `Path` is the
    /// accepted points only, src/ivp.rs:203-211).  Needs a solve `with_history(capacity)`.
    pub fn sample(&self, times: &[f64]) -> Result<Vec<f64>, IVPError> {
        let mut out = vec![0.0; self.n * times.len() * self.dim];
        let o = self.solved();
        let pptr = if self.params.is_empty() { std::ptr::null() } else { self.params.as_ptr() };
        check(unsafe { sys::bacon_ivp_sample_paths(&self.cfg, self.rhs, self.n, self.y0.as_ptr(), pptr, &o, times.len(),
                                                   times.as_ptr(), out.as_mut_ptr()) })?;
        Ok(out)
    }

    /// Zeros of `w . y - c` along every path, in order (`bacon_ivp_locate_events`): ([n][capacity][1 + dim] records
    /// (t*, y(t*)), [n] counts).  direction +1: rising only, -1: falling only, 0: both.
    pub fn locate_events(&self, w: &[f64], c: f64, direction: i32, capacity: usize) -> Result<(Vec<f64>, Vec<u32>), IVPError> {
        if w.len() != self.dim { return Err(IVPError::BadArgument); }
        let mut ev = vec![0.0; self.n * capacity * (1 + self.dim)];
        let mut cnt = vec![0u32; self.n];
        let o = self.solved();
        let pptr = if self.params.is_empty() { std::ptr::null() } else { self.params.as_ptr() };
        check(unsafe { sys::bacon_ivp_locate_events(&self.cfg, self.rhs, self.n, self.y0.as_ptr(), pptr, &o, w.as_ptr(), c,
                                                    direction, capacity as i32, ev.as_mut_ptr(), cnt.as_mut_ptr()) })?;
        Ok((ev, cnt))
    }
}

pub struct Solver<const METHOD: i32> {
    h: *mut sys::bacon_solver,
    dim: usize,
    rhs: Option<i32>,
    y0: Option<Vec<f64>>,
    event: Option<(Vec<f64>, f64, i32)>,
    n_gpus: i32,
}

impl<const METHOD: i32> Drop for Solver<METHOD> {
    fn drop(&mut self) { unsafe { sys::bacon_solver_free(self.h) } }
}

macro_rules! setter {
    ($name:ident, $ffi:ident) => {
        pub fn $name(self, v: f64) -> Result<Self, IVPError> { check(unsafe { sys::$ffi(self.h, v) })?; Ok(self) }
    };
}

impl<const METHOD: i32> Solver<METHOD> {
    /// `IVPSolver::new` for a `Const<C>` solver (src/ivp.rs:159, src/lib.rs:59-61): `dim` is the static dimension C.
    /// The dimension is checked against the RHS at solve time.
    pub fn new(dim: usize) -> Result<Self, IVPError> { Self::new_static(dim as i32) }
    /// `IVPSolver::new_dyn(size)` for a `Dyn` solver (src/ivp.rs:163, src/lib.rs:73-75).
    pub fn new_dyn(size: usize) -> Result<Self, IVPError> { Self::new_dyn_typed(sys::BACON_DIM_DYN, size) }
    /// The two constructors with the solver's type parameter spelled out (`dim_type` = C >= 1 for `Const<C>`,
    /// `BACON_DIM_DYN` for `Dyn`): `new()` on `Dyn` is `StaticOnDynamic`, `new_dyn` on `Const<C>` `DynamicOnStatic`.
    pub fn new_static(dim_type: i32) -> Result<Self, IVPError> {
        let mut h = std::ptr::null_mut();
        check(unsafe { sys::bacon_solver_new_static(METHOD, dim_type, &mut h) })?;
        Ok(Self { h, dim: dim_type as usize, rhs: None, y0: None, event: None, n_gpus: 1 })
    }
    pub fn new_dyn_typed(dim_type: i32, size: usize) -> Result<Self, IVPError> {
        let mut h = std::ptr::null_mut();
        check(unsafe { sys::bacon_solver_new_dyn(METHOD, dim_type, size as i32, &mut h) })?;
        Ok(Self { h, dim: size, rhs: None, y0: None, event: None, n_gpus: 1 })
    }
    pub fn dim(&self) -> usize { self.dim }
    /// Shard every ensemble solve over the first `n` GPUs of the box (trajectory i -> GPU i mod n).
    pub fn n_gpus(mut self, n: usize) -> Self { self.n_gpus = n as i32; self }
    /// First step size instead of the reference's (dt_max + dt_min)/2 (rk.rs:315).
    setter!(with_initial_dt, bacon_solver_with_initial_dt);
    /// Stop every trajectory at the first zero of `w . y - c` (direction +1 rising, -1 falling, 0 both).
    pub fn with_terminal_event(mut self, w: &[f64], c: f64, direction: i32) -> Result<Self, IVPError> {
        if w.len() != self.dim || !(-1..=1).contains(&direction) { return Err(IVPError::BadArgument); }
        self.event = Some((w.to_vec(), c, direction));
        Ok(self)
    }
    /// `with_derivative(closure)` for a GPU: the functor as CUDA C++ source text (contract: include/bacon_ivp_rhs.cuh),
    /// compiled by the library with NVRTC and inlined into the kernels.  `UserError` + `last_error()` = the compiler log.
    pub fn register_source(mut self, name: &str, type_name: &str, source: &str, n_params: usize) -> Result<Self, IVPError> {
        let (n, t, s) = (CString::new(name).map_err(|_| IVPError::BadArgument)?, CString::new(type_name).map_err(|_| IVPError::BadArgument)?,
                         CString::new(source).map_err(|_| IVPError::BadArgument)?);
        let id = unsafe { sys::bacon_rhs_register_source(n.as_ptr(), t.as_ptr(), s.as_ptr(), self.dim as i32, n_params as i32) };
        if id < 0 { check(-id)?; }
        self.rhs = Some(id);
        Ok(self)
    }

    setter!(with_tolerance, bacon_solver_with_tolerance);        // rk.rs:168
    setter!(with_maximum_dt, bacon_solver_with_maximum_dt);      // rk.rs:179
    setter!(with_minimum_dt, bacon_solver_with_minimum_dt);      // rk.rs:197
    setter!(with_initial_time, bacon_solver_with_initial_time);  // rk.rs:212
    setter!(with_ending_time, bacon_solver_with_ending_time);    // rk.rs:224
    // README.md:33-38
    setter!(with_dt_max, bacon_solver_with_maximum_dt);
    setter!(with_dt_min, bacon_solver_with_minimum_dt);
    setter!(with_start, bacon_solver_with_initial_time);
    setter!(with_end, bacon_solver_with_ending_time);

    pub fn with_initial_conditions_slice(mut self, start: &[f64]) -> Result<Self, IVPError> {   // ivp.rs:177
        if start.len() != self.dim { return Err(IVPError::BadArgument); }
        self.y0 = Some(start.to_vec());
        Ok(self)
    }
    pub fn with_initial_conditions(self, start: &[f64]) -> Result<Self, IVPError> { self.with_initial_conditions_slice(start) }
    /// `with_derivative` (ivp.rs:186): the RHS is a registered CUDA device functor, looked up by name.
    pub fn with_derivative(mut self, rhs: &str) -> Result<Self, IVPError> {
        let c = CString::new(rhs).map_err(|_| IVPError::BadArgument)?;
        let id = unsafe { sys::bacon_rhs_lookup(c.as_ptr()) };
        if id < 0 { return Err(IVPError::BadArgument); }
        self.rhs = Some(id);
        Ok(self)
    }
    pub fn with_history(self, cap: usize) -> Result<Self, IVPError> { check(unsafe { sys::bacon_solver_with_history(self.h, cap as i32) })?; Ok(self) }
    pub fn with_flags(self, flags: u32) -> Result<Self, IVPError> { check(unsafe { sys::bacon_solver_with_flags(self.h, flags) })?; Ok(self) }
    pub fn with_semantics(self, s: i32) -> Result<Self, IVPError> { check(unsafe { sys::bacon_solver_with_semantics(self.h, s) })?; Ok(self) }
    pub fn build(self) -> Self { self }

    /// N initial conditions x N parameter sets: y0 is [dim][n], params [n_params][n] (or [n_params] when shared).
    pub fn solve_ivp_ensemble(&self, y0: &[f64], params: &[f64], shared_params: bool) -> Result<EnsembleResult, IVPError> {
        self.solve_ivp_ensemble_from(y0, params, shared_params, Restart::default())
    }

    /// The same, going on from a restart record (`y0` = the previous leg's `y_end`).
    pub fn solve_ivp_ensemble_from(&self, y0: &[f64], params: &[f64], shared_params: bool, restart: Restart) -> Result<EnsembleResult, IVPError> {
        let rhs = self.rhs.ok_or(IVPError::MissingParameters)?;
        let mut cfg = sys::bacon_ivp_config::default();
        check(unsafe { sys::bacon_solver_config(self.h, &mut cfg) })?;
        let (mut d, mut np) = (0, 0);
        check(unsafe { sys::bacon_rhs_info(rhs, std::ptr::null_mut(), &mut d, &mut np) })?;
        cfg.n_params = np;
        if shared_params { cfg.flags |= sys::BACON_FLAG_SHARED_PARAMS; }
        if y0.len() % self.dim != 0 { return Err(IVPError::BadArgument); }
        let n = y0.len() / self.dim;
        let cap = cfg.history_capacity as usize;
        let mut r = EnsembleResult {
            n, dim: self.dim, capacity: cap,
            y_end: vec![0.0; self.dim * n], t_end: vec![0.0; n], dt_end: vec![0.0; n], status: vec![-1; n],
            n_accept: vec![0; n], n_reject: vec![0; n], n_rhs: vec![0; n],
            hist: vec![0.0; n * cap * (1 + self.dim)], hist_len: vec![0; n],
            cfg, rhs, y0: if cap > 0 { y0.to_vec() } else { Vec::new() }, params: if cap > 0 { params.to_vec() } else { Vec::new() },
            t_start: restart.t_start_each.map(|t| t.to_vec()).unwrap_or_default(),
        };
        if restart.t_start_each.map_or(false, |t| t.len() != n) || restart.dt_start_each.map_or(false, |t| t.len() != n) {
            return Err(IVPError::BadArgument);
        }
        let out = sys::bacon_ivp_result {
            y_end: r.y_end.as_mut_ptr(), t_end: r.t_end.as_mut_ptr(), dt_end: r.dt_end.as_mut_ptr(),
            status: r.status.as_mut_ptr(), n_accept: r.n_accept.as_mut_ptr(), n_reject: r.n_reject.as_mut_ptr(),
            n_rhs: r.n_rhs.as_mut_ptr(),
            hist: if cap > 0 { r.hist.as_mut_ptr() } else { std::ptr::null_mut() },
            hist_len: if cap > 0 { r.hist_len.as_mut_ptr() } else { std::ptr::null_mut() },
            t_start: std::ptr::null(),
        };
        let pptr = if params.is_empty() { std::ptr::null() } else { params.as_ptr() };
        let opts = sys::bacon_ivp_options {
            t_start_each: restart.t_start_each.map_or(std::ptr::null(), |t| t.as_ptr()),
            dt_start_each: restart.dt_start_each.map_or(std::ptr::null(), |t| t.as_ptr()),
            event_w: self.event.as_ref().map_or(std::ptr::null(), |e| e.0.as_ptr()),
            event_c: self.event.as_ref().map_or(0.0, |e| e.1),
            event_direction: self.event.as_ref().map_or(0, |e| e.2),
            reserved: 0,
        };
        check(unsafe { sys::bacon_ivp_solve_ensemble_ex(&cfg, rhs, n, y0.as_ptr(), pptr, &opts, &out, self.n_gpus) })?;
        Ok(r)
    }

    /// `solve(data)` + `collect_vec` (rk.rs:249-343, ivp.rs:209-211) for the trajectory set by with_initial_conditions.
    pub fn solve(self, data: &[f64]) -> Result<Path, IVPError> {
        let y0 = self.y0.clone().ok_or(IVPError::MissingParameters)?;
        let s = self.with_history(1 << 16)?;
        let mut r = s.solve_ivp_ensemble(&y0, data, false)?;
        let s = if r.n_accept[0] as usize > (1 << 16) {  // collect_vec grows its Vec (ivp.rs:209-211): once more, with the capacity reported
            let s = s.with_history(r.n_accept[0] as usize)?;
            r = s.solve_ivp_ensemble(&y0, data, false)?;
            s
        } else { s };
        let _ = s;
        if r.status[0] != STOPPED_AT_EVENT { check(r.status[0])?; }
        Ok(r.path(0))
    }
    pub fn solve_ivp(self, rhs: &str, data: &[f64]) -> Result<Path, IVPError> { self.with_derivative(rhs)?.solve(data) }   // README.md:40
}

pub type RungeKutta45 = Solver<{ sys::BACON_RK45 }>;  // rk.rs:561
pub type RungeKutta23 = Solver<{ sys::BACON_RK23 }>;  // rk.rs:656
pub type BDF6 = Solver<{ sys::BACON_BDF6 }>;          // bdf.rs:706
pub type BDF2 = Solver<{ sys::BACON_BDF2 }>;          // bdf.rs:762
pub type Adams5 = Solver<{ sys::BACON_ADAMS5 }>;      // adams.rs:633
pub type Adams3 = Solver<{ sys::BACON_ADAMS3 }>;      // adams.rs:693
pub type Euler = Solver<{ sys::BACON_EULER }>;        // ivp.rs:269
pub type RK45 = RungeKutta45;                         // README.md:24
pub type RK23 = RungeKutta23;
