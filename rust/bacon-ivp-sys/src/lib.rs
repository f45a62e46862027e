//! Raw bindings to `include/bacon_ivp.h` (ABI version 6).  UNVERIFIED: no Rust toolchain in the build image.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_double, c_int, c_void};

pub const BACON_RK45: c_int = 0;
pub const BACON_RK23: c_int = 1;
pub const BACON_BDF6: c_int = 2;
pub const BACON_BDF2: c_int = 3;
pub const BACON_ADAMS5: c_int = 4;
pub const BACON_ADAMS3: c_int = 5;
pub const BACON_EULER: c_int = 6;

/// the `Dyn` type parameter of bacon_solver_new_static / bacon_solver_new_dyn
pub const BACON_DIM_DYN: c_int = 0;
/// per-trajectory status, not an error: the integration stopped at a terminal event
pub const BACON_STOPPED_AT_EVENT: i32 = 19;

pub const BACON_SEM_CORRECTED: i32 = 0;
pub const BACON_SEM_LITERAL: i32 = 1;
pub const BACON_FLAG_STRICT_FP: u32 = 1;
pub const BACON_FLAG_SHARED_PARAMS: u32 = 2;
pub const BACON_FLAG_BDF_NEWTON: u32 = 4;
pub const BACON_FLAG_PARAMS_AOS: u32 = 8;
pub const BACON_FLAG_ZERO_COPY: u32 = 16;

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct bacon_ivp_config {
    pub method: i32,
    pub dim: i32,
    pub n_params: i32,
    pub semantics: i32,
    pub flags: u32,
    pub history_capacity: i32,
    pub dt_min: c_double,
    pub dt_max: c_double,
    pub tol: c_double,
    pub t_start: c_double,
    pub t_end: c_double,
    pub max_attempts: u64,
    pub dt_init: c_double,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct bacon_ivp_result {
    pub y_end: *mut c_double,
    pub t_end: *mut c_double,
    pub dt_end: *mut c_double,
    pub status: *mut i32,
    pub n_accept: *mut u32,
    pub n_reject: *mut u32,
    pub n_rhs: *mut u32,
    pub hist: *mut c_double,
    pub hist_len: *mut u32,
    /// INPUT of the path queries only: per-trajectory start times of a resumed leg (NULL = cfg.t_start)
    pub t_start: *const c_double,
}

/// Optional inputs of a solve: restart record (per-trajectory start time and first dt) and terminal event.
#[repr(C)]
#[derive(Clone, Copy)]
pub struct bacon_ivp_options {
    pub t_start_each: *const c_double,
    pub dt_start_each: *const c_double,
    pub event_w: *const c_double,
    pub event_c: c_double,
    pub event_direction: i32,
    pub reserved: i32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct bacon_ivp_launch_info {
    pub kernel_ms: f32,
    pub h2d_ms: f32,
    pub d2h_ms: f32,
    pub grid: i32,
    pub block: i32,
    pub regs_per_thread: i32,
    pub n_kernels: i32,
}

#[repr(C)]
pub struct bacon_solver {
    _private: [u8; 0],
}

extern "C" {
    pub fn bacon_abi_version() -> c_int;
    pub fn bacon_solver_new(method: c_int, dim: c_int) -> *mut bacon_solver;
    pub fn bacon_solver_new_static(method: c_int, dim_type: c_int, out: *mut *mut bacon_solver) -> c_int;
    pub fn bacon_solver_new_dyn(method: c_int, dim_type: c_int, size: c_int, out: *mut *mut bacon_solver) -> c_int;
    pub fn bacon_solver_with_initial_dt(s: *mut bacon_solver, dt: c_double) -> c_int;
    pub fn bacon_solver_free(s: *mut bacon_solver);
    pub fn bacon_solver_with_tolerance(s: *mut bacon_solver, tol: c_double) -> c_int;
    pub fn bacon_solver_with_maximum_dt(s: *mut bacon_solver, max: c_double) -> c_int;
    pub fn bacon_solver_with_minimum_dt(s: *mut bacon_solver, min: c_double) -> c_int;
    pub fn bacon_solver_with_initial_time(s: *mut bacon_solver, t: c_double) -> c_int;
    pub fn bacon_solver_with_ending_time(s: *mut bacon_solver, t: c_double) -> c_int;
    pub fn bacon_solver_with_semantics(s: *mut bacon_solver, semantics: c_int) -> c_int;
    pub fn bacon_solver_with_flags(s: *mut bacon_solver, flags: u32) -> c_int;
    pub fn bacon_solver_with_history(s: *mut bacon_solver, capacity: c_int) -> c_int;
    pub fn bacon_solver_with_max_attempts(s: *mut bacon_solver, cap: u64) -> c_int;
    pub fn bacon_solver_config(s: *const bacon_solver, out: *mut bacon_ivp_config) -> c_int;
    pub fn bacon_ivp_validate(cfg: *const bacon_ivp_config) -> c_int;
    pub fn bacon_rhs_lookup(name: *const c_char) -> c_int;
    /// A user right-hand side as CUDA C++ source text, compiled by the library with NVRTC (no nvcc at build time).
    pub fn bacon_rhs_register_source(name: *const c_char, type_name: *const c_char, source: *const c_char,
                                     dim: c_int, n_params: c_int) -> c_int;
    pub fn bacon_rhs_count() -> c_int;
    pub fn bacon_rhs_info(id: c_int, name: *mut *const c_char, dim: *mut c_int, n_params: *mut c_int) -> c_int;
    pub fn bacon_ivp_solve_ensemble(cfg: *const bacon_ivp_config, rhs_id: c_int, n: usize, y0: *const c_double,
                                    params: *const c_double, out: *const bacon_ivp_result) -> c_int;
    pub fn bacon_ivp_solve_ensemble_device(cfg: *const bacon_ivp_config, rhs_id: c_int, n: usize, d_y0: *const c_double,
                                           d_params: *const c_double, d_out: *const bacon_ivp_result,
                                           stream: *mut c_void) -> c_int;
    pub fn bacon_ivp_solve_ensemble_multi(cfg: *const bacon_ivp_config, rhs_id: c_int, n: usize, y0: *const c_double,
                                          params: *const c_double, out: *const bacon_ivp_result, n_gpus: c_int) -> c_int;
    pub fn bacon_ivp_solve_ensemble_ex(cfg: *const bacon_ivp_config, rhs_id: c_int, n: usize, y0: *const c_double,
                                       params: *const c_double, options: *const bacon_ivp_options,
                                       out: *const bacon_ivp_result, n_gpus: c_int) -> c_int;
    pub fn bacon_ivp_solve_ensemble_device_ex(cfg: *const bacon_ivp_config, rhs_id: c_int, n: usize,
                                              d_y0: *const c_double, d_params: *const c_double,
                                              options: *const bacon_ivp_options, d_out: *const bacon_ivp_result,
                                              stream: *mut c_void) -> c_int;
    /// Queries on stored paths (cubic Hermite continuous extension of a dense-output solve; not in bacon 0.16.2).
    pub fn bacon_ivp_sample_paths(cfg: *const bacon_ivp_config, rhs_id: c_int, n: usize, y0: *const c_double,
                                  params: *const c_double, solved: *const bacon_ivp_result, n_times: usize,
                                  times: *const c_double, samples: *mut c_double) -> c_int;
    pub fn bacon_ivp_sample_paths_device(cfg: *const bacon_ivp_config, rhs_id: c_int, n: usize, d_y0: *const c_double,
                                         d_params: *const c_double, d_solved: *const bacon_ivp_result, n_times: usize,
                                         d_times: *const c_double, d_samples: *mut c_double, stream: *mut c_void) -> c_int;
    pub fn bacon_ivp_locate_events(cfg: *const bacon_ivp_config, rhs_id: c_int, n: usize, y0: *const c_double,
                                   params: *const c_double, solved: *const bacon_ivp_result, w: *const c_double,
                                   c: c_double, direction: c_int, capacity: c_int, events: *mut c_double,
                                   n_events: *mut u32) -> c_int;
    pub fn bacon_ivp_locate_events_device(cfg: *const bacon_ivp_config, rhs_id: c_int, n: usize, d_y0: *const c_double,
                                          d_params: *const c_double, d_solved: *const bacon_ivp_result,
                                          w: *const c_double, c: c_double, direction: c_int, capacity: c_int,
                                          d_events: *mut c_double, d_n_events: *mut u32, stream: *mut c_void) -> c_int;
    pub fn bacon_ivp_last_launch(out: *mut bacon_ivp_launch_info) -> c_int;
    pub fn bacon_last_error() -> *const c_char;
    pub fn bacon_status_name(status: c_int) -> *const c_char;
    pub fn bacon_fp64_peak_tflops(iters: c_int, stream: *mut c_void) -> c_double;
    pub fn bacon_device_sm_count() -> c_int;
    pub fn bacon_host_alloc(bytes: usize) -> *mut c_void;
    pub fn bacon_host_free(p: *mut c_void);
}
