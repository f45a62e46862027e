"""ctypes mirror of include/bacon_ivp.h (structs, enums).  No logic here."""
import ctypes as C

ABI_VERSION = 6

# bacon_method  (rk.rs:561, rk.rs:656, bdf.rs:706, bdf.rs:762)
RK45, RK23, BDF6, BDF2, ADAMS5, ADAMS3, EULER = 0, 1, 2, 3, 4, 5, 6
N_METHODS = 7
METHOD_NAMES = {RK45: "RK45", RK23: "RK23", BDF6: "BDF6", BDF2: "BDF2", ADAMS5: "Adams5", ADAMS3: "Adams3",
                EULER: "Euler"}

# bacon_status  (1..12 = IVPError, src/ivp.rs:50-76)
OK = 0
E_MISSING_PARAMETERS = 1
E_USER = 2
E_TOLERANCE_OOB = 3
E_TIME_DELTA_OOB = 4
E_TIME_END_OOB = 5
E_TIME_START_OOB = 6
E_FROM_PRIMITIVE = 7
E_MIN_DT_EXCEEDED = 8
E_MAX_ITER = 9
E_SINGULAR = 10
E_DYNAMIC_ON_STATIC = 11
E_STATIC_ON_DYNAMIC = 12
E_NONFINITE = 13
E_MAX_ATTEMPTS = 14
E_HISTORY_OVERFLOW = 15
E_CUDA = 16
E_BAD_ARGUMENT = 17
E_UNSUPPORTED = 18
STOPPED_AT_EVENT = 19  # per-trajectory, not an error: the integration stopped at a terminal event
DIM_DYN = 0            # the `Dyn` type parameter of bacon_solver_new_static / bacon_solver_new_dyn

STATUS_NAMES = {
    0: "Ok", 1: "MissingParameters", 2: "UserError", 3: "ToleranceOOB", 4: "TimeDeltaOOB",
    5: "TimeEndOOB", 6: "TimeStartOOB", 7: "FromPrimitiveFailure", 8: "MinimumTimeDeltaExceeded",
    9: "MaximumIterationsExceeded", 10: "SingularMatrix", 11: "DynamicOnStatic",
    12: "StaticOnDynamic", 13: "NonFinite", 14: "MaxAttempts", 15: "HistoryOverflow",
    16: "CudaError", 17: "BadArgument", 18: "Unsupported", 19: "StoppedAtEvent",
}

SEM_CORRECTED, SEM_LITERAL = 0, 1
FLAG_STRICT_FP, FLAG_SHARED_PARAMS, FLAG_BDF_NEWTON, FLAG_PARAMS_AOS, FLAG_ZERO_COPY = 1, 2, 4, 8, 16


class Config(C.Structure):
    """bacon_ivp_config"""
    _fields_ = [
        ("method", C.c_int32), ("dim", C.c_int32), ("n_params", C.c_int32),
        ("semantics", C.c_int32), ("flags", C.c_uint32), ("history_capacity", C.c_int32),
        ("dt_min", C.c_double), ("dt_max", C.c_double), ("tol", C.c_double),
        ("t_start", C.c_double), ("t_end", C.c_double), ("max_attempts", C.c_uint64),
        ("dt_init", C.c_double),
    ]


class Result(C.Structure):
    """bacon_ivp_result — raw addresses (host or device)."""
    _fields_ = [
        ("y_end", C.c_void_p), ("t_end", C.c_void_p), ("dt_end", C.c_void_p),
        ("status", C.c_void_p), ("n_accept", C.c_void_p), ("n_reject", C.c_void_p),
        ("n_rhs", C.c_void_p), ("hist", C.c_void_p), ("hist_len", C.c_void_p),
        ("t_start", C.c_void_p),
    ]


class Options(C.Structure):
    """bacon_ivp_options — optional inputs of a solve (restart record, terminal event)."""
    _fields_ = [
        ("t_start_each", C.c_void_p), ("dt_start_each", C.c_void_p), ("event_w", C.c_void_p),
        ("event_c", C.c_double), ("event_direction", C.c_int32), ("reserved", C.c_int32),
    ]


class LaunchInfo(C.Structure):
    """bacon_ivp_launch_info"""
    _fields_ = [
        ("kernel_ms", C.c_float), ("h2d_ms", C.c_float), ("d2h_ms", C.c_float),
        ("grid", C.c_int32), ("block", C.c_int32), ("regs_per_thread", C.c_int32),
        ("n_kernels", C.c_int32),
    ]


# every symbol include/bacon_ivp.h declares (checked by tests/test_abi.py)
EXPORTED_SYMBOLS = [
    "bacon_abi_version", "bacon_solver_new", "bacon_solver_free", "bacon_solver_with_tolerance",
    "bacon_solver_with_maximum_dt", "bacon_solver_with_minimum_dt", "bacon_solver_with_initial_time",
    "bacon_solver_with_ending_time", "bacon_solver_with_semantics", "bacon_solver_with_flags",
    "bacon_solver_with_history", "bacon_solver_with_max_attempts", "bacon_solver_config",
    "bacon_ivp_validate", "bacon_rhs_register", "bacon_rhs_register_source", "bacon_rhs_lookup", "bacon_rhs_count",
    "bacon_rhs_info", "bacon_ivp_solve_ensemble", "bacon_ivp_solve_ensemble_device",
    "bacon_ivp_solve_ensemble_multi", "bacon_ivp_last_launch", "bacon_last_error",
    "bacon_status_name", "bacon_fp64_peak_tflops", "bacon_device_sm_count", "bacon_host_alloc", "bacon_host_free",
    "bacon_ivp_sample_paths", "bacon_ivp_sample_paths_device", "bacon_ivp_locate_events",
    "bacon_ivp_locate_events_device",
    "bacon_solver_new_static", "bacon_solver_new_dyn", "bacon_solver_with_initial_dt",
    "bacon_ivp_solve_ensemble_ex", "bacon_ivp_solve_ensemble_device_ex",
]
