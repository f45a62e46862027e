"""Loader of the C-ABI shared library (bacon_b200/libbacon_ivp.so).

There is no CPU fallback: if the library is missing this raises, and every solve
ends in a CUDA kernel launch or an error code (include/bacon_ivp.h).
"""
import ctypes as C
import os

from . import _abi

_HERE = os.path.dirname(os.path.abspath(__file__))
# BACON_IVP_LIB: load another build of the same ABI (A/B timing of kernel variants on one box)
LIB_PATH = os.environ.get("BACON_IVP_LIB") or os.path.join(_HERE, "libbacon_ivp.so")
_LIB = None


class LibraryMissing(RuntimeError):
    pass


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise LibraryMissing(
            f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C bacon_b200/csrc`). There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, i32, u32, u64, dbl, sz = C.c_void_p, C.c_int, C.c_uint32, C.c_uint64, C.c_double, C.c_size_t
    cfgp, resp, optp = C.POINTER(_abi.Config), C.POINTER(_abi.Result), C.POINTER(_abi.Options)
    sig = {
        "bacon_abi_version": (i32, []),
        "bacon_solver_new": (vp, [i32, i32]),
        "bacon_solver_new_static": (i32, [i32, i32, C.POINTER(vp)]),
        "bacon_solver_new_dyn": (i32, [i32, i32, i32, C.POINTER(vp)]),
        "bacon_solver_with_initial_dt": (i32, [vp, dbl]),
        "bacon_ivp_solve_ensemble_ex": (i32, [cfgp, i32, sz, vp, vp, optp, resp, i32]),
        "bacon_ivp_solve_ensemble_device_ex": (i32, [cfgp, i32, sz, vp, vp, optp, resp, vp]),
        "bacon_solver_free": (None, [vp]),
        "bacon_solver_with_tolerance": (i32, [vp, dbl]),
        "bacon_solver_with_maximum_dt": (i32, [vp, dbl]),
        "bacon_solver_with_minimum_dt": (i32, [vp, dbl]),
        "bacon_solver_with_initial_time": (i32, [vp, dbl]),
        "bacon_solver_with_ending_time": (i32, [vp, dbl]),
        "bacon_solver_with_semantics": (i32, [vp, i32]),
        "bacon_solver_with_flags": (i32, [vp, u32]),
        "bacon_solver_with_history": (i32, [vp, i32]),
        "bacon_solver_with_max_attempts": (i32, [vp, u64]),
        "bacon_solver_config": (i32, [vp, cfgp]),
        "bacon_ivp_validate": (i32, [cfgp]),
        "bacon_rhs_register": (i32, [vp]),
        "bacon_rhs_lookup": (i32, [C.c_char_p]),
        "bacon_rhs_register_source": (i32, [C.c_char_p, C.c_char_p, C.c_char_p, i32, i32]),
        "bacon_rhs_count": (i32, []),
        "bacon_rhs_info": (i32, [i32, C.POINTER(C.c_char_p), C.POINTER(i32), C.POINTER(i32)]),
        "bacon_ivp_solve_ensemble": (i32, [cfgp, i32, sz, vp, vp, resp]),
        "bacon_ivp_solve_ensemble_device": (i32, [cfgp, i32, sz, vp, vp, resp, vp]),
        "bacon_ivp_solve_ensemble_multi": (i32, [cfgp, i32, sz, vp, vp, resp, i32]),
        "bacon_ivp_last_launch": (i32, [C.POINTER(_abi.LaunchInfo)]),
        "bacon_last_error": (C.c_char_p, []),
        "bacon_status_name": (C.c_char_p, [i32]),
        "bacon_fp64_peak_tflops": (dbl, [i32, vp]),
        "bacon_device_sm_count": (i32, []),
        "bacon_host_alloc": (vp, [sz]),
        "bacon_host_free": (None, [vp]),
        "bacon_ivp_sample_paths": (i32, [cfgp, i32, sz, vp, vp, resp, sz, vp, vp]),
        "bacon_ivp_sample_paths_device": (i32, [cfgp, i32, sz, vp, vp, resp, sz, vp, vp, vp]),
        "bacon_ivp_locate_events": (i32, [cfgp, i32, sz, vp, vp, resp, vp, dbl, i32, i32, vp, vp]),
        "bacon_ivp_locate_events_device": (i32, [cfgp, i32, sz, vp, vp, resp, vp, dbl, i32, i32, vp, vp, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _LIB = L
    return L


def last_error():
    return lib().bacon_last_error().decode()
