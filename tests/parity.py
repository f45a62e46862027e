"""Shared helpers of the parity tests: run the same ensemble through the CUDA path
(C ABI, host buffers) and through the CPU oracle and compare per trajectory."""
import numpy as np

from bacon_b200 import _abi

METHODS = {"RK45": _abi.RK45, "RK23": _abi.RK23, "BDF6": _abi.BDF6, "BDF2": _abi.BDF2, "Adams5": _abi.ADAMS5,
           "Adams3": _abi.ADAMS3, "Euler": _abi.EULER}


def make_solver(engine, method, dim, *, dt_min, dt_max, tol, t_start, t_end, rhs, semantics=0, flags=0,
                history=0, max_attempts=0):
    cls = {"RK45": engine.RungeKutta45, "RK23": engine.RungeKutta23, "BDF6": engine.BDF6, "BDF2": engine.BDF2,
           "Adams5": engine.Adams5, "Adams3": engine.Adams3, "Euler": engine.Euler}[method]
    s = (cls.new(dim).with_minimum_dt(dt_min).with_maximum_dt(dt_max).with_tolerance(tol)
         .with_initial_time(t_start).with_ending_time(t_end).with_derivative(rhs)
         .with_semantics(semantics).with_flags(flags).with_history(history).with_max_attempts(max_attempts))
    return s


def run_both(engine, oracle, method, rhs, y0, params=None, *, shared_params=False, strict=False, semantics=0,
             history=0, max_attempts=0, pow_mode=None, n_gpus=1, extra_flags=0, **cfg):
    dim = y0.shape[0]
    flags = (_abi.FLAG_STRICT_FP if strict else 0) | extra_flags
    s = make_solver(engine, method, dim, rhs=rhs, semantics=semantics, flags=flags, history=history,
                    max_attempts=max_attempts, **cfg)
    gpu = s.solve_ivp_ensemble(y0, params, shared_params=shared_params, n_gpus=n_gpus)
    if pow_mode is None:
        pow_mode = 1 if (strict or semantics == 1) else 0  # strict kernels take (.)^(1/4) as sqrt(sqrt())
    ref = oracle.solve_ensemble(METHODS[method], rhs, y0, params, shared_params=shared_params, semantics=semantics,
                                history_capacity=history, max_attempts=max_attempts, pow_mode=pow_mode, **cfg)
    return gpu, ref


def rel_err(y_gpu, y_ref):
    """|| y_gpu - y_ref ||_2 / || y_ref ||_2 per trajectory; arrays are (dim, n)."""
    num = np.sqrt(((y_gpu - y_ref) ** 2).sum(axis=0))
    den = np.sqrt((y_ref ** 2).sum(axis=0))
    return num / np.maximum(den, 1e-300)


def band(tol):
    """north_star: final state within a relative error of max(10*tol, 1e-12) per trajectory."""
    return max(10.0 * tol, 1e-12)
