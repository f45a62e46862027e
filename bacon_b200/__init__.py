"""bacon_b200 — B200-native ensemble IVP engine behind `bacon_sci::ivp`'s builder API.

The product is bacon_b200/libbacon_ivp.so (hand-written CUDA for sm_100a + a C ABI,
include/bacon_ivp.h).  This package is the thin host-side mirror of the reference's
solver front end over that ABI.
"""
from . import _abi  # noqa: F401
from .ivp import (BDF2, BDF6, RK23, RK45, Adams3, Adams5, EnsembleResult, Euler, IVPError,  # noqa: F401
                  RungeKutta23, RungeKutta45, fp64_peak_tflops, last_launch, pinned_empty, register_rhs_source,
                  solve_ivp)
