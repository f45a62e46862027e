"""The reference's own IVP test problems (parameters copied as data, file:line cited), replayed
through the oracle (CPU tests) and through the CUDA path (GPU tests).

rk.rs:682-758   rungekutta_quadratic / rungekutta_sine, each for RungeKutta23 and RungeKutta45
bdf.rs:785-1063 bdf6_{exp,unstable,quadratic,sin}, bdf2_{exp,unstable,quadratic,sin}
adams.rs:714-922 adams5_{exp,quadratic,sine}, adams3_{exp,quadratic,sine}   (live: y-dependent RHS included)
ivp.rs:539-653  euler_{cos,exp,quadratic,sin} (euler_dynamic_cos is euler_cos through new_dyn)
Each asserts |y(t_i) - exact(t_i)| <= eps for EVERY yielded point.
"""
import numpy as np

RK_CASES = [
    # name, rhs, y0, cfg, exact, eps, expected accepted (SURVEY.md §8c predicted-answer table)
    ("rungekutta_quadratic", "quadratic", 1.0, dict(dt_min=1e-4, dt_max=0.1, tol=1e-5, t_start=0.0, t_end=10.0),
     lambda t: 1.0 - t * t, 1e-4, 101),                                   # rk.rs:685-719
    ("rungekutta_sine", "cos", 0.0, dict(dt_min=1e-3, dt_max=1e-2, tol=1e-4, t_start=0.0, t_end=10.0),
     lambda t: np.sin(t), 1e-2, 1001),                                     # rk.rs:724-758
]

_BDF = dict(dt_min=1e-5, dt_max=0.1, tol=1e-5, t_start=0.0)
BDF_CASES = [
    # name, method, rhs, y0, t_end, exact, eps, yielded points in REF_CORRECTED, internal final time in REF_LITERAL
    ("bdf6_exp", "BDF6", "exp", 1.0, 7.0, lambda t: np.exp(t), 0.01, 83, 7.3),            # bdf.rs:786-818
    ("bdf6_unstable", "BDF6", "decay", 1.0, 10.0, lambda t: np.exp(-t), 0.01, 99, 14.45),  # bdf.rs:821-853
    ("bdf6_quadratic", "BDF6", "quadratic", 1.0, 2.0, lambda t: 1.0 - t * t, 0.01, 18, 7.3),  # bdf.rs:856-888
    ("bdf6_sin", "BDF6", "cos", 0.0, 6.0, lambda t: np.sin(t), 0.01, 64, 7.3),              # bdf.rs:891-923
    ("bdf2_exp", "BDF2", "exp", 1.0, 7.0, lambda t: np.exp(t), 0.01, 21097, 9.175),         # bdf.rs:926-958
    ("bdf2_unstable", "BDF2", "decay", 1.0, 10.0, lambda t: np.exp(-t), 0.01, None, 12.1875),  # bdf.rs:961-993
    ("bdf2_quadratic", "BDF2", "quadratic", 1.0, 1.0, lambda t: 1.0 - t * t, 0.01, None, 3.1),  # bdf.rs:996-1028
    ("bdf2_sin", "BDF2", "cos", 0.0, 6.0, lambda t: np.sin(t), 0.01, 1690, 6.15),           # bdf.rs:1031-1063
]


def bdf_cfg(t_end):
    return dict(_BDF, t_end=t_end)


ADAMS_CASES = [
    # name, method, rhs, y0, cfg, exact, eps, (yielded, rejected) in REF_LITERAL (as written), same in REF_CORRECTED (D10 fixed)
    ("adams5_exp", "Adams5", "exp", 1.0, dict(dt_min=1e-5, dt_max=0.1, tol=5e-4, t_start=0.0, t_end=2.0),
     lambda t: np.exp(t), 0.01, (541, 67), (19, 0)),                                                    # adams.rs:715-747
    ("adams5_quadratic", "Adams5", "quadratic", 1.0, dict(dt_min=1e-7, dt_max=1e-3, tol=0.01, t_start=0.0, t_end=5.0),
     lambda t: 1.0 - t * t, 0.01, (4999, 0), (4999, 0)),                                                  # adams.rs:750-782
    ("adams5_sine", "Adams5", "cos", 0.0, dict(dt_min=1e-5, dt_max=1e-3, tol=0.01, t_start=0.0, t_end=2.0 * np.pi),
     lambda t: np.sin(t), 0.01, (6283, 0), (6283, 0)),                                                    # adams.rs:785-817
    ("adams3_exp", "Adams3", "exp", 1.0, dict(dt_min=1e-5, dt_max=0.1, tol=1e-3, t_start=0.0, t_end=2.0),
     lambda t: np.exp(t), 0.01, (100, 21), (26, 1)),                                                    # adams.rs:820-852
    ("adams3_quadratic", "Adams3", "quadratic", 1.0, dict(dt_min=1e-5, dt_max=1e-3, tol=0.1, t_start=0.0, t_end=5.0),
     lambda t: 1.0 - t * t, 0.01, (5000, 0), (5000, 0)),                                                  # adams.rs:855-887
    ("adams3_sine", "Adams3", "cos", 0.0, dict(dt_min=1e-5, dt_max=1e-3, tol=0.01, t_start=0.0, t_end=2.0 * np.pi),
     lambda t: np.sin(t), 0.01, (6284, 0), (6284, 0)),                                                    # adams.rs:890-922
]

EULER_CASES = [
    # name, rhs, y0 (list), dt, component-0 exact, eps; all on t in [0, 1] (ivp.rs:497-512 helper)
    ("euler_cos", "harmonic", [1.0, 0.0], 0.01, lambda t: np.cos(t), 0.01),     # ivp.rs:587-602 (y'' = -y, p = w = 1)
    ("euler_exp", "exp", [1.0], 0.005, lambda t: np.exp(t), 0.01),             # ivp.rs:605-619
    ("euler_quadratic", "quadratic", [1.0], 0.01, lambda t: 1.0 - t * t, 0.01),  # ivp.rs:622-636
    ("euler_sin", "cos", [0.0], 0.01, lambda t: np.sin(t), 0.01),              # ivp.rs:639-653
]
