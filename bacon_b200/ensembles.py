"""Seeded synthetic ensembles of BASELINE.json's configs.

Counter-based generator (SplitMix64): value k of trajectory i depends only on
(seed, i, k), so any shard of any ensemble size reproduces the same trajectories —
ranks generate their own shard, the CPU baseline its sub-sample, with no exchange.
"""
import numpy as np

_GOLDEN = np.uint64(0x9E3779B97F4A7C15)


def splitmix64(x):
    x = np.asarray(x, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = x + _GOLDEN
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def uniform01(seed, index, stream):
    """U[0,1) for (trajectory index, stream k): 53 mantissa bits of SplitMix64(seed ^ mix(index, k))."""
    index = np.asarray(index, dtype=np.uint64)
    with np.errstate(over="ignore"):
        ctr = splitmix64(index * np.uint64(64) + np.uint64(stream)) ^ np.uint64(seed)
    return (splitmix64(ctr) >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def normal(seed, index, stream):
    """N(0,1) by Box-Muller from two streams (2*stream, 2*stream+1)."""
    u1 = uniform01(seed, index, 2 * stream)
    u2 = uniform01(seed, index, 2 * stream + 1)
    return np.sqrt(-2.0 * np.log1p(-u1)) * np.cos(2.0 * np.pi * u2)


# ---- config 2: Lorenz-63, RK45, tol 1e-8 (the headline)
LORENZ = dict(rhs="lorenz", method="RK45", dim=3, seed=0x5EED0001, params=(10.0, 28.0, 8.0 / 3.0),
              t_start=0.0, t_end=5.0, tol=1e-8, dt_min=1e-9, dt_max=0.1, n=1 << 20)


def lorenz_y0(indices):
    """y0 ~ U([-15,15] x [-20,20] x [5,40]); returns (3, len(indices)) float64."""
    i = np.asarray(indices, dtype=np.uint64)
    s = LORENZ["seed"]
    return np.stack([-15.0 + 30.0 * uniform01(s, i, 0), -20.0 + 40.0 * uniform01(s, i, 1),
                     5.0 + 35.0 * uniform01(s, i, 2)])


# ---- config 3: Van der Pol mu-sweep, RK23, tol 1e-10
VDP = dict(rhs="vdp", method="RK23", dim=2, t_start=0.0, t_end=0.25, tol=1e-10, dt_min=1e-12, dt_max=0.1,
           n=1 << 22)


def vdp_problem(indices, n_total):
    i = np.asarray(indices, dtype=np.float64)
    y0 = np.stack([np.full(i.shape, 2.0), np.zeros(i.shape)])
    mu = 0.1 + 4.9 * i / float(max(n_total - 1, 1))
    return y0, mu[None, :]


# ---- config 4: 32-dim linear ODE with per-trajectory A, RK45, dense output
LINEAR32 = dict(rhs="linear32", method="RK45", dim=32, seed=0x5EED0004, t_start=0.0, t_end=4.0, tol=1e-8,
                dt_min=1e-9, dt_max=0.1, n=1 << 18, history_capacity=256)


def linear32_problem(indices):
    """A_i = -0.5 I + 0.5 (G - G^T)/sqrt(32) + 0.1 G'/sqrt(32); y0 ~ N(0, I).
    Returns y0 (32, m) and A (m, 32, 32) row-major per trajectory (AoS parameter block)."""
    i = np.asarray(indices, dtype=np.uint64)
    m = i.shape[0]
    s = LINEAR32["seed"]
    G = np.empty((m, 32, 32))
    Gp = np.empty((m, 32, 32))
    for r in range(32):
        for c in range(32):
            G[:, r, c] = normal(s, i, 1 + r * 32 + c)
            Gp[:, r, c] = normal(s, i, 1 + 1024 + r * 32 + c)
    A = 0.5 * (G - np.transpose(G, (0, 2, 1))) / np.sqrt(32.0) + 0.1 * Gp / np.sqrt(32.0)
    A[:, np.arange(32), np.arange(32)] += -0.5
    y0 = np.stack([normal(s, i, 1 + 2048 + d) for d in range(32)])
    return y0, np.ascontiguousarray(A)


# ---- config 5: Robertson kinetics, BDF6, tol 1e-6
ROBERTSON = dict(rhs="robertson", method="BDF6", dim=3, seed=0x5EED0005, t_start=0.0, t_end=0.5, tol=1e-6,
                 dt_min=1e-10, dt_max=1e-4, n=1 << 20)


def robertson_problem(indices):
    i = np.asarray(indices, dtype=np.uint64)
    s = ROBERTSON["seed"]
    base = (0.04, 3e7, 1e4)
    k = np.stack([base[j] * (1.0 + 0.1 * (2.0 * uniform01(s, i, j) - 1.0)) for j in range(3)])
    y0 = np.stack([np.ones(i.shape), np.zeros(i.shape), np.zeros(i.shape)])
    return y0, k


# ---- algorithmic flop counts (FMA = 2; sqrt, /, pow = 1; structural zeros not counted)
def rk_flops(method, dim, f_rhs, n_attempts, n_accepts):
    """RK pair with s stages: per attempt s*F_rhs + [2 nnzA + s + (2 nnz_e - 1) + 2] D + 2(s-1) + 6,
    per accepted step + 2 nnz_b D + 1."""
    s, nnz_a, nnz_e, nnz_b = {"RK45": (6, 15, 5, 4), "RK23": (4, 5, 4, 3)}[method]
    per_attempt = s * f_rhs + (2 * nnz_a + s + (2 * nnz_e - 1) + 2) * dim + 2 * (s - 1) + 6
    per_accept = 2 * nnz_b * dim + 1
    return float(n_attempts) * per_attempt + float(n_accepts) * per_accept


F_RHS = {"lorenz": 8, "vdp": 5, "robertson": 8, "linear32": 2 * 32 * 32}
