// rhs_builtin.cuh — built-in right-hand sides: the device form of the reference's
// `Derivative` closures (src/ivp.rs:34-48).  A RHS is a stateless device functor
//   static constexpr int DIM, NPARAM;
//   __device__ void operator()(double t, const double (&y)[DIM], const double* p, double (&dy)[DIM]) const;
//   (optional, fast RK path)     __device__ void scaled(double h, double t, const double (&y)[DIM], const double* p,
//                                                       double (&k)[DIM]) const;   k = h * f(t, y)
//   (optional, BDF Newton path)  __device__ void jac(double t, const double (&y)[DIM], const double* p,
//                                                    double (&J)[DIM][DIM]) const;   J[r][c] = d f_r / d y_c
// The per-trajectory parameter block `p` is read-only: the reference clones the
// user data for every evaluation (rk.rs:380, bdf.rs:351), so a RHS cannot carry
// state from one stage to the next there either.
// Expression trees match oracle/oracle_capi.cpp term by term, so the strict build
// (no FMA contraction) is bit-comparable with the CPU oracle.
#pragma once

namespace bacon {

struct RhsLorenz {  // p = (sigma, rho, beta)
    static constexpr int DIM = 3, NPARAM = 3;
    __device__ __forceinline__ void operator()(double, const double (&y)[3], const double* p, double (&dy)[3]) const {
        dy[0] = p[0] * (y[1] - y[0]);
        dy[1] = y[0] * (p[1] - y[2]) - y[1];
        dy[2] = y[0] * y[1] - p[2] * y[2];
    }
    // k = h * f: the first component is linear in sigma, so h*sigma (the same for all stages of an attempt) replaces
    // a multiplication per stage
    __device__ __forceinline__ void scaled(double h, double, const double (&y)[3], const double* p, double (&k)[3]) const {
        const double hs = h * p[0];
        k[0] = hs * (y[1] - y[0]);
        k[1] = h * (y[0] * (p[1] - y[2]) - y[1]);
        k[2] = h * (y[0] * y[1] - p[2] * y[2]);
    }
    __device__ __forceinline__ void jac(double, const double (&y)[3], const double* p, double (&J)[3][3]) const {
        J[0][0] = -p[0];       J[0][1] = p[0];  J[0][2] = 0.0;
        J[1][0] = p[1] - y[2]; J[1][1] = -1.0;  J[1][2] = -y[0];
        J[2][0] = y[1];        J[2][1] = y[0];  J[2][2] = -p[2];
    }
};

struct RhsVdp {  // Van der Pol, p = (mu)
    static constexpr int DIM = 2, NPARAM = 1;
    __device__ __forceinline__ void operator()(double, const double (&y)[2], const double* p, double (&dy)[2]) const {
        dy[0] = y[1];
        dy[1] = (p[0] * (1.0 - y[0] * y[0])) * y[1] - y[0];
    }
    __device__ __forceinline__ void jac(double, const double (&y)[2], const double* p, double (&J)[2][2]) const {
        J[0][0] = 0.0;                               J[0][1] = 1.0;
        J[1][0] = -2.0 * p[0] * y[0] * y[1] - 1.0;   J[1][1] = p[0] * (1.0 - y[0] * y[0]);
    }
};

struct RhsRobertson {  // stiff kinetics, p = (k1, k2, k3)
    static constexpr int DIM = 3, NPARAM = 3;
    __device__ __forceinline__ void operator()(double, const double (&y)[3], const double* p, double (&dy)[3]) const {
        const double a = p[0] * y[0];
        const double b = (p[2] * y[1]) * y[2];
        const double c = (p[1] * y[1]) * y[1];
        dy[0] = b - a;
        dy[1] = (a - b) - c;
        dy[2] = c;
    }
    __device__ __forceinline__ void jac(double, const double (&y)[3], const double* p, double (&J)[3][3]) const {
        const double k3y2 = p[2] * y[2], k3y1 = p[2] * y[1], k2y1 = 2.0 * p[1] * y[1];
        J[0][0] = -p[0]; J[0][1] = k3y2;           J[0][2] = k3y1;
        J[1][0] = p[0];  J[1][1] = -k3y2 - k2y1;   J[1][2] = -k3y1;
        J[2][0] = 0.0;   J[2][1] = k2y1;           J[2][2] = 0.0;
    }
};

template <int N> struct RhsLinear {  // y' = A y, p = A row-major [N][N]
    static constexpr int DIM = N, NPARAM = N * N;
    __device__ __forceinline__ void operator()(double, const double (&y)[N], const double* p, double (&dy)[N]) const {
#pragma unroll
        for (int i = 0; i < N; ++i) {
            double s = p[i * N] * y[0];
#pragma unroll
            for (int j = 1; j < N; ++j) s += p[i * N + j] * y[j];
            dy[i] = s;
        }
    }
    __device__ __forceinline__ void jac(double, const double (&)[N], const double* p, double (&J)[N][N]) const {
#pragma unroll
        for (int i = 0; i < N; ++i)
#pragma unroll
            for (int j = 0; j < N; ++j) J[i][j] = p[i * N + j];
    }
};

struct RhsExp {  // y' = y  (README.md:26-28, rk.rs:539-541, bdf.rs:769-771)
    static constexpr int DIM = 1, NPARAM = 0;
    __device__ __forceinline__ void operator()(double, const double (&y)[1], const double*, double (&dy)[1]) const { dy[0] = y[0]; }
    __device__ __forceinline__ void jac(double, const double (&)[1], const double*, double (&J)[1][1]) const { J[0][0] = 1.0; }
};
struct RhsDecay {  // y' = -y  (bdf.rs:781-783)
    static constexpr int DIM = 1, NPARAM = 0;
    __device__ __forceinline__ void operator()(double, const double (&y)[1], const double*, double (&dy)[1]) const { dy[0] = -y[0]; }
    __device__ __forceinline__ void jac(double, const double (&)[1], const double*, double (&J)[1][1]) const { J[0][0] = -1.0; }
};
struct RhsQuadratic {  // y' = -2t  (rk.rs:664-666, bdf.rs:773-775)
    static constexpr int DIM = 1, NPARAM = 0;
    __device__ __forceinline__ void operator()(double t, const double (&)[1], const double*, double (&dy)[1]) const { dy[0] = -2.0 * t; }
    __device__ __forceinline__ void jac(double, const double (&)[1], const double*, double (&J)[1][1]) const { J[0][0] = 0.0; }
};
struct RhsCos {  // y' = cos t  (rk.rs:668-670, bdf.rs:777-779)
    static constexpr int DIM = 1, NPARAM = 0;
    __device__ __forceinline__ void operator()(double t, const double (&)[1], const double*, double (&dy)[1]) const { dy[0] = cos(t); }
    __device__ __forceinline__ void jac(double, const double (&)[1], const double*, double (&J)[1][1]) const { J[0][0] = 0.0; }
};
struct RhsHarmonic {  // y'' = -w^2 y, p = (w)
    static constexpr int DIM = 2, NPARAM = 1;
    __device__ __forceinline__ void operator()(double, const double (&y)[2], const double* p, double (&dy)[2]) const {
        dy[0] = y[1];
        dy[1] = -(p[0] * p[0]) * y[0];
    }
    __device__ __forceinline__ void jac(double, const double (&)[2], const double* p, double (&J)[2][2]) const {
        J[0][0] = 0.0;               J[0][1] = 1.0;
        J[1][0] = -(p[0] * p[0]);    J[1][1] = 0.0;
    }
};

}  // namespace bacon
