// tableaux.cuh — embedded Runge-Kutta pairs and BDF coefficients of the reference,
// REF_CORRECTED semantics (D1 matrix orientation, D2 1859/4104,
// D3 safety 84/100 repaired; nothing else changed).
//
// The fast kernels take these as compile-time constants: after full unrolling
// every coefficient is an immediate / constant-bank operand of a DFMA and the
// structural zeros of the Butcher matrix generate no instruction at all.  The
// strict kernels read a runtime copy (either semantics) from __constant__ memory.
#pragma once

namespace bacon {

// Device-side copies in __constant__ memory.  The fast kernels know the STRUCTURE (which entries are
// zero) at compile time from the constexpr accessors below, and read the VALUES as constant-bank
// operands of DFMA/DMUL (c[0x3][..]) through cv/av/bv/ev: no immediates to materialise per use.
#ifdef __CUDACC__
static __constant__ double kRKF45_c[6] = {0.0, 1.0 / 4.0, 3.0 / 8.0, 12.0 / 13.0, 1.0, 1.0 / 2.0};
static __constant__ double kRKF45_a[6][6] = {
    {0, 0, 0, 0, 0, 0},
    {1.0 / 4.0, 0, 0, 0, 0, 0},
    {3.0 / 32.0, 9.0 / 32.0, 0, 0, 0, 0},
    {1932.0 / 2197.0, -7200.0 / 2197.0, 7296.0 / 2197.0, 0, 0, 0},
    {439.0 / 216.0, -8.0, 3680.0 / 513.0, -845.0 / 4104.0, 0, 0},
    {-8.0 / 27.0, 2.0, -3544.0 / 2565.0, 1859.0 / 4104.0, -11.0 / 40.0, 0}};
static __constant__ double kRKF45_b[6] = {25.0 / 216.0, 0.0, 1408.0 / 2565.0, 2197.0 / 4104.0, -(1.0 / 5.0), 0.0};
static __constant__ double kRKF45_e[6] = {1.0 / 360.0, 0.0, -128.0 / 4275.0, -2197.0 / 75240.0, 1.0 / 50.0, 2.0 / 55.0};
static __constant__ double kBS23_c[4] = {0.0, 1.0 / 2.0, 3.0 / 4.0, 1.0};
static __constant__ double kBS23_a[4][4] = {{0, 0, 0, 0}, {1.0 / 2.0, 0, 0, 0}, {0, 3.0 / 4.0, 0, 0},
                                            {2.0 / 9.0, 1.0 / 3.0, 4.0 / 9.0, 0}};
static __constant__ double kBS23_b[4] = {2.0 / 9.0, 1.0 / 3.0, 4.0 / 9.0, 0.0};
static __constant__ double kBS23_e[4] = {-5.0 / 72.0, 1.0 / 12.0, 1.0 / 9.0, -(1.0 / 8.0)};
#endif

// Runge-Kutta-Fehlberg 4(5): src/ivp/rk.rs:430-526
//   c  = t_coefficients   rk.rs:443-450
//   a  = k_coefficients   rk.rs:459-502 (rows as the source comments label them)
//   b  = avg_coefficients rk.rs:506-513 (4th-order weights)
//   e  = error_coefficients rk.rs:517-524
struct TabRKF45 {
    static constexpr int O = 6;
    __host__ __device__ static constexpr double c(int i) {
        constexpr double v[6] = {0.0, 1.0 / 4.0, 3.0 / 8.0, 12.0 / 13.0, 1.0, 1.0 / 2.0};
        return v[i];
    }
    __host__ __device__ static constexpr double a(int i, int j) {
        constexpr double v[6][6] = {
            {0, 0, 0, 0, 0, 0},
            {1.0 / 4.0, 0, 0, 0, 0, 0},
            {3.0 / 32.0, 9.0 / 32.0, 0, 0, 0, 0},
            {1932.0 / 2197.0, -7200.0 / 2197.0, 7296.0 / 2197.0, 0, 0, 0},
            {439.0 / 216.0, -8.0, 3680.0 / 513.0, -845.0 / 4104.0, 0, 0},
            {-8.0 / 27.0, 2.0, -3544.0 / 2565.0, 1859.0 / 4104.0, -11.0 / 40.0, 0}};
        return v[i][j];
    }
    __host__ __device__ static constexpr double b(int i) {
        constexpr double v[6] = {25.0 / 216.0, 0.0, 1408.0 / 2565.0, 2197.0 / 4104.0, -(1.0 / 5.0), 0.0};
        return v[i];
    }
    __host__ __device__ static constexpr double e(int i) {
        constexpr double v[6] = {1.0 / 360.0, 0.0, -128.0 / 4275.0, -2197.0 / 75240.0, 1.0 / 50.0, 2.0 / 55.0};
        return v[i];
    }
    static constexpr double safety = 84.0 / 100.0;  // rk.rs:266-268 (intent)
    static constexpr float log2_safety = -0.2515387670f;  // log2(0.84)
#ifdef __CUDACC__
    __device__ __forceinline__ static double cv(int i) { return kRKF45_c[i]; }
    __device__ __forceinline__ static double av(int i, int j) { return kRKF45_a[i][j]; }
    __device__ __forceinline__ static double bv(int i) { return kRKF45_b[i]; }
    __device__ __forceinline__ static double ev(int i) { return kRKF45_e[i]; }
#endif
};

// Bogacki-Shampine 3(2): src/ivp/rk.rs:563-621 ("the second adaptive RK").
// FSAL is not exploited by the reference (4 evaluations per attempt) and the
// controller exponent stays 1/4 (rk.rs:401); both are kept.
struct TabBS23 {
    static constexpr int O = 4;
    __host__ __device__ static constexpr double c(int i) {
        constexpr double v[4] = {0.0, 1.0 / 2.0, 3.0 / 4.0, 1.0};
        return v[i];
    }
    __host__ __device__ static constexpr double a(int i, int j) {
        constexpr double v[4][4] = {{0, 0, 0, 0}, {1.0 / 2.0, 0, 0, 0}, {0, 3.0 / 4.0, 0, 0},
                                    {2.0 / 9.0, 1.0 / 3.0, 4.0 / 9.0, 0}};
        return v[i][j];
    }
    __host__ __device__ static constexpr double b(int i) {
        constexpr double v[4] = {2.0 / 9.0, 1.0 / 3.0, 4.0 / 9.0, 0.0};
        return v[i];
    }
    __host__ __device__ static constexpr double e(int i) {
        constexpr double v[4] = {-5.0 / 72.0, 1.0 / 12.0, 1.0 / 9.0, -(1.0 / 8.0)};
        return v[i];
    }
    static constexpr double safety = 84.0 / 100.0;
    static constexpr float log2_safety = -0.2515387670f;  // log2(0.84)
#ifdef __CUDACC__
    __device__ __forceinline__ static double cv(int i) { return kBS23_c[i]; }
    __device__ __forceinline__ static double av(int i, int j) { return kBS23_a[i][j]; }
    __device__ __forceinline__ static double bv(int i) { return kBS23_b[i]; }
    __device__ __forceinline__ static double ev(int i) { return kBS23_e[i]; }
#endif
};

// Runtime tableau for the strict (oracle-order) kernels.  `a` is what the
// stepper's row_iter() sees: REF_LITERAL fills it column-major exactly like
// BSMatrix::from_vec (rk.rs:459) so stage i reads A[j][i] of the listed rows.
struct RkTableauRt {
    double c[6];
    double a[6][6];
    double b[6];
    double e[6];
    double safety;
};

// which of the strict kernels' four constant tableaux (rk_strict.cuh: c_rk_tabs) a (method, semantics) pair uses
__host__ __device__ inline int rk_tab_slot(int order, int semantics) {
    return (order == 6 ? 0 : 2) + (semantics == BACON_SEM_LITERAL ? 1 : 0);
}

template <class Tab> inline void fill_runtime_tableau(RkTableauRt& T, bool literal) {
    constexpr int O = Tab::O;
    for (int i = 0; i < 6; ++i) {
        T.c[i] = T.b[i] = T.e[i] = 0.0;
        for (int j = 0; j < 6; ++j) T.a[i][j] = 0.0;
    }
    double listed[6][6] = {};
    for (int i = 0; i < O; ++i) {
        T.c[i] = Tab::c(i);
        T.b[i] = Tab::b(i);
        T.e[i] = Tab::e(i);
        for (int j = 0; j < O; ++j) listed[i][j] = Tab::a(i, j);
    }
    if (literal && O == 6) listed[5][3] = 1859.0 / 4014.0;  // rk.rs:499 as written (D2)
    // from_vec consumes the listed numbers in order and fills column by column (D1)
    for (int r = 0; r < O; ++r)
        for (int cc = 0; cc < O; ++cc) {
            if (literal) {
                const int flat = cc * O + r;  // element (r,cc) = listed_flat[cc*O + r]
                T.a[r][cc] = listed[flat / O][flat % O];
            } else {
                T.a[r][cc] = listed[r][cc];
            }
        }
    T.safety = literal ? 100.0 / 100.0 : 84.0 / 100.0;  // rk.rs:266-268 (D3)
}

// BDF coefficients: src/ivp/bdf.rs:641-673 (BDF6 / BDF5), :708-730 (BDF2 / BDF1).
// Element 0 multiplies dt*f(t,y); elements 1.. multiply y_n, y_{n-1}, ...
struct CoefBDF6 {
    static constexpr int O = 7;
    __host__ __device__ static constexpr double higher(int i) {
        constexpr double v[7] = {60.0 / 147.0, -360.0 / 147.0, 450.0 / 147.0, -400.0 / 147.0,
                                 225.0 / 147.0, -72.0 / 147.0, 10.0 / 147.0};
        return v[i];
    }
    __host__ __device__ static constexpr double lower(int i) {
        constexpr double v[7] = {60.0 / 137.0, -300.0 / 137.0, 300.0 / 137.0, -200.0 / 137.0,
                                 75.0 / 137.0, -12.0 / 137.0, 0.0};
        return v[i];
    }
};
struct CoefBDF2 {
    static constexpr int O = 3;
    __host__ __device__ static constexpr double higher(int i) {
        constexpr double v[3] = {2.0 / 3.0, -4.0 / 3.0, 1.0 / 3.0};
        return v[i];
    }
    __host__ __device__ static constexpr double lower(int i) {
        constexpr double v[3] = {1.0, -1.0, 0.0};
        return v[i];
    }
};

}  // namespace bacon
