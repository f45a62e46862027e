// rk_fast.cuh — K1/K2: the adaptive embedded Runge-Kutta stepper of the reference
// (RungeKuttaSolver::step, src/ivp/rk.rs:361-423) driven by the IVPIterator loop
// (src/ivp.rs:220-238), one trajectory per thread, REF_CORRECTED semantics.
//
// What is kept from the reference, statement by statement:
//   * Done when t >= end (rk.rs:362-364); last step clamped: t+dt >= end -> dt = end-t (:366-368)
//   * O stage evaluations per attempt, no FSAL (:370-384)
//   * error = || sum_j e_j k_j ||_2 / dt, absolute, Euclidean (:386-390)
//   * accept iff error <= tol; then t += dt, y += sum_j b_j k_j (:392-398)
//   * dt *= clamp(safety*(tol/error)^(1/4), 0.1, 4) on EVERY attempt (:400-408), dt = min(dt, dt_max) (:410-412)
//   * dt < dt_min && t < end -> MinimumTimeDeltaExceeded, before the point is yielded (:414-416)
//   * dt0 = (dt_max + dt_min)/2 (rk.rs:315)
// What is re-expressed for the FP64 pipe (results agree with the oracle to rounding,
// inside the parity band max(10*tol, 1e-12); the strict kernel is the bit-exact one):
//   * stages keep the unscaled derivative f_j (k_j = dt*f_j): Y_i = fma(dt, sum_j a_ij f_j, y),
//     so error = ||sum_j e_j f_j|| needs neither the division by dt nor O*D multiplies by dt
//   * accept test on the squared norm; (tol/error)^(1/4) = (error^2/tol^2)^(-1/8) from an SFU
//     seed (MUFU sqrt, sqrt, rsqrt in fp32) + one Newton step in fp64 (|rel err| < 1e-12)
//   * tableau coefficients are compile-time constants: zeros emit nothing, the rest are
//     constant-bank operands of DFMA
#pragma once
#include "ivp_common.cuh"
#include "tableaux.cuh"

namespace bacon {

// x^(-1/8), x in [1e-6, 1e8]
__device__ __forceinline__ double inv_eighth_root(double x) {
    float s = (float)x;
    asm("sqrt.approx.ftz.f32 %0, %0;" : "+f"(s));
    asm("sqrt.approx.ftz.f32 %0, %0;" : "+f"(s));
    asm("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(s));
    double z = (double)s;
    const double z2 = z * z, z4 = z2 * z2, z8 = z4 * z4;
    const double r = fma(-x, z8, 1.0);  // 1 - x z^8
    return fma(z * 0.125, r, z);        // Newton on z^-8 = x
}

// a <= b / a < b for doubles that are >= +0 (or NaN, which orders above everything): integer pipe
__device__ __forceinline__ bool pos_le(double a, double b) { return __double_as_longlong(a) <= __double_as_longlong(b); }
__device__ __forceinline__ bool pos_lt(double a, double b) { return __double_as_longlong(a) < __double_as_longlong(b); }

template <class Tab, int I> __host__ __device__ constexpr int first_nz_a() {
    for (int j = 0; j < I; ++j)
        if (Tab::a(I, j) != 0.0) return j;
    return -1;
}
template <class Tab> __host__ __device__ constexpr int first_nz_b() {
    for (int j = 0; j < Tab::O; ++j)
        if (Tab::b(j) != 0.0) return j;
    return -1;
}
template <class Tab> __host__ __device__ constexpr int first_nz_e() {
    for (int j = 0; j < Tab::O; ++j)
        if (Tab::e(j) != 0.0) return j;
    return -1;
}

// Stepper concept used by ensemble_kernel (drive.cuh):
//   D;  ctor(args);  reset(args, idx, live);  int attempt(bool& yielded)  (-1 = keep going, else a
//   bacon_status);  t, dt, n_acc, n_rej, n_rhs();  out_t()/out_y() = the yielded point;  end_y().
template <class Rhs, class Tab> struct RkFastStepper {
    static constexpr int D = Rhs::DIM;
    static constexpr int P = Rhs::NPARAM;
    static constexpr int O = Tab::O;

    // ensemble-wide constants (registers / uniform registers)
    double t_start, t_end, dt_min, dt_max, tol2, inv_tol2, dt0;
    uint32_t cap;
    // one trajectory
    double y[D], p[P > 0 ? P : 1];
    double t, dt;
    uint32_t n_acc, n_rej, n_att;
    bool clamped;  // the previous attempt used dt = t_end - t

    __device__ __forceinline__ explicit RkFastStepper(const bacon_launch_args& a) {
        t_start = a.cfg.t_start;
        t_end = a.cfg.t_end;
        dt_min = a.cfg.dt_min;
        dt_max = a.cfg.dt_max;
        tol2 = a.cfg.tol * a.cfg.tol;
        inv_tol2 = 1.0 / tol2;
        dt0 = (dt_max + dt_min) * 0.5;  // rk.rs:315
        cap = (a.cfg.max_attempts == 0 || a.cfg.max_attempts > 0xFFFFFFFEull) ? 0xFFFFFFFEu
                                                                               : (uint32_t)a.cfg.max_attempts;
        t = t_start;
        dt = dt0;
        n_acc = n_rej = n_att = 0;
        clamped = false;
    }
    __device__ __forceinline__ void reset(const bacon_launch_args& a, unsigned long long idx, bool live) {
        t = t_start;
        dt = dt0;
        n_acc = n_rej = n_att = 0;
        clamped = false;
        if (live) load_problem<D, P>(a, idx, y, p);
    }
    __device__ __forceinline__ uint32_t n_rhs() const { return n_att * (uint32_t)O; }
    __device__ __forceinline__ double out_t() const { return t; }
    __device__ __forceinline__ const double (&out_y() const)[D] { return y; }
    __device__ __forceinline__ const double (&end_y() const)[D] { return y; }

    // one IVPStepper::step call (rk.rs:361-423)
    __device__ __forceinline__ int attempt(bool& yielded) {
        const Rhs rhs{};
        yielded = false;
        if (n_att >= cap) return BACON_E_MAX_ATTEMPTS;
        // rk.rs:362-364.  t can only have reached t_end in an attempt whose step was clamped to land on it
        // (otherwise t_new = t + dt < t_end was just tested), so the FP64 compare is skipped otherwise.
        if (clamped && t >= t_end) return BACON_OK;
        double h = dt;
        clamped = t + h >= t_end;
        if (clamped) h = t_end - t;  // rk.rs:366-368

        double f[O][D];
        rhs(t, y, p, f[0]);
        static_for<1, O>([&](auto I) {
            constexpr int i = decltype(I)::value;
            constexpr int j0 = first_nz_a<Tab, i>();
            double Y[D];
#pragma unroll
            for (int d = 0; d < D; ++d) {
                double s = Tab::av(i, j0) * f[j0][d];
                static_for<j0 + 1, i>([&](auto J) {
                    constexpr int j = decltype(J)::value;
                    if constexpr (Tab::a(i, j) != 0.0) s = fma(Tab::av(i, j), f[j][d], s);
                });
                Y[d] = fma(h, s, y[d]);
            }
            rhs(fma(Tab::cv(i), h, t), Y, p, f[i]);
        });

        // embedded error, squared: q = || sum_j e_j f_j ||^2   ( = (||sum_j e_j k_j|| / dt)^2 )
        double q = 0.0;
        constexpr int e0 = first_nz_e<Tab>();
#pragma unroll
        for (int d = 0; d < D; ++d) {
            double s = Tab::ev(e0) * f[e0][d];
            static_for<e0 + 1, O>([&](auto J) {
                constexpr int j = decltype(J)::value;
                if constexpr (Tab::e(j) != 0.0) s = fma(Tab::ev(j), f[j][d], s);
            });
            q = (d == 0) ? s * s : fma(s, s, q);
        }

        n_att++;
        // Comparisons of non-negative doubles are done on their bit patterns (integer pipe): the FP64
        // pipe is the bound of this kernel and a DSETP costs it as much as a DFMA.  A NaN q has a bit
        // pattern above every finite value, so it is "not accepted" and is diagnosed on that rare path.
        const bool accepted = pos_le(q, tol2);  // rk.rs:392
        if (!accepted && q != q) return BACON_E_NONFINITE;  // the reference would Redo forever (D8)
        if (accepted) {
            t += h;
            constexpr int b0 = first_nz_b<Tab>();
#pragma unroll
            for (int d = 0; d < D; ++d) {
                double s = Tab::bv(b0) * f[b0][d];
                static_for<b0 + 1, O>([&](auto J) {
                    constexpr int j = decltype(J)::value;
                    if constexpr (Tab::b(j) != 0.0) s = fma(Tab::bv(j), f[j][d], s);
                });
                y[d] = fma(h, s, y[d]);
            }
        }
        // rk.rs:400-412; outside [1e-6, 1e8] the clamp of delta to [0.1, 4] decides anyway
        double x = q * inv_tol2;
        x = pos_lt(x, 1e-6) ? 1e-6 : x;
        x = pos_lt(1e8, x) ? 1e8 : x;
        double delta = Tab::safety * inv_eighth_root(x);
        delta = pos_lt(delta, 0.1) ? 0.1 : delta;
        delta = pos_lt(4.0, delta) ? 4.0 : delta;
        dt = h * delta;
        dt = pos_lt(dt_max, dt) ? dt_max : dt;
        if (pos_lt(dt, dt_min)) {
            if (t < t_end) {  // rk.rs:414-416 (fails before the point is yielded)
                if (!accepted) n_rej++;
                return BACON_E_MIN_DT_EXCEEDED;
            }
        }
        if (accepted) {
            n_acc++;
            yielded = true;
        } else {
            n_rej++;
        }
        return -1;
    }
};

}  // namespace bacon
