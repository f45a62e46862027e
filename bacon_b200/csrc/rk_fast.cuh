// rk_fast.cuh — K1/K2: the adaptive embedded Runge-Kutta stepper of the reference
// (RungeKuttaSolver::step, src/ivp/rk.rs:361-423) driven by the IVPIterator loop
// (src/ivp.rs:220-238), one trajectory per thread, REF_CORRECTED semantics.
//
// What is kept from the reference, statement by statement:
//   * Done when t >= end (rk.rs:362-364); last step clamped: t+dt >= end -> dt = end-t (:366-368)
//   * O stage evaluations per attempt, no FSAL; k_j = f_j * dt stored per stage (:370-384)
//   * stage argument y + sum_j a_ij k_j (:371-374), error = || sum_j e_j k_j ||_2 / dt (:386-390)
//   * accept iff error <= tol; then t += dt, y += sum_j b_j k_j (:392-398)
//   * dt *= clamp(safety*(tol/error)^(1/4), 0.1, 4) on EVERY attempt (:400-408), dt = min(dt, dt_max) (:410-412)
//   * dt < dt_min && t < end -> MinimumTimeDeltaExceeded, before the point is yielded (:414-416)
//   * dt0 = (dt_max + dt_min)/2 (rk.rs:315)
// How it is laid out for the FP64 pipe of sm_100a (measured with tools/fp64_peak.cu: a DFMA whose three
// sources are three different registers issues at 2/3 rate, 24.7 vs 37.0 TFLOP/s; DMUL/DADD and DFMAs
// with a uniform-register/constant operand run at full rate):
//   * every tableau combination is a chain  acc = fma(coef, k_j, acc)  started FROM y (stage
//     argument, solution update), so the coefficient is a uniform-register operand (values in
//     __constant__ memory, loaded once outside the loop); structural zeros emit nothing
//   * the accept test compares squared quantities on the integer pipe: ||E||^2 <= (tol*dt)^2
//   * the step-size factor (tol/error)^(1/4) = ((tol*dt)^2/||E||^2)^(1/8) is evaluated on the SFU in the
//     log domain: exponents by integer arithmetic, 2x MUFU.LG2 on the mantissas, 1x MUFU.EX2.
//     Relative accuracy ~3e-7: below the ~1e-6 rounding noise of the embedded error estimate itself (a
//     cancelling sum, sum_j e_j = 0), which already makes dt differ at that level between any two
//     correct evaluations.  It only ever sets the NEXT step size; no FP64 division, sqrt or pow.
//   * the hot path holds no FP64 compare at all: "t + dt >= end" is tested as dt >= end - t on the bit
//     patterns (the two differ only when the sum rounds onto `end`, where both forms take the same step
//     h up to one ulp), and everything rare (last step, Done, reject, NaN, dt < dt_min, attempt cap) sits
//     behind two branches.  Measured with tools/fp64_mix.cu: integer/ALU instructions issued next to a
//     saturated FP64 pipe are not free (each costs about half a DFMA slot), so the loop is trimmed to
//     what the reference's step needs.
// Results agree with the oracle inside the parity band max(10*tol, 1e-12); the strict kernels
// (rk_strict.cuh) are the bit-exact form.
#pragma once
#include "ivp_common.cuh"
#include "tableaux.cuh"

namespace bacon {

// a <= b / a < b for doubles that are >= +0 (or NaN, which orders above everything): integer pipe.
// (A DSETP occupies the FP64 pipe for as long as a DFMA does.)
__device__ __forceinline__ bool pos_le(double a, double b) { return __double_as_longlong(a) <= __double_as_longlong(b); }
__device__ __forceinline__ bool pos_lt(double a, double b) { return __double_as_longlong(a) < __double_as_longlong(b); }

// clamp(safety * (num/den)^(1/8), 0.1, 4) as a double; num, den >= 0.  SFU, log domain: the exponent
// difference is integer arithmetic on the high words, the mantissas (top 20 bits rebuilt as floats in
// [1, 2)) go through MUFU.LG2, the result through MUFU.EX2.  den = 0 -> 4, den = inf -> 0.1.
// Relative accuracy ~3e-7 (the clamp bounds are the floats nearest 0.1 and 4).
__device__ __forceinline__ double step_factor(double num, double den, float safety) {
    const int hn = __double2hiint(num), hd = __double2hiint(den);
    const int de = (hn >> 20) - (hd >> 20);
    float mn = __uint_as_float(0x3f800000u | (((unsigned)hn << 3) & 0x007ffff8u));
    float md = __uint_as_float(0x3f800000u | (((unsigned)hd << 3) & 0x007ffff8u));
    asm("lg2.approx.ftz.f32 %0, %0;" : "+f"(mn));
    asm("lg2.approx.ftz.f32 %0, %0;" : "+f"(md));
    float l = 0.125f * ((float)de + (mn - md));
    asm("ex2.approx.ftz.f32 %0, %0;" : "+f"(l));
    const float d = fminf(fmaxf(safety * l, 0.1f), 4.0f);
    const unsigned fb = __float_as_uint(d);  // float -> double on the integer pipe (d is a normal float)
    return __hiloint2double((int)((fb >> 3) + (896u << 20)), (int)(fb << 29));
}

// x^(-1/8), x in [1e-6, 1e8]: SFU seed + one Newton step in fp64 (used by the warp-per-trajectory kernel)
__device__ __forceinline__ double inv_eighth_root(double x) {
    float s = (float)x;
    asm("sqrt.approx.ftz.f32 %0, %0;" : "+f"(s));
    asm("sqrt.approx.ftz.f32 %0, %0;" : "+f"(s));
    asm("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(s));
    const double z = (double)s;
    const double z2 = z * z, z4 = z2 * z2, z8 = z4 * z4;
    const double r = fma(-x, z8, 1.0);  // 1 - x z^8
    return fma(z * 0.125, r, z);        // Newton on z^-8 = x
}

// optional member of a RHS functor:  void scaled(double h, double t, const double (&y)[D], const double* p,
// double (&k)[D]) const  ->  k = h * f(t, y).  Only the fast (non-strict) RK kernels use it.
template <class Rhs, class = void> struct has_scaled { static constexpr bool value = false; };
template <class Rhs> struct has_scaled<Rhs, decltype(void(&Rhs::scaled))> { static constexpr bool value = true; };

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
// is k_j used by a later stage, the error estimate or the solution update?
template <class Tab> __host__ __device__ constexpr bool stage_used(int j) {
    for (int i = j + 1; i < Tab::O; ++i)
        if (Tab::a(i, j) != 0.0) return true;
    return Tab::e(j) != 0.0 || Tab::b(j) != 0.0;
}

// Stepper concept used by ensemble_kernel (drive.cuh):
//   D;  ctor(args);  reset(args, idx, live);  int attempt(bool& yielded)  (-1 = keep going, else a
//   bacon_status);  t, dt, n_acc, n_rej, n_rhs();  out_t()/out_y() = the yielded point;  end_y().
template <class Rhs, class Tab> struct RkFastStepper {
    static constexpr int D = Rhs::DIM;
    static constexpr int P = Rhs::NPARAM;
    static constexpr int O = Tab::O;

    // ensemble-wide constants (registers / uniform registers)
    double t_start, t_end, dt_min, dt_max, tol, dt0;
    uint32_t cap;
    // one trajectory
    double y[D], p[P > 0 ? P : 1];
    double t, dt;
    uint32_t n_acc, n_att;
    uint32_t n_rej;  // valid once attempt() has returned a status (>= 0)

    __device__ __forceinline__ explicit RkFastStepper(const bacon_launch_args& a) {
        t_start = a.cfg.t_start;
        t_end = a.cfg.t_end;
        dt_min = a.cfg.dt_min;
        dt_max = a.cfg.dt_max;
        tol = a.cfg.tol;
        dt0 = (dt_max + dt_min) * 0.5;  // rk.rs:315
        cap = (a.cfg.max_attempts == 0 || a.cfg.max_attempts > 0xFFFFFFFEull) ? 0xFFFFFFFEu
                                                                               : (uint32_t)a.cfg.max_attempts;
        t = t_start;
        dt = dt0;
        n_acc = n_rej = n_att = 0;
    }
    __device__ __forceinline__ void reset(const bacon_launch_args& a, unsigned long long idx, bool live) {
        t = t_start;
        dt = dt0;
        n_acc = n_rej = n_att = 0;
        if (live) load_problem<D, P>(a, idx, y, p);
    }
    __device__ __forceinline__ uint32_t n_rhs() const { return n_att * (uint32_t)O; }
    __device__ __forceinline__ double out_t() const { return t; }
    __device__ __forceinline__ const double (&out_y() const)[D] { return y; }
    __device__ __forceinline__ const double (&end_y() const)[D] { return y; }

    // one IVPStepper::step call (rk.rs:361-423).  Exits (return >= 0) set n_rej = attempts that were
    // rejected and reported as such; on the hot path it is implied by n_att - n_acc.
    __device__ __forceinline__ int attempt(bool& yielded) {
        const Rhs rhs{};
        yielded = false;
        // rk.rs:362-368.  rem <= 0 (Done) and rem <= dt (clamp the last step) are ONE signed compare of bit
        // patterns: dt > 0, and a negative or zero double is <= every positive one as a signed integer.
        const double rem = t_end - t;
        double h = dt;
        const bool last = __double_as_longlong(rem) <= __double_as_longlong(dt);
        if (n_att >= cap || last) {
            n_rej = n_att - n_acc;
            if (n_att >= cap) return BACON_E_MAX_ATTEMPTS;
            if (!(t < t_end)) return BACON_OK;  // rk.rs:362-364
            if (t + dt >= t_end) h = rem;       // rk.rs:366-368, the reference's own test on this rare path
        }

        // stages: k_i = h * f(t + c_i h, y + sum_j a_ij k_j)   (rk.rs:370-384)
        double k[O][D];
        static_for<0, O>([&](auto I) {
            constexpr int i = decltype(I)::value;
            double Y[D], fi[D];
#pragma unroll
            for (int d = 0; d < D; ++d) {
                double s = y[d];
                static_for<0, i>([&](auto J) {
                    constexpr int j = decltype(J)::value;
                    if constexpr (Tab::a(i, j) != 0.0) s = fma(Tab::av(i, j), k[j][d], s);
                });
                Y[d] = s;
            }
            if constexpr (has_scaled<Rhs>::value) {
                // the functor returns h * f itself (it can fold h into a linear term: Lorenz saves a DMUL per stage)
                if constexpr (i == 0) rhs.scaled(h, t, Y, p, fi);
                else rhs.scaled(h, fma(Tab::cv(i), h, t), Y, p, fi);
#pragma unroll
                for (int d = 0; d < D; ++d) k[i][d] = stage_used<Tab>(i) ? fi[d] : 0.0;
            } else {
                if constexpr (i == 0) rhs(t, Y, p, fi);
                else rhs(fma(Tab::cv(i), h, t), Y, p, fi);
#pragma unroll
                for (int d = 0; d < D; ++d) k[i][d] = stage_used<Tab>(i) ? h * fi[d] : 0.0;
            }
        });

        // embedded error, squared: q = || sum_j e_j k_j ||^2   (rk.rs:386-390 is sqrt(q)/h)
        double q = 0.0;
        constexpr int e0 = first_nz_e<Tab>();
#pragma unroll
        for (int d = 0; d < D; ++d) {
            double s = Tab::ev(e0) * k[e0][d];
            static_for<e0 + 1, O>([&](auto J) {
                constexpr int j = decltype(J)::value;
                if constexpr (Tab::e(j) != 0.0) s = fma(Tab::ev(j), k[j][d], s);
            });
            q = (d == 0) ? s * s : fma(s, s, q);
        }
        const double th = tol * h;
        const double th2 = th * th;

        n_att++;
        // error <= tol  <=>  q <= (tol h)^2   (rk.rs:392).  A NaN q has a bit pattern above every finite
        // value: it is "not accepted" and diagnosed on the reject path.
        const bool accepted = pos_le(q, th2);
        // rk.rs:400-412: (tol/error)^(1/4) = (th2/q)^(1/8)
        double dtn = h * step_factor(th2, q, (float)Tab::safety);
        dtn = pos_lt(dt_max, dtn) ? dt_max : dtn;
        dt = dtn;
        if (accepted) {
            t += h;
#pragma unroll
            for (int d = 0; d < D; ++d) {
                double s = y[d];
                static_for<0, O>([&](auto J) {
                    constexpr int j = decltype(J)::value;
                    if constexpr (Tab::b(j) != 0.0) s = fma(Tab::bv(j), k[j][d], s);
                });
                y[d] = s;
            }
        }
        if (!accepted || pos_lt(dtn, dt_min)) {  // rare
            if (!accepted && q != q) {           // the reference would Redo forever (D8)
                n_rej = n_att - n_acc - 1;
                return BACON_E_NONFINITE;
            }
            if (pos_lt(dtn, dt_min) && t < t_end) {  // rk.rs:414-416: fails before the point is yielded
                n_rej = n_att - n_acc - (accepted ? 1u : 0u);
                return BACON_E_MIN_DT_EXCEEDED;
            }
        }
        if (accepted) {
            n_acc++;
            yielded = true;
        }
        return -1;
    }
};

}  // namespace bacon
