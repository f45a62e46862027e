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

// the float 1.m built from the top 20 mantissa bits of a double's high word: two instructions (mask, shift-add).
// Inline PTX because the compiler otherwise rewrites the expression as shift / and / or.
__device__ __forceinline__ float mantissa_as_float(int hi) {
    unsigned m;
    asm("and.b32 %0, %1, 0x000fffff;\n\tmad.lo.u32 %0, %0, 8, 0x3f800000;" : "=r"(m) : "r"(hi));
    return __uint_as_float(m);
}

// The step-size factor clamp(safety * (num/den)^(1/8), 0.1, 4), num, den >= 0, on the SFU in the log domain.
// step_factor_log2: log2(safety * (num/den)^(1/8)), unclamped — the exponent difference is integer arithmetic on
// the high words, the mantissas (top 20 bits rebuilt as floats in [1, 2)) go through MUFU.LG2.  den = 0 -> large
// positive, den = inf -> large negative.  exp2_as_double: MUFU.EX2 and a float -> double conversion on the integer
// pipe (the argument is inside [-3.33, 2], so the result is a normal float).  Relative accuracy ~3e-7.
__device__ __forceinline__ float lg2_approx(float x) {
    asm("lg2.approx.ftz.f32 %0, %0;" : "+f"(x));
    return x;
}
__device__ __forceinline__ float step_factor_log2(double num, double den, float log2_safety) {
    const int hn = __double2hiint(num), hd = __double2hiint(den);
    const int de = (hn >> 20) - (hd >> 20);
    float mn = mantissa_as_float(hn), md = mantissa_as_float(hd);
    asm("lg2.approx.ftz.f32 %0, %0;" : "+f"(mn));
    asm("lg2.approx.ftz.f32 %0, %0;" : "+f"(md));
    return fmaf(0.125f, (float)de + (mn - md), log2_safety);
}
__device__ __forceinline__ double exp2_as_double(float l) {
    asm("ex2.approx.ftz.f32 %0, %0;" : "+f"(l));
    const unsigned fb = __float_as_uint(l);
    return __hiloint2double((int)((fb >> 3) + (896u << 20)), (int)(fb << 29));
}
constexpr float LOG2_TENTH = -3.3219280949f;  // log2(0.1); log2(4) = 2

// A positive double as a float scaled by a power of two chosen per ensemble (only RATIOS of such values are used):
// bits = ((hi << 3) | (lo >> 29)) - (C << 23).  The funnel shift leaves the low 9 exponent bits in the top 9 bits
// and 23 mantissa bits below; the subtraction re-biases the exponent.  Two integer instructions, no conversion on the
// FP64 pipe.  Inside the window (resulting exponent field 1..254) the result is exact up to truncation (2^-23); see
// RkFastStepper::attempt for what happens outside.
__device__ __forceinline__ float scaled_float(double x, int c_shifted) {
    return __int_as_float((int)__funnelshift_l((unsigned)__double2loint(x), (unsigned)__double2hiint(x), 3) - c_shifted);
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
//   bacon_status);  t, dt, n_rej, n_rhs();  accepted count: field n_acc, or acc_running() / acc_of(raw) + status_of(raw)
//   when attempt() returns an encoded status (StepperCodec, drive.cuh);  out_t()/out_y() = the yielded point;  end_y().
template <class Rhs, class Tab> struct RkFastStepper {
    using RhsT = Rhs;
    static constexpr int D = Rhs::DIM;
    static constexpr int P = Rhs::NPARAM;
    static constexpr int O = Tab::O;

    // ensemble-wide constants (registers / uniform registers)
    double t_start, t_end, dt_min, dt_max, tol, dt0;
    uint32_t cap;  // max_attempts
    int c_shifted;   // exponent re-bias of scaled_float() for this ensemble's (tol dt)^2 range, << 23
    int hi_dt_max;   // high word of dt_max — or INT_MIN (every attempt takes the exact path) if that range has no window
    float eighth;    // 0.125f pinned in a register (ptxas otherwise re-materialises it every attempt)
    // one trajectory
    double y[D], p[P > 0 ? P : 1];
    double t, dt;
    // Counters.  `tick` counts attempt() CALLS and is the only counter the common path touches.  Every lane of a warp
    // makes the same calls, so the lanes' ticks stay equal, and every CHECK_EVERY calls the whole warp takes the rare
    // block together, rewinds tick to 0 and returns RAW_CHECKPOINT to the driver (which uses it to notice that the
    // work counter ran dry without polling anything per attempt).  `tick0` is the trajectory's origin on that axis:
    // attempts that ran their stages = tick - tick0 (calls that ran none move tick0 along).  The same compare,
    // tick > next_check, enforces max_attempts.  Accepted = attempts - n_rej (- 1 when the trajectory failed in an
    // attempt that counts neither way: see attempt()).
    static constexpr uint32_t CHECK_EVERY = 64;
    static constexpr bool UNROLL_DRIVER = true;  // drive.cuh: two attempts per loop iteration
    uint32_t tick, tick0, next_check, n_rej;
    static constexpr int VOID_ATTEMPT = 0x100;  // ORed into the status attempt() returns: the last attempt counts neither way

    __device__ __forceinline__ explicit RkFastStepper(const bacon_launch_args& a) {
        t_start = a.cfg.t_start;
        t_end = a.cfg.t_end;
        dt_min = a.cfg.dt_min;
        dt_max = a.cfg.dt_max;
        tol = a.cfg.tol;
        dt0 = (dt_max + dt_min) * 0.5;  // rk.rs:315
        {
            // Window of scaled_float(): (tol dt_max)^2 maps to exponent field 250; (tol dt_min)^2 / 16 (a clamped last
            // step h in [dt_min/4, dt_min) still lands inside; a smaller h fails the dt_min test below whatever the
            // factor) must stay above field 24, so that everything below the window is < th2 * 2^-20, where the factor
            // is clamped to 4 anyway.  Anchoring the top keeps the 9-bit exponent field free of aliases down to 2^-1020.
            const double lo2 = (tol * dt_min) * (tol * dt_min), hi2 = (tol * dt_max) * (tol * dt_max);
            const int e_lo = (__double2hiint(lo2) >> 20) & 0x7ff, e_hi = (__double2hiint(hi2) >> 20) & 0x7ff;
            const bool ok = e_lo >= 1 && e_hi <= 1022 && e_hi - e_lo <= 250 - 28;
            c_shifted = (int)((unsigned)(e_hi - 512 - 250) << 23);
            hi_dt_max = ok ? __double2hiint(dt_max) : (int)0x80000000;
            eighth = __int_as_float(0x3e000000 + (a.cfg.dim >> 30));  // + 0, but not a constant ptxas can see
        }
        cap = (a.cfg.max_attempts == 0 || a.cfg.max_attempts > 0xFFFFFFFEull) ? 0xFFFFFFFEu
                                                                               : (uint32_t)a.cfg.max_attempts;
        tick = 0;
        t = t_start;
        dt = dt0;
        rearm(0);
    }
    // the trajectory has made n attempts so far: place it on the tick axis
    __device__ __forceinline__ void rearm(uint32_t n) {
        tick0 = tick - n;
        const uint32_t left = cap - n, room = CHECK_EVERY - tick;
        next_check = tick + (left < room ? left : room);
    }
    // the next checkpoint in at most k calls (the driver shortens the spacing at the end of the ensemble: drive.cuh)
    __device__ __forceinline__ void hurry(uint32_t k) { next_check = next_check < tick + k ? next_check : tick + k; }
    __device__ __forceinline__ uint32_t n_att() const { return tick - tick0; }
    __device__ __forceinline__ void reset(const bacon_launch_args& a, unsigned long long idx, bool live) {
        t = t_start;
        dt = dt0;
        n_rej = 0;
        rearm(0);
        if (live) load_problem<D, P>(a, idx, y, p);
    }
    // restart record (bacon_ivp_options), applied by the kernels compiled with the optional inputs (drive.cuh: EVENT)
    __device__ __forceinline__ void apply_restart(const bacon_launch_args& a, unsigned long long idx) {
        trajectory_start(a, idx, dt_min, dt_max, t, dt);
    }
    __device__ __forceinline__ uint32_t n_rhs() const { return n_att() * (uint32_t)O; }
    __device__ __forceinline__ double out_t() const { return t; }
    __device__ __forceinline__ const double (&out_y() const)[D] { return y; }
    __device__ __forceinline__ const double (&end_y() const)[D] { return y; }

    // one IVPStepper::step call (rk.rs:361-423).  The common case — not the last step, accepted, new dt strictly
    // inside (dt_min, dt_max) — is proven by 32-bit compares of the HIGH words (a positive double's high word is
    // monotone in its value, so hi(a) < hi(b) proves a < b); everything else, including every case the high words
    // cannot decide, goes to one of two rare blocks that run the reference's own tests exactly.  Exits (return >= 0)
    // return the bacon_status, ORed with VOID_ATTEMPT when the failing attempt is reported neither as accepted nor as
    // rejected (status_of / acc_of decode it for the driver).
    __device__ __forceinline__ int attempt(bool& yielded) {
        const Rhs rhs{};
        // rk.rs:362-368.  rem <= 0 (Done) and rem <= dt (clamp the last step) both imply hi(rem) <= hi(dt) as signed
        // integers (dt > 0).
        const double rem = t_end - t;
        double h = dt;
        if (__builtin_expect((tick >= next_check) | (__double2hiint(rem) <= __double2hiint(dt)), 0)) {  // rare
            if (tick >= next_check) {
                if (n_att() >= cap) {
                    tick++, tick0++;  // a call that runs no stages
                    return BACON_E_MAX_ATTEMPTS;
                }
                const uint32_t n = n_att();  // a checkpoint: the whole warp is here; the tick axis restarts
                tick = 0;
                rearm(n);
                return -2;  // RAW_CHECKPOINT
            }
            if (!(t < t_end)) {  // rk.rs:362-364
                tick++, tick0++;
                return BACON_OK;
            }
            if (t + dt >= t_end) h = rem;  // rk.rs:366-368, the reference's own test
        }
        tick++;

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
                    if constexpr (Tab::a(i, j) != 0.0) s = fma(coef_a<i, j>(), k[j][d], s);
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
        // rk.rs:400-408: (tol/error)^(1/4) = (th2/q)^(1/8), on the SFU in the log domain.  An accepted step has a factor
        // >= safety > 0.1, so the lower clamp, like the clamp to dt_max (:410-412), is left to the rare block.
        // q and th2 are turned into floats by scaled_float().  th2 is always inside the window (ctor).  q below the
        // window (or +0) gives NaN or -inf from LG2, hence lf = NaN or +inf, and fminf(., 2) = 2: the factor 4 that such
        // a q deserves.  q above the window is > th2: the rare block recomputes the factor from the doubles.
        const float lf = fmaf(eighth, lg2_approx(scaled_float(th2, c_shifted)) - lg2_approx(scaled_float(q, c_shifted)),
                              Tab::log2_safety);
        double dtn = h * exp2_as_double(fminf(lf, 2.0f));

        // error <= tol  <=>  q <= (tol h)^2   (rk.rs:392).  q >= +0 or NaN (either sign): as UNSIGNED integers a NaN
        // orders above every finite value, so it is never "accepted".
        const unsigned hq = (unsigned)__double2hiint(q), hth = (unsigned)__double2hiint(th2);
        const int hdn = __double2hiint(dtn);
        if (__builtin_expect((hq >= hth) | (hdn >= hi_dt_max) | (hdn <= __double2hiint(dt_min)), 0)) {
            // rare: the exact tests
            const bool accepted =
                (unsigned long long)__double_as_longlong(q) <= (unsigned long long)__double_as_longlong(th2);
            dtn = h * exp2_as_double(fminf(fmaxf(step_factor_log2(th2, q, Tab::log2_safety), LOG2_TENTH), 2.0f));
            dtn = pos_lt(dt_max, dtn) ? dt_max : dtn;  // rk.rs:410-412
            dt = dtn;
            if (!accepted) {
                if (q != q) return BACON_E_NONFINITE | VOID_ATTEMPT;  // the reference would Redo forever (D8)
                n_rej++;
                if (pos_lt(dtn, dt_min) && t < t_end) return BACON_E_MIN_DT_EXCEEDED;  // rk.rs:414-416
                return -1;
            }
            advance(h, k);
            // rk.rs:414-416: fails before the point is yielded (the attempt is counted neither way)
            if (pos_lt(dtn, dt_min) && t < t_end) return BACON_E_MIN_DT_EXCEEDED | VOID_ATTEMPT;
            yielded = true;
            return -1;
        }
        dt = dtn;
        advance(h, k);
        yielded = true;
        return -1;
    }

    // suspend / resume (end-of-ensemble regrouping, drive.cuh): everything that belongs to the trajectory
    static constexpr int STATE_DOUBLES = D + (P > 0 ? P : 0) + 3;
    __device__ __forceinline__ double remaining() const { return t_end - t; }
    __device__ __forceinline__ void save(double (&st)[STATE_DOUBLES + 1]) const {
#pragma unroll
        for (int d = 0; d < D; ++d) st[d] = y[d];
#pragma unroll
        for (int k = 0; k < P; ++k) st[D + k] = p[k];
        st[D + P] = t;
        st[D + P + 1] = dt;
        st[D + P + 2] = __hiloint2double((int)n_att(), (int)n_rej);
    }
    __device__ __forceinline__ void load(const double (&st)[STATE_DOUBLES + 1]) {
#pragma unroll
        for (int d = 0; d < D; ++d) y[d] = st[d];
#pragma unroll
        for (int k = 0; k < P; ++k) p[k] = st[D + k];
        t = st[D + P];
        dt = st[D + P + 1];
        n_rej = (uint32_t)__double2loint(st[D + P + 2]);
        tick = 0;  // the lanes of the new warp start a common tick axis
        rearm((uint32_t)__double2hiint(st[D + P + 2]));
    }
    __device__ __forceinline__ static int status_of(int raw) { return raw & (VOID_ATTEMPT - 1); }
    __device__ __forceinline__ uint32_t acc_of(int raw) const { return n_att() - n_rej - ((raw & VOID_ATTEMPT) ? 1u : 0u); }
    __device__ __forceinline__ uint32_t acc_running() const { return n_att() - n_rej; }

    // rk.rs:393-398: t += dt, y += sum_j b_j k_j
    __device__ __forceinline__ void advance(double h, const double (&k)[O][D]) {
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

    // a_ij as a DFMA operand: entries whose low word is zero (1/4, 2, -8, 3/4 ...) are encodable as an immediate and
    // stay literals; the others are read from __constant__ memory into uniform registers once, outside the loop
    // (RKF45 has 24 distinct coefficients: with all of them in uniform registers ptxas runs out and re-loads one
    // per attempt).
    template <int I, int J> __device__ __forceinline__ static double coef_a() {
        constexpr double v = Tab::a(I, J);
        constexpr double s = v * 1048576.0;
        if constexpr (s == (double)(long long)s && (v >= 1.0 / 1024 || v <= -1.0 / 1024) && v <= 1024.0 && v >= -1024.0) return v;
        else return Tab::av(I, J);
    }
};

}  // namespace bacon
