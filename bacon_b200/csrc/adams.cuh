// adams.cuh — the reference's Adams predictor-corrector stepper (AdamsSolver, src/ivp/adams.rs:72-122
// fields, :340-394 RK4 start-up, :398-566 step state machine) and its fixed-step Euler stepper
// (EulerSolver, src/ivp.rs:306-344), one trajectory per thread.  SURVEY.md §8f rows N1 and N3: the
// callers either side of the RK/BDF path (README.md:45-47 `solve_ivp` = Adams5 -> RK45 -> BDF6).
//
// Adams, kept as structure exactly like the BDF path: O-1 explicit RK4 warm-up steps after EVERY dt
// change (adams.rs:452-463, :528-529, :557-558), the speculative first predictor-corrector step after a
// warm-up (yield_memory == O) with rollback by dt*(O-1) on rejection (:538-542), warm-up points yielded
// one per step() call (:415-427), error = (19/270)/dt * ||corrector - predictor||_2 (:485-486), step-size
// factor q = (tol / (2 error))^(1/O) clamped to [0.1, 4], applied on a reject and on an accept whose
// error < tol/10 only (:511-529, :544-551).
//
// One defect on this path, found while restating it (D10, not in SURVEY.md's table): when the speculative
// step is accepted (adams.rs:498-501) the early `Redo` return skips `prev_derivatives.push_back(implicit
// derivative)`, although the branch that later yields that step (:429-440) says "the derivatives memory
// deque already has the derivatives for this step".  As written, the first regular step after EVERY warm-up
// extrapolates with derivatives that lag one step behind: its error estimate is O(dt), it is rejected, dt
// shrinks, a new warm-up starts — the step count grows like 1/tol (2.3e6 steps for one oscillator period at
// tol 1e-6).  REF_LITERAL keeps that; REF_CORRECTED pushes the derivative.  Both run the same kernel.
//
// STRICT = the oracle's operation order, every product/sum individually rounded; x^(1/O) is then the
// deterministic Newton root both sides share (oracle PowMode::DetRoot; f64::powf itself is libm and not
// reproducible bit for bit on a GPU).  Non-strict = FMA contraction and an SFU log-domain root
// (relative accuracy ~1e-7; it only ever scales the NEXT step size).
#pragma once
#include "bdf.cuh"
#include "ivp_common.cuh"

namespace bacon {

// Adams-Bashforth / Adams-Moulton weights (adams.rs:635-650 Adams5, :695-712 Adams3); error
// coefficient 19/270 for both (:652-654, :714-716).
struct CoefAdams5 {
    static constexpr int O = 5;
    __host__ __device__ static constexpr double predictor(int i) {
        constexpr double v[5] = {55.0 / 24.0, -59.0 / 24.0, 37.0 / 24.0, -9.0 / 24.0, 0.0};
        return v[i];
    }
    __host__ __device__ static constexpr double corrector(int i) {
        constexpr double v[5] = {251.0 / 720.0, 646.0 / 720.0, -264.0 / 720.0, 106.0 / 720.0, -19.0 / 720.0};
        return v[i];
    }
    static constexpr double error = 19.0 / 270.0;
};
struct CoefAdams3 {
    static constexpr int O = 3;
    __host__ __device__ static constexpr double predictor(int i) {
        constexpr double v[3] = {1.0 + 1.0 / 2.0, -(1.0 / 2.0), 0.0};
        return v[i];
    }
    __host__ __device__ static constexpr double corrector(int i) {
        constexpr double v[3] = {5.0 / 12.0, 2.0 / 3.0, -(1.0 / 12.0)};
        return v[i];
    }
    static constexpr double error = 19.0 / 270.0;
};

// x^(1/N) with only correctly rounded +, *, / and integer arithmetic on the bit pattern: the same bits on
// the CPU oracle (det_root in oracle/bacon_oracle.hpp, built -ffp-contract=off) and here.
template <int N> __device__ __forceinline__ double det_root(double x) {
    if (!(x > 0.0)) return x;                        // 0 -> 0 (NaN is diagnosed before this is called)
    if (x > 1.7976931348623157e308) return x;        // +inf -> +inf
    const long long one = 0x3ff0000000000000ll;
    const long long bits = __double_as_longlong(x);
    double z = __longlong_as_double(one + (bits - one) / N);
#pragma unroll 1
    for (int it = 0; it < 8; ++it) {
        double zn1 = z;
#pragma unroll
        for (int k = 2; k < N; ++k) zn1 = __dmul_rn(zn1, z);
        z = __ddiv_rn(__dadd_rn(__dmul_rn((double)(N - 1), z), __ddiv_rn(x, zn1)), (double)N);
    }
    return z;
}

// x^(1/N) on the SFU: exponent by integer arithmetic, MUFU.LG2 on the mantissa, MUFU.EX2; x > 0 finite.
template <int N> __device__ __forceinline__ double sfu_root(double x) {
    if (!(x > 0.0) || x > 1.7976931348623157e308) return x;
    const int hi = __double2hiint(x);
    const unsigned lo = (unsigned)__double2loint(x);
    float m = __uint_as_float(0x3f800000u | (((unsigned)hi & 0xfffffu) << 3) | (lo >> 29));
    asm("lg2.approx.ftz.f32 %0, %0;" : "+f"(m));
    const float l = ((float)((hi >> 20) - 1023) + m) * (1.0f / (float)N);
    // split l into integer and fraction so the result keeps full range: 2^l = 2^floor(l) * 2^frac
    const float fl = floorf(l);
    float fr = l - fl;
    asm("ex2.approx.ftz.f32 %0, %0;" : "+f"(fr));
    return ldexp((double)fr, (int)fl);
}

template <class Rhs, class Coef, bool STRICT> struct AdamsStepper {
    using RhsT = Rhs;
    static constexpr int D = Rhs::DIM;
    static constexpr int P = Rhs::NPARAM;
    static constexpr int O = Coef::O;
    static constexpr int H = O - 1;  // entries the two deques hold once filled
    using A = Ar<STRICT>;

    double t_start, t_end, dt_min, dt_max, tol, dt0, order;
    uint32_t cap;
    // one trajectory (adams.rs:72-122)
    double y[D], p[P > 0 ? P : 1];
    double hy[H][D], ht[H];  // prev_values, oldest first
    double hf[H][D];         // prev_derivatives, oldest first
    double save[D];
    double oy[D], ot;        // the point of the last Ok(...)
    double t, dt;
    bool have;               // the deques are non-empty (they hold 0 or O-1 entries)
    bool fix_d10;            // REF_CORRECTED
    uint32_t ym;             // yield_memory (adams.rs:119)
    uint32_t n_acc, n_rej, n_att, n_f;

    __device__ __forceinline__ explicit AdamsStepper(const bacon_launch_args& a) {
        t_start = a.cfg.t_start;
        t_end = a.cfg.t_end;
        dt_min = a.cfg.dt_min;
        dt_max = a.cfg.dt_max;
        tol = a.cfg.tol;
        dt0 = A::mul(A::add(dt_max, dt_min), 0.5);  // adams.rs:297
        order = (double)O;                           // adams.rs:288
        fix_d10 = a.cfg.semantics == BACON_SEM_CORRECTED;
        cap = (a.cfg.max_attempts == 0 || a.cfg.max_attempts > 0xFFFFFFFEull) ? 0xFFFFFFFEu
                                                                               : (uint32_t)a.cfg.max_attempts;
        reset_scalars();
    }
    __device__ __forceinline__ void reset_scalars() {
        t = t_start;
        dt = dt0;
        have = false;
        ym = 0;
        n_acc = n_rej = n_att = n_f = 0;
        ot = t_start;
#pragma unroll
        for (int d = 0; d < D; ++d) { save[d] = 0.0; oy[d] = 0.0; }
#pragma unroll
        for (int k = 0; k < H; ++k) {
            ht[k] = 0.0;
#pragma unroll
            for (int d = 0; d < D; ++d) { hy[k][d] = 0.0; hf[k][d] = 0.0; }
        }
    }
    __device__ __forceinline__ void reset(const bacon_launch_args& a, unsigned long long idx, bool live) {
        reset_scalars();
        if (live) load_problem<D, P>(a, idx, y, p);
    }
    __device__ __forceinline__ void apply_restart(const bacon_launch_args& a, unsigned long long idx) {
        trajectory_start(a, idx, dt_min, dt_max, t, dt);
        ot = t;
    }
    __device__ __forceinline__ uint32_t n_rhs() const { return n_f; }
    __device__ __forceinline__ double out_t() const { return ot; }
    __device__ __forceinline__ const double (&out_y() const)[D] { return oy; }
    __device__ __forceinline__ const double (&end_y() const)[D] { return y; }

    __device__ __forceinline__ void f(double tt, const double (&x)[D], double (&dy)[D]) {
        n_f++;
        Rhs{}(tt, x, p, dy);
    }

    // one classical RK4 step at fixed dt (adams.rs:342-383); `store_slot` >= 0: before the step, record
    // (t, y) and f(t, y) in that deque slot (the `if i != 0` block, :372-380)
    __device__ __forceinline__ void rk4_step(int store_slot) {
        double k1[D], k2[D], k3[D], k4[D], in[D], dy[D];
        const double half = 0.5, two = 2.0, one_sixth = A::div(1.0, 6.0);
        f(t, y, dy);
#pragma unroll
        for (int d = 0; d < D; ++d) { k1[d] = A::mul(dy[d], dt); in[d] = A::madd(k1[d], half, y[d]); }
        const double tm = A::madd(half, dt, t);
        f(tm, in, dy);
#pragma unroll
        for (int d = 0; d < D; ++d) { k2[d] = A::mul(dy[d], dt); in[d] = A::madd(k2[d], half, y[d]); }
        f(tm, in, dy);
#pragma unroll
        for (int d = 0; d < D; ++d) { k3[d] = A::mul(dy[d], dt); in[d] = A::add(y[d], k3[d]); }
        f(A::add(t, dt), in, dy);
#pragma unroll
        for (int d = 0; d < D; ++d) k4[d] = A::mul(dy[d], dt);
        if (store_slot >= 0) record(store_slot);
#pragma unroll
        for (int d = 0; d < D; ++d) {
            const double s = A::add(A::madd(k3[d], two, A::madd(k2[d], two, k1[d])), k4[d]);
            y[d] = A::madd(s, one_sixth, y[d]);
        }
        t = A::add(t, dt);
    }
    // prev_derivatives.push_back(f(t, y)); prev_values.push_back((t, y)) into a fixed slot
    __device__ __forceinline__ void record(int slot) {
        double der[D];
        f(t, y, der);
#pragma unroll
        for (int k = 0; k < H; ++k)
            if (slot == k) {
                ht[k] = t;
#pragma unroll
                for (int d = 0; d < D; ++d) { hy[k][d] = y[d]; hf[k][d] = der[d]; }
            }
    }

    __device__ __forceinline__ double root(double x) const {
        if constexpr (STRICT) return det_root<O>(x);
        else return sfu_root<O>(x);
    }

    // one IVPStepper::step call (adams.rs:410-566)
    __device__ __forceinline__ int attempt(bool& yielded) {
        yielded = false;
        if (n_att >= cap) return BACON_E_MAX_ATTEMPTS;
        n_att++;
        if (ym > 0 && ym < (uint32_t)O) {  // emit a stored warm-up point (:415-427)
            const uint32_t get = (uint32_t)O - ym - 1;
            ym -= 1;
            if (ym == 0) ym = O + 1;
#pragma unroll
            for (int k = 0; k < H; ++k)
                if (get == (uint32_t)k) {
                    ot = ht[k];
#pragma unroll
                    for (int d = 0; d < D; ++d) oy[d] = hy[k][d];
                }
            n_acc++;
            yielded = true;
            return -1;
        }
        if (ym == (uint32_t)O + 1) {  // the speculative step becomes a regular point (:434-440)
            ym = 0;
            // prev_values.push_back((t, y)); pop_front() — the derivative deque already holds this step's
#pragma unroll
            for (int k = 0; k + 1 < H; ++k) {
                ht[k] = ht[k + 1];
#pragma unroll
                for (int d = 0; d < D; ++d) hy[k][d] = hy[k + 1][d];
            }
            ht[H - 1] = t;
#pragma unroll
            for (int d = 0; d < D; ++d) hy[H - 1][d] = y[d];
            emit_state();
            yielded = true;
            return -1;
        }
        if (t >= t_end) return BACON_OK;  // :442-444

        if (A::add(t, dt) >= t_end) {  // last step by RK4 (:446-450); what it pushes is never read again
            dt = A::sub(t_end, t);
            rk4_step(-1);
            n_f++;  // the derivative the reference evaluates for the entry it pushes (:385-389)
            emit_state();
            yielded = true;
            return -1;
        }

        if (!have) {  // (re)start with O-1 explicit RK4 steps (:452-463)
#pragma unroll
            for (int d = 0; d < D; ++d) save[d] = y[d];
            const double om1 = A::sub(order, 1.0);
            if (A::madd(dt, om1, t) >= t_end) dt = A::div(A::sub(t_end, t), om1);
#pragma unroll
            for (int i = 0; i < H; ++i) rk4_step(i == 0 ? -1 : i - 1);
            record(H - 1);
            have = true;
            ym = O;
            return -1;  // Redo
        }

        // predictor (Adams-Bashforth, :465-470) and corrector (Adams-Moulton, :472-483)
        double pred[D], corr[D], imp[D];
#pragma unroll
        for (int d = 0; d < D; ++d) {
            double sp = A::mul(hf[0][d], Coef::predictor(O - 2));
            static_for<1, O - 1>([&](auto I) {
                constexpr int i = decltype(I)::value;
                sp = A::madd(hf[i][d], Coef::predictor(O - i - 2), sp);
            });
            pred[d] = A::madd(sp, dt, y[d]);
        }
        f(A::add(t, dt), pred, imp);
        double ss = 0.0;
#pragma unroll
        for (int d = 0; d < D; ++d) {
            double sp = A::mul(imp[d], Coef::corrector(0));
            static_for<0, O - 1>([&](auto I) {
                constexpr int i = decltype(I)::value;
                sp = A::madd(hf[i][d], Coef::corrector(O - i - 1), sp);
            });
            corr[d] = A::madd(sp, dt, y[d]);
            const double df = A::sub(corr[d], pred[d]);
            ss = A::madd(df, df, ss);
        }
        const double error = A::mul(A::div(Coef::error, dt), sqrt(ss));  // :486
        if (error != error) return BACON_E_NONFINITE;  // the reference cannot leave this state (cf. D8)

        if (error <= tol) {  // :488-531
#pragma unroll
            for (int d = 0; d < D; ++d) y[d] = corr[d];
            t = A::add(t, dt);
            if (ym == (uint32_t)O) {
                ym -= 1;
                if (fix_d10) {  // REF_CORRECTED: the derivative of the speculative step joins the deque (D10)
#pragma unroll
                    for (int k = 0; k + 1 < H; ++k)
#pragma unroll
                        for (int d = 0; d < D; ++d) hf[k][d] = hf[k + 1][d];
#pragma unroll
                    for (int d = 0; d < D; ++d) hf[H - 1][d] = imp[d];
                }
                return -1;  // Redo: the warm-up points are yielded first
            }
#pragma unroll
            for (int k = 0; k + 1 < H; ++k) {  // push_back + pop_front on both deques (:503-509)
                ht[k] = ht[k + 1];
#pragma unroll
                for (int d = 0; d < D; ++d) { hy[k][d] = hy[k + 1][d]; hf[k][d] = hf[k + 1][d]; }
            }
            ht[H - 1] = t;
#pragma unroll
            for (int d = 0; d < D; ++d) { hy[H - 1][d] = y[d]; hf[H - 1][d] = imp[d]; }
            if (error < A::mul(0.1, tol)) {  // :511-529
                const double q = root(A::div(tol, A::mul(2.0, error)));
                dt = (q > 4.0) ? A::mul(dt, 4.0) : A::mul(dt, q);
                if (dt > dt_max) dt = dt_max;
                have = false;
            }
            emit_state();
            yielded = true;
            return -1;
        }
        n_rej++;
        if (ym == (uint32_t)O) {  // undo the O-1 warm-up steps (:538-542)
            t = A::sub(t, A::mul(dt, A::sub(order, 1.0)));
#pragma unroll
            for (int d = 0; d < D; ++d) y[d] = save[d];
        }
        const double q = root(A::div(tol, A::mul(2.0, error)));  // :544-551
        dt = (q < 0.1) ? A::mul(dt, 0.1) : A::mul(dt, q);
        if (dt < dt_min) return BACON_E_MIN_DT_EXCEEDED;  // :553-555
        have = false;                                      // :557-558
        return -1;
    }
    __device__ __forceinline__ void emit_state() {
        ot = t;
#pragma unroll
        for (int d = 0; d < D; ++d) oy[d] = y[d];
        n_acc++;
    }
};

// EulerSolver::step (ivp.rs:320-338): fixed dt (clamped on the last step), yields the OLD (t, y).
// cfg.dt_max carries the builder's dt (ivp.rs:396-421).
template <class Rhs, bool STRICT> struct EulerStepper {
    using RhsT = Rhs;
    static constexpr int D = Rhs::DIM;
    static constexpr int P = Rhs::NPARAM;
    using A = Ar<STRICT>;

    double t_start, t_end, dt0;
    uint32_t cap;
    double y[D], p[P > 0 ? P : 1];
    double oy[D], ot;
    double t, dt;
    uint32_t n_acc, n_rej, n_att;

    __device__ __forceinline__ explicit EulerStepper(const bacon_launch_args& a) {
        t_start = a.cfg.t_start;
        t_end = a.cfg.t_end;
        dt0 = a.cfg.dt_max;
        cap = (a.cfg.max_attempts == 0 || a.cfg.max_attempts > 0xFFFFFFFEull) ? 0xFFFFFFFEu
                                                                               : (uint32_t)a.cfg.max_attempts;
        t = t_start;
        dt = dt0;
        ot = t_start;
        n_acc = n_rej = n_att = 0;
#pragma unroll
        for (int d = 0; d < D; ++d) oy[d] = 0.0;
    }
    __device__ __forceinline__ void reset(const bacon_launch_args& a, unsigned long long idx, bool live) {
        t = t_start;
        dt = dt0;
        n_acc = n_rej = n_att = 0;
        if (live) load_problem<D, P>(a, idx, y, p);
    }
    __device__ __forceinline__ void apply_restart(const bacon_launch_args& a, unsigned long long idx) {
        if (a.t0_each) t = a.t0_each[idx];  // (Euler's dt is the builder's, the record's is not used)
        ot = t;
    }
    __device__ __forceinline__ uint32_t n_rhs() const { return n_acc; }
    __device__ __forceinline__ double out_t() const { return ot; }
    __device__ __forceinline__ const double (&out_y() const)[D] { return oy; }
    __device__ __forceinline__ const double (&end_y() const)[D] { return y; }

    __device__ __forceinline__ int attempt(bool& yielded) {
        yielded = false;
        if (n_att >= cap) return BACON_E_MAX_ATTEMPTS;
        n_att++;
        if (t >= t_end) return BACON_OK;                      // ivp.rs:321-323
        if (A::add(t, dt) >= t_end) dt = A::sub(t_end, t);    // :324-326
        double dy[D];
        Rhs{}(t, y, p, dy);                                   // :328-329
        ot = t;                                               // :331-332
#pragma unroll
        for (int d = 0; d < D; ++d) {
            oy[d] = y[d];
            y[d] = A::madd(dy[d], dt, y[d]);                  // :334
        }
        t = A::add(t, dt);                                    // :335
        n_acc++;
        yielded = true;
        return -1;
    }
};

}  // namespace bacon
