// bdf.cuh — K3: the reference's BDF stepper (BDFSolver, src/ivp/bdf.rs:72-122 fields,
// :346-387 RK4 start-up, :390-411 finite-difference Jacobian, :414-475 Broyden "secant",
// :495-634 step state machine) one trajectory per thread, REF_CORRECTED semantics
// (D4 central difference, D5 lower formula uses the lower coefficients,
// D6 rollback by dt*order, D7 implicit derivative at t_{n+1}).
//
// Kept as structure (not typos): O explicit RK4 warm-up steps after EVERY dt change and
// after every step whose error < tol/10 (bdf.rs:602-611), the speculative first implicit
// step after a warm-up (yield_memory == O+1) with rollback on rejection, warm-up points
// yielded one per step() call, error = ||y(6) - y(5)||_2 absolute, halve/double control.
//
// The implicit equations  g(y) = y - dt*beta*f(t_{n+1}, y) + sum_k a_k y_{n+1-k} = 0  are solved
//   * Broyden (default): exactly the reference's iteration — FD Jacobian with h = dt, inverse by
//     partial-pivot LU (nalgebra lu().try_inverse()), Sherman-Morrison updates, stop at
//     ||shift|| <= tol, at most 998 iterations.  (The full-pivot-LU / QR fallbacks of bdf.rs:433-444
//     only trigger on an exactly zero pivot column, i.e. an exactly singular Jacobian: reported
//     as SingularMatrix here.)
//   * Newton (BACON_FLAG_BDF_NEWTON): g'(y) = I - dt*beta*J_f from the RHS's analytic `jac`
//     (central finite differences with h = dt when the functor has none), factored IN REGISTERS
//     by partial-pivot LU every iteration; same stopping rule.  No tensor cores: a 3x3 solve is
//     ~30 flops, the rest of the step is RHS evaluations.
//
// STRICT = the oracle's operation order with every product/sum individually rounded
// (bit-comparable with oracle/, built -ffp-contract=off).  Non-strict = the same algorithm,
// products feeding sums contracted to FMA.
#pragma once
#include "ivp_common.cuh"
#include "tableaux.cuh"

namespace bacon {

template <bool STRICT> struct Ar {
    static __device__ __forceinline__ double mul(double a, double b) { return STRICT ? __dmul_rn(a, b) : a * b; }
    static __device__ __forceinline__ double add(double a, double b) { return STRICT ? __dadd_rn(a, b) : a + b; }
    static __device__ __forceinline__ double sub(double a, double b) { return STRICT ? __dadd_rn(a, -b) : a - b; }
    static __device__ __forceinline__ double div(double a, double b) { return STRICT ? __ddiv_rn(a, b) : a / b; }
    // a*b + c : two roundings (strict) or one FMA
    static __device__ __forceinline__ double madd(double a, double b, double c) {
        return STRICT ? __dadd_rn(__dmul_rn(a, b), c) : fma(a, b, c);
    }
};

template <class Rhs, class = void> struct has_jac { static constexpr bool value = false; };
template <class Rhs>
struct has_jac<Rhs, decltype(void(&Rhs::jac))> { static constexpr bool value = true; };

// jac.lu().try_inverse() (bdf.rs:429-430) restated from nalgebra 0.32: partial pivoting on the first
// largest |x| of the column, gauss step multiplies by the reciprocal of the pivot, unit-lower then
// upper triangular solves against the row-permuted identity.  false = zero pivot.
template <int D, bool STRICT> __device__ __forceinline__ bool inverse_lu_partial(const double (&a)[D][D], double (&inv)[D][D]) {
    using A = Ar<STRICT>;
    double lu[D][D];
#pragma unroll
    for (int r = 0; r < D; ++r)
#pragma unroll
        for (int c = 0; c < D; ++c) {
            lu[r][c] = a[r][c];
            inv[r][c] = (r == c) ? 1.0 : 0.0;
        }
    bool ok = true;
#pragma unroll
    for (int i = 0; i < D; ++i) {
        // pivot search + row swap of lu and of the right-hand side (p.permute_rows(identity)): the
        // swaps of P are applied in order, which is what swapping inv's rows here does
        int piv = i;
        double best = fabs(lu[i][i]);
#pragma unroll
        for (int r = i + 1; r < D; ++r) {
            const double v = fabs(lu[r][i]);
            if (v > best) { best = v; piv = r; }
        }
        if (best == 0.0) { ok = false; continue; }  // nalgebra skips the column; try_inverse then fails
#pragma unroll
        for (int r = i + 1; r < D; ++r) {
            if (piv == r) {
#pragma unroll
                for (int c = 0; c < D; ++c) {
                    const double tmp = lu[i][c]; lu[i][c] = lu[r][c]; lu[r][c] = tmp;
                    const double t2 = inv[i][c]; inv[i][c] = inv[r][c]; inv[r][c] = t2;
                }
            }
        }
        const double inv_diag = A::div(1.0, lu[i][i]);
#pragma unroll
        for (int r = i + 1; r < D; ++r) lu[r][i] = A::mul(lu[r][i], inv_diag);
#pragma unroll
        for (int c = i + 1; c < D; ++c) {
            const double prc = lu[i][c];
#pragma unroll
            for (int r = i + 1; r < D; ++r) lu[r][c] = A::madd(-prc, lu[r][i], lu[r][c]);
        }
    }
    if (!ok) return false;
#pragma unroll
    for (int k = 0; k < D; ++k) {
#pragma unroll
        for (int i = 0; i < D - 1; ++i) {  // unit lower triangular
            const double coeff = inv[i][k];
#pragma unroll
            for (int r = i + 1; r < D; ++r) inv[r][k] = A::madd(-coeff, lu[r][i], inv[r][k]);
        }
#pragma unroll
        for (int i = D - 1; i >= 0; --i) {  // upper triangular
            const double diag = lu[i][i];
            if (diag == 0.0) ok = false;
            const double coeff = A::div(inv[i][k], diag);
            inv[i][k] = coeff;
#pragma unroll
            for (int r = 0; r < i; ++r) inv[r][k] = A::madd(-coeff, lu[r][i], inv[r][k]);
        }
    }
    return ok;
}

// swap rows i and j (j > i, predicated over the static candidates: no dynamic register indexing)
template <int D> __device__ __forceinline__ void swap_rows(double (&m)[D][D], int i_static, int j) {
#pragma unroll
    for (int r = 0; r < D; ++r)
        if (r > i_static && r == j) {
#pragma unroll
            for (int c = 0; c < D; ++c) { const double t = m[i_static][c]; m[i_static][c] = m[r][c]; m[r][c] = t; }
        }
}
template <int D> __device__ __forceinline__ void swap_cols(double (&m)[D][D], int i_static, int j) {
#pragma unroll
    for (int c = 0; c < D; ++c)
        if (c > i_static && c == j) {
#pragma unroll
            for (int r = 0; r < D; ++r) { const double t = m[r][i_static]; m[r][i_static] = m[r][c]; m[r][c] = t; }
        }
}
// L U X = B in place, column by column of B (nalgebra solve.rs; oracle: lu_solve_inplace).  false = zero diagonal.
template <int D, bool STRICT> __device__ __forceinline__ bool lu_solve_inplace(const double (&lu)[D][D], double (&b)[D][D]) {
    using A = Ar<STRICT>;
    bool ok = true;
#pragma unroll
    for (int k = 0; k < D; ++k) {
#pragma unroll
        for (int i = 0; i < D - 1; ++i) {
            const double coeff = b[i][k];
#pragma unroll
            for (int r = i + 1; r < D; ++r) b[r][k] = A::madd(-coeff, lu[r][i], b[r][k]);
        }
#pragma unroll
        for (int i = D - 1; i >= 0; --i) {
            const double diag = lu[i][i];
            if (diag == 0.0) ok = false;
            const double coeff = A::div(b[i][k], diag);
            b[i][k] = coeff;
#pragma unroll
            for (int r = 0; r < i; ++r) b[r][k] = A::madd(-coeff, lu[r][i], b[r][k]);
        }
    }
    return ok;
}

// jac.full_piv_lu().try_inverse() (bdf.rs:433-434), the second link of the reference's inversion chain: complete
// pivoting on the first largest |x| of the trailing block in column-major scan order (nalgebra icamax_full), the same
// gauss step, then  P, L, U solves and the inverse column permutation.  Reached only when the partially pivoted LU hit an
// exactly zero pivot — which REF_LITERAL's Jacobian `(above + below) / 2h` (rank one up to rounding) does.
template <int D, bool STRICT> __device__ __noinline__ bool inverse_lu_full(const double (&a)[D][D], double (&inv)[D][D]) {
    using A = Ar<STRICT>;
    double lu[D][D];
    int cperm[D];
#pragma unroll
    for (int r = 0; r < D; ++r) {
        cperm[r] = -1;
#pragma unroll
        for (int c = 0; c < D; ++c) {
            lu[r][c] = a[r][c];
            inv[r][c] = (r == c) ? 1.0 : 0.0;
        }
    }
    bool stopped = false;
#pragma unroll
    for (int i = 0; i < D; ++i) {
        int pr = i, pc = i;
        double best = -1.0, pivot = 0.0;
#pragma unroll
        for (int c = i; c < D; ++c)
#pragma unroll
            for (int r = i; r < D; ++r) {
                const double v = fabs(lu[r][c]);
                if (!stopped && v > best) { best = v; pr = r; pc = c; pivot = lu[r][c]; }
            }
        if (pivot == 0.0) stopped = true;  // "the remaining of the matrix is zero": the loop ends here
        if (!stopped) {
            if (pc != i) {
                cperm[i] = pc;
                swap_cols<D>(lu, i, pc);
            }
            if (pr != i) {
                swap_rows<D>(lu, i, pr);
                swap_rows<D>(inv, i, pr);  // (P's swaps applied in order to the identity)
            }
            const double inv_diag = A::div(1.0, lu[i][i]);
#pragma unroll
            for (int r = i + 1; r < D; ++r) lu[r][i] = A::mul(lu[r][i], inv_diag);
#pragma unroll
            for (int c = i + 1; c < D; ++c) {
                const double prc = lu[i][c];
#pragma unroll
                for (int r = i + 1; r < D; ++r) lu[r][c] = A::madd(-prc, lu[r][i], lu[r][c]);
            }
        }
    }
    if (!lu_solve_inplace<D, STRICT>(lu, inv)) return false;
#pragma unroll
    for (int k = D - 1; k >= 0; --k)  // q.inv_permute_rows(b)
        if (cperm[k] >= 0) swap_rows<D>(inv, k, cperm[k]);
    return true;
}

// jac.qr().try_inverse() (bdf.rs:437-438), the last link: Householder QR, then R X = Q^T (oracle: inverse_qr).
template <int D, bool STRICT> __device__ __noinline__ bool inverse_qr(const double (&a)[D][D], double (&inv)[D][D]) {
    using A = Ar<STRICT>;
    double r[D][D], qt[D][D];
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
        for (int c = 0; c < D; ++c) {
            r[i][c] = a[i][c];
            qt[i][c] = (i == c) ? 1.0 : 0.0;
        }
    bool ok = true;
#pragma unroll
    for (int k = 0; k < D; ++k) {
        double nrm = 0.0;
#pragma unroll
        for (int i = k; i < D; ++i) nrm = A::madd(r[i][k], r[i][k], nrm);
        nrm = sqrt(nrm);
        if (nrm == 0.0) ok = false;
        if (ok) {
            const double alpha = (r[k][k] >= 0.0) ? -nrm : nrm;
            double v[D];
#pragma unroll
            for (int i = 0; i < D; ++i) v[i] = i >= k ? r[i][k] : 0.0;
            v[k] = A::sub(v[k], alpha);
            double vn = 0.0;
#pragma unroll
            for (int i = k; i < D; ++i) vn = A::madd(v[i], v[i], vn);
            if (vn != 0.0) {
#pragma unroll
                for (int c = 0; c < D; ++c) {
                    double dot = 0.0;
#pragma unroll
                    for (int i = k; i < D; ++i) dot = A::madd(v[i], r[i][c], dot);
                    const double f = A::div(A::mul(2.0, dot), vn);
#pragma unroll
                    for (int i = k; i < D; ++i) r[i][c] = A::sub(r[i][c], A::mul(f, v[i]));
                    double dq = 0.0;
#pragma unroll
                    for (int i = k; i < D; ++i) dq = A::madd(v[i], qt[i][c], dq);
                    const double fq = A::div(A::mul(2.0, dq), vn);
#pragma unroll
                    for (int i = k; i < D; ++i) qt[i][c] = A::sub(qt[i][c], A::mul(fq, v[i]));
                }
            }
        }
    }
    if (!ok) return false;
#pragma unroll
    for (int i = 0; i < D; ++i)
        if (r[i][i] == 0.0) ok = false;
    if (!ok) return false;
#pragma unroll
    for (int c = 0; c < D; ++c)
#pragma unroll
        for (int i = D - 1; i >= 0; --i) {
            const double coeff = A::div(qt[i][c], r[i][i]);
            qt[i][c] = coeff;
#pragma unroll
            for (int rr = 0; rr < i; ++rr) qt[rr][c] = A::madd(-coeff, r[rr][i], qt[rr][c]);
        }
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
        for (int c = 0; c < D; ++c) inv[i][c] = qt[i][c];
    return true;
}

// The reference's inversion chain (bdf.rs:429-444): LU with partial pivoting, else with full pivoting, else QR, else
// SingularMatrix.  The strict build runs the whole chain (bit-exact with the oracle also where a pivot is exactly zero);
// the fast build stops after the first link (the others only matter on an exactly singular Jacobian).
template <int D, bool STRICT> __device__ __forceinline__ bool inverse_chain(const double (&a)[D][D], double (&inv)[D][D]) {
    if (inverse_lu_partial<D, STRICT>(a, inv)) return true;
    if constexpr (STRICT) {
        if (inverse_lu_full<D, STRICT>(a, inv)) return true;
        return inverse_qr<D, STRICT>(a, inv);
    } else {
        return false;
    }
}

// 1/x for the Newton path's pivots: SFU seed (2^-23) + two Newton steps, 5 instructions against the ~20 of an IEEE
// division; relative error ~1e-14 on normal x (a Newton iteration's linear solve needs far less), 0 or inf for
// zero / overflow like the division.
__device__ __forceinline__ double fast_rcp(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    const double e0 = fma(-x, r, 1.0);
    r = fma(r, e0, r);
    const double e1 = fma(-x, r, 1.0);
    return fma(r, e1, r);
}

// In-register partial-pivot LU solve of M x = b (Newton path); M is destroyed.  false = singular.
// One reciprocal per pivot, reused by the back substitution (an FP64 division is a ~15-instruction dependent chain,
// and this kernel is latency-bound: DESIGN.md §4, K3).
template <int D> __device__ __forceinline__ bool lu_solve(double (&M)[D][D], double (&b)[D]) {
    bool ok = true;
    double inv_diag[D];
#pragma unroll
    for (int i = 0; i < D; ++i) {
        int piv = i;
        double best = fabs(M[i][i]);
#pragma unroll
        for (int r = i + 1; r < D; ++r) {
            const double v = fabs(M[r][i]);
            if (v > best) { best = v; piv = r; }
        }
        if (best == 0.0) ok = false;
#pragma unroll
        for (int r = i + 1; r < D; ++r) {
            if (piv == r) {
#pragma unroll
                for (int c = 0; c < D; ++c) { const double tmp = M[i][c]; M[i][c] = M[r][c]; M[r][c] = tmp; }
                const double tb = b[i]; b[i] = b[r]; b[r] = tb;
            }
        }
        inv_diag[i] = fast_rcp(M[i][i]);
#pragma unroll
        for (int r = i + 1; r < D; ++r) {
            const double l = M[r][i] * inv_diag[i];
#pragma unroll
            for (int c = i + 1; c < D; ++c) M[r][c] = fma(-l, M[i][c], M[r][c]);
            b[r] = fma(-l, b[i], b[r]);
        }
    }
#pragma unroll
    for (int i = D - 1; i >= 0; --i) {
        double s = b[i];
#pragma unroll
        for (int c = i + 1; c < D; ++c) s = fma(-M[i][c], b[c], s);
        b[i] = s * inv_diag[i];
    }
    return ok;
}

template <class Rhs, class Coef, bool STRICT, bool NEWTON> struct BdfStepper {
    using RhsT = Rhs;
    static constexpr int D = Rhs::DIM;
    static constexpr int P = Rhs::NPARAM;
    static constexpr int O = Coef::O;
    using A = Ar<STRICT>;

    double t_start, t_end, dt_min, dt_max, tol, dt0, order;
    uint32_t cap;
    // one trajectory (bdf.rs:72-122)
    double y[D], p[P > 0 ? P : 1];
    // prev_values, oldest first; `have` = deque non-empty (it holds 0 or O entries).  Only the STATES are kept: the
    // times of the deque are never used by the formulas, and the one place that reads them — yielding the stored warm-up
    // points (bdf.rs:500-512) — gets the same bits by repeating the warm-up's own additions t <- t + dt from the time
    // the block started (`tb`): dt cannot change between the warm-up and those yields.  Seven doubles fewer to hold and
    // to shift at every accepted step (the shifts were a fifth of this kernel's instructions).
    double hy[O][D], tb;
    double save[D];
    double oy[D], ot;        // the point of the last Ok(...)
    double t, dt;
    bool have;
    // REF_LITERAL (the source as written; strict Broyden build only): FD Jacobian `above + below`
    // (bdf.rs:407), the lower formula summed with the HIGHER coefficients (:568), g evaluated at t_n (:403-454), and the
    // warm-up rollback `time -= dt - order` (:622).  Every reference BDF test ends with an empty path in this mode.
    bool literal;
    __device__ __forceinline__ bool lit() const { return (STRICT && !NEWTON) ? literal : false; }
    uint32_t ym;             // yield_memory (bdf.rs:119)
    uint32_t n_acc, n_rej, n_att, n_f;

    __device__ __forceinline__ explicit BdfStepper(const bacon_launch_args& a) {
        t_start = a.cfg.t_start;
        t_end = a.cfg.t_end;
        dt_min = a.cfg.dt_min;
        dt_max = a.cfg.dt_max;
        tol = a.cfg.tol;
        dt0 = A::mul(A::add(dt_max, dt_min), 0.5);  // bdf.rs:302
        order = (double)O;                           // bdf.rs:293
        literal = STRICT && !NEWTON && a.cfg.semantics == BACON_SEM_LITERAL;
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
        tb = t_start;
#pragma unroll
        for (int k = 0; k < O; ++k) {
#pragma unroll
            for (int d = 0; d < D; ++d) hy[k][d] = 0.0;
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

    __device__ __forceinline__ void push_pop(double, const double (&x)[D]) {  // push_back + pop_front
#pragma unroll
        for (int k = 0; k + 1 < O; ++k) {
#pragma unroll
            for (int d = 0; d < D; ++d) hy[k][d] = hy[k + 1][d];
        }
#pragma unroll
        for (int d = 0; d < D; ++d) hy[O - 1][d] = x[d];
    }

    // one classical RK4 step at fixed dt (bdf.rs:348-381)
    __device__ __forceinline__ void rk4_step() {
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
        for (int d = 0; d < D; ++d) {
            k4[d] = A::mul(dy[d], dt);
            const double s = A::add(A::madd(k3[d], two, A::madd(k2[d], two, k1[d])), k4[d]);
            y[d] = A::madd(s, one_sixth, y[d]);
        }
        t = A::add(t, dt);
    }

    // higher_func / lower_func (bdf.rs:548-575): g(x) = x - dt*beta*f(tg, x) + sum_ind coef[ind]*prev[O-ind]
    template <bool HIGHER> __device__ __forceinline__ void g_eval(double tg, const double (&x)[D], double (&out)[D]) {
        double dy[D];
        f(tg, x, dy);
        const double beta = HIGHER ? Coef::higher(0) : Coef::lower(0);
#pragma unroll
        for (int d = 0; d < D; ++d) {
            double sp = A::mul(A::mul(-dy[d], dt), beta);
            static_for<1, O>([&](auto I) {
                constexpr int ind = decltype(I)::value;
                constexpr double c = HIGHER ? Coef::higher(ind) : Coef::lower(ind);
                if constexpr (STRICT && !NEWTON && !HIGHER) {
                    sp = A::madd(hy[O - ind][d], lit() ? Coef::higher(ind) : c, sp);  // bdf.rs:568 (D5)
                } else if constexpr (STRICT || c != 0.0) {
                    sp = A::madd(hy[O - ind][d], c, sp);
                }
            });
            out[d] = A::add(sp, x[d]);
        }
    }

    // roots::secant as embedded in bdf.rs:414-475.  Returns a bacon_status.
    template <bool HIGHER> __device__ __forceinline__ int broyden(double (&res)[D]) {
        const double tg = lit() ? t : A::add(t, dt);  // bdf.rs:403,405,423,454 pass self.time (D7)
        double guess[D], g[D];
#pragma unroll
        for (int d = 0; d < D; ++d) guess[d] = y[d];
        g_eval<HIGHER>(tg, guess, g);
        double jac[D][D], jinv[D][D];
        {
            const double h = dt;
            const double denom = A::div(1.0, A::mul(2.0, h));
#pragma unroll
            for (int ind = 0; ind < D; ++ind) {
                double above[D], below[D];
                guess[ind] = A::add(guess[ind], h);
                g_eval<HIGHER>(tg, guess, above);
                guess[ind] = A::sub(guess[ind], A::mul(2.0, h));
                g_eval<HIGHER>(tg, guess, below);
                guess[ind] = A::add(guess[ind], h);
#pragma unroll
                for (int r = 0; r < D; ++r)  // bdf.rs:407: `(above + below) * denom` as written (D4)
                    jac[r][ind] = A::mul(lit() ? A::add(above[r], below[r]) : A::sub(above[r], below[r]), denom);
            }
        }
        if (!inverse_chain<D, STRICT>(jac, jinv)) return BACON_E_SINGULAR;  // bdf.rs:429-444
        double shift[D];
        neg_matvec(jinv, g, shift);
#pragma unroll
        for (int d = 0; d < D; ++d) guess[d] = A::add(guess[d], shift[d]);
        for (int n = 2; n < 1000; ++n) {  // bdf.rs:449
            double g_last[D], diff[D], adj[D], u[D];
#pragma unroll
            for (int d = 0; d < D; ++d) g_last[d] = g[d];
            g_eval<HIGHER>(tg, guess, g);
#pragma unroll
            for (int d = 0; d < D; ++d) diff[d] = A::sub(g[d], g_last[d]);
            neg_matvec(jinv, diff, adj);
            double pp = 0.0;
#pragma unroll
            for (int d = 0; d < D; ++d) pp = A::madd(-shift[d], adj[d], pp);
#pragma unroll
            for (int c = 0; c < D; ++c) {
                double s = 0.0;
#pragma unroll
                for (int r = 0; r < D; ++r) s = A::madd(shift[r], jinv[r][c], s);
                u[c] = s;
            }
#pragma unroll
            for (int r = 0; r < D; ++r)
#pragma unroll
                for (int c = 0; c < D; ++c)
                    jinv[r][c] = A::add(jinv[r][c], A::div(A::mul(A::add(shift[r], adj[r]), u[c]), pp));
            neg_matvec(jinv, g, shift);
            double ss = 0.0;
#pragma unroll
            for (int d = 0; d < D; ++d) {
                guess[d] = A::add(guess[d], shift[d]);
                ss = A::madd(shift[d], shift[d], ss);
            }
            if (sqrt(ss) <= tol) {  // bdf.rs:468 (sqrt is correctly rounded in both builds)
#pragma unroll
                for (int d = 0; d < D; ++d) res[d] = guess[d];
                return BACON_OK;
            }
            if (!STRICT && ss != ss) return BACON_E_MAX_ITER;  // NaN never converges: the reference runs out of iterations
        }
        return BACON_E_MAX_ITER;  // bdf.rs:474
    }
    __device__ __forceinline__ void neg_matvec(const double (&M)[D][D], const double (&v)[D], double (&r)[D]) {
#pragma unroll
        for (int i = 0; i < D; ++i) {
            double s = 0.0;
#pragma unroll
            for (int j = 0; j < D; ++j) s = A::madd(-M[i][j], v[j], s);
            r[i] = s;
        }
    }

    // Newton on g with an in-register LU of g'(x) = I - dt*beta*J_f(tg, x).  The history part of g (bdf.rs:553-559),
    // sum_ind coef[ind] * prev[O - ind], does not depend on x: it is summed once per solve, so an evaluation of g is
    // the right-hand side plus two FMAs per component and the 7-deep history is dead while Newton iterates.
    template <bool HIGHER> __device__ __forceinline__ int newton(double (&res)[D]) {
        const double tg = t + dt;
        const double beta = HIGHER ? Coef::higher(0) : Coef::lower(0);
        const double hb = dt * beta;
        double x[D], hsum[D];
#pragma unroll
        for (int d = 0; d < D; ++d) {
            x[d] = y[d];
            double sp = 0.0;
            static_for<1, O>([&](auto I) {
                constexpr int ind = decltype(I)::value;
                constexpr double c = HIGHER ? Coef::higher(ind) : Coef::lower(ind);
                if constexpr (c != 0.0) sp = fma(hy[O - ind][d], c, sp);
            });
            hsum[d] = sp;
        }
        auto g_of = [&](const double (&at)[D], double (&out)[D]) {
            double dy[D];
            f(tg, at, dy);
#pragma unroll
            for (int d = 0; d < D; ++d) out[d] = fma(-hb, dy[d], at[d] + hsum[d]);
        };
        for (int n = 2; n < 1000; ++n) {
            double g[D], M[D][D];
            g_of(x, g);
            if constexpr (has_jac<Rhs>::value) {
                Rhs{}.jac(tg, x, p, M);
#pragma unroll
                for (int r = 0; r < D; ++r)
#pragma unroll
                    for (int c = 0; c < D; ++c) M[r][c] = (r == c ? 1.0 : 0.0) - hb * M[r][c];
            } else {  // central differences of g with h = dt (bdf.rs:390-411 semantics)
                const double inv2h = 1.0 / (2.0 * dt);
#pragma unroll
                for (int c = 0; c < D; ++c) {
                    double xa[D], up[D], dn[D];
#pragma unroll
                    for (int d = 0; d < D; ++d) xa[d] = x[d];
                    xa[c] = x[c] + dt;
                    g_of(xa, up);
                    xa[c] = x[c] - dt;
                    g_of(xa, dn);
#pragma unroll
                    for (int r = 0; r < D; ++r) M[r][c] = (up[r] - dn[r]) * inv2h;
                }
            }
            double rhs[D];
#pragma unroll
            for (int d = 0; d < D; ++d) rhs[d] = -g[d];
            if (!lu_solve<D>(M, rhs)) return BACON_E_SINGULAR;
            double ss = 0.0;
#pragma unroll
            for (int d = 0; d < D; ++d) {
                x[d] += rhs[d];
                ss = fma(rhs[d], rhs[d], ss);
            }
            if (ss <= tol * tol) {
#pragma unroll
                for (int d = 0; d < D; ++d) res[d] = x[d];
                return BACON_OK;
            }
            if (ss != ss) return BACON_E_MAX_ITER;
        }
        return BACON_E_MAX_ITER;
    }

    // one IVPStepper::step call (bdf.rs:495-634)
    __device__ __forceinline__ int attempt(bool& yielded) {
        yielded = false;
        if (n_att >= cap) return BACON_E_MAX_ATTEMPTS;
        n_att++;
        if (ym > 0 && ym <= (uint32_t)O) {  // A: emit a stored warm-up point (bdf.rs:500-512)
            const uint32_t get = (uint32_t)O - ym;
            ym -= 1;
            if (ym == 0) ym = O + 2;
            ot = tb;  // time of stored point `get` = tb + dt, get + 1 times over, rounded as the warm-up rounded it
#pragma unroll
            for (int k = 0; k < O; ++k) {
                if ((uint32_t)k <= get) ot = A::add(ot, dt);
                if (get == (uint32_t)k) {
#pragma unroll
                    for (int d = 0; d < D; ++d) oy[d] = hy[k][d];
                }
            }
            n_acc++;
            yielded = true;
            return -1;
        }
        if (ym == (uint32_t)O + 2) {  // B: the speculative implicit step becomes a regular point (:519-525)
            ym = 0;
            push_pop(t, y);
            emit_state();
            yielded = true;
            return -1;
        }
        if (t >= t_end) return BACON_OK;  // C (:527-529)

        if (A::add(t, dt) >= t_end) {  // D: last step by RK4 (:531-535)
            dt = A::sub(t_end, t);
            rk4_step();
            emit_state();
            yielded = true;
            return -1;
        }

        if (!have) {  // E: (re)start with O explicit RK4 steps (:537-546)
#pragma unroll
            for (int d = 0; d < D; ++d) save[d] = y[d];
            if (A::madd(dt, order, t) >= t_end) dt = A::div(A::sub(t_end, t), order);
#pragma unroll
            tb = t;
#pragma unroll
            for (int k = 0; k < O; ++k) {
                rk4_step();
#pragma unroll
                for (int d = 0; d < D; ++d) hy[k][d] = y[d];
            }
            have = true;
            ym = O + 1;
            return -1;  // Redo
        }

        // F: implicit step, orders 6 and 5 (:548-581)
        double hi[D], lo[D];
        int rc = NEWTON ? newton<true>(hi) : broyden<true>(hi);
        if (rc != BACON_OK) return rc;
        rc = NEWTON ? newton<false>(lo) : broyden<false>(lo);
        if (rc != BACON_OK) return rc;
        double ss = 0.0;
#pragma unroll
        for (int d = 0; d < D; ++d) {
            const double df = A::sub(hi[d], lo[d]);
            ss = A::madd(df, df, ss);
        }
        // error = ||hi - lo||_2 (:580-581).  The fast build compares squares (no square root on the path); the strict one
        // takes the root like the oracle.
        const double error = STRICT ? sqrt(ss) : ss;
        const double tol_cmp = STRICT ? tol : tol * tol, tenth_cmp = STRICT ? A::mul(0.1, tol) : 0.01 * (tol * tol);

        if (error <= tol_cmp) {  // :583-613
#pragma unroll
            for (int d = 0; d < D; ++d) y[d] = hi[d];
            t = A::add(t, dt);
            if (ym == (uint32_t)O + 1) {
                ym -= 1;
                return -1;  // Redo: the warm-up points are yielded first
            }
            push_pop(t, y);
            if (error < tenth_cmp) {  // :602-611
                dt = A::mul(dt, 2.0);
                if (dt > dt_max) dt = dt_max;
                have = false;
            }
            emit_state();
            yielded = true;
            return -1;
        }
        n_rej++;
        if (ym == (uint32_t)O + 1) {  // :620-624 (intent: undo the O warm-up steps, D6)
            t = lit() ? A::sub(t, A::sub(dt, order)) : A::sub(t, A::mul(dt, order));  // :622 as written (D6)
#pragma unroll
            for (int d = 0; d < D; ++d) y[d] = save[d];
        }
        dt = A::mul(dt, 0.5);  // :626
        if (dt < dt_min) return BACON_E_MIN_DT_EXCEEDED;
        have = false;  // :632
        return -1;
    }
    __device__ __forceinline__ void emit_state() {
        ot = t;
#pragma unroll
        for (int d = 0; d < D; ++d) oy[d] = y[d];
        n_acc++;
    }
};

}  // namespace bacon
