// bacon_oracle.hpp — CPU ORACLE (test infrastructure, NOT product code).
//
// A statement-for-statement C++ restatement of the hot path of aftix/bacon
// (bacon-sci 0.16.2): the adaptive Runge-Kutta stepper (src/ivp/rk.rs), the BDF
// stepper (src/ivp/bdf.rs) and the IVPIterator drive loop (src/ivp.rs:220-238).
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// `--impl reference` legs may compile, link or call this; the product path
// (bacon_b200/csrc) never does.
//
// PARITY PINNING:
//  * The Rust reference cannot be built here (no cargo/rustc, no network), so
//    this restatement is checked against the assertions of the reference's OWN
//    tests replayed through it: rk.rs:682-758 (pins c, b, e and the dt
//    controller) and src/tests/roots/mod.rs:169-221 (pins the Broyden + LU code
//    bdf.rs:414-475 shares with roots::secant).
//  * Every coefficient table below (RK stage matrix, nodes, weights, error
//    weights, safety factor; BDF and Adams coefficients) is held bit for bit
//    against the numbers parsed out of the reference's own source text
//    (tests/golden/reference_coefficients.json, tests/test_oracle.py).
//  * The reference's RK tests use y-independent right-hand sides, so what the
//    step DOES with the stage matrix (rk.rs:370-384) is NOT pinned by any
//    reference test: "stage-matrix parity unpinned" beyond its coefficients.
//    All eight BDF tests (bdf.rs:785-1063) iterate over an empty path: "BDF
//    parity unpinned".  Both are anchored on closed forms and SciPy instead
//    (tests/golden/), and on SECOND, independent readings of the same Rust
//    source in plain Python (tests/rk_second_reading.py,
//    tests/bdf_second_reading.py, tests/adams_second_reading.py) that agree
//    with this file bit for bit.
//  * Arithmetic that lives in nalgebra 0.32 (crates.io, un-vendored, patch
//    level unpinned — no Cargo.lock in the reference): column-major from_vec,
//    row_iter, norm() = sqrt(sum x^2) accumulated in storage order, partial-
//    pivot lu()/try_inverse(), full_piv_lu(), qr().  Restated from the
//    published algorithms; ulp-level operation order is unpinned.
//
// Build with -O2 -ffp-contract=off: rustc never contracts a*b+c into an FMA.
//
// Mode::Literal  = the source as written (defects D1-D7 of SURVEY.md §8c kept).
// Mode::Corrected = D1-D7 repaired, nothing else.
#pragma once
#include <array>
#include <cmath>
#include <cstring>
#include <cstdint>
#include <deque>
#include <utility>
#include <vector>

namespace bacon_oracle {

enum class Mode : int { Corrected = 0, Literal = 1 };

// how (tol/error)^(1/4) is evaluated (rk.rs:401 calls f64::powf -> libm pow).
// For the Adams path (adams.rs:527, :553: powf(1/order)) mode 1 selects det_root below.
enum class PowMode : int { LibmPow = 0, SqrtSqrt = 1 };

// x^(1/N) from correctly rounded +, *, / and integer arithmetic on the bit pattern only: reproducible
// bit for bit on the GPU (bacon_b200/csrc/adams.cuh det_root), unlike libm pow.  8 Newton steps from a
// linear-in-the-exponent guess; agrees with pow(x, 1/N) to a few ulp (tests/test_oracle.py).
template <int N> inline double det_root(double x) {
    if (!(x > 0.0)) return x;
    if (x > 1.7976931348623157e308) return x;
    const long long one = 0x3ff0000000000000ll;
    long long bits;
    std::memcpy(&bits, &x, 8);
    const long long gb = one + (bits - one) / N;
    double z;
    std::memcpy(&z, &gb, 8);
    for (int it = 0; it < 8; ++it) {
        double zn1 = z;
        for (int k = 2; k < N; ++k) zn1 = zn1 * z;
        z = ((double)(N - 1) * z + x / zn1) / (double)N;
    }
    return z;
}

// status codes shared with include/bacon_ivp.h (bacon_status)
enum : int {
    ST_OK = 0, ST_USER = 2, ST_MIN_DT = 8, ST_MAX_ITER = 9, ST_SINGULAR = 10,
    ST_NONFINITE = 13, ST_MAX_ATTEMPTS = 14, ST_HISTORY_OVERFLOW = 15, ST_STOPPED_AT_EVENT = 19
};

// result of one IVPStepper::step call (src/ivp.rs:96, :20-28)
enum class StepKind : int { Ok, Redo, Done, Failure };

template <int D> using Vec = std::array<double, D>;

// ------------------------------------------------------------------------
// Runge-Kutta tableaux (rk.rs:430-526 RKF45, rk.rs:563-621 BS23)
// ------------------------------------------------------------------------
template <int O> struct RkTableau {
    double c[O];     // t_coefficients
    double A[O][O];  // k_coefficients as the stepper's row_iter() sees them: A[row][col]
    double b[O];     // avg_coefficients
    double e[O];     // error_coefficients
    double safety;   // "point_eighty_four" (rk.rs:266-268)
};

// `listed` holds the O*O numbers in the order the source lists them ("Row 0",
// "Row 1", ...).  BSMatrix::from_vec (rk.rs:459, :580) fills COLUMN by column,
// so as written M(r,c) = listed[c*O + r] (defect D1); the evident intent is
// M(r,c) = listed[r*O + c].
template <int O> inline void fill_matrix(RkTableau<O>& T, const double* listed, Mode mode) {
    for (int r = 0; r < O; ++r)
        for (int c = 0; c < O; ++c)
            T.A[r][c] = (mode == Mode::Literal) ? listed[c * O + r] : listed[r * O + c];
}

inline RkTableau<6> tableau_rkf45(Mode mode) {
    RkTableau<6> T{};
    const double c[6] = {0.0, 1.0 / 4.0, 3.0 / 8.0, 12.0 / 13.0, 1.0, 1.0 / 2.0};  // rk.rs:443-450
    // rk.rs:499: 1859/4014 as written (D2); Fehlberg's a64 is 1859/4104.
    const double a64 = (mode == Mode::Literal) ? 1859.0 / 4014.0 : 1859.0 / 4104.0;
    const double listed[36] = {
        0, 0, 0, 0, 0, 0,                                                        // Row 0  rk.rs:460-466
        1.0 / 4.0, 0, 0, 0, 0, 0,                                                // Row 1  :467-473
        3.0 / 32.0, 9.0 / 32.0, 0, 0, 0, 0,                                      // Row 2  :474-480
        1932.0 / 2197.0, -7200.0 / 2197.0, 7296.0 / 2197.0, 0, 0, 0,             // Row 3  :481-487
        439.0 / 216.0, -8.0, 3680.0 / 513.0, -845.0 / 4104.0, 0, 0,              // Row 4  :488-494
        -8.0 / 27.0, 2.0, -3544.0 / 2565.0, a64, -11.0 / 40.0, 0};               // Row 5  :495-501
    const double b[6] = {25.0 / 216.0, 0.0, 1408.0 / 2565.0, 2197.0 / 4104.0, -(1.0 / 5.0), 0.0};  // :506-513
    const double e[6] = {1.0 / 360.0, 0.0, -128.0 / 4275.0, -2197.0 / 75240.0, 1.0 / 50.0,
                         2.0 / 55.0};  // :517-524
    for (int i = 0; i < 6; ++i) { T.c[i] = c[i]; T.b[i] = b[i]; T.e[i] = e[i]; }
    fill_matrix<6>(T, listed, mode);
    // rk.rs:266-268: eighty_four = from_u8(100) (D3) -> 100/100
    T.safety = (mode == Mode::Literal) ? 100.0 / 100.0 : 84.0 / 100.0;
    return T;
}

inline RkTableau<4> tableau_bs23(Mode mode) {
    RkTableau<4> T{};
    const double c[4] = {0.0, 1.0 / 2.0, 3.0 / 4.0, 1.0};                        // rk.rs:569-574
    const double listed[16] = {0, 0, 0, 0,                                       // Row 0 :581-585
                               1.0 / 2.0, 0, 0, 0,                               // Row 1 :586-590
                               0, 3.0 / 4.0, 0, 0,                               // Row 2 :591-595
                               2.0 / 9.0, 1.0 / 3.0, 4.0 / 9.0, 0};              // Row 3 :596-600
    const double b[4] = {2.0 / 9.0, 1.0 / 3.0, 4.0 / 9.0, 0.0};                  // :605-610
    const double e[4] = {-5.0 / 72.0, 1.0 / 12.0, 1.0 / 9.0, -(1.0 / 8.0)};      // :614-619
    for (int i = 0; i < 4; ++i) { T.c[i] = c[i]; T.b[i] = b[i]; T.e[i] = e[i]; }
    fill_matrix<4>(T, listed, mode);
    T.safety = (mode == Mode::Literal) ? 100.0 / 100.0 : 84.0 / 100.0;
    return T;
}

// nalgebra norm(): sqrt of the sum of squares accumulated in storage order.
template <int D> inline double norm2(const Vec<D>& v) {
    double s = 0.0;
    for (int d = 0; d < D; ++d) s += v[d] * v[d];
    return std::sqrt(s);
}

// Per-trajectory counters the harness reports.
struct Counters {
    uint64_t n_accept = 0;  // Ok(...) results of step() (points yielded)
    uint64_t n_reject = 0;  // rejected attempts
    uint64_t n_rhs = 0;     // derivative calls
    uint64_t n_steps = 0;   // step() calls
    uint64_t n_iter = 0;    // Broyden iterations (BDF)
    uint64_t n_implicit = 0;  // implicit BDF steps attempted
};

// ------------------------------------------------------------------------
// RungeKuttaSolver (rk.rs:75-116 fields, :361-423 step)
// Rhs: struct with static constexpr int DIM and
//      bool operator()(double t, const double* y, const double* p, double* dy) const
//      (false = UserError, ivp.rs:31)
// ------------------------------------------------------------------------
template <int D, int O, class Rhs> struct RungeKuttaSolver {
    double dt_max, dt_min, time, end, tolerance;
    double dt;
    Vec<D> state;
    RkTableau<O> tab;
    double half_steps[O][D];  // column i of the reference's D x O matrix = half_steps[i]
    Vec<D> scratch_pad;
    double one_tenth, one_fourth, four;
    Rhs rhs;
    const double* params;
    PowMode pow_mode;
    Counters cnt;
    int fail_code = 0;

    // RungeKutta::solve (rk.rs:249-343)
    RungeKuttaSolver(const RkTableau<O>& T, Rhs f, const double* p, const double* y0, double t0,
                     double t1, double dtmin, double dtmax, double tol, PowMode pm)
        : dt_max(dtmax), dt_min(dtmin), time(t0), end(t1), tolerance(tol), tab(T), rhs(f), params(p),
          pow_mode(pm) {
        const double two = 2.0;
        const double half = 1.0 / two;  // rk.rs:258-259
        one_tenth = 1.0 / 10.0;         // :261-262
        four = 4.0;                     // :263
        one_fourth = 1.0 / four;        // :264
        dt = (dtmax + dtmin) * half;    // :315
        for (int d = 0; d < D; ++d) state[d] = y0[d];
        for (int i = 0; i < O; ++i)
            for (int d = 0; d < D; ++d) half_steps[i][d] = 0.0;  // :323-327
        scratch_pad.fill(0.0);
    }

    // rk.rs:361-423
    StepKind step() {
        cnt.n_steps++;
        if (time >= end) return StepKind::Done;  // :362-364
        if (time + dt >= end) dt = end - time;   // :366-368

        for (int i = 0; i < O; ++i) {  // :370 row_iter().enumerate()
            scratch_pad = state;       // :371
            for (int j = 0; j < O; ++j)  // :372-374 (dense row, structural zeros included)
                for (int d = 0; d < D; ++d) scratch_pad[d] += half_steps[j][d] * tab.A[i][j];
            const double step_time = time + tab.c[i] * dt;  // :376
            double dy[D];
            cnt.n_rhs++;
            if (!rhs(step_time, scratch_pad.data(), params, dy)) {  // :377-381 (`?`)
                fail_code = ST_USER;
                return StepKind::Failure;
            }
            for (int d = 0; d < D; ++d) half_steps[i][d] = dy[d] * dt;  // :381-383
        }

        for (int d = 0; d < D; ++d) scratch_pad[d] = half_steps[0][d] * tab.e[0];  // :386
        for (int ind = 1; ind < O; ++ind)                                          // :387-389
            for (int d = 0; d < D; ++d) scratch_pad[d] += half_steps[ind][d] * tab.e[ind];
        const double error = norm2<D>(scratch_pad) / dt;  // :390

        if (std::isnan(error)) {  // D8: the reference returns Redo forever
            fail_code = ST_NONFINITE;
            return StepKind::Failure;
        }

        if (error <= tolerance) {  // :392-398
            time += dt;
            for (int ind = 0; ind < O; ++ind)
                for (int d = 0; d < D; ++d) state[d] += half_steps[ind][d] * tab.b[ind];
        }

        const double ratio = tolerance / error;  // :400-401
        const double root = (pow_mode == PowMode::LibmPow) ? std::pow(ratio, one_fourth)
                                                           : std::sqrt(std::sqrt(ratio));
        const double delta = tab.safety * root;
        if (delta <= one_tenth) dt *= one_tenth;  // :402-408
        else if (delta >= four) dt *= four;
        else dt *= delta;

        if (dt > dt_max) dt = dt_max;  // :410-412

        if (dt < dt_min && time < end) {  // :414-416
            // the reference returns Failure BEFORE reporting an accepted point (:418)
            fail_code = ST_MIN_DT;
            if (!(error <= tolerance)) cnt.n_reject++;
            return StepKind::Failure;
        }

        if (error <= tolerance) {  // :418-422
            cnt.n_accept++;
            return StepKind::Ok;
        }
        cnt.n_reject++;
        return StepKind::Redo;
    }
};

// ------------------------------------------------------------------------
// small dense linear algebra restated from nalgebra 0.32 (un-vendored):
//   LU::new + try_inverse (partial pivoting), FullPivLU, QR (Householder).
// Matrices are M[row][col].
// ------------------------------------------------------------------------
template <int D> struct Mat { double m[D][D]; };

template <int D> inline Mat<D> mat_identity() {
    Mat<D> I{};
    for (int i = 0; i < D; ++i) for (int j = 0; j < D; ++j) I.m[i][j] = (i == j) ? 1.0 : 0.0;
    return I;
}

// lower-unit then upper triangular solve of LU * X = B in place, column by
// column of B, column-oriented updates (nalgebra solve.rs).  false = zero diagonal.
template <int D> inline bool lu_solve_inplace(const Mat<D>& lu, Mat<D>& b) {
    for (int k = 0; k < D; ++k) {
        for (int i = 0; i < D - 1; ++i) {  // solve_lower_triangular_with_diag_mut(diag = 1)
            const double coeff = b.m[i][k] / 1.0;
            for (int r = i + 1; r < D; ++r) b.m[r][k] = (-coeff) * lu.m[r][i] + b.m[r][k];
        }
        for (int i = D - 1; i >= 0; --i) {  // solve_upper_triangular_mut
            const double diag = lu.m[i][i];
            if (diag == 0.0) return false;
            const double coeff = b.m[i][k] / diag;
            b.m[i][k] = coeff;
            for (int r = 0; r < i; ++r) b.m[r][k] = (-coeff) * lu.m[r][i] + b.m[r][k];
        }
    }
    return true;
}

// jac.lu().try_inverse()  (bdf.rs:429-430)
template <int D> inline bool inverse_lu_partial(const Mat<D>& a, Mat<D>& inv) {
    Mat<D> lu = a;
    int perm_a[D], perm_b[D], nperm = 0;
    for (int i = 0; i < D; ++i) {
        int piv = i;  // icamax: first index of the largest |x| in column i, rows i..
        double best = std::fabs(lu.m[i][i]);
        for (int r = i + 1; r < D; ++r)
            if (std::fabs(lu.m[r][i]) > best) { best = std::fabs(lu.m[r][i]); piv = r; }
        const double diag = lu.m[piv][i];
        if (diag == 0.0) continue;  // "no non-zero entries on this column"
        if (piv != i) {
            perm_a[nperm] = i; perm_b[nperm] = piv; nperm++;
            for (int c = 0; c < D; ++c) std::swap(lu.m[i][c], lu.m[piv][c]);
        }
        const double inv_diag = 1.0 / diag;  // gauss_step multiplies by the reciprocal
        for (int r = i + 1; r < D; ++r) lu.m[r][i] *= inv_diag;
        for (int c = i + 1; c < D; ++c) {
            const double pivot_row_c = lu.m[i][c];
            for (int r = i + 1; r < D; ++r) lu.m[r][c] = (-pivot_row_c) * lu.m[r][i] + lu.m[r][c];
        }
    }
    inv = mat_identity<D>();
    for (int k = 0; k < nperm; ++k)  // p.permute_rows(b)
        for (int c = 0; c < D; ++c) std::swap(inv.m[perm_a[k]][c], inv.m[perm_b[k]][c]);
    return lu_solve_inplace<D>(lu, inv);
}

// jac.full_piv_lu().try_inverse()  (bdf.rs:433-434)
template <int D> inline bool inverse_lu_full(const Mat<D>& a, Mat<D>& inv) {
    Mat<D> lu = a;
    int pr_a[D], pr_b[D], npr = 0, pc_a[D], pc_b[D], npc = 0;
    for (int i = 0; i < D; ++i) {
        int pr = i, pc = i;  // icamax_full over the trailing block, column-major scan
        double best = -1.0;
        for (int c = i; c < D; ++c)
            for (int r = i; r < D; ++r)
                if (std::fabs(lu.m[r][c]) > best) { best = std::fabs(lu.m[r][c]); pr = r; pc = c; }
        if (lu.m[pr][pc] == 0.0) break;  // "the remaining of the matrix is zero"
        if (pc != i) {
            pc_a[npc] = i; pc_b[npc] = pc; npc++;
            for (int r = 0; r < D; ++r) std::swap(lu.m[r][i], lu.m[r][pc]);
        }
        if (pr != i) {
            pr_a[npr] = i; pr_b[npr] = pr; npr++;
            for (int c = 0; c < D; ++c) std::swap(lu.m[i][c], lu.m[pr][c]);
        }
        const double inv_diag = 1.0 / lu.m[i][i];
        for (int r = i + 1; r < D; ++r) lu.m[r][i] *= inv_diag;
        for (int c = i + 1; c < D; ++c) {
            const double pivot_row_c = lu.m[i][c];
            for (int r = i + 1; r < D; ++r) lu.m[r][c] = (-pivot_row_c) * lu.m[r][i] + lu.m[r][c];
        }
    }
    inv = mat_identity<D>();
    for (int k = 0; k < npr; ++k)
        for (int c = 0; c < D; ++c) std::swap(inv.m[pr_a[k]][c], inv.m[pr_b[k]][c]);
    if (!lu_solve_inplace<D>(lu, inv)) return false;
    for (int k = npc - 1; k >= 0; --k)  // q.inv_permute_rows(b)
        for (int c = 0; c < D; ++c) std::swap(inv.m[pc_a[k]][c], inv.m[pc_b[k]][c]);
    return true;
}

// jac.qr().try_inverse()  (bdf.rs:437-438): Householder QR, then R X = Q^T.
template <int D> inline bool inverse_qr(const Mat<D>& a, Mat<D>& inv) {
    Mat<D> r = a;
    Mat<D> qt = mat_identity<D>();  // accumulates Q^T
    for (int k = 0; k < D; ++k) {
        double nrm = 0.0;
        for (int i = k; i < D; ++i) nrm += r.m[i][k] * r.m[i][k];
        nrm = std::sqrt(nrm);
        if (nrm == 0.0) return false;
        const double alpha = (r.m[k][k] >= 0.0) ? -nrm : nrm;
        double v[D];
        for (int i = 0; i < D; ++i) v[i] = 0.0;
        for (int i = k; i < D; ++i) v[i] = r.m[i][k];
        v[k] -= alpha;
        double vn = 0.0;
        for (int i = k; i < D; ++i) vn += v[i] * v[i];
        if (vn != 0.0) {
            for (int c = 0; c < D; ++c) {
                double dot = 0.0;
                for (int i = k; i < D; ++i) dot += v[i] * r.m[i][c];
                const double f = 2.0 * dot / vn;
                for (int i = k; i < D; ++i) r.m[i][c] -= f * v[i];
                double dq = 0.0;
                for (int i = k; i < D; ++i) dq += v[i] * qt.m[i][c];
                const double fq = 2.0 * dq / vn;
                for (int i = k; i < D; ++i) qt.m[i][c] -= fq * v[i];
            }
        }
    }
    for (int i = 0; i < D; ++i) if (r.m[i][i] == 0.0) return false;
    inv = qt;
    for (int c = 0; c < D; ++c)
        for (int i = D - 1; i >= 0; --i) {
            const double coeff = inv.m[i][c] / r.m[i][i];
            inv.m[i][c] = coeff;
            for (int rr = 0; rr < i; ++rr) inv.m[rr][c] = (-coeff) * r.m[rr][i] + inv.m[rr][c];
        }
    return true;
}

// the LU -> full-pivot LU -> QR fallback chain of bdf.rs:429-444
template <int D> inline bool inverse_chain(const Mat<D>& a, Mat<D>& inv) {
    if (inverse_lu_partial<D>(a, inv)) return true;
    if (inverse_lu_full<D>(a, inv)) return true;
    return inverse_qr<D>(a, inv);
}

template <int D> inline Vec<D> neg_matvec(const Mat<D>& M, const Vec<D>& v) {  // -&M * &v
    Vec<D> r;
    for (int i = 0; i < D; ++i) {
        double s = 0.0;  // nalgebra gemv: column-oriented axpy accumulation
        for (int j = 0; j < D; ++j) s += (-M.m[i][j]) * v[j];
        r[i] = s;
    }
    return r;
}

// Broyden's "good" method exactly as bdf.rs:414-475 / roots/mod.rs:289-337.
// G: bool g(const double* x, double* out).  h = finite-difference step.
// central = false reproduces `(above + below) * denom` (D4).
template <int D, class G>
inline int broyden_secant(G&& g, const Vec<D>& initial, double h, double tol, int n_max,
                          bool central, Vec<D>& result, uint64_t* n_iter = nullptr) {
    int n = 2;                   // bdf.rs:418
    Vec<D> guess = initial;      // :420
    Vec<D> derivative;
    if (!g(guess.data(), derivative.data())) return ST_USER;  // :421-426

    Mat<D> jac;  // jac_finite_diff, bdf.rs:390-411
    {
        const double denom = 1.0 / (2.0 * h);  // :399
        for (int ind = 0; ind < D; ++ind) {
            Vec<D> above, below;
            guess[ind] += h;  // :402
            if (!g(guess.data(), above.data())) return ST_USER;
            guess[ind] -= 2.0 * h;  // :404
            if (!g(guess.data(), below.data())) return ST_USER;
            guess[ind] += h;  // :406
            for (int r = 0; r < D; ++r)  // :407
                jac.m[r][ind] = central ? (above[r] - below[r]) * denom : (above[r] + below[r]) * denom;
        }
    }
    Mat<D> jac_inv;
    if (!inverse_chain<D>(jac, jac_inv)) return ST_SINGULAR;  // :429-444

    Vec<D> shift = neg_matvec<D>(jac_inv, derivative);  // :446
    for (int d = 0; d < D; ++d) guess[d] += shift[d];   // :447

    while (n < n_max) {  // :449
        if (n_iter) (*n_iter)++;
        const Vec<D> derivative_last = derivative;  // :450
        if (!g(guess.data(), derivative.data())) return ST_USER;  // :451-456
        Vec<D> difference;
        for (int d = 0; d < D; ++d) difference[d] = derivative[d] - derivative_last[d];  // :458
        const Vec<D> adjustment = neg_matvec<D>(jac_inv, difference);                    // :459
        double p = 0.0;  // :461  (-s^T * adjustment)[(0,0)]
        for (int d = 0; d < D; ++d) p += (-shift[d]) * adjustment[d];
        double u[D];  // :462  s^T * jac_inv  (row vector)
        for (int c = 0; c < D; ++c) {
            double s = 0.0;
            for (int r = 0; r < D; ++r) s += shift[r] * jac_inv.m[r][c];
            u[c] = s;
        }
        for (int r = 0; r < D; ++r)  // :464  jac_inv += (shift + adjustment) * u / p
            for (int c = 0; c < D; ++c) jac_inv.m[r][c] += ((shift[r] + adjustment[r]) * u[c]) / p;
        shift = neg_matvec<D>(jac_inv, derivative);        // :465
        for (int d = 0; d < D; ++d) guess[d] += shift[d];  // :466
        if (norm2<D>(shift) <= tol) {                      // :468
            result = guess;
            return ST_OK;
        }
        n += 1;  // :471
    }
    return ST_MAX_ITER;  // :474
}

// ------------------------------------------------------------------------
// BDF coefficients (bdf.rs:641-673 BDF6, :708-730 BDF2)
// ------------------------------------------------------------------------
template <int O> struct BdfCoefficients { double higher[O]; double lower[O]; };

inline BdfCoefficients<7> coefficients_bdf6() {
    BdfCoefficients<7> C{};
    const double h[7] = {60.0 / 147.0, -360.0 / 147.0, 450.0 / 147.0, -400.0 / 147.0,
                         225.0 / 147.0, -72.0 / 147.0, 10.0 / 147.0};
    const double l[7] = {60.0 / 137.0, -300.0 / 137.0, 300.0 / 137.0, -200.0 / 137.0,
                         75.0 / 137.0, -12.0 / 137.0, 0.0};
    for (int i = 0; i < 7; ++i) { C.higher[i] = h[i]; C.lower[i] = l[i]; }
    return C;
}
inline BdfCoefficients<3> coefficients_bdf2() {
    BdfCoefficients<3> C{};
    const double h[3] = {2.0 / 3.0, -4.0 / 3.0, 1.0 / 3.0};
    const double l[3] = {1.0, -1.0, 0.0};
    for (int i = 0; i < 3; ++i) { C.higher[i] = h[i]; C.lower[i] = l[i]; }
    return C;
}

// ------------------------------------------------------------------------
// BDFSolver (bdf.rs:72-122 fields, :346-387 RK4 start-up, :495-634 step)
// ------------------------------------------------------------------------
template <int D, int O, class Rhs> struct BDFSolver {
    double dt_max, dt_min, time, end, tolerance;
    double dt;
    Vec<D> state;
    BdfCoefficients<O> coef;
    std::deque<std::pair<double, Vec<D>>> prev_values;
    Vec<D> save_state;
    double one_tenth, one_sixth, half, two, order;
    size_t yield_memory = 0;
    Rhs rhs;
    const double* params;
    Mode mode;
    bool newton = false;  // NOT in the reference: Newton + LU on the analytic Jacobian (north-star item 4),
                          // kept here so the device Newton path has a like-for-like CPU counterpart
    Counters cnt;
    int fail_code = 0;
    // value of the last Ok(...) (what the iterator yields)
    double out_t = 0.0;
    Vec<D> out_y{};

    // BDF::solve (bdf.rs:257-331)
    BDFSolver(const BdfCoefficients<O>& C, Rhs f, const double* p, const double* y0, double t0,
              double t1, double dtmin, double dtmax, double tol, Mode m)
        : dt_max(dtmax), dt_min(dtmin), time(t0), end(t1), tolerance(tol), coef(C), rhs(f), params(p),
          mode(m) {
        two = 2.0;
        half = 1.0 / two;           // :267 two.recip()
        one_sixth = 1.0 / 6.0;      // :268-270
        one_tenth = 1.0 / 10.0;     // :271-273
        order = static_cast<double>(O);  // :293
        dt = (dtmax + dtmin) * half;     // :302
        for (int d = 0; d < D; ++d) state[d] = y0[d];
        save_state.fill(0.0);
    }

    bool f(double t, const double* y, double* dy) {
        cnt.n_rhs++;
        return rhs(t, y, params, dy);
    }

    // bdf.rs:346-387
    bool runge_kutta(int iterations) {
        for (int i = 0; i < iterations; ++i) {
            Vec<D> k1, k2, k3, k4, inter;
            double dy[D];
            if (!f(time, state.data(), dy)) return false;  // :348-352
            for (int d = 0; d < D; ++d) k1[d] = dy[d] * dt;
            for (int d = 0; d < D; ++d) inter[d] = state[d] + k1[d] * half;  // :353
            if (!f(time + half * dt, inter.data(), dy)) return false;       // :355-359
            for (int d = 0; d < D; ++d) k2[d] = dy[d] * dt;
            for (int d = 0; d < D; ++d) inter[d] = state[d] + k2[d] * half;  // :360
            if (!f(time + half * dt, inter.data(), dy)) return false;       // :362-366
            for (int d = 0; d < D; ++d) k3[d] = dy[d] * dt;
            for (int d = 0; d < D; ++d) inter[d] = state[d] + k3[d];        // :367
            if (!f(time + dt, inter.data(), dy)) return false;              // :369-373
            for (int d = 0; d < D; ++d) k4[d] = dy[d] * dt;
            if (i != 0) prev_values.push_back({time, state});               // :375-378
            for (int d = 0; d < D; ++d)                                      // :380
                state[d] += (((k1[d] + k2[d] * two) + k3[d] * two) + k4[d]) * one_sixth;
            time += dt;  // :381
        }
        prev_values.push_back({time, state});  // :383-384
        return true;
    }

    // higher_func / lower_func closures (bdf.rs:548-575)
    bool g_eval(bool higher, double t, const double* y, double* out) {
        double dy[D];
        if (!f(t, y, dy)) return false;
        const double beta = higher ? coef.higher[0] : coef.lower[0];  // :553 / :567
        // :568 loops over higher_coefficients in BOTH closures (D5)
        const double* hist = (higher || mode == Mode::Literal) ? coef.higher : coef.lower;
        Vec<D> sp;
        for (int d = 0; d < D; ++d) sp[d] = ((-dy[d]) * dt) * beta;
        for (int ind = 1; ind < O; ++ind)  // :554-556
            for (int d = 0; d < D; ++d) sp[d] += prev_values[O - ind].second[d] * hist[ind];
        for (int d = 0; d < D; ++d) out[d] = sp[d] + y[d];  // :557-560
        return true;
    }

    // Newton on g(x) = 0 with g'(x) = I - dt*beta*J_f(t_{n+1}, x), dense LU with partial pivoting,
    // stop when ||delta||_2 <= tol (same rule as bdf.rs:468), at most 998 iterations.
    int newton_solve(bool higher, Vec<D>& result) {
        const double tg = time + dt;
        const double beta = higher ? coef.higher[0] : coef.lower[0];
        Vec<D> x = state;
        for (int n = 2; n < 1000; ++n) {
            cnt.n_iter++;
            double g[D];
            if (!g_eval(higher, tg, x.data(), g)) return ST_USER;
            double J[D][D];
            rhs.jac(tg, x.data(), params, &J[0][0]);
            double M[D][D], b[D];
            for (int r = 0; r < D; ++r) {
                b[r] = -g[r];
                for (int c = 0; c < D; ++c) M[r][c] = (r == c ? 1.0 : 0.0) - (dt * beta) * J[r][c];
            }
            for (int i = 0; i < D; ++i) {
                int piv = i;
                double best = std::fabs(M[i][i]);
                for (int r = i + 1; r < D; ++r)
                    if (std::fabs(M[r][i]) > best) { best = std::fabs(M[r][i]); piv = r; }
                if (best == 0.0) return ST_SINGULAR;
                if (piv != i) {
                    for (int c = 0; c < D; ++c) std::swap(M[i][c], M[piv][c]);
                    std::swap(b[i], b[piv]);
                }
                const double inv_diag = 1.0 / M[i][i];
                for (int r = i + 1; r < D; ++r) {
                    const double l = M[r][i] * inv_diag;
                    for (int c = i + 1; c < D; ++c) M[r][c] -= l * M[i][c];
                    b[r] -= l * b[i];
                }
            }
            double ss = 0.0;
            for (int i = D - 1; i >= 0; --i) {
                double acc = b[i];
                for (int c = i + 1; c < D; ++c) acc -= M[i][c] * b[c];
                b[i] = acc / M[i][i];
            }
            for (int d = 0; d < D; ++d) { x[d] += b[d]; ss += b[d] * b[d]; }
            if (ss <= tolerance * tolerance) { result = x; return ST_OK; }
        }
        return ST_MAX_ITER;
    }

    int secant(bool higher, Vec<D>& result) {
        if (newton) return newton_solve(higher, result);
        // bdf.rs:403,405,423,454 pass self.time (t_n) to g (D7); intent t_{n+1}
        const double tg = (mode == Mode::Literal) ? time : time + dt;
        auto g = [&](const double* x, double* out) { return g_eval(higher, tg, x, out); };
        return broyden_secant<D>(g, state, dt, tolerance, 1000, mode != Mode::Literal, result,
                                 &cnt.n_iter);
    }

    // bdf.rs:495-634
    StepKind step() {
        cnt.n_steps++;
        if (yield_memory > 0 && yield_memory <= (size_t)O) {  // :500-512
            const size_t get_item = O - yield_memory;
            yield_memory -= 1;
            if (yield_memory == 0) yield_memory = O + 2;
            out_t = prev_values[get_item].first;
            out_y = prev_values[get_item].second;
            cnt.n_accept++;
            return StepKind::Ok;
        }
        if (yield_memory == (size_t)O + 2) {  // :519-525
            yield_memory = 0;
            prev_values.push_back({time, state});
            prev_values.pop_front();
            out_t = time; out_y = state;
            cnt.n_accept++;
            return StepKind::Ok;
        }
        if (time >= end) return StepKind::Done;  // :527-529

        if (time + dt >= end) {  // :531-535
            dt = end - time;
            if (!runge_kutta(1)) { fail_code = ST_USER; return StepKind::Failure; }
            out_t = time; out_y = prev_values.back().second;
            cnt.n_accept++;
            return StepKind::Ok;
        }

        if (prev_values.empty()) {  // :537-546
            save_state = state;
            if (time + dt * order >= end) dt = (end - time) / order;
            if (!runge_kutta(O)) { fail_code = ST_USER; return StepKind::Failure; }
            yield_memory = O + 1;
            return StepKind::Redo;
        }

        cnt.n_implicit++;
        Vec<D> higher_step, lower_step;
        int rc = secant(true, higher_step);  // :577
        if (rc != ST_OK) { fail_code = rc; return StepKind::Failure; }
        rc = secant(false, lower_step);      // :578
        if (rc != ST_OK) { fail_code = rc; return StepKind::Failure; }

        Vec<D> difference;
        for (int d = 0; d < D; ++d) difference[d] = higher_step[d] - lower_step[d];  // :580
        const double error = norm2<D>(difference);                                    // :581

        if (error <= tolerance) {  // :583
            state = higher_step;
            time += dt;
            if (yield_memory == (size_t)O + 1) {  // :593-596
                yield_memory -= 1;
                return StepKind::Redo;
            }
            prev_values.push_back({time, state});  // :598-600
            prev_values.pop_front();
            if (error < one_tenth * tolerance) {  // :602-611
                dt *= two;
                if (dt > dt_max) dt = dt_max;
                prev_values.clear();
            }
            out_t = time; out_y = state;
            cnt.n_accept++;
            return StepKind::Ok;  // :613
        }

        cnt.n_reject++;
        if (yield_memory == (size_t)O + 1) {  // :620-624
            if (mode == Mode::Literal) time -= dt - order;  // :622 as written (D6)
            else time -= dt * order;
            state = save_state;
        }
        dt *= half;  // :626
        if (dt < dt_min) {  // :628-630
            fail_code = ST_MIN_DT;
            return StepKind::Failure;
        }
        prev_values.clear();  // :632
        return StepKind::Redo;
    }
};

// ------------------------------------------------------------------------
// Adams predictor-corrector coefficients (adams.rs:635-650 Adams5, :695-712 Adams3)
//   predictor: Adams-Bashforth weights, newest derivative first, padded with a zero
//   corrector: Adams-Moulton weights, implicit derivative first
//   error_coefficient: 19/270 for BOTH (adams.rs:652-654, :714-716)
// ------------------------------------------------------------------------
template <int O> struct AdamsCoefficients { double predictor[O]; double corrector[O]; double error; };

inline AdamsCoefficients<5> coefficients_adams5() {
    AdamsCoefficients<5> c{};
    const double tf = 24.0, s = 720.0;
    const double p[5] = {55.0 / tf, -59.0 / tf, 37.0 / tf, -9.0 / tf, 0.0};
    const double q[5] = {251.0 / s, 646.0 / s, -264.0 / s, 106.0 / s, -19.0 / s};
    for (int i = 0; i < 5; ++i) { c.predictor[i] = p[i]; c.corrector[i] = q[i]; }
    c.error = 19.0 / 270.0;
    return c;
}
inline AdamsCoefficients<3> coefficients_adams3() {
    AdamsCoefficients<3> c{};
    const double p[3] = {1.0 + 1.0 / 2.0, -(1.0 / 2.0), 0.0};            // adams.rs:697-703
    const double q[3] = {5.0 / 12.0, 2.0 / 3.0, -(1.0 / 12.0)};          // adams.rs:705-711
    for (int i = 0; i < 3; ++i) { c.predictor[i] = p[i]; c.corrector[i] = q[i]; }
    c.error = 19.0 / 270.0;
    return c;
}

// ------------------------------------------------------------------------
// AdamsSolver (adams.rs:72-122 fields, :340-394 RK4 start-up, :398-566 step).
// One defect on this path (D10, see step()): Literal = as written, Corrected = D10 repaired.
// ------------------------------------------------------------------------
template <int D, int O, class Rhs> struct AdamsSolver {
    double dt_max, dt_min, time, end, tolerance;
    double dt;
    Vec<D> state;
    AdamsCoefficients<O> coef;
    std::deque<std::pair<double, Vec<D>>> prev_values;
    std::deque<Vec<D>> prev_derivatives;
    Vec<D> implicit_derivs{}, save_state{};
    double one_tenth, one_sixth, half, two, four, order;
    size_t yield_memory = 0;
    Rhs rhs;
    const double* params;
    PowMode pow_mode;  // q = (tol / (2 error))^(1/order): libm pow, as f64::powf (adams.rs:527, :553)
    Mode mode = Mode::Corrected;
    Counters cnt;
    int fail_code = 0;
    double out_t = 0.0;
    Vec<D> out_y{};

    // Adams::solve (adams.rs:249-337)
    AdamsSolver(const AdamsCoefficients<O>& C, Rhs f, const double* p, const double* y0, double t0, double t1,
                double dtmin, double dtmax, double tol, PowMode pm)
        : dt_max(dtmax), dt_min(dtmin), time(t0), end(t1), tolerance(tol), coef(C), rhs(f), params(p),
          pow_mode(pm) {
        two = 2.0;
        half = 1.0 / two;        // :259
        one_sixth = 1.0 / 6.0;   // :260-262
        one_tenth = 1.0 / 10.0;  // :263-265
        four = two * two;        // :266
        order = static_cast<double>(O);  // :288
        dt = (dtmax + dtmin) * half;     // :297
        for (int d = 0; d < D; ++d) state[d] = y0[d];
    }

    bool f(double t, const double* y, double* dy) {
        cnt.n_rhs++;
        return rhs(t, y, params, dy);
    }
    double root(double x) const {  // x.powf(order.recip()) (adams.rs:526-527, :544-545)
        return pow_mode == PowMode::LibmPow ? std::pow(x, 1.0 / order) : det_root<O>(x);
    }

    // adams.rs:340-394: like bdf.rs:346-387, and additionally stores f(t, y) of every stored point
    bool runge_kutta(int iterations) {
        for (int i = 0; i < iterations; ++i) {
            Vec<D> k1, k2, k3, k4, inter;
            double dy[D];
            if (!f(time, state.data(), dy)) return false;
            for (int d = 0; d < D; ++d) k1[d] = dy[d] * dt;
            for (int d = 0; d < D; ++d) inter[d] = state[d] + k1[d] * half;
            if (!f(time + half * dt, inter.data(), dy)) return false;
            for (int d = 0; d < D; ++d) k2[d] = dy[d] * dt;
            for (int d = 0; d < D; ++d) inter[d] = state[d] + k2[d] * half;
            if (!f(time + half * dt, inter.data(), dy)) return false;
            for (int d = 0; d < D; ++d) k3[d] = dy[d] * dt;
            for (int d = 0; d < D; ++d) inter[d] = state[d] + k3[d];
            if (!f(time + dt, inter.data(), dy)) return false;
            for (int d = 0; d < D; ++d) k4[d] = dy[d] * dt;
            if (i != 0) {  // :372-380
                Vec<D> der;
                if (!f(time, state.data(), der.data())) return false;
                prev_derivatives.push_back(der);
                prev_values.push_back({time, state});
            }
            for (int d = 0; d < D; ++d)  // :382
                state[d] += (((k1[d] + k2[d] * two) + k3[d] * two) + k4[d]) * one_sixth;
            time += dt;  // :383
        }
        Vec<D> der;  // :385-391
        if (!f(time, state.data(), der.data())) return false;
        prev_derivatives.push_back(der);
        prev_values.push_back({time, state});
        return true;
    }

    // adams.rs:410-566
    StepKind step() {
        cnt.n_steps++;
        if (yield_memory > 0 && yield_memory < (size_t)O) {  // :415-427
            const size_t get_item = O - yield_memory - 1;
            yield_memory -= 1;
            if (yield_memory == 0) yield_memory = O + 1;
            out_t = prev_values[get_item].first;
            out_y = prev_values[get_item].second;
            cnt.n_accept++;
            return StepKind::Ok;
        }
        if (yield_memory == (size_t)O + 1) {  // :434-440
            yield_memory = 0;
            prev_values.push_back({time, state});
            prev_values.pop_front();
            out_t = time; out_y = state;
            cnt.n_accept++;
            return StepKind::Ok;
        }
        if (time >= end) return StepKind::Done;  // :442-444

        if (time + dt >= end) {  // :446-450
            dt = end - time;
            if (!runge_kutta(1)) { fail_code = ST_USER; return StepKind::Failure; }
            out_t = time; out_y = prev_values.back().second;
            cnt.n_accept++;
            return StepKind::Ok;
        }

        if (prev_values.empty()) {  // :452-463
            save_state = state;
            if (time + dt * (order - 1.0) >= end) dt = (end - time) / (order - 1.0);
            if (!runge_kutta(O - 1)) { fail_code = ST_USER; return StepKind::Failure; }
            yield_memory = O;
            return StepKind::Redo;
        }

        // predictor (Adams-Bashforth) :465-470
        Vec<D> sp, predictor, corrector;
        for (int d = 0; d < D; ++d) sp[d] = prev_derivatives[0][d] * coef.predictor[O - 2];
        for (int i = 1; i < O - 1; ++i) {
            const double c = coef.predictor[O - i - 2];
            for (int d = 0; d < D; ++d) sp[d] += prev_derivatives[i][d] * c;
        }
        for (int d = 0; d < D; ++d) predictor[d] = state[d] + sp[d] * dt;

        // corrector (Adams-Moulton) :472-483
        if (!f(time + dt, predictor.data(), implicit_derivs.data())) { fail_code = ST_USER; return StepKind::Failure; }
        for (int d = 0; d < D; ++d) sp[d] = implicit_derivs[d] * coef.corrector[0];
        for (int i = 0; i < O - 1; ++i) {
            const double c = coef.corrector[O - i - 1];
            for (int d = 0; d < D; ++d) sp[d] += prev_derivatives[i][d] * c;
        }
        for (int d = 0; d < D; ++d) corrector[d] = state[d] + sp[d] * dt;

        Vec<D> difference;
        for (int d = 0; d < D; ++d) difference[d] = corrector[d] - predictor[d];  // :485
        const double error = coef.error / dt * norm2<D>(difference);             // :486

        if (std::isnan(error)) {  // same hazard as D8: neither branch below can make progress
            fail_code = ST_NONFINITE;
            return StepKind::Failure;
        }

        if (error <= tolerance) {  // :488
            state = corrector;
            time += dt;
            if (yield_memory == (size_t)O) {  // :498-501
                yield_memory -= 1;
                // D10 (found while restating; not in SURVEY.md's table): the early return skips the push below,
                // although the sentinel branch (:429-440) assumes "the derivatives memory deque already has the
                // derivatives for this step".  As written the step after every warm-up block therefore
                // extrapolates with derivatives that lag one step behind: an O(dt) error estimate, a reject,
                // a smaller dt and a new warm-up — the step count grows like 1/tol.  Corrected = push it.
                if (mode == Mode::Corrected) {
                    prev_derivatives.push_back(implicit_derivs);
                    prev_derivatives.pop_front();
                }
                return StepKind::Redo;
            }
            prev_derivatives.push_back(implicit_derivs);  // :503-509
            prev_values.push_back({time, state});
            prev_values.pop_front();
            prev_derivatives.pop_front();

            if (error < one_tenth * tolerance) {  // :511-529
                const double q = root(tolerance / (two * error));
                if (q > four) dt *= four;
                else dt *= q;
                if (dt > dt_max) dt = dt_max;
                prev_values.clear();
                prev_derivatives.clear();
            }
            out_t = time; out_y = state;
            cnt.n_accept++;
            return StepKind::Ok;  // :531
        }

        cnt.n_reject++;
        if (yield_memory == (size_t)O) {  // :538-542
            time -= dt * (order - 1.0);
            state = save_state;
        }
        const double q = root(tolerance / (two * error));  // :544-545
        if (q < one_tenth) dt *= one_tenth;  // :547-551
        else dt *= q;
        if (dt < dt_min) {  // :553-555
            fail_code = ST_MIN_DT;
            return StepKind::Failure;
        }
        prev_values.clear();  // :557-558
        prev_derivatives.clear();
        return StepKind::Redo;
    }
};

// ------------------------------------------------------------------------
// EulerSolver (ivp.rs:306-344): fixed step, yields the OLD (time, state) of every step —
// the initial condition is the first point and the final state is never yielded.
// Builder (ivp.rs:386-471): with_tolerance is a no-op (:390-392); each of with_maximum_dt /
// with_minimum_dt sets dt, or averages with the dt already set (:396-421).
// ------------------------------------------------------------------------
template <int D, class Rhs> struct EulerSolver {
    double dt, time, end;
    Vec<D> state;
    Rhs rhs;
    const double* params;
    Counters cnt;
    int fail_code = 0;
    double out_t = 0.0;
    Vec<D> out_y{};

    EulerSolver(Rhs f, const double* p, const double* y0, double t0, double t1, double step)
        : dt(step), time(t0), end(t1), rhs(f), params(p) {
        for (int d = 0; d < D; ++d) state[d] = y0[d];
    }

    StepKind step() {
        cnt.n_steps++;
        if (time >= end) return StepKind::Done;       // ivp.rs:321-323
        if (time + dt >= end) dt = end - time;        // :324-326
        double dy[D];
        cnt.n_rhs++;
        if (!rhs(time, state.data(), params, dy)) {   // :328-329
            fail_code = ST_USER;
            return StepKind::Failure;
        }
        out_t = time;                                  // :331-332
        out_y = state;
        for (int d = 0; d < D; ++d) state[d] += dy[d] * dt;  // :334
        time += dt;                                    // :335
        cnt.n_accept++;
        return StepKind::Ok;                           // :337
    }
};

// ------------------------------------------------------------------------
// Cubic Hermite interpolant between two knots and the zero of a scalar Hermite cubic: CPU statement of
// bacon_b200/csrc/path_query.cuh (hermite_eval, hermite_root), operation for operation.  NOT in the reference
// (its Path is the accepted points, ivp.rs:203-211); used by the path queries (oracle_capi.cpp) and by the
// terminal event of the drive loop below.
// ------------------------------------------------------------------------
template <int D>
inline void hermite_eval(double th, double h, const double* ya, const double* yb, const double* fa, const double* fb,
                         double* out) {
    const double om = 1.0 - th, tt = th * (th - 1.0), c0 = 1.0 - 2.0 * th, c1 = th - 1.0;
    for (int d = 0; d < D; ++d) {
        const double dy = yb[d] - ya[d];
        const double v = (c0 * dy + c1 * (h * fa[d])) + th * (h * fb[d]);
        out[d] = (om * ya[d] + th * yb[d]) + tt * v;
    }
}

// zero in [0, 1] of the scalar Hermite cubic: safeguarded Newton, as path_query.cuh (hermite_root)
inline double hermite_root(double ga, double gb, double A, double B) {
    if (gb == 0.0) return 1.0;
    const double dg = gb - ga, vs = (A + B) - 2.0 * dg;
    double lo = 0.0, hi = 1.0;
    double th = ga / (ga - gb);
    for (int it = 0; it < 60; ++it) {
        const double tm1 = th - 1.0;
        const double v = ((1.0 - 2.0 * th) * dg + tm1 * A) + th * B;
        const double val = ((1.0 - th) * ga + th * gb) + (th * tm1) * v;
        if (val == 0.0) break;
        if ((val < 0.0) == (ga < 0.0)) lo = th;
        else hi = th;
        const double der = (dg + (2.0 * th - 1.0) * v) + (th * tm1) * vs;
        double tn = th - val / der;
        if (!(tn > lo && tn < hi)) tn = 0.5 * (lo + hi);
        const double moved = std::fabs(tn - th);
        th = tn;
        if (moved <= 1e-15) break;
    }
    return th;
}


// ------------------------------------------------------------------------
// IVPIterator drive loop (ivp.rs:220-238) + collect_vec (ivp.rs:209-211)
// ------------------------------------------------------------------------
template <int D> struct Solution {
    int status = ST_OK;
    double t_end = 0.0;  // IVPStepper::time() when the loop stopped
    double dt_end = 0.0;
    Vec<D> y_end{};      // stepper state when the loop stopped
    Counters cnt;
    std::vector<double> path_t;        // yielded times
    std::vector<Vec<D>> path_y;        // yielded states
};

// Optional inputs of a solve (include/bacon_ivp.h: bacon_ivp_config::dt_init, bacon_ivp_options): a first dt other
// than the reference's (dt_max + dt_min)/2 — clamped into [dt_min, dt_max] — and a terminal event g(y) = w . y - c.
// NOT in the reference; the statement of what the device kernels compiled for them do (drive.cuh: EventWatch).
struct DriveOpts {
    bool set_dt = false;
    double dt = 0.0;
    const double* ev_w = nullptr;
    double ev_c = 0.0;
    int ev_dir = 0;
};
inline double clamp_first_dt(double d, double dt_min, double dt_max) {
    return !(d >= dt_min) ? dt_min : (d > dt_max ? dt_max : d);
}

// `point(s, t, y)` reads the point a step() == Ok yielded.
template <int D, class Stepper, class Point>
inline void drive(Stepper& s, uint64_t max_attempts, bool keep_path, Solution<D>& sol, Point&& point,
                  const DriveOpts* o = nullptr) {
    const uint64_t cap = max_attempts ? max_attempts : 0xFFFFFFFEull;
    sol.status = ST_OK;
    const bool ev = o && o->ev_w;
    auto g_of = [&](const Vec<D>& y) {
        double acc = o->ev_w[0] * y[0];
        for (int d = 1; d < D; ++d) acc += o->ev_w[d] * y[d];
        return acc - o->ev_c;
    };
    double tp = s.time, gp = 0.0;  // the last knot (knot 0 = the initial condition)
    Vec<D> yp = s.state;
    if (ev) gp = g_of(yp);
    for (;;) {
        if (s.cnt.n_steps >= cap) { sol.status = ST_MAX_ATTEMPTS; break; }
        const StepKind k = s.step();
        if (k == StepKind::Ok) {
            double t;
            Vec<D> y;
            point(s, t, y);
            if (ev) {
                const double gb = g_of(y);
                const bool rising = gp < 0.0 && gb >= 0.0, falling = gp > 0.0 && gb <= 0.0;
                if (o->ev_dir > 0 ? rising : (o->ev_dir < 0 ? falling : (rising || falling))) {
                    // the trajectory ends at the crossing, located on the Hermite cubic of this interval; the point that
                    // crossed is not yielded
                    double fa[D], fb[D];
                    s.rhs(tp, yp.data(), s.params, fa);
                    s.rhs(t, y.data(), s.params, fb);
                    const double h = t - tp;
                    double da = o->ev_w[0] * fa[0], db = o->ev_w[0] * fb[0];
                    for (int d = 1; d < D; ++d) {
                        da += o->ev_w[d] * fa[d];
                        db += o->ev_w[d] * fb[d];
                    }
                    const double th = hermite_root(gp, gb, h * da, h * db);
                    hermite_eval<D>(th, h, yp.data(), y.data(), fa, fb, sol.y_end.data());
                    sol.t_end = tp + th * h;
                    sol.dt_end = s.dt;
                    sol.cnt = s.cnt;
                    sol.cnt.n_accept -= 1;
                    sol.status = ST_STOPPED_AT_EVENT;
                    return;
                }
                gp = gb;
                tp = t;
                yp = y;
            }
            if (keep_path) {
                sol.path_t.push_back(t);
                sol.path_y.push_back(y);
            }
            continue;
        }
        if (k == StepKind::Redo) continue;
        if (k == StepKind::Done) break;
        sol.status = s.fail_code;  // Failure: yielded once, then the iterator is finished
        break;
    }
    sol.t_end = s.time;
    sol.dt_end = s.dt;
    sol.y_end = s.state;
    sol.cnt = s.cnt;
}

template <int D, int O, class Rhs>
inline Solution<D> solve_rk(const RkTableau<O>& T, Rhs rhs, const double* params, const double* y0,
                            double t0, double t1, double dtmin, double dtmax, double tol,
                            PowMode pm, uint64_t max_attempts, bool keep_path, const DriveOpts* o = nullptr) {
    RungeKuttaSolver<D, O, Rhs> s(T, rhs, params, y0, t0, t1, dtmin, dtmax, tol, pm);
    if (o && o->set_dt) s.dt = clamp_first_dt(o->dt, dtmin, dtmax);
    Solution<D> sol;
    drive<D>(s, max_attempts, keep_path, sol, [](RungeKuttaSolver<D, O, Rhs>& st, double& t, Vec<D>& y) {
        t = st.time;    // rk.rs:419
        y = st.state;
    }, o);
    return sol;
}

template <int D, int O, class Rhs>
inline Solution<D> solve_bdf(const BdfCoefficients<O>& C, Rhs rhs, const double* params,
                             const double* y0, double t0, double t1, double dtmin, double dtmax,
                             double tol, Mode mode, uint64_t max_attempts, bool keep_path,
                             bool newton = false, const DriveOpts* o = nullptr) {
    BDFSolver<D, O, Rhs> s(C, rhs, params, y0, t0, t1, dtmin, dtmax, tol, mode);
    s.newton = newton;
    if (o && o->set_dt) s.dt = clamp_first_dt(o->dt, dtmin, dtmax);
    Solution<D> sol;
    drive<D>(s, max_attempts, keep_path, sol, [](BDFSolver<D, O, Rhs>& st, double& t, Vec<D>& y) {
        t = st.out_t;
        y = st.out_y;
    }, o);
    return sol;
}

template <int D, int O, class Rhs>
inline Solution<D> solve_adams(const AdamsCoefficients<O>& C, Rhs rhs, const double* params, const double* y0,
                               double t0, double t1, double dtmin, double dtmax, double tol, PowMode pm,
                               Mode mode, uint64_t max_attempts, bool keep_path, const DriveOpts* o = nullptr) {
    AdamsSolver<D, O, Rhs> s(C, rhs, params, y0, t0, t1, dtmin, dtmax, tol, pm);
    s.mode = mode;
    if (o && o->set_dt) s.dt = clamp_first_dt(o->dt, dtmin, dtmax);
    Solution<D> sol;
    drive<D>(s, max_attempts, keep_path, sol, [](AdamsSolver<D, O, Rhs>& st, double& t, Vec<D>& y) {
        t = st.out_t;
        y = st.out_y;
    }, o);
    return sol;
}

// Euler: `dt` is what the builder arrives at (ivp.rs:396-421); the C API passes (dt_max + dt_min)/2 when both
// were given, or the one that was.  (A first dt in DriveOpts is not used: Euler's step is the builder's.)
template <int D, class Rhs>
inline Solution<D> solve_euler(Rhs rhs, const double* params, const double* y0, double t0, double t1, double dt,
                               uint64_t max_attempts, bool keep_path, const DriveOpts* o = nullptr) {
    EulerSolver<D, Rhs> s(rhs, params, y0, t0, t1, dt);
    Solution<D> sol;
    drive<D>(s, max_attempts, keep_path, sol, [](EulerSolver<D, Rhs>& st, double& t, Vec<D>& y) {
        t = st.out_t;
        y = st.out_y;
    }, o);
    return sol;
}

}  // namespace bacon_oracle
