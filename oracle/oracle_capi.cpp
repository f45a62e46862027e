// oracle_capi.cpp — C entry points of the CPU ORACLE (test infrastructure only).
//
// Wraps bacon_oracle.hpp (the restatement of src/ivp/rk.rs, src/ivp/bdf.rs, src/ivp/adams.rs,
// the Euler stepper src/ivp.rs:306-344 and the drive loop src/ivp.rs:220-238) behind the same structs as include/bacon_ivp.h so tests can
// run the oracle and the CUDA path on identical buffers.  OpenMP over
// trajectories = the "rayon over trajectories" CPU baseline of BASELINE.md
// (kind "port": the Rust reference cannot be compiled in this image).
//
// Also exports roots::secant (src/roots/mod.rs:289-337) on the reference's own
// test functions (src/tests/roots/mod.rs:41-83) so the shared Broyden + LU code
// can be pinned against the reference's known answers.
#include <cmath>
#include <cstring>
#include <string>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/bacon_ivp.h"
#include "bacon_oracle.hpp"

namespace bo = bacon_oracle;

// ---- right-hand sides.  Expression trees are written out explicitly; the
// device functors in bacon_b200/csrc/rhs_builtin.cuh use the same trees so a
// strict-FP (no FMA) device build is bit-comparable. ----------------------
struct RhsLorenz {  // p = (sigma, rho, beta)
    static constexpr int DIM = 3, NPARAM = 3;
    bool operator()(double, const double* y, const double* p, double* dy) const {
        dy[0] = p[0] * (y[1] - y[0]);
        dy[1] = y[0] * (p[1] - y[2]) - y[1];
        dy[2] = y[0] * y[1] - p[2] * y[2];
        return true;
    }
    void jac(double, const double* y, const double* p, double* J) const {
        J[0] = -p[0]; J[1] = p[0]; J[2] = 0.0;
        J[3] = p[1] - y[2]; J[4] = -1.0; J[5] = -y[0];
        J[6] = y[1]; J[7] = y[0]; J[8] = -p[2];
    }
};
struct RhsVdp {  // p = (mu)
    static constexpr int DIM = 2, NPARAM = 1;
    bool operator()(double, const double* y, const double* p, double* dy) const {
        dy[0] = y[1];
        dy[1] = (p[0] * (1.0 - y[0] * y[0])) * y[1] - y[0];
        return true;
    }
    void jac(double, const double* y, const double* p, double* J) const {
        J[0] = 0.0; J[1] = 1.0;
        J[2] = -2.0 * p[0] * y[0] * y[1] - 1.0; J[3] = p[0] * (1.0 - y[0] * y[0]);
    }
};
struct RhsRobertson {  // p = (k1, k2, k3)
    static constexpr int DIM = 3, NPARAM = 3;
    bool operator()(double, const double* y, const double* p, double* dy) const {
        const double a = p[0] * y[0];
        const double b = (p[2] * y[1]) * y[2];
        const double c = (p[1] * y[1]) * y[1];
        dy[0] = b - a;
        dy[1] = (a - b) - c;
        dy[2] = c;
        return true;
    }
    void jac(double, const double* y, const double* p, double* J) const {
        const double k3y2 = p[2] * y[2], k3y1 = p[2] * y[1], k2y1 = 2.0 * p[1] * y[1];
        J[0] = -p[0]; J[1] = k3y2; J[2] = k3y1;
        J[3] = p[0]; J[4] = -k3y2 - k2y1; J[5] = -k3y1;
        J[6] = 0.0; J[7] = k2y1; J[8] = 0.0;
    }
};
template <int N> struct RhsLinear {  // p = A row-major [N][N]
    static constexpr int DIM = N, NPARAM = N * N;
    bool operator()(double, const double* y, const double* p, double* dy) const {
        for (int i = 0; i < N; ++i) {
            double s = p[i * N] * y[0];
            for (int j = 1; j < N; ++j) s += p[i * N + j] * y[j];
            dy[i] = s;
        }
        return true;
    }
    void jac(double, const double*, const double* p, double* J) const {
        for (int i = 0; i < N * N; ++i) J[i] = p[i];
    }
};
struct RhsExp {  // y' = y   (README.md:26-28, rk.rs:539-541, bdf.rs:769-771)
    static constexpr int DIM = 1, NPARAM = 0;
    bool operator()(double, const double* y, const double*, double* dy) const { dy[0] = y[0]; return true; }
    void jac(double, const double*, const double*, double* J) const { J[0] = 1.0; }
};
struct RhsDecay {  // y' = -y  (bdf.rs:781-783)
    static constexpr int DIM = 1, NPARAM = 0;
    bool operator()(double, const double* y, const double*, double* dy) const { dy[0] = -y[0]; return true; }
    void jac(double, const double*, const double*, double* J) const { J[0] = -1.0; }
};
struct RhsQuadratic {  // y' = -2t (rk.rs:664-666, bdf.rs:773-775)
    static constexpr int DIM = 1, NPARAM = 0;
    bool operator()(double t, const double*, const double*, double* dy) const { dy[0] = -2.0 * t; return true; }
    void jac(double, const double*, const double*, double* J) const { J[0] = 0.0; }
};
struct RhsCos {  // y' = cos t (rk.rs:668-670, bdf.rs:777-779)
    static constexpr int DIM = 1, NPARAM = 0;
    bool operator()(double t, const double*, const double*, double* dy) const { dy[0] = std::cos(t); return true; }
    void jac(double, const double*, const double*, double* J) const { J[0] = 0.0; }
};
struct RhsHarmonic {  // y'' = -w^2 y ; p = (w)
    static constexpr int DIM = 2, NPARAM = 1;
    bool operator()(double, const double* y, const double* p, double* dy) const {
        dy[0] = y[1];
        dy[1] = -(p[0] * p[0]) * y[0];
        return true;
    }
    void jac(double, const double*, const double* p, double* J) const {
        J[0] = 0.0; J[1] = 1.0; J[2] = -(p[0] * p[0]); J[3] = 0.0;
    }
};

struct RunArgs {
    const bacon_ivp_config* cfg;
    bo::PowMode pm;
    size_t n;
    const double* y0;
    const double* params;
    const bacon_ivp_result* out;
};

template <int D>
static void store(const RunArgs& a, size_t i, const bo::Solution<D>& s) {
    const bacon_ivp_result& o = *a.out;
    int status = s.status;
    const int cap = a.cfg->history_capacity;
    if (cap > 0 && o.hist) {
        const size_t np = s.path_t.size();
        const size_t keep = np < (size_t)cap ? np : (size_t)cap;
        for (size_t k = 0; k < keep; ++k) {  // one (t, y) record per yielded point: Vec<(f64, BVector)> (ivp.rs:203)
            double* rec = o.hist + (i * cap + k) * (size_t)(D + 1);
            rec[0] = s.path_t[k];
            for (int d = 0; d < D; ++d) rec[1 + d] = s.path_y[k][d];
        }
        if (o.hist_len) o.hist_len[i] = (uint32_t)keep;
        if (np > (size_t)cap && status == bo::ST_OK) status = bo::ST_HISTORY_OVERFLOW;
    }
    for (int d = 0; d < D; ++d) o.y_end[(size_t)d * a.n + i] = s.y_end[d];
    if (o.t_end) o.t_end[i] = s.t_end;
    if (o.dt_end) o.dt_end[i] = s.dt_end;
    o.status[i] = status;
    if (o.n_accept) o.n_accept[i] = (uint32_t)s.cnt.n_accept;
    if (o.n_reject) o.n_reject[i] = (uint32_t)s.cnt.n_reject;
    if (o.n_rhs) o.n_rhs[i] = (uint32_t)s.cnt.n_rhs;
}

template <class Rhs> static void run_one(const RunArgs& a, size_t i) {
    constexpr int D = Rhs::DIM;
    constexpr int P = Rhs::NPARAM;
    const bacon_ivp_config& c = *a.cfg;
    double y0[D];
    for (int d = 0; d < D; ++d) y0[d] = a.y0[(size_t)d * a.n + i];
    std::vector<double> p(P > 0 ? P : 1);
    const bool shared = (c.flags & BACON_FLAG_SHARED_PARAMS) != 0;
    const bool aos = (c.flags & BACON_FLAG_PARAMS_AOS) != 0;
    for (int k = 0; k < P; ++k)
        p[k] = shared ? a.params[k] : (aos ? a.params[i * (size_t)P + k] : a.params[(size_t)k * a.n + i]);
    const bo::Mode mode = c.semantics == BACON_SEM_LITERAL ? bo::Mode::Literal : bo::Mode::Corrected;
    const bool keep = c.history_capacity > 0;
    const bool newton = (c.flags & BACON_FLAG_BDF_NEWTON) != 0;
    bo::Solution<D> s;
    switch (c.method) {
        case BACON_RK45:
            s = bo::solve_rk<D, 6>(bo::tableau_rkf45(mode), Rhs{}, p.data(), y0, c.t_start, c.t_end,
                                   c.dt_min, c.dt_max, c.tol, a.pm, c.max_attempts, keep);
            break;
        case BACON_RK23:
            s = bo::solve_rk<D, 4>(bo::tableau_bs23(mode), Rhs{}, p.data(), y0, c.t_start, c.t_end,
                                   c.dt_min, c.dt_max, c.tol, a.pm, c.max_attempts, keep);
            break;
        case BACON_BDF6:
            s = bo::solve_bdf<D, 7>(bo::coefficients_bdf6(), Rhs{}, p.data(), y0, c.t_start, c.t_end,
                                    c.dt_min, c.dt_max, c.tol, mode, c.max_attempts, keep, newton);
            break;
        case BACON_BDF2:
            s = bo::solve_bdf<D, 3>(bo::coefficients_bdf2(), Rhs{}, p.data(), y0, c.t_start, c.t_end,
                                    c.dt_min, c.dt_max, c.tol, mode, c.max_attempts, keep, newton);
            break;
        case BACON_ADAMS5:
            s = bo::solve_adams<D, 5>(bo::coefficients_adams5(), Rhs{}, p.data(), y0, c.t_start, c.t_end,
                                      c.dt_min, c.dt_max, c.tol, a.pm, mode, c.max_attempts, keep);
            break;
        case BACON_ADAMS3:
            s = bo::solve_adams<D, 3>(bo::coefficients_adams3(), Rhs{}, p.data(), y0, c.t_start, c.t_end,
                                      c.dt_min, c.dt_max, c.tol, a.pm, mode, c.max_attempts, keep);
            break;
        default:  // BACON_EULER: config.dt_max carries the builder's dt
            s = bo::solve_euler<D>(Rhs{}, p.data(), y0, c.t_start, c.t_end, c.dt_max, c.max_attempts, keep);
            break;
    }
    store<D>(a, i, s);
}

typedef void (*run_fn)(const RunArgs&, size_t);
struct Entry { const char* name; int dim; int n_params; run_fn run; };
static const Entry kTable[] = {
    {"lorenz", 3, 3, run_one<RhsLorenz>},       {"vdp", 2, 1, run_one<RhsVdp>},
    {"robertson", 3, 3, run_one<RhsRobertson>}, {"linear32", 32, 1024, run_one<RhsLinear<32>>},
    {"exp", 1, 0, run_one<RhsExp>},             {"decay", 1, 0, run_one<RhsDecay>},
    {"quadratic", 1, 0, run_one<RhsQuadratic>}, {"cos", 1, 0, run_one<RhsCos>},
    {"harmonic", 2, 1, run_one<RhsHarmonic>},   {"linear4", 4, 16, run_one<RhsLinear<4>>},
};
static const int kTableSize = sizeof(kTable) / sizeof(kTable[0]);

extern "C" {

int oracle_rhs_lookup(const char* name) {
    for (int i = 0; i < kTableSize; ++i)
        if (std::strcmp(kTable[i].name, name) == 0) return i;
    return -1;
}

int oracle_rhs_info(int id, int* dim, int* n_params) {
    if (id < 0 || id >= kTableSize) return -1;
    *dim = kTable[id].dim;
    *n_params = kTable[id].n_params;
    return 0;
}

int oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

// Same contract as bacon_ivp_solve_ensemble (host buffers).  pow_mode: 0 = libm
// pow (what rk.rs:401 calls), 1 = sqrt(sqrt(x)).  n_threads <= 0: all cores.
int oracle_ivp_solve_ensemble(const bacon_ivp_config* cfg, int rhs_id, size_t n, const double* y0,
                              const double* params, const bacon_ivp_result* out, int pow_mode,
                              int n_threads) {
    if (!cfg || !out || !y0 || !out->y_end || !out->status) return BACON_E_BAD_ARGUMENT;
    if (rhs_id < 0 || rhs_id >= kTableSize) return BACON_E_BAD_ARGUMENT;
    const Entry& e = kTable[rhs_id];
    if (cfg->dim != e.dim || cfg->n_params != e.n_params) return BACON_E_BAD_ARGUMENT;
    if (e.n_params > 0 && !params) return BACON_E_BAD_ARGUMENT;
    if (cfg->method < 0 || cfg->method >= BACON_N_METHODS) return BACON_E_BAD_ARGUMENT;
    RunArgs a{cfg, pow_mode ? bo::PowMode::SqrtSqrt : bo::PowMode::LibmPow, n, y0, params, out};
#ifdef _OPENMP
    const int nt = n_threads > 0 ? n_threads : omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 16) num_threads(nt)
#endif
    for (long long i = 0; i < (long long)n; ++i) e.run(a, (size_t)i);
    (void)n_threads;
    return 0;
}

// ---- roots::secant on the reference's test functions -------------------
// which: 0 = newton_complex (3-dim, tests/roots/mod.rs:41-47), 1 = exp_newton
// (2-dim, :64-70), 2 = cos_secant (1-dim, :81-83).  central: 0 = as written
// (`above + below`, roots/mod.rs:249), 1 = central difference.
int oracle_roots_secant(int which, const double* start, double h, double tol, int n_max, int central,
                        double* solution, unsigned long long* iterations) {
    uint64_t it = 0;
    int rc = -1;
    if (which == 0) {
        auto g = [](const double* x, double* o) {
            o[0] = 3.0 * x[0] - std::cos(x[1] * x[2]) - 0.5;
            o[1] = x[0] * x[0] - 81.0 * ((x[1] + 0.1) * (x[1] + 0.1)) + std::sin(x[2]) + 1.06;
            o[2] = std::exp(-x[0] * x[1]) + 20.0 * x[2] + (M_PI * 10.0 - 3.0) / 3.0;
            return true;
        };
        bo::Vec<3> s{start[0], start[1], start[2]}, r{};
        rc = bo::broyden_secant<3>(g, s, h, tol, n_max, central != 0, r, &it);
        for (int d = 0; d < 3; ++d) solution[d] = r[d];
    } else if (which == 1) {
        auto g = [](const double* x, double* o) {
            o[0] = std::exp(x[0]) - x[0] * x[0];
            o[1] = std::exp(x[1]) - x[1] * x[1] * x[1];
            return true;
        };
        bo::Vec<2> s{start[0], start[1]}, r{};
        rc = bo::broyden_secant<2>(g, s, h, tol, n_max, central != 0, r, &it);
        for (int d = 0; d < 2; ++d) solution[d] = r[d];
    } else if (which == 2) {
        auto g = [](const double* x, double* o) { o[0] = std::cos(x[0]) - x[0]; return true; };
        bo::Vec<1> s{start[0]}, r{};
        rc = bo::broyden_secant<1>(g, s, h, tol, n_max, central != 0, r, &it);
        solution[0] = r[0];
    }
    if (iterations) *iterations = it;
    return rc;
}

// x^(1/n) by libm pow and by the deterministic Newton root the strict Adams kernels share (n = 3 or 5).
void oracle_nth_root(const double* x, size_t n_values, int n, double* by_pow, double* by_det) {
    for (size_t i = 0; i < n_values; ++i) {
        by_pow[i] = std::pow(x[i], 1.0 / (double)n);
        by_det[i] = n == 3 ? bo::det_root<3>(x[i]) : bo::det_root<5>(x[i]);
    }
}

// x^(1/4) both ways, for the pow-vs-sqrt(sqrt) agreement test.
void oracle_fourth_root(const double* x, size_t n, double* by_pow, double* by_sqrt) {
    for (size_t i = 0; i < n; ++i) {
        by_pow[i] = std::pow(x[i], 0.25);
        by_sqrt[i] = std::sqrt(std::sqrt(x[i]));
    }
}

}  // extern "C"
