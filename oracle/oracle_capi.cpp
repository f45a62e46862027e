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
    const bacon_ivp_options* opts;  // optional inputs (restart record, terminal event); may be NULL
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
    // optional inputs: cfg.dt_init / per-trajectory start time and first dt / terminal event
    double t_start = c.t_start;
    bo::DriveOpts dopt;
    if (c.dt_init > 0.0) { dopt.set_dt = true; dopt.dt = c.dt_init; }
    if (a.opts) {
        if (a.opts->t_start_each) t_start = a.opts->t_start_each[i];
        if (a.opts->dt_start_each) { dopt.set_dt = true; dopt.dt = a.opts->dt_start_each[i]; }
        dopt.ev_w = a.opts->event_w;
        dopt.ev_c = a.opts->event_c;
        dopt.ev_dir = a.opts->event_direction;
    }
    const bo::DriveOpts* o = (dopt.set_dt || dopt.ev_w) ? &dopt : nullptr;
    bo::Solution<D> s;
    switch (c.method) {
        case BACON_RK45:
            s = bo::solve_rk<D, 6>(bo::tableau_rkf45(mode), Rhs{}, p.data(), y0, t_start, c.t_end,
                                   c.dt_min, c.dt_max, c.tol, a.pm, c.max_attempts, keep, o);
            break;
        case BACON_RK23:
            s = bo::solve_rk<D, 4>(bo::tableau_bs23(mode), Rhs{}, p.data(), y0, t_start, c.t_end,
                                   c.dt_min, c.dt_max, c.tol, a.pm, c.max_attempts, keep, o);
            break;
        case BACON_BDF6:
            s = bo::solve_bdf<D, 7>(bo::coefficients_bdf6(), Rhs{}, p.data(), y0, t_start, c.t_end,
                                    c.dt_min, c.dt_max, c.tol, mode, c.max_attempts, keep, newton, o);
            break;
        case BACON_BDF2:
            s = bo::solve_bdf<D, 3>(bo::coefficients_bdf2(), Rhs{}, p.data(), y0, t_start, c.t_end,
                                    c.dt_min, c.dt_max, c.tol, mode, c.max_attempts, keep, newton, o);
            break;
        case BACON_ADAMS5:
            s = bo::solve_adams<D, 5>(bo::coefficients_adams5(), Rhs{}, p.data(), y0, t_start, c.t_end,
                                      c.dt_min, c.dt_max, c.tol, a.pm, mode, c.max_attempts, keep, o);
            break;
        case BACON_ADAMS3:
            s = bo::solve_adams<D, 3>(bo::coefficients_adams3(), Rhs{}, p.data(), y0, t_start, c.t_end,
                                      c.dt_min, c.dt_max, c.tol, a.pm, mode, c.max_attempts, keep, o);
            break;
        default:  // BACON_EULER: config.dt_max carries the builder's dt
            s = bo::solve_euler<D>(Rhs{}, p.data(), y0, t_start, c.t_end, c.dt_max, c.max_attempts, keep, o);
            break;
    }
    store<D>(a, i, s);
}


// ---- queries on stored paths (not in the reference, whose Path is the accepted points only,
// src/ivp.rs:203-211).  CPU statement of bacon_b200/csrc/path_query.cuh, same operation order (this file is built
// -ffp-contract=off, the strict device build -fmad=false): cubic Hermite interpolant between neighbouring knots
// with the right-hand side's slopes; knot 0 = the initial condition, knots 1..m = the records, the exit state
// closes the path when it lies beyond the last record. ------------------------------------------------------
struct PathArgs {
    const bacon_ivp_config* cfg;
    size_t n;
    const double* y0;
    const double* params;
    const bacon_ivp_result* solved;
    size_t n_times;
    const double* times;
    double* samples;
    const double* w;
    double c;
    int direction, capacity;
    double* events;
    uint32_t* n_events;
};

template <int D> struct Knots {
    const PathArgs& a;
    size_t i;
    const double* rec;
    uint32_t m;
    bool closing;
    double t0, tc;
    Knots(const PathArgs& a_, size_t i_) : a(a_), i(i_) {
        const uint32_t cap = (uint32_t)a.cfg->history_capacity;
        rec = a.solved->hist + i * (size_t)cap * (1 + D);
        const uint32_t len = a.solved->hist_len[i];
        m = len < cap ? len : cap;
        t0 = a.solved->t_start ? a.solved->t_start[i] : a.cfg->t_start;
        closing = false;
        tc = 0.0;
        // a path cut short by its capacity ends at its last record (no closing knot across the unrecorded span)
        const bool cut = a.solved->n_accept ? a.solved->n_accept[i] > cap
                                            : (a.solved->status ? a.solved->status[i] == BACON_E_HISTORY_OVERFLOW : false);
        if (a.solved->t_end && a.solved->y_end && !cut) {
            tc = a.solved->t_end[i];
            closing = tc > (m > 0 ? rec[(size_t)(m - 1) * (1 + D)] : t0);
        }
    }
    uint32_t last() const { return m + (closing ? 1u : 0u); }
    double time(uint32_t k) const { return k == 0 ? t0 : (k <= m ? rec[(size_t)(k - 1) * (1 + D)] : tc); }
    void state(uint32_t k, double* y) const {
        for (int d = 0; d < D; ++d)
            y[d] = k == 0 ? a.y0[(size_t)d * a.n + i]
                          : (k <= m ? rec[(size_t)(k - 1) * (1 + D) + 1 + d] : a.solved->y_end[(size_t)d * a.n + i]);
    }
};

using bo::hermite_eval;
using bo::hermite_root;

template <class Rhs> static void load_params(const PathArgs& a, size_t i, std::vector<double>& p) {
    constexpr int P = Rhs::NPARAM;
    p.assign(P > 0 ? P : 1, 0.0);
    const bool shared = (a.cfg->flags & BACON_FLAG_SHARED_PARAMS) != 0;
    const bool aos = (a.cfg->flags & BACON_FLAG_PARAMS_AOS) != 0;
    for (int k = 0; k < P; ++k)
        p[k] = shared ? a.params[k] : (aos ? a.params[i * (size_t)P + k] : a.params[(size_t)k * a.n + i]);
}

template <class Rhs> static void sample_one(const PathArgs& a, size_t i) {
    constexpr int D = Rhs::DIM;
    const Knots<D> kn(a, i);
    const uint32_t K = kn.last();
    std::vector<double> p;
    load_params<Rhs>(a, i, p);
    for (size_t j = 0; j < a.n_times; ++j) {
        const double tau = a.times[j];
        double* out = a.samples + (i * a.n_times + j) * D;
        if (K == 0 || !(tau >= kn.t0 && tau <= kn.time(K))) {
            if (tau == kn.t0) kn.state(0, out);
            else for (int d = 0; d < D; ++d) out[d] = std::nan("");
            continue;
        }
        uint32_t lo = 1, hi = K;
        while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            if (kn.time(mid) >= tau) hi = mid;
            else lo = mid + 1;
        }
        const double ta = kn.time(lo - 1), tb = kn.time(lo);
        double ya[D], yb[D], fa[D], fb[D];
        kn.state(lo - 1, ya);
        kn.state(lo, yb);
        Rhs{}(ta, ya, p.data(), fa);
        Rhs{}(tb, yb, p.data(), fb);
        const double h = tb - ta;
        const double th = h > 0.0 ? (tau - ta) / h : 0.0;
        hermite_eval<D>(th, h, ya, yb, fa, fb, out);
    }
}

template <int D> static double event_fn(const PathArgs& a, const double* y) {
    double s = a.w[0] * y[0];
    for (int d = 1; d < D; ++d) s += a.w[d] * y[d];
    return s - a.c;
}

template <class Rhs> static void events_one(const PathArgs& a, size_t i) {
    constexpr int D = Rhs::DIM;
    const Knots<D> kn(a, i);
    const uint32_t K = kn.last();
    std::vector<double> p;
    load_params<Rhs>(a, i, p);
    double ya[D], yb[D];
    kn.state(0, ya);
    double ga = event_fn<D>(a, ya);
    uint32_t count = 0;
    for (uint32_t k = 1; k <= K; ++k) {
        kn.state(k, yb);
        const double gb = event_fn<D>(a, yb);
        const bool rising = ga < 0.0 && gb >= 0.0, falling = ga > 0.0 && gb <= 0.0;
        const bool hit = a.direction > 0 ? rising : (a.direction < 0 ? falling : (rising || falling));
        if (hit) {
            if (count < (uint32_t)a.capacity) {
                const double ta = kn.time(k - 1), tb = kn.time(k);
                double fa[D], fb[D];
                Rhs{}(ta, ya, p.data(), fa);
                Rhs{}(tb, yb, p.data(), fb);
                const double h = tb - ta;
                double da = a.w[0] * fa[0], db = a.w[0] * fb[0];
                for (int d = 1; d < D; ++d) {
                    da += a.w[d] * fa[d];
                    db += a.w[d] * fb[d];
                }
                const double th = hermite_root(ga, gb, h * da, h * db);
                double* dst = a.events + (i * (size_t)a.capacity + count) * (1 + D);
                dst[0] = ta + th * h;
                hermite_eval<D>(th, h, ya, yb, fa, fb, dst + 1);
            }
            ++count;
        }
        for (int d = 0; d < D; ++d) ya[d] = yb[d];
        ga = gb;
    }
    a.n_events[i] = count;
}

typedef void (*run_fn)(const RunArgs&, size_t);
typedef void (*path_fn)(const PathArgs&, size_t);
struct Entry { const char* name; int dim; int n_params; run_fn run; path_fn sample; path_fn events; };
#define ORACLE_ENTRY(name, dim, np, R) {name, dim, np, run_one<R>, sample_one<R>, events_one<R>}
static const Entry kTable[] = {
    ORACLE_ENTRY("lorenz", 3, 3, RhsLorenz),       ORACLE_ENTRY("vdp", 2, 1, RhsVdp),
    ORACLE_ENTRY("robertson", 3, 3, RhsRobertson), ORACLE_ENTRY("linear32", 32, 1024, RhsLinear<32>),
    ORACLE_ENTRY("exp", 1, 0, RhsExp),             ORACLE_ENTRY("decay", 1, 0, RhsDecay),
    ORACLE_ENTRY("quadratic", 1, 0, RhsQuadratic), ORACLE_ENTRY("cos", 1, 0, RhsCos),
    ORACLE_ENTRY("harmonic", 2, 1, RhsHarmonic),   ORACLE_ENTRY("linear4", 4, 16, RhsLinear<4>),
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
int oracle_ivp_solve_ensemble_ex(const bacon_ivp_config* cfg, int rhs_id, size_t n, const double* y0,
                                 const double* params, const bacon_ivp_options* opts, const bacon_ivp_result* out,
                                 int pow_mode, int n_threads) {
    if (!cfg || !out || !y0 || !out->y_end || !out->status) return BACON_E_BAD_ARGUMENT;
    if (rhs_id < 0 || rhs_id >= kTableSize) return BACON_E_BAD_ARGUMENT;
    const Entry& e = kTable[rhs_id];
    if (cfg->dim != e.dim || cfg->n_params != e.n_params) return BACON_E_BAD_ARGUMENT;
    if (e.n_params > 0 && !params) return BACON_E_BAD_ARGUMENT;
    if (cfg->method < 0 || cfg->method >= BACON_N_METHODS) return BACON_E_BAD_ARGUMENT;
    RunArgs a{cfg, pow_mode ? bo::PowMode::SqrtSqrt : bo::PowMode::LibmPow, n, y0, params, out, opts};
#ifdef _OPENMP
    const int nt = n_threads > 0 ? n_threads : omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 16) num_threads(nt)
#endif
    for (long long i = 0; i < (long long)n; ++i) e.run(a, (size_t)i);
    (void)n_threads;
    return 0;
}

int oracle_ivp_solve_ensemble(const bacon_ivp_config* cfg, int rhs_id, size_t n, const double* y0,
                              const double* params, const bacon_ivp_result* out, int pow_mode,
                              int n_threads) {
    return oracle_ivp_solve_ensemble_ex(cfg, rhs_id, n, y0, params, nullptr, out, pow_mode, n_threads);
}

// Same contracts as bacon_ivp_sample_paths / bacon_ivp_locate_events (host buffers).
static int path_check(const bacon_ivp_config* cfg, int rhs_id, const double* y0, const double* params,
                      const bacon_ivp_result* solved) {
    if (!cfg || !solved || !y0 || !solved->hist || !solved->hist_len || cfg->history_capacity <= 0) return BACON_E_BAD_ARGUMENT;
    if (rhs_id < 0 || rhs_id >= kTableSize) return BACON_E_BAD_ARGUMENT;
    const Entry& e = kTable[rhs_id];
    if (cfg->dim != e.dim || cfg->n_params != e.n_params) return BACON_E_BAD_ARGUMENT;
    if (e.n_params > 0 && !params) return BACON_E_BAD_ARGUMENT;
    return 0;
}
int oracle_sample_paths(const bacon_ivp_config* cfg, int rhs_id, size_t n, const double* y0, const double* params,
                        const bacon_ivp_result* solved, size_t n_times, const double* times, double* samples) {
    if (const int rc = path_check(cfg, rhs_id, y0, params, solved)) return rc;
    PathArgs a{cfg, n, y0, params, solved, n_times, times, samples, nullptr, 0.0, 0, 0, nullptr, nullptr};
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 16)
#endif
    for (long long i = 0; i < (long long)n; ++i) kTable[rhs_id].sample(a, (size_t)i);
    return 0;
}
int oracle_locate_events(const bacon_ivp_config* cfg, int rhs_id, size_t n, const double* y0, const double* params,
                         const bacon_ivp_result* solved, const double* w, double c, int direction, int capacity,
                         double* events, uint32_t* n_events) {
    if (const int rc = path_check(cfg, rhs_id, y0, params, solved)) return rc;
    if (!w || !n_events || capacity < 0 || (capacity > 0 && !events)) return BACON_E_BAD_ARGUMENT;
    PathArgs a{cfg, n, y0, params, solved, 0, nullptr, nullptr, w, c, direction, capacity, events, n_events};
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 16)
#endif
    for (long long i = 0; i < (long long)n; ++i) kTable[rhs_id].events(a, (size_t)i);
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

// The coefficient tables of a method as the steppers use them, flattened (tests/test_oracle.py holds them against the
// numbers parsed out of the reference's own source text, tests/golden/reference_coefficients.json):
//   RK45 / RK23 (O stages): c[O], A[O][O] row-major = what the stepper's row_iter() sees, b[O], e[O], safety
//   BDF6 / BDF2           : higher[O], lower[O]
//   Adams5 / Adams3       : predictor[O], corrector[O], error
// Returns how many doubles the method has (and writes them if `cap` is large enough), -1 for an unknown method.
int oracle_coefficients(int method, int literal, double* out, int cap) {
    std::vector<double> v;
    const bo::Mode mode = literal ? bo::Mode::Literal : bo::Mode::Corrected;
    auto rk = [&](const auto& T, int O) {
        for (int i = 0; i < O; ++i) v.push_back(T.c[i]);
        for (int r = 0; r < O; ++r)
            for (int c = 0; c < O; ++c) v.push_back(T.A[r][c]);
        for (int i = 0; i < O; ++i) v.push_back(T.b[i]);
        for (int i = 0; i < O; ++i) v.push_back(T.e[i]);
        v.push_back(T.safety);
    };
    auto bdf = [&](const auto& Cf, int O) {
        for (int i = 0; i < O; ++i) v.push_back(Cf.higher[i]);
        for (int i = 0; i < O; ++i) v.push_back(Cf.lower[i]);
    };
    auto adams = [&](const auto& Cf, int O) {
        for (int i = 0; i < O; ++i) v.push_back(Cf.predictor[i]);
        for (int i = 0; i < O; ++i) v.push_back(Cf.corrector[i]);
        v.push_back(Cf.error);
    };
    switch (method) {
        case BACON_RK45: rk(bo::tableau_rkf45(mode), 6); break;
        case BACON_RK23: rk(bo::tableau_bs23(mode), 4); break;
        case BACON_BDF6: bdf(bo::coefficients_bdf6(), 7); break;
        case BACON_BDF2: bdf(bo::coefficients_bdf2(), 3); break;
        case BACON_ADAMS5: adams(bo::coefficients_adams5(), 5); break;
        case BACON_ADAMS3: adams(bo::coefficients_adams3(), 3); break;
        default: return -1;
    }
    if (out && cap >= (int)v.size())
        for (size_t i = 0; i < v.size(); ++i) out[i] = v[i];
    return (int)v.size();
}

// x^(1/4) both ways, for the pow-vs-sqrt(sqrt) agreement test.
void oracle_fourth_root(const double* x, size_t n, double* by_pow, double* by_sqrt) {
    for (size_t i = 0; i < n; ++i) {
        by_pow[i] = std::pow(x[i], 0.25);
        by_sqrt[i] = std::sqrt(std::sqrt(x[i]));
    }
}

}  // extern "C"
