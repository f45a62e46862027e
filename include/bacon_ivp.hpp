// bacon_ivp.hpp — C++ host-side mirror of bacon_sci::ivp's solver front end over the C ABI
// (header only; link with -lbacon_ivp).  The reference is compiled code (Rust) and no Rust
// toolchain exists in the build image, so this is the compiled-language façade that is
// actually built and tested; rust/ holds the same façade written in Rust (unverified).
//
// Names and error behaviour follow the `IVPSolver` builder trait (src/ivp.rs:134-190) as
// implemented in src/ivp/rk.rs:118-343 and src/ivp/bdf.rs:124-332.  Rust's
// `Result<Self, IVPError>` becomes "returns *this or throws bacon::IVPError"; the error
// carries the same variant (src/ivp.rs:50-76).  README aliases (README.md:24-40) included.
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "bacon_ivp.h"

namespace bacon {

struct IVPError : std::runtime_error {
    int code;  // bacon_status; 1..12 are the IVPError variants of src/ivp.rs:50-76
    IVPError(int c, const std::string& what) : std::runtime_error(std::string(bacon_status_name(c)) + ": " + what), code(c) {}
};

inline void check(int rc) {
    if (rc != 0) throw IVPError(rc, bacon_last_error());
}

// One trajectory's `Path` (src/ivp.rs:203): accepted (t, y) points.
using Path = std::vector<std::pair<double, std::vector<double>>>;

struct EnsembleResult {
    size_t n = 0;
    int dim = 0, capacity = 0;
    std::vector<double> y_end, t_end, dt_end, hist;  // layouts of bacon_ivp_result; hist = [n][capacity][1 + dim]
    std::vector<int32_t> status;
    std::vector<uint32_t> n_accept, n_reject, n_rhs, hist_len;
    bacon_ivp_launch_info launch{};
    // what the path queries below need of the solve (kept by solve_ivp_ensemble)
    bacon_ivp_config cfg{};
    int rhs = -1;
    std::vector<double> y0, params, t_start;
    double y(size_t i, int d) const { return y_end[(size_t)d * n + i]; }
    Path path(size_t i) const {
        Path p;
        for (uint32_t k = 0; k < hist_len[i]; ++k) {
            const double* rec = &hist[(i * capacity + k) * (size_t)(1 + dim)];  // (t, y[0..dim))
            p.emplace_back(rec[0], std::vector<double>(rec + 1, rec + 1 + dim));
        }
        return p;
    }

    // ---- queries on the stored paths (bacon_ivp_sample_paths / bacon_ivp_locate_events; not in the reference, whose
    // Path is the accepted points only, src/ivp.rs:203-211).  Need a solve with_history(capacity).
    bacon_ivp_result solved() const {
        bacon_ivp_result o{};
        o.hist = const_cast<double*>(hist.data());
        o.hist_len = const_cast<uint32_t*>(hist_len.data());
        o.t_end = const_cast<double*>(t_end.data());
        o.y_end = const_cast<double*>(y_end.data());
        o.n_accept = const_cast<uint32_t*>(n_accept.data());  // a path cut short by its capacity has no closing knot
        o.status = const_cast<int32_t*>(status.data());
        if (!t_start.empty()) o.t_start = t_start.data();     // a resumed leg: per-trajectory start times
        return o;
    }
    // the state of every trajectory at `times`: [n][times.size()][dim], NaN outside a trajectory's path
    std::vector<double> sample(const std::vector<double>& times) const {
        std::vector<double> out(n * times.size() * (size_t)dim);
        const bacon_ivp_result o = solved();
        const int rc = bacon_ivp_sample_paths(&cfg, rhs, n, y0.data(), params.empty() ? nullptr : params.data(), &o,
                                              times.size(), times.data(), out.data());
        if (rc != 0) throw IVPError(rc, bacon_last_error());
        return out;
    }
    // zeros of w . y - c along every path: events [n][capacity][1 + dim] = (t*, y(t*)), counts [n]
    std::pair<std::vector<double>, std::vector<uint32_t>> locate_events(const std::vector<double>& w, double c,
                                                                        int direction, int capacity) const {
        if ((int)w.size() != dim) throw IVPError(BACON_E_BAD_ARGUMENT, "w must have dim entries");
        std::vector<double> ev(n * (size_t)capacity * (size_t)(1 + dim));
        std::vector<uint32_t> cnt(n);
        const bacon_ivp_result o = solved();
        const int rc = bacon_ivp_locate_events(&cfg, rhs, n, y0.data(), params.empty() ? nullptr : params.data(), &o,
                                               w.data(), c, direction, capacity, ev.data(), cnt.data());
        if (rc != 0) throw IVPError(rc, bacon_last_error());
        return {std::move(ev), std::move(cnt)};
    }
};

// The restart record of a solve (t_end, dt_end per trajectory; y_end is the next leg's y0): the C-ABI form of the
// reference's in-memory resumable iterator (src/ivp.rs:220-238).  Either pointer may be null.
struct Restart {
    const double* t_start_each = nullptr;
    const double* dt_start_each = nullptr;
};

constexpr int Dyn = BACON_DIM_DYN;  // the `Dyn` type parameter (src/lib.rs:68)

// D is the reference's type parameter: a static dimension C >= 1 (`Const<C>`, e.g. RungeKutta45<3>) or Dyn.
template <int METHOD, int D = Dyn> class Solver {
    bacon_solver* h_ = nullptr;
    int dim_ = 0;
    int rhs_ = -1;
    std::vector<double> y0_;
    std::vector<double> ev_w_;
    double ev_c_ = 0.0;
    int ev_dir_ = 0;

    Solver() = default;

  public:
    // (round-1 form: a run-time dimension, no Dimension check; the same as `Solver<M, Dyn>::new_dyn(dim)`)
    explicit Solver(int dim) : h_(bacon_solver_new(METHOD, dim)), dim_(dim) {
        if (!h_) throw IVPError(BACON_E_BAD_ARGUMENT, bacon_last_error());
    }
    ~Solver() { bacon_solver_free(h_); }
    Solver(const Solver&) = delete;
    Solver& operator=(const Solver&) = delete;
    Solver(Solver&& o) noexcept
        : h_(o.h_), dim_(o.dim_), rhs_(o.rhs_), y0_(std::move(o.y0_)), ev_w_(std::move(o.ev_w_)), ev_c_(o.ev_c_), ev_dir_(o.ev_dir_) {
        o.h_ = nullptr;
    }

    // IVPSolver::new (ivp.rs:159): Const<C> -> a solver of dimension C; Dyn -> throws StaticOnDynamic (lib.rs:69-71)
    static Solver make() {
        Solver s;
        check(bacon_solver_new_static(METHOD, D, &s.h_));
        s.dim_ = D;
        return s;
    }
    // IVPSolver::new_dyn (ivp.rs:163): Dyn -> a solver of dimension `size`; Const<C> -> throws DynamicOnStatic (lib.rs:63-65)
    static Solver new_dyn(int size) {
        Solver s;
        check(bacon_solver_new_dyn(METHOD, D, size, &s.h_));
        s.dim_ = size;
        return s;
    }
    int dim() const { return dim_; }

    Solver& with_tolerance(double tol) { check(bacon_solver_with_tolerance(h_, tol)); return *this; }          // rk.rs:168
    Solver& with_maximum_dt(double v) { check(bacon_solver_with_maximum_dt(h_, v)); return *this; }            // rk.rs:179
    Solver& with_minimum_dt(double v) { check(bacon_solver_with_minimum_dt(h_, v)); return *this; }            // rk.rs:197
    Solver& with_initial_time(double v) { check(bacon_solver_with_initial_time(h_, v)); return *this; }        // rk.rs:212
    Solver& with_ending_time(double v) { check(bacon_solver_with_ending_time(h_, v)); return *this; }          // rk.rs:224
    Solver& with_initial_conditions_slice(const std::vector<double>& y0) {                                      // ivp.rs:177
        if ((int)y0.size() != dim_) throw IVPError(BACON_E_BAD_ARGUMENT, "initial conditions do not match dim()");
        y0_ = y0;
        return *this;
    }
    Solver& with_initial_conditions(const std::vector<double>& y0) { return with_initial_conditions_slice(y0); }
    Solver& with_derivative(const std::string& rhs_name) {                                                      // ivp.rs:186
        rhs_ = bacon_rhs_lookup(rhs_name.c_str());
        if (rhs_ < 0) throw IVPError(BACON_E_BAD_ARGUMENT, "no right-hand side named '" + rhs_name + "'");
        return *this;
    }
    // the closest thing to `with_derivative(closure)`: the functor as CUDA C++ source text, compiled by the library
    // (NVRTC) and inlined into the kernels; throws IVPError(UserError) with the compiler log if it does not compile
    Solver& with_derivative_source(const std::string& name, const std::string& type_name, const std::string& source,
                                   int n_params) {
        const int id = bacon_rhs_register_source(name.c_str(), type_name.c_str(), source.c_str(), dim_, n_params);
        if (id < 0) check(-id);
        rhs_ = id;
        return *this;
    }
    // README.md:33-39
    Solver& with_dt_max(double v) { return with_maximum_dt(v); }
    Solver& with_dt_min(double v) { return with_minimum_dt(v); }
    Solver& with_start(double v) { return with_initial_time(v); }
    Solver& with_end(double v) { return with_ending_time(v); }
    Solver& build() { return *this; }
    // engine knobs
    Solver& with_semantics(int s) { check(bacon_solver_with_semantics(h_, s)); return *this; }
    Solver& with_flags(uint32_t f) { check(bacon_solver_with_flags(h_, f)); return *this; }
    Solver& with_history(int cap) { check(bacon_solver_with_history(h_, cap)); return *this; }
    Solver& with_max_attempts(uint64_t cap) { check(bacon_solver_with_max_attempts(h_, cap)); return *this; }
    // first step size instead of (dt_max + dt_min)/2 (rk.rs:315); clamped into [dt_min, dt_max] at solve time
    Solver& with_initial_dt(double dt) { check(bacon_solver_with_initial_dt(h_, dt)); return *this; }
    // stop every trajectory at the first zero of w . y - c (direction +1 rising, -1 falling, 0 both): status
    // BACON_STOPPED_AT_EVENT, t_end / y_end = the event point.  An empty w removes it.  Not in the reference.
    Solver& with_terminal_event(const std::vector<double>& w, double c = 0.0, int direction = 0) {
        if (!w.empty() && (int)w.size() != dim_) throw IVPError(BACON_E_BAD_ARGUMENT, "w must have dim() entries");
        ev_w_ = w;
        ev_c_ = c;
        ev_dir_ = direction;
        return *this;
    }

    bacon_ivp_config config() const {
        bacon_ivp_config c;
        check(bacon_solver_config(h_, &c));
        return c;
    }

    // N initial conditions x N parameter sets.  y0: [dim][n]; params: [n_params][n] (or [n_params] shared).
    EnsembleResult solve_ivp_ensemble(size_t n, const double* y0, const double* params, bool shared_params = false,
                                      int n_gpus = 1, Restart restart = {}) const {
        if (rhs_ < 0) throw IVPError(BACON_E_MISSING_PARAMETERS, "with_derivative was not called");
        bacon_ivp_config c = config();
        int d = 0, np = 0;
        check(bacon_rhs_info(rhs_, nullptr, &d, &np));
        c.n_params = np;
        if (shared_params) c.flags |= BACON_FLAG_SHARED_PARAMS;
        EnsembleResult r;
        r.n = n;
        r.dim = c.dim;
        r.capacity = c.history_capacity;
        r.y_end.resize((size_t)c.dim * n);
        r.t_end.resize(n);
        r.dt_end.resize(n);
        r.status.assign(n, -1);
        r.n_accept.resize(n);
        r.n_reject.resize(n);
        r.n_rhs.resize(n);
        bacon_ivp_result o{};
        o.y_end = r.y_end.data();
        o.t_end = r.t_end.data();
        o.dt_end = r.dt_end.data();
        o.status = r.status.data();
        o.n_accept = r.n_accept.data();
        o.n_reject = r.n_reject.data();
        o.n_rhs = r.n_rhs.data();
        if (c.history_capacity > 0) {
            r.hist.resize(n * c.history_capacity * (size_t)(1 + c.dim));
            r.hist_len.resize(n);
            o.hist = r.hist.data();
            o.hist_len = r.hist_len.data();
        }
        bacon_ivp_options opt{};
        opt.t_start_each = restart.t_start_each;
        opt.dt_start_each = restart.dt_start_each;
        if (!ev_w_.empty()) {
            opt.event_w = ev_w_.data();
            opt.event_c = ev_c_;
            opt.event_direction = ev_dir_;
        }
        check(bacon_ivp_solve_ensemble_ex(&c, rhs_, n, y0, params, &opt, &o, n_gpus));
        check(bacon_ivp_last_launch(&r.launch));
        if (restart.t_start_each) r.t_start.assign(restart.t_start_each, restart.t_start_each + n);
        if (c.history_capacity > 0) {  // for the path queries
            r.cfg = c;
            r.rhs = rhs_;
            r.y0.assign(y0, y0 + (size_t)c.dim * n);
            if (np > 0 && params) r.params.assign(params, params + (size_t)np * (shared_params ? 1 : n));
        }
        return r;
    }

    // solve(data) + collect_vec (rk.rs:249-343, ivp.rs:209-211): the single trajectory set by with_initial_conditions
    Path solve(const std::vector<double>& data = {}, int capacity = 1 << 16) {
        if (y0_.empty()) throw IVPError(BACON_E_MISSING_PARAMETERS, "with_initial_conditions was not called");
        // collect_vec grows its Vec (ivp.rs:209-211): a path longer than `capacity` is integrated once more with the
        // capacity the first pass reported
        EnsembleResult r;
        for (int pass = 0; pass < 2; ++pass) {
            with_history(capacity);
            r = solve_ivp_ensemble(1, y0_.data(), data.empty() ? nullptr : data.data());
            with_history(0);
            if ((int64_t)r.n_accept[0] <= (int64_t)capacity) break;
            capacity = (int)r.n_accept[0];
        }
        if (r.status[0] != BACON_OK && r.status[0] != BACON_STOPPED_AT_EVENT)
            throw IVPError(r.status[0], "trajectory failed after " + std::to_string(r.hist_len[0]) + " point(s)");
        return r.path(0);
    }
    Path solve_ivp(const std::string& rhs_name, const std::vector<double>& data = {}) {  // README.md:40
        return with_derivative(rhs_name).solve(data);
    }
};

// `RungeKutta45<3>` is the reference's `RungeKutta45<'a, f64, U3, T>`; `RungeKutta45<>` its `Dyn` form
template <int D = Dyn> using RungeKutta45 = Solver<BACON_RK45, D>;  // rk.rs:561
template <int D = Dyn> using RungeKutta23 = Solver<BACON_RK23, D>;  // rk.rs:656
template <int D = Dyn> using BDF6 = Solver<BACON_BDF6, D>;          // bdf.rs:706
template <int D = Dyn> using BDF2 = Solver<BACON_BDF2, D>;          // bdf.rs:762
template <int D = Dyn> using Adams5 = Solver<BACON_ADAMS5, D>;      // adams.rs:633
template <int D = Dyn> using Adams3 = Solver<BACON_ADAMS3, D>;      // adams.rs:693
template <int D = Dyn> using Euler = Solver<BACON_EULER, D>;        // ivp.rs:269 (with_tolerance is a no-op, dt = average of the bounds given)
template <int D = Dyn> using RK45 = RungeKutta45<D>;                // README.md:24
template <int D = Dyn> using RK23 = RungeKutta23<D>;

}  // namespace bacon
