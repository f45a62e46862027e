// selftest.cpp — the oracle under AddressSanitizer + UndefinedBehaviorSanitizer (the reference is safe
// Rust; its restatement in C++ should at least be clean under the sanitizers).  TEST INFRASTRUCTURE ONLY.
//   g++ -O1 -g -std=c++17 -ffp-contract=off -fsanitize=address,undefined -fno-sanitize-recover=all selftest.cpp -o _selftest
// Runs every stepper family of the C entry points on small ensembles — both semantics, dense output below and above
// the capacity, the optional inputs (first dt, restart record, terminal event) and the path queries — and exits 0.
#include <cstdio>
#include <vector>

#include "oracle_capi.cpp"

struct Buffers {
    std::vector<double> y_end, t_end, dt_end, hist;
    std::vector<int32_t> status;
    std::vector<uint32_t> n_accept, n_reject, n_rhs, hist_len;
    bacon_ivp_result res{};
    Buffers(size_t n, int dim, int cap)
        : y_end(n * dim), t_end(n), dt_end(n), hist(n * (size_t)cap * (1 + dim)), status(n), n_accept(n), n_reject(n), n_rhs(n), hist_len(n) {
        res.y_end = y_end.data(); res.t_end = t_end.data(); res.dt_end = dt_end.data(); res.status = status.data();
        res.n_accept = n_accept.data(); res.n_reject = n_reject.data(); res.n_rhs = n_rhs.data();
        if (cap > 0) { res.hist = hist.data(); res.hist_len = hist_len.data(); }
    }
};

static int failures = 0;
#define CHECK(cond) do { if (!(cond)) { std::printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #cond); ++failures; } } while (0)

static bacon_ivp_config config(int method, int dim, int np, int cap, double dt_min, double dt_max, double tol, double t1, int sem = 0, uint32_t flags = 0) {
    bacon_ivp_config c{};
    c.method = method; c.dim = dim; c.n_params = np; c.semantics = sem; c.flags = flags; c.history_capacity = cap;
    c.dt_min = dt_min; c.dt_max = dt_max; c.tol = tol; c.t_start = 0.0; c.t_end = t1; c.max_attempts = 200000;
    return c;
}

int main() {
    const size_t n = 24;
    // Lorenz, SoA parameters
    std::vector<double> y0(3 * n), p(3 * n);
    for (size_t i = 0; i < n; ++i) {
        y0[i] = -8.0 + i; y0[n + i] = 7.0 - 0.5 * i; y0[2 * n + i] = 20.0 + 0.3 * i;
        p[i] = 10.0; p[n + i] = 28.0; p[2 * n + i] = 8.0 / 3.0;
    }
    const int lor = oracle_rhs_lookup("lorenz");
    for (int method : {BACON_RK45, BACON_RK23, BACON_BDF6, BACON_BDF2, BACON_ADAMS5, BACON_ADAMS3, BACON_EULER})
        for (int sem = 0; sem < 2; ++sem)
            for (int cap : {0, 16, 4096}) {
                const bool euler = method == BACON_EULER;
                bacon_ivp_config c = config(method, 3, 3, cap, euler ? 1e-3 : 1e-7, euler ? 1e-3 : 0.05, 1e-6, 0.15, sem);
                Buffers b(n, 3, cap);
                CHECK(oracle_ivp_solve_ensemble(&c, lor, n, y0.data(), p.data(), &b.res, 1, 2) == 0);
                if (cap > 0) {  // path queries on whatever was stored (overflowed paths included)
                    const double times[3] = {0.0, 0.07, 0.15};
                    std::vector<double> smp(n * 3 * 3), ev(n * 2 * 4);
                    std::vector<uint32_t> cnt(n);
                    const double w[3] = {0.0, 0.0, 1.0};
                    CHECK(oracle_sample_paths(&c, lor, n, y0.data(), p.data(), &b.res, 3, times, smp.data()) == 0);
                    CHECK(oracle_locate_events(&c, lor, n, y0.data(), p.data(), &b.res, w, 24.0, 0, 2, ev.data(), cnt.data()) == 0);
                }
                // optional inputs: first dt, then a second leg from the restart record with a terminal event
                c.dt_init = 0.01;
                bacon_ivp_options o{};
                const double w[3] = {1.0, -1.0, 0.0};
                o.t_start_each = b.t_end.data(); o.dt_start_each = b.dt_end.data();
                o.event_w = w; o.event_c = 0.25; o.event_direction = 0;
                c.t_end = 0.3;
                Buffers b2(n, 3, cap);
                CHECK(oracle_ivp_solve_ensemble_ex(&c, lor, n, b.y_end.data(), p.data(), &o, &b2.res, 1, 2) == 0);
            }
    // Robertson: Broyden (both semantics: LITERAL walks the LU -> full-pivot LU -> QR chain) and Newton
    {
        std::vector<double> r0(3 * n, 0.0), k(3 * n);
        for (size_t i = 0; i < n; ++i) { r0[i] = 1.0; k[i] = 0.04 * (1 + 0.01 * i); k[n + i] = 3e7; k[2 * n + i] = 1e4; }
        const int rob = oracle_rhs_lookup("robertson");
        for (int sem = 0; sem < 2; ++sem)
            for (uint32_t fl : {0u, (uint32_t)BACON_FLAG_BDF_NEWTON}) {
                bacon_ivp_config c = config(BACON_BDF6, 3, 3, 64, 1e-10, 1e-4, 1e-6, 0.003, sem, fl);
                c.max_attempts = 3000;
                Buffers b(n, 3, 64);
                CHECK(oracle_ivp_solve_ensemble(&c, rob, n, r0.data(), k.data(), &b.res, 1, 2) == 0);
            }
    }
    // linear32, AoS parameters, dense output
    {
        const size_t m = 4;
        std::vector<double> z0(32 * m), A(m * 1024, 0.0);
        for (size_t i = 0; i < m; ++i)
            for (int d = 0; d < 32; ++d) {
                z0[d * m + i] = 0.1 * (d + 1) - 0.05 * i;
                A[i * 1024 + d * 32 + d] = -0.5;
                A[i * 1024 + d * 32 + (d + 1) % 32] = 0.3;
                A[i * 1024 + ((d + 1) % 32) * 32 + d] = -0.3;
            }
        bacon_ivp_config c = config(BACON_RK45, 32, 1024, 128, 1e-8, 0.1, 1e-8, 1.0, 0, BACON_FLAG_PARAMS_AOS);
        Buffers b(m, 32, 128);
        const int lin = oracle_rhs_lookup("linear32");
        CHECK(oracle_ivp_solve_ensemble(&c, lin, m, z0.data(), A.data(), &b.res, 1, 1) == 0);
        for (size_t i = 0; i < m; ++i) CHECK(b.status[i] == 0);
        double w[32] = {1.0};
        std::vector<double> ev(m * 4 * 33);
        std::vector<uint32_t> cnt(m);
        CHECK(oracle_locate_events(&c, lin, m, z0.data(), A.data(), &b.res, w, 0.0, 0, 4, ev.data(), cnt.data()) == 0);
    }
    // roots::secant on the reference's test functions
    {
        double start[3] = {0.5, 0.5, 0.5}, sol[3];
        unsigned long long it = 0;
        for (int which = 0; which < 3; ++which) oracle_roots_secant(which, start, 0.1, 1e-8, 1000, 1, sol, &it);
    }
    std::printf(failures ? "selftest: %d failure(s)\n" : "selftest ok\n", failures);
    return failures ? 1 : 0;
}
