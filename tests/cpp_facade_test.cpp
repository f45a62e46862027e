// Exercises include/bacon_ivp.hpp.  `validate`: builder rules only (no GPU).  `solve`: README example on the GPU.
#include <cmath>
#include <cstdio>
#include <cstring>

#include "bacon_ivp.hpp"

#define EXPECT_THROW_CODE(expr, want)                                                \
    do {                                                                             \
        int got = 0;                                                                 \
        try { expr; } catch (const bacon::IVPError& e) { got = e.code; }             \
        if (got != (want)) { std::printf("FAIL %s: code %d, want %d\n", #expr, got, (want)); return 1; } \
    } while (0)

int main(int argc, char** argv) {
    using namespace bacon;
    const bool solve = argc > 1 && std::strcmp(argv[1], "solve") == 0;
    EXPECT_THROW_CODE(RK45<>(1).with_tolerance(0.0), BACON_E_TOLERANCE_OOB);
    EXPECT_THROW_CODE(RK45<>(1).with_dt_min(-1.0), BACON_E_TIME_DELTA_OOB);
    EXPECT_THROW_CODE(BDF6<>(1).with_maximum_dt(0.0), BACON_E_TIME_DELTA_OOB);
    EXPECT_THROW_CODE(RK23<>(1).with_end(1.0).with_start(2.0), BACON_E_TIME_START_OOB);
    EXPECT_THROW_CODE(RK23<>(1).with_start(2.0).with_end(1.0), BACON_E_TIME_END_OOB);
    EXPECT_THROW_CODE(RK45<>(1).with_dt_min(0.01).config(), BACON_E_MISSING_PARAMETERS);
    EXPECT_THROW_CODE(RK45<>(1).with_derivative("nope"), BACON_E_BAD_ARGUMENT);
    {
        RK45<> s(1);
        s.with_minimum_dt(0.5).with_maximum_dt(0.1).with_tolerance(1e-3).with_start(0).with_end(1);
        const bacon_ivp_config c = s.config();
        if (c.dt_min != 0.1 || c.dt_max != 0.1) { std::printf("FAIL min/max ordering\n"); return 1; }
    }
    // IVPSolver::new / new_dyn with the reference's Dimension check (src/lib.rs:53-76)
    if (RK45<3>::make().dim() != 3 || BDF6<>::new_dyn(2).dim() != 2) { std::printf("FAIL constructors\n"); return 1; }
    EXPECT_THROW_CODE(RK45<>::make(), BACON_E_STATIC_ON_DYNAMIC);        // RK45::<Dyn>::new()
    EXPECT_THROW_CODE(Euler<2>::new_dyn(2), BACON_E_DYNAMIC_ON_STATIC);  // Euler::<U2>::new_dyn(2)
    EXPECT_THROW_CODE(RK45<1>::make().with_initial_dt(0.0), BACON_E_TIME_DELTA_OOB);
    if (solve) {  // README.md:24-40
        RK45<> s(1);
        s.with_dt_min(0.01).with_dt_max(0.1).with_tolerance(1e-4).with_initial_conditions({1.0}).with_start(0.0).with_end(10.0).build();
        const Path path = s.solve_ivp("exp");
        if (path.size() != 128) { std::printf("FAIL path size %zu\n", path.size()); return 1; }
        for (const auto& pt : path)
            if (std::fabs(pt.second[0] - std::exp(pt.first)) > 2e-2 * std::exp(pt.first)) { std::printf("FAIL accuracy\n"); return 1; }
        if (path.back().first != 10.0) { std::printf("FAIL end time\n"); return 1; }
        std::printf("solve ok: %zu points, y(10) = %.10g\n", path.size(), path.back().second[0]);
        // path queries: y = exp(t) sampled between the accepted points, and the crossing of y = 100 (t = ln 100)
        s.with_history(256);
        const double y0[1] = {1.0};
        const EnsembleResult r = s.solve_ivp_ensemble(1, y0, nullptr);
        const std::vector<double> ys = r.sample({0.0, 1.234, 5.0, 10.0, 11.0});
        if (ys[0] != 1.0 || std::fabs(ys[1] / std::exp(1.234) - 1.0) > 1e-5 || std::fabs(ys[2] / std::exp(5.0) - 1.0) > 1e-5 ||
            ys[3] != r.y(0, 0) || ys[4] == ys[4]) { std::printf("FAIL sample %g %g %g %g %g\n", ys[0], ys[1], ys[2], ys[3], ys[4]); return 1; }
        const auto ev = r.locate_events({1.0}, 100.0, +1, 2);
        if (ev.second[0] != 1 || std::fabs(ev.first[0] - std::log(100.0)) > 1e-5 || std::fabs(ev.first[1] - 100.0) > 1e-9) {
            std::printf("FAIL event %u %g %g\n", ev.second[0], ev.first[0], ev.first[1]); return 1; }
        std::printf("queries ok: y(1.234) = %.8g, y = 100 at t = %.8g\n", ys[1], ev.first[0]);
        // the same crossing as a terminal event, then the restart record carries the solve on to t = 10
        s.with_terminal_event({1.0}, 100.0, +1);
        const EnsembleResult stop = s.solve_ivp_ensemble(1, y0, nullptr);
        if (stop.status[0] != BACON_STOPPED_AT_EVENT || std::fabs(stop.t_end[0] - std::log(100.0)) > 1e-5 ||
            std::fabs(stop.y(0, 0) - 100.0) > 1e-9) { std::printf("FAIL terminal event %d %g\n", stop.status[0], stop.t_end[0]); return 1; }
        s.with_terminal_event({});
        Restart rs;
        rs.t_start_each = stop.t_end.data();
        rs.dt_start_each = stop.dt_end.data();
        const EnsembleResult rest = s.solve_ivp_ensemble(1, stop.y_end.data(), nullptr, false, 1, rs);
        if (rest.status[0] != BACON_OK || rest.t_end[0] != 10.0 || std::fabs(rest.y(0, 0) / std::exp(10.0) - 1.0) > 1e-3) {
            std::printf("FAIL restart %d %g %g\n", rest.status[0], rest.t_end[0], rest.y(0, 0)); return 1; }
        std::printf("terminal event + restart ok: stopped at t = %.8g, resumed to y(10) = %.8g\n", stop.t_end[0], rest.y(0, 0));
    }
    std::printf("ok\n");
    return 0;
}
