// rk_strict.cuh — the reference's RungeKuttaSolver::step (src/ivp/rk.rs:361-423) in
// the reference's OWN operation order, for bit-level parity with the CPU oracle
// (oracle/bacon_oracle.hpp, built -ffp-contract=off; pow_mode = sqrt(sqrt)):
//   * k_j = f_j * dt stored ("half_steps", rk.rs:383), never re-zeroed between attempts
//   * dense O x O stage sums including structural zeros (rk.rs:372-374)
//   * error = sqrt(sum s^2) / dt (rk.rs:390), ratio = tol/error, (ratio)^(1/4) as sqrt(sqrt())
//   * every product/sum individually rounded (__dmul_rn/__dadd_rn: never contracted to FMA);
//     the translation unit is additionally built with -fmad=false so the RHS functor is not
//     contracted either.
// The Butcher tableau is a RUNTIME table in __constant__ memory, filled for either
// semantics: REF_CORRECTED, or REF_LITERAL (column-major from_vec, rk.rs:459 vs row_iter
// rk.rs:370; 1859/4014 rk.rs:499; safety 100/100 rk.rs:267) — so this kernel also is the
// device form of the source exactly as written.
#pragma once
#include "ivp_common.cuh"
#include "tableaux.cuh"

namespace bacon {

// One constant buffer per (tableau, semantics): its content never changes, so strict or LITERAL solves that run
// concurrently on different streams (the device entry point is asynchronous) cannot overwrite each other's tableau
// under a running kernel; every launch re-writes its slot with the same bytes.
static __constant__ RkTableauRt c_rk_tabs[4];

template <class Rhs, int O_> struct RkStrictStepper {
    using RhsT = Rhs;
    static constexpr int D = Rhs::DIM;
    static constexpr int P = Rhs::NPARAM;
    static constexpr int O = O_;

    double t_start, t_end, dt_min, dt_max, tol, dt0;
    uint32_t cap;
    double y[D], p[P > 0 ? P : 1];
    double hs[O][D];  // half_steps columns (rk.rs:323-327)
    double t, dt;
    uint32_t n_acc, n_rej, n_att;
    int tab_slot;

    __device__ __forceinline__ explicit RkStrictStepper(const bacon_launch_args& a) {
        tab_slot = rk_tab_slot(O, a.cfg.semantics);
        t_start = a.cfg.t_start;
        t_end = a.cfg.t_end;
        dt_min = a.cfg.dt_min;
        dt_max = a.cfg.dt_max;
        tol = a.cfg.tol;
        dt0 = __dmul_rn(__dadd_rn(dt_max, dt_min), 0.5);  // rk.rs:315
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
#pragma unroll
        for (int i = 0; i < O; ++i)
#pragma unroll
            for (int d = 0; d < D; ++d) hs[i][d] = 0.0;
        if (live) load_problem<D, P>(a, idx, y, p);
    }
    __device__ __forceinline__ void apply_restart(const bacon_launch_args& a, unsigned long long idx) {
        trajectory_start(a, idx, dt_min, dt_max, t, dt);
    }
    __device__ __forceinline__ uint32_t n_rhs() const { return n_att * (uint32_t)O; }
    __device__ __forceinline__ double out_t() const { return t; }
    __device__ __forceinline__ const double (&out_y() const)[D] { return y; }
    __device__ __forceinline__ const double (&end_y() const)[D] { return y; }

    __device__ __forceinline__ int attempt(bool& yielded) {
        const Rhs rhs{};
        const RkTableauRt& c_rk_tab = c_rk_tabs[tab_slot];
        yielded = false;
        if (n_att >= cap) return BACON_E_MAX_ATTEMPTS;
        if (t >= t_end) return BACON_OK;                            // rk.rs:362-364
        if (__dadd_rn(t, dt) >= t_end) dt = __dadd_rn(t_end, -t);   // rk.rs:366-368

#pragma unroll
        for (int i = 0; i < O; ++i) {  // rk.rs:370-384
            double sp[D];
#pragma unroll
            for (int d = 0; d < D; ++d) sp[d] = y[d];
#pragma unroll
            for (int j = 0; j < O; ++j)
#pragma unroll
                for (int d = 0; d < D; ++d) sp[d] = __dadd_rn(sp[d], __dmul_rn(hs[j][d], c_rk_tab.a[i][j]));
            const double step_time = __dadd_rn(t, __dmul_rn(c_rk_tab.c[i], dt));
            double dy[D];
            rhs(step_time, sp, p, dy);
#pragma unroll
            for (int d = 0; d < D; ++d) hs[i][d] = __dmul_rn(dy[d], dt);
        }
        n_att++;

        double sp[D];
#pragma unroll
        for (int d = 0; d < D; ++d) sp[d] = __dmul_rn(hs[0][d], c_rk_tab.e[0]);  // rk.rs:386
#pragma unroll
        for (int ind = 1; ind < O; ++ind)
#pragma unroll
            for (int d = 0; d < D; ++d) sp[d] = __dadd_rn(sp[d], __dmul_rn(hs[ind][d], c_rk_tab.e[ind]));
        double ss = 0.0;
#pragma unroll
        for (int d = 0; d < D; ++d) ss = __dadd_rn(ss, __dmul_rn(sp[d], sp[d]));
        const double error = __ddiv_rn(__dsqrt_rn(ss), dt);  // rk.rs:390

        if (error != error) return BACON_E_NONFINITE;  // D8

        const bool accepted = error <= tol;  // rk.rs:392-398
        if (accepted) {
            t = __dadd_rn(t, dt);
#pragma unroll
            for (int ind = 0; ind < O; ++ind)
#pragma unroll
                for (int d = 0; d < D; ++d) y[d] = __dadd_rn(y[d], __dmul_rn(hs[ind][d], c_rk_tab.b[ind]));
        }

        const double ratio = __ddiv_rn(tol, error);  // rk.rs:400-408
        const double delta = __dmul_rn(c_rk_tab.safety, __dsqrt_rn(__dsqrt_rn(ratio)));
        if (delta <= 0.1) dt = __dmul_rn(dt, 0.1);
        else if (delta >= 4.0) dt = __dmul_rn(dt, 4.0);
        else dt = __dmul_rn(dt, delta);
        if (dt > dt_max) dt = dt_max;  // rk.rs:410-412

        if (dt < dt_min && t < t_end) {  // rk.rs:414-416
            if (!accepted) n_rej++;
            return BACON_E_MIN_DT_EXCEEDED;
        }
        if (accepted) {
            n_acc++;
            yielded = true;
        } else {
            n_rej++;
        }
        return -1;
    }
};

}  // namespace bacon
