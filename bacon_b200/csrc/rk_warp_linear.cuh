// rk_warp_linear.cuh — K5: wide-state variant of the RK stepper for y' = A y with a
// per-trajectory dense A (BASELINE config 4: D = 32, dense output).  7 x 32 doubles of
// stepper state plus a 32 x 32 matrix cannot live in one thread, so ONE WARP integrates
// one trajectory: lane i owns component i of y and of every stage derivative and 32 entries of A in registers (strict
// build: row i; fast build: half of row i and half of row i ^ 16, see matvec below).  The stage vector is exchanged
// through a 256-byte shared buffer per warp (one conflict-free STS.64 per lane, then 16 broadcast LDS.128 — 8 per
// half-warp in the fast build), so a right-hand side costs 32 DFMA per lane = exactly the 2*32*32 algorithmic flops.
// The controller (error norm, accept, dt) is warp-uniform: every lane holds the same
// t, dt and counters, the norm is an xor-shuffle tree (fast) or the oracle's sequential
// sum (strict).  Dense output needs no staging here: an accepted point is one coalesced
// 256-byte row of hist_y written by the warp.
//
// Same reference statements as rk_fast.cuh / rk_strict.cuh (src/ivp/rk.rs:361-423).
#pragma once
#include "ivp_common.cuh"
#include "path_query.cuh"
#include "rk_fast.cuh"
#include "rk_strict.cuh"
#include "tableaux.cuh"

namespace bacon {

constexpr int WARP_BLOCK = 128;  // 4 warps = 4 trajectories in flight per CTA
#ifndef LIN32_MINB
#define LIN32_MINB 4  // resident CTAs per SM the plain kernels are compiled for (128 registers)
#endif

// EVENT: the instantiation that watches a terminal event (bacon_ivp_options::event_w; drive.cuh has the thread-per-
// trajectory form): g = w . y summed over the lanes in the oracle's order on every accepted point, the crossing located
// on the Hermite cubic of that step with f = A y at both ends, all of it warp-uniform.
template <class Tab, bool STRICT, bool HIST, int MINB, bool EVENT = false>
__global__ void __launch_bounds__(WARP_BLOCK, MINB) rk_warp_linear32_kernel(const __grid_constant__ bacon_launch_args a) {
    constexpr int N = 32;
    constexpr int O = Tab::O;
    __shared__ __align__(16) double s_y[WARP_BLOCK / 32][N];

    const unsigned lane = lane_id();
    const unsigned warp = threadIdx.x >> 5;
    double* sy = s_y[warp];
    const unsigned long long n = a.n;
    const double t_start = a.cfg.t_start, t_end = a.cfg.t_end;
    const double dt_min = a.cfg.dt_min, dt_max = a.cfg.dt_max, tol = a.cfg.tol;
    const double tol2 = tol * tol, inv_tol2 = 1.0 / tol2;
    double dt0 = STRICT ? __dmul_rn(__dadd_rn(dt_max, dt_min), 0.5) : (dt_max + dt_min) * 0.5;
    initial_dt_given(a, dt_min, dt_max, dt0);
    const uint32_t cap = (a.cfg.max_attempts == 0 || a.cfg.max_attempts > 0xFFFFFFFEull) ? 0xFFFFFFFEu
                                                                                         : (uint32_t)a.cfg.max_attempts;
    const uint32_t hcap = (uint32_t)a.cfg.history_capacity;
    const bool aos = (a.cfg.flags & BACON_FLAG_PARAMS_AOS) != 0;
    const bool shared = (a.cfg.flags & BACON_FLAG_SHARED_PARAMS) != 0;
    [[maybe_unused]] const RkTableauRt& c_rk_tab = c_rk_tabs[rk_tab_slot(O, a.cfg.semantics)];  // (strict only)

    // y' = A y : dy_i = sum_j A[i][j] Y[j], Y broadcast from shared memory.
    // Strict build: lane i holds row i and sums it in the oracle's order.  Fast build: lane (h, r) = (lane >> 4,
    // lane & 15) holds the column half h of its own row r + 16 h (A[0..15]) and of its partner's row r + 16 (1 - h)
    // (A[16..31]), so it needs 16 values of Y instead of 32, and each of them feeds two DFMA.  The kernel is bound by the shared-memory data pipe, not by the FP64 pipe
    // (ncu: l1tex__data_pipe_lsu_wavefronts 91 % busy with the row layout, profiles/r04d_cfg4_source.md): a broadcast
    // LDS.128 takes two wavefronts, and it still takes two when each HALF-warp reads its own address
    // (tools/lds_pattern_probe.cu) — 16 wavefronts per product instead of 32.  The two halves of a row are joined by one
    // 64-bit shuffle with lane ^ 16; every lane ends up with component `lane` as before.
    const unsigned half = lane >> 4;
    auto matvec = [&](const double (&A)[N], double Yi) -> double {
        __syncwarp();
        sy[lane] = Yi;
        __syncwarp();
        if constexpr (STRICT) {  // the oracle's order: s = A[i][0]*y[0]; s += A[i][j]*y[j]
            const double2* v = reinterpret_cast<const double2*>(sy);
            double s = 0.0;
#pragma unroll
            for (int j2 = 0; j2 < N / 2; ++j2) {
                const double2 yy = v[j2];
                s = (j2 == 0) ? __dmul_rn(A[0], yy.x) : __dadd_rn(s, __dmul_rn(A[2 * j2], yy.x));
                s = __dadd_rn(s, __dmul_rn(A[2 * j2 + 1], yy.y));
            }
            return s;
        } else {  // two rows x two chains
            const double2* v = reinterpret_cast<const double2*>(sy + 16 * half);
            double c00 = 0.0, c01 = 0.0, c10 = 0.0, c11 = 0.0;
#pragma unroll
            for (int j2 = 0; j2 < N / 4; ++j2) {
                const double2 p = v[j2];
                c00 = fma(A[2 * j2], p.x, c00);
                c10 = fma(A[16 + 2 * j2], p.x, c10);
                c01 = fma(A[2 * j2 + 1], p.y, c01);
                c11 = fma(A[16 + 2 * j2 + 1], p.y, c11);
            }
            // (this lane's own row first, the partner's second: nothing to select)
            return (c00 + c01) + __shfl_xor_sync(FULL_MASK, c10 + c11, 16);
        }
    };
    // where A[j] of this lane sits in the row-major 32 x 32 matrix
    auto a_index = [&](int j) -> unsigned {
        if constexpr (STRICT) return lane * N + j;
        else return ((lane & 15u) + 16u * (j < 16 ? half : 1u - half)) * N + 16u * half + (unsigned)(j & 15);
    };

    for (;;) {
        // ---- one trajectory per warp from the work counter
        unsigned long long idx = 0;
        if (lane == 0) idx = atomicAdd(a.work_counter, 1ull);
        idx = __shfl_sync(FULL_MASK, idx, 0);
        if (idx >= n) break;

        double A[N];
        {
            const double* P = a.params;
            if (shared) {
#pragma unroll
                for (int j = 0; j < N; ++j) A[j] = P[a_index(j)];
            } else if (aos) {  // [n][N*N] row-major per trajectory: 128-bit loads along the lane's row(s)
                const double* M = P + (size_t)idx * (N * N);
#pragma unroll
                for (int j2 = 0; j2 < N / 2; ++j2) {
                    const double2 v = *reinterpret_cast<const double2*>(M + a_index(2 * j2));
                    A[2 * j2] = v.x;
                    A[2 * j2 + 1] = v.y;
                }
            } else {  // [N*N][n] SoA
#pragma unroll
                for (int j = 0; j < N; ++j) A[j] = P[(size_t)a_index(j) * n + idx];
            }
        }
        double y = a.y0[(size_t)lane * n + idx];
        double t = t_start, dt = dt0;
        trajectory_start(a, idx, dt_min, dt_max, t, dt);  // restart record (bacon_ivp_options)
        uint32_t n_acc = 0, n_rej = 0, n_att = 0;
        int st = -1;
        double k[O];  // strict: half_steps (k_j = f_j*dt, persistent); fast: unscaled f_j
#pragma unroll
        for (int j = 0; j < O; ++j) k[j] = 0.0;
        // terminal event: the last knot (knot 0 = the initial condition)
        auto g_of = [&](double y_lane) -> double {  // w . y - c, sequential in d like the oracle; every lane gets it
            double s = a.ev_w[0] * __shfl_sync(FULL_MASK, y_lane, 0);
#pragma unroll
            for (int d = 1; d < N; ++d) s += a.ev_w[d] * __shfl_sync(FULL_MASK, y_lane, d);
            return s - a.ev_c;
        };
        [[maybe_unused]] double ev_tp = t, ev_yp = y, ev_gp = 0.0;
        const bool ev_on = EVENT && a.ev_on != 0;  // (a launch with only a restart record runs these kernels too)
        if (ev_on) ev_gp = g_of(y);
        // fast build: k[0] = f(y) is carried from attempt to attempt (below); the first one is computed here
        if constexpr (!STRICT) k[0] = matvec(A, y);

        while (st < 0) {
            if (n_att >= cap) { st = BACON_E_MAX_ATTEMPTS; break; }
            if (t >= t_end) { st = BACON_OK; break; }  // rk.rs:362-364
            bool accepted;
            if constexpr (STRICT) {
                if (__dadd_rn(t, dt) >= t_end) dt = __dadd_rn(t_end, -t);  // rk.rs:366-368
#pragma unroll
                for (int i = 0; i < O; ++i) {  // rk.rs:370-384, dense rows
                    double sp = y;
#pragma unroll
                    for (int j = 0; j < O; ++j) sp = __dadd_rn(sp, __dmul_rn(k[j], c_rk_tab.a[i][j]));
                    k[i] = __dmul_rn(matvec(A, sp), dt);
                }
                n_att++;
                double sp = __dmul_rn(k[0], c_rk_tab.e[0]);
#pragma unroll
                for (int j = 1; j < O; ++j) sp = __dadd_rn(sp, __dmul_rn(k[j], c_rk_tab.e[j]));
                __syncwarp();
                sy[lane] = sp;
                __syncwarp();
                double ss = 0.0;  // nalgebra norm(): sequential sum of squares
#pragma unroll
                for (int d = 0; d < N; ++d) ss = __dadd_rn(ss, __dmul_rn(sy[d], sy[d]));
                const double error = __ddiv_rn(__dsqrt_rn(ss), dt);
                if (error != error) { st = BACON_E_NONFINITE; break; }
                accepted = error <= tol;
                if (accepted) {
                    t = __dadd_rn(t, dt);
#pragma unroll
                    for (int j = 0; j < O; ++j) y = __dadd_rn(y, __dmul_rn(k[j], c_rk_tab.b[j]));
                }
                const double delta = __dmul_rn(c_rk_tab.safety, __dsqrt_rn(__dsqrt_rn(__ddiv_rn(tol, error))));
                if (delta <= 0.1) dt = __dmul_rn(dt, 0.1);
                else if (delta >= 4.0) dt = __dmul_rn(dt, 4.0);
                else dt = __dmul_rn(dt, delta);
                if (dt > dt_max) dt = dt_max;
            } else {
                double h = dt;
                if (t + h >= t_end) h = t_end - t;
                // (k[0] = A y is already there: computed at the end of the attempt that produced y)
                static_for<1, O>([&](auto I) {
                    constexpr int i = decltype(I)::value;
                    constexpr int j0 = first_nz_a<Tab, i>();
                    double s = Tab::a(i, j0) * k[j0];
                    static_for<j0 + 1, i>([&](auto J) {
                        constexpr int j = decltype(J)::value;
                        if constexpr (Tab::a(i, j) != 0.0) s = fma(Tab::a(i, j), k[j], s);
                    });
                    k[i] = matvec(A, fma(h, s, y));
                });
                n_att++;
                constexpr int e0 = first_nz_e<Tab>();
                double s = Tab::e(e0) * k[e0];
                static_for<e0 + 1, O>([&](auto J) {
                    constexpr int j = decltype(J)::value;
                    if constexpr (Tab::e(j) != 0.0) s = fma(Tab::e(j), k[j], s);
                });
                double q = s * s;
                // The point an accepted step lands on and the NEXT attempt's first stage there, before the verdict:
                // 99 % of the attempts are accepted, the 32 DFMA of this product depend on nothing below, and the
                // norm's shuffle tree and the step factor are two long dependent chains without FP64 work (ncu,
                // profiles/r04d_cfg4_source.md: 38 % of a warp's time per attempt) — issued together they overlap.
                // A rejected attempt keeps y and with it k[0] = A y; the numbers are those of the plain order.
                constexpr int b0 = first_nz_b<Tab>();
                double sb = Tab::b(b0) * k[b0];
                static_for<b0 + 1, O>([&](auto J) {
                    constexpr int j = decltype(J)::value;
                    if constexpr (Tab::b(j) != 0.0) sb = fma(Tab::b(j), k[j], sb);
                });
                const double y_try = fma(h, sb, y);
                const double k0_try = matvec(A, y_try);
#pragma unroll
                for (int m = 16; m >= 1; m >>= 1) q += __shfl_xor_sync(FULL_MASK, q, m);  // identical on every lane
                if (q != q) { st = BACON_E_NONFINITE; break; }
                accepted = q <= tol2;
                if (accepted) {
                    t += h;
                    y = y_try;
                    k[0] = k0_try;
                }
                const double x = fmax(q * inv_tol2, 1e-6);
                const double delta = fmin(fmax(Tab::safety * inv_eighth_root(x), 0.1), 4.0);
                dt = fmin(h * delta, dt_max);
            }
            if (dt < dt_min && t < t_end) {  // rk.rs:414-416
                if (!accepted) n_rej++;
                st = BACON_E_MIN_DT_EXCEEDED;
                break;
            }
            if constexpr (EVENT) {
                if (accepted && ev_on) {
                    const double gb = g_of(y);
                    if (event_crossing(ev_gp, gb, a.ev_direction)) {  // rare: the trajectory ends at the crossing
                        const double fa[1] = {matvec(A, ev_yp)}, fb[1] = {matvec(A, y)};
                        const double hh = t - ev_tp;
                        double da = a.ev_w[0] * __shfl_sync(FULL_MASK, fa[0], 0), db = a.ev_w[0] * __shfl_sync(FULL_MASK, fb[0], 0);
#pragma unroll
                        for (int d = 1; d < N; ++d) {
                            da += a.ev_w[d] * __shfl_sync(FULL_MASK, fa[0], d);
                            db += a.ev_w[d] * __shfl_sync(FULL_MASK, fb[0], d);
                        }
                        const double th = hermite_root(ev_gp, gb, hh * da, hh * db);
                        const double ya[1] = {ev_yp}, yb[1] = {y};
                        double ye[1];
                        hermite_eval<1>(th, hh, ya, yb, fa, fb, ye);
                        y = ye[0];
                        t = ev_tp + th * hh;
                        st = BACON_STOPPED_AT_EVENT;  // (the point that crossed is not yielded)
                        break;
                    }
                    ev_gp = gb;
                    ev_tp = t;
                    ev_yp = y;
                }
            }
            if (accepted) {
                if (HIST && n_acc < hcap) {  // the yielded point (rk.rs:418-419): one coalesced row
                    const size_t row = (size_t)idx * hcap + n_acc;
                    double* rec = a.out.hist + row * (N + 1);  // (t, y[0..N)) record, 264 contiguous bytes per point
                    rec[1 + lane] = y;
                    if (lane == 0) rec[0] = t;
                }
                n_acc++;
            } else {
                n_rej++;
            }
        }

        // ---- retire
        if (HIST && st == BACON_OK && n_acc > hcap) st = BACON_E_HISTORY_OVERFLOW;
        a.out.y_end[(size_t)lane * n + idx] = y;
        if (lane == 0) {
            if (a.out.t_end) a.out.t_end[idx] = t;
            if (a.out.dt_end) a.out.dt_end[idx] = dt;
            a.out.status[idx] = st;
            if (a.out.n_accept) a.out.n_accept[idx] = n_acc;
            if (a.out.n_reject) a.out.n_reject[idx] = n_rej;
            if (a.out.n_rhs) a.out.n_rhs[idx] = n_att * (uint32_t)O;
            if (HIST && a.out.hist_len) a.out.hist_len[idx] = n_acc < hcap ? n_acc : hcap;
        }
    }
}

template <class K> inline int launch_persistent_warp(K kernel, bacon_launch_args* a) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, WARP_BLOCK, 0) != cudaSuccess || per_sm < 1)
        return BACON_E_CUDA;
    cudaFuncAttributes fa;
    if (cudaFuncGetAttributes(&fa, kernel) != cudaSuccess) return BACON_E_CUDA;
    long long grid = (long long)per_sm * a->sm_count;
    const long long need = (long long)((a->n + (WARP_BLOCK / 32) - 1) / (WARP_BLOCK / 32));
    if (grid > need) grid = need;
    if (a->grid_override > 0) grid = a->grid_override;
    if (grid < 1) grid = 1;
    a->grid = (int)grid;
    a->block = WARP_BLOCK;
    a->regs_per_thread = fa.numRegs;
    a->n_kernels = 1;
    kernel<<<(unsigned)grid, WARP_BLOCK, 0, (cudaStream_t)a->stream>>>(*a);
    return cudaGetLastError() == cudaSuccess ? 0 : BACON_E_CUDA;
}

template <class Tab, bool STRICT, bool EVENT = false> int launch_rk_warp_linear32(bacon_launch_args* a) {
    if (STRICT) {
        RkTableauRt T;
        fill_runtime_tableau<Tab>(T, a->cfg.semantics == BACON_SEM_LITERAL);
        if (cudaMemcpyToSymbolAsync(c_rk_tabs, &T, sizeof(T), sizeof(T) * rk_tab_slot(Tab::O, a->cfg.semantics),
                                    cudaMemcpyHostToDevice, (cudaStream_t)a->stream) != cudaSuccess)
            return BACON_E_CUDA;
    } else if (a->cfg.semantics != BACON_SEM_CORRECTED) {
        return BACON_E_UNSUPPORTED;
    }
    if constexpr (EVENT) {
        if (a->cfg.history_capacity > 0 && a->out.hist)
            return launch_persistent_warp(rk_warp_linear32_kernel<Tab, STRICT, true, 4, true>, a);
        return launch_persistent_warp(rk_warp_linear32_kernel<Tab, STRICT, false, 4, true>, a);
    } else {
        if (a->ev_on) return BACON_E_UNSUPPORTED;
        if (a->cfg.history_capacity > 0 && a->out.hist)
            return launch_persistent_warp(rk_warp_linear32_kernel<Tab, STRICT, true, LIN32_MINB>, a);
        return launch_persistent_warp(rk_warp_linear32_kernel<Tab, STRICT, false, LIN32_MINB>, a);
    }
}

}  // namespace bacon
