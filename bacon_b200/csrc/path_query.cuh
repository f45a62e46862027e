// path_query.cuh — queries on stored paths: the continuous extension of a dense-output solve.
//
// NOT in the reference: its `Path` is the accepted points and nothing between them (src/ivp.rs:203-211).  These two
// kernel families are the step after the path: they read the record-layout history a solve left in HBM
// (hist[n][cap][1 + D], hist_stage.cuh) and evaluate, between two neighbouring knots (t_a, y_a), (t_b, y_b), the
// cubic Hermite interpolant with the right-hand side's own slopes f_a = f(t_a, y_a), f_b = f(t_b, y_b):
//     theta = (tau - t_a) / h,  h = t_b - t_a
//     y(tau) = (1 - theta) y_a + theta y_b + theta (theta - 1) [(1 - 2 theta)(y_b - y_a) + (theta - 1) h f_a + theta h f_b]
// (exact at both knots, local error O(h^4)).  Knot 0 of a trajectory is its initial condition (the steppers do not yield
// it: rk.rs:418-419), knots 1..m are its records, and the stepper's exit state (t_end, y_end), when it lies beyond the
// last record (a failed trajectory, or the unyielded last block of SURVEY.md D9; Euler, whose records are the OLD
// points, ivp.rs:331-337), closes the path.
//
//   path_sample_kernel   one thread per (trajectory, sample time): a search over the trajectory's knot times
//                        (bisection down to 64 knots, then interpolation: PathView::first_knot_at_or_after), two
//                        records in, D doubles out; neighbouring lanes take neighbouring times of ONE trajectory, so
//                        their probes share lines and their stores are contiguous.
//   path_events_kernel   one warp per trajectory: the warp streams the path 32 records at a time (coalesced: a
//                        record is 8(1 + D) bytes, a D = 3 chunk is 1 KB), each lane tests g(y) = w . y - c for a sign
//                        change against its left neighbour, and the rare lane that finds one locates the root of the
//                        interpolant (safeguarded Newton).  Events come out in path order (ballot + prefix count);
//                        crossings wait in a per-warp queue and are located 32 at a time, one per lane.
// Both are HBM-bound: the events kernel reads every record once (8(1 + D) bytes per accepted step), the sample kernel
// touches 2 records + D outputs per sample.  Operation order matches oracle/oracle_capi.cpp (oracle_sample_paths,
// oracle_locate_events): the strict build (-fmad=false) is bit-comparable with it.
#pragma once
#ifndef __CUDACC_RTC__
#include <cuda_runtime.h>

#include <cstdlib>
#endif

#include "ivp_common.cuh"

#define BACON_PATH_SAMPLE 0
#define BACON_PATH_EVENTS 1
#define BACON_PATH_MAX_DIM 32

// Plain C layout: crosses the C ABI inside bacon_rhs_desc::path_query.
struct bacon_path_args {
    bacon_ivp_config cfg;       // of the solve that produced the paths
    unsigned long long n;
    const double* y0;           // [dim][n] device
    const double* params;       // as given to the solve (layout flags in cfg.flags)
    const double* hist;         // [n][cap][1 + dim]
    const uint32_t* hist_len;   // [n]
    const double* t_end;        // [n] or NULL
    const double* y_end;        // [dim][n] or NULL (both or neither)
    const uint32_t* n_accept;   // [n] or NULL: points the solve yielded (> capacity: the stored path was cut short)
    const int32_t* status;      // [n] or NULL (used when n_accept is not given)
    const double* t_start_each; // [n] or NULL: per-trajectory start times (a resumed leg, bacon_ivp_options); else cfg.t_start
    int32_t op;                 // BACON_PATH_SAMPLE / BACON_PATH_EVENTS
    // sampling
    unsigned long long n_times;
    const double* times;        // [n_times] device
    double* samples;            // [n][n_times][dim]
    // events of g(y) = w . y - c
    double ev_w[BACON_PATH_MAX_DIM];
    double ev_c;
    int32_t ev_direction;       // +1 rising, -1 falling, 0 both
    int32_t ev_capacity;
    double* events;             // [n][ev_capacity][1 + dim]
    uint32_t* n_events;         // [n]
    void* stream;               // cudaStream_t
    // filled by the launcher
    int32_t grid, block, regs_per_thread;
};

namespace bacon {

constexpr int PATH_BLOCK = 128;
// Knot search of the sampling kernels (PathView::first_knot_at_or_after): bisection above this bracket width,
// interpolation below it, at most this many times (an adversarial path cannot make it a linear scan).
#ifndef PATH_INTERP_BELOW
#define PATH_INTERP_BELOW 64u
#endif
#ifndef PATH_INTERP_MAX
#define PATH_INTERP_MAX 8u
#endif
// Events kernel: chunks of 32 records in flight per lane, and resident CTAs per SM it is compiled for.  Measured on
// config 2's history (profiles/r01o_path_queries.md): 4 chunks at 4 CTAs/SM (114 registers, nothing spilled) 4.67 TB/s;
// 2 chunks at 8 CTAs (64 registers, spills on the cold path) 4.84; 4 chunks at 6 CTAs (80 registers) spills inside the
// loop: 1.0 TB/s.  The default is the spill-free one; wider records take fewer chunks to stay inside 128 registers.
#ifndef BACON_EV_MINB
#define BACON_EV_MINB 4
#endif
#ifndef BACON_EV_WIDE_MINB
#define BACON_EV_WIDE_MINB 4
#endif
// Wide records (1 + D > 8, linear32): each lane sums its record's g straight from memory (BACON_EV_WIDE_STAGE 0), or
// the warp first copies the chunk into shared memory with coalesced loads (1).  Measured on config 4's history (2^18
// paths of ~122 records of 264 B, 8.5 GB; profiles/r01o_path_queries.md): 3.66 ms direct against 3.87 staged, and
// 4.06 at 6 or 8 resident CTAs either way — neither the access pattern nor the occupancy is the limit there; a warp
// has only four chunks to stream per path, so its dependent prologue (length, last record, closing knot) dominates.
#ifndef BACON_EV_WIDE_STAGE
#define BACON_EV_WIDE_STAGE 0
#endif
template <int D> __host__ __device__ constexpr int ev_minb() { return 1 + D > 8 ? BACON_EV_WIDE_MINB : BACON_EV_MINB; }
template <int D> __host__ __device__ constexpr int ev_unroll() {
#ifdef BACON_EV_UNROLL
    return BACON_EV_UNROLL;
#else
    return 1 + D <= 4 ? 4 : (1 + D <= 8 ? 2 : 1);
#endif
}

template <int D>
__device__ __forceinline__ void hermite_eval(double th, double h, const double (&ya)[D], const double (&yb)[D],
                                             const double (&fa)[D], const double (&fb)[D], double (&out)[D]) {
    const double om = 1.0 - th, tt = th * (th - 1.0), c0 = 1.0 - 2.0 * th, c1 = th - 1.0;
#pragma unroll
    for (int d = 0; d < D; ++d) {
        const double dy = yb[d] - ya[d];
        const double v = (c0 * dy + c1 * (h * fa[d])) + th * (h * fb[d]);
        out[d] = (om * ya[d] + th * yb[d]) + tt * v;
    }
}

// The zero in [0, 1] of the scalar Hermite cubic p through p(0) = ga, p(1) = gb (opposite signs, ga != 0) with end
// slopes A = h p'(0), B = h p'(1): Newton from the secant point, kept inside the bracket the signs maintain (a step
// that leaves it is replaced by the midpoint), until a step moves theta by no more than 1e-15.  A handful of iterations
// where bisection to the last bit takes 52; oracle/oracle_capi.cpp (hermite_root) repeats it operation for operation.
__device__ __forceinline__ double hermite_root(double ga, double gb, double A, double B) {
    if (gb == 0.0) return 1.0;
    const double dg = gb - ga, vs = (A + B) - 2.0 * dg;  // v'(theta) is constant
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
        const double moved = fabs(tn - th);
        th = tn;
        if (moved <= 1e-15) break;
    }
    return th;
}

// One whole (t, y) record with the widest loads its alignment allows (history base and records of 1 + D = 4k doubles are
// 32-byte aligned: one 256-bit load, the mirror of hist_stage.cuh's store; 128-bit when 1 + D is even).
template <int R> __device__ __forceinline__ void load_record(const double* __restrict__ src, double (&r)[R]) {
    if constexpr (R % 4 == 0) {
#pragma unroll
        for (int j = 0; j < R; j += 4) {
#if defined(__CUDACC_VER_MAJOR__) && (__CUDACC_VER_MAJOR__ * 100 + __CUDACC_VER_MINOR__ < 1209)
            // 256-bit vector loads need PTX ISA 8.8 (CUDA 12.9): an older NVRTC (bacon_rhs_register_source picks up whatever
            // libnvrtc.so.12 the process has) reads the same sector as two 128-bit halves (as hist_stage.cuh's store)
            asm volatile("ld.global.v2.f64 {%0, %1}, [%4];\n\tld.global.v2.f64 {%2, %3}, [%4+16];"
                         : "=d"(r[j]), "=d"(r[j + 1]), "=d"(r[j + 2]), "=d"(r[j + 3]) : "l"(src + j));
#else
            asm volatile("ld.global.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(r[j]), "=d"(r[j + 1]), "=d"(r[j + 2]), "=d"(r[j + 3]) : "l"(src + j));
#endif
        }
    } else if constexpr (R % 2 == 0) {
#pragma unroll
        for (int j = 0; j < R; j += 2) {
            const double2 v = *reinterpret_cast<const double2*>(src + j);
            r[j] = v.x;
            r[j + 1] = v.y;
        }
    } else {
#pragma unroll
        for (int j = 0; j < R; ++j) r[j] = src[j];
    }
}

// One trajectory's knots: 0 = (t_start, y0), 1..m = records, m + 1 = the closing knot when there is one.
template <int D> struct PathView {
    static constexpr int R = 1 + D;
    const double* rec;
    const double* y0;
    const double* y_end;
    unsigned long long n, i;
    uint32_t m;
    bool closing;
    double t0, tc;

    __device__ __forceinline__ PathView(const bacon_path_args& a, unsigned long long i_) {
        i = i_;
        n = a.n;
        const uint32_t cap = (uint32_t)a.cfg.history_capacity;
        rec = a.hist + (size_t)i * cap * R;
        y0 = a.y0;
        y_end = a.y_end;
        const uint32_t len = a.hist_len[i];
        m = len < cap ? len : cap;
        t0 = a.t_start_each ? a.t_start_each[i] : a.cfg.t_start;
        closing = false;
        tc = 0.0;
        // A path that overflowed its capacity ENDS at its last record: (t_end, y_end) lies a whole unrecorded span
        // later, and one cubic must not bridge it — times past the last record give NaN, as the header promises.
        const bool cut = a.n_accept ? a.n_accept[i] > cap : (a.status ? a.status[i] == BACON_E_HISTORY_OVERFLOW : false);
        if (a.t_end && a.y_end && !cut) {
            tc = a.t_end[i];
            closing = tc > (m > 0 ? rec[(size_t)(m - 1) * R] : t0);
        }
    }
    __device__ __forceinline__ uint32_t last() const { return m + (closing ? 1u : 0u); }
    __device__ __forceinline__ double time(uint32_t k) const {
        return k == 0 ? t0 : (k <= m ? rec[(size_t)(k - 1) * R] : tc);
    }
    // The first knot at or after tau, for t0 <= tau <= t_last = time(K), K >= 1.  Every probe of a knot's time costs a
    // whole DRAM burst for 8 bytes, so the search is laid out for bursts, not for comparisons: plain bisection while
    // the bracket is wider than PATH_INTERP_BELOW knots (the sample times of a warp are neighbours, so these probes are
    // shared and come from L2), then interpolation on the bracket's end times — the step size of an adaptive path
    // varies slowly, over 64 knots the guess is within 1.2 knots (median; 3.6 at the 90th percentile on Lorenz) — at
    // most PATH_INTERP_MAX times, then bisection again.  The last probes fall into the bursts of the two records the
    // caller reads anyway: 5.6 -> 2.9 distinct bursts per sample on config 2's paths, 12 -> 8.5 dependent probes.
    // Whatever the probe sequence, the answer is the same knot.
    __device__ __forceinline__ uint32_t first_knot_at_or_after(double tau, uint32_t K, double t_last) const {
        uint32_t lo = 1, hi = K, n_interp = 0;
        double tl = t0, th = t_last;  // the times of knots lo - 1 and hi
        while (lo < hi) {
            const uint32_t width = hi - lo;
            uint32_t mid = (lo + hi) >> 1;
            if (width <= PATH_INTERP_BELOW && width > 1 && n_interp < PATH_INTERP_MAX && th > tl) {
                ++n_interp;
                const float x = __fdividef((float)(tau - tl), (float)(th - tl)) * (float)(width + 1);
                const uint32_t g = (lo - 1) + (uint32_t)ceilf(fminf(fmaxf(x, 0.0f), (float)(width + 1)));
                mid = g < lo ? lo : (g > hi - 1 ? hi - 1 : g);
            }
            const double tm = time(mid);
            if (tm >= tau) {
                hi = mid;
                th = tm;
            } else {
                lo = mid + 1;
                tl = tm;
            }
        }
        return lo;
    }
    // time and state of knot k; a record comes in with one vector load
    __device__ __forceinline__ double knot(uint32_t k, double (&y)[D]) const {
        if (k >= 1 && k <= m) {
            double r[R];
            load_record<R>(rec + (size_t)(k - 1) * R, r);
#pragma unroll
            for (int d = 0; d < D; ++d) y[d] = r[1 + d];
            return r[0];
        }
        state(k, y);
        return k == 0 ? t0 : tc;
    }
    __device__ __forceinline__ void state(uint32_t k, double (&y)[D]) const {
        if (k == 0) {
#pragma unroll
            for (int d = 0; d < D; ++d) y[d] = y0[(size_t)d * n + i];
        } else if (k <= m) {
            const double* r = rec + (size_t)(k - 1) * R + 1;
#pragma unroll
            for (int d = 0; d < D; ++d) y[d] = r[d];
        } else {
#pragma unroll
            for (int d = 0; d < D; ++d) y[d] = y_end[(size_t)d * n + i];
        }
    }
};

template <int P>
__device__ __forceinline__ void load_path_params(const bacon_path_args& a, unsigned long long i, double (&p)[(P > 0 ? P : 1)]) {
    p[0] = 0.0;
    if constexpr (P > 0) {
        const bool shared = (a.cfg.flags & BACON_FLAG_SHARED_PARAMS) != 0;
        const bool aos = (a.cfg.flags & BACON_FLAG_PARAMS_AOS) != 0;
#pragma unroll
        for (int k = 0; k < P; ++k)
            p[k] = shared ? a.params[k] : (aos ? a.params[(size_t)i * P + k] : a.params[(size_t)k * a.n + i]);
    }
}

__device__ __forceinline__ double path_nan() { return __longlong_as_double(0x7ff8000000000000ll); }

template <class Rhs, bool STRICT>
__global__ void __launch_bounds__(PATH_BLOCK) path_sample_kernel(const __grid_constant__ bacon_path_args a) {
    constexpr int D = Rhs::DIM;
    constexpr int P = Rhs::NPARAM;
    const unsigned long long g = (unsigned long long)blockIdx.x * PATH_BLOCK + threadIdx.x;
    if (g >= a.n * a.n_times) return;
    const unsigned long long i = g / a.n_times, j = g - i * a.n_times;
    const PathView<D> pv(a, i);
    const double tau = a.times[j];
    double* out = a.samples + (size_t)g * D;
    const uint32_t K = pv.last();
    double res[D];
    const double t_last = pv.time(K);
    if (K == 0 || !(tau >= pv.t0 && tau <= t_last)) {
        if (tau == pv.t0) {
            pv.state(0, res);
        } else {
#pragma unroll
            for (int d = 0; d < D; ++d) res[d] = path_nan();
        }
    } else {
        // (Interpolating from the whole path's ends instead needs 7 probes on average but up to 17 on the slowest lane of
        // a warp, and the warp waits for that lane: measured 15 % slower than plain bisection in round 1.)
        const uint32_t lo = pv.first_knot_at_or_after(tau, K, t_last);
        double ya[D], yb[D], fa[D], fb[D], p[P > 0 ? P : 1];
        const double ta = pv.knot(lo - 1, ya), tb = pv.knot(lo, yb);
        load_path_params<P>(a, i, p);
        const Rhs rhs{};
        rhs(ta, ya, p, fa);
        rhs(tb, yb, p, fb);
        const double h = tb - ta;
        const double th = h > 0.0 ? (tau - ta) / h : 0.0;
        hermite_eval<D>(th, h, ya, yb, fa, fb, res);
    }
#pragma unroll
    for (int d = 0; d < D; ++d) out[d] = res[d];
}

template <int D> __device__ __forceinline__ double event_fn(const bacon_path_args& a, const double (&y)[D]) {
    double s = a.ev_w[0] * y[0];
#pragma unroll
    for (int d = 1; d < D; ++d) s += a.ev_w[d] * y[d];
    return s - a.ev_c;
}

__device__ __forceinline__ bool event_crossing(double ga, double gb, int direction) {
    const bool rising = ga < 0.0 && gb >= 0.0, falling = ga > 0.0 && gb <= 0.0;
    return direction > 0 ? rising : (direction < 0 ? falling : (rising || falling));
}

// (not inlined: rare, and its registers should not weigh on the streaming loop)
template <class Rhs>
__device__ __noinline__ void locate_event(const bacon_path_args& a, unsigned long long i, uint32_t k, double* dst) {
    constexpr int D = Rhs::DIM;
    constexpr int P = Rhs::NPARAM;
    const PathView<D> pv(a, i);  // (rebuilt here: passing the caller's by reference would put it on the stack)
    double ya[D], yb[D], fa[D], fb[D], p[P > 0 ? P : 1];
    const double ta = pv.knot(k - 1, ya), tb = pv.knot(k, yb);
    load_path_params<P>(a, i, p);
    const Rhs rhs{};
    rhs(ta, ya, p, fa);
    rhs(tb, yb, p, fb);
    const double h = tb - ta;
    // g is linear, so g(interpolant) is the Hermite cubic through (g_a, w . f_a), (g_b, w . f_b)
    const double ga[1] = {event_fn<D>(a, ya) }, gb[1] = {event_fn<D>(a, yb)};
    double da[1], db[1];
    {
        double s = a.ev_w[0] * fa[0], r = a.ev_w[0] * fb[0];
#pragma unroll
        for (int d = 1; d < D; ++d) {
            s += a.ev_w[d] * fa[d];
            r += a.ev_w[d] * fb[d];
        }
        da[0] = s;
        db[0] = r;
    }
    const double th = hermite_root(ga[0], gb[0], h * da[0], h * db[0]);
    double ys[D];
    hermite_eval<D>(th, h, ya, yb, fa, fb, ys);
    dst[0] = ta + th * h;
#pragma unroll
    for (int d = 0; d < D; ++d) dst[1 + d] = ys[d];
}

// how a warp's queue of crossings is located: one crossing per lane (thread-sized states)
template <class Rhs> struct LaneLocate {
    static constexpr bool DEFERRED = false;  // crossings are located inside the streaming kernel
    static __device__ __forceinline__ void flush(const bacon_path_args& a, unsigned long long i, uint32_t n_pend, const uint32_t* pk,
                                                 const uint32_t* ps, double* ev, unsigned lane) {
        if (lane < n_pend) locate_event<Rhs>(a, i, pk[lane], ev + (size_t)ps[lane] * (1 + Rhs::DIM));
    }
};

template <class Rhs, bool STRICT, class Locate = LaneLocate<Rhs>>
__global__ void __launch_bounds__(PATH_BLOCK, ev_minb<Rhs::DIM>()) path_events_kernel(const __grid_constant__ bacon_path_args a) {
    constexpr int D = Rhs::DIM;
    const unsigned long long i = ((unsigned long long)blockIdx.x * PATH_BLOCK + threadIdx.x) >> 5;
    if (i >= a.n) return;  // (whole warps leave together)
    const unsigned lane = lane_id();
    const PathView<D> pv(a, i);
    const uint32_t K = pv.last();
    // g of a knot straight from memory (knot 0; and every record of a wide one, whose components are not worth holding)
    auto knot_g = [&](uint32_t k) -> double {
        const double* src = k == 0 ? pv.y0 + i : (k <= pv.m ? pv.rec + (size_t)(k - 1) * (1 + D) + 1 : pv.y_end + i);
        const size_t stride = (k >= 1 && k <= pv.m) ? 1 : (size_t)pv.n;
        double s = a.ev_w[0] * src[0];
#pragma unroll
        for (int d = 1; d < D; ++d) s += a.ev_w[d] * src[(size_t)d * stride];
        return s - a.ev_c;
    };
    const double g_carry = knot_g(0);
    uint32_t count = 0;
    const uint32_t cap = (uint32_t)a.ev_capacity;
    double* ev = a.events + (size_t)i * cap * (1 + D);

    // EV_UNROLL chunks of 32 records per iteration: every lane has EV_UNROLL whole-record loads in flight (4 KB per warp
    // for D = 3) before the first is used.  Crossings are found on warp-wide sign masks (three votes per chunk; the left
    // neighbour's sign is the mask shifted by one, the previous chunk's last lane carried in bit 0) — all of it in
    // uniform registers, no shuffles.  Iterations that lie wholly inside the records run without bounds tests.
    constexpr int R = 1 + D;
    constexpr int EV_UNROLL = ev_unroll<D>();
    __shared__ uint32_t pend_k[PATH_BLOCK / 32][32], pend_slot[PATH_BLOCK / 32][32];
    constexpr bool WIDE_REC = R > 8 && BACON_EV_WIDE_STAGE;
    __shared__ double stage[WIDE_REC ? PATH_BLOCK / 32 : 1][WIDE_REC ? 32 * R : 1];
    const unsigned wid = threadIdx.x >> 5;
    uint32_t n_pend = 0;  // warp-uniform
    auto flush = [&]() {
        __syncwarp();
        Locate::flush(a, i, n_pend, pend_k[wid], pend_slot[wid], ev, lane);
        __syncwarp();
    };
    unsigned carry_neg = g_carry < 0.0 ? 1u : 0u, carry_pos = g_carry > 0.0 ? 1u : 0u;
    auto body = [&](auto full_tag, uint32_t base) {
        constexpr bool FULL = decltype(full_tag)::value;
        constexpr bool WIDE = R > 8;
        constexpr int RR = WIDE ? 1 : R;  // (wide records are not held: their g is summed straight from memory)
        double rec[EV_UNROLL][RR];
        unsigned hits[EV_UNROLL];
        if constexpr (WIDE && WIDE_REC) {
            // a chunk of 32 wide records is one contiguous block: the warp copies it into shared memory with coalesced
            // loads (R per lane in flight), and each lane then sums its own record from there, in component order
            const uint32_t nrec = FULL ? 32u : (base > pv.m ? 0u : (pv.m - base + 1 < 32u ? pv.m - base + 1 : 32u));
            const double* src = pv.rec + (size_t)(base - 1) * R;
            __syncwarp();
            if (FULL) {
#pragma unroll
                for (int it = 0; it < R; ++it) stage[wid][it * 32 + lane] = src[it * 32 + lane];
            } else {
                for (uint32_t idx = lane; idx < nrec * R; idx += 32) stage[wid][idx] = src[idx];
            }
            __syncwarp();
        } else if constexpr (!WIDE) {
#pragma unroll
            for (int u = 0; u < EV_UNROLL; ++u) {
                const uint32_t k = base + 32 * u + lane;
                if (FULL || k <= pv.m) {
                    load_record<RR>(pv.rec + (size_t)(k - 1) * R, rec[u]);
                } else if (k <= K) {  // the closing knot
#pragma unroll
                    for (int d = 0; d < D; ++d) rec[u][(1 + d) % RR] = pv.y_end[(size_t)d * pv.n + i];
                }
            }
        }
#pragma unroll
        for (int u = 0; u < EV_UNROLL; ++u) {
            const uint32_t k = base + 32 * u + lane;
            const bool have = FULL || k <= K;
            double g = 0.0;
            if constexpr (WIDE) {
                if (WIDE_REC && have && k <= pv.m) {
                    const double* r = &stage[wid][(k - base) * R + 1];
                    double sum = a.ev_w[0] * r[0];
#pragma unroll
                    for (int d = 1; d < D; ++d) sum += a.ev_w[d] * r[d];
                    g = sum - a.ev_c;
                } else if (have) {
                    g = knot_g(k);  // the closing knot
                }
            } else {
                double y[D];
#pragma unroll
                for (int d = 0; d < D; ++d) y[d] = have ? rec[u][(1 + d) % RR] : 0.0;
                g = event_fn<D>(a, y);
            }
            const unsigned neg = __ballot_sync(FULL_MASK, have && g < 0.0), pos = __ballot_sync(FULL_MASK, have && g > 0.0);
            const unsigned ord = __ballot_sync(FULL_MASK, have && g == g);  // (a NaN is neither side of the surface)
            const unsigned rising = ((neg << 1) | carry_neg) & ~neg & ord, falling = ((pos << 1) | carry_pos) & ~pos & ord;
            carry_neg = neg >> 31;
            carry_pos = pos >> 31;
            hits[u] = a.ev_direction > 0 ? rising : (a.ev_direction < 0 ? falling : (rising | falling));
        }
        // Crossings are queued (record index, output slot) and located a warp-load at a time, one per lane: locating
        // one costs ~20 streaming iterations' worth of instructions, and done on the spot it ran with 1 lane of 32 —
        // 3/4 of all instructions issued on config 2's history (profiles/r01o_path_queries.md).  Crossings past the
        // capacity are only counted.
        unsigned any = 0;
#pragma unroll
        for (int u = 0; u < EV_UNROLL; ++u) any |= hits[u];
        if (any) {
#pragma unroll 1
            for (int u = 0; u < EV_UNROLL; ++u) {
                unsigned h = 0;
#pragma unroll
                for (int v = 0; v < EV_UNROLL; ++v) h = u == v ? hits[v] : h;  // (no dynamic indexing: registers)
                if (h == 0) continue;
                const uint32_t nh = (uint32_t)__popc(h);
                const uint32_t room = count < cap ? cap - count : 0u;
                const uint32_t n_enq = nh < room ? nh : room;  // slots are handed out in order: the stored ones come first
                if (n_pend + n_enq > 32u) {
                    flush();
                    n_pend = 0;
                }
                const uint32_t rank = (uint32_t)__popc(h & lanemask_lt());
                if (((h >> lane) & 1u) && rank < n_enq) {
                    pend_k[wid][n_pend + rank] = base + 32 * u + lane;
                    pend_slot[wid][n_pend + rank] = count + rank;
                }
                n_pend += n_enq;
                count += nh;
            }
        }
    };
    uint32_t base = 1;
    for (; base + 32 * EV_UNROLL - 1 <= pv.m; base += 32 * EV_UNROLL) body(IC<1>{}, base);
    for (; base <= K; base += 32 * EV_UNROLL) body(IC<0>{}, base);
    if (n_pend) flush();
    if (lane == 0) a.n_events[i] = count;
}

// ---- events on WIDE records (1 + D > 8; linear32's 264-byte records): the stream staged through shared memory by TMA.
// What limited the one-lane-per-record kernel above on such paths (config 4: ~122 records, 2.3 TB/s = 36 % of the HBM
// peak; ncu: 67 % of the stall samples long_scoreboard, 0.27 eligible warps per scheduler, 3.1 of 4 warps resident at
// 128 registers; profiles/r02z_cfg4_events_full.md) is memory-level parallelism: a lane that sums its own 264-byte
// record from global memory touches a new 32-byte sector every four terms, and with the registers of a wide kernel only a
// few of those loads are in flight at once — nine dependent round trips per chunk, ~33 us per path per warp.  Here ONE
// elected lane asks the TMA unit for whole chunks (32 records = 8448 contiguous bytes, cp.async.bulk into a per-warp ring of
// EVW_STAGES buffers, completion on an mbarrier), up to EVW_STAGES chunks ahead, so 25 KB per warp are in flight without
// a single register; the lanes then read their record from shared memory (row stride 33 doubles: conflict-free) and
// sum g in the oracle's order.  Crossing test and queue are the kernel's above.  Location: a Locate policy with
// DEFERRED leaves only the knot index of every crossing in its event slot and a second kernel (path_locate_deferred_kernel)
// locates them all at once — for linear32 a location needs the trajectory's 8 KB matrix and four
// 32 x 32 matrix-vector products by the whole warp, and done inside the stream, one crossing after the other at the
// end of each path, it held the streaming kernel at 3.2 ms whatever its staging depth or occupancy (config 4: 0.9
// crossings per path; A/B in profiles/r03_path_queries.md).
#ifndef BACON_EVW_STAGES
#define BACON_EVW_STAGES 2
#endif
#ifndef BACON_EVW_MINB
#define BACON_EVW_MINB 3
#endif
constexpr int EVW_STAGES = BACON_EVW_STAGES;
template <int R> __host__ __device__ constexpr size_t evw_smem_bytes() { return (size_t)(PATH_BLOCK / 32) * EVW_STAGES * 32 * R * sizeof(double); }

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    unsigned done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    }
}
// global -> shared bulk copy by the TMA unit: 16-byte aligned on both sides, a multiple of 16 bytes
__device__ __forceinline__ void tma_bulk_load(void* dst_smem, const void* src_gmem, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

template <class Rhs, bool STRICT, class Locate>
__global__ void __launch_bounds__(PATH_BLOCK, BACON_EVW_MINB) path_events_wide_kernel(const __grid_constant__ bacon_path_args a) {
    constexpr int D = Rhs::DIM;
    constexpr int R = 1 + D;
    constexpr int NW = PATH_BLOCK / 32;
    extern __shared__ __align__(128) double evw_ring[];  // [NW][EVW_STAGES][32 * R]
    __shared__ __align__(8) unsigned long long bars[NW][EVW_STAGES];
    __shared__ uint32_t pend_k[NW][32], pend_slot[NW][32];
    const unsigned wid = threadIdx.x >> 5, lane = lane_id();
    double* ring = evw_ring + (size_t)wid * EVW_STAGES * 32 * R;
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < EVW_STAGES; ++s) mbar_init(&bars[wid][s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");  // (the barriers become visible to the async proxy)
    }
    __syncwarp();
    const unsigned long long i = ((unsigned long long)blockIdx.x * PATH_BLOCK + threadIdx.x) >> 5;
    if (i >= a.n) return;  // (whole warps leave together)
    const PathView<D> pv(a, i);
    const uint32_t K = pv.last();
    const uint32_t n_chunks = (pv.m + 31u) >> 5;
    auto issue = [&](uint32_t c) {  // lane 0: chunk c on its way into stage c % EVW_STAGES
        const uint32_t st = c % EVW_STAGES;
        const uint32_t left = pv.m - 32u * c, nrec = left < 32u ? left : 32u;
        const uint32_t bytes = (nrec & ~1u) * R * 8u;  // an even number of records is a multiple of 16 bytes
        if (bytes) {
            mbar_expect_tx(&bars[wid][st], bytes);
            tma_bulk_load(ring + (size_t)st * 32 * R, pv.rec + (size_t)32 * c * R, bytes, &bars[wid][st]);
        } else {
            mbar_arrive(&bars[wid][st]);  // (a last chunk of one record: nothing for the TMA unit)
        }
    };
    if (lane == 0) {
        for (uint32_t c = 0; c < n_chunks && c < (uint32_t)EVW_STAGES; ++c) issue(c);
    }
    // g of a knot that is not a record (knot 0, the closing knot): straight from memory, oracle order
    auto knot_g = [&](uint32_t k) -> double {
        const double* src = k == 0 ? pv.y0 + i : pv.y_end + i;
        double sum = a.ev_w[0] * src[0];
#pragma unroll
        for (int d = 1; d < D; ++d) sum += a.ev_w[d] * src[(size_t)d * pv.n];
        return sum - a.ev_c;
    };
    const double g0 = knot_g(0);
    unsigned carry_neg = g0 < 0.0 ? 1u : 0u, carry_pos = g0 > 0.0 ? 1u : 0u;
    uint32_t count = 0, n_pend = 0;  // warp-uniform
    const uint32_t cap = (uint32_t)a.ev_capacity;
    double* ev = a.events + (size_t)i * cap * R;
    auto flush = [&]() {
        // (a DEFERRED policy never queues anything here, and its flush() must not add its shared memory to this kernel)
        if constexpr (!Locate::DEFERRED) {
            __syncwarp();
            Locate::flush(a, i, n_pend, pend_k[wid], pend_slot[wid], ev, lane);
            __syncwarp();
        }
    };
    // 32 knots base .. base + 31 (lane l holds knot base + l, `have` = it exists): crossings against the left neighbour
    auto scan = [&](uint32_t base, bool have, double g) {
        const unsigned neg = __ballot_sync(FULL_MASK, have && g < 0.0), pos = __ballot_sync(FULL_MASK, have && g > 0.0);
        const unsigned ord = __ballot_sync(FULL_MASK, have && g == g);  // (a NaN is neither side of the surface)
        const unsigned rising = ((neg << 1) | carry_neg) & ~neg & ord, falling = ((pos << 1) | carry_pos) & ~pos & ord;
        carry_neg = neg >> 31;
        carry_pos = pos >> 31;
        const unsigned h = a.ev_direction > 0 ? rising : (a.ev_direction < 0 ? falling : (rising | falling));
        if (h == 0) return;
        const uint32_t nh = (uint32_t)__popc(h);
        const uint32_t room = count < cap ? cap - count : 0u;
        const uint32_t n_enq = nh < room ? nh : room;
        const uint32_t rank = (uint32_t)__popc(h & lanemask_lt());
        if constexpr (Locate::DEFERRED) {  // the slot holds the knot index until the second kernel has been there
            if (((h >> lane) & 1u) && rank < n_enq) {  // (and, next to it, the number of records: flush_known)
                ev[(size_t)(count + rank) * R] = __longlong_as_double((long long)(base + lane));
                ev[(size_t)(count + rank) * R + 1] = __longlong_as_double((long long)pv.m);
            }
            count += nh;
            return;
        }
        if (n_pend + n_enq > 32u) {
            flush();
            n_pend = 0;
        }
        if (((h >> lane) & 1u) && rank < n_enq) {
            pend_k[wid][n_pend + rank] = base + lane;
            pend_slot[wid][n_pend + rank] = count + rank;
        }
        n_pend += n_enq;
        count += nh;
    };
    for (uint32_t c = 0; c < n_chunks; ++c) {
        const uint32_t st = c % EVW_STAGES;
        double* buf = ring + (size_t)st * 32 * R;
        const uint32_t left = pv.m - 32u * c, nrec = left < 32u ? left : 32u;
        mbar_wait(&bars[wid][st], (c / EVW_STAGES) & 1u);
        if (nrec & 1u) {  // the odd last record of the path: by the lanes
            const double* src = pv.rec + ((size_t)32 * c + (nrec - 1)) * R;
            for (uint32_t e = lane; e < (uint32_t)R; e += 32) buf[(size_t)(nrec - 1) * R + e] = src[e];
        }
        __syncwarp();
        const bool have = lane < nrec;
        double g = 0.0;
        if (have) {
            const double* r = buf + (size_t)lane * R + 1;
            double sum = a.ev_w[0] * r[0];
#pragma unroll
            for (int d = 1; d < D; ++d) sum += a.ev_w[d] * r[d];
            g = sum - a.ev_c;
        }
        scan(1u + 32u * c, have, g);
        __syncwarp();  // every lane is done with this stage: the next chunk may land in it
        if (lane == 0 && c + EVW_STAGES < n_chunks) issue(c + EVW_STAGES);
    }
    if (K > pv.m) scan(K, lane == 0, knot_g(K));  // the closing knot
    if (n_pend) flush();
    if (lane == 0) a.n_events[i] = count;
}

// second kernel of a DEFERRED location.  A warp takes 32 consecutive event slots (slot = trajectory * capacity + j):
// lane l reads whether its slot is filled (j < min(n_events, capacity)) and the knot index and record count parked
// there — coalesced, one round trip for 32 slots — then the warp locates the filled ones trajectory by trajectory through
// the policy's flush_known(), so a trajectory's matrix is loaded once for all its crossings.
#ifndef BACON_LOCATE_MINB
#define BACON_LOCATE_MINB 4  // (measured on config 4: 3 -> 2.42 ms, 4 -> 2.32, 5 -> 2.33)
#endif
template <class Locate>
__global__ void __launch_bounds__(PATH_BLOCK, BACON_LOCATE_MINB) path_locate_deferred_kernel(const __grid_constant__ bacon_path_args a) {
    constexpr int R = 1 + Locate::DIM;
    __shared__ uint32_t pk[PATH_BLOCK / 32][32], ps[PATH_BLOCK / 32][32], pm[PATH_BLOCK / 32][32];
    const unsigned wid = threadIdx.x >> 5, lane = lane_id();
    const unsigned long long cap = (unsigned long long)a.ev_capacity, total = a.n * cap;
    const unsigned long long s = ((((unsigned long long)blockIdx.x * PATH_BLOCK + threadIdx.x) >> 5) << 5) + lane;
    bool valid = false;
    uint32_t k = 0, j = 0, m = 0;
    unsigned long long i = 0;
    if (s < total) {
        i = s / cap;
        j = (uint32_t)(s - i * cap);
        const uint32_t found = a.n_events[i];
        valid = j < (found < (uint32_t)cap ? found : (uint32_t)cap);
        if (valid) {
            k = (uint32_t)__double_as_longlong(a.events[(size_t)s * R]);
            m = (uint32_t)__double_as_longlong(a.events[(size_t)s * R + 1]);
        }
    }
    unsigned mask = __ballot_sync(FULL_MASK, valid);
    while (mask) {
        const int lead = __ffs(mask) - 1;
        const unsigned long long i_cur = __shfl_sync(FULL_MASK, i, lead);
        const bool mine = valid && i == i_cur;
        const unsigned same = __ballot_sync(FULL_MASK, mine);
        if (mine) {
            const unsigned rank = __popc(same & lanemask_lt());
            pk[wid][rank] = k;
            ps[wid][rank] = j;
            pm[wid][rank] = m;
        }
        __syncwarp();
        Locate::flush_known(a, i_cur, (uint32_t)__popc(same), pk[wid], ps[wid], pm[wid], a.events + (size_t)i_cur * cap * R, lane);
        __syncwarp();
        mask &= ~same;
    }
}

#ifndef __CUDACC_RTC__  // (runtime-compiled functors are launched through the driver API: rtc.cu)
// host-side launcher of both queries for one right-hand side and one build flavour
// the events kernel for this record width: wide records (1 + D > 8) stream through the TMA-staged kernel when every
// path starts on a 16-byte boundary (an even capacity; hist itself is at least 32-byte aligned)
template <class Rhs, bool STRICT, class Locate>
int launch_path_events(bacon_path_args* a, unsigned blocks, cudaStream_t st, cudaFuncAttributes* fa) {
    constexpr int R = 1 + Rhs::DIM;
    const bool no_tma = getenv("BACON_EV_NO_TMA") != nullptr;  // (A/B switch for measurements and tests; read per call)
    if constexpr (R > 8) {
        const bool aligned = ((size_t)a->cfg.history_capacity * R) % 2 == 0 && (reinterpret_cast<uintptr_t>(a->hist) & 15u) == 0;
        if (aligned && !no_tma) {
            auto kernel = path_events_wide_kernel<Rhs, STRICT, Locate>;
            constexpr size_t smem = evw_smem_bytes<R>();
            if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return BACON_E_CUDA;
            if (cudaFuncGetAttributes(fa, kernel) != cudaSuccess) return BACON_E_CUDA;
            kernel<<<blocks, PATH_BLOCK, smem, st>>>(*a);
            if constexpr (Locate::DEFERRED) {
                const unsigned long long slots = (unsigned long long)a->n * (unsigned long long)a->ev_capacity;
                const unsigned long long b2 = (slots + PATH_BLOCK - 1) / PATH_BLOCK;  // (a warp per 32 slots)
                if (b2 > 0x7fffffffull) return BACON_E_BAD_ARGUMENT;
                if (b2 > 0) path_locate_deferred_kernel<Locate><<<(unsigned)b2, PATH_BLOCK, 0, st>>>(*a);
            }
            return 0;
        }
    }
    auto kernel = path_events_kernel<Rhs, STRICT, Locate>;
    if (cudaFuncGetAttributes(fa, kernel) != cudaSuccess) return BACON_E_CUDA;
    kernel<<<blocks, PATH_BLOCK, 0, st>>>(*a);
    return 0;
}

template <class Rhs, bool STRICT> int launch_path_query(bacon_path_args* a) {
    static_assert(Rhs::DIM <= BACON_PATH_MAX_DIM, "event weights are passed by value");
    cudaStream_t st = (cudaStream_t)a->stream;
    cudaFuncAttributes fa;
    unsigned long long blocks = 0;
    if (a->op == BACON_PATH_SAMPLE) {
        auto kernel = path_sample_kernel<Rhs, STRICT>;
        if (cudaFuncGetAttributes(&fa, kernel) != cudaSuccess) return BACON_E_CUDA;
        blocks = (a->n * a->n_times + PATH_BLOCK - 1) / PATH_BLOCK;
        if (blocks == 0 || blocks > 0x7fffffffull) return BACON_E_BAD_ARGUMENT;
        kernel<<<(unsigned)blocks, PATH_BLOCK, 0, st>>>(*a);
    } else if (a->op == BACON_PATH_EVENTS) {
        blocks = (a->n * 32 + PATH_BLOCK - 1) / PATH_BLOCK;
        if (blocks == 0 || blocks > 0x7fffffffull) return BACON_E_BAD_ARGUMENT;
        if (const int rc = launch_path_events<Rhs, STRICT, LaneLocate<Rhs>>(a, (unsigned)blocks, st, &fa)) return rc;
    } else {
        return BACON_E_BAD_ARGUMENT;
    }
    if (cudaGetLastError() != cudaSuccess) return BACON_E_CUDA;
    a->grid = (int)blocks;
    a->block = PATH_BLOCK;
    a->regs_per_thread = fa.numRegs;
    return 0;
}
#endif

}  // namespace bacon
