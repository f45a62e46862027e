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
//   path_sample_kernel   one thread per (trajectory, sample time): binary search over the trajectory's knots, two
//                        records in, D doubles out; neighbouring lanes take neighbouring times of ONE trajectory, so
//                        their probes share sectors and their stores are contiguous.
//   path_events_kernel   one warp per trajectory: the warp streams the path 32 records at a time (coalesced: a
//                        record is 8(1 + D) bytes, a D = 3 chunk is 1 KB), each lane tests g(y) = w . y - c for a sign
//                        change against its left neighbour, and the rare lane that finds one locates the root of the
//                        interpolant by bisection.  Events come out in path order (ballot + prefix count).
// Both are HBM-bound: the events kernel reads every record once (8(1 + D) bytes per accepted step), the sample kernel
// touches 2 records + D outputs per sample.  Operation order matches oracle/oracle_capi.cpp (oracle_sample_paths,
// oracle_locate_events): the strict build (-fmad=false) is bit-comparable with it.
#pragma once
#ifndef __CUDACC_RTC__
#include <cuda_runtime.h>
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
        t0 = a.cfg.t_start;
        closing = false;
        tc = 0.0;
        if (a.t_end && a.y_end) {
            tc = a.t_end[i];
            closing = tc > (m > 0 ? rec[(size_t)(m - 1) * R] : t0);
        }
    }
    __device__ __forceinline__ uint32_t last() const { return m + (closing ? 1u : 0u); }
    __device__ __forceinline__ double time(uint32_t k) const {
        return k == 0 ? t0 : (k <= m ? rec[(size_t)(k - 1) * R] : tc);
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
    if (K == 0 || !(tau >= pv.t0 && tau <= pv.time(K))) {
        if (tau == pv.t0) {
            pv.state(0, res);
        } else {
#pragma unroll
            for (int d = 0; d < D; ++d) res[d] = path_nan();
        }
    } else {
        uint32_t lo = 1, hi = K;  // the first knot at or after tau
        while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            if (pv.time(mid) >= tau) hi = mid;
            else lo = mid + 1;
        }
        const double ta = pv.time(lo - 1), tb = pv.time(lo);
        double ya[D], yb[D], fa[D], fb[D], p[P > 0 ? P : 1];
        pv.state(lo - 1, ya);
        pv.state(lo, yb);
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
    const double ta = pv.time(k - 1), tb = pv.time(k);
    double ya[D], yb[D], fa[D], fb[D], p[P > 0 ? P : 1];
    pv.state(k - 1, ya);
    pv.state(k, yb);
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
    double th = 1.0;
    if (gb[0] != 0.0) {
        double lo = 0.0, hi = 1.0;
        for (int it = 0; it < 80; ++it) {
            const double mid = 0.5 * (lo + hi);
            if (!(mid > lo && mid < hi)) break;
            double v[1];
            hermite_eval<1>(mid, h, ga, gb, da, db, v);
            if (v[0] == 0.0) {
                hi = mid;
                break;
            }
            if ((v[0] < 0.0) == (ga[0] < 0.0)) lo = mid;
            else hi = mid;
        }
        th = hi;
    }
    double ys[D];
    hermite_eval<D>(th, h, ya, yb, fa, fb, ys);
    dst[0] = ta + th * h;
#pragma unroll
    for (int d = 0; d < D; ++d) dst[1 + d] = ys[d];
}

template <class Rhs, bool STRICT>
__global__ void __launch_bounds__(PATH_BLOCK, 6) path_events_kernel(const __grid_constant__ bacon_path_args a) {
    constexpr int D = Rhs::DIM;
    const unsigned long long i = ((unsigned long long)blockIdx.x * PATH_BLOCK + threadIdx.x) >> 5;
    if (i >= a.n) return;  // (whole warps leave together)
    const unsigned lane = lane_id();
    const PathView<D> pv(a, i);
    const uint32_t K = pv.last();
    double y[D];
    pv.state(0, y);
    double g_carry = event_fn<D>(a, y);
    uint32_t count = 0;
    const uint32_t cap = (uint32_t)a.ev_capacity;
    double* ev = a.events + (size_t)i * cap * (1 + D);

    // software pipeline: the next chunk's record is in flight while this one is tested
    uint32_t k = 1 + lane;
    bool have = k <= K;
    if (have) pv.state(k, y);
    for (uint32_t base = 1; base <= K; base += 32) {
        const double gk = have ? event_fn<D>(a, y) : 0.0;
        const uint32_t k_now = k;
        const bool have_now = have;
        k += 32;
        have = k <= K;
        if (have) pv.state(k, y);
        double gprev = __shfl_up_sync(FULL_MASK, gk, 1);
        if (lane == 0) gprev = g_carry;
        g_carry = __shfl_sync(FULL_MASK, gk, 31);
        const bool hit = have_now && event_crossing(gprev, gk, a.ev_direction);
        const unsigned m = __ballot_sync(FULL_MASK, hit);
        if (hit) {
            const uint32_t slot = count + (uint32_t)__popc(m & lanemask_lt());
            if (slot < cap) locate_event<Rhs>(a, i, k_now, ev + (size_t)slot * (1 + D));
        }
        count += (uint32_t)__popc(m);
    }
    if (lane == 0) a.n_events[i] = count;
}

// host-side launcher of both queries for one right-hand side and one build flavour
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
        auto kernel = path_events_kernel<Rhs, STRICT>;
        if (cudaFuncGetAttributes(&fa, kernel) != cudaSuccess) return BACON_E_CUDA;
        blocks = (a->n * 32 + PATH_BLOCK - 1) / PATH_BLOCK;
        if (blocks == 0 || blocks > 0x7fffffffull) return BACON_E_BAD_ARGUMENT;
        kernel<<<(unsigned)blocks, PATH_BLOCK, 0, st>>>(*a);
    } else {
        return BACON_E_BAD_ARGUMENT;
    }
    if (cudaGetLastError() != cudaSuccess) return BACON_E_CUDA;
    a->grid = (int)blocks;
    a->block = PATH_BLOCK;
    a->regs_per_thread = fa.numRegs;
    return 0;
}

}  // namespace bacon
