// path_query_warp.cuh — the path queries (path_query.cuh) for the wide-state right-hand side y' = A y, D = 32
// (`linear32`, BASELINE config 4: the configuration whose dense output the queries read).  As in rk_warp_linear.cuh a
// state does not fit one thread, so the interpolation is done by a WARP: lane d owns component d of both knots and of
// both slopes, and row d of the trajectory's matrix; f = A y takes the other components by shuffle, in the oracle's
// order (s = A[d][0] y[0]; s += A[d][k] y[k]), so the strict build stays bit-comparable with oracle/oracle_capi.cpp.
//   path_sample_warp32_kernel   one warp per (trajectory, up to 32 sample times): one bisection per lane, then the warp interpolates;
//                               both knots come in as coalesced 256-byte rows, the sample leaves as one.
//   events                      the streaming kernel is path_query.cuh's (one lane per record, g accumulated over the
//                               record's 32 components); queued crossings are located by the whole warp, one at a time.
#pragma once
#include "path_query.cuh"
#include "rhs_builtin.cuh"

namespace bacon {

// row `lane` of trajectory i's matrix, any parameter layout
__device__ __forceinline__ void load_matrix_row32(const bacon_path_args& a, unsigned long long i, unsigned lane, double (&A)[32]) {
    constexpr int N = 32;
    const double* P = a.params;
    if (a.cfg.flags & BACON_FLAG_SHARED_PARAMS) {
#pragma unroll
        for (int j = 0; j < N; ++j) A[j] = P[lane * N + j];
    } else if (a.cfg.flags & BACON_FLAG_PARAMS_AOS) {
        const double2* row = reinterpret_cast<const double2*>(P + (size_t)i * (N * N) + (size_t)lane * N);
#pragma unroll
        for (int j2 = 0; j2 < N / 2; ++j2) {
            const double2 v = row[j2];
            A[2 * j2] = v.x;
            A[2 * j2 + 1] = v.y;
        }
    } else {
#pragma unroll
        for (int j = 0; j < N; ++j) A[j] = P[(size_t)(lane * N + j) * a.n + i];
    }
}

// (A y)[lane], y spread over the lanes; sequential in k like RhsLinear::operator()
__device__ __forceinline__ double warp_matvec32(const double (&A)[32], double y_lane) {
    double s = A[0] * __shfl_sync(FULL_MASK, y_lane, 0);
#pragma unroll
    for (int k = 1; k < 32; ++k) s += A[k] * __shfl_sync(FULL_MASK, y_lane, k);
    return s;
}

// The same sums for both knots of an interval at once: the two chains are independent, so their shuffles and FMAs
// interleave and the 32 dependent steps are walked once instead of twice (each sum keeps its own order, so its bits).
__device__ __forceinline__ void warp_matvec32_pair(const double (&A)[32], double ya, double yb, double& fa, double& fb) {
    double sa = A[0] * __shfl_sync(FULL_MASK, ya, 0), sb = A[0] * __shfl_sync(FULL_MASK, yb, 0);
#pragma unroll
    for (int k = 1; k < 32; ++k) {
        sa += A[k] * __shfl_sync(FULL_MASK, ya, k);
        sb += A[k] * __shfl_sync(FULL_MASK, yb, k);
    }
    fa = sa;
    fb = sb;
}
// f = A y and w . y for both knots: one broadcast of y_k feeds both sums
__device__ __forceinline__ void warp_matvec_dot32_pair(const bacon_path_args& a, const double (&A)[32], double ya, double yb,
                                                       double& fa, double& fb, double& wa, double& wb) {
    double ya_k = __shfl_sync(FULL_MASK, ya, 0), yb_k = __shfl_sync(FULL_MASK, yb, 0);
    double sa = A[0] * ya_k, sb = A[0] * yb_k, da = a.ev_w[0] * ya_k, db = a.ev_w[0] * yb_k;
#pragma unroll
    for (int k = 1; k < 32; ++k) {
        ya_k = __shfl_sync(FULL_MASK, ya, k);
        yb_k = __shfl_sync(FULL_MASK, yb, k);
        sa += A[k] * ya_k;
        sb += A[k] * yb_k;
        da += a.ev_w[k] * ya_k;
        db += a.ev_w[k] * yb_k;
    }
    fa = sa;
    fb = sb;
    wa = da;
    wb = db;
}
__device__ __forceinline__ void warp_seqdot32_pair(const bacon_path_args& a, double va, double vb, double& wa, double& wb) {
    double da = a.ev_w[0] * __shfl_sync(FULL_MASK, va, 0), db = a.ev_w[0] * __shfl_sync(FULL_MASK, vb, 0);
#pragma unroll
    for (int d = 1; d < 32; ++d) {
        da += a.ev_w[d] * __shfl_sync(FULL_MASK, va, d);
        db += a.ev_w[d] * __shfl_sync(FULL_MASK, vb, d);
    }
    wa = da;
    wb = db;
}

// w . v over the lanes, sequential in d like event_fn; every lane gets the sum
__device__ __forceinline__ double warp_seqdot32(const bacon_path_args& a, double v_lane) {
    double s = a.ev_w[0] * __shfl_sync(FULL_MASK, v_lane, 0);
#pragma unroll
    for (int d = 1; d < 32; ++d) s += a.ev_w[d] * __shfl_sync(FULL_MASK, v_lane, d);
    return s;
}

// component d of knot k
__device__ __forceinline__ double knot_component32(const PathView<32>& pv, uint32_t k, unsigned d) {
    if (k == 0) return pv.y0[(size_t)d * pv.n + pv.i];
    if (k <= pv.m) return pv.rec[(size_t)(k - 1) * 33 + 1 + d];
    return pv.y_end[(size_t)d * pv.n + pv.i];
}

// Sample times one warp takes of its trajectory.  The 8 KB matrix (row d in lane d's registers) is loaded once per warp
// and serves them all, and the bisections of all its times run side by side, one per lane (one warp per sample re-read
// the matrix per sample and searched with all 32 lanes in lockstep: 14.1 ms for 2^18 x 16 samples; 8 times per warp,
// searched one after the other: 7.6 ms; profiles/r01o_path_queries.md).
constexpr int WARP32_TIMES = 32;

template <bool STRICT>
__global__ void __launch_bounds__(PATH_BLOCK) path_sample_warp32_kernel(const __grid_constant__ bacon_path_args a) {
    const unsigned long long chunks = (a.n_times + WARP32_TIMES - 1) / WARP32_TIMES;
    const unsigned long long g = ((unsigned long long)blockIdx.x * PATH_BLOCK + threadIdx.x) >> 5;
    if (g >= a.n * chunks) return;  // (whole warps)
    const unsigned lane = lane_id();
    const unsigned long long i = g / chunks, j0 = (g - i * chunks) * WARP32_TIMES;
    const unsigned n_here = (unsigned)(j0 + WARP32_TIMES < a.n_times ? WARP32_TIMES : a.n_times - j0);
    const PathView<32> pv(a, i);
    const uint32_t K = pv.last();
    const double t_last = pv.time(K);
    // phase 1: lane l finds the interval of time j0 + l (0 = outside the path, ~0 = exactly t_start on an empty path)
    double my_tau = 0.0;
    uint32_t my_lo = 0;
    if (lane < n_here) {
        my_tau = a.times[j0 + lane];
        if (K > 0 && my_tau >= pv.t0 && my_tau <= t_last) {
            my_lo = pv.first_knot_at_or_after(my_tau, K, t_last);
        } else if (my_tau == pv.t0) {
            my_lo = ~0u;
        }
    }
    double A[32];
    load_matrix_row32(a, i, lane, A);
    // phase 2: the warp interpolates them one after the other.  (Requesting the knots of sample l + 1 before sample l is
    // computed was measured slower: 3.90 -> 4.85 ms on config 4, the extra live rows cost more than the overlap gains.)
#pragma unroll 2
    for (unsigned l = 0; l < n_here; ++l) {
        const double tau = __shfl_sync(FULL_MASK, my_tau, l);
        const uint32_t lo = __shfl_sync(FULL_MASK, my_lo, l);
        double* out = a.samples + ((size_t)i * a.n_times + j0 + l) * 32;
        if (lo == 0 || lo == ~0u) {
            out[lane] = lo ? knot_component32(pv, 0, lane) : path_nan();
            continue;
        }
        const double ta = pv.time(lo - 1), tb = pv.time(lo);
        const double ya[1] = {knot_component32(pv, lo - 1, lane)}, yb[1] = {knot_component32(pv, lo, lane)};
        double fa[1], fb[1];
        warp_matvec32_pair(A, ya[0], yb[0], fa[0], fb[0]);
        const double h = tb - ta;
        const double th = h > 0.0 ? (tau - ta) / h : 0.0;
        double res[1];
        hermite_eval<1>(th, h, ya, yb, fa, fb, res);
        out[lane] = res[0];
    }
}

// the queue of crossings of one trajectory, located by the whole warp one after the other
template <bool STRICT> struct WarpLocateLinear32 {
    static constexpr int DIM = 32;
    // with the TMA-staged streaming kernel (path_query.cuh: path_events_wide_kernel) the crossings are located by a
    // second kernel, one warp per crossing, instead of one after the other at the end of each path
    static constexpr bool DEFERRED = true;
    static __device__ __noinline__ void flush(const bacon_path_args& a, unsigned long long i, uint32_t n_pend, const uint32_t* pk,
                                              const uint32_t* ps, double* ev, unsigned lane) {
        const PathView<32> pv(a, i);
        double A[32];
        load_matrix_row32(a, i, lane, A);
#pragma unroll 1
        for (uint32_t e = 0; e < n_pend; ++e) {
            const uint32_t k = pk[e];
            double* dst = ev + (size_t)ps[e] * 33;
            const double ta = pv.time(k - 1), tb = pv.time(k);
            const double ya[1] = {knot_component32(pv, k - 1, lane)}, yb[1] = {knot_component32(pv, k, lane)};
            double fa[1], fb[1], wa, wb, da, db;
            warp_matvec_dot32_pair(a, A, ya[0], yb[0], fa[0], fb[0], wa, wb);
            warp_seqdot32_pair(a, fa[0], fb[0], da, db);
            const double h = tb - ta;
            const double ga = wa - a.ev_c, gb = wb - a.ev_c;
            const double th = hermite_root(ga, gb, h * da, h * db);
            double ys[1];
            hermite_eval<1>(th, h, ya, yb, fa, fb, ys);
            if (lane == 0) dst[0] = ta + th * h;
            dst[1 + lane] = ys[0];
        }
    }
};

template <bool STRICT> int launch_path_query_linear32(bacon_path_args* a) {
    cudaStream_t st = (cudaStream_t)a->stream;
    cudaFuncAttributes fa;
    unsigned long long blocks = 0;
    if (a->op == BACON_PATH_SAMPLE) {
        auto kernel = path_sample_warp32_kernel<STRICT>;
        if (cudaFuncGetAttributes(&fa, kernel) != cudaSuccess) return BACON_E_CUDA;
        blocks = (a->n * ((a->n_times + WARP32_TIMES - 1) / WARP32_TIMES) * 32 + PATH_BLOCK - 1) / PATH_BLOCK;
        if (blocks == 0 || blocks > 0x7fffffffull) return BACON_E_BAD_ARGUMENT;
        kernel<<<(unsigned)blocks, PATH_BLOCK, 0, st>>>(*a);
    } else if (a->op == BACON_PATH_EVENTS) {
        blocks = (a->n * 32 + PATH_BLOCK - 1) / PATH_BLOCK;
        if (blocks == 0 || blocks > 0x7fffffffull) return BACON_E_BAD_ARGUMENT;
        if (const int rc = launch_path_events<RhsLinear<32>, STRICT, WarpLocateLinear32<STRICT>>(a, (unsigned)blocks, st, &fa)) return rc;
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
