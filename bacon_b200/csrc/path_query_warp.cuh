// path_query_warp.cuh — the path queries (path_query.cuh) for the wide-state right-hand side y' = A y, D = 32
// (`linear32`, BASELINE config 4: the configuration whose dense output the queries read).  As in rk_warp_linear.cuh a
// state does not fit one thread, so the interpolation is done by a WARP: lane d owns component d of both knots and of
// both slopes, and row d of the trajectory's matrix; f = A y takes the other components through shared memory, in the oracle's
// order (s = A[d][0] y[0]; s += A[d][k] y[k]), so the strict build stays bit-comparable with oracle/oracle_capi.cpp.
//   path_sample_warp32_kernel   one warp per (trajectory, up to 32 sample times): one knot search per lane, then the warp interpolates;
//                               the knots of 16 samples come in at once as asynchronous copies of coalesced 256-byte
//                               rows, the sample leaves as one.
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

// The other lanes' components reach a lane through shared memory, two vectors at a time (both knots of an interval):
// every lane publishes its component of each (one STS.64 per vector), then reads all of them back with broadcast
// LDS.128.  Round 2's form took them by shuffle; both go through the L1/shared data pipe, where a 64-bit shuffle costs
// two wavefronts per double and a broadcast LDS.128 two per PAIR of doubles (tools/lds_pattern_probe.cu) — 128 against
// 68 wavefronts for the two products of a sample, in kernels that issue little else (profiles/r04_linear32.md).  The
// sums keep the oracle's order (s = A[d][0] y[0]; s += A[d][k] y[k]), so the bits are those of the shuffle form.
struct WarpPairBuf {
    double v[2][32];
};
__device__ __forceinline__ const double2* warp_pair_publish(WarpPairBuf& buf, unsigned lane, double va, double vb) {
    __syncwarp();  // (the previous pair has been read by every lane)
    buf.v[0][lane] = va;
    buf.v[1][lane] = vb;
    __syncwarp();
    return reinterpret_cast<const double2*>(&buf.v[0][0]);
}

// (A ya)[lane] and (A yb)[lane] with both vectors in shared memory (16-byte aligned): the two chains are independent,
// so their FMAs interleave and the 32 dependent steps are walked once instead of twice (each sum keeps its own order,
// so its bits)
__device__ __forceinline__ void warp_matvec32_pair_at(const double2* va, const double2* vb, const double (&A)[32], double& fa, double& fb) {
    double sa = 0.0, sb = 0.0;
#pragma unroll
    for (int k2 = 0; k2 < 16; ++k2) {
        const double2 pa = va[k2], pb = vb[k2];
        if (k2 == 0) {
            sa = A[0] * pa.x;
            sb = A[0] * pb.x;
        } else {
            sa += A[2 * k2] * pa.x;
            sb += A[2 * k2] * pb.x;
        }
        sa += A[2 * k2 + 1] * pa.y;
        sb += A[2 * k2 + 1] * pb.y;
    }
    fa = sa;
    fb = sb;
}
__device__ __forceinline__ void warp_matvec32_pair(WarpPairBuf& buf, unsigned lane, const double (&A)[32], double ya, double yb,
                                                   double& fa, double& fb) {
    const double2* v = warp_pair_publish(buf, lane, ya, yb);
    warp_matvec32_pair_at(v, v + 16, A, fa, fb);
}

// Asynchronous copies into shared memory (LDGSTS): no register holds the data in flight, so a warp can have all the
// knots of its samples — or its whole matrix — on the way at once.
__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async8(void* dst_smem, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// One warp's staging area: the 32 x 32 matrix with rows 34 doubles apart, or 32 state vectors of 32 doubles followed by
// their 32 knot times.
struct WarpStage {
    double v[32 * 34];
};
// Row `lane` of a row-major (AoS) matrix through the staging area: the warp moves the 8 KB as 16 coalesced 512-byte
// requests (a row per lane straight from global memory touches 32 lines per request: 512 wavefronts of the L1 data
// pipe against 192 this way) and every lane reads its own row back with LDS.128 (272 bytes apart: conflict-free).
// begin() only issues the copies; what the warp does before end() overlaps them.
__device__ __forceinline__ bool stage_matrix32_begin(const bacon_path_args& a, unsigned long long i, unsigned lane, WarpStage& st) {
    if ((a.cfg.flags & BACON_FLAG_SHARED_PARAMS) || !(a.cfg.flags & BACON_FLAG_PARAMS_AOS)) return false;
    const double* M = a.params + (size_t)i * 1024;
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 16; ++k) cp_async16(&st.v[(2 * k + (lane >> 4)) * 34 + 2 * (lane & 15)], M + 64 * k + 2 * lane);
    return true;
}
__device__ __forceinline__ void stage_matrix32_end(const bacon_path_args& a, unsigned long long i, unsigned lane, bool staged,
                                                   WarpStage& st, double (&A)[32]) {
    if (!staged) {
        load_matrix_row32(a, i, lane, A);
        return;
    }
    cp_async_wait_all();
    __syncwarp();
    const double2* row = reinterpret_cast<const double2*>(&st.v[lane * 34]);
#pragma unroll
    for (int j2 = 0; j2 < 16; ++j2) {
        const double2 v = row[j2];
        A[2 * j2] = v.x;
        A[2 * j2 + 1] = v.y;
    }
    __syncwarp();  // (the area is free again)
}

// f = A y and w . y for both knots: one broadcast of y_k feeds both sums
__device__ __forceinline__ void warp_matvec_dot32_pair(WarpPairBuf& buf, unsigned lane, const bacon_path_args& a, const double (&A)[32],
                                                       double ya, double yb, double& fa, double& fb, double& wa, double& wb) {
    const double2* v = warp_pair_publish(buf, lane, ya, yb);
    double sa = 0.0, sb = 0.0, da = 0.0, db = 0.0;
#pragma unroll
    for (int k2 = 0; k2 < 16; ++k2) {
        const double2 pa = v[k2], pb = v[16 + k2];
        if (k2 == 0) {
            sa = A[0] * pa.x;
            sb = A[0] * pb.x;
            da = a.ev_w[0] * pa.x;
            db = a.ev_w[0] * pb.x;
        } else {
            sa += A[2 * k2] * pa.x;
            sb += A[2 * k2] * pb.x;
            da += a.ev_w[2 * k2] * pa.x;
            db += a.ev_w[2 * k2] * pb.x;
        }
        sa += A[2 * k2 + 1] * pa.y;
        sb += A[2 * k2 + 1] * pb.y;
        da += a.ev_w[2 * k2 + 1] * pa.y;
        db += a.ev_w[2 * k2 + 1] * pb.y;
    }
    fa = sa;
    fb = sb;
    wa = da;
    wb = db;
}
// w . va and w . vb over the lanes, sequential in d like event_fn; every lane gets both sums
__device__ __forceinline__ void warp_seqdot32_pair(WarpPairBuf& buf, unsigned lane, const bacon_path_args& a, double va, double vb,
                                                   double& wa, double& wb) {
    const double2* v = warp_pair_publish(buf, lane, va, vb);
    double da = 0.0, db = 0.0;
#pragma unroll
    for (int d2 = 0; d2 < 16; ++d2) {
        const double2 pa = v[d2], pb = v[16 + d2];
        if (d2 == 0) {
            da = a.ev_w[0] * pa.x;
            db = a.ev_w[0] * pb.x;
        } else {
            da += a.ev_w[2 * d2] * pa.x;
            db += a.ev_w[2 * d2] * pb.x;
        }
        da += a.ev_w[2 * d2 + 1] * pa.y;
        db += a.ev_w[2 * d2 + 1] * pb.y;
    }
    wa = da;
    wb = db;
}

// component d of knot k
__device__ __forceinline__ double knot_component32(const PathView<32>& pv, uint32_t k, unsigned d) {
    if (k == 0) return pv.y0[(size_t)d * pv.n + pv.i];
    if (k <= pv.m) return pv.rec[(size_t)(k - 1) * 33 + 1 + d];
    return pv.y_end[(size_t)d * pv.n + pv.i];
}

// Sample times one warp takes of its trajectory.  The 8 KB matrix (row d in lane d's registers) is loaded once per warp
// and serves them all, and the knot searches of all its times run side by side, one per lane (one warp per sample re-read
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
    // the matrix is on its way while the lanes search (phase 1)
    __shared__ __align__(16) WarpStage s_stage[PATH_BLOCK / 32];
    WarpStage& st = s_stage[threadIdx.x >> 5];
    const bool staged = stage_matrix32_begin(a, i, lane, st);
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
    stage_matrix32_end(a, i, lane, staged, st, A);
    // phase 2: the knots of up to 16 samples are requested at once (asynchronous copies into the staging area: the
    // warp pays the memory latency once per group, not once per sample — doing the same through registers was measured
    // slower in round 1, 3.90 -> 4.85 ms: the extra live rows cost more than the overlap gained), then the warp
    // interpolates them one after the other, the products reading the knot vectors where they landed.
    double* const knots = st.v;          // vector q = 2 u + side of sample u: knots[32 q .. 32 q + 31]
    double* const knot_t = st.v + 1024;  // its time
    for (unsigned base = 0; base < n_here; base += 16) {
        const unsigned cnt = n_here - base < 16u ? n_here - base : 16u;
        __syncwarp();
        for (unsigned u = 0; u < cnt; ++u) {
            const uint32_t lo = __shfl_sync(FULL_MASK, my_lo, base + u);
            if (lo == 0 || lo == ~0u) continue;
#pragma unroll
            for (unsigned side = 0; side < 2; ++side) {
                const uint32_t k = lo - 1 + side, q = 2 * u + side;
                if (k >= 1 && k <= pv.m) {
                    const double* r = pv.rec + (size_t)(k - 1) * 33;
                    cp_async8(&knots[32 * q + lane], r + 1 + lane);
                    if (lane == 0) cp_async8(&knot_t[q], r);
                } else {  // the initial condition or the closing knot: other arrays
                    knots[32 * q + lane] = knot_component32(pv, k, lane);
                    if (lane == 0) knot_t[q] = pv.time(k);
                }
            }
        }
        cp_async_wait_all();
        __syncwarp();
        for (unsigned u = 0; u < cnt; ++u) {
            const unsigned l = base + u;
            const double tau = __shfl_sync(FULL_MASK, my_tau, l);
            const uint32_t lo = __shfl_sync(FULL_MASK, my_lo, l);
            double* out = a.samples + ((size_t)i * a.n_times + j0 + l) * 32;
            if (lo == 0 || lo == ~0u) {
                out[lane] = lo ? knot_component32(pv, 0, lane) : path_nan();
                continue;
            }
            const double* va = knots + 64 * u;
            const double ta = knot_t[2 * u], tb = knot_t[2 * u + 1];
            const double ya[1] = {va[lane]}, yb[1] = {va[32 + lane]};
            double fa[1], fb[1];
            warp_matvec32_pair_at(reinterpret_cast<const double2*>(va), reinterpret_cast<const double2*>(va + 32), A, fa[0], fb[0]);
            const double h = tb - ta;
            const double th = h > 0.0 ? (tau - ta) / h : 0.0;
            double res[1];
            hermite_eval<1>(th, h, ya, yb, fa, fb, res);
            out[lane] = res[0];
        }
    }
}

// the queue of crossings of one trajectory, located by the whole warp one after the other
template <bool STRICT> struct WarpLocateLinear32 {
    static constexpr int DIM = 32;
    // with the TMA-staged streaming kernel (path_query.cuh: path_events_wide_kernel) the crossings are located by a
    // second kernel, one warp per crossing, instead of one after the other at the end of each path
    static constexpr bool DEFERRED = true;
    static constexpr uint32_t PREFETCH = 4;  // crossings of a trajectory whose knots are requested together with its matrix
    // pm != nullptr: pm[e] = the number of records of the path (parked next to the knot index by the streaming kernel),
    // which says where knot k lives without a look at the trajectory's bookkeeping: the two knots of the first PREFETCH
    // crossings are then requested by asynchronous copies at the same time as the matrix, and the warp waits ONCE per
    // trajectory instead of once for the matrix and once per crossing (the location kernel is a chain of dependent
    // round trips: 19 % of the warp slots resident, data pipe 60 % busy).
    template <bool KNOWN>
    static __device__ __forceinline__ void locate_all(const bacon_path_args& a, unsigned long long i, uint32_t n_pend, const uint32_t* pk,
                                                      const uint32_t* ps, const uint32_t* pm, double* ev, unsigned lane) {
        __shared__ __align__(16) WarpPairBuf s_pair[PATH_BLOCK / 32];
        __shared__ __align__(16) WarpStage s_stage[PATH_BLOCK / 32];
        __shared__ __align__(16) double s_knots[PATH_BLOCK / 32][2 * PREFETCH][32];
        __shared__ double s_knot_t[PATH_BLOCK / 32][2 * PREFETCH];
        const unsigned w = threadIdx.x >> 5;
        WarpPairBuf& pair = s_pair[w];
        WarpStage& st = s_stage[w];
        const bool staged = stage_matrix32_begin(a, i, lane, st);
        const double* rec = a.hist + (size_t)i * (size_t)a.cfg.history_capacity * 33;
        if constexpr (KNOWN) {
            for (uint32_t e = 0; e < n_pend && e < PREFETCH; ++e) {
                const uint32_t k = pk[e], m = pm[e];
#pragma unroll
                for (uint32_t side = 0; side < 2; ++side) {
                    const uint32_t kk = k - 1 + side;
                    if (kk >= 1 && kk <= m) {
                        const double* r = rec + (size_t)(kk - 1) * 33;
                        cp_async8(&s_knots[w][2 * e + side][lane], r + 1 + lane);
                        if (lane == 0) cp_async8(&s_knot_t[w][2 * e + side], r);
                    }
                }
            }
        }
        const PathView<32> pv(a, i);
        double A[32];
        stage_matrix32_end(a, i, lane, staged, st, A);  // (waits for every copy of this lane, then the warp meets)
        if (!staged) {
            cp_async_wait_all();
            __syncwarp();
        }
#pragma unroll 1
        for (uint32_t e = 0; e < n_pend; ++e) {
            const uint32_t k = pk[e];
            double* dst = ev + (size_t)ps[e] * 33;
            double t2[2], y2[2];
#pragma unroll
            for (uint32_t side = 0; side < 2; ++side) {
                const uint32_t kk = k - 1 + side;
                bool here = false;
                if constexpr (KNOWN) here = e < PREFETCH && kk >= 1 && kk <= pm[e];
                if (here) {
                    t2[side] = s_knot_t[w][2 * e + side];
                    y2[side] = s_knots[w][2 * e + side][lane];
                } else {
                    t2[side] = pv.time(kk);
                    y2[side] = knot_component32(pv, kk, lane);
                }
            }
            const double ta = t2[0], tb = t2[1];
            const double ya[1] = {y2[0]}, yb[1] = {y2[1]};
            double fa[1], fb[1], wa, wb, da, db;
            warp_matvec_dot32_pair(pair, lane, a, A, ya[0], yb[0], fa[0], fb[0], wa, wb);
            warp_seqdot32_pair(pair, lane, a, fa[0], fb[0], da, db);
            const double h = tb - ta;
            const double ga = wa - a.ev_c, gb = wb - a.ev_c;
            const double th = hermite_root(ga, gb, h * da, h * db);
            double ys[1];
            hermite_eval<1>(th, h, ya, yb, fa, fb, ys);
            if (lane == 0) dst[0] = ta + th * h;
            dst[1 + lane] = ys[0];
        }
        __syncwarp();  // (the staging areas are free again)
    }
    static __device__ __noinline__ void flush(const bacon_path_args& a, unsigned long long i, uint32_t n_pend, const uint32_t* pk,
                                              const uint32_t* ps, double* ev, unsigned lane) {
        locate_all<false>(a, i, n_pend, pk, ps, nullptr, ev, lane);
    }
    static __device__ __noinline__ void flush_known(const bacon_path_args& a, unsigned long long i, uint32_t n_pend, const uint32_t* pk,
                                                    const uint32_t* ps, const uint32_t* pm, double* ev, unsigned lane) {
        locate_all<true>(a, i, n_pend, pk, ps, pm, ev, lane);
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
