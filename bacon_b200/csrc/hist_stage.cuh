// hist_stage.cuh — dense output (K4): every accepted (t, y) of every trajectory, i.e. what
// `IVPIterator::collect_vec` materialises (src/ivp.rs:209-211; the initial condition is not yielded by the
// RK/BDF/Adams steppers: rk.rs:418-419).
//
// HBM layout (bacon_ivp.h): hist[n][cap][1 + D] — one contiguous `Path` per trajectory, one (t, y) record
// per accepted point: the memory image of the reference's Vec<(f64, SVector<f64, D>)> (ivp.rs:203).
//
// Lanes of a warp sit at different trajectories and different step counts, so nothing a warp writes in one
// step is contiguous ACROSS lanes; what is contiguous is each lane's own record.  For D = 3 a record is
// 32 bytes = one DRAM sector, written by the lane with a single 256-bit store (STG.E.256, sm_100): every
// store instruction fills 32 whole sectors, no read-modify-write, no staging, one instruction per accepted
// step.  Other dimensions use the widest stores the record's alignment allows (128-bit when 1 + D is even,
// else 64-bit); consecutive records of a trajectory complete each other's sectors in L2 before eviction.
//
// History of this file: round 1 first staged 8 points per lane in shared memory (pair-interleaved layout)
// and copied full lane buffers out warp-cooperatively into separate hist_t[n][cap] / hist_y[n][cap][D]
// arrays.  Measured on B200 (bench_configs.py --config 2, profiles/): the copy-out loop serialised over
// full lanes (about 100 extra warp instructions per warp step against 195 for the integrator) and capped
// dense output at 1.2 TB/s; the record layout removes the loop altogether.
#pragma once
#include "ivp_common.cuh"

namespace bacon {

template <int D, bool ENABLED> struct HistStage;

template <int D> struct HistStage<D, false> {
    __device__ __forceinline__ explicit HistStage(const bacon_launch_args&) {}
    __device__ __forceinline__ void push(bool, uint32_t, unsigned long long, double, const double (&)[D]) {}
    __device__ __forceinline__ void retire(bool, unsigned long long, uint32_t) {}
};

template <int D> struct HistStage<D, true> {
    static constexpr int R = 1 + D;  // doubles per record

    double* hist;
    uint32_t* hist_len;
    uint32_t cap;

    __device__ __forceinline__ explicit HistStage(const bacon_launch_args& a) {
        hist = a.out.hist;
        hist_len = a.out.hist_len;
        cap = (uint32_t)a.cfg.history_capacity;
    }

    // called by every lane once per step() call; lanes that yielded a point write its record
    __device__ __forceinline__ void push(bool yielded, uint32_t n_acc_before, unsigned long long idx, double t,
                                         const double (&y)[D]) {
        if (!(yielded && n_acc_before < cap)) return;
        double* dst = hist + ((size_t)idx * cap + n_acc_before) * R;
        double rec[R];
        rec[0] = t;
#pragma unroll
        for (int d = 0; d < D; ++d) rec[1 + d] = y[d];
        if constexpr (R % 4 == 0) {  // 32-byte records: whole sectors
#pragma unroll
            for (int j = 0; j < R; j += 4)
                asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(dst + j), "d"(rec[j]), "d"(rec[j + 1]),
                             "d"(rec[j + 2]), "d"(rec[j + 3])
                             : "memory");
        } else if constexpr (R % 2 == 0) {
#pragma unroll
            for (int j = 0; j < R; j += 2) *reinterpret_cast<double2*>(dst + j) = make_double2(rec[j], rec[j + 1]);
        } else {
#pragma unroll
            for (int j = 0; j < R; ++j) dst[j] = rec[j];
        }
    }

    __device__ __forceinline__ void retire(bool fin, unsigned long long idx, uint32_t n_acc) {
        if (fin && hist_len) hist_len[idx] = n_acc < cap ? n_acc : cap;
    }
};

}  // namespace bacon
