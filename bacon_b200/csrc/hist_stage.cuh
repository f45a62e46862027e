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
// store instruction fills 32 whole sectors, and the four sectors of a 128-byte line arrive within a few
// microseconds of each other, which is young enough for L2 to merge them into one line write.  Other
// dimensions use the widest stores the record's alignment allows (128-bit when 1 + D is even, else 64-bit).
//
// What decides the speed is WHERE the 32 lanes of one store instruction write (profiles/r01i_dense_output.md):
// 32 different 2 MB pages cost four times as much as a few pages (address translation), so the driver hands
// trajectories out in per-warp blocks of consecutive indices (WarpQueue, drive.cuh).  tools/
// hist_compute_probe.cu is the synthetic twin that shows the ceiling: 128 DFMA + one such store per step
// runs at 96 % of the FP64 rate while writing 4.4 TB/s when the lanes of a warp are neighbours, 1.16 TB/s
// when they are not.  Staging whole 128-byte lines per lane in shared memory or registers (needed when
// NOTHING separates the stores: tools/hist_write_probe.cu) buys nothing here and was dropped, like the
// round-1 staging of 8 points per lane with a warp-cooperative copy-out into separate t / y arrays.
#pragma once
#include "ivp_common.cuh"

namespace bacon {

template <int D, bool ENABLED> struct HistStage;

template <int D> struct HistStage<D, false> {
    __device__ __forceinline__ explicit HistStage(const bacon_launch_args&) {}
    __device__ __forceinline__ void begin(unsigned long long) {}
    __device__ __forceinline__ void mute() {}
    __device__ __forceinline__ void push(bool, uint32_t, double, const double (&)[D]) {}
    __device__ __forceinline__ void retire(unsigned long long, uint32_t) {}
};

template <int D> struct HistStage<D, true> {
    static constexpr int R = 1 + D;  // doubles per record

    double* hist;
    uint32_t* hist_len;
    uint32_t cap, cap_all;  // records kept per trajectory: for the lane's current trajectory (0 = muted) / as configured
    double* base;           // the current trajectory's first record

    __device__ __forceinline__ explicit HistStage(const bacon_launch_args& a) {
        hist = a.out.hist;
        hist_len = a.out.hist_len;
        cap = cap_all = (uint32_t)a.cfg.history_capacity;
        base = hist;
    }
    // this lane now runs trajectory idx
    __device__ __forceinline__ void begin(unsigned long long idx) {
        cap = cap_all;
        base = hist + (size_t)idx * cap * R;
    }
    // ... or a ghost (drive.cuh): nothing is written until the next begin()
    __device__ __forceinline__ void mute() { cap = 0; }

    // called by every lane once per step() call; lanes that yielded a point write its record
    __device__ __forceinline__ void push(bool yielded, uint32_t n_acc_before, double t, const double (&y)[D]) {
        if (!(yielded && n_acc_before < cap)) return;
        double* dst = base + (size_t)n_acc_before * R;
        double r[R];
        r[0] = t;
#pragma unroll
        for (int d = 0; d < D; ++d) r[1 + d] = y[d];
        if constexpr (R % 4 == 0) {  // 32-byte records: whole sectors
#pragma unroll
            for (int j = 0; j < R; j += 4) {
#if defined(__CUDACC_VER_MAJOR__) && (__CUDACC_VER_MAJOR__ * 100 + __CUDACC_VER_MINOR__ < 1209)
                // 256-bit vector stores need PTX ISA 8.8 (CUDA 12.9): an older NVRTC (bacon_rhs_register_source picks up
                // whatever libnvrtc.so.12 the process has) gets the same sector as two 128-bit halves
                asm volatile("st.global.v2.f64 [%0], {%1, %2};\n\tst.global.v2.f64 [%0+16], {%3, %4};" ::"l"(dst + j), "d"(r[j]),
                             "d"(r[j + 1]), "d"(r[j + 2]), "d"(r[j + 3]));
#else
                asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(dst + j), "d"(r[j]), "d"(r[j + 1]),
                             "d"(r[j + 2]), "d"(r[j + 3]));
#endif
            }
        } else if constexpr (R % 2 == 0) {
#pragma unroll
            for (int j = 0; j < R; j += 2) *reinterpret_cast<double2*>(dst + j) = make_double2(r[j], r[j + 1]);
        } else {
#pragma unroll
            for (int j = 0; j < R; ++j) dst[j] = r[j];
        }
    }

    __device__ __forceinline__ void retire(unsigned long long idx, uint32_t n_acc) {
        if (hist_len) hist_len[idx] = n_acc < cap ? n_acc : cap;
    }
};

}  // namespace bacon
