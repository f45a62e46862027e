// hist_stage.cuh — dense output (K4): every accepted (t, y) of every trajectory,
// i.e. what `IVPIterator::collect_vec` materialises (src/ivp.rs:209-211, the
// initial condition is not yielded: rk.rs:418-419).
//
// HBM layout (bacon_ivp.h): hist_t[n][cap], hist_y[n][cap][D] — one contiguous
// `Path` per trajectory.  Lanes of a warp sit at different trajectories and
// different step counts, so a lane storing its own point would issue scattered
// 8-byte writes.  Instead each lane stages S points in shared memory
// (pair-interleaved layout: conflict-free 128-bit reads for the copy-out, at
// worst 2-way conflicts for the per-lane 64-bit stage writes) and the WARP
// copies a full lane buffer out as aligned 128-bit stores: S*D*8 contiguous
// bytes of hist_y and S*8 of hist_t per lane buffer, whole 32-byte sectors only.
#pragma once
#include "ivp_common.cuh"

namespace bacon {

template <int D, bool ENABLED> struct HistStage;

template <int D> struct HistStage<D, false> {
    static constexpr size_t smem_bytes(int) { return 0; }
    __device__ __forceinline__ HistStage(const bacon_launch_args&, unsigned char*) {}
    __device__ __forceinline__ void begin() {}
    __device__ __forceinline__ void push(bool, uint32_t, unsigned long long, double, const double (&)[D]) {}
    __device__ __forceinline__ void retire(bool, unsigned long long, uint32_t) {}
};

template <int D> struct HistStage<D, true> {
    static constexpr int S = 8;                   // staged points per lane (even)
    static constexpr int W = S * (D + 1);         // doubles per lane buffer: y part [S][D] then t part [S]
    static constexpr int PAIRS = W / 2;
    static constexpr int NY2 = S * D / 2;         // double2 elements of the y part
    static constexpr int NT2 = S / 2;             // double2 elements of the t part
    static constexpr size_t smem_bytes(int warps) { return (size_t)warps * PAIRS * 33 * 16; }

    double2* wbuf;     // this warp's staging area
    double* hist_t;
    double* hist_y;
    uint32_t* hist_len;
    uint32_t cap;
    uint32_t cnt;      // points staged by this lane
    uint32_t written;  // points of this lane's trajectory already in HBM (multiple of S)
    bool vec_ok;       // 128-bit path allowed (cap even -> every lane buffer is 16-byte aligned)
    unsigned lane;

    __device__ __forceinline__ HistStage(const bacon_launch_args& a, unsigned char* smem) {
        lane = lane_id();
        wbuf = reinterpret_cast<double2*>(smem) + (size_t)(threadIdx.x >> 5) * PAIRS * 33;
        hist_t = a.out.hist_t;
        hist_y = a.out.hist_y;
        hist_len = a.out.hist_len;
        cap = (uint32_t)a.cfg.history_capacity;
        vec_ok = (cap % 2u) == 0u && ((reinterpret_cast<uintptr_t>(hist_t) | reinterpret_cast<uintptr_t>(hist_y)) & 15u) == 0u;
        cnt = 0;
        written = 0;
    }
    __device__ __forceinline__ void begin() { cnt = 0; written = 0; }

    // element j (0..W) of lane l's buffer
    __device__ __forceinline__ double& at(int j, unsigned l) {
        return reinterpret_cast<double*>(wbuf + (size_t)(j >> 1) * 33 + l)[j & 1];
    }

    // warp-collective: copy the FULL buffers of the lanes in `mask` to HBM
    __device__ __forceinline__ void flush_full(unsigned mask, unsigned long long idx) {
        while (mask) {
            const int L = __ffs(mask) - 1;
            mask &= mask - 1;
            const unsigned long long iL = __shfl_sync(FULL_MASK, idx, L);
            const uint32_t wL = __shfl_sync(FULL_MASK, written, L);
            const size_t row = (size_t)iL * cap + wL;
            if (vec_ok) {
                for (int j = lane; j < NY2 + NT2; j += 32) {
                    const double2 v = wbuf[(size_t)j * 33 + L];
                    double2* dst = (j < NY2) ? reinterpret_cast<double2*>(hist_y + row * D) + j
                                             : reinterpret_cast<double2*>(hist_t + row) + (j - NY2);
                    *dst = v;
                }
            } else {
                for (int j = lane; j < W; j += 32) {
                    const double v = at(j, L);
                    if (j < S * D) hist_y[row * D + j] = v;
                    else hist_t[row + (j - S * D)] = v;
                }
            }
        }
        __syncwarp();
    }

    // called by every lane of the warp once per attempt; `accepted` lanes stage (t, y)
    __device__ __forceinline__ void push(bool accepted, uint32_t n_acc_before, unsigned long long idx,
                                         double t, const double (&y)[D]) {
        const bool keep = accepted && n_acc_before < cap;
        if (keep) {
#pragma unroll
            for (int d = 0; d < D; ++d) at(cnt * D + d, lane) = y[d];
            at(S * D + cnt, lane) = t;
            cnt++;
        }
        const unsigned full = __ballot_sync(FULL_MASK, cnt == (uint32_t)S);
        if (full) {
            __syncwarp();
            flush_full(full, idx);
            if (cnt == (uint32_t)S) { written += S; cnt = 0; }
        }
    }

    // warp-collective: lanes with `fin` write their partial buffer and hist_len
    __device__ __forceinline__ void retire(bool fin, unsigned long long idx, uint32_t n_acc) {
        unsigned mask = __ballot_sync(FULL_MASK, fin && cnt > 0);
        __syncwarp();
        while (mask) {
            const int L = __ffs(mask) - 1;
            mask &= mask - 1;
            const unsigned long long iL = __shfl_sync(FULL_MASK, idx, L);
            const uint32_t wL = __shfl_sync(FULL_MASK, written, L);
            const int cL = (int)__shfl_sync(FULL_MASK, cnt, L);
            const size_t row = (size_t)iL * cap + wL;
            for (int j = lane; j < cL * D; j += 32) hist_y[row * D + j] = at(j, L);
            for (int j = lane; j < cL; j += 32) hist_t[row + j] = at(S * D + j, L);
        }
        __syncwarp();
        if (fin) {
            if (hist_len) hist_len[idx] = n_acc < cap ? n_acc : cap;
            cnt = 0;
            written = 0;
        }
    }
};

}  // namespace bacon
