// ivp_common.cuh — shared device-side plumbing of the ensemble IVP kernels.
//
// One trajectory = one `solve(data)` + `IVPIterator::collect_vec` of the
// reference (src/ivp.rs:209-238).  A kernel owns a grid of persistent lanes; a
// lane integrates one trajectory at a time with its whole stepper state in
// registers, and when the trajectory retires (Done / Failure) the lane stores
// its record and pulls the next trajectory index from the launch's work counter
// (drive.cuh; the first 32 indices of a warp come from ONE warp-aggregated
// atomicAdd, warp_fetch below; rk_warp_linear.cuh uses it throughout).
#pragma once
#ifndef __CUDACC_RTC__
#include <cuda_runtime.h>
#include <stdint.h>
#endif

#include "../../include/bacon_ivp.h"

// Everything a launcher needs.  Plain C layout: it crosses the C ABI inside
// bacon_rhs_desc::launch (user RHS translation units fill it the same way).
struct bacon_launch_args {
    bacon_ivp_config cfg;
    unsigned long long n;              // trajectories in this launch
    const double* y0;                  // [dim][n] device
    const double* params;              // [n_params][n] or [n_params] device (or NULL)
    bacon_ivp_result out;              // device pointers
    unsigned long long* work_counter;  // device scratch, zeroed by the engine before the launch
    void* stream;                      // cudaStream_t
    int32_t sm_count;                  // SMs of the current device
    int32_t grid_override;             // >0: force this grid (tests)
    // filled by the launcher
    int32_t grid, block, regs_per_thread, n_kernels;
    // Late inputs (zero-copy host path): trajectories >= late_from are read from y0_late / params_late (a device copy
    // that a DMA fills while the first trajectories, read straight from the caller's pinned memory, already run)
    // once *late_ready != 0.  All NULL / 0 otherwise.  late_from is set by the launcher (= the lanes of the grid).
    const double* y0_late;
    const double* params_late;
    const unsigned int* late_ready;
    unsigned long long late_from;
};

namespace bacon {

constexpr unsigned FULL_MASK = 0xffffffffu;

// compile-time loop with a constexpr index (tableau zeros vanish at compile time)
template <int I> struct IC { static constexpr int value = I; constexpr operator int() const { return I; } };
template <int B, int E, class F> __host__ __device__ __forceinline__ void static_for(F&& f) {
    if constexpr (B < E) {
        f(IC<B>{});
        static_for<B + 1, E>(static_cast<F&&>(f));
    }
}

__device__ __forceinline__ unsigned lane_id() {
    unsigned l;
    asm volatile("mov.u32 %0, %%laneid;" : "=r"(l));
    return l;
}
__device__ __forceinline__ unsigned lanemask_lt() {
    unsigned m;
    asm volatile("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// Warp-aggregated work fetch: lanes with `want` set receive consecutive
// trajectory indices; one atomicAdd per warp.  Must be called by all 32 lanes.
__device__ __forceinline__ unsigned long long warp_fetch(unsigned long long* counter, bool want) {
    const unsigned m = __ballot_sync(FULL_MASK, want);
    if (m == 0) return ~0ull;
    const int leader = __ffs(m) - 1;
    unsigned long long base = 0;
    if ((int)lane_id() == leader) base = atomicAdd(counter, (unsigned long long)__popc(m));
    base = __shfl_sync(FULL_MASK, base, leader);
    return base + (unsigned long long)__popc(m & lanemask_lt());
}

// Per-trajectory record written at retirement.  Scattered 4/8-byte stores: a few
// dozen bytes per trajectory against thousands of steps of register-only work.
template <int D>
__device__ __forceinline__ void store_result(const bacon_ivp_result& o, unsigned long long n,
                                             unsigned long long i, const double (&y)[D], double t,
                                             double dt, int status, uint32_t n_acc, uint32_t n_rej,
                                             uint32_t n_rhs) {
#pragma unroll
    for (int d = 0; d < D; ++d) o.y_end[(size_t)d * n + i] = y[d];
    if (o.t_end) o.t_end[i] = t;
    if (o.dt_end) o.dt_end[i] = dt;
    o.status[i] = status;
    if (o.n_accept) o.n_accept[i] = n_acc;
    if (o.n_reject) o.n_reject[i] = n_rej;
    if (o.n_rhs) o.n_rhs[i] = n_rhs;
}

// (not inlined: a loop inside the persistent loop's body makes ptxas re-load the tableau from the constant bank on every
// attempt instead of keeping it in uniform registers)
static __device__ __noinline__ void wait_until_set(const unsigned int* flag) {
    while (*(volatile const unsigned int*)flag == 0) {
    }
}

template <int D, int P>
__device__ __forceinline__ void load_problem(const bacon_launch_args& a, unsigned long long i,
                                             double (&y)[D], double (&p)[(P > 0 ? P : 1)]) {
    const double* y0 = a.y0;
    const double* params = a.params;
    if (a.y0_late && i >= a.late_from) {  // a refill on the zero-copy host path: the DMA-ed copy, not the host link
        wait_until_set(a.late_ready);
        y0 = a.y0_late;
        if (a.params_late) params = a.params_late;
    }
#pragma unroll
    for (int d = 0; d < D; ++d) y[d] = y0[(size_t)d * a.n + i];
    if constexpr (P > 0) {
        const bool shared = (a.cfg.flags & BACON_FLAG_SHARED_PARAMS) != 0;
        const bool aos = (a.cfg.flags & BACON_FLAG_PARAMS_AOS) != 0;
#pragma unroll
        for (int k = 0; k < P; ++k)
            p[k] = shared ? params[k] : (aos ? params[(size_t)i * P + k] : params[(size_t)k * a.n + i]);
    }
}

}  // namespace bacon
