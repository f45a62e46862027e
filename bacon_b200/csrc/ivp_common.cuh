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
    // Optional inputs (bacon_ivp_options): restart record and terminal event.  NULL / 0 = the plain solve.
    const double* t0_each;   // [n] device: per-trajectory initial time
    const double* dt0_each;  // [n] device: per-trajectory first dt (clamped into [dt_min, dt_max])
    int32_t ev_on;           // != 0: the launcher picked the kernels compiled with the terminal-event test
    int32_t ev_direction;    // +1 rising, -1 falling, 0 both
    double ev_c;
    double ev_w[32];         // g(y) = ev_w . y - ev_c  (dim <= 32)
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

// The zero-copy host path's "late inputs are in HBM" flag, raised by a DMA on a second stream.  CUDA promises no forward
// progress between streams (a persistent grid that fills the machine may be all that runs), so the wait is BOUNDED:
// after LATE_POLLS polls (a few hundred microseconds; the DMA of config 2's 25 MB takes ~1 ms and the first refill
// comes ~3 ms into the launch) the lane gives up and reads the caller's mapped host memory instead, which is valid for
// every index.  The poll is a system-scope RELAXED load (served by L2, the point of coherence of the copy engine's
// writes), not an acquire: ld.acquire.sys carries a system-scope fence, and on this path every retiring lane has just
// posted its result record to host memory over PCIe — the fence waits for those writes, once per refill, with 31
// warp-mates at the loop's latch (measured: config 2 end to end 30.9 -> 136.6 ms).  What orders the data instead: the
// DMA of y0_late / params_late precedes the flag's memset in stream order; this kernel never touches those buffers
// before it has seen the flag set (no stale line can sit in its L1, which starts empty at launch); and the loads that
// follow depend on the flag's value through a branch the hardware does not speculate past (the asm is volatile with
// a memory clobber, so the compiler keeps the order too).
// (not inlined: a loop inside the persistent loop's body makes ptxas re-load the tableau from the constant bank on every
// attempt instead of keeping it in uniform registers)
constexpr unsigned LATE_POLLS = 1024;
static __device__ __noinline__ bool wait_until_set(const unsigned int* flag) {
    for (unsigned k = 0; k < LATE_POLLS; ++k) {
        unsigned v;
        asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
        if (v != 0) return true;
    }
    return false;
}

// Where a trajectory starts: (cfg.t_start, the reference's dt0 = (dt_max + dt_min)/2 computed by the stepper in its own
// arithmetic) unless the launch carries cfg.dt_init or a restart record (bacon_ivp_options).  Called once per
// trajectory by the kernels compiled for the optional inputs only (drive.cuh: EVENT / Stepper::apply_restart): in the
// plain fast RK kernels a first dt that is not a function of dt_min and dt_max costs a register pair across the
// persistent loop, i.e. one constant re-load per attempt (tools/sass_count.py: 55 -> 57 other instructions per trip).
__device__ __forceinline__ void trajectory_start(const bacon_launch_args& a, unsigned long long i, double dt_min,
                                                 double dt_max, double& t, double& dt) {
    if (a.t0_each) t = a.t0_each[i];
    double d = a.cfg.dt_init;  // bacon_solver_with_initial_dt; 0 = the reference's default, already in dt
    if (a.dt0_each) d = a.dt0_each[i];
    if (a.dt0_each || d > 0.0) dt = !(d >= dt_min) ? dt_min : (d > dt_max ? dt_max : d);  // (NaN -> dt_min)
}
// cfg.dt_init for kernels that read it once at start (rk_warp_linear.cuh); 0 = not given
__device__ __forceinline__ bool initial_dt_given(const bacon_launch_args& a, double dt_min, double dt_max, double& dt0) {
    const double d = a.cfg.dt_init;
    if (!(d > 0.0)) return false;
    dt0 = d < dt_min ? dt_min : (d > dt_max ? dt_max : d);
    return true;
}

template <int D, int P>
__device__ __forceinline__ void load_problem(const bacon_launch_args& a, unsigned long long i,
                                             double (&y)[D], double (&p)[(P > 0 ? P : 1)]) {
    const double* y0 = a.y0;
    const double* params = a.params;
    if (a.y0_late && i >= a.late_from) {  // a refill on the zero-copy host path: the DMA-ed copy, not the host link
        if (wait_until_set(a.late_ready)) {
            y0 = a.y0_late;
            if (a.params_late) params = a.params_late;
        }
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
