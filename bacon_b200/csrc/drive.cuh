// drive.cuh — the persistent-lane ensemble driver shared by every thread-per-
// trajectory kernel: the device form of `IVPIterator::next` (src/ivp.rs:220-238)
// run for n trajectories at once.
//
//   loop { match solver.step() { Ok -> yield, Redo -> continue, Done -> stop,
//                                Failure(e) -> yield Err(e) once, stop } }
//
// Divergent trajectory lengths: what diverges is the END of a trajectory.  A lane
// whose trajectory retired (Done / Failure) stores its record and re-arms itself
// with the next trajectory index from the global work counter, so a warp only
// idles lanes once the whole ensemble has been handed out.  All lanes start
// together and trajectories of one ensemble are of similar length, so the lanes
// stay roughly in phase: the launcher (launch.cuh) picks the number of resident
// CTAs that leaves the LAST round of trajectories as full as possible.
// (Measured and dropped in round 1: suspending the CTA when the counter runs dry,
// sorting its 128 trajectories by remaining time through shared memory and
// re-dealing one quartile per warp — 71.3 % of FP64 peak against 72.9 % without,
// profiles/r01g_ab.md.)
#pragma once
#include "hist_stage.cuh"
#include "ivp_common.cuh"

namespace bacon {

constexpr int ENSEMBLE_BLOCK = 128;

// How the driver reads the accepted-step count and the exit status of a stepper.  Default: a field `n_acc` and a plain
// bacon_status.  Steppers that keep the count off their hot path (RkFastStepper) provide acc_running(), acc_of(raw) and
// status_of(raw) for the encoded value their attempt() returns.
template <class S, class = void> struct StepperCodec {
    __device__ __forceinline__ static int status(const S&, int raw) { return raw; }
    __device__ __forceinline__ static uint32_t acc(const S& s, int) { return s.n_acc; }
    __device__ __forceinline__ static uint32_t acc_running(const S& s) { return s.n_acc; }
};
template <class S> struct StepperCodec<S, decltype(void(&S::acc_running))> {
    __device__ __forceinline__ static int status(const S&, int raw) { return S::status_of(raw); }
    __device__ __forceinline__ static uint32_t acc(const S& s, int raw) { return s.acc_of(raw); }
    __device__ __forceinline__ static uint32_t acc_running(const S& s) { return s.acc_running(); }
};

constexpr int RAW_RUNNING = -1;  // attempt(): the trajectory goes on

template <class Stepper, bool HIST, int MINB>
__global__ void __launch_bounds__(ENSEMBLE_BLOCK, MINB) ensemble_kernel(const __grid_constant__ bacon_launch_args a) {
    constexpr int D = Stepper::D;
    using Codec = StepperCodec<Stepper>;

    Stepper s(a);
    HistStage<D, HIST> hist(a);
    const unsigned long long n = a.n;

    unsigned long long idx = warp_fetch(a.work_counter, true);
    if (idx >= n) return;
    s.reset(a, idx, true);

    // No vote and no liveness test in the loop: a lane whose trajectory ends leaves the common path on its own (the
    // branch is inside attempt()), stores its record, takes the next trajectory index with its own atomicAdd (one per
    // trajectory: ~3e7/s for the whole GPU, and the compiler aggregates lanes that arrive together) and rejoins its
    // warp at the next attempt.  A lane that finds the counter dry is done.
    for (;;) {
        bool yielded = false;
        const uint32_t n_acc_before = HIST ? Codec::acc_running(s) : 0u;
        const int raw = s.attempt(yielded);
        hist.push(yielded, n_acc_before, idx, s.out_t(), s.out_y());
        if (raw != RAW_RUNNING) {  // rare
            const uint32_t n_acc = Codec::acc(s, raw);
            hist.retire(true, idx, n_acc);
            int st = Codec::status(s, raw);
            if (HIST && st == BACON_OK && n_acc > (uint32_t)a.cfg.history_capacity) st = BACON_E_HISTORY_OVERFLOW;
            store_result<D>(a.out, n, idx, s.end_y(), s.t, s.dt, st, n_acc, s.n_rej, s.n_rhs());
            idx = atomicAdd(a.work_counter, 1ull);
            if (idx >= n) return;
            s.reset(a, idx, true);
        }
    }
}

}  // namespace bacon
