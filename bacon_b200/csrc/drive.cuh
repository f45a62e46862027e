// drive.cuh — the persistent-lane ensemble driver shared by every thread-per-
// trajectory kernel: the device form of `IVPIterator::next` (src/ivp.rs:220-238)
// run for n trajectories at once.
//
//   loop { match solver.step() { Ok -> yield, Redo -> continue, Done -> stop,
//                                Failure(e) -> yield Err(e) once, stop } }
//
// Divergent trajectory lengths: accept/reject is predicated inside the stepper;
// what diverges is the END of a trajectory.  After every attempt the warp
// ballots the lanes whose trajectory retired (Done / Failure), those lanes store
// their record and are re-armed with the next trajectory indices from the global
// work counter (one atomicAdd per warp), so a warp only idles lanes once the
// whole ensemble has been handed out.
#pragma once
#include "hist_stage.cuh"
#include "ivp_common.cuh"

namespace bacon {

constexpr int ENSEMBLE_BLOCK = 128;

template <class Stepper, bool HIST, int MINB>
__global__ void __launch_bounds__(ENSEMBLE_BLOCK, MINB) ensemble_kernel(const __grid_constant__ bacon_launch_args a) {
    constexpr int D = Stepper::D;

    Stepper s(a);
    HistStage<D, HIST> hist(a);
    const unsigned long long n = a.n;

    unsigned long long idx = warp_fetch(a.work_counter, true);
    bool live = idx < n;
    s.reset(a, idx, live);

    if (!__any_sync(FULL_MASK, live)) return;
    for (;;) {
        int st = -1;
        bool yielded = false;
        const uint32_t n_acc_before = s.n_acc;
        if (live) st = s.attempt(yielded);
        hist.push(yielded, n_acc_before, idx, s.out_t(), s.out_y());

        // one vote per attempt: st >= 0 only on live lanes whose trajectory retired in this attempt
        const bool fin = st >= 0;
        if (__any_sync(FULL_MASK, fin)) {
            hist.retire(fin, idx, s.n_acc);
            if (fin) {
                if (HIST && st == BACON_OK && s.n_acc > (uint32_t)a.cfg.history_capacity)
                    st = BACON_E_HISTORY_OVERFLOW;
                store_result<D>(a.out, n, idx, s.end_y(), s.t, s.dt, st, s.n_acc, s.n_rej, s.n_rhs());
            }
            const unsigned long long nxt = warp_fetch(a.work_counter, fin);
            if (fin) {
                idx = nxt;
                live = idx < n;
                s.reset(a, idx, live);
            }
            if (!__any_sync(FULL_MASK, live)) return;  // the warp can only run dry right after a retirement
        }
    }
}

}  // namespace bacon
