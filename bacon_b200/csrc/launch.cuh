// launch.cuh — host-side launchers, instantiated in the translation unit that
// defines the RHS functor (built-ins: rhs_builtin_*.cu; user RHS: their own .cu,
// see INTEGRATION.md).  A launcher picks the kernel variant (final state only /
// dense output), sizes a PERSISTENT grid (resident CTAs per SM x SM count: every
// lane stays on the machine and pulls trajectories from the work counter) and
// enqueues it on the caller's stream.
#pragma once
#include <cuda_runtime.h>

#include <cstdlib>

#include "adams.cuh"
#include "bdf.cuh"
#include "drive.cuh"
#include "path_query.cuh"
#include "rk_fast.cuh"
#include "rk_strict.cuh"

namespace bacon {

// grid of a persistent launch: resident CTAs per SM (occupancy of `kernel`) x SM count, capped by the work (one
// bundle of 32 trajectories per warp: a small ensemble spreads over as many CTAs as it has bundles)
template <int BLOCK, class K> inline int persistent_grid(K kernel, bacon_launch_args* a, size_t smem) {
    cudaError_t e;
    if (smem > 32 * 1024) {  // (static + dynamic above 48 KB needs the opt-in: the kernel holds a few hundred static bytes)
        e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return BACON_E_CUDA;
    }
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, BLOCK, smem);
    if (e != cudaSuccess || per_sm < 1) return BACON_E_CUDA;
    cudaFuncAttributes fa;
    e = cudaFuncGetAttributes(&fa, kernel);
    if (e != cudaSuccess) return BACON_E_CUDA;
    if (const char* env = getenv("BACON_IVP_BLOCKS_PER_SM")) {  // tuning knob: fewer resident CTAs than the occupancy limit
        const int want = atoi(env);
        if (want >= 1 && want < per_sm) per_sm = want;
    }
    long long grid = (long long)per_sm * a->sm_count;
    const long long need = (long long)((a->n + 31) / 32);
    if (grid > need) grid = need;
    if (a->grid_override > 0) grid = a->grid_override;
    if (const char* env = getenv("BACON_IVP_GRID")) {  // a small grid makes small ensembles refill, run dry and regroup
        const int want = atoi(env);                     // (tests, compute-sanitizer runs)
        if (want >= 1 && want < grid) grid = want;
    }
    if (grid < 1) grid = 1;
    a->grid = (int)grid;
    a->block = BLOCK;
    a->regs_per_thread = fa.numRegs;
    a->late_from = (unsigned long long)grid * BLOCK;  // the first trajectory of every lane comes before this
    return 0;
}

// Steppers that can suspend and resume a trajectory (Stepper::save / load) run as ONE CTA per SM holding all the
// lanes the register budget allows (128 x MINB), so that the end-of-ensemble regrouping (drive.cuh) is SM-wide;
// the others keep 128-lane CTAs, MINB of them per SM.
template <class Stepper, class = void> struct StepperSuspends { static constexpr bool value = false; };
template <class Stepper> struct StepperSuspends<Stepper, decltype(void(Stepper::STATE_DOUBLES))> { static constexpr bool value = true; };

// EVENT: the instantiation that watches a terminal event (drive.cuh: EventWatch); 128-lane CTAs, no regrouping.
// EVENT is part of every launcher's template signature: the plain and the event launchers of one stepper may be
// compiled in different translation units and must not share a symbol.
template <class Stepper, bool HIST, int MINB, bool EVENT = false> inline int launch_stepper_hist(bacon_launch_args* a) {
    constexpr bool WIDE = !EVENT && StepperSuspends<Stepper>::value && StepperMigrates<Stepper, ENSEMBLE_BLOCK * MINB>::value;
    constexpr int BLOCK = WIDE ? ENSEMBLE_BLOCK * MINB : ENSEMBLE_BLOCK;
    auto kernel = ensemble_kernel<Stepper, HIST, BLOCK, (WIDE ? 1 : MINB), EVENT>;
    constexpr size_t smem = ensemble_smem_bytes<Stepper, BLOCK>();
    if (const int rc = persistent_grid<BLOCK>(kernel, a, smem)) return rc;
    a->n_kernels = 1;
    kernel<<<(unsigned)a->grid, BLOCK, smem, (cudaStream_t)a->stream>>>(*a);
    if (cudaGetLastError() != cudaSuccess) return BACON_E_CUDA;
    return 0;
}

// An ensemble that does not fit the machine's lanes ONCE but would with one more warp per sub-partition (final state
// only; e.g. the 2^20-trajectory ensemble sharded over 8 GPUs: 131072 per GPU against 148 x 768 = 113664 lanes):
// with 768 lanes per SM the 17408 trajectories left over start when the first lanes free up, i.e. when everything
// else is nearly finished, and then run their ~3600 attempts almost alone (measured 4.9 ms against 3.7 ms of work).
// The same kernel compiled for 128 x (MINB + 1) lanes per SM (Lorenz / RKF45: 69 registers instead of 78, no
// spills) holds all of them from the start; a sub-partition saturates its FP64 pipe from ~4 resident warps
// (tools/tau_probe.py), so the seventh costs nothing.  On larger ensembles the narrower kernel is 1.7 % faster
// (profiles/r02_strong_scaling.md), so this one is used for that window of sizes only.
template <class Stepper, int MINB> inline bool fits_one_more_warp(const bacon_launch_args* a) {
    if constexpr (!(StepperSuspends<Stepper>::value && StepperMigrates<Stepper, ENSEMBLE_BLOCK * (MINB + 1)>::value &&
                    ENSEMBLE_BLOCK * (MINB + 1) <= 1024 && MINB >= 6)) {
        return false;
    } else {
        static const bool off = getenv("BACON_IVP_NO_FIT") != nullptr;  // (A/B switch for measurements)
        const unsigned long long lanes = (unsigned long long)a->sm_count * ENSEMBLE_BLOCK * MINB;
        const unsigned long long lanes_fit = (unsigned long long)a->sm_count * ENSEMBLE_BLOCK * (MINB + 1);
        return !off && a->grid_override <= 0 && a->n > lanes && a->n <= lanes_fit;
    }
}

template <class Stepper, int MINB, bool EVENT = false> inline int launch_stepper(bacon_launch_args* a) {
    // Dense output: one resident CTA fewer when the budget is tight (6 -> 5: 80 -> 96 registers, no re-loads of launch
    // constants inside the loop).  Measured 1-2 % faster than 6 (profiles/r01i_dense_output.md).
#ifndef BACON_HIST_MINB_DROP
#define BACON_HIST_MINB_DROP 1
#endif
    constexpr int MINB_HIST = MINB >= 6 ? MINB - BACON_HIST_MINB_DROP : MINB;
    if (a->cfg.history_capacity > 0 && a->out.hist) return launch_stepper_hist<Stepper, true, MINB_HIST, EVENT>(a);
    if constexpr (!EVENT && StepperSuspends<Stepper>::value && MINB >= 6 && ENSEMBLE_BLOCK * (MINB + 1) <= 1024) {
        if (fits_one_more_warp<Stepper, MINB>(a)) return launch_stepper_hist<Stepper, false, MINB + 1>(a);
    }
    return launch_stepper_hist<Stepper, false, MINB, EVENT>(a);
}

// ---- fast (FMA, compile-time tableau), REF_CORRECTED only
// resident 128-thread CTAs per SM the fast RK kernels are compiled for (register budget): 6 (80 registers) when the
// state y, the stages k and the parameters fit (Lorenz/RKF45 needs 74), else 4 (128 registers)
template <class Rhs, class Tab> constexpr int rk_fast_minb() {
#ifdef BACON_RK_MINB
    return BACON_RK_MINB;
#else
    return Rhs::DIM * (Tab::O + 1) + Rhs::NPARAM <= 28 ? 6 : 4;
#endif
}
template <class Rhs, class Tab, bool EVENT = false, int MINB = rk_fast_minb<Rhs, Tab>()> int launch_rk_fast(bacon_launch_args* a) {
    if (a->cfg.semantics != BACON_SEM_CORRECTED) return BACON_E_UNSUPPORTED;
    return launch_stepper<RkFastStepper<Rhs, Tab>, MINB, EVENT>(a);
}

// ---- strict (oracle operation order, runtime tableau in __constant__), either semantics
template <class Rhs, class Tab, bool EVENT = false, int MINB = 1> int launch_rk_strict(bacon_launch_args* a) {
    RkTableauRt T;
    fill_runtime_tableau<Tab>(T, a->cfg.semantics == BACON_SEM_LITERAL);
    if (cudaMemcpyToSymbolAsync(c_rk_tabs, &T, sizeof(T), sizeof(T) * rk_tab_slot(Tab::O, a->cfg.semantics),
                                cudaMemcpyHostToDevice, (cudaStream_t)a->stream) != cudaSuccess)
        return BACON_E_CUDA;
    return launch_stepper<RkStrictStepper<Rhs, Tab::O>, MINB, EVENT>(a);
}

// ---- BDF (Broyden as in the reference, or Newton + in-register LU with BACON_FLAG_BDF_NEWTON)
// Resident CTAs per SM the kernels are compiled for.  Newton + in-register LU: 4 (128 registers, ~170 bytes spilled on
// cold paths) — the kernel is latency-bound (ncu: 3 warps per sub-partition at 168 registers, 45 % of stalls are
// fixed-latency waits), measured 7.38e10 steps/s against 7.00e10 at 3 and 7.09e10 at 5.  The reference's Broyden
// iteration keeps two D x D matrices alive and stays at 2.
#ifndef BACON_BDF_NEWTON_MINB
#define BACON_BDF_NEWTON_MINB 4  // (A/B switch for measurements)
#endif
template <class Rhs, class Coef, bool STRICT, bool EVENT = false, int MINB = 2, int MINB_NEWTON = BACON_BDF_NEWTON_MINB> int launch_bdf(bacon_launch_args* a) {
    // REF_LITERAL (bdf.rs:407, :568, :622 as written) is the reference's own Broyden iteration in the strict build
    if (a->cfg.semantics != BACON_SEM_CORRECTED && !(STRICT && !(a->cfg.flags & BACON_FLAG_BDF_NEWTON))) return BACON_E_UNSUPPORTED;
    if (a->cfg.flags & BACON_FLAG_BDF_NEWTON) {
        // the strict build is the reference's own (Broyden) iteration.  `if constexpr`: a strict translation
        // unit (-fmad=false) must not instantiate a kernel that a fast one also defines under the same name.
        if constexpr (STRICT) return BACON_E_UNSUPPORTED;
        else return launch_stepper<BdfStepper<Rhs, Coef, false, true>, MINB_NEWTON, EVENT>(a);
    }
    return launch_stepper<BdfStepper<Rhs, Coef, STRICT, false>, MINB, EVENT>(a);
}

// ---- Adams predictor-corrector and Euler.  No D1-D9 defect on these paths:
// both semantics run the same kernel.
template <class Rhs, class Coef, bool STRICT, bool EVENT = false, int MINB = 2> int launch_adams(bacon_launch_args* a) {
    return launch_stepper<AdamsStepper<Rhs, Coef, STRICT>, MINB, EVENT>(a);
}
template <class Rhs, bool STRICT, bool EVENT = false, int MINB = 4> int launch_euler(bacon_launch_args* a) {
    return launch_stepper<EulerStepper<Rhs, STRICT>, MINB, EVENT>(a);
}

// ---- registration: fills the launcher table of one RHS for THIS translation unit's build
// flavour (fast, or strict when compiled with -DBACON_STRICT_FP -fmad=false) and hands it
// to the engine through the C ABI (bacon_rhs_register merges the two flavours by name).
template <class Rhs, int S, bool EVENT> void fill_launchers(bacon_launch_fn (&t)[2][BACON_N_METHODS]) {
#ifndef BACON_SKIP_RK
    // `if constexpr`, not `?:` — a conditional expression would instantiate BOTH launchers in BOTH builds, and the
    // two objects would then carry same-named kernels compiled with different -fmad settings (one of which the
    // runtime picks arbitrarily).  Each kernel must exist in exactly one object: tests/test_abi.py checks it.
    if constexpr (S) {
        t[S][BACON_RK45] = &launch_rk_strict<Rhs, TabRKF45, EVENT>;
        t[S][BACON_RK23] = &launch_rk_strict<Rhs, TabBS23, EVENT>;
    } else {
        t[S][BACON_RK45] = &launch_rk_fast<Rhs, TabRKF45, EVENT>;
        t[S][BACON_RK23] = &launch_rk_fast<Rhs, TabBS23, EVENT>;
    }
#endif
#ifndef BACON_SKIP_BDF
    t[S][BACON_BDF6] = &launch_bdf<Rhs, CoefBDF6, S != 0, EVENT>;
    t[S][BACON_BDF2] = &launch_bdf<Rhs, CoefBDF2, S != 0, EVENT>;
#endif
#ifndef BACON_SKIP_ADAMS
    t[S][BACON_ADAMS5] = &launch_adams<Rhs, CoefAdams5, S != 0, EVENT>;
    t[S][BACON_ADAMS3] = &launch_adams<Rhs, CoefAdams3, S != 0, EVENT>;
    t[S][BACON_EULER] = &launch_euler<Rhs, S != 0, EVENT>;
#endif
}

template <class Rhs> int register_rhs(const char* name) {
    bacon_rhs_desc d{};
    d.name = name;
    d.dim = Rhs::DIM;
    d.n_params = Rhs::NPARAM;
#ifdef BACON_STRICT_FP
    constexpr int S = 1;
#else
    constexpr int S = 0;
#endif
    // The library's own build compiles the plain kernels (-DBACON_SKIP_EVENTS) and the terminal-event kernels
    // (-DBACON_ONLY_EVENTS) in separate objects, and the families -DBACON_SKIP_RK / _BDF / _ADAMS leave out, only to
    // compile them in parallel; a user translation unit defines none of these and gets everything.
#ifndef BACON_ONLY_EVENTS
    fill_launchers<Rhs, S, false>(d.launch);
#ifndef BACON_SKIP_RK
    // queries on stored paths (path_query.cuh) serve every method; they live in the object that holds the RK kernels
    d.path_query[S] = &launch_path_query<Rhs, S != 0>;
#endif
#endif
#ifndef BACON_SKIP_EVENTS
    fill_launchers<Rhs, S, true>(d.launch_event);
#endif
    return bacon_rhs_register(&d);
}

}  // namespace bacon

#define BACON_CAT2(a, b) a##b
#define BACON_CAT(a, b) BACON_CAT2(a, b)
// BACON_REGISTER_RHS(MyRhs, "my_rhs"); at namespace scope of a .cu file
#define BACON_REGISTER_RHS(RhsType, name) \
    static const int BACON_CAT(bacon_rhs_id_, __LINE__) = ::bacon::register_rhs<RhsType>(name)
