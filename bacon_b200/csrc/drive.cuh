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
// idles lanes once the whole ensemble has been handed out.  What is left then is
// regrouped inside the kernel: the CTA's trajectories change lanes through shared
// memory so that half-empty warps fold away (ensemble_kernel, below).
#pragma once
#include "hist_stage.cuh"
#include "ivp_common.cuh"
#include "path_query.cuh"
#ifdef BACON_DRIVE_TRACE
#include <cstdio>
#endif

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

template <class S, class = void> struct StepperUnrolls { static constexpr bool value = false; };
template <class S> struct StepperUnrolls<S, decltype(void(S::UNROLL_DRIVER))> { static constexpr bool value = S::UNROLL_DRIVER; };

constexpr int RAW_RUNNING = -1;     // attempt(): the trajectory goes on
constexpr int RAW_CHECKPOINT = -2;  // attempt(): nothing was computed, the whole warp reports in (RkFastStepper::tick)

// does the stepper support suspending a trajectory and resuming it on another lane (regrouping, below)?  The CTA's
// exchange buffer holds STATE_DOUBLES + 1 words per lane in shared memory.
template <class S, int BLOCK, class = void> struct StepperMigrates { static constexpr bool value = false; };
#ifndef BACON_NO_MIGRATE  // (A/B switch for measurements)
template <class S, int BLOCK> struct StepperMigrates<S, BLOCK, decltype(void(S::STATE_DOUBLES))> {
    static constexpr bool value = BLOCK >= 256 && (S::STATE_DOUBLES + 1) * 8 * BLOCK <= 160 * 1024;
};
#endif
template <class S, int BLOCK> constexpr size_t ensemble_smem_bytes() {
    if constexpr (StepperMigrates<S, BLOCK>::value) return (size_t)(S::STATE_DOUBLES + 1) * 8 * BLOCK;
    else return 0;
}

// Work hand-out: per-warp blocks of consecutive trajectories.  Why blocks: with dense output every lane stores into its
// own trajectory's history, and ONE store instruction whose 32 lanes touch 32 different 2 MB pages is 4 times slower
// than one whose lanes stay within a few pages (address translation; tools/hist_compute_probe.cu: 1.16 against 4.4
// TB/s).  Lanes that each take "the next index" from one global counter end up with 32 unrelated trajectories per
// warp; lanes that take it from their WARP's block of <= 32 consecutive indices stay neighbours (a trajectory's history
// is cap * 8(1+D) bytes: 147 KB in config 2, so a block is 2-3 pages).  Block sizes shrink towards the end of the
// ensemble (guided self-scheduling) so that no warp sits on unstarted work while others idle.
// One 64-bit word per warp in shared memory: base << 16 | size << 8 | used.  Loop-free and safe under divergence: the
// lane whose atomicAdd finds the block exactly exhausted refills it from the global counter; a lane that arrives while
// that refill is in flight takes a single index from the global counter instead.
// The launch's work counter counts the trajectories handed out AFTER the static first deal: index = first + counter.
struct WarpQueue {
    unsigned long long* word;       // shared memory
    unsigned long long* global;     // the launch's work counter
    unsigned long long n;
    unsigned long long first;       // lanes of the grid: the indices below were dealt statically
    unsigned long long per_block;   // BACON_WQ_DIV x warps of the grid: remaining / per_block = next block size

    __device__ __noinline__ unsigned long long fetch() {
        const unsigned long long s = atomicAdd(word, 1ull);
        const unsigned used = (unsigned)(s & 0xff), size = (unsigned)((s >> 8) & 0xff);
        if (used < size) return (s >> 16) + used;
        if (used > size) return first + atomicAdd(global, 1ull);  // a refill is in flight (rare): take a single index instead
        // used == size: this lane refills
        const unsigned long long seen = first + *(volatile unsigned long long*)global;
        const unsigned long long rem = seen < n ? n - seen : 0;
        unsigned long long b = rem / per_block;
        b = b < 1 ? 1 : (b > 32 ? 32 : b);
        const unsigned long long nb = first + atomicAdd(global, b);
        // (a block may reach past n, or start there when the counter is dry: callers test idx < n, and every later
        // fetch from such a block returns an index >= n as well)
        atomicExch(word, (nb << 16) | (b << 8) | 1ull);
        return nb;
    }
    // nothing left to start in this warp's block and nothing left in the global counter
    __device__ __forceinline__ bool dry() const {
        const unsigned long long v = *(volatile unsigned long long*)word;
        const unsigned used = (unsigned)(v & 0xff), size = (unsigned)((v >> 8) & 0xff);
        return (used >= size || (v >> 16) + used >= n) && first + *(volatile unsigned long long*)global >= n;
    }
};

__device__ __forceinline__ void named_barrier(int id, int threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// Regrouping, control side (see ensemble_kernel).  Everything warp- or CTA-collective lives in functions that are NOT
// inlined, for two reasons.  (1) The lanes of one warp come here along different paths — live lanes from the
// persistent loop, lanes that hold nothing from idle_until_regrouped — and must meet at the SAME vote / shuffle /
// barrier instructions.  (2) ptxas keeps the tableau in uniform registers across the persistent loop only while the
// loop's function holds no other loop: with the waiting loop inlined it re-loads all coefficients from the constant
// bank on every attempt (28 LDCU per pair of attempts, tools/sass_count.py).
// The two polled flags of RegroupCtl (request, dry) are written and read without a barrier in between ON PURPOSE: every
// write that races stores the same value, a reader that misses it sees it at its next checkpoint, and what is
// exchanged afterwards is ordered by the named barrier of the meeting.  They are accessed with morally strong relaxed
// operations at CTA scope (the PTX memory model's race-free form of exactly this), not `volatile`.
__device__ __forceinline__ int flag_load(const int* p) {
    int v;
    asm volatile("ld.relaxed.cta.shared.s32 %0, [%1];" : "=r"(v) : "r"((unsigned)__cvta_generic_to_shared(p)) : "memory");
    return v;
}
__device__ __forceinline__ void flag_store(int* p, int v) {
    asm volatile("st.relaxed.cta.shared.s32 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(p)), "r"(v) : "memory");
}
struct RegroupCtl {  // one per CTA, in shared memory
    int live;        // lanes of the CTA that hold a trajectory (kept by the lanes that run out of work: atomicSub)
    int request;     // != 0: a regrouping is asked for; every running warp comes to the meeting at its next checkpoint
    int dry;         // != 0: the launch's work counter has been seen dry (it stays dry)
    int sorts;       // exchanges so far that dealt the trajectories sorted by remaining time
    int cnt[32];     // per warp, during a meeting: its live lanes
};
struct RegroupPlan {
    int action;    // RG_EXIT, RG_EXCHANGE or RG_NOTHING_TO_FREE
    int sort;      // RG_EXCHANGE: deal the trajectories sorted by remaining time (else: compacted, order kept)
    int slot;      // RG_EXCHANGE, live lanes: where this lane's state goes in the compacted order
    int total;     // live trajectories of the CTA
    int w_active;  // in: warps of the CTA that are running; out (regroup_meet): the same after this regrouping
};
constexpr int RG_EXIT = 1, RG_EXCHANGE = 2, RG_NOTHING_TO_FREE = 3;

#ifdef BACON_DRIVE_TRACE  // (diagnosis: the regroupings of CTA 0 on a time axis; read back with bacon_debug_trace, rhs_builtin.cu)
static __device__ unsigned long long g_trace[3 * 4096];
static __device__ unsigned int g_trace_n;
__device__ __forceinline__ void trace_put(unsigned long long a, unsigned long long b) {
    unsigned long long ns;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
    const unsigned k = atomicAdd(&g_trace_n, 1u);
    if (k < 4096) {
        g_trace[3 * k] = ns;
        g_trace[3 * k + 1] = a;
        g_trace[3 * k + 2] = b;
    }
}
#endif

// How many exchanges of a launch deal the trajectories SORTED by remaining time (the others keep the order, which a
// sorted deal leaves sorted): the lanes of a warp then end together, whole warps fall silent one after the other and
// little is left to compact.  (tools/sorted_probe.py: an ensemble sorted by step count beforehand runs 9 % faster than
// in seeded order at 131072 per GPU, with or without regrouping: that is the ceiling of any regrouping.)
#ifndef BACON_REGROUP_SORTS
#define BACON_REGROUP_SORTS 1
#endif
// Regroup when this many lanes of the CTA hold nothing: a meeting costs every warp up to CHECK_EVERY_DRY attempts of
// waiting, an empty lane costs its share of a warp instruction; 64 = two warps freed per meeting.
#ifndef BACON_REGROUP_AT
#define BACON_REGROUP_AT 96
#endif

// A lane ran out of work: count it out, and ask for a regrouping once enough lanes of the CTA hold nothing (or its
// live trajectories fit in one warp fewer and that is all there is to gain).
static __device__ __noinline__ void regroup_count_out(RegroupCtl* ctl, int w_active) {
    const int left = atomicSub(&ctl->live, 1) - 1;
    const int empty = 32 * w_active - left;
#ifdef BACON_NO_REGROUP  // (A/B switch for measurements: the warps meet once, when the CTA has nothing left)
    if (left == 0) flag_store(&ctl->request, 1);
#else
    if (empty >= BACON_REGROUP_AT || (empty >= 32 && w_active <= 4) || left == 0) flag_store(&ctl->request, 1);
#endif
}

// Has the launch's work counter handed out everything?  One shared-memory load once somebody in the CTA has seen it.
// (Not inlined: inlined, the persistent loop carries four more register moves per pair of attempts.)
static __device__ __noinline__ bool counter_dry(RegroupCtl* ctl, const unsigned long long* counter, unsigned long long n_rest) {
    if (flag_load(&ctl->dry) != 0) return true;
    if (*(volatile const unsigned long long*)counter < n_rest) return false;
    flag_store(&ctl->dry, 1);
    return true;
}

// First meeting of a regrouping.  The vote is the warp's meeting point: lanes that hold nothing have been waiting HERE
// (blocked, not spinning) since they ran out of work; the live lanes of the warp come when they find ctl->request set
// at a checkpoint.  Then the CTA's running warps meet at the named barrier (every one of them arrives within
// CHECK_EVERY_DRY attempts) and the exchange is planned on exact counts.
static __device__ __noinline__ void regroup_plan(RegroupPlan* p, RegroupCtl* ctl, bool live) {
    const int lane = (int)(threadIdx.x & 31), warp = (int)(threadIdx.x >> 5);
    const int w_active = p->w_active;
    volatile int* cnt = ctl->cnt;
    const unsigned live_mask = __ballot_sync(FULL_MASK, live);
#ifdef BACON_DRIVE_TRACE
    if (blockIdx.x == 0 && lane == 0) trace_put(1001, warp);
#endif
    if (lane == 0) {
        cnt[warp] = __popc(live_mask);
        flag_store(&ctl->request, 1);  // (it is: set again so that a meeting can never be one-sided)
    }
    named_barrier(1, w_active * 32);
    const int c = lane < w_active ? cnt[lane] : 0;
    int incl = c;
#pragma unroll
    for (int m = 1; m < 32; m <<= 1) {
        const int up = __shfl_up_sync(FULL_MASK, incl, m);
        if (lane >= m) incl += up;
    }
    const int total = __shfl_sync(FULL_MASK, incl, 31);
    const int my_off = __shfl_sync(FULL_MASK, incl - c, warp);
#ifdef BACON_DRIVE_TRACE
    if (blockIdx.x == 0 && threadIdx.x == 0) trace_put(w_active, total);
#endif
    p->total = total;
    p->slot = my_off + __popc(live_mask & lanemask_lt());
    p->action = total == 0 ? RG_EXIT : (((total + 31) >> 5) < w_active ? RG_EXCHANGE : RG_NOTHING_TO_FREE);
    p->sort = *(volatile int*)&ctl->sorts < BACON_REGROUP_SORTS;
}
// One more meeting of the running warps (all 32 lanes of a warp come together, whichever way they came)
static __device__ __noinline__ void regroup_sync(int w_active) {
    __syncwarp();
    named_barrier(1, w_active * 32);
}
// Rank of a trajectory among the CTA's `total` live ones, longest remaining time first (ties: compacted order).
// keys[] = remaining times as float bit patterns (monotone for the non-negative values that matter), in compacted
// order, padded with INT_MIN to a multiple of 4.
static __device__ __noinline__ int regroup_rank(const int* keys, int total, int key, int slot) {
    int r = 0;
    const int4* k4 = reinterpret_cast<const int4*>(keys);
    for (int j = 0; j < total; j += 4) {
        const int4 k = k4[j >> 2];
        r += (k.x > key || (k.x == key && j < slot)) ? 1 : 0;
        r += (k.y > key || (k.y == key && j + 1 < slot)) ? 1 : 0;
        r += (k.z > key || (k.z == key && j + 2 < slot)) ? 1 : 0;
        r += (k.w > key || (k.w == key && j + 3 < slot)) ? 1 : 0;
    }
    return r;
}
// Second meeting of a regrouping: the live lanes' states are in shared memory (or there was nothing to exchange).
// Afterwards the lowest ceil(total / 32) warps go on: lane l of warp w owns slot 32 w + l.
static __device__ __noinline__ void regroup_meet(RegroupPlan* p, RegroupCtl* ctl) {
    if (threadIdx.x == 0) {  // (nobody is stepping: every running warp is between the two meetings)
        flag_store(&ctl->request, 0);
        *(volatile int*)&ctl->live = p->total;
        if (p->action == RG_EXCHANGE && p->sort) *(volatile int*)&ctl->sorts = *(volatile int*)&ctl->sorts + 1;
    }
    __syncwarp();
    named_barrier(1, p->w_active * 32);
    if (p->action == RG_EXCHANGE) p->w_active = (p->total + 31) >> 5;
}
// A lane that holds nothing waits for the regrouping that gives it a trajectory: returns its slot in the exchange
// buffer, or -1 when its warp is freed (or the CTA is finished).
static __device__ __noinline__ int idle_until_regrouped(RegroupPlan* p, RegroupCtl* ctl) {
    for (;;) {
        regroup_plan(p, ctl, false);  // (blocks until the next regrouping)
        if (p->action == RG_EXIT) return -1;
        if (p->action == RG_EXCHANGE && p->sort) regroup_sync(p->w_active);  // (the live lanes post their keys)
        regroup_meet(p, ctl);
        if (p->action != RG_EXCHANGE) continue;
        const int warp = (int)(threadIdx.x >> 5), slot = (int)threadIdx.x;  // = 32 warp + lane
        if (warp >= p->w_active) return -1;
        if (slot < p->total) return slot;
    }
}

// steppers whose next checkpoint can be brought forward (RkFastStepper::hurry)
template <class S, class = void> struct StepperRetimes { static constexpr bool value = false; };
template <class S> struct StepperRetimes<S, decltype(void(&S::hurry))> { static constexpr bool value = true; };
#ifndef BACON_CHECK_EVERY_DRY
#define BACON_CHECK_EVERY_DRY 16
#endif
constexpr unsigned CHECK_EVERY_DRY = BACON_CHECK_EVERY_DRY;  // attempts between checkpoints once the work counter is dry

// Terminal event (bacon_ivp_options::event_w; NOT in the reference): the kernels instantiated with EVENT watch
// g(y) = w . y - c on the points a trajectory yields, knot 0 being its initial condition.  When g changes sign between the
// last knot and the point just yielded (same rule as the events query on stored paths, path_query.cuh:
// event_crossing), the crossing is located on the cubic Hermite interpolant of that interval with the right-hand
// side's own slopes (hermite_root: the very function the path query uses, so `stop at the first event` and `first
// event of the stored path` agree bit for bit in the strict build), and the trajectory retires there:
// status BACON_STOPPED_AT_EVENT, t_end = t*, y_end = y(t*); the point that crossed is not yielded.  D + 2 more doubles per
// lane (the last knot), D FMAs and a compare per yielded point: a separate instantiation, so the plain kernels
// pay nothing; event kernels run as 128-lane CTAs without regrouping (the last knot would have to migrate too).
// The same instantiation serves the other optional input, the restart record (per-trajectory start time and first dt:
// Stepper::apply_restart): even two more launch constants read inside the plain kernels' refill path cost the fast RK
// loop a uniform register pair, i.e. one constant re-load per attempt (tools/sass_count.py).
template <int D, bool ON> struct EventWatch {
    __device__ __forceinline__ void begin(const bacon_launch_args&, double, const double (&)[D]) {}
};
template <int D> struct EventWatch<D, true> {
    double gp, tp, yp[D];  // the last knot: g, time, state
    __device__ __forceinline__ double g_of(const bacon_launch_args& a, const double (&y)[D]) const {
        double s = a.ev_w[0] * y[0];
#pragma unroll
        for (int d = 1; d < D; ++d) s += a.ev_w[d] * y[d];
        return s - a.ev_c;
    }
    __device__ __forceinline__ void begin(const bacon_launch_args& a, double t0, const double (&y0)[D]) {
        tp = t0;
#pragma unroll
        for (int d = 0; d < D; ++d) yp[d] = y0[d];
        gp = g_of(a, y0);
    }
    // the point (t, y) was yielded: true = g crossed on the way here (the knot is kept); else (t, y) is the new knot
    __device__ __forceinline__ bool crossed(const bacon_launch_args& a, double t, const double (&y)[D], double& g_new) {
        if (!a.ev_on) return false;  // (this launch only carries a restart record)
        g_new = g_of(a, y);
        if (event_crossing(gp, g_new, a.ev_direction)) return true;
        gp = g_new;
        tp = t;
#pragma unroll
        for (int d = 0; d < D; ++d) yp[d] = y[d];
        return false;
    }
    // the event point between the knot and (tb, yb): same operations, same order as locate_event (path_query.cuh)
    template <class Rhs, int P>
    __device__ __noinline__ void locate(const bacon_launch_args& a, const double (&p)[P], double tb, const double (&yb)[D],
                                        double gb, double& te, double (&ye)[D]) const {
        double fa[D], fb[D];
        const Rhs rhs{};
        rhs(tp, yp, p, fa);
        rhs(tb, yb, p, fb);
        const double h = tb - tp;
        double da = a.ev_w[0] * fa[0], db = a.ev_w[0] * fb[0];
#pragma unroll
        for (int d = 1; d < D; ++d) {
            da += a.ev_w[d] * fa[d];
            db += a.ev_w[d] * fb[d];
        }
        const double th = hermite_root(gp, gb, h * da, h * db);
        hermite_eval<D>(th, h, yp, yb, fa, fb, ye);
        te = tp + th * h;
    }
};

// The kernel.  One CTA of BLOCK lanes; a lane integrates one trajectory at a time.
//
// First deal (static): bundle j = 32 consecutive trajectories; warp w of CTA b starts on bundle w * gridDim.x + b, so a
// small ensemble spreads over all CTAs (and, inside a CTA, over its lowest warps = evenly over the SM's four
// sub-partitions) and a large one starts without a single atomic.  After that a lane whose trajectory retires
// (Done / Failure) takes the next index from the launch's work counter.  No vote and no liveness test in the loop: the
// lane leaves the common path on its own (the branch is inside attempt()), stores its record, re-arms and rejoins its
// warp at the next attempt.
//
// End of the ensemble (steppers that migrate).  Once the counter is dry, lanes that retire have nothing to take, and a
// warp instruction occupies the FP64 pipe for the same time with 1 active lane as with 32: left alone, every warp
// would run until its LONGEST lane ends with ever fewer lanes active (measured in round 1: a fixed 1.75 ms per launch,
// half a trajectory time, whatever the ensemble size; and for an ensemble that fits the grid once, every warp pays
// max-of-32 instead of the mean step count: +12 % on Lorenz).  So the CTA REGROUPS.  A lane that runs out of work counts
// itself out of the CTA's live total (one shared-memory atomic per trajectory) and waits, blocked at its warp's vote;
// once BACON_REGROUP_AT lanes of the CTA hold nothing it raises a request flag.  The lanes of a warp pass a checkpoint
// together every CHECK_EVERY attempts (the tick axis of the stepper: nothing is polled per attempt; every
// CHECK_EVERY_DRY once the counter is dry, so that a meeting gathers quickly): there they read the flag — one
// shared-memory load — and if it is set all running warps meet at a named barrier, the live lanes write their stepper
// state (Stepper::save, STATE_DOUBLES + 1 words) to shared memory in compacted order, the lowest ceil(live / 32) warps
// read them back (Stepper::load) and the freed warps exit.  Between two regroupings the warps run freely (no
// barrier), so a sub-partition that holds one warp fewer simply runs its warps faster.  What the regrouping buys is
// measured by tools/tau_probe.py: the time of one attempt of a warp grows with the warps resident on its
// sub-partition (0.27 us alone, 0.47 with three, 0.89 with six: the FP64 pipe is saturated from four), so every
// half-empty warp that is folded away speeds up all the others; what it costs is the wait of the early arrivals at a
// meeting, at most CHECK_EVERY_DRY attempts (profiles/r02_strong_scaling.md).  With one CTA per SM (BLOCK = all resident
// lanes of the SM) this is an SM-wide re-deal; per-SM work is even by the law of large numbers (886 trajectories per SM
// at 131072 per GPU: 0.3 % spread).  A trajectory's numbers do not depend on where it ran
// (tests/test_gpu_rk.py::test_regrouping_and_dense_output_do_not_change_a_trajectory: bitwise).
template <class Stepper, bool HIST, int BLOCK, int MINB, bool EVENT = false>
__global__ void __launch_bounds__(BLOCK, MINB) ensemble_kernel(const __grid_constant__ bacon_launch_args a) {
    constexpr int D = Stepper::D;
    constexpr bool MIGRATE = StepperMigrates<Stepper, BLOCK>::value;
    static_assert(!(EVENT && MIGRATE), "event kernels do not regroup");
    constexpr int NW = BLOCK / 32;
    static_assert(BLOCK % 32 == 0 && NW <= 32, "a CTA is at most 32 warps");
    using Codec = StepperCodec<Stepper>;

    extern __shared__ double xch[];  // MIGRATE: [STATE_DOUBLES + 1][BLOCK] exchange buffer of a regrouping
    __shared__ unsigned long long wq_words[NW];
    __shared__ RegroupCtl ctl;  // MIGRATE
    __shared__ __align__(16) int rg_keys[MIGRATE ? BLOCK + 4 : 4];  // MIGRATE: sort keys of an exchange

    Stepper s(a);
    HistStage<D, HIST> hist(a);
    [[maybe_unused]] EventWatch<D, EVENT> ev;
    const unsigned long long n = a.n;
    const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned long long lanes = (unsigned long long)gridDim.x * BLOCK;
    const unsigned long long n_rest = n > lanes ? n - lanes : 0;  // what the work counter hands out

    // ---- static first deal
    const unsigned long long bundles = ((n < lanes ? n : lanes) + 31) >> 5;
    int w_active = 0;  // warps of this CTA that hold a bundle
    if (bundles > blockIdx.x) {
        const unsigned long long mine = (bundles - blockIdx.x + gridDim.x - 1) / gridDim.x;
        w_active = mine < (unsigned long long)NW ? (int)mine : NW;
    }
    unsigned long long idx = ((unsigned long long)warp * gridDim.x + blockIdx.x) * 32 + lane;
    bool live = (int)warp < w_active && idx < n;
    if constexpr (MIGRATE) {
        if (threadIdx.x == 0) {
            ctl.request = 0;
            ctl.dry = 0;
            ctl.sorts = 0;
            // live lanes of the CTA after the first deal: its bundles are full except the ensemble's last one
            const unsigned long long last = bundles - 1;  // (bundles >= 1: n >= 1)
            int cta_live = 32 * w_active;
            if (w_active > 0 && last % gridDim.x == blockIdx.x && n < lanes) cta_live -= (int)(32 * bundles - n);
            ctl.live = cta_live;
        }
        __syncthreads();
    }
    if ((int)warp >= w_active) return;

#ifndef BACON_WQ_DIV
#define BACON_WQ_DIV 1  // block size = remaining / (BACON_WQ_DIV x warps of the grid), clamped to [1, 32]
#endif
    WarpQueue wq{&wq_words[warp], a.work_counter, n, lanes, (unsigned long long)BACON_WQ_DIV * NW * gridDim.x};
    if (HIST && lane == 0) *wq.word = (1ull << 8) | 1ull;  // the queue starts as an exhausted block
    __syncwarp();

    if (live) {
        s.reset(a, idx, true);
        hist.begin(idx);
        if constexpr (EVENT) {
            s.apply_restart(a, idx);
            ev.begin(a, s.t, s.end_y());
        }
    }
    auto step = [&]() -> bool {  // one IVPIterator::next; true = this lane leaves the loop (nothing to run, or its warp reports in)
        bool yielded = false;
        const uint32_t n_acc_before = (HIST || EVENT) ? Codec::acc_running(s) : 0u;
        const int raw = s.attempt(yielded);
        if constexpr (EVENT) {
            double g_new;
            if (yielded && ev.crossed(a, s.out_t(), s.out_y(), g_new)) {  // rare: the trajectory ends at the event
                double te, ye[D];
                ev.template locate<typename Stepper::RhsT>(a, s.p, s.out_t(), s.out_y(), g_new, te, ye);
                hist.retire(idx, n_acc_before);
                store_result<D>(a.out, n, idx, ye, te, s.dt, BACON_STOPPED_AT_EVENT, n_acc_before, s.n_rej, s.n_rhs());
                idx = HIST ? wq.fetch() : lanes + atomicAdd(a.work_counter, 1ull);
                if (idx >= n) {
                    live = false;
                    return true;
                }
                s.reset(a, idx, true);
                hist.begin(idx);
                s.apply_restart(a, idx);
                ev.begin(a, s.t, s.end_y());
                return false;
            }
        }
        hist.push(yielded, n_acc_before, s.out_t(), s.out_y());
        if (raw != RAW_RUNNING) {  // rare
            if (raw == RAW_CHECKPOINT) {
                if constexpr (MIGRATE) {
                    if (HIST ? wq.dry() : counter_dry(&ctl, a.work_counter, n_rest)) {
                        if constexpr (StepperRetimes<Stepper>::value) s.hurry(CHECK_EVERY_DRY);
                        if (flag_load(&ctl.request) != 0) return true;  // the warp goes to the meeting
                    }
                }
            } else {
                const uint32_t n_acc = Codec::acc(s, raw);
                hist.retire(idx, n_acc);
                int st = Codec::status(s, raw);
                if (HIST && st == BACON_OK && n_acc > (uint32_t)a.cfg.history_capacity) st = BACON_E_HISTORY_OVERFLOW;
                store_result<D>(a.out, n, idx, s.end_y(), s.t, s.dt, st, n_acc, s.n_rej, s.n_rhs());
                // (final state only: which trajectory a lane runs next does not matter — one global atomicAdd)
                idx = HIST ? wq.fetch() : lanes + atomicAdd(a.work_counter, 1ull);
                if (idx >= n) {
                    live = false;
                    if constexpr (MIGRATE) regroup_count_out(&ctl, w_active);
                    return true;
                }
                s.reset(a, idx, true);
                hist.begin(idx);
                if constexpr (EVENT) {
            s.apply_restart(a, idx);
            ev.begin(a, s.t, s.end_y());
        }
            }
        }
        return false;
    };
    // A regrouping is asked for (live lanes: they are at a checkpoint), or this lane holds nothing; true = the lane
    // leaves the kernel.  (No loop in here: see above.)
    auto report = [&]() -> bool {
        if constexpr (!MIGRATE) {
            return true;  // (the lane has nothing left to run)
        } else {
            constexpr int W = Stepper::STATE_DOUBLES + 1;
            RegroupPlan plan;
            plan.w_active = w_active;
            int slot = -1;
            if (live) {
                regroup_plan(&plan, &ctl, true);
                if (plan.action == RG_EXIT) return true;
                if (plan.action == RG_EXCHANGE) {
                    int dst = plan.slot;
                    if (plan.sort) {
                        const int key = __float_as_int((float)s.remaining());
                        rg_keys[plan.slot] = key;
                        if (plan.slot == plan.total - 1)  // pad to a multiple of 4 (no loop in this function: see RegroupCtl)
                            rg_keys[plan.total] = rg_keys[plan.total + 1] = rg_keys[plan.total + 2] = (int)0x80000000;
                        regroup_sync(plan.w_active);
                        dst = regroup_rank(rg_keys, plan.total, key, plan.slot);
                    }
                    double st[W];
                    s.save(st);
                    st[W - 1] = __longlong_as_double((long long)idx);
#pragma unroll
                    for (int w = 0; w < W; ++w) xch[w * BLOCK + dst] = st[w];
                }
                regroup_meet(&plan, &ctl);
                if (plan.action != RG_EXCHANGE) return false;
                w_active = plan.w_active;
                if ((int)warp >= w_active) return true;  // freed
                live = (int)threadIdx.x < plan.total;
                if (live) slot = (int)threadIdx.x;
            }
            if (!live) {
                slot = idle_until_regrouped(&plan, &ctl);
                if (slot < 0) return true;
                w_active = plan.w_active;
                live = true;
            }
            double st[W];
#pragma unroll
            for (int w = 0; w < W; ++w) st[w] = xch[w * BLOCK + slot];
            idx = (unsigned long long)__double_as_longlong(st[W - 1]);
            s.load(st);
            if constexpr (StepperRetimes<Stepper>::value) s.hurry(CHECK_EVERY_DRY);
            hist.begin(idx);
            return false;
        }
    };
    if (!live && report()) return;  // (the ragged end of the last bundle)
    for (;;) {
        if (step() && report()) return;
        // steppers with a short body (RkFastStepper) take two per iteration: the copies that carry dt and the call
        // counter round the loop, and the loop's own branch, are then paid once per pair (31 -> 28.5 non-FP64
        // instructions per attempt)
        if constexpr (StepperUnrolls<Stepper>::value) {
            if (step() && report()) return;
        }
    }
}

}  // namespace bacon
