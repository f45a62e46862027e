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
// idles lanes once the whole ensemble has been handed out.  What is left of the
// tail is suspended and re-dealt by a second kernel (below).
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

template <class S, class = void> struct StepperUnrolls { static constexpr bool value = false; };
template <class S> struct StepperUnrolls<S, decltype(void(S::UNROLL_DRIVER))> { static constexpr bool value = S::UNROLL_DRIVER; };

constexpr int RAW_RUNNING = -1;     // attempt(): the trajectory goes on
constexpr int RAW_CHECKPOINT = -2;  // attempt(): nothing was computed, the whole warp reports in (RkFastStepper::tick)

// does the stepper support suspending a trajectory and resuming it on another lane (tail compaction)?
template <class S, class = void> struct StepperMigrates { static constexpr bool value = false; };
#ifndef BACON_NO_MIGRATE  // (A/B switch for measurements)
template <class S> struct StepperMigrates<S, decltype(void(S::STATE_DOUBLES))> {
    static constexpr bool value = (S::STATE_DOUBLES + 1) * 8 * (2 * ENSEMBLE_BLOCK) <= 40 * 1024;  // the tail CTA's exchange buffer
};
#endif

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
struct WarpQueue {
    unsigned long long* word;       // shared memory
    unsigned long long* global;     // the launch's work counter
    unsigned long long n;
    unsigned long long per_block;   // BACON_WQ_DIV x warps of the grid: remaining / per_block = next block size

    __device__ __noinline__ unsigned long long fetch() {
        const unsigned long long s = atomicAdd(word, 1ull);
        const unsigned used = (unsigned)(s & 0xff), size = (unsigned)((s >> 8) & 0xff);
        if (used < size) return (s >> 16) + used;
        if (used > size) return atomicAdd(global, 1ull);  // a refill is in flight (rare): take a single index instead
        // used == size: this lane refills
        const unsigned long long seen = *(volatile unsigned long long*)global;
        const unsigned long long rem = seen < n ? n - seen : 0;
        unsigned long long b = rem / per_block;
        b = b < 1 ? 1 : (b > 32 ? 32 : b);
        const unsigned long long nb = atomicAdd(global, b);
        // (a block may reach past n, or start there when the counter is dry: callers test idx < n, and every later
        // fetch from such a block returns an index >= n as well)
        atomicExch(word, (nb << 16) | (b << 8) | 1ull);
        return nb;
    }
    // nothing left to start in this warp's block and nothing left in the global counter
    __device__ __forceinline__ bool dry() const {
        const unsigned long long v = *(volatile unsigned long long*)word;
        return (unsigned)(v & 0xff) >= (unsigned)((v >> 8) & 0xff) && *(volatile unsigned long long*)global >= n;
    }
};

// Main kernel.  No vote and no liveness test in the loop: a lane whose trajectory ends leaves the common path on its own
// (the branch is inside attempt()), stores its record, takes the next trajectory index from its warp's block (WarpQueue)
// and rejoins its warp at the next attempt.  A lane that finds the queue dry is done.
//
// Tail (steppers that migrate, `tail` != nullptr).  The lanes decohere over the run, so when the counter runs dry the
// remaining work per lane is spread evenly between nothing and a whole trajectory; left alone, each warp would run
// until its LONGEST lane ends with ever fewer lanes active — and a warp instruction occupies the FP64 pipe for the same
// time with 1 active lane as with 32: the tail costs half a trajectory time whatever the ensemble size (measured:
// 1.75 ms of a 31.6 ms launch, profiles/r01g_tail.md).  So at its next checkpoint after the counter ran dry (every
// CHECK_EVERY attempts the whole warp reports in: nothing is polled per attempt) a warp SUSPENDS: every lane writes
// its trajectory's state to its slot of `tail` and the kernel ends; ensemble_tail_kernel re-deals and finishes them.
// Slot layout: tail[w][grid*128] doubles, w < STATE_DOUBLES + 1 (the last word is the trajectory index, -1 = none).
template <class Stepper, bool HIST, int MINB>
__global__ void __launch_bounds__(ENSEMBLE_BLOCK, MINB)
    ensemble_kernel(const __grid_constant__ bacon_launch_args a, double* __restrict__ tail) {
    constexpr int D = Stepper::D;
    constexpr bool MIGRATE = StepperMigrates<Stepper>::value;
    using Codec = StepperCodec<Stepper>;

    Stepper s(a);
    HistStage<D, HIST> hist(a);
    const unsigned long long n = a.n;
    const size_t lanes = (size_t)gridDim.x * ENSEMBLE_BLOCK, me = (size_t)blockIdx.x * ENSEMBLE_BLOCK + threadIdx.x;

    // a lane leaves the kernel: its slot of `tail` says what it leaves behind (nothing, or a suspended trajectory)
    auto leave = [&](bool suspended, unsigned long long idx) {
        if constexpr (MIGRATE) {
            if (tail) {
                constexpr int W = Stepper::STATE_DOUBLES + 1;
                if (suspended) {
                    double st[W];
                    s.save(st);
#pragma unroll
                    for (int w = 0; w < W - 1; ++w) tail[(size_t)w * lanes + me] = st[w];
                }
                tail[(size_t)(W - 1) * lanes + me] = __longlong_as_double(suspended ? (long long)idx : -1ll);
            }
        }
    };

    __shared__ unsigned long long wq_words[ENSEMBLE_BLOCK / 32];
#ifndef BACON_WQ_DIV
#define BACON_WQ_DIV 1  // block size = remaining / (BACON_WQ_DIV x warps of the grid), clamped to [1, 32]
#endif
    WarpQueue wq{&wq_words[threadIdx.x >> 5], a.work_counter, n, (unsigned long long)BACON_WQ_DIV * (ENSEMBLE_BLOCK / 32) * gridDim.x};
    // the first 32 trajectories of the warp: one aggregated fetch; the queue starts as an exhausted block
    unsigned long long idx = warp_fetch(a.work_counter, true);
    if ((threadIdx.x & 31) == 0) *wq.word = (1ull << 8) | 1ull;
    __syncwarp();

    if (idx >= n) {
        leave(false, 0);
        return;
    }
    s.reset(a, idx, true);
    hist.begin(idx);
    auto step = [&]() -> bool {  // one IVPIterator::next; true = this lane leaves the kernel
        bool yielded = false;
        const uint32_t n_acc_before = HIST ? Codec::acc_running(s) : 0u;
        const int raw = s.attempt(yielded);
        hist.push(yielded, n_acc_before, s.out_t(), s.out_y());
        if (raw != RAW_RUNNING) {  // rare
            if (raw == RAW_CHECKPOINT) {
                if (MIGRATE && tail && (HIST ? wq.dry() : *(volatile unsigned long long*)a.work_counter >= n)) {
                    leave(true, idx);  // suspend
                    return true;
                }
            } else {
                const uint32_t n_acc = Codec::acc(s, raw);
                hist.retire(idx, n_acc);
                int st = Codec::status(s, raw);
                if (HIST && st == BACON_OK && n_acc > (uint32_t)a.cfg.history_capacity) st = BACON_E_HISTORY_OVERFLOW;
                store_result<D>(a.out, n, idx, s.end_y(), s.t, s.dt, st, n_acc, s.n_rej, s.n_rhs());
                // (final state only: which trajectory a lane runs next does not matter — one global atomicAdd)
                idx = HIST ? wq.fetch() : atomicAdd(a.work_counter, 1ull);
                if (idx >= n) {
                    leave(false, 0);
                    return true;
                }
                s.reset(a, idx, true);
                hist.begin(idx);
            }
        }
        return false;
    };
    for (;;) {
        if (step()) return;
        // steppers with a short body (RkFastStepper) take two per iteration: the copies that carry dt and the call
        // counter round the loop, and the loop's own branch, are then paid once per pair (31 -> 28.5 non-FP64
        // instructions per attempt)
        if constexpr (StepperUnrolls<Stepper>::value) {
            if (step()) return;
        }
    }
}

// Which eighth of the CTA's sorted trajectories a warp takes in the tail.  Measured with tools/smsp_probe.cu on B200:
// a warp runs on sub-partition %warpid % 4 (two warps with the same value share one FP64 pipe), a 256-thread CTA holds
// hardware warp slots 8k..8k+7 (k = %warpid / 8 = the CTA's slot on its SM; its warps w and w+4 share a
// sub-partition), and the hardware already rotates which warp of the CTA starts on sub-partition 0.  The two warps of
// a CTA on sub-partition j take the octiles o and 7 - o, o = (j + k) mod 4: every sub-partition holds the same amount
// of work whatever the number of resident CTAs, and its warps retire at evenly spread times.
constexpr int TAIL_BLOCK = 2 * ENSEMBLE_BLOCK;
__device__ __forceinline__ int tail_octile() {
    unsigned hw;
    asm volatile("mov.u32 %0, %%warpid;" : "=r"(hw));
    const int j = hw & 3, second = (hw >> 2) & 1, k = hw >> 3;
    const int o = (j + k) & 3;
    return second ? 7 - o : o;
}

// Tail kernel (one CTA of 256 lanes per two CTAs of the main kernel): the CTA sorts the 256 suspended trajectories of its
// slots by remaining time and deals them so that each warp holds one octile — warps retire one after the other and
// the sub-partitions thin out — then runs them to their end, no refills.
template <class Stepper, bool HIST, int MINB>
__global__ void __launch_bounds__(TAIL_BLOCK, (MINB + 1) / 2)
    ensemble_tail_kernel(const __grid_constant__ bacon_launch_args a, const double* __restrict__ tail, unsigned long long lanes) {
    constexpr int D = Stepper::D;
    using Codec = StepperCodec<Stepper>;
    constexpr int W = Stepper::STATE_DOUBLES + 1;
    __shared__ double slots[W][TAIL_BLOCK];
    __shared__ int keys[TAIL_BLOCK];

    Stepper s(a);
    HistStage<D, HIST> hist(a);
    const size_t g = (size_t)blockIdx.x * TAIL_BLOCK + threadIdx.x;  // slot written by lane g of the main kernel
    const int me = threadIdx.x;

    double st[W];
    st[W - 1] = g < lanes ? tail[(size_t)(W - 1) * lanes + g] : __longlong_as_double(-1ll);
    const bool had = __double_as_longlong(st[W - 1]) >= 0;
#pragma unroll
    for (int w = 0; w < W - 1; ++w) st[w] = had ? tail[(size_t)w * lanes + g] : 0.0;
    // sort key: the float's bit pattern as an integer — monotone for the non-negative values that matter, and a total
    // order (with the index as tie-break) whatever the value, so `rank` is always a permutation; empty slots sort low
    int key = -1;
    if (had) {
        s.load(st);
        key = __float_as_int((float)s.remaining());
    }
    keys[me] = key;
    __syncthreads();
    int rank = 0;
    for (int j = 0; j < TAIL_BLOCK; ++j) {
        const int kj = keys[j];
        rank += (kj < key || (kj == key && j < me)) ? 1 : 0;
    }
#pragma unroll
    for (int w = 0; w < W; ++w) slots[w][rank] = st[w];
    __syncthreads();
    const int src = tail_octile() * 32 + (me & 31);
#pragma unroll
    for (int w = 0; w < W; ++w) st[w] = slots[w][src];
    const long long moved = __double_as_longlong(st[W - 1]);
    if (moved < 0) return;
    const unsigned long long idx = (unsigned long long)moved;
    s.load(st);
    hist.begin(idx);

    for (;;) {
        bool yielded = false;
        const uint32_t n_acc_before = HIST ? Codec::acc_running(s) : 0u;
        const int raw = s.attempt(yielded);
        hist.push(yielded, n_acc_before, s.out_t(), s.out_y());
        if (raw >= 0) {
            const uint32_t n_acc = Codec::acc(s, raw);
            hist.retire(idx, n_acc);
            int stt = Codec::status(s, raw);
            if (HIST && stt == BACON_OK && n_acc > (uint32_t)a.cfg.history_capacity) stt = BACON_E_HISTORY_OVERFLOW;
            store_result<D>(a.out, a.n, idx, s.end_y(), s.t, s.dt, stt, n_acc, s.n_rej, s.n_rhs());
            return;
        }
    }
}

}  // namespace bacon
