/* bacon_ivp.h — C ABI of the B200 ensemble IVP engine (libbacon_ivp.so).
 *
 * This is the drop-in boundary for ONE path of aftix/bacon (bacon-sci 0.16.2):
 * the solvers of `bacon_sci::ivp` (RungeKutta45, RungeKutta23, BDF6/BDF2, and the
 * callers either side of them: Adams5/Adams3, Euler), batched over N independent
 * trajectories.  The reference has no FFI of its
 * own (it is pure safe Rust); every entry point below names the reference
 * interface it replaces (file:line relative to the reference tree).
 *
 * All entry points are `extern "C"`, take plain pointers and sizes, never
 * throw, and never touch torch/nalgebra types.  Call-level failures are the
 * return code (0 = success, otherwise a bacon_status); per-trajectory failures
 * go to `status[i]` and never abort the ensemble (the reference aborts the one
 * trajectory it is integrating, src/ivp.rs:232-235).
 */
#ifndef BACON_IVP_H
#define BACON_IVP_H

#ifdef __CUDACC_RTC__ /* NVRTC has no system headers (bacon_rhs_register_source compiles user functors with it) */
typedef signed int int32_t;
typedef unsigned int uint32_t;
typedef long long int64_t;
typedef unsigned long long uint64_t;
typedef unsigned long size_t;
#else
#include <stddef.h>
#include <stdint.h>
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define BACON_IVP_ABI_VERSION 6

/* ---- solver families: src/ivp/rk.rs:561 (RungeKutta45), rk.rs:656
 * (RungeKutta23), src/ivp/bdf.rs:706 (BDF6), bdf.rs:762 (BDF2),
 * src/ivp/adams.rs:633 (Adams5), adams.rs:693 (Adams3), src/ivp.rs:269 (Euler) */
typedef enum bacon_method {
    BACON_RK45 = 0,
    BACON_RK23 = 1,
    BACON_BDF6 = 2,
    BACON_BDF2 = 3,
    BACON_ADAMS5 = 4,
    BACON_ADAMS3 = 5,
    BACON_EULER = 6, /* fixed step: config.dt_max carries the builder's dt, tol/dt_min unused */
    BACON_N_METHODS = 7
} bacon_method;

/* ---- status codes.  1..12 mirror `IVPError` variant by variant
 * (src/ivp.rs:50-76); 0 is `IVPStatus::Done` (src/ivp.rs:24) reached without
 * error; 13.. are conditions the reference either hangs on or cannot hit. -- */
typedef enum bacon_status {
    BACON_OK = 0,
    BACON_E_MISSING_PARAMETERS = 1,  /* ivp.rs:52 */
    BACON_E_USER = 2,                /* ivp.rs:54 */
    BACON_E_TOLERANCE_OOB = 3,       /* ivp.rs:56 */
    BACON_E_TIME_DELTA_OOB = 4,      /* ivp.rs:58 */
    BACON_E_TIME_END_OOB = 5,        /* ivp.rs:60 */
    BACON_E_TIME_START_OOB = 6,      /* ivp.rs:62 */
    BACON_E_FROM_PRIMITIVE = 7,      /* ivp.rs:64 */
    BACON_E_MIN_DT_EXCEEDED = 8,     /* ivp.rs:66 */
    BACON_E_MAX_ITER = 9,            /* ivp.rs:68 */
    BACON_E_SINGULAR = 10,           /* ivp.rs:70 */
    BACON_E_DYNAMIC_ON_STATIC = 11,  /* ivp.rs:72 */
    BACON_E_STATIC_ON_DYNAMIC = 12,  /* ivp.rs:74 */
    BACON_E_NONFINITE = 13,          /* NaN error estimate: reference loops forever (rk.rs:392-422) */
    BACON_E_MAX_ATTEMPTS = 14,       /* hard per-trajectory attempt cap hit */
    BACON_E_HISTORY_OVERFLOW = 15,   /* more accepted points than history_capacity */
    BACON_E_CUDA = 16,               /* CUDA runtime error; see bacon_last_error() */
    BACON_E_BAD_ARGUMENT = 17,       /* NULL pointer, unknown rhs/method, dim mismatch */
    BACON_E_UNSUPPORTED = 18,        /* combination not built (e.g. REF_LITERAL with the Newton BDF)  */
    BACON_STOPPED_AT_EVENT = 19      /* per-trajectory, not an error: the integration stopped at a terminal event
                                        (bacon_ivp_options::event_w); t_end / y_end are the event point */
} bacon_status;

/* ---- semantics: SURVEY.md §8c.  REF_LITERAL reproduces the source as
 * written (transposed Butcher matrix rk.rs:459/370, 1859/4014 rk.rs:499,
 * safety factor 100/100 rk.rs:267, ...); REF_CORRECTED applies D1-D7 only. -- */
#define BACON_SEM_CORRECTED 0
#define BACON_SEM_LITERAL 1

/* ---- flags ------------------------------------------------------------ */
#define BACON_FLAG_STRICT_FP 1u      /* no FMA contraction: bit-comparable with the CPU oracle */
#define BACON_FLAG_SHARED_PARAMS 2u  /* params is [n_params], shared by all trajectories        */
#define BACON_FLAG_BDF_NEWTON 4u     /* BDF: Newton + analytic-Jacobian LU instead of Broyden   */
#define BACON_FLAG_PARAMS_AOS 8u     /* params is [n][n_params] (one contiguous block per trajectory) */
#define BACON_FLAG_ZERO_COPY 16u     /* host entry point, 1 GPU, final state only, every buffer from
                                        bacon_host_alloc: the kernel reads y0/params and writes the
                                        per-trajectory records straight from/to pinned host memory
                                        (no staging copy; ignored when a buffer is not pinned)          */

/* One POD block = everything the reference builder collects before `solve`
 * (rk.rs:60-68 / bdf.rs:57-65), shared by the whole ensemble. */
typedef struct bacon_ivp_config {
    int32_t method;           /* bacon_method                                            */
    int32_t dim;              /* state dimension D; must equal the RHS's DIM             */
    int32_t n_params;         /* per-trajectory parameter count; must equal RHS NPARAM   */
    int32_t semantics;        /* BACON_SEM_*                                             */
    uint32_t flags;           /* BACON_FLAG_*                                            */
    int32_t history_capacity; /* accepted points kept per trajectory; 0 = final only     */
    double dt_min, dt_max;    /* with_minimum_dt / with_maximum_dt  (ivp.rs:171-172)     */
    double tol;               /* with_tolerance                      (ivp.rs:169)        */
    double t_start, t_end;    /* with_initial_time / with_ending_time (ivp.rs:173-174)   */
    uint64_t max_attempts;    /* per-trajectory cap on step() calls; 0 = 2^32-2          */
    double dt_init;           /* first step size; 0 = the reference's (dt_max + dt_min)/2 (rk.rs:315,
                                 bdf.rs:302, adams.rs:297); > 0: bacon_solver_with_initial_dt.  Unused by Euler */
} bacon_ivp_config;

/* Output block.  Any pointer may be NULL (that output is skipped) except
 * y_end and status.  Layouts (n = number of trajectories):
 *   y_end   [dim][n]           SoA, trajectory index fastest
 *   hist    [n][cap][1 + dim]  dense output: one contiguous `Path` per trajectory, one (t, y[0..dim))
 *                              record per accepted point, in the order the reference yields them — the
 *                              memory image of its `Vec<(f64, SVector<f64, D>)>` (ivp.rs:203).  A record is
 *                              32 bytes for dim = 3: the kernel writes it with ONE 256-bit store (a whole
 *                              DRAM sector), no staging.  Device pointers must be 32-byte aligned.
 * For the host entry points these are host pointers, for *_device device
 * pointers. */
typedef struct bacon_ivp_result {
    double* y_end;
    double* t_end;       /* [n] stepper time at Done / failure (IVPStepper::time, ivp.rs:124)  */
    double* dt_end;      /* [n] controller dt at exit (restart record)                         */
    int32_t* status;     /* [n] bacon_status                                                   */
    uint32_t* n_accept;  /* [n] points yielded (`Ok`, ivp.rs:229)                               */
    uint32_t* n_reject;  /* [n] rejected step attempts                                          */
    uint32_t* n_rhs;     /* [n] derivative evaluations                                          */
    double* hist;        /* [n][cap][1 + dim], required when history_capacity > 0               */
    uint32_t* hist_len;  /* [n] records written (<= history_capacity)                           */
    const double* t_start; /* INPUT of the path queries only: [n] per-trajectory start times when the solve was
                              given bacon_ivp_options::t_start_each (knot 0 of every path); NULL = cfg.t_start */
} bacon_ivp_result;

/* Optional inputs of a solve (the *_ex entry points); all-zero = the plain solve.
 *  - restart record: the reference's iterator is resumable in memory (`IVPIterator` keeps the solver, ivp.rs:220-238;
 *    a caller can stop pulling points and go on later).  Across the C ABI the same thing is the per-trajectory record
 *    (t_end, y_end, dt_end) a solve returns: pass y_end as y0, t_end as t_start_each and dt_end as dt_start_each and the
 *    integration goes on where it stopped.  dt is clamped into [dt_min, dt_max] (a finished leg's last, shortened step
 *    can leave dt_end below dt_min).  Exact for the one-step methods (RK: the stepper's whole state is (t, y, dt));
 *    the multistep methods (BDF, Adams) restart with their RK4 warm-up, as after any change of dt.
 *  - terminal event: stop a trajectory where g(y) = w . y - c changes sign between two yielded points (direction +1:
 *    rising only, -1: falling only, 0: both; a right end exactly on the surface counts, a left end does not).  The crossing
 *    is located on the cubic Hermite interpolant of that step exactly like bacon_ivp_locate_events; the trajectory ends
 *    there with status BACON_STOPPED_AT_EVENT, t_end = t*, y_end = y(t*); the step that crossed is not yielded (n_accept
 *    and the history count the points before it).  NOT in the reference. */
typedef struct bacon_ivp_options {
    const double* t_start_each;  /* [n] per-trajectory initial time; NULL = cfg.t_start (host or device as y0)   */
    const double* dt_start_each; /* [n] per-trajectory first dt; NULL = cfg.dt_init / the reference's default    */
    const double* event_w;       /* HOST array of dim doubles (both variants); NULL = no terminal event          */
    double event_c;
    int32_t event_direction;
    int32_t reserved;
} bacon_ivp_options;

/* Launch record filled by the last solve on this thread (timing + totals). */
typedef struct bacon_ivp_launch_info {
    float kernel_ms;        /* CUDA-event time of the ensemble kernel on its stream */
    float h2d_ms, d2h_ms;   /* host entry point only                               */
    int32_t grid, block;    /* launch geometry                                     */
    int32_t regs_per_thread;
    int32_t n_kernels;      /* kernels launched by the call                        */
} bacon_ivp_launch_info;

/* ---- configuration: replaces the builder setters.  Same rules, same order
 * of checks as rk.rs:168-256 (identical in bdf.rs:176-264). ---------------- */
typedef struct bacon_solver bacon_solver; /* opaque builder handle */

int bacon_abi_version(void);

/* IVPSolver::new / new_dyn (ivp.rs:159-163).  Static dimensions are the ones
 * a RHS was compiled for; `dim` is checked against the RHS at solve time. */
bacon_solver* bacon_solver_new(int method, int dim);
/* The reference's two constructors with their `Dimension` check (src/lib.rs:53-76, ivp.rs:159-163, rk.rs:136-166):
 * the solver's type parameter D is either `Const<C>` (dim_type = C >= 1) or `Dyn` (dim_type = BACON_DIM_DYN).
 *   new()          = bacon_solver_new_static(method, dim_type):   Const<C> -> dimension C;  Dyn -> StaticOnDynamic (12)
 *   new_dyn(size)  = bacon_solver_new_dyn(method, dim_type, size): Dyn -> dimension `size`;  Const<C> -> DynamicOnStatic (11)
 * On success *out is the handle (free with bacon_solver_free) and the return code is 0.  Either way the dimension is
 * checked against the right-hand side's DIM when a solve is called (nalgebra would panic at ivp.rs:178). */
#define BACON_DIM_DYN 0
int bacon_solver_new_static(int method, int dim_type, bacon_solver** out);
int bacon_solver_new_dyn(int method, int dim_type, int size, bacon_solver** out);
void bacon_solver_free(bacon_solver*);
int bacon_solver_with_tolerance(bacon_solver*, double tol);            /* rk.rs:168-174 */
int bacon_solver_with_maximum_dt(bacon_solver*, double max);           /* rk.rs:179-192 */
int bacon_solver_with_minimum_dt(bacon_solver*, double min);           /* rk.rs:197-210 */
int bacon_solver_with_initial_time(bacon_solver*, double initial);     /* rk.rs:212-222 */
int bacon_solver_with_ending_time(bacon_solver*, double ending);       /* rk.rs:224-234 */
/* Not in the reference (its first dt is always (dt_max + dt_min)/2, rk.rs:315): the first step size, so that a
 * restart record's dt can be handed back (see bacon_ivp_options).  TimeDeltaOOB unless dt > 0; a value outside
 * [dt_min, dt_max] is clamped at solve time. */
int bacon_solver_with_initial_dt(bacon_solver*, double dt);
int bacon_solver_with_semantics(bacon_solver*, int semantics);
int bacon_solver_with_flags(bacon_solver*, uint32_t flags);
int bacon_solver_with_history(bacon_solver*, int capacity);
int bacon_solver_with_max_attempts(bacon_solver*, uint64_t cap);
/* `solve` front half (rk.rs:249-256): MissingParameters if any of dt_max,
 * dt_min, tolerance, initial time, ending time is unset; fills *out. */
int bacon_solver_config(const bacon_solver*, bacon_ivp_config* out);

/* Stateless re-check of a filled config (bounds only; used by the solve calls). */
int bacon_ivp_validate(const bacon_ivp_config*);

/* ---- right-hand sides: replaces `Derivative` (ivp.rs:34-48) and
 * `with_derivative` (ivp.rs:186).  Built in: "lorenz", "vdp", "robertson",
 * "linear32", "exp", "decay", "quadratic", "cos", "harmonic".  User RHS are
 * CUDA device functors compiled against bacon_ivp_rhs.cuh and registered
 * through bacon_rhs_register. ------------------------- */
struct bacon_launch_args; /* defined in bacon_b200/csrc/ivp_common.cuh; opaque to C callers */
typedef int (*bacon_launch_fn)(struct bacon_launch_args*); /* fills grid/block/regs on return */

struct bacon_path_args;   /* defined in bacon_b200/csrc/path_query.cuh; opaque to C callers */
typedef int (*bacon_path_fn)(struct bacon_path_args*);

typedef struct bacon_rhs_desc {
    const char* name;
    int32_t dim;
    int32_t n_params;
    /* [strict_fp 0/1][method]; NULL = not built for that slot */
    bacon_launch_fn launch[2][BACON_N_METHODS];
    /* [strict_fp 0/1]: the path queries below (sampling, events) for this RHS; NULL = not built */
    bacon_path_fn path_query[2];
    /* the same kernels compiled with the terminal-event test (bacon_ivp_options::event_w); NULL = not built */
    bacon_launch_fn launch_event[2][BACON_N_METHODS];
} bacon_rhs_desc;

int bacon_rhs_register(const bacon_rhs_desc*); /* returns rhs id >= 0, or -bacon_status */
int bacon_rhs_lookup(const char* name);        /* rhs id >= 0, or -1                     */
/* The same plug-in without nvcc on the caller's side: `source` is CUDA C++ text that defines the functor type
 * `type_name` (contract: include/bacon_ivp_rhs.cuh); the library compiles it with NVRTC together with its own kernel
 * headers — the right-hand side is inlined into the stage loops exactly like a built-in — one program per (method,
 * strict/fast, dense output) actually used, cached per device.  Replaces `with_derivative(closure)` (src/ivp.rs:186)
 * for callers that cannot run a CUDA compiler at build time.  Returns the rhs id >= 0, or -bacon_status:
 * BACON_E_USER with the compiler log in bacon_last_error() when the source does not compile, BACON_E_UNSUPPORTED when
 * libnvrtc / the driver are not there. */
int bacon_rhs_register_source(const char* name, const char* type_name, const char* source, int dim, int n_params);
int bacon_rhs_count(void);
int bacon_rhs_info(int rhs_id, const char** name, int* dim, int* n_params);

/* ---- the solve: replaces `solve(data)` + `IVPIterator::collect_vec`
 * (rk.rs:249-343, ivp.rs:209-238) for n trajectories at once. --------------
 * y0 [dim][n] SoA; params [n_params][n] SoA (or [n][n_params] with
 * BACON_FLAG_PARAMS_AOS, or [n_params] with BACON_FLAG_SHARED_PARAMS, or NULL
 * when n_params == 0). */

/* Host buffers: H2D, kernel, D2H on the current CUDA device. */
int bacon_ivp_solve_ensemble(const bacon_ivp_config*, int rhs_id, size_t n, const double* y0,
                             const double* params, const bacon_ivp_result* out);

/* Device buffers already resident in HBM; `stream` is a cudaStream_t (NULL =
 * default stream).  Asynchronous: returns after enqueueing.  */
int bacon_ivp_solve_ensemble_device(const bacon_ivp_config*, int rhs_id, size_t n,
                                    const double* d_y0, const double* d_params,
                                    const bacon_ivp_result* d_out, void* stream);

/* Host buffers, trajectories dealt round-robin (i mod G) over the first
 * n_gpus visible devices of this process (one stream per device). */
int bacon_ivp_solve_ensemble_multi(const bacon_ivp_config*, int rhs_id, size_t n, const double* y0,
                                   const double* params, const bacon_ivp_result* out, int n_gpus);

/* The same three solves with optional inputs (restart record, terminal event): see bacon_ivp_options.  NULL options =
 * the plain call.  n_gpus as in bacon_ivp_solve_ensemble_multi (1 = the current device). */
int bacon_ivp_solve_ensemble_ex(const bacon_ivp_config*, int rhs_id, size_t n, const double* y0, const double* params,
                                const bacon_ivp_options* options, const bacon_ivp_result* out, int n_gpus);
int bacon_ivp_solve_ensemble_device_ex(const bacon_ivp_config*, int rhs_id, size_t n, const double* d_y0,
                                       const double* d_params, const bacon_ivp_options* options,
                                       const bacon_ivp_result* d_out, void* stream);

/* ---- queries on stored paths: the continuous extension.
 * NOT in the reference — its `Path` is the accepted points and nothing between them (src/ivp.rs:203-211); this is the
 * step after the path.  Both calls take what a dense-output solve left behind (the same cfg, y0, params, and its
 * bacon_ivp_result with hist + hist_len; t_end + y_end, when given, close a path whose last points the stepper did not
 * yield) and treat every trajectory's path as the knots (t_start, y0), (t_1, y_1) ... (t_m, y_m).
 * Between two knots the state is the cubic Hermite interpolant through both points with the right-hand side's own
 * slopes f(t_k, y_k), f(t_k+1, y_k+1) (local error O(h^4): the order of every adaptive stepper's propagated solution
 * here or better), evaluated by a second, HBM-bound kernel family compiled per right-hand side (path_query.cuh).
 * Works for every method; BACON_FLAG_STRICT_FP selects the uncontracted build (bit-comparable with the CPU oracle).
 *
 * bacon_ivp_sample_paths*: samples[n][n_times][dim] = the state of every trajectory at `times` (any order; a time
 *   outside a trajectory's path — before t_start, after its last knot, e.g. a trajectory that failed early — gives NaN).
 * bacon_ivp_locate_events*: the zeros of g(y) = w . y - c along every path, in order: where g changes sign between two
 *   knots (direction +1: rising only, -1: falling only, 0: both; an interval whose right knot is exactly zero counts, one
 *   whose left knot is does not), the root of the interpolant's g is located by a bracketed Newton iteration (theta to 1e-15).
 *   events[n][capacity][1 + dim] receives the first `capacity` (t*, y(t*)) records of a trajectory, n_events[n] the
 *   number found (it may exceed capacity).  `w` is a HOST array of dim doubles in both variants.
 * Host variants stage through the current device; *_device take device pointers and enqueue on `stream`. */
int bacon_ivp_sample_paths(const bacon_ivp_config*, int rhs_id, size_t n, const double* y0, const double* params,
                           const bacon_ivp_result* solved, size_t n_times, const double* times, double* samples);
int bacon_ivp_sample_paths_device(const bacon_ivp_config*, int rhs_id, size_t n, const double* d_y0,
                                  const double* d_params, const bacon_ivp_result* d_solved, size_t n_times,
                                  const double* d_times, double* d_samples, void* stream);
int bacon_ivp_locate_events(const bacon_ivp_config*, int rhs_id, size_t n, const double* y0, const double* params,
                            const bacon_ivp_result* solved, const double* w, double c, int direction, int capacity,
                            double* events, uint32_t* n_events);
int bacon_ivp_locate_events_device(const bacon_ivp_config*, int rhs_id, size_t n, const double* d_y0,
                                   const double* d_params, const bacon_ivp_result* d_solved, const double* w, double c,
                                   int direction, int capacity, double* d_events, uint32_t* d_n_events, void* stream);

/* Page-locked host memory for the host entry points (cached by size inside the
 * library; cudaHostAlloc is slow).  Buffers from here make the H2D/D2H legs of
 * bacon_ivp_solve_ensemble asynchronous DMA transfers; pageable buffers work
 * too but are copied through the driver's bounce buffer.  The reference's
 * counterpart is the `Vec` that `collect_vec` returns (ivp.rs:209-211): the
 * caller owns the result storage. */
void* bacon_host_alloc(size_t bytes); /* NULL on failure (see bacon_last_error) */
void bacon_host_free(void* p);        /* NULL is a no-op                         */

int bacon_ivp_last_launch(bacon_ivp_launch_info* out);
const char* bacon_last_error(void); /* thread-local message for the last rc != 0 */
const char* bacon_status_name(int status);

/* ---- measurement helpers (not on the product path) -------------------- */
/* Register-resident DFMA loop on the current device: returns achieved FP64
 * TFLOP/s (the roofline denominator of the RK kernels), <0 on error. */
double bacon_fp64_peak_tflops(int iters, void* stream);
int bacon_device_sm_count(void);

#ifdef __cplusplus
}
#endif
#endif /* BACON_IVP_H */
