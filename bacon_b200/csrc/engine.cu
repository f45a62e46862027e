// engine.cu — the C ABI of libbacon_ivp.so (include/bacon_ivp.h): builder handle,
// RHS registry, and the three solve entry points (device-resident, host buffers,
// host buffers sharded over the GPUs of one box).  No torch, no CPU fallback:
// every solve ends in a CUDA kernel launch or an error code.
//
// Reference interfaces replaced (file:line relative to aftix/bacon):
//   builder setters + validation        src/ivp/rk.rs:168-256 (same in bdf.rs:176-264)
//   `solve(data)` + collect_vec         src/ivp/rk.rs:249-343, src/ivp.rs:209-238
//   `Derivative` / with_derivative      src/ivp.rs:34-48, :186
//   IVPError / IVPStatus                src/ivp.rs:20-28, :50-76
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <cstring>
#include <deque>
#include <functional>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../include/bacon_ivp.h"
#include "ivp_common.cuh"
#include "path_query.cuh"

namespace {

thread_local std::string g_last_error;
thread_local bacon_ivp_launch_info g_last_launch = {};

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

// a launcher returned rc: keep the message it left (a runtime-compiled right-hand side reports the compiler log that way)
int launch_failed(int rc, int dev) {
    if (!g_last_error.empty()) return rc;
    if (dev >= 0) return fail(rc, "kernel launch failed on device %d: %s", dev, cudaGetErrorString(cudaGetLastError()));
    return fail(rc, "kernel launch failed: %s", cudaGetErrorString(cudaGetLastError()));
}

#define CUDA_TRY(expr)                                                                          \
    do {                                                                                        \
        cudaError_t e__ = (expr);                                                               \
        if (e__ != cudaSuccess)                                                                 \
            return fail(BACON_E_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__),   \
                        __FILE__, __LINE__);                                                    \
    } while (0)

}  // namespace
namespace bacon_internal { void set_last_error(const char* msg) { g_last_error = msg; } }  // for rtc.cu
namespace {

// ---------------------------------------------------------------- RHS registry
struct RhsEntry {
    std::string name;
    int dim, n_params;
    bacon_launch_fn launch[2][BACON_N_METHODS];
    bacon_path_fn path_query[2];
    bacon_launch_fn launch_event[2][BACON_N_METHODS];
};
struct Registry {
    std::mutex mu;
    std::deque<RhsEntry> entries;  // (a deque: bacon_rhs_info hands out name.c_str(), which must survive later registrations)
};
Registry& registry() {
    static Registry* r = new Registry();  // leaked on purpose: RHS translation units register during static init
    return *r;
}

// ---------------------------------------------------------------- per-device context
constexpr int kCounterSlots = 256;
struct DeviceCtx {
    bool ready = false;
    int sm_count = 0;
    unsigned long long* counters = nullptr;  // ring of work counters, one per in-flight launch
    int next_counter = 0;
    std::mutex use;                          // held by a host-buffer solve while it uses stream / events / staging below
    cudaStream_t stream = nullptr;           // used by the host-buffer entry points
    cudaStream_t copy_stream = nullptr;      // background DMA of the late inputs (zero-copy path)
    cudaEvent_t ev[4] = {};
    // grow-only device staging for the host-buffer entry points
    void* d_buf = nullptr;
    size_t d_cap = 0;
    // pinned host staging for the multi-GPU entry point
    void* h_buf = nullptr;
    size_t h_cap = 0;
};
std::mutex g_ctx_mu;
DeviceCtx g_ctx[64];

int get_ctx(int dev, DeviceCtx** out) {
    if (dev < 0 || dev >= 64) return fail(BACON_E_BAD_ARGUMENT, "device index %d out of range", dev);
    DeviceCtx& c = g_ctx[dev];
    if (!c.ready) {
        CUDA_TRY(cudaDeviceGetAttribute(&c.sm_count, cudaDevAttrMultiProcessorCount, dev));
        CUDA_TRY(cudaMalloc(&c.counters, sizeof(unsigned long long) * kCounterSlots));
        CUDA_TRY(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
        CUDA_TRY(cudaStreamCreateWithFlags(&c.copy_stream, cudaStreamNonBlocking));
        for (auto& e : c.ev) CUDA_TRY(cudaEventCreate(&e));
        c.ready = true;
    }
    *out = &c;
    return 0;
}

int ensure_device_buf(DeviceCtx& c, size_t bytes) {
    if (bytes <= c.d_cap) return 0;
    if (c.d_buf) CUDA_TRY(cudaFree(c.d_buf));
    c.d_buf = nullptr;
    c.d_cap = 0;
    CUDA_TRY(cudaMalloc(&c.d_buf, bytes));
    c.d_cap = bytes;
    return 0;
}
int ensure_host_buf(DeviceCtx& c, size_t bytes) {
    if (bytes <= c.h_cap) return 0;
    if (c.h_buf) CUDA_TRY(cudaFreeHost(c.h_buf));
    c.h_buf = nullptr;
    c.h_cap = 0;
    CUDA_TRY(cudaMallocHost(&c.h_buf, bytes));
    c.h_cap = bytes;
    return 0;
}

// ---------------------------------------------------------------- pinned host blocks (bacon_host_alloc)
struct PinnedPool {
    std::mutex mu;
    std::multimap<size_t, void*> free_blocks;   // size -> block, reused by exact (rounded) size
    std::unordered_map<void*, size_t> live;     // blocks handed out
    size_t cached = 0;
    static constexpr size_t kMaxCached = (size_t)2 << 30;
};
PinnedPool& pinned_pool() {
    static PinnedPool* p = new PinnedPool();
    return *p;
}

// host pointer -> is it page-locked, and what does the device call it?
bool pinned_device_pointer(const void* host, void** dev) {
    if (!host) { *dev = nullptr; return true; }
    cudaPointerAttributes at{};
    if (cudaPointerGetAttributes(&at, host) != cudaSuccess) { cudaGetLastError(); return false; }
    if (at.type != cudaMemoryTypeHost || !at.devicePointer) return false;
    *dev = at.devicePointer;
    return true;
}

// thread-local timing events of the device entry point (one pair per device)
struct TlEvents {
    cudaEvent_t start[64] = {}, stop[64] = {};
    int last_dev = -1;
    bool pending = false;
};
thread_local TlEvents g_tl;

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// carve the outputs (and optionally inputs) of one shard out of a device buffer
struct Carve {
    unsigned char* base;
    size_t off = 0;
    explicit Carve(void* b) : base((unsigned char*)b) {}
    template <class T> T* take(size_t count) {
        T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
        off += align_up(count * sizeof(T), 256);
        return p;
    }
};

struct ShardLayout {
    double* y0;
    double* params;
    double* t0;   // bacon_ivp_options::t_start_each / dt_start_each of the shard, when given
    double* dt0;
    bacon_ivp_result out;
    size_t bytes;
};

// the caller's current device is restored on every way out of a multi-device call
struct DeviceGuard {
    int dev = 0;
    bool ok;
    DeviceGuard() { ok = cudaGetDevice(&dev) == cudaSuccess; }
    ~DeviceGuard() {
        if (ok) cudaSetDevice(dev);
    }
};

// run fn(task) for task in [0, n_tasks) on up to `max_threads` host threads (packing / scattering strided shards)
void parallel_for(size_t n_tasks, const std::function<void(size_t)>& fn, unsigned max_threads = 16) {
    unsigned hw = std::thread::hardware_concurrency();
    if (hw == 0) hw = 1;
    const size_t nt = std::min<size_t>(std::min<size_t>(hw, max_threads), n_tasks);
    if (nt <= 1) {
        for (size_t t = 0; t < n_tasks; ++t) fn(t);
        return;
    }
    std::vector<std::thread> pool;
    for (size_t w = 0; w < nt; ++w)
        pool.emplace_back([&, w] {
            for (size_t t = w; t < n_tasks; t += nt) fn(t);
        });
    for (auto& th : pool) th.join();
}

// which outputs the caller asked for decides what is allocated and copied back
ShardLayout layout_shard(void* base, const bacon_ivp_config& cfg, size_t n, bool shared_params,
                         const bacon_ivp_result& want, const bacon_ivp_options* opts) {
    Carve c(base);
    ShardLayout L{};
    L.y0 = c.take<double>((size_t)cfg.dim * n);
    L.params = cfg.n_params > 0 ? c.take<double>(shared_params ? (size_t)cfg.n_params : (size_t)cfg.n_params * n) : nullptr;
    L.t0 = (opts && opts->t_start_each) ? c.take<double>(n) : nullptr;
    L.dt0 = (opts && opts->dt_start_each) ? c.take<double>(n) : nullptr;
    L.out.y_end = c.take<double>((size_t)cfg.dim * n);
    L.out.t_end = want.t_end ? c.take<double>(n) : nullptr;
    L.out.dt_end = want.dt_end ? c.take<double>(n) : nullptr;
    L.out.status = c.take<int32_t>(n);
    L.out.n_accept = want.n_accept ? c.take<uint32_t>(n) : nullptr;
    L.out.n_reject = want.n_reject ? c.take<uint32_t>(n) : nullptr;
    L.out.n_rhs = want.n_rhs ? c.take<uint32_t>(n) : nullptr;
    const size_t cap = cfg.history_capacity > 0 ? (size_t)cfg.history_capacity : 0;
    L.out.hist = cap ? c.take<double>(n * cap * (size_t)(1 + cfg.dim)) : nullptr;
    L.out.hist_len = (cap && want.hist_len) ? c.take<uint32_t>(n) : nullptr;
    L.bytes = c.off;
    return L;
}

int check_common(const bacon_ivp_config* cfg, int rhs_id, const double* y0, const double* params,
                 const bacon_ivp_options* opts, const bacon_ivp_result* out, RhsEntry* entry, bacon_launch_fn* fn) {
    if (!cfg || !out) return fail(BACON_E_BAD_ARGUMENT, "cfg and out must not be NULL");
    const int v = bacon_ivp_validate(cfg);
    if (v != 0) return v;
    {
        Registry& r = registry();
        std::lock_guard<std::mutex> lk(r.mu);
        if (rhs_id < 0 || rhs_id >= (int)r.entries.size()) return fail(BACON_E_BAD_ARGUMENT, "unknown rhs id %d", rhs_id);
        *entry = r.entries[rhs_id];
    }
    // (the reference has no counterpart of this failure: a solver and a derivative of different dimensions do not
    // type-check there, or panic inside nalgebra, ivp.rs:178; the Dimension errors 11 / 12 belong to the constructors)
    if (cfg->dim != entry->dim)
        return fail(BACON_E_BAD_ARGUMENT, "cfg.dim=%d but rhs '%s' has DIM=%d", cfg->dim, entry->name.c_str(), entry->dim);
    if (cfg->n_params != entry->n_params)
        return fail(BACON_E_BAD_ARGUMENT, "cfg.n_params=%d but rhs '%s' has NPARAM=%d", cfg->n_params,
                    entry->name.c_str(), entry->n_params);
    if (!y0 || !out->y_end || !out->status) return fail(BACON_E_BAD_ARGUMENT, "y0, out.y_end and out.status are required");
    if (entry->n_params > 0 && !params) return fail(BACON_E_BAD_ARGUMENT, "rhs '%s' needs params", entry->name.c_str());
    if (cfg->history_capacity > 0 && !out->hist)
        return fail(BACON_E_BAD_ARGUMENT, "history_capacity > 0 needs out.hist ([n][capacity][1 + dim])");
    // REF_LITERAL is the source as written, operation order included: only the strict kernels implement it
    const int strict = ((cfg->flags & BACON_FLAG_STRICT_FP) || cfg->semantics == BACON_SEM_LITERAL) ? 1 : 0;
    // any optional input (restart record, terminal event) selects the kernels compiled for them (drive.cuh: EVENT)
    const bool with_opts = cfg->dt_init > 0.0 || (opts && (opts->event_w || opts->t_start_each || opts->dt_start_each));
    if (opts && opts->event_w) {
        if (cfg->dim > 32) return fail(BACON_E_UNSUPPORTED, "terminal events: dim <= 32");
        if (opts->event_direction < -1 || opts->event_direction > 1)
            return fail(BACON_E_BAD_ARGUMENT, "event_direction must be -1, 0 or +1");
    }
    *fn = with_opts ? entry->launch_event[strict][cfg->method] : entry->launch[strict][cfg->method];
    if (!*fn)
        return fail(BACON_E_UNSUPPORTED, "rhs '%s' was not built for method %d (%s%s)", entry->name.c_str(), cfg->method,
                    strict ? "strict" : "fast", with_opts ? ", optional inputs" : "");
    return 0;
}

// the optional inputs of a launch; t0 / dt0 are DEVICE pointers here
void apply_options(bacon_launch_args& a, const bacon_ivp_options* opts, const double* d_t0, const double* d_dt0) {
    if (!opts) return;
    a.t0_each = d_t0;
    a.dt0_each = d_dt0;
    if (opts->event_w) {
        a.ev_on = 1;
        a.ev_direction = opts->event_direction;
        a.ev_c = opts->event_c;
        for (int d = 0; d < a.cfg.dim && d < 32; ++d) a.ev_w[d] = opts->event_w[d];
    }
}

// What every launch needs: the problem, the stream, and a zeroed work counter from the context's ring (the ring's
// cursor is guarded by g_ctx_mu; fill_args_locked is for callers that hold it already).
int fill_args_locked(bacon_launch_args& a, const bacon_ivp_config* cfg, size_t n, const double* y0, const double* params,
                     const bacon_ivp_result& out, cudaStream_t stream, DeviceCtx& ctx) {
    a = bacon_launch_args{};
    a.cfg = *cfg;
    a.n = n;
    a.y0 = y0;
    a.params = params;
    a.out = out;
    a.stream = stream;
    a.sm_count = ctx.sm_count;
    a.work_counter = ctx.counters + ctx.next_counter;
    ctx.next_counter = (ctx.next_counter + 1) % kCounterSlots;
    CUDA_TRY(cudaMemsetAsync(a.work_counter, 0, sizeof(unsigned long long), stream));
    return 0;
}

int fill_args(bacon_launch_args& a, const bacon_ivp_config* cfg, size_t n, const double* y0, const double* params,
              const bacon_ivp_result& out, cudaStream_t stream, DeviceCtx& ctx) {
    std::lock_guard<std::mutex> lk(g_ctx_mu);
    return fill_args_locked(a, cfg, n, y0, params, out, stream, ctx);
}

int launch_on_device(const bacon_ivp_config* cfg, bacon_launch_fn fn, size_t n, const double* d_y0,
                     const double* d_params, const bacon_ivp_options* opts, const bacon_ivp_result* d_out,
                     cudaStream_t stream, int dev, DeviceCtx& ctx, cudaEvent_t ev_start, cudaEvent_t ev_stop,
                     bacon_launch_args* filled) {
    bacon_launch_args a;
    {
        std::lock_guard<std::mutex> lk(g_ctx_mu);
        if (const int rc = fill_args_locked(a, cfg, n, d_y0, d_params, *d_out, stream, ctx)) return rc;
    }
    apply_options(a, opts, opts ? opts->t_start_each : nullptr, opts ? opts->dt_start_each : nullptr);
    (void)dev;
    if (ev_start) CUDA_TRY(cudaEventRecord(ev_start, stream));
    g_last_error.clear();
    const int rc = fn(&a);
    if (rc != 0) return launch_failed(rc, -1);
    if (ev_stop) CUDA_TRY(cudaEventRecord(ev_stop, stream));
    if (filled) *filled = a;
    return 0;
}

}  // namespace

// ===================================================================== C ABI
extern "C" {

int bacon_abi_version(void) { return BACON_IVP_ABI_VERSION; }

const char* bacon_last_error(void) { return g_last_error.c_str(); }

const char* bacon_status_name(int s) {
    static const char* names[] = {"Ok", "MissingParameters", "UserError", "ToleranceOOB", "TimeDeltaOOB", "TimeEndOOB",
                                  "TimeStartOOB", "FromPrimitiveFailure", "MinimumTimeDeltaExceeded",
                                  "MaximumIterationsExceeded", "SingularMatrix", "DynamicOnStatic", "StaticOnDynamic",
                                  "NonFinite", "MaxAttempts", "HistoryOverflow", "CudaError", "BadArgument", "Unsupported",
                                  "StoppedAtEvent"};
    if (s < 0 || s > BACON_STOPPED_AT_EVENT) return "Unknown";
    return names[s];
}

// ---------------------------------------------------------------- builder (rk.rs:118-256)
struct bacon_solver {
    int method, dim;
    bool has_tol, has_max, has_min, has_t0, has_t1;
    double tol, dt_max, dt_min, t0, t1;
    int semantics;
    uint32_t flags;
    int history;
    uint64_t max_attempts;
    bool has_euler_dt;  // Euler keeps ONE dt: the first bound given, then averaged with later ones (ivp.rs:396-421)
    double euler_dt;
    double dt_init;     // bacon_solver_with_initial_dt; 0 = the reference's (dt_max + dt_min)/2
};

// IVPSolver::new / new_dyn with the reference's Dimension check (lib.rs:53-76): dim_type >= 1 is Const<C>, BACON_DIM_DYN is Dyn
int bacon_solver_new_static(int method, int dim_type, bacon_solver** out) {
    if (!out) return fail(BACON_E_BAD_ARGUMENT, "NULL out");
    *out = nullptr;
    if (dim_type == BACON_DIM_DYN)  // Dyn::dim() (lib.rs:69-71)
        return fail(BACON_E_STATIC_ON_DYNAMIC, "attempted to build a static solver with dynamic dimension");
    if (dim_type < 0) return fail(BACON_E_BAD_ARGUMENT, "dim_type must be a dimension >= 1 or BACON_DIM_DYN");
    *out = bacon_solver_new(method, dim_type);  // Const<C>::dim() (lib.rs:59-61)
    return *out ? 0 : BACON_E_BAD_ARGUMENT;
}
int bacon_solver_new_dyn(int method, int dim_type, int size, bacon_solver** out) {
    if (!out) return fail(BACON_E_BAD_ARGUMENT, "NULL out");
    *out = nullptr;
    if (dim_type != BACON_DIM_DYN) {  // Const<C>::dim_dyn(size) (lib.rs:63-65)
        if (dim_type < 0) return fail(BACON_E_BAD_ARGUMENT, "dim_type must be a dimension >= 1 or BACON_DIM_DYN");
        return fail(BACON_E_DYNAMIC_ON_STATIC, "attempted to build a dynamic solver with static dimension");
    }
    *out = bacon_solver_new(method, size);  // Dyn::dim_dyn(size) (lib.rs:73-75)
    return *out ? 0 : BACON_E_BAD_ARGUMENT;
}

bacon_solver* bacon_solver_new(int method, int dim) {
    if (method < 0 || method >= BACON_N_METHODS) {
        fail(BACON_E_BAD_ARGUMENT, "unknown method %d", method);
        return nullptr;
    }
    if (dim < 1) {  // Dimension::dim_dyn of a zero-sized system is meaningless here (lib.rs:53-76)
        fail(BACON_E_BAD_ARGUMENT, "dim must be >= 1");
        return nullptr;
    }
    bacon_solver* s = new bacon_solver();
    std::memset(s, 0, sizeof(*s));
    s->method = method;
    s->dim = dim;
    return s;
}
void bacon_solver_free(bacon_solver* s) { delete s; }

int bacon_solver_with_tolerance(bacon_solver* s, double tol) {
    if (!s) return fail(BACON_E_BAD_ARGUMENT, "NULL solver");
    if (s->method == BACON_EULER) return 0;  // "Unused for Euler, call is a no-op" (ivp.rs:389-392)
    if (tol <= 0.0) return fail(BACON_E_TOLERANCE_OOB, "tolerance must be > 0");  // rk.rs:169-171
    s->tol = tol;
    s->has_tol = true;
    return 0;
}
int bacon_solver_with_maximum_dt(bacon_solver* s, double max) {
    if (!s) return fail(BACON_E_BAD_ARGUMENT, "NULL solver");
    if (max <= 0.0) return fail(BACON_E_TIME_DELTA_OOB, "maximum dt must be > 0");  // rk.rs:180-182
    if (s->method == BACON_EULER) {  // ivp.rs:396-406
        s->euler_dt = s->has_euler_dt ? (s->euler_dt + max) / 2.0 : max;
        s->has_euler_dt = true;
        return 0;
    }
    s->dt_max = max;
    s->has_max = true;
    if (s->has_min && s->dt_min > max) s->dt_min = max;  // rk.rs:185-189
    return 0;
}
int bacon_solver_with_minimum_dt(bacon_solver* s, double min) {
    if (!s) return fail(BACON_E_BAD_ARGUMENT, "NULL solver");
    if (min <= 0.0) return fail(BACON_E_TIME_DELTA_OOB, "minimum dt must be > 0");  // rk.rs:198-200
    if (s->method == BACON_EULER) {  // ivp.rs:411-421
        s->euler_dt = s->has_euler_dt ? (s->euler_dt + min) / 2.0 : min;
        s->has_euler_dt = true;
        return 0;
    }
    s->dt_min = min;
    s->has_min = true;
    if (s->has_max && s->dt_max < min) s->dt_max = min;  // rk.rs:203-207
    return 0;
}
int bacon_solver_with_initial_time(bacon_solver* s, double initial) {
    if (!s) return fail(BACON_E_BAD_ARGUMENT, "NULL solver");
    s->t0 = initial;  // stored first, then checked (rk.rs:213-219)
    s->has_t0 = true;
    if (s->has_t1 && s->t1 <= initial) return fail(BACON_E_TIME_START_OOB, "initial time must be before the ending time");
    return 0;
}
int bacon_solver_with_ending_time(bacon_solver* s, double ending) {
    if (!s) return fail(BACON_E_BAD_ARGUMENT, "NULL solver");
    s->t1 = ending;  // rk.rs:225-231
    s->has_t1 = true;
    if (s->has_t0 && s->t0 >= ending) return fail(BACON_E_TIME_END_OOB, "ending time must be after the initial time");
    return 0;
}
int bacon_solver_with_initial_dt(bacon_solver* s, double dt) {
    if (!s) return fail(BACON_E_BAD_ARGUMENT, "NULL solver");
    if (!(dt > 0.0)) return fail(BACON_E_TIME_DELTA_OOB, "initial dt must be > 0");
    s->dt_init = dt;
    return 0;
}
int bacon_solver_with_semantics(bacon_solver* s, int semantics) {
    if (!s || (semantics != BACON_SEM_CORRECTED && semantics != BACON_SEM_LITERAL))
        return fail(BACON_E_BAD_ARGUMENT, "bad semantics");
    s->semantics = semantics;
    return 0;
}
int bacon_solver_with_flags(bacon_solver* s, uint32_t flags) {
    if (!s) return fail(BACON_E_BAD_ARGUMENT, "NULL solver");
    s->flags = flags;
    return 0;
}
int bacon_solver_with_history(bacon_solver* s, int capacity) {
    if (!s || capacity < 0) return fail(BACON_E_BAD_ARGUMENT, "history capacity must be >= 0");
    s->history = capacity;
    return 0;
}
int bacon_solver_with_max_attempts(bacon_solver* s, uint64_t cap) {
    if (!s) return fail(BACON_E_BAD_ARGUMENT, "NULL solver");
    s->max_attempts = cap;
    return 0;
}
int bacon_solver_config(const bacon_solver* s, bacon_ivp_config* out) {
    if (!s || !out) return fail(BACON_E_BAD_ARGUMENT, "NULL argument");
    if (s->method == BACON_EULER) {  // ivp.rs:451-459: dt, initial time, ending time
        if (!s->has_euler_dt || !s->has_t0 || !s->has_t1)
            return fail(BACON_E_MISSING_PARAMETERS, "a time step (with_maximum_dt / with_minimum_dt), initial time and ending time are required");
        std::memset(out, 0, sizeof(*out));
        out->method = s->method;
        out->dim = s->dim;
        out->semantics = s->semantics;
        out->flags = s->flags;
        out->history_capacity = s->history;
        out->dt_min = out->dt_max = s->euler_dt;
        out->tol = 1.0;  // unused
        out->t_start = s->t0;
        out->t_end = s->t1;
        out->max_attempts = s->max_attempts;
        return 0;
    }
    // rk.rs:250-254, in the reference's order
    if (!s->has_max || !s->has_min || !s->has_tol || !s->has_t0 || !s->has_t1)
        return fail(BACON_E_MISSING_PARAMETERS, "dt_max, dt_min, tolerance, initial time and ending time are all required");
    std::memset(out, 0, sizeof(*out));
    out->method = s->method;
    out->dim = s->dim;
    out->n_params = 0;  // filled from the RHS by the caller
    out->semantics = s->semantics;
    out->flags = s->flags;
    out->history_capacity = s->history;
    out->dt_min = s->dt_min;
    out->dt_max = s->dt_max;
    out->tol = s->tol;
    out->t_start = s->t0;
    out->t_end = s->t1;
    out->max_attempts = s->max_attempts;
    out->dt_init = s->dt_init;
    return 0;
}

int bacon_ivp_validate(const bacon_ivp_config* c) {
    if (!c) return fail(BACON_E_BAD_ARGUMENT, "NULL config");
    if (c->method < 0 || c->method >= BACON_N_METHODS) return fail(BACON_E_BAD_ARGUMENT, "unknown method %d", c->method);
    if (c->dim < 1) return fail(BACON_E_BAD_ARGUMENT, "dim must be >= 1");
    if (c->n_params < 0 || c->history_capacity < 0) return fail(BACON_E_BAD_ARGUMENT, "negative size");
    if (c->semantics != BACON_SEM_CORRECTED && c->semantics != BACON_SEM_LITERAL)
        return fail(BACON_E_BAD_ARGUMENT, "bad semantics %d", c->semantics);
    if (c->tol <= 0.0) return fail(BACON_E_TOLERANCE_OOB, "tolerance must be > 0");
    if (c->dt_max <= 0.0 || c->dt_min <= 0.0) return fail(BACON_E_TIME_DELTA_OOB, "dt bounds must be > 0");
    if (c->dt_min > c->dt_max) return fail(BACON_E_TIME_DELTA_OOB, "dt_min > dt_max");
    if (c->t_end <= c->t_start) return fail(BACON_E_TIME_END_OOB, "t_end must be after t_start");
    if (!(c->dt_init >= 0.0)) return fail(BACON_E_TIME_DELTA_OOB, "dt_init must be > 0 (or 0 = the default)");
    return 0;
}

// ---------------------------------------------------------------- registry
int bacon_rhs_register(const bacon_rhs_desc* d) {
    if (!d || !d->name || d->dim < 1 || d->n_params < 0) return -BACON_E_BAD_ARGUMENT;
    Registry& r = registry();
    std::lock_guard<std::mutex> lk(r.mu);
    for (size_t i = 0; i < r.entries.size(); ++i) {
        RhsEntry& e = r.entries[i];
        if (e.name == d->name) {  // a second translation unit (e.g. the strict build) adds its launchers
            if (e.dim != d->dim || e.n_params != d->n_params) return -BACON_E_BAD_ARGUMENT;
            for (int s = 0; s < 2; ++s)
                for (int m = 0; m < BACON_N_METHODS; ++m)
                    if (d->launch[s][m]) e.launch[s][m] = d->launch[s][m];
            for (int s = 0; s < 2; ++s)
                if (d->path_query[s]) e.path_query[s] = d->path_query[s];
            for (int s = 0; s < 2; ++s)
                for (int m = 0; m < BACON_N_METHODS; ++m)
                    if (d->launch_event[s][m]) e.launch_event[s][m] = d->launch_event[s][m];
            return (int)i;
        }
    }
    RhsEntry e;
    e.name = d->name;
    e.dim = d->dim;
    e.n_params = d->n_params;
    std::memcpy(e.launch, d->launch, sizeof(e.launch));
    std::memcpy(e.path_query, d->path_query, sizeof(e.path_query));
    std::memcpy(e.launch_event, d->launch_event, sizeof(e.launch_event));
    r.entries.push_back(e);
    return (int)r.entries.size() - 1;
}
int bacon_rhs_lookup(const char* name) {
    if (!name) return -1;
    Registry& r = registry();
    std::lock_guard<std::mutex> lk(r.mu);
    for (size_t i = 0; i < r.entries.size(); ++i)
        if (r.entries[i].name == name) return (int)i;
    return -1;
}
int bacon_rhs_count(void) {
    Registry& r = registry();
    std::lock_guard<std::mutex> lk(r.mu);
    return (int)r.entries.size();
}
int bacon_rhs_info(int id, const char** name, int* dim, int* n_params) {
    Registry& r = registry();
    std::lock_guard<std::mutex> lk(r.mu);
    if (id < 0 || id >= (int)r.entries.size()) return fail(BACON_E_BAD_ARGUMENT, "unknown rhs id %d", id);
    if (name) *name = r.entries[id].name.c_str();
    if (dim) *dim = r.entries[id].dim;
    if (n_params) *n_params = r.entries[id].n_params;
    return 0;
}

// ---------------------------------------------------------------- solves
int bacon_ivp_solve_ensemble_device_ex(const bacon_ivp_config* cfg, int rhs_id, size_t n, const double* d_y0,
                                       const double* d_params, const bacon_ivp_options* opts,
                                       const bacon_ivp_result* d_out, void* stream) {
    RhsEntry entry;
    bacon_launch_fn fn = nullptr;
    int rc = check_common(cfg, rhs_id, d_y0, d_params, opts, d_out, &entry, &fn);
    if (rc != 0) return rc;
    if (cfg->history_capacity > 0 && (reinterpret_cast<uintptr_t>(d_out->hist) & 31u))
        return fail(BACON_E_BAD_ARGUMENT, "d_out.hist must be 32-byte aligned (records are written with 256-bit stores)");
    g_last_launch = bacon_ivp_launch_info{};
    if (n == 0) return 0;
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    DeviceCtx* ctx = nullptr;
    {
        std::lock_guard<std::mutex> lk(g_ctx_mu);
        rc = get_ctx(dev, &ctx);
    }
    if (rc != 0) return rc;
    if (!g_tl.start[dev]) {
        CUDA_TRY(cudaEventCreate(&g_tl.start[dev]));
        CUDA_TRY(cudaEventCreate(&g_tl.stop[dev]));
    }
    bacon_launch_args filled{};
    rc = launch_on_device(cfg, fn, n, d_y0, d_params, opts, d_out, (cudaStream_t)stream, dev, *ctx, g_tl.start[dev],
                          g_tl.stop[dev], &filled);
    if (rc != 0) return rc;
    g_tl.last_dev = dev;
    g_tl.pending = true;
    g_last_launch.grid = filled.grid;
    g_last_launch.block = filled.block;
    g_last_launch.regs_per_thread = filled.regs_per_thread;
    g_last_launch.n_kernels = filled.n_kernels;
    return 0;
}
int bacon_ivp_solve_ensemble_device(const bacon_ivp_config* cfg, int rhs_id, size_t n, const double* d_y0,
                                    const double* d_params, const bacon_ivp_result* d_out, void* stream) {
    return bacon_ivp_solve_ensemble_device_ex(cfg, rhs_id, n, d_y0, d_params, nullptr, d_out, stream);
}

int bacon_ivp_solve_ensemble(const bacon_ivp_config* cfg, int rhs_id, size_t n, const double* y0, const double* params,
                             const bacon_ivp_result* out) {
    return bacon_ivp_solve_ensemble_ex(cfg, rhs_id, n, y0, params, nullptr, out, 1);
}
int bacon_ivp_solve_ensemble_multi(const bacon_ivp_config* cfg, int rhs_id, size_t n, const double* y0,
                                   const double* params, const bacon_ivp_result* out, int n_gpus) {
    return bacon_ivp_solve_ensemble_ex(cfg, rhs_id, n, y0, params, nullptr, out, n_gpus);
}

// Round-robin sharding (trajectory i -> GPU i mod G): parameter
// sweeps stay balanced, no data-path collective.  With G == 1 the shard IS the
// caller's buffer and no repacking happens.  Concurrency: a call holds the `use` mutex of every device it runs on
// (taken in device order), so host solves on different devices of one process run side by side.
int bacon_ivp_solve_ensemble_ex(const bacon_ivp_config* cfg, int rhs_id, size_t n, const double* y0,
                                const double* params, const bacon_ivp_options* opts, const bacon_ivp_result* out,
                                int n_gpus) {
    RhsEntry entry;
    bacon_launch_fn fn = nullptr;
    int rc = check_common(cfg, rhs_id, y0, params, opts, out, &entry, &fn);
    if (rc != 0) return rc;
    g_last_launch = bacon_ivp_launch_info{};
    g_tl.pending = false;
    if (n == 0) return 0;
    int have = 0;
    CUDA_TRY(cudaGetDeviceCount(&have));
    if (n_gpus < 1 || n_gpus > have || n_gpus > 64) return fail(BACON_E_BAD_ARGUMENT, "n_gpus=%d but %d device(s) visible", n_gpus, have);
    // BACON_IVP_TIMING=1: the host-side phases of this call on stderr (tools/multi_entry_bench.py)
    static const bool timing = getenv("BACON_IVP_TIMING") != nullptr;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto ms_since = [&](std::chrono::steady_clock::time_point t0) { return std::chrono::duration<double, std::milli>(now() - t0).count(); };
    const auto t_call = now();
    double t_setup = 0, t_pack = 0, t_enq = 0, t_wait = 0, t_scatter = 0;
    DeviceGuard restore;  // the caller's current device comes back on every way out
    if (!restore.ok) return fail(BACON_E_CUDA, "cudaGetDevice failed");
    const int dev0 = restore.dev;
    const int G = n_gpus;
    const bool shared = (cfg->flags & BACON_FLAG_SHARED_PARAMS) != 0;
    const int D = cfg->dim, P = cfg->n_params;
    const size_t cap = cfg->history_capacity > 0 ? (size_t)cfg->history_capacity : 0;
    const bool has_opts = cfg->dt_init > 0.0 || (opts && (opts->t_start_each || opts->dt_start_each || opts->event_w));

    struct Shard {
        int dev = 0;
        size_t n = 0;
        DeviceCtx* ctx = nullptr;
        ShardLayout dl{}, hl{};  // device / pinned-host layouts
        bacon_launch_args filled{};
    };
    std::vector<Shard> shards(G);
    for (int g = 0; g < G; ++g) {
        Shard& s = shards[g];
        s.dev = (G == 1) ? dev0 : g;
        s.n = (n + G - 1 - g) / G;  // indices g, g+G, ...
        std::lock_guard<std::mutex> lk(g_ctx_mu);
        CUDA_TRY(cudaSetDevice(s.dev));
        rc = get_ctx(s.dev, &s.ctx);
        if (rc != 0) return rc;
    }
    std::vector<std::unique_lock<std::mutex>> held;  // (device order: G == 1 holds one, G > 1 holds 0 .. G-1)
    for (int g = 0; g < G; ++g) held.emplace_back(shards[g].ctx->use);

    // Zero-copy: per-trajectory I/O is a few dozen bytes against thousands of register-only steps, so with
    // page-locked caller buffers the persistent kernel can read its initial conditions and post its
    // retirement records over the host link itself; nothing is staged and nothing is left to copy at the end.
    // (Not for large per-trajectory parameter blocks: those are streamed once by DMA instead.)
    const size_t in_doubles = (size_t)D + (shared ? 0 : (size_t)P);
    if (G == 1 && cap == 0 && !has_opts && in_doubles <= 32 && (cfg->flags & BACON_FLAG_ZERO_COPY)) {
        void *zy0 = nullptr, *zp = nullptr;
        bacon_ivp_result z{};
        bool ok = pinned_device_pointer(y0, &zy0) && pinned_device_pointer(P > 0 ? params : nullptr, &zp);
        void* t = nullptr;
#define ZC(field, type)                                    \
    ok = ok && pinned_device_pointer(out->field, &t);      \
    z.field = (type*)t
        ZC(y_end, double);
        ZC(t_end, double);
        ZC(dt_end, double);
        ZC(status, int32_t);
        ZC(n_accept, uint32_t);
        ZC(n_reject, uint32_t);
        ZC(n_rhs, uint32_t);
#undef ZC
        if (ok) {
            DeviceCtx* ctx = shards[0].ctx;
            cudaStream_t st = ctx->stream;
            bacon_launch_args a;
            rc = fill_args(a, cfg, n, (const double*)zy0, (const double*)zp, z, st, *ctx);
            if (rc != 0) return rc;
            // A refill over the host link costs a lane ~2 us and its 31 warp-mates wait at the loop's latch, ~9 times per
            // lane: 0.6 ms of a 31 ms launch.  So only the FIRST trajectory of every lane is read from the caller's
            // memory (nothing to wait for); meanwhile a DMA copies all inputs into device memory on a second stream
            // and then raises a flag the refills check (with a bounded wait: ivp_common.cuh, wait_until_set).
            const size_t y0_bytes = sizeof(double) * (size_t)D * n, p_bytes = (shared || P == 0) ? 0 : sizeof(double) * (size_t)P * n;
            rc = ensure_device_buf(*ctx, y0_bytes + p_bytes + 256);
            if (rc != 0) return rc;
            char* dbuf = (char*)ctx->d_buf;
            unsigned int* flag = (unsigned int*)(dbuf + y0_bytes + p_bytes);
            CUDA_TRY(cudaMemsetAsync(flag, 0, sizeof(unsigned int), st));
            CUDA_TRY(cudaEventRecord(ctx->ev[3], st));
            CUDA_TRY(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev[3], 0));
            CUDA_TRY(cudaMemcpyAsync(dbuf, y0, y0_bytes, cudaMemcpyHostToDevice, ctx->copy_stream));
            if (p_bytes) CUDA_TRY(cudaMemcpyAsync(dbuf + y0_bytes, params, p_bytes, cudaMemcpyHostToDevice, ctx->copy_stream));
            CUDA_TRY(cudaMemsetAsync(flag, 1, sizeof(unsigned int), ctx->copy_stream));
            a.y0_late = (const double*)dbuf;
            a.params_late = p_bytes ? (const double*)(dbuf + y0_bytes) : nullptr;
            a.late_ready = flag;
            CUDA_TRY(cudaEventRecord(ctx->ev[1], st));
            g_last_error.clear();
            rc = fn(&a);
            if (rc != 0) return launch_failed(rc, -1);
            CUDA_TRY(cudaEventRecord(ctx->ev[2], st));
            CUDA_TRY(cudaStreamSynchronize(st));
            CUDA_TRY(cudaStreamSynchronize(ctx->copy_stream));
            float ms = 0.f;
            CUDA_TRY(cudaEventElapsedTime(&ms, ctx->ev[1], ctx->ev[2]));
            g_last_launch.kernel_ms = ms;
            g_last_launch.grid = a.grid;
            g_last_launch.block = a.block;
            g_last_launch.regs_per_thread = a.regs_per_thread;
            g_last_launch.n_kernels = a.n_kernels;
            return 0;
        }
    }

    for (int g = 0; g < G; ++g) {
        Shard& s = shards[g];
        if (s.n == 0) continue;
        CUDA_TRY(cudaSetDevice(s.dev));
        s.dl = layout_shard(nullptr, *cfg, s.n, shared, *out, opts);
        rc = ensure_device_buf(*s.ctx, s.dl.bytes);
        if (rc != 0) return rc;
        s.dl = layout_shard(s.ctx->d_buf, *cfg, s.n, shared, *out, opts);
        if (G > 1) {
            rc = ensure_host_buf(*s.ctx, s.dl.bytes);
            if (rc != 0) return rc;
            s.hl = layout_shard(s.ctx->h_buf, *cfg, s.n, shared, *out, opts);
        }
    }
    t_setup = ms_since(t_call);
    auto t_phase = now();
    // Dealing trajectory i to shard i mod G (and back) on host threads.  A task is a contiguous range of GLOBAL indices
    // and walks it in order — k outer, g inner — so the caller's arrays are read (written) sequentially, once, and
    // every shard's buffer is a sequential stream of its own; walking shard by shard instead touches every cache line
    // of the caller's arrays G times (measured on 8 GPUs, 2^20 Lorenz trajectories: 10.4 ms per call against 5.3 ms of
    // GPU time; profiles/r03_multi_entry.md).
    constexpr size_t DEAL_CHUNK = (size_t)1 << 12;  // trajectories per shard and task
    const size_t k_max = (n + G - 1) / G;
    const size_t deal_tasks = (k_max + DEAL_CHUNK - 1) / DEAL_CHUNK;
    const size_t k_full = n / G;  // rows in which every shard has a trajectory
    if (G > 1) {
        parallel_for(deal_tasks, [&](size_t task) {
            const size_t k0 = task * DEAL_CHUNK, k1 = std::min(k_max, k0 + DEAL_CHUNK);
            // rows k < k_full hold all G shards (no test in the inner loop); the last row may be ragged
            const size_t kf = std::min(k1, k_full);
            auto deal = [&](const double* src, auto&& dst_of /* (g) -> double* */) {
                double* dst[64];
                for (int g = 0; g < G; ++g) dst[g] = dst_of(g);
                for (size_t k = k0; k < kf; ++k) {
                    const double* row = src + k * G;
                    for (int g = 0; g < G; ++g) dst[g][k] = row[g];
                }
                for (size_t k = std::max(k0, kf); k < k1; ++k)
                    for (int g = 0; g < G; ++g)
                        if (k < shards[g].n) dst[g][k] = src[k * G + g];
            };
            for (int d = 0; d < D; ++d) deal(y0 + (size_t)d * n, [&](int g) { return shards[g].hl.y0 + (size_t)d * shards[g].n; });
            if (P > 0 && !shared) {
                if (cfg->flags & BACON_FLAG_PARAMS_AOS) {
                    for (size_t k = k0; k < k1; ++k)
                        for (int g = 0; g < G; ++g)
                            if (k < shards[g].n)
                                std::memcpy(shards[g].hl.params + k * P, params + (k * G + g) * (size_t)P, sizeof(double) * P);
                } else {
                    for (int p = 0; p < P; ++p) deal(params + (size_t)p * n, [&](int g) { return shards[g].hl.params + (size_t)p * shards[g].n; });
                }
            }
            if (opts && opts->t_start_each) deal(opts->t_start_each, [&](int g) { return shards[g].hl.t0; });
            if (opts && opts->dt_start_each) deal(opts->dt_start_each, [&](int g) { return shards[g].hl.dt0; });
        }, 32);
        if (P > 0 && shared)
            for (int g = 0; g < G; ++g)
                if (shards[g].n) std::memcpy(shards[g].hl.params, params, sizeof(double) * P);
    }

    t_pack = ms_since(t_phase);
    t_phase = now();
    // enqueue H2D -> kernel -> D2H on every device's own stream, then wait for all
    for (int g = 0; g < G; ++g) {
        Shard& s = shards[g];
        if (s.n == 0) continue;
        CUDA_TRY(cudaSetDevice(s.dev));
        cudaStream_t st = s.ctx->stream;
        const double* src_y0 = (G == 1) ? y0 : s.hl.y0;
        const double* src_p = (G == 1) ? params : s.hl.params;
        CUDA_TRY(cudaEventRecord(s.ctx->ev[0], st));
        CUDA_TRY(cudaMemcpyAsync(s.dl.y0, src_y0, sizeof(double) * D * s.n, cudaMemcpyHostToDevice, st));
        if (P > 0)
            CUDA_TRY(cudaMemcpyAsync(s.dl.params, src_p, sizeof(double) * (shared ? (size_t)P : (size_t)P * s.n),
                                     cudaMemcpyHostToDevice, st));
        if (s.dl.t0)
            CUDA_TRY(cudaMemcpyAsync(s.dl.t0, (G == 1) ? opts->t_start_each : s.hl.t0, sizeof(double) * s.n, cudaMemcpyHostToDevice, st));
        if (s.dl.dt0)
            CUDA_TRY(cudaMemcpyAsync(s.dl.dt0, (G == 1) ? opts->dt_start_each : s.hl.dt0, sizeof(double) * s.n, cudaMemcpyHostToDevice, st));
        {
            bacon_launch_args a;
            rc = fill_args(a, cfg, s.n, s.dl.y0, s.dl.params, s.dl.out, st, *s.ctx);
            if (rc != 0) return rc;
            apply_options(a, opts, s.dl.t0, s.dl.dt0);
            if (cap)  // slots beyond hist_len read as zero on the host (the staging buffer is reused between calls)
                CUDA_TRY(cudaMemsetAsync(s.dl.out.hist, 0, sizeof(double) * s.n * cap * (D + 1), st));
            CUDA_TRY(cudaEventRecord(s.ctx->ev[1], st));
            g_last_error.clear();
            rc = fn(&a);
            if (rc != 0) return launch_failed(rc, s.dev);
            CUDA_TRY(cudaEventRecord(s.ctx->ev[2], st));
            s.filled = a;
        }
        const bacon_ivp_result& dst = (G == 1) ? *out : s.hl.out;
        const bacon_ivp_result& src = s.dl.out;
#define D2H(field, type, count)                                                                       \
    if (dst.field && src.field)                                                                       \
    CUDA_TRY(cudaMemcpyAsync(dst.field, src.field, sizeof(type) * (count), cudaMemcpyDeviceToHost, st))
        D2H(y_end, double, (size_t)D * s.n);
        D2H(t_end, double, s.n);
        D2H(dt_end, double, s.n);
        D2H(status, int32_t, s.n);
        D2H(n_accept, uint32_t, s.n);
        D2H(n_reject, uint32_t, s.n);
        D2H(n_rhs, uint32_t, s.n);
        if (cap) {
            D2H(hist, double, s.n * cap * (D + 1));
            D2H(hist_len, uint32_t, s.n);
        }
#undef D2H
        CUDA_TRY(cudaEventRecord(s.ctx->ev[3], st));
    }

    t_enq = ms_since(t_phase);
    t_phase = now();
    float k_ms = 0.f, h2d_ms = 0.f, d2h_ms = 0.f;
    for (int g = 0; g < G; ++g) {
        Shard& s = shards[g];
        if (s.n == 0) continue;
        CUDA_TRY(cudaSetDevice(s.dev));
        CUDA_TRY(cudaStreamSynchronize(s.ctx->stream));
        float a = 0, b = 0, c = 0;
        CUDA_TRY(cudaEventElapsedTime(&a, s.ctx->ev[0], s.ctx->ev[1]));
        CUDA_TRY(cudaEventElapsedTime(&b, s.ctx->ev[1], s.ctx->ev[2]));
        CUDA_TRY(cudaEventElapsedTime(&c, s.ctx->ev[2], s.ctx->ev[3]));
        h2d_ms = a > h2d_ms ? a : h2d_ms;
        k_ms = b > k_ms ? b : k_ms;  // max over devices
        d2h_ms = c > d2h_ms ? c : d2h_ms;
    }
    t_wait = ms_since(t_phase);
    t_phase = now();
    if (G > 1) {  // scatter the shards back into the caller's arrays (same tasks, the caller's arrays written in order)
        parallel_for(deal_tasks, [&](size_t task) {
            const size_t k0 = task * DEAL_CHUNK, k1 = std::min(k_max, k0 + DEAL_CHUNK);
            const size_t kf = std::min(k1, k_full);
            auto gather = [&](auto* dst, auto&& src_of /* (g) -> const T* */) {
                decltype(src_of(0)) src[64];
                for (int g = 0; g < G; ++g) src[g] = src_of(g);
                for (size_t k = k0; k < kf; ++k) {
                    auto* row = dst + k * G;
                    for (int g = 0; g < G; ++g) row[g] = src[g][k];
                }
                for (size_t k = std::max(k0, kf); k < k1; ++k)
                    for (int g = 0; g < G; ++g)
                        if (k < shards[g].n) dst[k * G + g] = src[g][k];
            };
            for (int d = 0; d < D; ++d)
                gather(out->y_end + (size_t)d * n, [&](int g) -> const double* { return shards[g].hl.out.y_end + (size_t)d * shards[g].n; });
#define SCATTER(field, type) \
    if (out->field && shards[0].hl.out.field) gather(out->field, [&](int g) -> const type* { return shards[g].hl.out.field; })
            SCATTER(t_end, double);
            SCATTER(dt_end, double);
            SCATTER(status, int32_t);
            SCATTER(n_accept, uint32_t);
            SCATTER(n_reject, uint32_t);
            SCATTER(n_rhs, uint32_t);
            if (cap) {
                SCATTER(hist_len, uint32_t);
                const size_t path = cap * (size_t)(D + 1);  // doubles per trajectory
                for (size_t k = k0; k < k1; ++k)
                    for (int g = 0; g < G; ++g)
                        if (k < shards[g].n)
                            std::memcpy(out->hist + (k * G + g) * path, shards[g].hl.out.hist + k * path, sizeof(double) * path);
            }
#undef SCATTER
        }, 32);
    }
    t_scatter = ms_since(t_phase);
    if (timing)
        fprintf(stderr, "[bacon_ivp] G=%d n=%zu: setup %.3f pack %.3f enqueue %.3f wait %.3f (kernel %.3f h2d %.3f d2h %.3f) scatter %.3f total %.3f ms\n",
                G, n, t_setup, t_pack, t_enq, t_wait, k_ms, h2d_ms, d2h_ms, t_scatter, ms_since(t_call));
    g_last_launch.kernel_ms = k_ms;
    g_last_launch.h2d_ms = h2d_ms;
    g_last_launch.d2h_ms = d2h_ms;
    g_last_launch.grid = shards[0].filled.grid;
    g_last_launch.block = shards[0].filled.block;
    g_last_launch.regs_per_thread = shards[0].filled.regs_per_thread;
    g_last_launch.n_kernels = G * shards[0].filled.n_kernels;
    return 0;
}

// ---------------------------------------------------------------- queries on stored paths
}  // extern "C"
namespace {

struct PathQuery {
    int op;
    size_t n_times;
    const double* times;
    double* samples;
    const double* w;  // host, [dim]
    double c;
    int direction, capacity;
    double* events;
    uint32_t* n_events;
};

int path_query_check(const bacon_ivp_config* cfg, int rhs_id, const double* y0, const double* params,
                     const bacon_ivp_result* solved, const PathQuery& q, RhsEntry* entry, bacon_path_fn* fn) {
    if (!cfg || !solved) return fail(BACON_E_BAD_ARGUMENT, "cfg and the solved result must not be NULL");
    const int v = bacon_ivp_validate(cfg);
    if (v != 0) return v;
    {
        Registry& r = registry();
        std::lock_guard<std::mutex> lk(r.mu);
        if (rhs_id < 0 || rhs_id >= (int)r.entries.size()) return fail(BACON_E_BAD_ARGUMENT, "unknown rhs id %d", rhs_id);
        *entry = r.entries[rhs_id];
    }
    if (cfg->dim != entry->dim || cfg->n_params != entry->n_params)
        return fail(BACON_E_BAD_ARGUMENT, "cfg (dim %d, n_params %d) does not match rhs '%s' (%d, %d)", cfg->dim,
                    cfg->n_params, entry->name.c_str(), entry->dim, entry->n_params);
    if (cfg->history_capacity <= 0 || !solved->hist || !solved->hist_len)
        return fail(BACON_E_BAD_ARGUMENT, "a path query needs the history of a dense-output solve (history_capacity > 0, hist, hist_len)");
    if (!y0) return fail(BACON_E_BAD_ARGUMENT, "y0 is required (knot 0 of every path)");
    if (entry->n_params > 0 && !params) return fail(BACON_E_BAD_ARGUMENT, "rhs '%s' needs params", entry->name.c_str());
    if (q.op == BACON_PATH_SAMPLE) {
        if (q.n_times > 0 && (!q.times || !q.samples)) return fail(BACON_E_BAD_ARGUMENT, "times and samples are required");
    } else {
        if (!q.w || !q.n_events) return fail(BACON_E_BAD_ARGUMENT, "w and n_events are required");
        if (q.capacity < 0 || (q.capacity > 0 && !q.events)) return fail(BACON_E_BAD_ARGUMENT, "capacity > 0 needs events");
        if (q.direction < -1 || q.direction > 1) return fail(BACON_E_BAD_ARGUMENT, "direction must be -1, 0 or +1");
        if (cfg->dim > BACON_PATH_MAX_DIM) return fail(BACON_E_UNSUPPORTED, "events: dim <= %d", BACON_PATH_MAX_DIM);
    }
    const int strict = ((cfg->flags & BACON_FLAG_STRICT_FP) || cfg->semantics == BACON_SEM_LITERAL) ? 1 : 0;
    *fn = entry->path_query[strict] ? entry->path_query[strict] : entry->path_query[1 - strict];
    if ((cfg->flags & BACON_FLAG_STRICT_FP) && !entry->path_query[1]) *fn = nullptr;  // strict was asked for by name
    if (!*fn)
        return fail(BACON_E_UNSUPPORTED, "rhs '%s' has no path-query kernels (%s build)", entry->name.c_str(),
                    strict ? "strict" : "fast");
    return 0;
}

int path_query_device(const bacon_ivp_config* cfg, int rhs_id, size_t n, const double* d_y0, const double* d_params,
                      const bacon_ivp_result* d_solved, const PathQuery& q, void* stream) {
    RhsEntry entry;
    bacon_path_fn fn = nullptr;
    int rc = path_query_check(cfg, rhs_id, d_y0, d_params, d_solved, q, &entry, &fn);
    if (rc != 0) return rc;
    g_last_launch = bacon_ivp_launch_info{};
    if (n == 0 || (q.op == BACON_PATH_SAMPLE && q.n_times == 0)) return 0;
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return fail(BACON_E_UNSUPPORTED, "device ordinal %d", dev);
    if (!g_tl.start[dev]) {
        CUDA_TRY(cudaEventCreate(&g_tl.start[dev]));
        CUDA_TRY(cudaEventCreate(&g_tl.stop[dev]));
    }
    bacon_path_args a{};
    a.cfg = *cfg;
    a.n = n;
    a.y0 = d_y0;
    a.params = d_params;
    a.hist = d_solved->hist;
    a.hist_len = d_solved->hist_len;
    if (d_solved->t_end && d_solved->y_end) {
        a.t_end = d_solved->t_end;
        a.y_end = d_solved->y_end;
    }
    a.n_accept = d_solved->n_accept;
    a.status = d_solved->status;
    a.t_start_each = d_solved->t_start;
    a.op = q.op;
    a.n_times = q.n_times;
    a.times = q.times;
    a.samples = q.samples;
    if (q.op == BACON_PATH_EVENTS) {
        for (int d = 0; d < cfg->dim; ++d) a.ev_w[d] = q.w[d];
        a.ev_c = q.c;
        a.ev_direction = q.direction;
        a.ev_capacity = q.capacity;
        a.events = q.events;
        a.n_events = q.n_events;
    }
    a.stream = stream;
    CUDA_TRY(cudaEventRecord(g_tl.start[dev], (cudaStream_t)stream));
    g_last_error.clear();
    rc = fn(&a);
    if (rc != 0) return launch_failed(rc, dev);
    CUDA_TRY(cudaEventRecord(g_tl.stop[dev], (cudaStream_t)stream));
    g_tl.last_dev = dev;
    g_tl.pending = true;
    g_last_launch.grid = a.grid;
    g_last_launch.block = a.block;
    g_last_launch.regs_per_thread = a.regs_per_thread;
    g_last_launch.n_kernels = 1;
    return 0;
}

// plain device allocation for the host variants (queries are not on the timed path of anything)
struct DevBlock {
    void* p = nullptr;
    ~DevBlock() {
        if (p) cudaFree(p);
    }
    int put(const void* host, size_t bytes) {
        CUDA_TRY(cudaMalloc(&p, bytes ? bytes : 1));
        if (host && bytes) CUDA_TRY(cudaMemcpy(p, host, bytes, cudaMemcpyHostToDevice));
        return 0;
    }
    int get(void* host, size_t bytes) const {
        if (host && bytes) CUDA_TRY(cudaMemcpy(host, p, bytes, cudaMemcpyDeviceToHost));
        return 0;
    }
};

int path_query_host(const bacon_ivp_config* cfg, int rhs_id, size_t n, const double* y0, const double* params,
                    const bacon_ivp_result* solved, const PathQuery& q) {
    RhsEntry entry;
    bacon_path_fn fn = nullptr;
    int rc = path_query_check(cfg, rhs_id, y0, params, solved, q, &entry, &fn);
    if (rc != 0) return rc;
    if (n == 0 || (q.op == BACON_PATH_SAMPLE && q.n_times == 0)) return 0;
    const size_t dim = (size_t)cfg->dim, cap = (size_t)cfg->history_capacity, np = (size_t)cfg->n_params;
    DevBlock b_y0, b_par, b_hist, b_len, b_tend, b_yend, b_times, b_out, b_cnt, b_acc, b_stat, b_t0;
    if ((rc = b_y0.put(y0, 8 * dim * n))) return rc;
    if (np > 0 && (rc = b_par.put(params, 8 * np * ((cfg->flags & BACON_FLAG_SHARED_PARAMS) ? 1 : n)))) return rc;
    if ((rc = b_hist.put(solved->hist, 8 * n * cap * (1 + dim)))) return rc;
    if ((rc = b_len.put(solved->hist_len, 4 * n))) return rc;
    bacon_ivp_result d{};
    d.hist = (double*)b_hist.p;
    d.hist_len = (uint32_t*)b_len.p;
    if (solved->t_end && solved->y_end) {
        if ((rc = b_tend.put(solved->t_end, 8 * n))) return rc;
        if ((rc = b_yend.put(solved->y_end, 8 * dim * n))) return rc;
        d.t_end = (double*)b_tend.p;
        d.y_end = (double*)b_yend.p;
    }
    if (solved->n_accept) {
        if ((rc = b_acc.put(solved->n_accept, 4 * n))) return rc;
        d.n_accept = (uint32_t*)b_acc.p;
    }
    if (solved->status) {
        if ((rc = b_stat.put(solved->status, 4 * n))) return rc;
        d.status = (int32_t*)b_stat.p;
    }
    if (solved->t_start) {
        if ((rc = b_t0.put(solved->t_start, 8 * n))) return rc;
        d.t_start = (const double*)b_t0.p;
    }
    PathQuery dq = q;
    size_t out_bytes = 0;
    if (q.op == BACON_PATH_SAMPLE) {
        if ((rc = b_times.put(q.times, 8 * q.n_times))) return rc;
        out_bytes = 8 * n * q.n_times * dim;
        if ((rc = b_out.put(nullptr, out_bytes))) return rc;
        dq.times = (const double*)b_times.p;
        dq.samples = (double*)b_out.p;
    } else {
        out_bytes = 8 * n * (size_t)q.capacity * (1 + dim);
        if ((rc = b_out.put(nullptr, out_bytes))) return rc;
        if ((rc = b_cnt.put(nullptr, 4 * n))) return rc;
        if (out_bytes) CUDA_TRY(cudaMemset(b_out.p, 0, out_bytes));
        dq.events = (double*)b_out.p;
        dq.n_events = (uint32_t*)b_cnt.p;
    }
    rc = path_query_device(cfg, rhs_id, n, (const double*)b_y0.p, (const double*)b_par.p, &d, dq, nullptr);
    if (rc != 0) return rc;
    CUDA_TRY(cudaStreamSynchronize(nullptr));
    if (q.op == BACON_PATH_SAMPLE) return b_out.get(q.samples, out_bytes);
    if ((rc = b_out.get(q.events, out_bytes))) return rc;
    return b_cnt.get(q.n_events, 4 * n);
}

}  // namespace
extern "C" {

int bacon_ivp_sample_paths_device(const bacon_ivp_config* cfg, int rhs_id, size_t n, const double* d_y0,
                                  const double* d_params, const bacon_ivp_result* d_solved, size_t n_times,
                                  const double* d_times, double* d_samples, void* stream) {
    PathQuery q{};
    q.op = BACON_PATH_SAMPLE;
    q.n_times = n_times;
    q.times = d_times;
    q.samples = d_samples;
    return path_query_device(cfg, rhs_id, n, d_y0, d_params, d_solved, q, stream);
}
int bacon_ivp_sample_paths(const bacon_ivp_config* cfg, int rhs_id, size_t n, const double* y0, const double* params,
                           const bacon_ivp_result* solved, size_t n_times, const double* times, double* samples) {
    PathQuery q{};
    q.op = BACON_PATH_SAMPLE;
    q.n_times = n_times;
    q.times = times;
    q.samples = samples;
    return path_query_host(cfg, rhs_id, n, y0, params, solved, q);
}
int bacon_ivp_locate_events_device(const bacon_ivp_config* cfg, int rhs_id, size_t n, const double* d_y0,
                                   const double* d_params, const bacon_ivp_result* d_solved, const double* w, double c,
                                   int direction, int capacity, double* d_events, uint32_t* d_n_events, void* stream) {
    PathQuery q{};
    q.op = BACON_PATH_EVENTS;
    q.w = w;
    q.c = c;
    q.direction = direction;
    q.capacity = capacity;
    q.events = d_events;
    q.n_events = d_n_events;
    return path_query_device(cfg, rhs_id, n, d_y0, d_params, d_solved, q, stream);
}
int bacon_ivp_locate_events(const bacon_ivp_config* cfg, int rhs_id, size_t n, const double* y0, const double* params,
                            const bacon_ivp_result* solved, const double* w, double c, int direction, int capacity,
                            double* events, uint32_t* n_events) {
    PathQuery q{};
    q.op = BACON_PATH_EVENTS;
    q.w = w;
    q.c = c;
    q.direction = direction;
    q.capacity = capacity;
    q.events = events;
    q.n_events = n_events;
    return path_query_host(cfg, rhs_id, n, y0, params, solved, q);
}

void* bacon_host_alloc(size_t bytes) {
    if (bytes == 0) bytes = 1;
    const size_t want = align_up(bytes, (size_t)1 << 16);
    PinnedPool& P = pinned_pool();
    std::lock_guard<std::mutex> lk(P.mu);
    auto it = P.free_blocks.find(want);
    void* p = nullptr;
    if (it != P.free_blocks.end()) {
        p = it->second;
        P.free_blocks.erase(it);
        P.cached -= want;
    } else {
        cudaError_t e = cudaHostAlloc(&p, want, cudaHostAllocPortable | cudaHostAllocMapped);
        if (e != cudaSuccess) {
            // drop the cache and retry once
            for (auto& kv : P.free_blocks) cudaFreeHost(kv.second);
            P.free_blocks.clear();
            P.cached = 0;
            cudaGetLastError();
            e = cudaHostAlloc(&p, want, cudaHostAllocPortable | cudaHostAllocMapped);
        }
        if (e != cudaSuccess) {
            fail(BACON_E_CUDA, "cudaHostAlloc(%zu) failed: %s", want, cudaGetErrorString(e));
            return nullptr;
        }
    }
    P.live[p] = want;
    return p;
}

void bacon_host_free(void* p) {
    if (!p) return;
    PinnedPool& P = pinned_pool();
    std::lock_guard<std::mutex> lk(P.mu);
    auto it = P.live.find(p);
    if (it == P.live.end()) return;  // not ours
    const size_t sz = it->second;
    P.live.erase(it);
    if (P.cached + sz <= PinnedPool::kMaxCached) {
        P.free_blocks.emplace(sz, p);
        P.cached += sz;
    } else {
        cudaFreeHost(p);
    }
}

int bacon_ivp_last_launch(bacon_ivp_launch_info* out) {
    if (!out) return fail(BACON_E_BAD_ARGUMENT, "NULL argument");
    if (g_tl.pending && g_tl.last_dev >= 0) {  // device entry point: resolve the event pair lazily
        int dev = 0;
        CUDA_TRY(cudaGetDevice(&dev));
        CUDA_TRY(cudaSetDevice(g_tl.last_dev));
        CUDA_TRY(cudaEventSynchronize(g_tl.stop[g_tl.last_dev]));
        float ms = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&ms, g_tl.start[g_tl.last_dev], g_tl.stop[g_tl.last_dev]));
        CUDA_TRY(cudaSetDevice(dev));
        g_last_launch.kernel_ms = ms;
        g_tl.pending = false;
    }
    *out = g_last_launch;
    return 0;
}

int bacon_device_sm_count(void) {
    int dev = 0, sm = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1;
    if (cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
    return sm;
}

}  // extern "C"

// ---------------------------------------------------------------- FP64 peak probe
// Register-resident DFMA chains: the denominator of the RK kernels' roofline (MEASURED_PEAKS.json
// has HBM and bf16 only).  16 independent chains per thread, 8 resident warps per SM sub-partition,
// x = fma(x, a, x): two distinct register sources per DFMA, which is the form that reaches the pipe's
// full rate (tools/fp64_peak.cu: 37.0 TFLOP/s = 99.5 % of 148 SM x 64 DFMA/clk x 1.965 GHz on this
// pool's B200s; DFMAs with three distinct register sources top out near 34 TFLOP/s).
namespace {
constexpr int kPeakChains = 16;
__global__ void __launch_bounds__(256) fp64_peak_kernel(double* sink, int iters, double a) {
    double x[kPeakChains];
#pragma unroll
    for (int k = 0; k < kPeakChains; ++k) x[k] = 1e-3 * (threadIdx.x + k + 1);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < kPeakChains; ++k) x[k] = fma(x[k], a, x[k]);
    }
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < kPeakChains; ++k) s += x[k];
    if (s == 123.456) sink[0] = s;  // never true; keeps the chains alive
}
}  // namespace

extern "C" double bacon_fp64_peak_tflops(int iters, void* stream) {
    if (iters < 1) iters = 4096;
    int dev = 0, sm = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1.0;
    if (cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1.0;
    double* sink = nullptr;
    if (cudaMalloc(&sink, 8) != cudaSuccess) return -1.0;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = sm * 4, block = 256;
    fp64_peak_kernel<<<grid, block, 0, st>>>(sink, 64, -1e-9);  // warm-up
    double best = -1.0;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0, st);
        fp64_peak_kernel<<<grid, block, 0, st>>>(sink, iters, -1e-9);
        cudaEventRecord(e1, st);
        if (cudaEventSynchronize(e1) != cudaSuccess) { best = -1.0; break; }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double flops = 2.0 * kPeakChains * (double)iters * (double)grid * block;
        const double tf = flops / (ms * 1e-3) / 1e12;
        if (tf > best) best = tf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    return best;
}
