// rtc.cu — user right-hand sides compiled at run time: bacon_rhs_register_source.
//
// The registration-macro path (include/bacon_ivp_rhs.cuh) needs nvcc when the user's crate is built.  This path needs
// nothing but the library: the functor arrives as CUDA C++ source text, NVRTC compiles it TOGETHER with the kernel
// headers (embedded in the library at build time, build/embedded_headers.inc), so the right-hand side is inlined into
// the stage loops exactly as for a built-in — the same ensemble_kernel template, instantiated
// on demand: one NVRTC program per (method, strict/fast, dense output or not) the caller actually uses, cached per
// device.  libnvrtc and libcuda are loaded lazily (dlopen), so the library still loads where they are absent
// (BACON_E_UNSUPPORTED from this entry point only).  Launching mirrors launch.cuh through the driver API.
#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nvrtc.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/bacon_ivp.h"
#include "ivp_common.cuh"
#include "path_query.cuh"
#include "tableaux.cuh"

namespace bacon_internal { void set_last_error(const char* msg); }  // engine.cu

namespace {

struct EmbeddedHeader { const char* name; const char* text; };
const EmbeddedHeader kHeaders[] = {
#include "build/embedded_headers.inc"
};
constexpr int kNumHeaders = sizeof(kHeaders) / sizeof(kHeaders[0]);

int rtc_fail(int code, const char* fmt, ...) {
    char buf[4096];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    bacon_internal::set_last_error(buf);
    return code;
}

// ---------------------------------------------------------------- lazily loaded NVRTC + driver API
#define RTC_SYMS(X) X(nvrtcCreateProgram) X(nvrtcCompileProgram) X(nvrtcGetProgramLogSize) X(nvrtcGetProgramLog) \
    X(nvrtcGetCUBINSize) X(nvrtcGetCUBIN) X(nvrtcAddNameExpression) X(nvrtcGetLoweredName) X(nvrtcDestroyProgram)
#define DRV_SYMS(X) X(cuModuleLoadData) X(cuModuleGetFunction) X(cuModuleGetGlobal_v2) X(cuLaunchKernel)         \
    X(cuFuncGetAttribute) X(cuFuncSetAttribute) X(cuOccupancyMaxActiveBlocksPerMultiprocessor) X(cuMemcpyHtoDAsync_v2) X(cuGetErrorString)
struct Api {
    bool ok = false, drv_ok = false;  // NVRTC alone is enough to compile (and to reject) a source; launching needs the driver
    std::string why;
#define DECL(name) decltype(&::name) name = nullptr;
    RTC_SYMS(DECL)
    DRV_SYMS(DECL)
#undef DECL
};
Api& api() {
    static Api a;
    static std::once_flag once;
    std::call_once(once, [] {
        void* hn = nullptr;
        for (const char* n : {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12"})
            if ((hn = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break;
        void* hd = nullptr;
        for (const char* n : {"libcuda.so.1", "libcuda.so"})
            if ((hd = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break;
        if (!hn) { a.why = "libnvrtc.so.12 not found"; return; }
#define LOAD(h, name)                                                      \
    a.name = reinterpret_cast<decltype(a.name)>(dlsym(h, #name));          \
    if (!a.name) { a.why = "symbol " #name " not found"; return; }
#define LOADN(name) LOAD(hn, name)
#define LOADD(name) LOAD(hd, name)
        RTC_SYMS(LOADN)
        a.ok = true;
        if (!hd) { a.why = "libcuda.so.1 not found (no NVIDIA driver)"; return; }
        DRV_SYMS(LOADD)
#undef LOADN
#undef LOADD
#undef LOAD
        a.drv_ok = true;
    });
    return a;
}

// ---------------------------------------------------------------- one registered source
struct Compiled {  // one NVRTC program: the kernels of one (method, strict, newton, hist) for one architecture
    std::vector<char> cubin;
    std::string main, tail, tableau;  // lowered names ("" = not in this program; tail: the events kernel of a path program)
    int block = 128;                  // CTA size the ensemble kernel was instantiated for
    size_t smem = 0;                  // its dynamic shared memory (exchange buffer of the regrouping)
};
struct Variant {  // ... loaded on one device
    CUfunction main = nullptr, tail = nullptr;
    int block = 128;
    size_t smem = 0;
    CUdeviceptr tableau = 0;  // strict RK: address of bacon::c_rk_tab in this module
};
struct RtcRhs {
    std::string name, type_name, source;
    int dim = 0, n_params = 0;
    std::mutex mu;
    std::map<std::vector<int>, Compiled> programs;  // key: arch (major*10+minor), method, strict, newton, hist
    std::map<std::vector<int>, Variant> variants;   // key: device, method, strict, newton, hist
};
constexpr int kMaxRtc = 32;
std::mutex g_mu;
std::vector<std::unique_ptr<RtcRhs>> g_rtc;

const char* tab_name(int method) { return (method == BACON_RK45) ? "bacon::TabRKF45" : "bacon::TabBS23"; }

// what launch.cuh would instantiate for this call: stepper type, CTA size, resident CTAs per SM the kernel is compiled
// for, and the exchange buffer of the end-of-ensemble regrouping (drive.cuh)
struct Plan { std::string stepper; int minb; int block; size_t smem; int state_doubles; };
int make_plan(const RtcRhs& r, const bacon_ivp_config& c, bool strict, bool hist, bool fit, bool event, Plan* p) {
    const std::string T = r.type_name;
    const bool newton = (c.flags & BACON_FLAG_BDF_NEWTON) != 0;
    p->state_doubles = 0;
    switch (c.method) {
        case BACON_RK45:
        case BACON_RK23: {
            const int O = c.method == BACON_RK45 ? 6 : 4;
            if (strict) {
                p->stepper = "bacon::RkStrictStepper<" + T + ", " + std::to_string(O) + ">";
                p->minb = 1;
            } else {
                if (c.semantics != BACON_SEM_CORRECTED) return BACON_E_UNSUPPORTED;
                p->stepper = "bacon::RkFastStepper<" + T + ", " + tab_name(c.method) + ">";
                p->minb = r.dim * (O + 1) + r.n_params <= 28 ? 6 : 4;                 // rk_fast_minb (launch.cuh)
                p->state_doubles = r.dim + r.n_params + 3;                             // RkFastStepper::STATE_DOUBLES
            }
            break;
        }
        case BACON_BDF6:
        case BACON_BDF2: {
            if (c.semantics != BACON_SEM_CORRECTED && !(strict && !newton)) return BACON_E_UNSUPPORTED;  // launch_bdf
            if (newton && strict) return BACON_E_UNSUPPORTED;
            const char* coef = c.method == BACON_BDF6 ? "bacon::CoefBDF6" : "bacon::CoefBDF2";
            p->stepper = "bacon::BdfStepper<" + T + ", " + coef + ", " + (strict ? "true" : "false") + ", " +
                         (newton ? "true" : "false") + ">";
            p->minb = newton ? 4 : 2;
            break;
        }
        case BACON_ADAMS5:
        case BACON_ADAMS3:
            p->stepper = "bacon::AdamsStepper<" + T + ", " + (c.method == BACON_ADAMS5 ? "bacon::CoefAdams5" : "bacon::CoefAdams3") +
                         ", " + (strict ? "true" : "false") + ">";
            p->minb = 2;
            break;
        case BACON_EULER:
            p->stepper = "bacon::EulerStepper<" + T + ", " + (strict ? "true" : "false") + ">";
            p->minb = 4;
            break;
        default:
            return BACON_E_BAD_ARGUMENT;
    }
    if (hist && p->minb >= 6) p->minb -= 1;  // MINB_HIST (launch.cuh)
    if (fit && !hist && !event && p->state_doubles > 0 && p->minb >= 6 && p->minb < 8) p->minb += 1;  // fits_one_more_warp (launch.cuh)
    // launch_stepper_hist (launch.cuh): steppers that suspend run as one wide CTA per SM when its exchange buffer fits
    // (terminal-event kernels never do)
    p->block = 128;
    p->smem = 0;
    if (p->state_doubles > 0 && !event) {
        const int wide = 128 * p->minb;
        if (wide >= 256 && (size_t)(p->state_doubles + 1) * 8 * wide <= 160 * 1024) {  // StepperMigrates (drive.cuh)
            p->block = wide;
            p->minb = 1;
            p->smem = (size_t)(p->state_doubles + 1) * 8 * wide;
        }
    }
    return 0;
}

int compile_program(RtcRhs& r, const bacon_ivp_config& c, bool strict, bool hist, bool fit, bool event, int cc_major, int cc_minor, Compiled* out) {
    Api& A = api();
    Plan plan;
    if (const int rc = make_plan(r, c, strict, hist, fit, event, &plan))
        return rtc_fail(rc, "rhs '%s': method %d / semantics %d / flags 0x%x has no %s kernel", r.name.c_str(), c.method,
                        c.semantics, c.flags, strict ? "strict" : "fast");
    std::string src = "#include \"drive.cuh\"\n#include \"rk_fast.cuh\"\n#include \"rk_strict.cuh\"\n#include \"adams.cuh\"\n";
    src += "#line 1 \"" + r.name + ".cu\"\n" + r.source + "\n";
    src += "static_assert(" + r.type_name + "::DIM == " + std::to_string(r.dim) + " && " + r.type_name +
           "::NPARAM == " + std::to_string(r.n_params) + ", \"DIM / NPARAM of the functor differ from the registration\");\n";
    const std::string h = hist ? "true" : "false", m = std::to_string(plan.minb);
    const std::string k_main = "&bacon::ensemble_kernel<" + plan.stepper + ", " + h + ", " + std::to_string(plan.block) + ", " + m +
                               (event ? ", true>" : ", false>");
    const std::string k_tab = "&bacon::c_rk_tabs";
    const bool want_tab = strict && (c.method == BACON_RK45 || c.method == BACON_RK23);

    std::vector<const char*> hdr_text, hdr_name;
    for (int i = 0; i < kNumHeaders; ++i) { hdr_text.push_back(kHeaders[i].text); hdr_name.push_back(kHeaders[i].name); }
    nvrtcProgram prog = nullptr;
    if (A.nvrtcCreateProgram(&prog, src.c_str(), (r.name + ".cu").c_str(), kNumHeaders, hdr_text.data(), hdr_name.data()) != NVRTC_SUCCESS)
        return rtc_fail(BACON_E_CUDA, "nvrtcCreateProgram failed");
    A.nvrtcAddNameExpression(prog, k_main.c_str());
    if (want_tab) A.nvrtcAddNameExpression(prog, k_tab.c_str());
    const std::string arch = "--gpu-architecture=sm_" + std::to_string(cc_major) + std::to_string(cc_minor) + (cc_major >= 9 ? "a" : "");
    std::vector<const char*> opts = {arch.c_str(), "-std=c++17", "-default-device", "-lineinfo"};
    if (strict) { opts.push_back("--fmad=false"); opts.push_back("-DBACON_STRICT_FP"); }
    const nvrtcResult cr = A.nvrtcCompileProgram(prog, (int)opts.size(), opts.data());
    if (cr != NVRTC_SUCCESS) {
        size_t n = 0;
        A.nvrtcGetProgramLogSize(prog, &n);
        std::string log(n, '\0');
        if (n) A.nvrtcGetProgramLog(prog, &log[0]);
        A.nvrtcDestroyProgram(&prog);
        return rtc_fail(BACON_E_USER, "rhs '%s' does not compile:\n%.3500s", r.name.c_str(), log.c_str());
    }
    size_t nbin = 0;
    A.nvrtcGetCUBINSize(prog, &nbin);
    out->cubin.resize(nbin);
    A.nvrtcGetCUBIN(prog, out->cubin.data());
    const char* l = nullptr;
    A.nvrtcGetLoweredName(prog, k_main.c_str(), &l);
    out->main = l ? l : "";
    out->block = plan.block;
    out->smem = plan.smem;
    if (want_tab && A.nvrtcGetLoweredName(prog, k_tab.c_str(), &l) == NVRTC_SUCCESS) out->tableau = l;
    A.nvrtcDestroyProgram(&prog);
    return 0;
}

int load_variant(const RtcRhs& r, const Compiled& p, Variant* out) {
    Api& A = api();
    cudaFree(nullptr);  // make sure the primary context of the current device exists and is current
    CUmodule mod = nullptr;
    CUresult e = A.cuModuleLoadData(&mod, p.cubin.data());
    auto drv_fail = [&](const char* what) {
        const char* s = nullptr;
        A.cuGetErrorString(e, &s);
        return rtc_fail(BACON_E_CUDA, "rhs '%s': %s failed: %s", r.name.c_str(), what, s ? s : "?");
    };
    if (e != CUDA_SUCCESS) return drv_fail("cuModuleLoadData");
    if ((e = A.cuModuleGetFunction(&out->main, mod, p.main.c_str())) != CUDA_SUCCESS) return drv_fail("cuModuleGetFunction");
    if (!p.tail.empty() && (e = A.cuModuleGetFunction(&out->tail, mod, p.tail.c_str())) != CUDA_SUCCESS)
        return drv_fail("cuModuleGetFunction(tail)");
    out->block = p.block;
    out->smem = p.smem;
    if (p.smem > 32 * 1024 &&
        (e = A.cuFuncSetAttribute(out->main, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)p.smem)) != CUDA_SUCCESS)
        return drv_fail("cuFuncSetAttribute(max dynamic shared memory)");
    if (!p.tableau.empty()) {
        size_t bytes = 0;
        if ((e = A.cuModuleGetGlobal_v2(&out->tableau, &bytes, mod, p.tableau.c_str())) != CUDA_SUCCESS) return drv_fail("cuModuleGetGlobal");
    }
    return 0;
}

// ---------------------------------------------------------------- launch (mirrors launch_stepper_hist, launch.cuh)
int rtc_launch(int slot, bacon_launch_args* a) {
    Api& A = api();
    RtcRhs* r = nullptr;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        if (slot < 0 || slot >= (int)g_rtc.size()) return BACON_E_BAD_ARGUMENT;
        r = g_rtc[slot].get();
    }
    const bacon_ivp_config& c = a->cfg;
    const bool strict = (c.flags & BACON_FLAG_STRICT_FP) || c.semantics == BACON_SEM_LITERAL;
    const bool hist = c.history_capacity > 0 && a->out.hist;
    const bool event = a->ev_on != 0 || a->t0_each || a->dt0_each || c.dt_init > 0.0;  // the kernels compiled for the optional inputs
    int dev = 0;
    cudaGetDevice(&dev);
    if (!A.drv_ok) return rtc_fail(BACON_E_UNSUPPORTED, "rhs '%s' cannot be launched: %s", r->name.c_str(), A.why.c_str());
    Variant v;
    {
        std::lock_guard<std::mutex> lk(r->mu);
        const int newton = (c.flags & BACON_FLAG_BDF_NEWTON) ? 1 : 0;
        // the window of sizes served by the kernel compiled for one more warp per sub-partition (fits_one_more_warp, launch.cuh)
        bool fit = false;
        if (!strict && !hist && !event && (c.method == BACON_RK45 || c.method == BACON_RK23) && a->grid_override <= 0 && !getenv("BACON_IVP_NO_FIT")) {
            Plan base;
            if (make_plan(*r, c, strict, hist, false, false, &base) == 0 && base.block >= 768 && base.block + 128 <= 1024 &&
                (size_t)(base.state_doubles + 1) * 8 * (base.block + 128) <= 160 * 1024)
                fit = a->n > (unsigned long long)a->sm_count * base.block && a->n <= (unsigned long long)a->sm_count * (base.block + 128);
        }
        const std::vector<int> key = {dev, (int)c.method, strict ? 1 : 0, newton, hist ? 1 : 0, fit ? 1 : 0, event ? 1 : 0};
        auto it = r->variants.find(key);
        if (it == r->variants.end()) {
            cudaDeviceProp prop;
            if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return BACON_E_CUDA;
            const std::vector<int> pkey = {prop.major * 10 + prop.minor, (int)c.method, strict ? 1 : 0, newton, hist ? 1 : 0, fit ? 1 : 0, event ? 1 : 0};
            auto pit = r->programs.find(pkey);
            if (pit == r->programs.end()) {
                Compiled cp;
                if (const int rc = compile_program(*r, c, strict, hist, fit, event, prop.major, prop.minor, &cp)) return rc;
                pit = r->programs.emplace(pkey, std::move(cp)).first;
            }
            Variant nv;
            if (const int rc = load_variant(*r, pit->second, &nv)) return rc;
            it = r->variants.emplace(key, nv).first;
        }
        v = it->second;
    }
    const int BLOCK = v.block;
    int per_sm = 0, regs = 0;
    if (A.cuOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, v.main, BLOCK, v.smem) != CUDA_SUCCESS || per_sm < 1) return BACON_E_CUDA;
    A.cuFuncGetAttribute(&regs, CU_FUNC_ATTRIBUTE_NUM_REGS, v.main);
    if (const char* env = getenv("BACON_IVP_BLOCKS_PER_SM")) {
        const int want = atoi(env);
        if (want >= 1 && want < per_sm) per_sm = want;
    }
    long long grid = (long long)per_sm * a->sm_count;
    const long long need = (long long)((a->n + 31) / 32);  // one bundle of 32 trajectories per warp (persistent_grid)
    if (grid > need) grid = need;
    if (a->grid_override > 0) grid = a->grid_override;
    if (const char* env = getenv("BACON_IVP_GRID")) {
        const int want = atoi(env);
        if (want >= 1 && want < grid) grid = want;
    }
    if (grid < 1) grid = 1;
    a->grid = (int)grid;
    a->block = BLOCK;
    a->regs_per_thread = regs;
    a->n_kernels = 1;
    a->late_from = (unsigned long long)grid * BLOCK;
    CUstream st = (CUstream)a->stream;
    if (v.tableau) {  // strict RK: the tableau as the stepper's row_iter() sees it, either semantics (launch_rk_strict)
        bacon::RkTableauRt T;
        if (c.method == BACON_RK45) bacon::fill_runtime_tableau<bacon::TabRKF45>(T, c.semantics == BACON_SEM_LITERAL);
        else bacon::fill_runtime_tableau<bacon::TabBS23>(T, c.semantics == BACON_SEM_LITERAL);
        // (pageable source: the driver stages the bytes before it returns)
        const size_t slot = (size_t)bacon::rk_tab_slot(c.method == BACON_RK45 ? 6 : 4, c.semantics);
        if (A.cuMemcpyHtoDAsync_v2(v.tableau + slot * sizeof(T), &T, sizeof(T), st) != CUDA_SUCCESS) return BACON_E_CUDA;
    }
    bacon_launch_args args = *a;
    void* p_main[] = {&args};
    if (A.cuLaunchKernel(v.main, (unsigned)grid, 1, 1, BLOCK, 1, 1, (unsigned)v.smem, st, p_main, nullptr) != CUDA_SUCCESS) return BACON_E_CUDA;
    return 0;
}

// ---------------------------------------------------------------- path queries (path_query.cuh) for a runtime-compiled functor
// One more NVRTC program per (architecture, strict/fast): the sampling and the events kernel instantiated on the user's
// functor.  Compiled keeps them as main (sampling) / tail (events).
constexpr int kPathKey = 1000;  // in the method slot of the cache keys
int compile_path_program(RtcRhs& r, bool strict, int cc_major, int cc_minor, Compiled* out) {
    Api& A = api();
    std::string src = "#include \"path_query.cuh\"\n";
    src += "#line 1 \"" + r.name + ".cu\"\n" + r.source + "\n";
    const std::string s = strict ? "true" : "false", T = r.type_name;
    const std::string k_sample = "&bacon::path_sample_kernel<" + T + ", " + s + ">";
    const std::string k_events = "&bacon::path_events_kernel<" + T + ", " + s + ", bacon::LaneLocate<" + T + " > >";
    std::vector<const char*> hdr_text, hdr_name;
    for (int i = 0; i < kNumHeaders; ++i) { hdr_text.push_back(kHeaders[i].text); hdr_name.push_back(kHeaders[i].name); }
    nvrtcProgram prog = nullptr;
    if (A.nvrtcCreateProgram(&prog, src.c_str(), (r.name + "_paths.cu").c_str(), kNumHeaders, hdr_text.data(), hdr_name.data()) != NVRTC_SUCCESS)
        return rtc_fail(BACON_E_CUDA, "nvrtcCreateProgram failed");
    A.nvrtcAddNameExpression(prog, k_sample.c_str());
    A.nvrtcAddNameExpression(prog, k_events.c_str());
    const std::string arch = "--gpu-architecture=sm_" + std::to_string(cc_major) + std::to_string(cc_minor) + (cc_major >= 9 ? "a" : "");
    std::vector<const char*> opts = {arch.c_str(), "-std=c++17", "-default-device", "-lineinfo"};
    if (strict) { opts.push_back("--fmad=false"); opts.push_back("-DBACON_STRICT_FP"); }
    if (A.nvrtcCompileProgram(prog, (int)opts.size(), opts.data()) != NVRTC_SUCCESS) {
        size_t n = 0;
        A.nvrtcGetProgramLogSize(prog, &n);
        std::string log(n, '\0');
        if (n) A.nvrtcGetProgramLog(prog, &log[0]);
        A.nvrtcDestroyProgram(&prog);
        return rtc_fail(BACON_E_USER, "rhs '%s': the path-query kernels do not compile:\n%.3500s", r.name.c_str(), log.c_str());
    }
    size_t nbin = 0;
    A.nvrtcGetCUBINSize(prog, &nbin);
    out->cubin.resize(nbin);
    A.nvrtcGetCUBIN(prog, out->cubin.data());
    const char* l = nullptr;
    if (A.nvrtcGetLoweredName(prog, k_sample.c_str(), &l) == NVRTC_SUCCESS && l) out->main = l;
    if (A.nvrtcGetLoweredName(prog, k_events.c_str(), &l) == NVRTC_SUCCESS && l) out->tail = l;
    A.nvrtcDestroyProgram(&prog);
    if (out->main.empty() || out->tail.empty()) return rtc_fail(BACON_E_CUDA, "rhs '%s': path-query kernel names not found", r.name.c_str());
    return 0;
}

// mirrors launch_path_query (path_query.cuh)
int rtc_path_query(int slot, bacon_path_args* a) {
    Api& A = api();
    RtcRhs* r = nullptr;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        if (slot < 0 || slot >= (int)g_rtc.size()) return BACON_E_BAD_ARGUMENT;
        r = g_rtc[slot].get();
    }
    const bool strict = (a->cfg.flags & BACON_FLAG_STRICT_FP) || a->cfg.semantics == BACON_SEM_LITERAL;
    int dev = 0;
    cudaGetDevice(&dev);
    if (!A.drv_ok) return rtc_fail(BACON_E_UNSUPPORTED, "rhs '%s' cannot be launched: %s", r->name.c_str(), A.why.c_str());
    Variant v;
    {
        std::lock_guard<std::mutex> lk(r->mu);
        const std::vector<int> key = {dev, kPathKey, strict ? 1 : 0, 0, 0};
        auto it = r->variants.find(key);
        if (it == r->variants.end()) {
            cudaDeviceProp prop;
            if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return BACON_E_CUDA;
            const std::vector<int> pkey = {prop.major * 10 + prop.minor, kPathKey, strict ? 1 : 0, 0, 0};
            auto pit = r->programs.find(pkey);
            if (pit == r->programs.end()) {
                Compiled cp;
                if (const int rc = compile_path_program(*r, strict, prop.major, prop.minor, &cp)) return rc;
                pit = r->programs.emplace(pkey, std::move(cp)).first;
            }
            Variant nv;
            if (const int rc = load_variant(*r, pit->second, &nv)) return rc;
            it = r->variants.emplace(key, nv).first;
        }
        v = it->second;
    }
    CUfunction fn = nullptr;
    unsigned long long blocks = 0;
    if (a->op == BACON_PATH_SAMPLE) {
        fn = v.main;
        blocks = (a->n * a->n_times + bacon::PATH_BLOCK - 1) / bacon::PATH_BLOCK;
    } else if (a->op == BACON_PATH_EVENTS) {
        fn = v.tail;
        blocks = (a->n * 32 + bacon::PATH_BLOCK - 1) / bacon::PATH_BLOCK;
    }
    if (!fn || blocks == 0 || blocks > 0x7fffffffull) return BACON_E_BAD_ARGUMENT;
    int regs = 0;
    A.cuFuncGetAttribute(&regs, CU_FUNC_ATTRIBUTE_NUM_REGS, fn);
    bacon_path_args args = *a;
    void* params[] = {&args};
    if (A.cuLaunchKernel(fn, (unsigned)blocks, 1, 1, bacon::PATH_BLOCK, 1, 1, 0, (CUstream)a->stream, params, nullptr) != CUDA_SUCCESS)
        return BACON_E_CUDA;
    a->grid = (int)blocks;
    a->block = bacon::PATH_BLOCK;
    a->regs_per_thread = regs;
    return 0;
}
template <int SLOT> int path_trampoline(bacon_path_args* a) { return rtc_path_query(SLOT, a); }
template <int... I> void fill_path_trampolines(bacon_path_fn (&t)[kMaxRtc], std::integer_sequence<int, I...>) {
    ((t[I] = &path_trampoline<I>), ...);
}

// bacon_launch_fn carries no context: a fixed pool of trampolines, one per runtime-compiled right-hand side
template <int SLOT> int trampoline(bacon_launch_args* a) { return rtc_launch(SLOT, a); }
template <int... I> void fill_trampolines(bacon_launch_fn (&t)[kMaxRtc], std::integer_sequence<int, I...>) {
    ((t[I] = &trampoline<I>), ...);
}

}  // namespace

extern "C" int bacon_rhs_register_source(const char* name, const char* type_name, const char* source, int dim, int n_params) {
    if (!name || !type_name || !source || dim < 1 || n_params < 0)
        return -rtc_fail(BACON_E_BAD_ARGUMENT, "bacon_rhs_register_source: name, type_name, source, dim >= 1, n_params >= 0");
    Api& A = api();
    if (!A.ok) return -rtc_fail(BACON_E_UNSUPPORTED, "runtime compilation is not available: %s", A.why.c_str());
    static bacon_launch_fn tramps[kMaxRtc];
    static bacon_path_fn path_tramps[kMaxRtc];
    static std::once_flag once;
    std::call_once(once, [] {
        fill_trampolines(tramps, std::make_integer_sequence<int, kMaxRtc>{});
        fill_path_trampolines(path_tramps, std::make_integer_sequence<int, kMaxRtc>{});
    });
    int slot;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        if ((int)g_rtc.size() >= kMaxRtc) return -rtc_fail(BACON_E_UNSUPPORTED, "at most %d runtime-compiled right-hand sides", kMaxRtc);
        slot = (int)g_rtc.size();
        g_rtc.emplace_back(new RtcRhs());
        RtcRhs& r = *g_rtc.back();
        r.name = name; r.type_name = type_name; r.source = source; r.dim = dim; r.n_params = n_params;
    }
    // compile the default kernel now, so that a functor that does not compile is reported here and not at the first solve
    {
        bacon_ivp_config c{};
        c.method = BACON_RK45;
        c.dim = dim;
        c.n_params = n_params;
        c.semantics = BACON_SEM_CORRECTED;
        int major = 10, minor = 0, dev = 0;  // no device in sight: check the source against the architecture this library is for
        cudaDeviceProp prop;
        if (cudaGetDevice(&dev) == cudaSuccess && cudaGetDeviceProperties(&prop, dev) == cudaSuccess) {
            major = prop.major;
            minor = prop.minor;
        } else {
            (void)cudaGetLastError();
        }
        RtcRhs& r = *g_rtc[slot];
        std::lock_guard<std::mutex> lk(r.mu);
        Compiled cp;
        if (const int rc = compile_program(r, c, false, false, false, false, major, minor, &cp)) {
            std::lock_guard<std::mutex> lk2(g_mu);
            if ((int)g_rtc.size() == slot + 1) g_rtc.back()->source.clear();  // (the slot stays: trampolines are positional)
            return -rc;
        }
        r.programs.emplace(std::vector<int>{major * 10 + minor, (int)BACON_RK45, 0, 0, 0, 0, 0}, std::move(cp));
        // the path-query program is compiled at the first query; BACON_RTC_EAGER_PATHS=1 compiles it here as well (both
        // flavours), which is how the CPU test suite checks that path_query.cuh goes through NVRTC
        if (getenv("BACON_RTC_EAGER_PATHS")) {
            for (int strict = 0; strict < 2; ++strict) {
                Compiled pp;
                if (const int rc = compile_path_program(r, strict != 0, major, minor, &pp)) return -rc;
                r.programs.emplace(std::vector<int>{major * 10 + minor, kPathKey, strict, 0, 0}, std::move(pp));
            }
        }
    }
    bacon_rhs_desc d{};
    d.name = name;
    d.dim = dim;
    d.n_params = n_params;
    for (int s = 0; s < 2; ++s)
        for (int m = 0; m < BACON_N_METHODS; ++m) d.launch[s][m] = d.launch_event[s][m] = tramps[slot];  // (rtc_launch reads a->ev_on)
    d.path_query[0] = d.path_query[1] = path_tramps[slot];
    return bacon_rhs_register(&d);
}
