// rhs_builtin.cu — instantiates and registers the kernels of the built-in right-hand
// sides.  Compiled twice: as is (fast: FMA contraction, compile-time tableaux) and with
// -DBACON_STRICT_FP -fmad=false (strict: the oracle's operation order, no contraction).
// User RHS go through exactly the same two lines: see INTEGRATION.md.
// The build splits the instantiations into separate objects (families -DBACON_SKIP_RK / _BDF / _ADAMS, plain kernels
// -DBACON_SKIP_EVENTS against terminal-event kernels -DBACON_ONLY_EVENTS: see the Makefile) only to compile them in
// parallel; the registry merges launcher tables by RHS name.
#include "launch.cuh"
#include "rhs_builtin.cuh"
#include "path_query_warp.cuh"
#include "rk_warp_linear.cuh"

using namespace bacon;

BACON_REGISTER_RHS(RhsLorenz, "lorenz");
BACON_REGISTER_RHS(RhsVdp, "vdp");
BACON_REGISTER_RHS(RhsRobertson, "robertson");
BACON_REGISTER_RHS(RhsExp, "exp");
BACON_REGISTER_RHS(RhsDecay, "decay");
BACON_REGISTER_RHS(RhsQuadratic, "quadratic");
BACON_REGISTER_RHS(RhsCos, "cos");
BACON_REGISTER_RHS(RhsHarmonic, "harmonic");
BACON_REGISTER_RHS(RhsLinear<4>, "linear4");

#if !defined(BACON_SKIP_RK)
// linear32 (BASELINE config 4): warp-per-trajectory kernels (rk_warp_linear.cuh).  The plain kernels and the path
// queries live in the plain objects, the kernels that watch a terminal event in the event objects (Makefile).
namespace {
int register_linear32() {
    bacon_rhs_desc d{};
    d.name = "linear32";
    d.dim = 32;
    d.n_params = 32 * 32;
#ifdef BACON_STRICT_FP
    constexpr int S = 1;
#else
    constexpr int S = 0;
#endif
#ifndef BACON_ONLY_EVENTS
    d.launch[S][BACON_RK45] = &launch_rk_warp_linear32<TabRKF45, S != 0>;
    d.launch[S][BACON_RK23] = &launch_rk_warp_linear32<TabBS23, S != 0>;
    d.path_query[S] = &launch_path_query_linear32<S != 0>;
#endif
#ifndef BACON_SKIP_EVENTS
    d.launch_event[S][BACON_RK45] = &launch_rk_warp_linear32<TabRKF45, S != 0, true>;
    d.launch_event[S][BACON_RK23] = &launch_rk_warp_linear32<TabBS23, S != 0, true>;
#endif
    return bacon_rhs_register(&d);
}
const int bacon_rhs_id_linear32 = register_linear32();
}  // namespace
#endif

#if defined(BACON_DRIVE_TRACE) && !defined(BACON_SKIP_RK) && !defined(BACON_STRICT_FP)
// (diagnosis build only) the regrouping events CTA 0 recorded since the last call: (ns, tag or running warps, value)
extern "C" int bacon_debug_trace(unsigned long long* out, int cap) {
    unsigned int n = 0;
    cudaMemcpyFromSymbol(&n, g_trace_n, sizeof(n));
    if ((int)n > cap) n = (unsigned)cap;
    if (n > 4096) n = 4096;
    cudaMemcpyFromSymbol(out, g_trace, sizeof(unsigned long long) * 3 * n);
    const unsigned int zero = 0;
    cudaMemcpyToSymbol(g_trace_n, &zero, sizeof(zero));
    return (int)n;
}
#endif
