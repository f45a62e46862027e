// rhs_builtin.cu — instantiates and registers the kernels of the built-in right-hand
// sides.  Compiled twice: as is (fast: FMA contraction, compile-time tableaux) and with
// -DBACON_STRICT_FP -fmad=false (strict: the oracle's operation order, no contraction).
// User RHS go through exactly the same two lines: see INTEGRATION.md.
#include "launch.cuh"
#include "rhs_builtin.cuh"

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
