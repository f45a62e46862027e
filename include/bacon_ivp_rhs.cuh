/* bacon_ivp_rhs.cuh — device-side plug-in header for USER right-hand sides.
 *
 * Replaces the reference's `Derivative` trait (src/ivp.rs:34-48): where bacon takes any
 * `FnMut(f64, &[N], &mut T) -> Result<BVector<N, D>, UserError>` through `with_derivative`
 * (src/ivp.rs:186), this engine takes a CUDA device functor compiled into the kernels, so the
 * stage loop keeps the state in registers and the RHS is inlined (a device function pointer
 * would force both out).  A user translation unit looks like
 *
 *     #include "bacon_ivp_rhs.cuh"
 *     struct Brusselator {                       // stateless; parameters arrive per trajectory
 *         static constexpr int DIM = 2, NPARAM = 2;
 *         __device__ void operator()(double t, const double (&y)[2], const double* p, double (&dy)[2]) const {
 *             dy[0] = p[0] + y[0] * y[0] * y[1] - (p[1] + 1.0) * y[0];
 *             dy[1] = p[1] * y[0] - y[0] * y[0] * y[1];
 *         }
 *         // optional: analytic Jacobian for BACON_FLAG_BDF_NEWTON (finite differences otherwise)
 *         // __device__ void jac(double t, const double (&y)[2], const double* p, double (&J)[2][2]) const;
 *     };
 *     BACON_REGISTER_RHS(Brusselator, "brusselator");
 *
 * built with
 *     nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -shared \
 *          -I<repo>/include user_rhs.cu -o libuser_rhs.so -L<repo>/bacon_b200 -lbacon_ivp
 * (add -DBACON_STRICT_FP -fmad=false in a second object for the strict, oracle-order kernels).
 * Loading the library registers the functor under its name through the C ABI
 * (bacon_rhs_register); `bacon_rhs_lookup("brusselator")` then works like a built-in.
 * The per-trajectory parameter block plays the role of the reference's cloned `UserData`
 * (rk.rs:380): read-only, no state carried between evaluations.
 */
#ifndef BACON_IVP_RHS_CUH
#define BACON_IVP_RHS_CUH
#include "bacon_ivp.h"
#include "../bacon_b200/csrc/launch.cuh"
#endif
