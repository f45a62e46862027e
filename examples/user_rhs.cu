// A user right-hand side plugged in through the C ABI (see include/bacon_ivp_rhs.cuh).
#include "bacon_ivp_rhs.cuh"

// Brusselator: x' = a + x^2 y - (b+1) x,  y' = b x - x^2 y ;  p = (a, b)
struct Brusselator {
    static constexpr int DIM = 2, NPARAM = 2;
    __device__ void operator()(double, const double (&y)[2], const double* p, double (&dy)[2]) const {
        const double xxy = (y[0] * y[0]) * y[1];
        dy[0] = (p[0] + xxy) - (p[1] + 1.0) * y[0];
        dy[1] = p[1] * y[0] - xxy;
    }
    __device__ void jac(double, const double (&y)[2], const double* p, double (&J)[2][2]) const {
        J[0][0] = 2.0 * y[0] * y[1] - (p[1] + 1.0);  J[0][1] = y[0] * y[0];
        J[1][0] = p[1] - 2.0 * y[0] * y[1];          J[1][1] = -(y[0] * y[0]);
    }
};
BACON_REGISTER_RHS(Brusselator, "brusselator");

// A functor WITHOUT jac: the Newton path falls back to central differences (bdf.rs:390-411 semantics).
struct Pendulum {  // theta'' = -(g/l) sin(theta) ;  p = (g/l)
    static constexpr int DIM = 2, NPARAM = 1;
    __device__ void operator()(double, const double (&y)[2], const double* p, double (&dy)[2]) const {
        dy[0] = y[1];
        dy[1] = -p[0] * sin(y[0]);
    }
};
BACON_REGISTER_RHS(Pendulum, "pendulum");
