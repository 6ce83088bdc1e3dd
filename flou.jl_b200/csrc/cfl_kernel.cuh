// get_max_dt on the device (row f1): min over all nodes of cfl*dx_e / (|v| + c) for Euler
// (FlouCommon/Euler.jl:116-135: |u| in 1-D, sqrt(u^2+v^2[+w^2]) in 2-D/3-D) or cfl*dx_e/|a| for
// linear advection (LinearAdvection.jl:46-48); dx_e = (volume_e/npts)^(1/nd)
// (MultielementDiscontinuous.jl:162-178).  Positive doubles order like their bit patterns, so
// the global minimum is an atomicMin on the 64-bit pattern.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace flou {

__global__ void max_dt_kernel(const double *__restrict__ u, int64_t ndof, int npts, int nd, int euler,
                              double gamma, double anorm, double cfl,
                              const double *__restrict__ elem_dx, double cart_dx,
                              unsigned long long *__restrict__ out)
{
    double best = __longlong_as_double(0x7ff0000000000000LL);     // +inf
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < ndof;
         i += (int64_t)gridDim.x * blockDim.x) {
        const double dx = elem_dx ? elem_dx[i / npts] : cart_dx;
        double dt;
        if (euler) {
            const double rho = u[i];
            double m2 = 0.0;
            for (int d = 0; d < nd; d++) { const double m = u[i + ndof * (1 + d)]; m2 += m * m; }
            const double p = (gamma - 1.0) * (u[i + ndof * (nd + 1)] - m2 / (2.0 * rho));
            const double c = sqrt(gamma * p / rho);
            const double speed = (nd == 1) ? fabs(u[i + ndof] / rho) : sqrt(m2 / (rho * rho));
            dt = cfl * dx / (speed + c);
        } else {
            dt = cfl * dx / anorm;
        }
        if (dt < best) best = dt;      // NaN never wins: a crashed state is reported by status()
    }
    for (int o = 16; o > 0; o >>= 1) best = fmin(best, __shfl_xor_sync(0xffffffffu, best, o));
    if ((threadIdx.x & 31) == 0 && best > 0.0) atomicMin(out, (unsigned long long)__double_as_longlong(best));
}

}  // namespace flou
