// Monitors and the Zhang-Shu positivity limiter on the device.
//
//   kinetic_energy_monitor / entropy_monitor   src/FlouSpatial/Equations/Euler.jl:559-593
//       sum over elements of integrate(f(Q_i), geom_e) = sum_i (J w)_i f(Q_i)
//       (integrate: PhysicalRegions.jl:366-368, Jw = jac .* w :402,430,466)
//       f = kinetic_energy (FlouCommon/Euler.jl:162-175)  or  math_entropy (:213-217)
//   zhang_shu_limiter                          src/FlouSpatial/Equations/Euler.jl:616-660
//       per element: density scaled towards its mean so that rho >= min(minval, mean), then the
//       whole state scaled towards its mean so that p >= min(minval, mean pressure)
//
// The reference evaluates these on the host from `integrator.u` (callbacks FlouTime.jl:113-148;
// the limiter runs as `stage_limiter!` after every RK stage in examples/src/3D_Euler.jl:76-80):
// with the state resident on the device that would be a download + upload per stage.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace flou {

enum : int { MON_KINETIC_ENERGY = 0, MON_ENTROPY = 1 };

// pass 1: per-block partial sums in a fixed order (deterministic for a given grid)
__global__ void __launch_bounds__(256)
monitor_kernel(const double *__restrict__ u, int64_t ndof, int npts, int nd, int kind, double gamma,
               const double *__restrict__ w_nodes, const double *__restrict__ jac, double cjac,
               double *__restrict__ partial)
{
    double s = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < ndof;
         i += (int64_t)gridDim.x * blockDim.x) {
        const double jw = (jac ? jac[i] : cjac) * w_nodes[i % npts];
        const double rho = u[i];
        double m2 = 0.0;
        for (int d = 0; d < nd; d++) { const double m = u[i + ndof * (1 + d)]; m2 += m * m; }
        double f;
        if (kind == MON_KINETIC_ENERGY) {
            f = m2 / (2.0 * rho);
        } else {
            const double p = (gamma - 1.0) * (u[i + ndof * (nd + 1)] - m2 / (2.0 * rho));
            const double sp = log(p) - gamma * log(rho);          // entropy, Euler.jl:202-206
            f = -rho * sp / (gamma - 1.0);
        }
        s = fma(jw, f, s);
    }
    __shared__ double sw[8];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) sw[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int k = 0; k < (int)(blockDim.x >> 5); k++) t += sw[k];
        partial[blockIdx.x] = t;
    }
}

// pass 2: one block adds the partial sums, again in a fixed order
__global__ void __launch_bounds__(256)
monitor_reduce_kernel(const double *__restrict__ partial, int n, double *__restrict__ out)
{
    __shared__ double sw[256];
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) s += partial[i];
    sw[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) sw[threadIdx.x] += sw[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = sw[0];
}

__device__ __forceinline__ double warp_sum(double x)
{
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
}
__device__ __forceinline__ double warp_min(double x)
{
    for (int o = 16; o > 0; o >>= 1) x = fmin(x, __shfl_xor_sync(0xffffffffu, x, o));
    return x;
}

// Zhang-Shu limiter, one warp per element, in place.  The three sweeps re-read the element from
// global memory (its nv*npts doubles stay in L1/L2); every lane ends up with identical sums (the
// xor butterfly adds the same pairs in every lane), so the branch on theta is warp-uniform.
template <int NV>
__global__ void __launch_bounds__(256)
zhang_shu_kernel(double *__restrict__ u, int64_t ndof, int64_t nelem, int npts, double gamma,
                 double minval, const double *__restrict__ w_nodes, const double *__restrict__ jac,
                 double cjac)
{
    constexpr int ND = NV - 2;
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    for (int64_t e = warp0; e < nelem; e += nwarps) {
        const int64_t d0 = e * npts;
        // ---- mean state and minimum density
        double qs[NV], vol = 0.0, rmin = inf;
#pragma unroll
        for (int v = 0; v < NV; v++) qs[v] = 0.0;
        for (int i = lane; i < npts; i += 32) {
            const double jw = (jac ? jac[d0 + i] : cjac) * w_nodes[i];
            vol += jw;
#pragma unroll
            for (int v = 0; v < NV; v++) qs[v] = fma(jw, u[d0 + i + ndof * v], qs[v]);
            rmin = fmin(rmin, u[d0 + i]);
        }
        vol = warp_sum(vol);
        rmin = warp_min(rmin);
        double qb[NV];
#pragma unroll
        for (int v = 0; v < NV; v++) qb[v] = warp_sum(qs[v]) / vol;
        // ---- density limiting (Euler.jl:629-638)
        {
            const double m = fmin(minval, qb[0]);
            const double theta = fabs((qb[0] - m) / (qb[0] - rmin));
            if (theta <= 1.0)
                for (int i = lane; i < npts; i += 32) u[d0 + i] = theta * (u[d0 + i] - qb[0]) + qb[0];
        }
        __syncwarp();
        // ---- pressure of the (density-limited) nodes: mean and minimum (Euler.jl:640-648)
        double ps = 0.0, pmin = inf;
        for (int i = lane; i < npts; i += 32) {
            const double jw = (jac ? jac[d0 + i] : cjac) * w_nodes[i];
            const double rho = u[d0 + i];
            double m2 = 0.0;
#pragma unroll
            for (int d = 0; d < ND; d++) { const double mm = u[d0 + i + ndof * (1 + d)]; m2 += mm * mm; }
            const double p = (gamma - 1.0) * (u[d0 + i + ndof * (ND + 1)] - m2 / (2.0 * rho));
            ps = fma(jw, p, ps);
            pmin = fmin(pmin, p);
        }
        const double pb = warp_sum(ps) / vol;
        pmin = warp_min(pmin);
        {
            const double m = fmin(minval, pb);
            const double theta = fabs((pb - m) / (pb - pmin));
            if (theta <= 1.0)
                for (int i = lane; i < npts; i += 32) {
#pragma unroll
                    for (int v = 0; v < NV; v++) {
                        const double q = u[d0 + i + ndof * v];
                        u[d0 + i + ndof * v] = theta * (q - qb[v]) + qb[v];
                    }
                }
        }
        __syncwarp();
    }
}

}  // namespace flou
