// Projection of a state to equispaced nodes for the VTKHDF snapshots: project2equispaced!
// (src/FlouSpatial/StdRegions/StdRegions.jl:232-238) applied per element and variable by
// pointdata2VTKHDF (src/FlouSpatial/IO.jl:78-97).  The reference multiplies by the dense
// node2eq = kron(M, M, M) (StdHex.jl:57-61, StdQuad.jl:52-53) of the 1-D interp_matrix M;
// here the three 1-D sums are nested (same result to round-off, NP^ND instead of NP^(2 ND) reads of M).
// One CTA per element (grid-stride), the element's nodal values of one variable in shared memory, a
// thread per equispaced point.  HBM-bound: reads the state once, writes the projected state once.
#pragma once
#include <cstdint>

namespace flou {

__global__ void __launch_bounds__(256)
project_equispaced_kernel(const double *__restrict__ u, int64_t ndof, int nv, int nd, int np, int neq,
                          const double *__restrict__ M /* [neq][np] */, int64_t ne,
                          double *__restrict__ out /* [v][ne * neq^nd] */)
{
    extern __shared__ double psm[];
    double *sM = psm, *sq = psm + neq * np;
    int npts = np, nout = neq;
    for (int d = 1; d < nd; d++) { npts *= np; nout *= neq; }
    for (int i = threadIdx.x; i < neq * np; i += blockDim.x) sM[i] = M[i];
    const int64_t npoints = ne * (int64_t)nout;
    for (int64_t e = blockIdx.x; e < ne; e += gridDim.x) {
        for (int v = 0; v < nv; v++) {
            __syncthreads();
            for (int n = threadIdx.x; n < npts; n += blockDim.x) sq[n] = u[e * npts + n + ndof * v];
            __syncthreads();
            for (int p = threadIdx.x; p < nout; p += blockDim.x) {
                const int ix = p % neq, iy = (p / neq) % neq, iz = p / (neq * neq);
                const double *mx = sM + ix * np, *my = sM + iy * np, *mz = sM + iz * np;
                double r = 0.0;
                if (nd == 1) {
                    for (int jx = 0; jx < np; jx++) r = fma(mx[jx], sq[jx], r);
                } else if (nd == 2) {
                    for (int jy = 0; jy < np; jy++) {
                        double s = 0.0;
                        for (int jx = 0; jx < np; jx++) s = fma(mx[jx], sq[jy * np + jx], s);
                        r = fma(my[jy], s, r);
                    }
                } else {
                    for (int jz = 0; jz < np; jz++) {
                        double t = 0.0;
                        for (int jy = 0; jy < np; jy++) {
                            double s = 0.0;
                            for (int jx = 0; jx < np; jx++) s = fma(mx[jx], sq[(jz * np + jy) * np + jx], s);
                            t = fma(my[jy], s, t);
                        }
                        r = fma(mz[jz], t, r);
                    }
                }
                out[(int64_t)v * npoints + e * nout + p] = r;
            }
        }
    }
}

}  // namespace flou
