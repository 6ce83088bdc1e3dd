// Host-side view of the compiled kernel instances (one translation unit per (ND, NP)).
#pragma once
#include <cuda_runtime.h>
#include "stage_kernel.cuh"

namespace flou {

struct StageLauncher {
    cudaError_t (*launch)(const KParams &, cudaStream_t);         // fused single-kernel stage
    cudaError_t (*launch_elements)(const KParams &, cudaStream_t); // two-kernel path: volume+lift+RK, node per thread
    cudaError_t (*launch_lines)(const KParams &, cudaStream_t);    // two-kernel path: volume+lift+RK, line per thread
    cudaError_t (*launch_faces)(const KParams &, cudaStream_t);    // two-kernel path: Riemann fluxes
    cudaError_t (*prepare)();
    int (*resident)();          // persistent grid size (CTAs per SM x SMs)
    int epb, threads;
    size_t smem;
    int line_e, line_t;         // line kernel: elements and threads per CTA
    size_t line_smem;
    int (*line_resident)();     // persistent grid of the line kernel
};

struct EmitLauncher {
    cudaError_t (*launch)(const double *u, int64_t ndof, const int *list, int nslots,
                          int faces_per_elem, int colloc,
                          const double *lm, const double *lp, double *out, cudaStream_t);
};

// eq: EQ_*, vol: VOL_*, cart: 0/1.  Returns nullptr when the combination is not compiled.
const StageLauncher *get_stage_launcher(int nd, int np, int eq, int vol, int cart);
const EmitLauncher *get_emit_launcher(int nd, int np, int nv);

}  // namespace flou
