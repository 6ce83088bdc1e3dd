// Kernel instances for one (ND, NP) pair; compiled once per pair with
//   -DFLOU_ND=<1|2|3> -DFLOU_NP=<2..8>
// so the instances build in parallel.  dispatch.cu stitches the per-pair tables together.
#include <cstdlib>
#include "launch.h"
#include "face_kernel.cuh"
#include "line_kernel.cuh"
#include "line_kernel_ws.cuh"

#ifndef FLOU_GRID_MULT_DEFAULT
#define FLOU_GRID_MULT_DEFAULT 1
#endif

#ifndef FLOU_ND
#error "compile with -DFLOU_ND=.. -DFLOU_NP=.."
#endif

namespace flou {

template <class C>
static int resident_ctas();
template <class C>
static cudaError_t line_prepare();

template <class C>
static cudaError_t do_prepare()
{
    {
        // 3-D np = 8 on general geometry: one element of the node-per-thread kernels needs more than
        // an SM's shared memory (split form 228 / 232 KB); flou_b200_create then uses the line kernel
        int dev = 0, optin = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        if (optin > 0 && C::SMEM_BYTES > (size_t)optin) return line_prepare<C>();
    }
    cudaError_t e = cudaFuncSetAttribute(stage_kernel<C, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)C::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(stage_kernel<C, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)C::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    // ask for the largest shared-memory carve-out so several CTAs fit per SM
    e = cudaFuncSetAttribute(stage_kernel<C, false>, cudaFuncAttributePreferredSharedMemoryCarveout,
                             cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(stage_kernel<C, true>, cudaFuncAttributePreferredSharedMemoryCarveout,
                             cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    resident_ctas<C>();      // occupancy query outside any stream capture
    return line_prepare<C>();
}

// persistent grid: (CTAs that fit per SM) x (SM count), queried once per kernel instance
template <class C>
static int resident_ctas()
{
    static int n = 0;
    if (n == 0) {
        int dev = 0, sms = 0, per_sm = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, stage_kernel<C, true>, C::THREADS, C::SMEM_BYTES);
        n = (per_sm > 0 ? per_sm : 1) * (sms > 0 ? sms : 1);
    }
    return n;
}

// FLOU_B200_PDL=1: stage kernels launched as programmatic dependents of the kernel before them in the
// stream (pdl_prologue in the kernels); captured into the CUDA graph as programmatic edges
static bool pdl_enabled()
{
    static const bool v = [] { const char *e = std::getenv("FLOU_B200_PDL"); return e && e[0] == '1'; }();
    return v;
}

static cudaError_t launch_stage_kernel(void (*kernel)(KParams), int grid, int threads, size_t smem, cudaStream_t s,
                                       const KParams &P)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, P);
}

template <class C>
static cudaError_t do_launch(const KParams &P, cudaStream_t s)
{
    if (P.elem_count <= 0) return cudaSuccess;
    const int ngroups = (P.elem_count + C::EPB - 1) / C::EPB;
    const int grid = ngroups;      // one CTA per group of EPB consecutive elements
    return launch_stage_kernel(stage_kernel<C, false>, grid, C::THREADS, C::SMEM_BYTES, s, P);
}

template <class C>
static cudaError_t do_launch_elements(const KParams &P, cudaStream_t s)
{
    if (P.elem_count <= 0) return cudaSuccess;
    const int grid = (P.elem_count + C::EPB - 1) / C::EPB;
    return launch_stage_kernel(stage_kernel<C, true>, grid, C::THREADS, C::SMEM_BYTES, s, P);
}

// ---- line-per-thread element kernel (element kernel of the two-kernel stage): warp-specialised,
// TL line threads + one update warp
template <class C>
using LineOf = LCfg<C::ND, C::NP, C::EQ, C::VOL, C::CART, true, C::NB>;
// the kernel symbol of an instance: launch-bounds build or explicit register cap (only the one used
// is instantiated)
template <class L>
static constexpr auto line_kernel_of()
{
    if constexpr (L::MAXREG > 0) return &line_kernel_ws_mr<L>;
    else return &line_kernel_ws<L>;
}

template <class C>
static int line_resident_ctas()
{
    static int n = 0;
    if (n == 0) {
        using L = LineOf<C>;
        int dev = 0, sms = 0, per_sm = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, line_kernel_of<L>(), L::T, L::SMEM_BYTES);
        n = (per_sm > 0 ? per_sm : 1) * (sms > 0 ? sms : 1);
    }
    return n;
}

template <class C>
static cudaError_t line_prepare()
{
    using L = LineOf<C>;
    {
        // a group of one element can exceed the shared memory of an SM (3-D, np = 8: Euler strong
        // form 278 KB, StdAverage split form 230 KB): the host then uses the node-per-thread element
        // kernel (flou_b200_create compares line_smem with the same limit); nothing to prepare
        int dev = 0, optin = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        if (optin > 0 && L::SMEM_BYTES > (size_t)optin) return cudaSuccess;
    }
    cudaError_t e = cudaFuncSetAttribute(line_kernel_of<L>(), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)L::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(line_kernel_of<L>(), cudaFuncAttributePreferredSharedMemoryCarveout,
                             cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    line_resident_ctas<C>();
    return cudaSuccess;
}

template <class C>
static cudaError_t do_launch_lines(const KParams &P0, cudaStream_t s)
{
    using L = LineOf<C>;
    if (P0.elem_count <= 0) return cudaSuccess;
    KParams P = P0;
    const int ngroups = (P.elem_count + L::E - 1) / L::E;
    const int resident = line_resident_ctas<C>();
    const int grid = ngroups < resident ? ngroups : resident;      // persistent CTAs
    return launch_stage_kernel(line_kernel_of<L>(), grid, L::T, L::SMEM_BYTES, s, P);
}

template <class C>
static cudaError_t do_launch_faces(const KParams &P, cudaStream_t s)
{
    if (P.face_count <= 0) return cudaSuccess;
    const int64_t n = (int64_t)P.face_count * C::NFP;
    const int grid = (int)((n + 127) / 128);
    return launch_stage_kernel(face_flux_kernel<C::ND, C::NP, C::EQ, C::CART>, grid, 128, 0, s, P);
}

template <class C>
static constexpr StageLauncher make()
{
    return StageLauncher{&do_launch<C>, &do_launch_elements<C>, &do_launch_lines<C>, &do_launch_faces<C>, &do_prepare<C>,
                         &resident_ctas<C>, C::EPB, C::THREADS, C::SMEM_BYTES,
                         LineOf<C>::E, LineOf<C>::T, LineOf<C>::SMEM_BYTES, &line_resident_ctas<C>};
}

// operators that exist in the line-per-thread formulation only (HybridDivOperator): two-kernel
// stage, no fused / node-per-thread kernels
template <class C>
static cudaError_t lines_only_prepare() { return line_prepare<C>(); }
template <class C>
static constexpr StageLauncher make_lines()
{
    return StageLauncher{nullptr, nullptr, &do_launch_lines<C>, &do_launch_faces<C>, &lines_only_prepare<C>,
                         &line_resident_ctas<C>, LineOf<C>::E, LineOf<C>::T, LineOf<C>::SMEM_BYTES,
                         LineOf<C>::E, LineOf<C>::T, LineOf<C>::SMEM_BYTES, &line_resident_ctas<C>};
}

// configuration tag of the line-only instances for nodes without boundaries (split form, Gauss)
template <int ND_, int NP_, int EQ_, int VOL_, bool CART_>
struct NBCfg : KCfg<ND_, NP_, EQ_, VOL_, CART_> {
    static constexpr bool NB = true;
};

#define ND FLOU_ND
#define NP FLOU_NP

// [eq][vol][cart]; vol 4 / 5 = split form (StdAverage / Chandrasekhar), 6 = hybrid operator on nodes
// without boundaries.  The instances of one (ND, NP) pair are spread over three translation units
// (-DFLOU_PART=0|1|2) so that the build parallelises: 0 = strong / split form (all three kernel
// families) and the trace kernels, 1 = hybrid operator, 2 = split form on Gauss nodes.
#ifndef FLOU_PART
#define FLOU_PART 0
#endif
#define FLOU_NONE {StageLauncher{}, StageLauncher{}}
static const StageLauncher table[2][7][2] = {
    {   // linear advection: strong, split (StdAverage two-point flux); no Chandrasekhar
#if FLOU_PART == 0
        {make<KCfg<ND, NP, EQ_ADV, VOL_STRONG, false>>(), make<KCfg<ND, NP, EQ_ADV, VOL_STRONG, true>>()},
        {make<KCfg<ND, NP, EQ_ADV, VOL_SPLIT_STD, false>>(), make<KCfg<ND, NP, EQ_ADV, VOL_SPLIT_STD, true>>()},
#else
        FLOU_NONE, FLOU_NONE,
#endif
        FLOU_NONE, FLOU_NONE, FLOU_NONE, FLOU_NONE, FLOU_NONE,
    },
    {   // Euler
#if FLOU_PART == 0
        {make<KCfg<ND, NP, EQ_EULER, VOL_STRONG, false>>(), make<KCfg<ND, NP, EQ_EULER, VOL_STRONG, true>>()},
        {make<KCfg<ND, NP, EQ_EULER, VOL_SPLIT_STD, false>>(), make<KCfg<ND, NP, EQ_EULER, VOL_SPLIT_STD, true>>()},
        {make<KCfg<ND, NP, EQ_EULER, VOL_SPLIT_CHA, false>>(), make<KCfg<ND, NP, EQ_EULER, VOL_SPLIT_CHA, true>>()},
#else
        FLOU_NONE, FLOU_NONE, FLOU_NONE,
#endif
#if FLOU_PART == 1
        // HybridDivOperator (general geometry reads the sub-grid tables)
        {make_lines<KCfg<ND, NP, EQ_EULER, VOL_HYBRID, false>>(), make_lines<KCfg<ND, NP, EQ_EULER, VOL_HYBRID, true>>()},
#else
        FLOU_NONE,
#endif
#if FLOU_PART == 2
        // split form on Gauss nodes
        {make_lines<NBCfg<ND, NP, EQ_EULER, VOL_SPLIT_STD, false>>(), make_lines<NBCfg<ND, NP, EQ_EULER, VOL_SPLIT_STD, true>>()},
        {make_lines<NBCfg<ND, NP, EQ_EULER, VOL_SPLIT_CHA, false>>(), make_lines<NBCfg<ND, NP, EQ_EULER, VOL_SPLIT_CHA, true>>()},
#else
        FLOU_NONE, FLOU_NONE,
#endif
#if FLOU_PART == 1
        {make_lines<NBCfg<ND, NP, EQ_EULER, VOL_HYBRID, false>>(), make_lines<NBCfg<ND, NP, EQ_EULER, VOL_HYBRID, true>>()},
#else
        FLOU_NONE,
#endif
    },
};

template <int NV>
static cudaError_t emit_launch(const double *u, int64_t ndof, const int *list, int nslots,
                               int faces_per_elem, int colloc, const double *lm, const double *lp,
                               double *out, cudaStream_t s)
{
    if (nslots <= 0) return cudaSuccess;
    constexpr int NFP = ipow_c(NP, ND - 1);
    const int64_t n = (int64_t)nslots * NFP;
    const int threads = 128;
    const int grid = (int)((n + threads - 1) / threads);
    emit_traces_kernel<ND, NP, NV><<<grid, threads, 0, s>>>(u, ndof, list, nslots, faces_per_elem, colloc, lm, lp, out);
    return cudaGetLastError();
}

#if FLOU_PART == 0
static const EmitLauncher emit_adv = {&emit_launch<1>};
static const EmitLauncher emit_euler = {&emit_launch<ND + 2>};
#endif

#define CAT_(a, b, c, p) a##b##_##c##_p##p
#define CAT(a, b, c, p) CAT_(a, b, c, p)

const StageLauncher *CAT(stage_table_, FLOU_ND, FLOU_NP, FLOU_PART)(int eq, int vol, int cart)
{
    const StageLauncher *l = &table[eq][vol][cart ? 1 : 0];
    return (l->launch || l->launch_lines) ? l : nullptr;
}

#if FLOU_PART == 0
const EmitLauncher *CAT(emit_table_, FLOU_ND, FLOU_NP, 0)(int nv)
{
    return nv == 1 ? &emit_adv : &emit_euler;
}
#endif

}  // namespace flou
