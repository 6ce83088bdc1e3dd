// Face-flux kernel of the two-kernel stage: interface_fluxes! + applyBCs! evaluated ONCE per
// face point (src/FlouSpatial/Interfaces.jl:111-136 and :25-49), one thread per (face, master
// face dof).  Output: Fn[(slot*nv + v)*NFP + i] = rotate2phys(F*(rotate2face(Ql), rotate2face(Qr)))
// * jac, i.e. the reference's `Fn.sides[1]` (master-outward); the slave side is its negative with
// the `master2slave` permutation and is applied by the element kernel when it copies the block.
//
// Traces (project2faces!, Interfaces.jl:51-109) are not materialised for every face: with
// collocated (GLL) nodes the x-face traces come from the trace array the stage kernel writes
// with the state, y-/z-face traces are node layers read straight from u; partition-boundary
// faces read the halo buffer; Gauss nodes read the interpolated traces of emit_traces_kernel.
#pragma once
#include "stage_kernel.cuh"

namespace flou {

template <int ND, int NP, int NV>
__device__ __forceinline__ void load_trace(const KParams &P, int kind, int elem_or_slot, int lf, int k,
                                           double *Q)
{
    constexpr int NPTS = ipow_c(NP, ND), NFP = ipow_c(NP, ND - 1), NFACES = 2 * ND;
    if (kind == 1) {            // ghost: the remote element's own trace, its own face-dof order
        const double *src = P.ghost + (int64_t)elem_or_slot * (NV * NFP) + k;
#pragma unroll
        for (int v = 0; v < NV; v++) Q[v] = __ldg(src + v * NFP);
    } else if (lf < 2) {        // x-faces: trace array
        const double *src = P.tr_in + ((int64_t)elem_or_slot * 2 + lf) * (NV * NFP) + k;
#pragma unroll
        for (int v = 0; v < NV; v++) Q[v] = __ldg(src + v * NFP);
    } else if (P.colloc) {      // y-/z-faces: node layer of u
        int nb, ns;
        line_of<ND, NP>(lf >> 1, k, nb, ns);
        const double *src = P.u_in + (int64_t)elem_or_slot * NPTS + nb + ((lf & 1) ? (NP - 1) * ns : 0);
#pragma unroll
        for (int v = 0; v < NV; v++) Q[v] = __ldg(src + P.ndof * v);
    } else {                    // Gauss nodes: interpolated traces
        const double *src = P.tr_hi + ((int64_t)elem_or_slot * NFACES + lf) * (NV * NFP) + k;
#pragma unroll
        for (int v = 0; v < NV; v++) Q[v] = __ldg(src + v * NFP);
    }
}

#ifndef FLOU_FACE_MIN_BLOCKS
#define FLOU_FACE_MIN_BLOCKS 5
#endif

template <int ND, int NP, int EQ, bool CART>
__global__ void __launch_bounds__(128, FLOU_FACE_MIN_BLOCKS)
face_flux_kernel(const __grid_constant__ KParams P)
{
    constexpr int NV = (EQ == EQ_ADV) ? 1 : ND + 2;
    constexpr int NFP = ipow_c(NP, ND - 1);
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)P.face_count * NFP) return;
    const int fl = (int)(t / NFP), i = (int)(t - (int64_t)fl * NFP);
    const int f = P.face_first + fl;
    const FaceRec rec = P.faces[f];
    const int lfm = rec.info & 7, lfs = (rec.info >> 3) & 7, orient = (rec.info >> 6) & 7;
    const int mkind = (rec.info >> 9) & 3, skind = (rec.info >> 11) & 3;
    const int j = master2slave<ND, NP>(i, orient);

    // frame and face Jacobian of the master side (PhysicalRegions.jl:541-696 / 797-971)
    double fr[CART ? 1 : 3 * ND], fj;
    int dm = 0;
    double sn_ = 1.0;
    if (CART) {
        dm = lfm >> 1;
        sn_ = (lfm & 1) ? 1.0 : -1.0;
        fj = dm == 0 ? P.cfjac[0] : (dm == 1 ? P.cfjac[1] : P.cfjac[2]);
    } else {
        const int64_t fi = (int64_t)f * NFP + i;
#pragma unroll
        for (int c = 0; c < 3 * ND; c++)
            fr[c] = (c < ND * ND || ND == 3) ? __ldg(P.frames + fi + P.nfacedofs * c) : 0.0;
        fj = __ldg(P.fjac + fi);
    }

    double Ql[NV], Qr[NV];
    load_trace<ND, NP, NV>(P, mkind, rec.em, lfm, i, Ql);
    if (skind != 2) {
        load_trace<ND, NP, NV>(P, skind, rec.es, lfs, j, Qr);
    } else {
        // boundary face: exterior state from the BC functor (Interfaces.jl:44-48)
        const int ib = rec.info >> 13;
        const int bk = P.bc_kind[ib];
        if (bk == FLOU_B200_BC_INFLOW) {
#pragma unroll
            for (int v = 0; v < NV; v++) Qr[v] = P.bc_state[ib * NV + v];
        } else if (bk == FLOU_B200_BC_OUTFLOW) {
#pragma unroll
            for (int v = 0; v < NV; v++) Qr[v] = Ql[v];
        } else if (bk == FLOU_B200_BC_SLIP) {
            double R[NV];
            if (CART) rotate2face_cart<ND, EQ>(Ql, dm, sn_, R);
            else rotate2face<ND, EQ>(Ql, fr, R);
            if (NV > 1) R[NV > 1 ? 1 : 0] = -R[NV > 1 ? 1 : 0];
            if (CART) rotate2phys_cart<ND, EQ>(R, dm, sn_, Qr);
            else rotate2phys<ND, EQ>(R, fr, Qr);
        } else {
#pragma unroll
            for (int v = 0; v < NV; v++) Qr[v] = P.bc_table[((int64_t)rec.es * NFP + i) * NV + v];
        }
    }

    double Rl[NV], Rr[NV], Fn[NV], Fp[NV];
    if (CART) {
        rotate2face_cart<ND, EQ>(Ql, dm, sn_, Rl);
        rotate2face_cart<ND, EQ>(Qr, dm, sn_, Rr);
    } else {
        rotate2face<ND, EQ>(Ql, fr, Rl);
        rotate2face<ND, EQ>(Qr, fr, Rr);
    }
    if (EQ == EQ_EULER) {
        euler_numflux<ND>(P.fp, Rl, Rr, Fn);
    } else {
        double an = 0.0;
        if (CART) an = sn_ * pick<ND>(P.fp.a, dm);
        else {
#pragma unroll
            for (int c = 0; c < ND; c++) an += P.fp.a[c] * fr[c];
        }
        Fn[0] = an * (Rl[0] + Rr[0]) * 0.5;
        if (P.fp.numflux == FX_LXF) Fn[0] += fabs(an) * (Rl[0] - Rr[0]) * 0.5 * P.fp.intensity;
    }
    if (CART) rotate2phys_cart<ND, EQ>(Fn, dm, sn_, Fp);
    else rotate2phys<ND, EQ>(Fn, fr, Fp);
    double *dst = P.Fn + (int64_t)f * (NV * NFP) + i;
#pragma unroll
    for (int v = 0; v < NV; v++) dst[v * NFP] = Fp[v] * fj;
}

}  // namespace flou
