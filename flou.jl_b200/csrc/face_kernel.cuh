// Face-flux kernel of the two-kernel stage: interface_fluxes! + applyBCs! evaluated ONCE per
// face point (src/FlouSpatial/Interfaces.jl:111-136 and :25-49), one thread per (face, master
// face dof).  Output: Fn[slot][v][i] = rotate2phys(F*(rotate2face(Ql), rotate2face(Qr)))
// * jac (slot stride fn_block(nv, nfp): 16-byte aligned blocks for the element kernel's TMA copies),
// i.e. the reference's `Fn.sides[1]` (master-outward); the slave side is its negative with
// the `master2slave` permutation and is applied by the element kernel when it copies the block.
//
// Traces (project2faces!, Interfaces.jl:51-109) are not materialised for every face: with
// collocated (GLL) nodes the x-face traces come from the trace array the stage kernel writes
// with the state (or from u as well when the handle keeps no array: P.tr_in == nullptr), y-/z-face
// traces are node layers read straight from u; partition-boundary
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
    } else if (lf < 2 && P.tr_in) {     // x-faces: trace array (Gauss nodes; collocated nodes when kept)
        const double *src = P.tr_in + ((int64_t)elem_or_slot * 2 + lf) * (NV * NFP) + k;
#pragma unroll
        for (int v = 0; v < NV; v++) Q[v] = __ldg(src + v * NFP);
    } else if (P.colloc) {      // node layer of u
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

// The kernel waits on two dependent DRAM round trips (face record, then traces): occupancy pays
// more than the spills it costs.  cfg4s, 3-D Euler p=4: 5 CTAs/SM (95 registers) 0.658 ms,
// 6 (80) 0.610, 7 (72) 0.583, 8 (64) 0.591.

// FLOU_FACE_PREFETCH: faces ahead whose record is pulled into L2 (cfg4s: 0.585 -> 0.556 ms for
// 2048..8192, nothing at 32768).  Prefetching the traces of a face further ahead as well
// (cp.async.bulk.prefetch.L2 per run) was slower, 0.67 ms: with the record prefetch the kernel moves
// 3.3 GB at ~5.9 TB/s, i.e. it is bound by DRAM traffic, of which the 40-byte runs of y-face node
// layers waste a third.
#ifndef FLOU_FACE_PREFETCH
#define FLOU_FACE_PREFETCH 4096
#endif
#ifndef FLOU_FACE_MIN_BLOCKS
#define FLOU_FACE_MIN_BLOCKS 7
#endif

// Rotation to / from the frame of a Cartesian face whose master-side local face is the
// compile-time LFM: a fixed signed permutation (PhysicalRegions.jl:541-696), no selects.
template <int ND, int EQ, int LFM>
__device__ __forceinline__ void rot2face_c(const double *Q, double *R)
{
    constexpr int dm = LFM >> 1;
    constexpr double s = (LFM & 1) ? 1.0 : -1.0;
    R[0] = Q[0];
    if (EQ == EQ_ADV) return;
    R[1] = s * Q[1 + dm];
    if (ND == 2) R[2] = ((dm == 0) ? s : -s) * Q[1 + (1 - dm)];
    if (ND == 3) {
        constexpr int tm = (dm == 2) ? 0 : dm + 1, bm = (dm == 0) ? 2 : dm - 1;
        R[2] = s * Q[1 + tm];
        R[3] = Q[1 + bm];
    }
    R[ND + 1] = Q[ND + 1];
}

template <int ND, int EQ, int LFM>
__device__ __forceinline__ void rot2phys_c(const double *R, double *Q)
{
    constexpr int dm = LFM >> 1;
    constexpr double s = (LFM & 1) ? 1.0 : -1.0;
    Q[0] = R[0];
    if (EQ == EQ_ADV) return;
    if (ND == 1) Q[1] = s * R[1];
    if (ND == 2) {
        Q[1 + dm] = s * R[1];
        Q[1 + (1 - dm)] = ((dm == 0) ? s : -s) * R[2];
    }
    if (ND == 3) {
        constexpr int tm = (dm == 2) ? 0 : dm + 1, bm = (dm == 0) ? 2 : dm - 1;
        Q[1 + dm] = s * R[1];
        Q[1 + tm] = s * R[2];
        Q[1 + bm] = R[3];
    }
    Q[ND + 1] = R[ND + 1];
}

// One face point with the master's local face LFM known at compile time: the master trace path
// (trace array / node layer) and, on Cartesian meshes, the whole rotation are static.
template <int ND, int NP, int EQ, bool CART, int LFM>
__device__ __forceinline__ void face_point(const KParams &P, int f, int i, const FaceRec rec)
{
    constexpr int NV = (EQ == EQ_ADV) ? 1 : ND + 2;
    constexpr int NFP = ipow_c(NP, ND - 1);
    const int lfs = (rec.info >> 3) & 7, orient = (rec.info >> 6) & 7;
    const int mkind = (rec.info >> 9) & 3, skind = (rec.info >> 11) & 3;
    const int j = master2slave<ND, NP>(i, orient);

    // frame and face Jacobian of the master side (PhysicalRegions.jl:541-696 / 797-971)
    double fr[CART ? 1 : 3 * ND], fj;
    if (CART) {
        fj = P.cfjac[LFM >> 1];
    } else {
        const int64_t fi = (int64_t)f * NFP + i;
#pragma unroll
        for (int c = 0; c < 3 * ND; c++)
            fr[c] = (c < ND * ND || ND == 3) ? __ldg(P.frames + fi + P.nfacedofs * c) : 0.0;
        fj = __ldg(P.fjac + fi);
    }

    double Ql[NV], Qr[NV];
    if (mkind == 1) load_trace<ND, NP, NV>(P, 1, rec.em, LFM, i, Ql);
    else load_trace<ND, NP, NV>(P, 0, rec.em, LFM, i, Ql);          // LFM static: one path survives
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
            if (CART) rot2face_c<ND, EQ, LFM>(Ql, R);
            else rotate2face<ND, EQ>(Ql, fr, R);
            if (NV > 1) R[NV > 1 ? 1 : 0] = -R[NV > 1 ? 1 : 0];
            if (CART) rot2phys_c<ND, EQ, LFM>(R, Qr);
            else rotate2phys<ND, EQ>(R, fr, Qr);
        } else {
#pragma unroll
            for (int v = 0; v < NV; v++) Qr[v] = P.bc_table[((int64_t)rec.es * NFP + i) * NV + v];
        }
    }

    double Rl[NV], Rr[NV], Fn[NV], Fp[NV];
    if (CART) {
        rot2face_c<ND, EQ, LFM>(Ql, Rl);
        rot2face_c<ND, EQ, LFM>(Qr, Rr);
    } else {
        rotate2face<ND, EQ>(Ql, fr, Rl);
        rotate2face<ND, EQ>(Qr, fr, Rr);
    }
    if (EQ == EQ_EULER) {
        euler_numflux<ND>(P.fp, Rl, Rr, Fn);
    } else {
        double an = 0.0;
        if (CART) an = ((LFM & 1) ? 1.0 : -1.0) * P.fp.a[LFM >> 1];
        else {
#pragma unroll
            for (int c = 0; c < ND; c++) an += P.fp.a[c] * fr[c];
        }
        Fn[0] = an * (Rl[0] + Rr[0]) * 0.5;
        if (P.fp.numflux == FX_LXF) Fn[0] += fabs(an) * (Rl[0] - Rr[0]) * 0.5 * P.fp.intensity;
    }
    if (CART) rot2phys_c<ND, EQ, LFM>(Fn, Fp);
    else rotate2phys<ND, EQ>(Fn, fr, Fp);
    double *dst = P.Fn + (int64_t)f * fn_block(NV, NFP) + i;
#pragma unroll
    for (int v = 0; v < NV; v++) dst[v * NFP] = Fp[v] * fj;
}

template <int ND, int NP, int EQ, bool CART>
__global__ void __launch_bounds__(128, FLOU_FACE_MIN_BLOCKS)
face_flux_kernel(const __grid_constant__ KParams P)
{
    pdl_prologue();
    constexpr int NFP = ipow_c(NP, ND - 1);
    // face_reverse: the kernel starts where the element kernel of the previous stage ended (the
    // tail of u_out / the traces is still in L2) and ends where the next element kernel starts
    const int64_t blk = P.face_reverse ? (int64_t)gridDim.x - 1 - blockIdx.x : blockIdx.x;
    const int64_t t = blk * blockDim.x + threadIdx.x;
    if (t >= (int64_t)P.face_count * NFP) return;
    const int fl = (int)(t / NFP), i = (int)(t - (int64_t)fl * NFP);
    const int f = P.face_first + fl;
#if FLOU_FACE_PREFETCH > 0
    // the record of a face a later CTA will work on: pulled into L2 now, so that the first of the
    // two dependent round trips of that thread (record, then traces) is an L2 hit
    if (i == 0) {
        const int fp = P.face_reverse ? fl - FLOU_FACE_PREFETCH : fl + FLOU_FACE_PREFETCH;
        if (fp >= 0 && fp < P.face_count) asm volatile("prefetch.global.L2 [%0];" ::"l"(P.faces + P.face_first + fp));
    }
#endif
    const FaceRec rec = P.faces[f];
    // faces keep Flou's global order, in which the master's local face changes rarely: the switch
    // is warp-uniform almost everywhere
    switch (rec.info & 7) {
    case 0: face_point<ND, NP, EQ, CART, 0>(P, f, i, rec); break;
    case 1: face_point<ND, NP, EQ, CART, 1>(P, f, i, rec); break;
    case 2: if (ND >= 2) face_point<ND, NP, EQ, CART, (ND >= 2 ? 2 : 0)>(P, f, i, rec); break;
    case 3: if (ND >= 2) face_point<ND, NP, EQ, CART, (ND >= 2 ? 3 : 0)>(P, f, i, rec); break;
    case 4: if (ND >= 3) face_point<ND, NP, EQ, CART, (ND >= 3 ? 4 : 0)>(P, f, i, rec); break;
    default: if (ND >= 3) face_point<ND, NP, EQ, CART, (ND >= 3 ? 5 : 0)>(P, f, i, rec); break;
    }
}

}  // namespace flou
