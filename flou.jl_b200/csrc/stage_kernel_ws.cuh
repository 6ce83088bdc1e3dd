// Warp-specialised, persistent form of the fused stage kernel (same arithmetic as
// stage_kernel.cuh, which documents the reference mapping).
//
// A CTA = NCONS consumer threads (thread <-> node of EPB elements) + ONE producer warp and
// walks a contiguous range of element groups.  The producer streams, NSTAGE groups ahead,
// everything a group needs from HBM into a shared-memory ring with cp.async (LDGSTS):
//   state u (nv x npts), tmp (nv x npts), the neighbours' face traces (x-faces from the trace
//   array, y/z faces straight from u, ghosts from the halo buffer) and the connectivity records,
// and signals an mbarrier per ring slot (cp.async.mbarrier.arrive).  Consumers never touch global
// memory for input: they wait on the slot's `full` barrier, run phases 1-4 out of shared
// memory, write u_out / tmp / x-face traces from registers, and release the slot (`empty`).
// DRAM latency therefore sits off the compute warps' critical path.
#pragma once
#include "stage_kernel.cuh"

namespace flou {

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity)
{
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile(
        "{\n.reg .pred p;\nWAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(a), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
// arrive on `bar` once every cp.async issued so far by this thread has completed
__device__ __forceinline__ void cp_async_mbar_arrive(uint64_t *bar)
{
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void cp_async4(void *smem_dst, const void *gmem_src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}

#ifndef FLOU_WS_STAGES
#define FLOU_WS_STAGES 2
#endif

template <class C>
struct WSCfg {
    static constexpr int NV = C::NV, NPTS = C::NPTS, NFP = C::NFP, NFACES = C::NFACES, EPB = C::EPB;
    static constexpr int NSTAGE = FLOU_WS_STAGES;
    static constexpr int NCONS = C::THREADS;            // consumer threads (multiple of 32)
    static constexpr int THREADS = NCONS + 32;          // + one producer warp
    // ring slot, per element (doubles)
    static constexpr int SQ = 0;                         // u      [v][node]
    static constexpr int ST = SQ + NV * NPTS;            // tmp    [v][node]
    static constexpr int SF = ST + NV * NPTS;            // neighbour traces, then face fluxes [lf][v][k]
    static constexpr int SC = SF + NFACES * NV * NFP;    // Conn[NFACES]
    static constexpr int STG_ELEM = SC + NFACES;
    static constexpr int STG = EPB * STG_ELEM;           // one ring slot
    // scratch, per element (doubles): primitives, strong-form fluxes, metric, exchange/acc
    static constexpr int SA = 0;
    static constexpr int SFT = SA + C::NAUX * NPTS;
    static constexpr int SM = SFT + C::NFT_VOL * NPTS;
    static constexpr int SX = SM + C::NMET * NPTS;
    static constexpr int SACC = C::ACC_ALIAS ? SX + (C::ND & 1) * (C::XROUNDS * NV * NPTS) : SX + C::NXCH * NPTS;
    static constexpr int SCR_ELEM = SX + (C::NXCH + C::NACC) * NPTS;
    static constexpr int BAR_DOUBLES = 2 * NSTAGE;      // full[], empty[]
    static constexpr size_t SMEM_BYTES =
        sizeof(double) * (size_t)(C::OPS + BAR_DOUBLES + NSTAGE * STG + EPB * SCR_ELEM);
    static constexpr int SMEM_BLOCKS = (int)((227 * 1024) / (SMEM_BYTES + 1024));
    static constexpr int REG_BLOCKS = 65536 / (THREADS * 128);
    static constexpr int MIN_BLOCKS_ =
        SMEM_BLOCKS < REG_BLOCKS ? SMEM_BLOCKS : REG_BLOCKS;
    static constexpr int MIN_BLOCKS = MIN_BLOCKS_ < 1 ? 1 : MIN_BLOCKS_;
};

template <class C>
__global__ void __launch_bounds__(WSCfg<C>::THREADS, WSCfg<C>::MIN_BLOCKS)
stage_kernel_ws(const __grid_constant__ KParams P)
{
    using W = WSCfg<C>;
    constexpr int ND = C::ND, NP = C::NP, EQ = C::EQ, VOL = C::VOL, NV = C::NV;
    constexpr int NPTS = C::NPTS, NFP = C::NFP, NFACES = C::NFACES, NFT = C::NFT, EPB = C::EPB;
    constexpr int TPT = C::TPT, NSTAGE = W::NSTAGE;
    constexpr bool CART = C::CART, SPLIT = C::SPLIT;

    extern __shared__ double smem[];
    double *sD = smem;                       // [ii + NP*jj]
    double *sLm = sD + NP * NP, *sLp = sLm + NP, *sGl = sLp + NP, *sGr = sGl + NP;
    uint64_t *full = reinterpret_cast<uint64_t *>(sGr + NP);
    uint64_t *empty = full + NSTAGE;
    double *ring = reinterpret_cast<double *>(empty + NSTAGE);
    double *scratch = ring + (size_t)NSTAGE * W::STG;

    const int tid = threadIdx.x;
    for (int i = tid; i < NP * NP; i += W::THREADS) sD[i] = P.Dvol[i];
    if (tid < NP) { sLm[tid] = P.lm[tid]; sLp[tid] = P.lp[tid]; sGl[tid] = P.dgl[tid]; sGr[tid] = P.dgr[tid]; }
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < NSTAGE; s++) { mbar_init(full + s, 32); mbar_init(empty + s, 1); }
    }
    __syncthreads();

    const int64_t ndof = P.ndof;
    const bool need_tmp = (P.mode == MODE_STAGE);
    // contiguous range of groups for this CTA
    const int ngroups = (P.elem_count + EPB - 1) / EPB;
    const int per_cta = (ngroups + (int)gridDim.x - 1) / (int)gridDim.x;
    const int gbeg = blockIdx.x * per_cta;
    const int gend = min(ngroups, gbeg + per_cta);
    auto elem_of = [&](int idx) { return P.elem_list ? P.elem_list[idx] : P.elem_first + idx; };

    if (tid >= W::NCONS) {
        // ======================= producer warp =======================
        const int lane = tid - W::NCONS;
        for (int g = gbeg, it = 0; g < gend; g++, it++) {
            const int s = it % NSTAGE;
            mbar_wait(empty + s, ((it / NSTAGE) & 1) ^ 1);
            double *slot = ring + (size_t)s * W::STG;
            const int nact = min(EPB, P.elem_count - g * EPB);
            for (int tel = 0; tel < nact; tel++) {
                const int te = elem_of(g * EPB + tel);
                double *dst = slot + (size_t)tel * W::STG_ELEM;
                const double *usrc = P.u_in + (int64_t)te * NPTS;
                const double *tsrc = P.tmp + (int64_t)te * NPTS;
                for (int i = lane; i < NV * NPTS; i += 32) {
                    const int v = i / NPTS, n = i - v * NPTS;
                    cp_async8(dst + W::SQ + i, usrc + n + ndof * v);
                    if (need_tmp) cp_async8(dst + W::ST + i, tsrc + n + ndof * v);
                }
                // connectivity records + the neighbours' traces
                for (int r = lane; r < NFT; r += 32) {
                    const int lf = r / NFP, k = r - lf * NFP;
                    const int2 c = __ldg(reinterpret_cast<const int2 *>(P.conn) + ((int64_t)te * NFACES + lf));
                    if (k == 0) reinterpret_cast<int2 *>(dst + W::SC)[lf] = c;
                    const int kind = (c.y >> 7) & 3;
                    if (kind == FK_BOUNDARY) continue;
                    const int nlf = c.y & 7, orient = (c.y >> 3) & 7;
                    const bool master = (c.y >> 6) & 1;
                    const int kn = master ? master2slave<ND, NP>(k, orient) : slave2master<ND, NP>(k, orient);
                    double *d2 = dst + W::SF + lf * NV * NFP + k;
                    if (kind == FK_GHOST) {
                        const double *src = P.ghost + (int64_t)c.x * (NV * NFP) + kn;
#pragma unroll
                        for (int v = 0; v < NV; v++) cp_async8(d2 + v * NFP, src + v * NFP);
                    } else if (nlf < 2) {
                        const double *src = P.tr_in + ((int64_t)c.x * 2 + nlf) * (NV * NFP) + kn;
#pragma unroll
                        for (int v = 0; v < NV; v++) cp_async8(d2 + v * NFP, src + v * NFP);
                    } else if (P.colloc) {
                        int nb, ns;
                        line_of<ND, NP>(nlf >> 1, kn, nb, ns);
                        const double *src = P.u_in + (int64_t)c.x * NPTS + nb + ((nlf & 1) ? (NP - 1) * ns : 0);
#pragma unroll
                        for (int v = 0; v < NV; v++) cp_async8(d2 + v * NFP, src + ndof * v);
                    } else {
                        const double *src = P.tr_hi + ((int64_t)c.x * NFACES + nlf) * (NV * NFP) + kn;
#pragma unroll
                        for (int v = 0; v < NV; v++) cp_async8(d2 + v * NFP, src + v * NFP);
                    }
                }
            }
            // the conn records were written with plain stores: order them before the arrive
            __threadfence_block();
            cp_async_mbar_arrive(full + s);
        }
        return;
    }

    // ======================= consumer warps =======================
    auto cons_sync = [&]() { asm volatile("bar.sync 1, %0;" ::"n"(W::NCONS) : "memory"); };
    const int el = tid / NPTS, node = tid - el * NPTS;
    const bool node_thread = el < EPB;
    double *scr = scratch + (size_t)(node_thread ? el : 0) * W::SCR_ELEM;
    double *sA = scr + W::SA;
    double *sFt = scr + W::SFT;
    double *sM = scr + W::SM;
    double *sX = scr + W::SX;
    double *sAcc = scr + W::SACC;

    for (int g = gbeg, it = 0; g < gend; g++, it++) {
        const int s = it % NSTAGE;
        double *stage = ring + (size_t)s * W::STG;
        const int nact = min(EPB, P.elem_count - g * EPB);
        const bool active = node_thread && el < nact;
        const int e = active ? elem_of(g * EPB + el) : 0;
        const int64_t dof = (int64_t)e * NPTS + node;
        double *mine = stage + (size_t)(node_thread ? el : 0) * W::STG_ELEM;
        const double *sQ = mine + W::SQ;
        const double *sT = mine + W::ST;
        const double *sF_mine = mine + W::SF;
        mbar_wait(full + s, (it / NSTAGE) & 1);
        {
        // ---------------- phase 1: node primitives, contravariant fluxes
        double Q[NV];
        double met[CART ? 1 : ND * ND];
        double vi[ND], hvi[ND], pi = 0.0, bi = 0.0, qi = 0.0;
        if (active) {
#pragma unroll
            for (int v = 0; v < NV; v++) Q[v] = sQ[v * NPTS + node];
            if (!CART) {
#pragma unroll
                for (int m = 0; m < ND * ND; m++) {
                    met[m] = __ldg(P.metric + dof + ndof * m);
                    if (SPLIT) sM[m * NPTS + node] = met[m];
                }
            }
            if (EQ == EQ_EULER) {
                NodeAux<ND> A;
                node_aux<ND>(Q, P.fp.gamma, A);
                if (!(Q[0] > 0.0) || !(A.p > 0.0)) atomicOr(P.status, 1);
#pragma unroll
                for (int d = 0; d < ND; d++) vi[d] = A.vel[d];
                pi = A.p; bi = A.beta;
                if (VOL == VOL_SPLIT_CHA) {
                    double q = 0.0;
#pragma unroll
                    for (int d = 0; d < ND; d++) {
                        q = fma(A.vel[d], A.vel[d], q);
                        hvi[d] = 0.5 * A.vel[d];
                        sA[d * NPTS + node] = hvi[d];
                    }
                    qi = q;
                    sA[ND * NPTS + node] = q;
                    sA[(ND + 1) * NPTS + node] = A.beta;
                } else if (VOL == VOL_SPLIT_STD) {
#pragma unroll
                    for (int d = 0; d < ND; d++) sA[d * NPTS + node] = A.vel[d];
                    sA[ND * NPTS + node] = A.p;
                } else {
#pragma unroll
                    for (int d = 0; d < ND; d++) {
                        double Fc[NV], Ft[NV];
#pragma unroll
                        for (int v = 0; v < NV; v++) Ft[v] = 0.0;
#pragma unroll
                        for (int c = 0; c < ND; c++) {
                            if (CART && c != d) continue;
                            const double m = CART ? P.cmet[d] : met[c + ND * d];
                            euler_flux_dir<ND>(Q, A.vel, A.p, c, Fc);
#pragma unroll
                            for (int v = 0; v < NV; v++) Ft[v] += Fc[v] * m;
                        }
#pragma unroll
                        for (int v = 0; v < NV; v++) sFt[(d * NV + v) * NPTS + node] = Ft[v];
                    }
                }
            } else if (!SPLIT) {
#pragma unroll
                for (int d = 0; d < ND; d++) {
                    double an = 0.0;
#pragma unroll
                    for (int c = 0; c < ND; c++)
                        an += P.fp.a[c] * (CART ? (c == d ? P.cmet[d] : 0.0) : met[c + ND * d]);
                    sFt[d * NPTS + node] = an * Q[0];
                }
            }
        }
        cons_sync();     // every node's state and primitives are now visible

        // ---------------- phase 2: volume term
        double acc[NV];
#pragma unroll
        for (int v = 0; v < NV; v++) acc[v] = 0.0;
        if (!SPLIT) {
            if (active) {
#pragma unroll
                for (int d = 0; d < ND; d++) {
                    int k, ii, base, stride;
                    node_line<ND, NP>(node, d, k, ii);
                    line_of<ND, NP>(d, k, base, stride);
#pragma unroll
                    for (int jj = 0; jj < NP; jj++) {
                        const double dij = sD[ii + NP * jj];
                        const int l = base + jj * stride;
#pragma unroll
                        for (int v = 0; v < NV; v++) acc[v] -= dij * sFt[(d * NV + v) * NPTS + l];
                    }
                }
            }
        }
#ifdef FLOU_EXPERIMENT_SKIP_VOLUME
        else if (P.elem_count < 0) {
#else
        else {
#endif
            // split form  dQ_i -= sum_j D#[i,j] F#(i,j)   (OpDivergence.jl:248-282).  F# is
            // symmetric: node i evaluates the pairs (i, i+s mod NP), s = 1..NP/2, keeps them
            // for itself and leaves those with s <= (NP-1)/2 in shared memory for node i+s.
#pragma unroll
            for (int d = 0; d < ND; d++) {
                int k, ii, base, stride;
                node_line<ND, NP>(node, d, k, ii);
                line_of<ND, NP>(d, k, base, stride);
                double *xb = sX + (size_t)(d & 1) * (C::XROUNDS * NV * NPTS);
                if (active) {
                    double ni[ND];
#pragma unroll
                    for (int c = 0; c < ND; c++) ni[c] = CART ? (c == d ? P.cmet[d] : 0.0) : met[c + ND * d];
                    // diagonal entry: the node's own contravariant flux (OpDivergence.jl:252)
                    {
                        double F[NV];
                        if (EQ == EQ_EULER) {
#pragma unroll
                            for (int v = 0; v < NV; v++) F[v] = 0.0;
#pragma unroll
                            for (int c = 0; c < ND; c++) {
                                if (CART && c != d) continue;
                                double Fc[NV];
                                euler_flux_dir<ND>(Q, vi, pi, c, Fc);
#pragma unroll
                                for (int v = 0; v < NV; v++) F[v] += Fc[v] * ni[c];
                            }
                        } else {
                            double an = 0.0;
#pragma unroll
                            for (int c = 0; c < ND; c++) an += P.fp.a[c] * ni[c];
                            F[0] = an * Q[0];
                        }
                        const double dii = sD[ii + NP * ii];
#pragma unroll
                        for (int v = 0; v < NV; v++) acc[v] -= dii * F[v];
                    }
#pragma unroll
                    for (int s = 1; s <= C::ROUNDS; s++) {
                        int jj = ii + s;
                        if (jj >= NP) jj -= NP;
                        const int l = base + jj * stride;
                        const double dij = sD[ii + NP * jj];
                        double n[ND], F[NV];
#pragma unroll
                        for (int c = 0; c < ND; c++)
                            n[c] = CART ? ni[c] : 0.5 * (ni[c] + sM[(c + ND * d) * NPTS + l]);
                        if (EQ == EQ_EULER) {
                            double vl[ND];
#pragma unroll
                            for (int c = 0; c < ND; c++) vl[c] = sA[c * NPTS + l];
                            if (VOL == VOL_SPLIT_CHA) {
                                tp_chandrasekhar<ND>(Q[0], hvi, qi, bi, sQ[l], vl, sA[ND * NPTS + l],
                                                     sA[(ND + 1) * NPTS + l], P.fp.inv_gm1, n, F);
                            } else {
                                double Ql[NV];
#pragma unroll
                                for (int v = 0; v < NV; v++) Ql[v] = sQ[v * NPTS + l];
                                tp_stdavg<ND>(Q, vi, pi, Ql, vl, sA[ND * NPTS + l], n, F);
                            }
                        } else {
                            double an = 0.0;
#pragma unroll
                            for (int c = 0; c < ND; c++) an += P.fp.a[c] * n[c];
                            F[0] = an * (Q[0] + sQ[l]) * 0.5;
                        }
#pragma unroll
                        for (int v = 0; v < NV; v++) acc[v] -= dij * F[v];
                        if (s <= C::XROUNDS) {
#pragma unroll
                            for (int v = 0; v < NV; v++) xb[((s - 1) * NV + v) * NPTS + node] = F[v];
                        }
                    }
                }
                if (C::XROUNDS > 0) {
                    cons_sync();
                    if (active) {
#pragma unroll
                        for (int s = 1; s <= C::XROUNDS; s++) {
                            int jp = ii - s;
                            if (jp < 0) jp += NP;
                            const int lp = base + jp * stride;
                            const double dij = sD[ii + NP * jp];
#pragma unroll
                            for (int v = 0; v < NV; v++) acc[v] -= dij * xb[((s - 1) * NV + v) * NPTS + lp];
                        }
                    }
                }
            }
        }
        // park the volume accumulators: the face phase needs the registers
        if (active) {
#pragma unroll
            for (int v = 0; v < NV; v++) sAcc[v * NPTS + node] = acc[v];
        }

        // ---------------- phase 3: face tasks (traces, BCs, Riemann flux) -> shared memory
#ifdef FLOU_EXPERIMENT_SKIP_FACES
        if (P.elem_count < 0)
#endif
#pragma unroll 1
        for (int j = 0; j < TPT; j++) {
            const int task = tid + j * C::THREADS;
            if (task >= nact * NFT) break;
            const int tel = task / NFT, r = task - tel * NFT;
            const int lf = r / NFP, k = r - lf * NFP;
            const int d = lf >> 1, side = lf & 1;
            const double *tQ = stage + (size_t)tel * W::STG_ELEM + W::SQ;
            double *tF = stage + (size_t)tel * W::STG_ELEM + W::SF;
            const Conn cn = reinterpret_cast<const Conn *>(stage + (size_t)tel * W::STG_ELEM + W::SC)[lf];

            // own trace
            double Qown[NV];
            int base, stride;
            line_of<ND, NP>(d, k, base, stride);
            if (P.colloc) {
                const int n0 = base + (side ? (NP - 1) * stride : 0);
#pragma unroll
                for (int v = 0; v < NV; v++) Qown[v] = tQ[v * NPTS + n0];
            } else {
                const double *lv = side ? sLp : sLm;
#pragma unroll
                for (int v = 0; v < NV; v++) {
                    double s = 0.0;
#pragma unroll
                    for (int ii = 0; ii < NP; ii++) s += lv[ii] * tQ[v * NPTS + base + ii * stride];
                    Qown[v] = s;
                }
            }

            const int kind = (cn.info >> 7) & 3;
            const int nlf = cn.info & 7, orient = (cn.info >> 3) & 7;
            const bool master = (cn.info >> 6) & 1;
            // face dof seen from the other side / from the master
            const int kn = master ? master2slave<ND, NP>(k, orient) : slave2master<ND, NP>(k, orient);
            const int im = master ? k : kn;          // master face dof: frame and jac index

            double fr[CART ? 1 : 3 * ND], fj;
            int dm = 0;
            double sn_ = 1.0;                      // Cartesian: normal = sn_ * e_dm
            if (CART) {
                const int pm = master ? lf : nlf;
                dm = pm >> 1;
                sn_ = (pm & 1) ? 1.0 : -1.0;
                fj = dm == 0 ? P.cfjac[0] : (dm == 1 ? P.cfjac[1] : P.cfjac[2]);
            } else {
                const int te = elem_of(g * EPB + tel);
                const int64_t fi = (int64_t)P.faceid[(int64_t)te * NFACES + lf] * NFP + im;
#pragma unroll
                for (int c = 0; c < 3 * ND; c++)
                    fr[c] = (c < ND * ND || ND == 3) ? __ldg(P.frames + fi + P.nfacedofs * c) : 0.0;
                fj = __ldg(P.fjac + fi);
            }

            double Qnb[NV];
            if (kind != FK_BOUNDARY) {
#pragma unroll
                for (int v = 0; v < NV; v++) Qnb[v] = tF[(lf * NV + v) * NFP + k];
            } else {
                // boundary face: exterior state from the BC functor (Interfaces.jl:44-48)
                const int ib = cn.info >> 9;
                const int bk = P.bc_kind[ib];
                if (bk == FLOU_B200_BC_INFLOW) {
#pragma unroll
                    for (int v = 0; v < NV; v++) Qnb[v] = P.bc_state[ib * NV + v];
                } else if (bk == FLOU_B200_BC_OUTFLOW) {
#pragma unroll
                    for (int v = 0; v < NV; v++) Qnb[v] = Qown[v];
                } else if (bk == FLOU_B200_BC_SLIP) {
                    double R[NV];
                    if (CART) rotate2face_cart<ND, EQ>(Qown, dm, sn_, R);
                    else rotate2face<ND, EQ>(Qown, fr, R);
                    if (NV > 1) R[NV > 1 ? 1 : 0] = -R[NV > 1 ? 1 : 0];
                    if (CART) rotate2phys_cart<ND, EQ>(R, dm, sn_, Qnb);
                    else rotate2phys<ND, EQ>(R, fr, Qnb);
                } else {
#pragma unroll
                    for (int v = 0; v < NV; v++) Qnb[v] = P.bc_table[((int64_t)cn.nbr * NFP + k) * NV + v];
                }
            }

            // Riemann flux with (left = master, right = slave) exactly like the reference
            double Ql[NV], Qr[NV], Fn[NV], Fp[NV];
            {
                double Ro[NV], Rn[NV];
                if (CART) {
                    rotate2face_cart<ND, EQ>(Qown, dm, sn_, Ro);
                    rotate2face_cart<ND, EQ>(Qnb, dm, sn_, Rn);
                } else {
                    rotate2face<ND, EQ>(Qown, fr, Ro);
                    rotate2face<ND, EQ>(Qnb, fr, Rn);
                }
#pragma unroll
                for (int v = 0; v < NV; v++) { Ql[v] = master ? Ro[v] : Rn[v]; Qr[v] = master ? Rn[v] : Ro[v]; }
            }
            if (EQ == EQ_EULER) {
                euler_numflux<ND>(P.fp, Ql, Qr, Fn);
            } else {
                double an = 0.0;
                if (CART) an = sn_ * pick<ND>(P.fp.a, dm);
                else {
#pragma unroll
                    for (int c = 0; c < ND; c++) an += P.fp.a[c] * fr[c];
                }
                Fn[0] = an * (Ql[0] + Qr[0]) * 0.5;
                if (P.fp.numflux == FX_LXF) Fn[0] += fabs(an) * (Ql[0] - Qr[0]) * 0.5 * P.fp.intensity;
            }
            if (CART) rotate2phys_cart<ND, EQ>(Fn, dm, sn_, Fp);
            else rotate2phys<ND, EQ>(Fn, fr, Fp);
            const double sgn = master ? fj : -fj;
#pragma unroll
            for (int v = 0; v < NV; v++) tF[(lf * NV + v) * NFP + k] = Fp[v] * sgn;
        }
        cons_sync();

        // ---------------- phase 4: lift, mass matrix, RK stage update
        if (active) {
            const double *sF = sF_mine;
#pragma unroll
            for (int v = 0; v < NV; v++) acc[v] = sAcc[v * NPTS + node];
#ifdef FLOU_EXPERIMENT_SKIP_LIFT
            if (P.elem_count < 0)
#endif
#pragma unroll
            for (int d = 0; d < ND; d++) {
                int k, ii;
                node_line<ND, NP>(node, d, k, ii);
                const double gl = sGl[ii], gr = sGr[ii];
                // collocated nodes: the lifting weights vanish away from the two end nodes
                if (gl != 0.0) {
#pragma unroll
                    for (int v = 0; v < NV; v++) acc[v] -= gl * sF[((2 * d) * NV + v) * NFP + k];
                }
                if (gr != 0.0) {
#pragma unroll
                    for (int v = 0; v < NV; v++) acc[v] -= gr * sF[((2 * d + 1) * NV + v) * NFP + k];
                }
            }
            // mass matrix: dQ / jac (Diagonal ldiv!, MultielementDiscontinuous.jl:132-137)
            const double rjac = CART ? P.crjac : fast_rcp(__ldg(P.jac + dof));
            if (P.mode == MODE_RHS) {
#pragma unroll
                for (int v = 0; v < NV; v++) P.k_out[dof + ndof * v] = acc[v] * rjac;
            } else {
#pragma unroll
                for (int v = 0; v < NV; v++) {
                    const double kv = acc[v] * rjac;
                    double t;
                    if (P.mode == MODE_STAGE_FIRST) t = P.dt * kv;
                    else t = fma(P.dt, kv, P.rkA * sT[v * NPTS + node]);
                    P.tmp[dof + ndof * v] = t;
                    const double un = fma(P.rkB, t, sQ[v * NPTS + node]);
                    P.u_out[dof + ndof * v] = un;
                    acc[v] = un;
                }
                // traces of the new state for the next stage (collocated nodes: the boundary
                // node values; Gauss nodes are handled by emit_traces_kernel)
#ifdef FLOU_EXPERIMENT_SKIP_TRACE_WRITE
                if (P.colloc && P.elem_count < 0) {
#else
                if (P.colloc) {
#endif
                    {
                        int k, ii;
                        node_line<ND, NP>(node, 0, k, ii);
                        if (ii == 0 || ii == NP - 1) {
                            double *dst = P.tr_out + ((int64_t)e * 2 + (ii == 0 ? 0 : 1)) * (NV * NFP) + k;
#pragma unroll
                            for (int v = 0; v < NV; v++) dst[v * NFP] = acc[v];
                        }
                    }
                }
            }
        }

        }
        // every consumer is done with this ring slot: hand it back to the producer
        cons_sync();
        if (tid == 0) mbar_arrive(empty + s);
    }
}

}  // namespace flou
