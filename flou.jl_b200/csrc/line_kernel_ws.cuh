// Warp-specialised variant of the line-per-thread element kernel (line_kernel.cuh): the same three
// phases per group of E elements, but phase 3 (sum of the directional partial sums, mass matrix,
// RK update, traces) and every global->shared copy of the state run on ONE dedicated "update
// warp", while the TL line threads go from the line phase of one group straight into phase 1 of
// the next.  The pair-flux arithmetic of a CTA then never waits for its own memory phases (in the
// single-role kernel phases 1/3, their barriers and the copy issue were 55 % of the stall samples).
//
//   line threads (tid < TL), group i:
//     wait fullU[i%3]            state of group i in shared memory
//     phase 1                    node data -> sA[i&1]
//     bar.sync 1, TL             node data visible to the line threads
//     cp.async                   own face fluxes -> sF (own column)
//     line task                  ... wait freeP (partial sums of group i-1 consumed) ... -> sP
//     arrive fullP
//   update warp (tid >= TL), group i:
//     wait fullP                 partial sums of group i complete
//     phase 3                    sP, sT, sU[i%3] -> tmp, u_out, traces
//     arrive freeP
//     cp.async                   tmp of group i+1 -> sT, state of group i+3 -> sU[i%3], arrive-on-
//                                completion at fullU[i%3]
// Every mbarrier is one phase ahead of its waiters at most (see the ordering argument in
// profiles/r1_kernel_notes.md), so single-bit parities are safe.
#pragma once
#include "line_kernel.cuh"

namespace flou {

template <class C>
__global__ void __launch_bounds__(C::T, C::MINB)
line_kernel_ws(const __grid_constant__ KParams P)
{
    constexpr int ND = C::ND, NP = C::NP, EQ = C::EQ, VOL = C::VOL, NV = C::NV;
    constexpr int NPTS = C::NPTS, NFP = C::NFP, NFACES = C::NFACES, NLINES = C::NLINES;
    constexpr int E = C::E, TL = C::TL, N = C::N, LT = C::LT, NAUX = C::NAUX;
    constexpr bool FOLD = C::FOLD;
    static_assert(C::WS, "line_kernel_ws needs the WS shared-memory layout");

    extern __shared__ __align__(16) double lsmem[];
    double *const smem = lsmem;
    double *sU = smem + C::OFF_U, *sT = smem + C::OFF_T, *sA = smem + C::OFF_A;
    double *sP = smem + C::OFF_P, *sF = smem + C::OFF_F;
    const unsigned bar0 = (unsigned)__cvta_generic_to_shared(smem + C::OFF_BAR);
    const unsigned fullP = bar0 + 24, freeP = bar0 + 32;          // fullU[b] = bar0 + 8 b

    const int ngroups = (P.elem_count + E - 1) / E;
    const int64_t ndof = P.ndof;
    const bool need_tmp = (P.mode == MODE_STAGE);
    const int g0 = blockIdx.x, gs = gridDim.x;
    if (g0 >= ngroups) return;
    const int niter = (ngroups - g0 + gs - 1) / gs;

    // 16-byte aligned planes and an even node count per group: the update warp moves the state with
    // TMA bulk copies issued by one lane (and works on node pairs); otherwise 8-byte cp.async
    const bool wide = ((ndof & 1) == 0) && (((int64_t)P.elem_first * NPTS & 1) == 0) && ((N & 1) == 0) &&
                      ((reinterpret_cast<uintptr_t>(P.u_in) & 15) == 0) && ((reinterpret_cast<uintptr_t>(P.tmp) & 15) == 0) &&
                      ((reinterpret_cast<uintptr_t>(P.u_out) & 15) == 0) && ((reinterpret_cast<uintptr_t>(P.k_out) & 15) == 0) &&
                      (C::CART || (reinterpret_cast<uintptr_t>(P.jac) & 15) == 0);
    const unsigned fullT = bar0 + 40;
    if (threadIdx.x == 0) {
        for (int b = 0; b < 3; b++) mbar_init(bar0 + 8 * b, wide ? 1 : 32);
        mbar_init(fullP, TL);
        mbar_init(freeP, 32);
        mbar_init(fullT, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (threadIdx.x >= TL) {
        // =========================================================== update warp
        const int lane = threadIdx.x - TL;
        // NV planes of the nodes of group gg -> shared memory, completion on `bar`
        auto issue_planes = [&](const double *src, double *dst, int gg, unsigned bar) {
            const int nn = min(E, P.elem_count - gg * E) * NPTS;
            const double *s0 = src + (int64_t)(P.elem_first + gg * E) * NPTS;
            if (wide) {
                if (lane == 0) {
                    mbar_expect_tx(bar, (unsigned)(NV * nn * sizeof(double)));
#pragma unroll
                    for (int v = 0; v < NV; v++) bulk_g2s(dst + v * N, s0 + ndof * v, (unsigned)(nn * sizeof(double)), bar);
                }
            } else {
                for (int n = lane; n < nn; n += 32) {
#pragma unroll
                    for (int v = 0; v < NV; v++) cp_async8(dst + v * N + n, s0 + n + ndof * v);
                }
                if (bar != fullT) mbar_arrive_cp_async(bar);       // tmp: this warp's own wait_group
            }
        };
        for (int i = 0; i < 3 && i < niter; i++) issue_planes(P.u_in, sU + i * (NV * N), g0 + i * gs, bar0 + 8 * i);
        if (need_tmp) issue_planes(P.tmp, sT, g0, fullT);
        cp_async_commit();

        for (int i = 0; i < niter; i++) {
            const int g = g0 + i * gs, ub = i % 3;
            const int nact = min(E, P.elem_count - g * E), nn = nact * NPTS;
            const int64_t dof0 = (int64_t)(P.elem_first + g * E) * NPTS;
            double *U = sU + ub * (NV * N);
            mbar_wait(fullP, i & 1);
            if (wide) {
                mbar_wait(bar0 + 8 * ub, (i / 3) & 1);      // completed long ago: makes the TMA writes visible here
                if (need_tmp) mbar_wait(fullT, i & 1);
            } else {
                cp_async_wait<0>();      // tmp of this group and this lane's share of the state copies
                __syncwarp();
            }
            if (wide) {
                if (N >= 128) phase3_pairs<C, 2, 32, true>(P, U, sT, sP, lane, nn, dof0, g);
                else phase3_pairs<C, 1, 32, true>(P, U, sT, sP, lane, nn, dof0, g);
                __syncwarp();
                if (P.mode != MODE_RHS && P.colloc) trace_pass<C, 32>(P, U, lane, nact, g);
            } else if (N >= 64) phase3_nodes<C, 2, 32>(P, U, sT, sP, lane, nn, dof0, g);
            else phase3_nodes<C, 1, 32>(P, U, sT, sP, lane, nn, dof0, g);
            __syncwarp();                // every lane is done with sP, sT and sU[ub]
            mbar_arrive(freeP);
            if (need_tmp && i + 1 < niter) issue_planes(P.tmp, sT, g + gs, fullT);
            if (i + 3 < niter) issue_planes(P.u_in, sU + ub * (NV * N), g + 3 * gs, bar0 + 8 * ub);
            cp_async_commit();
        }
        return;
    }

    // =============================================================== line threads
    // face connectivity of the elements of a group, staged one group ahead (see line_kernel.cuh)
    int2 *sEC = reinterpret_cast<int2 *>(smem + C::OFF_EC);
    auto issue_ec = [&](int gg, int buf, int t) {
        const int nrec = min(E, P.elem_count - gg * E) * NFACES;
        if (t < nrec)
            cp_async8(reinterpret_cast<double *>(sEC + buf * (E * NFACES) + t),
                      reinterpret_cast<const double *>(P.econn + (int64_t)(P.elem_first + gg * E) * NFACES + t));
    };
    // face fluxes of a line: column `task` of sF belongs to the thread that owns the line
    auto issue_fn = [&](int task, const int2 *ec) {
        const int el = task / NLINES, r_ = task - el * NLINES;
        const int d = r_ / NFP, k = r_ - d * NFP;
        const int2 ecL = ec[el * NFACES + 2 * d], ecR = ec[el * NFACES + 2 * d + 1];
        const int iL = (ecL.y & 1) ? k : slave2master<ND, NP>(k, (ecL.y >> 1) & 7);
        const int iR = (ecR.y & 1) ? k : slave2master<ND, NP>(k, (ecR.y >> 1) & 7);
        const double *sL = P.Fn + (int64_t)ecL.x * (NV * NFP) + iL;
        const double *sR = P.Fn + (int64_t)ecR.x * (NV * NFP) + iR;
#pragma unroll
        for (int v = 0; v < NV; v++) {
            int vv = v;     // FOLD mode handles the momentum components in cyclic order starting at d
            if (FOLD && EQ == EQ_EULER && v >= 1 && v <= ND) { const int s_ = d + v - 1; vv = 1 + (s_ >= ND ? s_ - ND : s_); }
            cp_async8(sF + v * LT + task, sL + vv * NFP);
            cp_async8(sF + (NV + v) * LT + task, sR + vv * NFP);
        }
        sF[(2 * NV) * LT + task] = (ecL.y & 1) ? 1.0 : -1.0;
        sF[(2 * NV + 1) * LT + task] = (ecR.y & 1) ? 1.0 : -1.0;
    };

    int tid = threadIdx.x;
    issue_ec(g0, 0, tid);
    cp_async_commit();

    for (int i = 0; i < niter; i++) {
        // thread index re-read every iteration (see line_kernel.cuh: keeps loop invariants out of
        // the registers of the line phase)
        asm volatile("mov.u32 %0, %%tid.x;" : "=r"(tid));
        const int g = g0 + i * gs, ub = i % 3;
        const int nact = min(E, P.elem_count - g * E);
        const int nn = nact * NPTS, nl = nact * NLINES;
        const int64_t dof0 = (int64_t)(P.elem_first + g * E) * NPTS;
        const double *U = sU + ub * (NV * N);
        double *A = sA + (i & 1) * (NAUX * N);

        mbar_wait(bar0 + 8 * ub, (i / 3) & 1);
        // ---------------- phase 1
        if (N > TL && (tid & ~31) + TL < nn) phase1_nodes<C, 2, TL>(P, U, A, tid, nn, dof0);
        else phase1_nodes<C, 1, TL>(P, U, A, tid, nn, dof0);
        cp_async_wait<0>();          // connectivity records of this group (issued one iteration ago)
        asm volatile("bar.sync 1, %0;" ::"n"(TL) : "memory");

        // ---------------- face fluxes of this group's lines
        for (int task = tid; task < nl; task += TL) issue_fn(task, sEC + (i & 1) * (E * NFACES));
        cp_async_commit();

        // ---------------- phase 2: one tensor-product line per thread
        const unsigned fb = i > 0 ? freeP : 0u, fpar = (unsigned)((i - 1) & 1);
        for (int task = tid; task < nl; task += TL) {
            if (EQ == EQ_EULER && VOL == VOL_SPLIT_CHA && !C::NB) {
                if (line_task<C, true>(P, A, sP, sF, task, dof0, fb, fpar)) line_task_exact<C>(P, A, sP, sF, task, dof0, fb, fpar);
            } else {
                line_task<C, false>(P, A, sP, sF, task, dof0, fb, fpar);
            }
        }
        mbar_arrive(fullP);
        // connectivity of the next group: in flight during the wait for its state and phase 1
        asm volatile("mov.u32 %0, %%tid.x;" : "=r"(tid));
        if (i + 1 < niter) issue_ec(g0 + (i + 1) * gs, (i + 1) & 1, tid);
        cp_async_commit();
    }
}

}  // namespace flou
