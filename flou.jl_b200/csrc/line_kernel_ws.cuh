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
//     line task                  pair fluxes ... wait fullF (flux blocks of the group's faces) ... lift
//                                ... wait freeP (partial sums of group i-1 consumed) ... -> sP
//     arrive fullP
//   update warp (tid >= TL), group i:
//     wait fullP                 partial sums of group i complete (and sFn[i&1] no longer read)
//     TMA                        flux blocks of the faces of group i+2 -> sFn[i&1], one bulk copy per
//                                face slot (a lane each), complete_tx on fullF[i&1]; a full iteration
//                                of lead (requested one group ahead they arrived late: 13 % of the
//                                stall samples of the line threads were this wait)
//     phase 3                    sP, sT, sU[i%3] -> tmp, u_out, traces
//     arrive freeP
//     TMA / cp.async             tmp of group i+1 -> sT, state of group i+3 -> sU[i%3], completion
//                                on fullT / fullU[i%3]
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
    constexpr int E = C::E, TL = C::TL, N = C::N, NAUX = C::NAUX;
    static_assert(C::WS, "line_kernel_ws needs the WS shared-memory layout");

    extern __shared__ __align__(16) double lsmem[];
    double *const smem = lsmem;
    double *sU = smem + C::OFF_U, *sT = smem + C::OFF_T, *sA = smem + C::OFF_A;
    double *sP = smem + C::OFF_P, *sFn = smem + C::OFF_F;
    const unsigned bar0 = (unsigned)__cvta_generic_to_shared(smem + C::OFF_BAR);
    const unsigned fullP = bar0 + 24, freeP = bar0 + 32;          // fullU[b] = bar0 + 8 b
    const unsigned fullF = bar0 + 48;                             // fullF[b] = bar0 + 48 + 8 b
    constexpr int FNB = C::FNB, FSET = E * NFACES * FNB;

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
        mbar_init(fullF, 1);
        mbar_init(fullF + 8, 1);
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
        // Flux blocks of the faces of a group: the connectivity records {slot, side} of its
        // (element, local face) pairs are read one group ahead into registers (a lane each), the
        // block of every face slot then travels as one 16-byte aligned TMA bulk copy.
        constexpr int RN = (E * NFACES + 31) / 32;
        int slot_next[RN];
        auto load_slots = [&](int gg) {
            const int nrec = min(E, P.elem_count - gg * E) * NFACES;
#pragma unroll
            for (int q = 0; q < RN; q++) {
                const int r = lane + 32 * q;
                slot_next[q] = r < nrec ? __ldg(&P.econn[(int64_t)(P.elem_first + gg * E) * NFACES + r].x) : 0;
            }
        };
        auto issue_fn = [&](int gg, int buf) {
            const int nrec = min(E, P.elem_count - gg * E) * NFACES;
            const unsigned bar = fullF + 8 * buf;
            if (lane == 0) mbar_expect_tx(bar, (unsigned)(nrec * FNB * sizeof(double)));
            __syncwarp();
#pragma unroll
            for (int q = 0; q < RN; q++) {
                const int r = lane + 32 * q;
                if (r < nrec)
                    bulk_g2s_keep(sFn + buf * FSET + r * FNB, P.Fn + (int64_t)slot_next[q] * FNB, (unsigned)(FNB * sizeof(double)), bar);
            }
        };
        for (int i = 0; i < 3 && i < niter; i++) issue_planes(P.u_in, sU + i * (NV * N), g0 + i * gs, bar0 + 8 * i);
        if (need_tmp) issue_planes(P.tmp, sT, g0, fullT);
        cp_async_commit();
        load_slots(g0);
        issue_fn(g0, 0);
        if (niter > 1) { load_slots(g0 + gs); issue_fn(g0 + gs, 1); }
        if (niter > 2) load_slots(g0 + 2 * gs);

        for (int i = 0; i < niter; i++) {
            const int g = g0 + i * gs, ub = i % 3;
            const int nact = min(E, P.elem_count - g * E), nn = nact * NPTS;
            const int64_t dof0 = (int64_t)(P.elem_first + g * E) * NPTS;
            double *U = sU + ub * (NV * N);
            mbar_wait(fullP, i & 1);
            // every line thread is done with the flux blocks of group i: their buffer takes those of
            // group i+2 (in flight during a whole iteration of the line threads)
            if (i + 2 < niter) {
                issue_fn(g + 2 * gs, i & 1);
                if (i + 3 < niter) load_slots(g + 3 * gs);
            }
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

        // ---------------- phase 2: one tensor-product line per thread
        const unsigned fb = i > 0 ? freeP : 0u, fpar = (unsigned)((i - 1) & 1);
        const FaceSrc fs{sFn + (i & 1) * FSET, sEC + (i & 1) * (E * NFACES), fullF + 8 * (i & 1), (unsigned)((i >> 1) & 1)};
        for (int task = tid; task < nl; task += TL) {
            if (EQ == EQ_EULER && VOL == VOL_SPLIT_CHA && !C::NB) {
                if (line_task<C, true>(P, A, sP, fs, task, dof0, fb, fpar)) line_task_exact<C>(P, A, sP, fs, task, dof0, fb, fpar);
            } else {
                line_task<C, false>(P, A, sP, fs, task, dof0, fb, fpar);
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
