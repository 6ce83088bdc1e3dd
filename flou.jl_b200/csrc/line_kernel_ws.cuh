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

// grid-wide loop bounds re-read from the special registers where they are needed: values computed
// once in a common prologue were spilled to local memory by ptxas (the line phase needs every
// register) and reloaded by BOTH roles on their critical paths
__device__ __forceinline__ int ws_cta() { int v; asm volatile("mov.u32 %0, %%ctaid.x;" : "=r"(v)); return v; }
__device__ __forceinline__ int ws_nctas() { int v; asm volatile("mov.u32 %0, %%nctaid.x;" : "=r"(v)); return v; }

// FLOU_LINE_NO_ELECT: copies issued under `lane == 0` / a slot per lane (round-2 code before r2m)
#ifndef FLOU_LINE_NO_ELECT
#define FLOU_LINE_ELECT 1
#define FLOU_ELECT (tu < 32 && elect_one())
#else
#define FLOU_ELECT (lane == 0)
#endif
template <class C>
__device__ __forceinline__ void ws_body(const KParams &P)
{
    constexpr int EQ = C::EQ, VOL = C::VOL, NV = C::NV;
    constexpr int NPTS = C::NPTS, NFACES = C::NFACES, NLINES = C::NLINES;
    constexpr int E = C::E, TL = C::TL, N = C::N, NAUX = C::NAUX;
    static_assert(C::WS, "line_kernel_ws needs the WS shared-memory layout");
    using B = WsBars<C>;

    extern __shared__ __align__(16) double lsmem[];
    double *const smem = lsmem;
    constexpr int FNB = C::FNB, FSET = E * NFACES * FNB;

    pdl_prologue();
    if (ws_cta() * E >= P.elem_count) return;
    if (threadIdx.x == 0) {
        const bool wide = ws_wide<C>(P);
        // copies of a buffer: one issuing lane (TMA), or one per line warp (LINE_ISSUE), or the 32
        // lanes of the update warp (cp.async fallback)
        const unsigned nissue = wide ? (C::LINE_ISSUE ? TL / 32 : 1) : 32;
        for (int b = 0; b < 3; b++) mbar_init(B::fullU(b), nissue);
        mbar_init(B::fullP(), TL);
        mbar_init(B::freeP(), 32 * C::NUPD);
        mbar_init(B::fullT(), wide && C::LINE_ISSUE ? TL / 32 : 1);
        mbar_init(B::fullF(0), 1);
        mbar_init(B::fullF(1), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // iteration counters of the line warps (see line_task)
    volatile int *const sIt = reinterpret_cast<volatile int *>(smem + C::OFF_BAR + 8);
    if (threadIdx.x < 16) sIt[threadIdx.x] = 0;
    __syncthreads();

    if (threadIdx.x >= TL) {
        // =========================================================== update warp(s)
        // NUPD warps share phase 3 and the trace pass of a group; the first one issues every copy
        constexpr int TU = 32 * C::NUPD;
        const int tu = threadIdx.x - TL;                 // index among the update threads
        const int lane = tu < 32 ? tu : 1000;            // copy issue: first update warp only
        auto upd_sync = [&]() {
            if constexpr (C::NUPD == 1) __syncwarp();
            else asm volatile("bar.sync 2, %0;" ::"n"(TU) : "memory");
        };
        const int64_t ndof = P.ndof;
        // group gg exists
        auto live = [&](int gg) { return gg * E < P.elem_count; };
        const bool need_tmp = (P.mode == MODE_STAGE);
        const bool wide = ws_wide<C>(P);
        const unsigned long long pol = l2_evict_first();
        double *sU = smem + C::OFF_U, *sT = smem + C::OFF_T, *sP = smem + C::OFF_P, *sFn = smem + C::OFF_F;
        // NV planes of the nodes of group gg -> shared memory, completion on `bar`
        auto issue_planes = [&](const double *src, double *dst, int gg, unsigned bar, bool is_tmp) {
            if (tu >= 32) return;
            const int nn = min(E, P.elem_count - gg * E) * NPTS;
            const double *s0 = src + (int64_t)(P.elem_first + gg * E) * NPTS;
            if (wide) {
                if (FLOU_ELECT) {
                    mbar_expect_tx(bar, (unsigned)(NV * nn * sizeof(double)));
                    if (C::LINE_ISSUE) for (int k = 1; k < TL / 32; k++) mbar_arrive(bar);      // the barrier counts one arrival per line warp
#pragma unroll
                    for (int v = 0; v < NV; v++) bulk_g2s(dst + v * N, s0 + ndof * v, (unsigned)(nn * sizeof(double)), bar, pol);
                }
            } else {
                for (int n = lane; n < nn; n += 32) {
#pragma unroll
                    for (int v = 0; v < NV; v++) cp_async8(dst + v * N + n, s0 + n + ndof * v);
                }
                if (!is_tmp) mbar_arrive_cp_async(bar);       // tmp: this warp's own wait_group
            }
        };
        // Flux blocks of the faces of a group: the flux slots of its (element, local face) pairs
        // are staged in shared memory one group ahead (cp.async, a lane each: in registers they were
        // spilled across phase 3 and reloaded from local memory on this warp's critical path), the
        // block of every face slot then travels as one 16-byte aligned TMA bulk copy.
        constexpr int RN = (E * NFACES + 31) / 32;
        int *const sSlot = reinterpret_cast<int *>(smem + C::OFF_SLOT);
        auto load_slots = [&](int gg) {
            if (tu >= 32) return;
            const int nrec = min(E, P.elem_count - gg * E) * NFACES;
#pragma unroll
            for (int q = 0; q < RN; q++) {
                const int r = lane + 32 * q;
                if (r < nrec) cp_async4(sSlot + r, &P.econn[(int64_t)(P.elem_first + gg * E) * NFACES + r].x);
            }
            cp_async_commit();
        };
        auto issue_fn = [&](int gg, int buf) {
            const int nrec = min(E, P.elem_count - gg * E) * NFACES;
            const unsigned bar = B::fullF(buf);
            if (tu >= 32) return;
            cp_async_wait<0>();          // this lane's slots (requested an iteration ago)
            __syncwarp();
#ifdef FLOU_LINE_ELECT
            // one elected lane issues every copy: straight-line uniform-datapath code (a copy per lane
            // is serialised by a loop over the active lanes with an ELECT / branch pair per copy)
            if (elect_one()) {
                mbar_expect_tx(bar, (unsigned)(nrec * FNB * sizeof(double)));
                auto copy = [&](int r, int slot) {
                    bulk_g2s_keep(sFn + buf * FSET + r * FNB, P.Fn + (int64_t)slot * FNB, (unsigned)(FNB * sizeof(double)), bar);
                };
                if constexpr (E * NFACES <= 16) {
                    // slots first (the copies are volatile asm: loads are not moved across them)
                    int sl[E * NFACES];
#pragma unroll
                    for (int r = 0; r < E * NFACES; r++) sl[r] = sSlot[r];
                    if (nrec == E * NFACES) {
#pragma unroll
                        for (int r = 0; r < E * NFACES; r++) copy(r, sl[r]);
                    } else {
#pragma unroll
                        for (int r = 0; r < E * NFACES; r++) if (r < nrec) copy(r, sl[r]);
                    }
                } else {
#pragma unroll 4
                    for (int r = 0; r < nrec; r++) copy(r, sSlot[r]);
                }
            }
#else
            if (lane == 0) mbar_expect_tx(bar, (unsigned)(nrec * FNB * sizeof(double)));
#pragma unroll
            for (int q = 0; q < RN; q++) {
                const int r = lane + 32 * q;
                if (r < nrec)
                    bulk_g2s_keep(sFn + buf * FSET + r * FNB, P.Fn + (int64_t)sSlot[r] * FNB, (unsigned)(FNB * sizeof(double)), bar);
            }
#endif
            __syncwarp();
        };
        {
            const int g0 = ws_cta(), gs = ws_nctas();
            for (int i = 0; i < 3 && live(g0 + i * gs); i++) issue_planes(P.u_in, sU + i * (NV * N), g0 + i * gs, B::fullU(i), false);
            if (need_tmp) issue_planes(P.tmp, sT, g0, B::fullT(), true);
            cp_async_commit();
            load_slots(g0);
            issue_fn(g0, 0);
            if (live(g0 + gs)) { load_slots(g0 + gs); issue_fn(g0 + gs, 1); }
            if (live(g0 + 2 * gs)) load_slots(g0 + 2 * gs);
        }
        // iteration counter in shared memory, grid bounds re-read: nothing of the loop is carried
        // in registers (spilled) across phase 3
        volatile int *const itU = sIt + (threadIdx.x >> 5);
        for (;;) {
            const int i = *itU, gs = ws_nctas();
            const int g = ws_cta() + i * gs, ub = i % 3;
            if (!live(g)) break;
            const int nact = min(E, P.elem_count - g * E), nn = nact * NPTS;
            const int64_t dof0 = (int64_t)(P.elem_first + g * E) * NPTS;
            double *U = sU + ub * (NV * N);
            bool released = false;       // freeP arrived and tmp of the next group requested early (wide path)
            mbar_wait(B::fullP(), i & 1);
            // every line thread is done with the flux blocks of group i: their buffer takes those of
            // group i+2 (in flight during a whole iteration of the line threads)
            if (live(g + 2 * gs)) {
                issue_fn(g + 2 * gs, i & 1);
                if (live(g + 3 * gs)) load_slots(g + 3 * gs);
            }
            if (wide) {
                mbar_wait(B::fullU(ub), (i / 3) & 1);      // completed long ago: makes the TMA writes visible here
                if (need_tmp) mbar_wait(B::fullT(), i & 1);
            } else {
                cp_async_wait<0>();      // tmp of this group and this lane's share of the state copies
                upd_sync();
            }
            if (wide) {
                constexpr int RP3 = (N >= 128 && C::NUPD == 1) ? 2 : 1;
                // one warp-uniform branch on the pass mode, then code without mode selects
                if (P.source) phase3_pairs<C, RP3, TU, true, C::BULK_STORE>(P, U, sT, sP, tu, nn, dof0, g);
                else if (P.mode == MODE_STAGE) phase3_pairs<C, RP3, TU, true, C::BULK_STORE, MODE_STAGE, 0>(P, U, sT, sP, tu, nn, dof0, g);
                else if (P.mode == MODE_STAGE_FIRST) phase3_pairs<C, RP3, TU, true, C::BULK_STORE, MODE_STAGE_FIRST, 0>(P, U, sT, sP, tu, nn, dof0, g);
                else phase3_pairs<C, RP3, TU, true, C::BULK_STORE, MODE_RHS, 0>(P, U, sT, sP, tu, nn, dof0, g);
                if (!C::BULK_STORE) {
                    // sP and sT are free as soon as phase 3 is through: the line threads may store the
                    // next partial sums and tmp of the next group may land while the traces are written
                    upd_sync();
                    mbar_arrive(B::freeP());
                    released = true;
                    if (!C::LINE_ISSUE && need_tmp) {
                        const int gn = ws_cta() + (*itU + 1) * ws_nctas();      // re-read, not carried across phase 3
                        if (live(gn)) issue_planes(P.tmp, sT, gn, B::fullT(), true);
                    }
                    if (P.mode != MODE_RHS && P.colloc && P.tr_out) trace_pass<C, TU>(P, U, tu, nact, g);
                } else if (P.mode != MODE_RHS) {
                    // tmp (in sT) and the new state (in U) leave as TMA bulk stores, one per plane
                    fence_async_smem();
                    upd_sync();
                    if (tu == 0) {
#pragma unroll
                        for (int v = 0; v < NV; v++) {
                            bulk_s2g(P.tmp + dof0 + ndof * v, sT + v * N, (unsigned)(nn * sizeof(double)));
                            bulk_s2g(P.u_out + dof0 + ndof * v, U + v * N, (unsigned)(nn * sizeof(double)));
                        }
                        bulk_commit();
                    }
                    if (P.colloc) trace_pass<C, TU>(P, U, tu, nact, g);
                    if (tu == 0) bulk_wait_read();      // sT and U may be refilled
                }
            } else if (N >= 32 * C::NUPD) phase3_nodes<C, 2, TU>(P, U, sT, sP, tu, nn, dof0, g);
            else phase3_nodes<C, 1, TU>(P, U, sT, sP, tu, nn, dof0, g);
            upd_sync();                  // every update thread is done with sP, sT and sU[ub]
            if (!released) mbar_arrive(B::freeP());
            {
                const int i2 = *itU, gs2 = ws_nctas(), g2 = ws_cta() + i2 * gs2, ub2 = i2 % 3;
                if (!wide || !C::LINE_ISSUE) {     // LINE_ISSUE: a line thread issues the TMA copies (ws_issue_loads)
                    if (!released && need_tmp && live(g2 + gs2)) issue_planes(P.tmp, sT, g2 + gs2, B::fullT(), true);
                    if (live(g2 + 3 * gs2)) issue_planes(P.u_in, sU + ub2 * (NV * N), g2 + 3 * gs2, B::fullU(ub2), false);
                }
                cp_async_commit();
                __syncwarp();
                if ((tu & 31) == 0) *itU = i2 + 1;
                __syncwarp();
            }
        }
        return;
    }

    // =============================================================== line threads
    // face connectivity of the elements of a group, staged one group ahead
    int2 *const sEC = reinterpret_cast<int2 *>(smem + C::OFF_EC);
    auto issue_ec = [&](int gg, int buf, int t) {
        const int nrec = min(E, P.elem_count - gg * E) * NFACES;
        if (t < nrec)
            cp_async8(reinterpret_cast<double *>(sEC + buf * (E * NFACES) + t),
                      reinterpret_cast<const double *>(P.econn + (int64_t)(P.elem_first + gg * E) * NFACES + t));
    };
    issue_ec(ws_cta(), 0, threadIdx.x);
    cp_async_commit();
    const volatile int *const it = sIt + (threadIdx.x >> 5);

    for (;;) {
        // thread index, grid bounds and the iteration counter are re-read every iteration instead of
        // being carried across the line phase (see line_task)
        int tid;
        asm volatile("mov.u32 %0, %%tid.x;" : "=r"(tid));
        const int i = *it;
        const int g = ws_cta() + i * ws_nctas();
        if (g * E >= P.elem_count) break;
        const int ub = i % 3;
        const int nact = min(E, P.elem_count - g * E);
        const int nn = nact * NPTS, nl = nact * NLINES;
        const int64_t dof0 = (int64_t)(P.elem_first + g * E) * NPTS;
        const double *U = smem + C::OFF_U + ub * (NV * N);
        double *A = smem + C::OFF_A + (i & 1) * (NAUX * N);

        mbar_wait(B::fullU(ub), (i / 3) & 1);
        // ---------------- phase 1
        if (N > TL && (tid & ~31) + TL < nn) phase1_nodes<C, 2, TL>(P, U, A, tid, nn, dof0);
        else phase1_nodes<C, 1, TL>(P, U, A, tid, nn, dof0);
        cp_async_wait<0>();          // connectivity records of this group (issued one iteration ago)
        asm volatile("bar.sync 1, %0;" ::"n"(TL) : "memory");

        // ---------------- phase 2: one tensor-product line per thread
        auto one_task = [&](int task) {
            if (EQ == EQ_EULER && VOL == VOL_SPLIT_CHA && !C::NB) {
                if (line_task<C, true>(P, task, dof0, it)) {
                    // exact redo (rare): task and dof0 from scratch, nothing kept for it across the fast path
                    int t_;
                    asm volatile("mov.u32 %0, %%tid.x;" : "=r"(t_));
                    const int g_ = ws_cta() + *it * ws_nctas();
                    line_task_exact<C>(P, C::ONE_ROUND ? t_ : task, (int64_t)(P.elem_first + g_ * E) * NPTS, it);
                }
            } else {
                line_task<C, false>(P, task, dof0, it);
            }
        };
        if constexpr (C::ONE_ROUND) {
            if (tid < nl) one_task(tid);
        } else {
            for (int task = tid; task < nl; task += TL) one_task(task);
        }
        if constexpr (C::LINE_ISSUE) {
            // every line warp requests its planes of tmp (this group) and of the state two groups on,
            // as soon as the update warp is done with the buffers (warps without a line task in a tail
            // group wait here)
            int t3;
            asm volatile("mov.u32 %0, %%tid.x;" : "=r"(t3));
            const int i3 = *it;
            if ((t3 & 31) == 0 && i3 > 0 && ws_wide<C>(P)) {
                mbar_wait(B::freeP(), (unsigned)((i3 - 1) & 1));
                ws_issue_loads<C>(P, i3, t3 >> 5);
            }
            __syncwarp();
        }
        mbar_arrive(B::fullP());
        // connectivity of the next group: in flight during the wait for its state and phase 1
        {
            int t2;
            asm volatile("mov.u32 %0, %%tid.x;" : "=r"(t2));
            const int i2 = *it + 1, g2 = ws_cta() + i2 * ws_nctas();
            if (g2 * E < P.elem_count) issue_ec(g2, i2 & 1, t2);
            cp_async_commit();
            __syncwarp();            // every lane has read the counter of this iteration
            if ((t2 & 31) == 0) sIt[t2 >> 5] = i2;
            __syncwarp();
        }
    }
}

// The kernel: register budget either from the launch bounds (MINB CTAs of T threads per SM) or, for
// the instances with two update warps, as an explicit cap (C::MAXREG: 2 CTAs x 224 threads x 144
// registers fill the register file; the launch-bounds heuristic would stop at 128 and spill).
template <class C>
__global__ void __launch_bounds__(C::T, C::MINB) line_kernel_ws(const __grid_constant__ KParams P) { ws_body<C>(P); }
template <class C>
__global__ void __maxnreg__(C::MAXREG > 0 ? C::MAXREG : 255) line_kernel_ws_mr(const __grid_constant__ KParams P) { ws_body<C>(P); }

}  // namespace flou
