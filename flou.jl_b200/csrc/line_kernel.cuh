// Element kernel of the two-kernel RK stage, LINE-PER-THREAD formulation, persistent CTAs with an
// asynchronous-copy pipeline.
//
//   k = volume(u_in) + lift(Fn);  k /= jac;  tmp = A*tmp + dt*k;  u_out = u_in + B*tmp
//
// i.e. volume_contribution! (OpDivergence.jl:105-160 strong, :184-282 split),
// surface_contribution! (OpDivergence.jl:42-100), apply_massmatrix!
// (MultielementDiscontinuous.jl:132-137) and the LowStorageRK2N stage update (call site
// FlouTime.jl:34-38) in one pass; the Riemann fluxes Fn come from face_flux_kernel.
//
// Why lines: the reference's volume and surface operators are both loops over the tensor-product
// lines of an element (`tpdofs`, StdQuad.jl:116-124, StdHex.jl:135-145).  A thread that owns one
// line holds its NP nodes in registers, evaluates the NP(NP-1)/2 symmetric two-point fluxes of
// the split form exactly once each, applies D# (or Ds) from the constant bank and adds the two
// face fluxes at the line's ends -- no exchange buffers, no barriers inside the volume term, and
// one third of the shared-memory traffic of the node-per-thread kernel (which ncu showed to be
// bound by the shared-memory pipe at 83 % L1TEX, profiles/r1_kernel_notes.md).
//
// A persistent CTA walks over groups of E consecutive elements, g = blockIdx.x, +gridDim.x, ...:
//   phase 1  node tasks:  state of group g (already in shared memory) -> node primitives
//   phase 2  line tasks:  (element, direction, line) -> partial sums, one plane set per direction
//   phase 3  node tasks:  sum the ND partial sums, 1/jac, RK update, x-face traces of u_out
// with every global load in flight behind the flux arithmetic of phase 2.  This header holds the
// configuration (LCfg) and the three phases; the kernel itself -- warp-specialised, TMA bulk copies,
// mbarrier hand-offs -- is line_kernel_ws.cuh.
#pragma once
#include <type_traits>
#include "stage_kernel.cuh"
#include "face_kernel.cuh"     // rot2face_c / rot2phys_c: sub-grid frames of the hybrid operator

namespace flou {

// ---- mbarrier helpers (shared::cta addresses as 32-bit values)
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar)
{
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(bar) : "memory");
}
// arrive once every cp.async this thread issued before has completed (no pending-count increment:
// the arrival is part of the barrier's initial count)
__device__ __forceinline__ void mbar_arrive_cp_async(unsigned bar)
{
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
// one arrival plus `bytes` of expected asynchronous-copy traffic (TMA bulk copies complete_tx on it)
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes)
{
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(bar), "r"(bytes) : "memory");
}
// TMA 1-D bulk copy global -> shared, completion signalled on an mbarrier; 16-byte aligned
// addresses, size a multiple of 16 bytes
// The state and tmp are streamed exactly once per stage: L2 evict-first, so that the face-flux
// blocks, which both neighbours of a face read, survive in L2 until their second use.
__device__ __forceinline__ unsigned long long l2_evict_first()
{
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
// one lane of a converged warp (elect.sync: the compiler then knows that a single lane issues the
// copies that follow and emits no loop over the active lanes around every UBLKCP)
__device__ __forceinline__ bool elect_one()
{
    unsigned p;
    asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\tselp.u32 %0, 1, 0, q;\n\t}" : "=r"(p));
    return p != 0;
}
__device__ __forceinline__ void bulk_g2s(double *smem_dst, const double *gmem_src, unsigned bytes, unsigned bar,
                                         unsigned long long pol)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(d), "l"(gmem_src), "r"(bytes), "r"(bar), "l"(pol) : "memory");
}
// the same with the default L2 policy: face-flux blocks, which the second neighbour of the face
// reads again a little later
__device__ __forceinline__ void bulk_g2s_keep(double *smem_dst, const double *gmem_src, unsigned bytes, unsigned bar)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(d), "l"(gmem_src), "r"(bytes), "r"(bar) : "memory");
}
// TMA 1-D bulk copy shared -> global (bulk async-group completion): tmp and the new state are
// streamed out once per stage, L2 evict-first like the loads
__device__ __forceinline__ void bulk_s2g(double *gmem_dst, const double *smem_src, unsigned bytes)
{
    const unsigned s_ = (unsigned)__cvta_generic_to_shared(smem_src);
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;"
                 ::"l"(gmem_dst), "r"(s_), "r"(bytes), "l"(pol) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// the shared-memory sources of every committed bulk store have been read (they may be overwritten)
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// generic-proxy writes to shared memory made visible to the async proxy (TMA)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// Wait for the phase with the given parity.  A failed try_wait comes back within a few cycles on
// sm_100a, so a plain retry loop spins at ~4 cycles per iteration and takes issue slots from the
// compute warps of the same scheduler (ncu: 157 retries per wait of the line threads on freeP).
// FLOU_MBAR_WAIT selects the back-off: 0 = plain retry, 1 = try_wait with a suspend-time hint
// (measured: no effect), 2 = nanosleep between tries (default; element kernel -1.3 %).
#ifndef FLOU_MBAR_WAIT
#define FLOU_MBAR_WAIT 2
#endif
#ifndef FLOU_MBAR_NS
#define FLOU_MBAR_NS 40
#endif
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity)
{
#if FLOU_MBAR_WAIT == 1
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(bar), "r"(parity), "r"((unsigned)FLOU_MBAR_NS) : "memory");
#elif FLOU_MBAR_WAIT == 2
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "WAIT_%=:\n\t"
        "nanosleep.u32 %2;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@!p bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(bar), "r"(parity), "r"((unsigned)FLOU_MBAR_NS) : "memory");
#else
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
#endif
}

struct ETPick { int e, t; };


// elements per CTA / line threads per CTA: maximise the lane utilisation of the line phase
constexpr ETPick pick_et(int nlines, int per_elem_doubles, int max_kb)
{
    const int ts[7] = {128, 160, 96, 192, 64, 224, 256};
    ETPick best{1, 64};
    int best_util = 0;                                    // in 1/1000
    for (int q = 0; q < 7; q++) {
        const int t = ts[q];
        int e = t / nlines;
        if (e < 1) e = 1;
        const int emax = (max_kb * 1024 / 8) / per_elem_doubles;      // shared memory per CTA
        if (e > emax) e = emax < 1 ? 1 : emax;
        if (e > 64) e = 64;
        const int rounds = (e * nlines + t - 1) / t;
        const int util = (1000 * e * nlines) / (rounds * t);
        if (util > best_util + 20) { best = ETPick{e, t}; best_util = util; }
    }
    return best;
}

// Layout of the warp-specialised kernel (line_kernel_ws.cuh): TL line threads plus one update warp,
// three state buffers and two node-data buffers (WS is kept in the template signature for the
// instance names in profiles/; it is always true).
// NB = nodes without boundaries (Gauss): the split form takes its surface term from entropy-
// projected end states (splitdiv_nb_line); a separate instance, so that the GLL kernels carry none
// of that code (inlined into the exact path it cost the GLL instance 35 %: local-memory frame).
template <int ND_, int NP_, int EQ_, int VOL_, bool CART_, bool WS_ = false, bool NB_ = false>
struct LCfg {
    static constexpr int ND = ND_, NP = NP_, EQ = EQ_, VOL = VOL_;
    static constexpr bool CART = CART_, WS = WS_, NB = NB_;
    static constexpr int NV = (EQ == EQ_ADV) ? 1 : ND + 2;
    static constexpr int NPTS = ipow_c(NP, ND);
    static constexpr int NFP = ipow_c(NP, ND - 1);
    static constexpr int NFACES = 2 * ND;
    static constexpr int NLINES = ND * NFP;
    static constexpr bool HYBRID = (VOL == VOL_HYBRID);
    static constexpr bool SPLIT = (VOL == VOL_SPLIT_STD || VOL == VOL_SPLIT_CHA);
    static_assert(!HYBRID || EQ == EQ_EULER, "HybridDivOperator: Euler equations only");
    // Cartesian split form: the constant metric factor is applied when the partial sums are added
    static constexpr bool FOLD = CART && SPLIT;
    // node data in shared memory: Chandrasekhar (rho, v/2, beta); StdAverage (Q, v, p);
    // strong form: the ND contravariant fluxes; hybrid: the conservative state; advection: q
    static constexpr int NAUX = (EQ == EQ_ADV) ? (SPLIT ? 1 : ND)
                              : (VOL == VOL_SPLIT_CHA ? ND + 2 : (VOL == VOL_SPLIT_STD ? NV + ND + 1 : (HYBRID ? NV : ND * NV)));
    static constexpr int NPART = ND * NV;
    static constexpr int NUB = WS ? 3 : 2;                // state buffers
    static constexpr int NAB = WS ? 2 : 1;                // node-data buffers
    // face-flux block of one face slot, [v][i] padded to a 16-byte multiple (one TMA bulk copy)
    static constexpr int FNB = fn_block(NV, NFP);
    // shared memory per element (doubles): state buffers, tmp, node data, partial sums, the
    // flux blocks of its faces (two sets: the blocks of a group are requested two groups ahead)
    static constexpr int PER_ELEM = ((NUB + 1) * NV + NAB * NAUX + NPART) * NPTS + 2 * NFACES * FNB;
#if defined(FLOU_LINE_E) && defined(FLOU_LINE_T)
    static constexpr int E = FLOU_LINE_E, TL = FLOU_LINE_T;
#else
    // 112 KB per CTA: two CTAs (+ 1 KB each the driver reserves) fill the 228 KB of an SM
    static constexpr int E = pick_et(NLINES, PER_ELEM, 112).e, TL = pick_et(NLINES, PER_ELEM, 112).t;
#endif
    // update warps per CTA and (MAXREG > 0) an explicit register cap instead of the launch bounds
#ifdef FLOU_LINE_NUPD
    static constexpr int NUPD = FLOU_LINE_NUPD;
#else
    static constexpr int NUPD = 1;
#endif
#ifdef FLOU_LINE_MAXREG
    static constexpr int MAXREG = FLOU_LINE_MAXREG;
#else
    static constexpr int MAXREG = 0;
#endif
    // tmp and the new state leave the update warp as TMA bulk stores from shared memory instead of
    // 128-bit global stores.  Measured (profiles/r2_kernel_notes.md): neutral at cfg4 (13.98 vs
    // 13.66-14.16 ms), 7-9 % slower on the small groups of cfg2 / cfg3 (the wait for the stores'
    // shared-memory reads sits on the update warp's critical path): off by default.
#ifdef FLOU_LINE_BULK_STORE
    static constexpr bool BULK_STORE = true;
#else
    static constexpr bool BULK_STORE = false;
#endif
    // TMA loads of tmp / state issued by the line warps (a plane each) after freeP instead of by
    // the update warp (ws_issue_loads).  Issued by line thread 0 alone the freeP wait dropped from
    // 13 % to 7.5 % of the samples but warp 0 arrived late at the line threads' barrier (5.7 % ->
    // 9.8 %): 12.94 vs 12.63 ms.
#ifdef FLOU_LINE_ISSUE_BY_LINE
    static constexpr bool LINE_ISSUE = true;
#else
    static constexpr bool LINE_ISSUE = false;
#endif
    static constexpr int T = WS ? TL + 32 * NUPD : TL;    // threads per CTA
    static constexpr int N = E * NPTS;                    // nodes of a group = plane stride
    static constexpr int L = E * NLINES;                  // line tasks of a group
    static constexpr int ROUNDS = (L + TL - 1) / TL;
    static constexpr int LT = ROUNDS * TL;
    static constexpr bool ONE_ROUND = (ROUNDS == 1);
    static constexpr int OFF_U = 0;                       // [NUB][NV][N]    state of this / the next group(s)
    static constexpr int OFF_T = OFF_U + NUB * NV * N;    // [NV][N]         tmp
    static constexpr int OFF_A = OFF_T + NV * N;          // [NAB][NAUX][N]  node data
    static constexpr int OFF_P = OFF_A + NAB * NAUX * N;  // [ND*NV][N]      partial sums by direction
    static constexpr int OFF_F = (OFF_P + NPART * N + 1) & ~1;  // [2][E*NFACES][FNB]  flux blocks of the faces of this / the next group (16-byte aligned)
    static constexpr int OFF_EC = OFF_F + 2 * E * NFACES * FNB; // [2][E*NFACES] int2: face connectivity of this / the next group
    static constexpr int OFF_BAR = OFF_EC + 2 * E * NFACES;     // 8 mbarriers
    // + 16 iteration counters (int, one per warp) + the flux slots of the faces of an upcoming group (int each)
    static constexpr int OFF_SLOT = OFF_BAR + 8 + 8;
    static constexpr size_t SMEM_BYTES = sizeof(double) * (size_t)(OFF_SLOT + (E * NFACES + 1) / 2);
    static_assert(WS, "the line kernel exists in its warp-specialised form only");
    // registers: a line task holds NP nodes and NP*NV accumulators
    static constexpr int MINB =
#ifdef FLOU_LINE_MINB
        FLOU_LINE_MINB;
#else
        // light line tasks (3-D p=3: 163 registers) fit three CTAs per SM: cfg3 +12 % (profiles/r2_kernel_notes.md)
        (HYBRID || NP * (NAUX + NV) > 64) ? 1
        : (ND == 3 && NP * (NAUX + NV) <= 40 && 3 * (SMEM_BYTES + 1024) <= 233472 ? 3 : 2);
#endif
};

// ------------------------------------------------------------------ two-point fluxes of a line
// Chandrasekhar two-point flux (Equations/Euler.jl:474-536) along the (permuted) axis 0 with a
// unit metric; node data (rho, v/2, beta).  -|v1|^2/4 - |v2|^2/4 + |v_avg|^2 = 2 (v1/2).(v2/2).
// FAST: branch-free (series-only logarithmic means), `redo` is raised when a jump is too large for
// the series -- ten such fluxes then form one basic block the scheduler can interleave.
template <int ND, bool FAST>
__device__ __forceinline__ void tp_cha_axis(double r1, const double *hv1, double b1,
                                            double r2, const double *hv2, double b2,
                                            double inv_gm1, double *F, int &redo)
{
    const double rs = r1 + r2, bs = b1 + b2;
    const double irb = fast_rcp(rs * bs);
    const double irs = irb * bs, ibs = irb * rs;
    double Fr, Fb;
    if (FAST) logmean_F2_series(r1, r2, irs, b1, b2, ibs, Fr, Fb, redo);
    else logmean_F2(r1, r2, irs, b1, b2, ibs, Fr, Fb);
    const double rho = 0.5 * rs * fast_rcp(Fr);
    const double p = rs * ibs * 0.5;
    double vavg[ND], dot = hv1[0] * hv2[0];
#pragma unroll
    for (int c = 1; c < ND; c++) dot = fma(hv1[c], hv2[c], dot);
#pragma unroll
    for (int c = 0; c < ND; c++) vavg[c] = hv1[c] + hv2[c];
    const double h = fma(fma(Fb, inv_gm1, Fr), ibs, 2.0 * dot);
    const double mdot = rho * vavg[0];
    F[0] = mdot;
    F[1] = fma(mdot, vavg[0], p);
#pragma unroll
    for (int c = 1; c < ND; c++) F[1 + c] = mdot * vavg[c];
    F[ND + 1] = mdot * h;
}

// the same contracted with a general metric vector n (curved / unstructured elements)
template <int ND, bool FAST>
__device__ __forceinline__ void tp_cha_n(double r1, const double *hv1, double b1,
                                         double r2, const double *hv2, double b2,
                                         double inv_gm1, const double *n, double *F, int &redo)
{
    const double rs = r1 + r2, bs = b1 + b2;
    const double irb = fast_rcp(rs * bs);
    const double irs = irb * bs, ibs = irb * rs;
    double Fr, Fb;
    if (FAST) logmean_F2_series(r1, r2, irs, b1, b2, ibs, Fr, Fb, redo);
    else logmean_F2(r1, r2, irs, b1, b2, ibs, Fr, Fb);
    const double rho = 0.5 * rs * fast_rcp(Fr);
    const double p = rs * ibs * 0.5;
    double vavg[ND], dot = hv1[0] * hv2[0], vn = 0.0;
#pragma unroll
    for (int c = 1; c < ND; c++) dot = fma(hv1[c], hv2[c], dot);
#pragma unroll
    for (int c = 0; c < ND; c++) { vavg[c] = hv1[c] + hv2[c]; vn = fma(vavg[c], n[c], vn); }
    const double h = fma(fma(Fb, inv_gm1, Fr), ibs, 2.0 * dot);
    const double mdot = rho * vn;
    F[0] = mdot;
#pragma unroll
    for (int c = 0; c < ND; c++) F[1 + c] = fma(mdot, vavg[c], p * n[c]);
    F[ND + 1] = mdot * h;
}

// physical flux of ONE node from (rho, v/2, beta), contracted with n: the diagonal entry
// F#(i,i) = F~_i of the split form (OpDivergence.jl:252)
template <int ND>
__device__ __forceinline__ void diag_cha_n(double r, const double *hv, double b, double inv_gm1,
                                           const double *n, double *F)
{
    const double ib = fast_rcp(b);
    const double p = 0.5 * r * ib;
    double q = 0.0, vn = 0.0;
#pragma unroll
    for (int c = 0; c < ND; c++) { q = fma(hv[c], hv[c], q); vn = fma(2.0 * hv[c], n[c], vn); }
    const double h = fma((inv_gm1 + 1.0) * 0.5, ib, 2.0 * q);     // (E + p)/rho
    const double mdot = r * vn;
    F[0] = mdot;
#pragma unroll
    for (int c = 0; c < ND; c++) F[1 + c] = fma(mdot, 2.0 * hv[c], p * n[c]);
    F[ND + 1] = mdot * h;
}

// Pair (J, L) of a line and, recursively, every pair after it: template recursion instead of
// `#pragma unroll` so that every array index is a compile-time constant (nvcc declined to unroll
// the equivalent double loop and moved the node data to local memory).
template <class C, bool FAST, int J, int L>
__device__ __forceinline__ void cha_pair_rec(const KParams &P, const double (&r)[C::NP], const double (&hv)[C::NP][C::ND],
                                             const double (&b)[C::NP],
                                             const double (&mt)[C::CART ? 1 : C::NP][C::CART ? 1 : C::ND],
                                             double (&a_)[C::NP][C::NV], int &redo)
{
    constexpr int ND = C::ND, NP = C::NP, NV = C::NV;
    if constexpr (L < NP) {
        double F[NV];
        if constexpr (C::CART) {
            tp_cha_axis<ND, FAST>(r[J], hv[J], b[J], r[L], hv[L], b[L], P.fp.inv_gm1, F, redo);
        } else {
            double n[ND];
#pragma unroll
            for (int c = 0; c < ND; c++) n[c] = 0.5 * (mt[J][c] + mt[L][c]);
            tp_cha_n<ND, FAST>(r[J], hv[J], b[J], r[L], hv[L], b[L], P.fp.inv_gm1, n, F, redo);
        }
        const double djl = P.Dvol[J + NP * L], dlj = P.Dvol[L + NP * J];
#pragma unroll
        for (int v = 0; v < NV; v++) {
            a_[J][v] = fma(-djl, F[v], a_[J][v]);
            a_[L][v] = fma(-dlj, F[v], a_[L][v]);
        }
        if constexpr (L + 1 < NP) cha_pair_rec<C, FAST, J, L + 1>(P, r, hv, b, mt, a_, redo);
        else if constexpr (J + 2 < NP) cha_pair_rec<C, FAST, J + 1, J + 2>(P, r, hv, b, mt, a_, redo);
    }
}

// All NP(NP-1)/2 Chandrasekhar pair fluxes of one line, applied with D# to both partners.
// FAST = branch-free; returns true when some density or beta ratio is outside the series range
// of the logarithmic mean, in which case the caller redoes the line with FAST = false.
template <class C, bool FAST>
__device__ __forceinline__ bool cha_pairs(const KParams &P, const double (&r)[C::NP], const double (&hv)[C::NP][C::ND],
                                          const double (&b)[C::NP],
                                          const double (&mt)[C::CART ? 1 : C::NP][C::CART ? 1 : C::ND],
                                          double (&a_)[C::NP][C::NV])
{
    constexpr int ND = C::ND, NP = C::NP, NV = C::NV;
    int redo = 0;        // running maximum of the series arguments (high words)
    // the diagonal of D# is analytically zero on GLL nodes; when the host found entries that are
    // not round-off (diag_mask), the line goes to the exact path, which applies them
    if (FAST) {
        if (P.diag_mask) return true;
    } else {
#pragma unroll
        for (int j = 0; j < NP; j++) {
            if (P.diag_mask & (1 << j)) {
                double n[ND], F[NV];
#pragma unroll
                for (int c = 0; c < ND; c++) n[c] = C::CART ? (c == 0 ? 1.0 : 0.0) : mt[C::CART ? 0 : j][C::CART ? 0 : c];
                diag_cha_n<ND>(r[j], hv[j], b[j], P.fp.inv_gm1, n, F);
                const double djj = P.Dvol[j + NP * j];
#pragma unroll
                for (int v = 0; v < NV; v++) a_[j][v] = fma(-djj, F[v], a_[j][v]);
            }
        }
    }
    cha_pair_rec<C, FAST, 0, 1>(P, r, hv, b, mt, a_, redo);
    return FAST && series_out_of_range(redo);
}

// HybridDivOperator, volume term of ONE line in direction D (_vol_hybrid_tensorproduct!,
// OpDivergence.jl:557-612) on a Cartesian sub-grid (frames PhysicalRegions.jl:72-148: n = e_D,
// sub-cell face Jacobian = the element's metric factor of direction D):
//   Fbar[ii] = sum_{il < ii <= ik} 2 w[il] D[il,ik] F#(Q_il, Q_ik)       telescopic split form
//   Fv       = rotate2phys(F*(rotate2face(Q_{ii-1}), rotate2face(Q_ii))) * Js     sub-cell FV flux
//   b        = (W_ii - W_{ii-1}) . (Fbar[ii] - Fv),   delta = max((sqrt(b^2+c) - b)/sqrt(b^2+c), 1/2)
//   Fbar[ii] = (1 - delta) Fv + delta Fbar[ii];       dQ_ii += (Fbar[ii] - Fbar[ii+1]) / w[ii]
// Each pair flux is evaluated once and added to every sub-cell interface between its two nodes.
template <class C, int D>
__device__ __forceinline__ void hybrid_line(const KParams &P, const double (&Q)[C::NP][C::NV],
                                            int64_t gnode0, int stride, int64_t sub0,
                                            double (&acc)[C::NP][C::NV])
{
    constexpr int ND = C::ND, NP = C::NP, NV = C::NV, EQ = C::EQ;
    constexpr bool CART = C::CART;
    const double g = P.fp.gamma;
    const double js = CART ? P.cmet[D] : 0.0;
    double vel[NP][ND], hv[NP][ND], q[NP], pr[NP], beta[NP], W[NP][NV];
    // metric vector of direction D at the line's nodes, Ja[:, D] (general geometry; constant
    // cmet[D] e_D on Cartesian meshes, PhysicalRegions.jl:370-408)
    double mt[CART ? 1 : NP][CART ? 1 : ND];
    if (!CART) {
#pragma unroll
        for (int j = 0; j < NP; j++)
#pragma unroll
            for (int c = 0; c < ND; c++) mt[CART ? 0 : j][CART ? 0 : c] = __ldg(P.metric + gnode0 + j * stride + P.ndof * (c + ND * D));
    }
#pragma unroll
    for (int j = 0; j < NP; j++) {
        NodeAux<ND> A;
        node_aux<ND>(Q[j], g, A);
        double m2 = 0.0, q_ = 0.0;
#pragma unroll
        for (int c = 0; c < ND; c++) {
            vel[j][c] = A.vel[c]; hv[j][c] = 0.5 * A.vel[c];
            q_ += A.vel[c] * A.vel[c];
            m2 += Q[j][1 + c] * Q[j][1 + c];
        }
        q[j] = q_; pr[j] = A.p; beta[j] = A.beta;
        // vars_cons2entropy (FlouCommon/Euler.jl:273-307), entropy :202-206
        const double s = log(A.p) - g * log(Q[j][0]);
        W[j][0] = (g - s) / (g - 1.0) - m2 / Q[j][0] / (2.0 * A.p);
#pragma unroll
        for (int c = 0; c < ND; c++) W[j][1 + c] = Q[j][1 + c] / A.p;
        W[j][ND + 1] = -Q[j][0] / A.p;
    }

    double Fb[NP + 1][NV];
#pragma unroll
    for (int ii = 0; ii <= NP; ii++)
#pragma unroll
        for (int v = 0; v < NV; v++) Fb[ii][v] = 0.0;
#pragma unroll
    for (int il = 0; il < NP - 1; il++)
#pragma unroll
        for (int ik = il + 1; ik < NP; ik++) {
            double F[NV], n[ND];
#pragma unroll
            for (int c = 0; c < ND; c++)
                n[c] = CART ? ((c == D) ? js : 0.0) : 0.5 * (mt[CART ? 0 : il][CART ? 0 : c] + mt[CART ? 0 : ik][CART ? 0 : c]);
            if (P.tpflux == FX_CHA)
                tp_chandrasekhar<ND>(Q[il][0], hv[il], q[il], beta[il], Q[ik][0], hv[ik], q[ik], beta[ik],
                                     P.fp.inv_gm1, n, F);
            else
                tp_stdavg<ND>(Q[il], vel[il], pr[il], Q[ik], vel[ik], pr[ik], n, F);
            const double c = 2.0 * P.w1d[il] * P.Dvol[il + NP * ik];
#pragma unroll
            for (int ii = il + 1; ii <= ik; ii++)
#pragma unroll
                for (int v = 0; v < NV; v++) Fb[ii][v] = fma(c, F[v], Fb[ii][v]);
        }
#pragma unroll
    for (int ii = 1; ii < NP; ii++) {
        double Rl[NV], Rr[NV], Fn[NV], Fv[NV], jsi = js;
        if (CART) {
            rot2face_c<ND, EQ, 2 * D + 1>(Q[ii - 1], Rl);
            rot2face_c<ND, EQ, 2 * D + 1>(Q[ii], Rr);
            euler_numflux<ND>(P.fp, Rl, Rr, Fn);
            rot2phys_c<ND, EQ, 2 * D + 1>(Fn, Fv);
        } else {
            // frame and Jacobian of sub-cell interface ii of this line (PhysicalRegions.jl:179-292)
            double fr[3 * ND];
#pragma unroll
            for (int c = 0; c < 3 * ND; c++) fr[c] = (c < ND * ND || ND == 3) ? __ldg(P.sub_frames + (sub0 + ii) * (3 * ND) + c) : 0.0;
            jsi = __ldg(P.sub_jac + sub0 + ii);
            rotate2face<ND, EQ>(Q[ii - 1], fr, Rl);
            rotate2face<ND, EQ>(Q[ii], fr, Rr);
            euler_numflux<ND>(P.fp, Rl, Rr, Fn);
            rotate2phys<ND, EQ>(Fn, fr, Fv);
        }
        double b = 0.0;
#pragma unroll
        for (int v = 0; v < NV; v++) {
            Fv[v] *= jsi;
            b += (W[ii][v] - W[ii - 1][v]) * (Fb[ii][v] - Fv[v]);
        }
        double delta = sqrt(b * b + P.blend);                 // _hybrid_compute_delta (Fisher)
        delta = (delta - b) / delta;
        delta = fmax(delta, 0.5);
#pragma unroll
        for (int v = 0; v < NV; v++) Fb[ii][v] = (1.0 - delta) * Fv[v] + delta * Fb[ii][v];
    }
#pragma unroll
    for (int j = 0; j < NP; j++) {
#pragma unroll
        for (int v = 0; v < NV; v++) acc[j][v] = (Fb[j][v] - Fb[j + 1][v]) / P.w1d[j];
    }
}

// HybridDivOperator on nodes WITHOUT boundaries (Gauss), one line in direction D: everything is a
// surface contribution (_hybrid_nb_surface_contribution!, OpDivergence.jl:629-779):
//   F#            diagonal = contravariant flux of the node, off-diagonal = two-point fluxes
//   Fl_i, Fr_i    two-point fluxes against the entropy-projected end states, minus l'Fl + Fn_left
//                 and r'Fr - Fn_right (_flux_splitdiv_nb_tensorproduct!)
//   Fbar[0] = -Fn_left, Fbar[NP] = Fn_right,
//   Fbar[ii+1] = Fbar[ii] + w_ii (D# F#)_ii - l_ii Fl_ii + r_ii Fr_ii,  blended with the FV flux of
//                interface ii+1 (the blended value feeds the next step of the recursion),
//   dQ_ii += (Fbar[ii] - Fbar[ii+1]) / w_ii
// fnl / fnr: the element-side face fluxes of this line (sign applied).  Dvol = D# here.
template <class C, int D>
__device__ __forceinline__ void hybrid_nb_line(const KParams &P, const double (&Q)[C::NP][C::NV],
                                               const double (&fnl)[C::NV], const double (&fnr)[C::NV],
                                               int64_t gnode0, int stride, int64_t sub0,
                                               double (&acc)[C::NP][C::NV])
{
    constexpr int ND = C::ND, NP = C::NP, NV = C::NV, EQ = C::EQ;
    constexpr bool CART = C::CART;
    const double g = P.fp.gamma;
    double rho[NP], vel[NP][ND], pr[NP], W[NP][NV], mt[NP][ND], nend[2][ND], Wp[2][NV];
#pragma unroll
    for (int j = 0; j < NP; j++)
#pragma unroll
        for (int c = 0; c < ND; c++)
            mt[j][c] = CART ? ((c == D) ? P.cmet[D] : 0.0) : __ldg(P.metric + gnode0 + j * stride + P.ndof * (c + ND * D));
#pragma unroll
    for (int e = 0; e < 2; e++) {
#pragma unroll
        for (int c = 0; c < ND; c++) {
            if (CART) nend[e][c] = (c == D) ? P.cmet[D] : 0.0;
            else {
                const int64_t si = sub0 + (e ? NP : 0);
                nend[e][c] = __ldg(P.sub_frames + si * (3 * ND) + c) * __ldg(P.sub_jac + si);
            }
        }
#pragma unroll
        for (int v = 0; v < NV; v++) Wp[e][v] = 0.0;
    }
#pragma unroll
    for (int j = 0; j < NP; j++) {
        NodeAux<ND> A;
        node_aux<ND>(Q[j], g, A);
        rho[j] = Q[j][0]; pr[j] = A.p;
        double m2 = 0.0;
#pragma unroll
        for (int c = 0; c < ND; c++) { vel[j][c] = A.vel[c]; m2 += Q[j][1 + c] * Q[j][1 + c]; }
        const double s = log(A.p) - g * log(Q[j][0]);
        W[j][0] = (g - s) / (g - 1.0) - m2 / Q[j][0] / (2.0 * A.p);
#pragma unroll
        for (int c = 0; c < ND; c++) W[j][1 + c] = Q[j][1 + c] / A.p;
        W[j][ND + 1] = -Q[j][0] / A.p;
#pragma unroll
        for (int v = 0; v < NV; v++) { Wp[0][v] = fma(P.lm[j], W[j][v], Wp[0][v]); Wp[1][v] = fma(P.lp[j], W[j][v], Wp[1][v]); }
    }
    // two-point flux of two states given as (rho, vel, p), contracted with n
    auto tp = [&](double r1, const double *v1, double p1, double r2, const double *v2, double p2,
                  const double *n, double *F) {
        if (P.tpflux == FX_CHA) {
            double h1[ND], h2[ND], q1 = 0.0, q2 = 0.0;
#pragma unroll
            for (int c = 0; c < ND; c++) { h1[c] = 0.5 * v1[c]; h2[c] = 0.5 * v2[c]; q1 += v1[c] * v1[c]; q2 += v2[c] * v2[c]; }
            tp_chandrasekhar<ND>(r1, h1, q1, r1 / (2.0 * p1), r2, h2, q2, r2 / (2.0 * p2), P.fp.inv_gm1, n, F);
        } else {
            double Q1[NV], Q2[NV], q1 = 0.0, q2 = 0.0;
            Q1[0] = r1; Q2[0] = r2;
#pragma unroll
            for (int c = 0; c < ND; c++) { Q1[1 + c] = r1 * v1[c]; Q2[1 + c] = r2 * v2[c]; q1 += v1[c] * v1[c]; q2 += v2[c] * v2[c]; }
            Q1[ND + 1] = p1 * P.fp.inv_gm1 + 0.5 * r1 * q1;
            Q2[ND + 1] = p2 * P.fp.inv_gm1 + 0.5 * r2 * q2;
            tp_stdavg<ND>(Q1, v1, p1, Q2, v2, p2, n, F);
        }
    };
    // entropy-projected end states: vars_entropy2prim (FlouCommon/Euler.jl:309-334)
    double re[2], ve[2][ND], pe[2];
#pragma unroll
    for (int e = 0; e < 2; e++) {
        double q = 0.0;
#pragma unroll
        for (int c = 0; c < ND; c++) { ve[e][c] = -Wp[e][1 + c] / Wp[e][ND + 1]; q += ve[e][c] * ve[e][c]; }
        const double s = g - (g - 1.0) * (Wp[e][0] - Wp[e][ND + 1] * q / 2.0);
        pe[e] = pow(pow(-Wp[e][ND + 1], g) * exp(s), 1.0 / (1.0 - g));
        re[e] = -pe[e] * Wp[e][ND + 1];
    }
    // (D# F#)_ii for the first NP-1 nodes, and Fl, Fr
    double DF[NP][NV], Fl[NP][NV], Fr[NP][NV], lFl[NV], rFr[NV];
#pragma unroll
    for (int v = 0; v < NV; v++) { lFl[v] = 0.0; rFr[v] = 0.0; }
#pragma unroll
    for (int i = 0; i < NP; i++) {
        double Fc[NV];
#pragma unroll
        for (int v = 0; v < NV; v++) DF[i][v] = 0.0;
        // diagonal: contravariant flux of node i (volumeflux contracted with Ja_i[:, D])
#pragma unroll
        for (int c = 0; c < ND; c++) {
            if (CART && c != D) continue;
            euler_flux_dir<ND>(Q[i], vel[i], pr[i], c, Fc);
            const double dii = P.Dvol[i + NP * i] * mt[i][c];
#pragma unroll
            for (int v = 0; v < NV; v++) DF[i][v] = fma(dii, Fc[v], DF[i][v]);
        }
#pragma unroll
        for (int e = 0; e < 2; e++) {
            double n[ND];
#pragma unroll
            for (int c = 0; c < ND; c++) n[c] = 0.5 * (mt[i][c] + nend[e][c]);
            tp(rho[i], vel[i], pr[i], re[e], ve[e], pe[e], n, e ? Fr[i] : Fl[i]);
        }
#pragma unroll
        for (int v = 0; v < NV; v++) { lFl[v] = fma(P.lm[i], Fl[i][v], lFl[v]); rFr[v] = fma(P.lp[i], Fr[i][v], rFr[v]); }
    }
#pragma unroll
    for (int i = 0; i < NP - 1; i++)
#pragma unroll
        for (int l = i + 1; l < NP; l++) {
            double n[ND], F[NV];
#pragma unroll
            for (int c = 0; c < ND; c++) n[c] = 0.5 * (mt[i][c] + mt[l][c]);
            tp(rho[i], vel[i], pr[i], rho[l], vel[l], pr[l], n, F);
            const double dil = P.Dvol[i + NP * l], dli = P.Dvol[l + NP * i];
#pragma unroll
            for (int v = 0; v < NV; v++) { DF[i][v] = fma(dil, F[v], DF[i][v]); DF[l][v] = fma(dli, F[v], DF[l][v]); }
        }
    double Fb[NP + 1][NV];
#pragma unroll
    for (int v = 0; v < NV; v++) { Fb[0][v] = -fnl[v]; Fb[NP][v] = fnr[v]; }
#pragma unroll
    for (int ii = 0; ii < NP - 1; ii++) {
#pragma unroll
        for (int v = 0; v < NV; v++)
            Fb[ii + 1][v] = Fb[ii][v] + DF[ii][v] * P.w1d[ii] - P.lm[ii] * (Fl[ii][v] - (lFl[v] + fnl[v]))
                          + P.lp[ii] * (Fr[ii][v] - (rFr[v] - fnr[v]));
        double Rl[NV], Rr[NV], Fn[NV], Fv[NV], jsi = P.cmet[CART ? D : 0];
        if (CART) {
            rot2face_c<ND, EQ, 2 * D + 1>(Q[ii], Rl);
            rot2face_c<ND, EQ, 2 * D + 1>(Q[ii + 1], Rr);
            euler_numflux<ND>(P.fp, Rl, Rr, Fn);
            rot2phys_c<ND, EQ, 2 * D + 1>(Fn, Fv);
        } else {
            double fr[3 * ND];
#pragma unroll
            for (int c = 0; c < 3 * ND; c++) fr[c] = (c < ND * ND || ND == 3) ? __ldg(P.sub_frames + (sub0 + ii + 1) * (3 * ND) + c) : 0.0;
            jsi = __ldg(P.sub_jac + sub0 + ii + 1);
            rotate2face<ND, EQ>(Q[ii], fr, Rl);
            rotate2face<ND, EQ>(Q[ii + 1], fr, Rr);
            euler_numflux<ND>(P.fp, Rl, Rr, Fn);
            rotate2phys<ND, EQ>(Fn, fr, Fv);
        }
        double b = 0.0;
#pragma unroll
        for (int v = 0; v < NV; v++) {
            Fv[v] *= jsi;
            b += (W[ii + 1][v] - W[ii][v]) * (Fb[ii + 1][v] - Fv[v]);
        }
        double delta = sqrt(b * b + P.blend);
        delta = (delta - b) / delta;
        delta = fmax(delta, 0.5);
#pragma unroll
        for (int v = 0; v < NV; v++) Fb[ii + 1][v] = (1.0 - delta) * Fv[v] + delta * Fb[ii + 1][v];
    }
#pragma unroll
    for (int j = 0; j < NP; j++)
#pragma unroll
        for (int v = 0; v < NV; v++) acc[j][v] = (Fb[j][v] - Fb[j + 1][v]) / P.w1d[j];
}

// Shared-memory addresses of the kernel's mbarriers (line_kernel_ws.cuh), from the layout alone
template <class C>
struct WsBars {
    __device__ __forceinline__ static unsigned base()
    {
        extern __shared__ __align__(16) double lsmem[];
        return (unsigned)__cvta_generic_to_shared(lsmem + C::OFF_BAR);
    }
    __device__ __forceinline__ static unsigned fullU(int b) { return base() + 8 * b; }
    __device__ __forceinline__ static unsigned fullP() { return base() + 24; }
    __device__ __forceinline__ static unsigned freeP() { return base() + 32; }
    __device__ __forceinline__ static unsigned fullT() { return base() + 40; }
    __device__ __forceinline__ static unsigned fullF(int b) { return base() + 48 + 8 * b; }
};

// Surface term of the split form on nodes WITHOUT boundaries, one line
// (_flux_splitdiv_nb_tensorproduct! + _surf_splitdiv_nb_tensorproduct!, OpDivergence.jl:389-437):
//   W_i = vars_cons2entropy(Q_i);  Q(l) = vars_entropy2cons(l' W),  Q(r) = vars_entropy2cons(r' W)
//   Fl_i = F#(Q_i, Q(l)),  Fr_i = F#(Q_i, Q(r))
//   Fl_i -= l' Fl + Fn_left,   Fr_i -= r' Fr - Fn_right,   dQ_i += dg_l[i] Fl_i - dg_r[i] Fr_i
// in the folded Cartesian form of the caller: unit metric along the (permuted) axis 0, face fluxes
// fnl / fnr (master-outward, in the caller's variable order) pre-divided by the metric factor
// (wl, wr carry sign and 1/metric), components in the order pc.
template <class C>
__device__ __forceinline__ void splitdiv_nb_line(const KParams &P, const double *sA,
                                              const double (&fnl)[C::NV], const double (&fnr)[C::NV],
                                              int base, int stride, const int (&pc)[C::ND],
                                              double wl, double wr, int d, int64_t gnode0, int64_t sub0,
                                              double (&acc)[C::NP][C::NV])
{
    constexpr int ND = C::ND, NP = C::NP, NV = C::NV, N = C::N, VOL = C::VOL;
    constexpr bool CART = C::CART;
    const double g = P.fp.gamma;
    // general geometry: metric vector Ja[:, d] of the line's nodes and the sub-grid normals times
    // Jacobians at the two ends of the line (frames[dir][i1].n * Js[dir][i1], OpDivergence.jl:396-399)
    double mt[CART ? 1 : NP][CART ? 1 : ND], nend[2][ND];
    if (!CART) {
#pragma unroll
        for (int j = 0; j < NP; j++)
#pragma unroll
            for (int c = 0; c < ND; c++) mt[CART ? 0 : j][CART ? 0 : c] = __ldg(P.metric + gnode0 + j * stride + P.ndof * (c + ND * d));
#pragma unroll
        for (int e = 0; e < 2; e++) {
            const int64_t si = sub0 + (e ? NP : 0);
            const double js = __ldg(P.sub_jac + si);
#pragma unroll
            for (int c = 0; c < ND; c++) nend[e][c] = __ldg(P.sub_frames + si * (3 * ND) + c) * js;
        }
    }
    double rho[NP], vel[NP][ND], pr[NP], Wl[NV], Wr[NV];
#pragma unroll
    for (int v = 0; v < NV; v++) { Wl[v] = 0.0; Wr[v] = 0.0; }
#pragma unroll
    for (int j = 0; j < NP; j++) {
        const int node = base + j * stride;
        if (VOL == VOL_SPLIT_CHA) {          // node data (rho, v/2, beta)
            rho[j] = sA[node];
#pragma unroll
            for (int c = 0; c < ND; c++) vel[j][c] = 2.0 * sA[(1 + pc[c]) * N + node];
            pr[j] = rho[j] / (2.0 * sA[(ND + 1) * N + node]);
        } else {                             // node data (Q, v, p)
            rho[j] = sA[node];
#pragma unroll
            for (int c = 0; c < ND; c++) vel[j][c] = sA[(NV + pc[c]) * N + node];
            pr[j] = sA[(NV + ND) * N + node];
        }
        double q = 0.0, W[NV];
#pragma unroll
        for (int c = 0; c < ND; c++) q += vel[j][c] * vel[j][c];
        const double s = log(pr[j]) - g * log(rho[j]);                  // FlouCommon/Euler.jl:202-206
        W[0] = (g - s) / (g - 1.0) - rho[j] * q / (2.0 * pr[j]);        // :273-307
#pragma unroll
        for (int c = 0; c < ND; c++) W[1 + c] = rho[j] * vel[j][c] / pr[j];
        W[ND + 1] = -rho[j] / pr[j];
#pragma unroll
        for (int v = 0; v < NV; v++) { Wl[v] = fma(P.lm[j], W[v], Wl[v]); Wr[v] = fma(P.lp[j], W[v], Wr[v]); }
    }
    // vars_entropy2prim (FlouCommon/Euler.jl:309-334) of both projections: e = 0 left, 1 right
    double re[2], ve[2][ND], pe[2];
#pragma unroll
    for (int e = 0; e < 2; e++) {
        const double *W = e ? Wr : Wl;
        double q = 0.0;
#pragma unroll
        for (int c = 0; c < ND; c++) { ve[e][c] = -W[1 + c] / W[ND + 1]; q += ve[e][c] * ve[e][c]; }
        const double s = g - (g - 1.0) * (W[0] - W[ND + 1] * q / 2.0);
        pe[e] = pow(pow(-W[ND + 1], g) * exp(s), 1.0 / (1.0 - g));
        re[e] = -pe[e] * W[ND + 1];
    }
    double Fl[NP][NV], Fr[NP][NV], lFl[NV], rFr[NV];
#pragma unroll
    for (int v = 0; v < NV; v++) { lFl[v] = 0.0; rFr[v] = 0.0; }
#pragma unroll
    for (int j = 0; j < NP; j++) {
#pragma unroll
        for (int e = 0; e < 2; e++) {
            double *F = e ? Fr[j] : Fl[j];
            // averaged metric vector: unit axis 0 in the folded Cartesian form, (Ja_i + n_end)/2 else
            double n[ND];
#pragma unroll
            for (int c = 0; c < ND; c++)
                n[c] = CART ? ((c == 0) ? 1.0 : 0.0) : 0.5 * (mt[CART ? 0 : j][CART ? 0 : c] + nend[e][c]);
            if (VOL == VOL_SPLIT_CHA) {
                double h1[ND], h2[ND];
                int dummy = 0;
#pragma unroll
                for (int c = 0; c < ND; c++) { h1[c] = 0.5 * vel[j][c]; h2[c] = 0.5 * ve[e][c]; }
                if (CART)
                    tp_cha_axis<ND, false>(rho[j], h1, rho[j] / (2.0 * pr[j]), re[e], h2, re[e] / (2.0 * pe[e]),
                                           P.fp.inv_gm1, F, dummy);
                else
                    tp_cha_n<ND, false>(rho[j], h1, rho[j] / (2.0 * pr[j]), re[e], h2, re[e] / (2.0 * pe[e]),
                                        P.fp.inv_gm1, n, F, dummy);
            } else {
                double Q1[NV], Q2[NV], q1 = 0.0, q2 = 0.0;
                Q1[0] = rho[j]; Q2[0] = re[e];
#pragma unroll
                for (int c = 0; c < ND; c++) {
                    Q1[1 + c] = rho[j] * vel[j][c]; Q2[1 + c] = re[e] * ve[e][c];
                    q1 += vel[j][c] * vel[j][c]; q2 += ve[e][c] * ve[e][c];
                }
                Q1[ND + 1] = pr[j] * P.fp.inv_gm1 + 0.5 * rho[j] * q1;
                Q2[ND + 1] = pe[e] * P.fp.inv_gm1 + 0.5 * re[e] * q2;
                tp_stdavg<ND>(Q1, vel[j], pr[j], Q2, ve[e], pe[e], n, F);
            }
        }
#pragma unroll
        for (int v = 0; v < NV; v++) { lFl[v] = fma(P.lm[j], Fl[j][v], lFl[v]); rFr[v] = fma(P.lp[j], Fr[j][v], rFr[v]); }
    }
#pragma unroll
    for (int v = 0; v < NV; v++) {
        const double a0 = lFl[v] + wl * fnl[v], b0 = rFr[v] - wr * fnr[v];
#pragma unroll
        for (int j = 0; j < NP; j++)
            acc[j][v] += P.dgl[j] * (Fl[j][v] - a0) - P.dgr[j] * (Fr[j][v] - b0);
    }
}

// One line task: volume term of the line's NP nodes in direction d plus the lift of the two face
// fluxes at its ends, written as the partial sums of direction d.  FAST (Chandrasekhar only):
// branch-free pair fluxes; returns true when the line has to be redone with FAST = false.
// `it`: the warp's iteration counter in shared memory.  Everything that depends on the iteration --
// node-data buffer, flux-block buffer, mbarrier parities -- is derived from a fresh read of it where
// it is needed, so that nothing but the line's own data is live across the pair fluxes (the counter
// and the buffer pointers used to be spilled to local memory there; with the shared-memory carve-out
// this kernel asks for, L1 is too small to keep those lines: ncu showed 15 % of the stall samples on
// the reloads, profiles/r2_kernel_notes.md).
// 16-byte aligned planes and an even node count per group: the state moves with TMA bulk copies
// (and phase 3 works on node pairs); otherwise 8-byte cp.async
template <class C>
__device__ __forceinline__ bool ws_wide(const KParams &P)
{
    return ((P.ndof & 1) == 0) && (((int64_t)P.elem_first * C::NPTS & 1) == 0) && ((C::N & 1) == 0) &&
           ((reinterpret_cast<uintptr_t>(P.u_in) & 15) == 0) && ((reinterpret_cast<uintptr_t>(P.tmp) & 15) == 0) &&
           ((reinterpret_cast<uintptr_t>(P.u_out) & 15) == 0) && ((reinterpret_cast<uintptr_t>(P.k_out) & 15) == 0) &&
           (C::CART || (reinterpret_cast<uintptr_t>(P.jac) & 15) == 0);
}

// TMA loads of the next groups issued by the LINE warps (C::LINE_ISSUE; lane 0 of line warp w takes
// the planes w, w + NW, ... of both copies) once the update warp has released the buffers (freeP):
// tmp of the group the warp is working on and the state two groups further on, into the buffers
// phase 3 of the previous group has just finished with.  The update warp is the kernel's critical
// path and these copies with their descriptors are ~120 of its ~1100 instructions per group; spread
// over the line warps they cost each ~25.  (`wide` only: the cp.async fallback stays on the update
// warp.)  The mbarriers then count NW arrivals, one per issuing warp.
template <class C>
__device__ __forceinline__ void ws_issue_loads(const KParams &P, int i_, int w)
{
    extern __shared__ __align__(16) double lsmem[];
    constexpr int E = C::E, NV = C::NV, N = C::N, NPTS = C::NPTS, NW = C::TL / 32;
    int cta, ncta;
    asm volatile("mov.u32 %0, %%ctaid.x;" : "=r"(cta));
    asm volatile("mov.u32 %0, %%nctaid.x;" : "=r"(ncta));
    const int64_t ndof = P.ndof;
    const unsigned long long pol = l2_evict_first();
    const int nplanes = w < NV ? (NV - w + NW - 1) / NW : 0;      // planes w, w + NW, ... of this warp
    const int g = cta + i_ * ncta;
    if (P.mode == MODE_STAGE) {
        const int nn = min(E, P.elem_count - g * E) * NPTS;
        const double *s0 = P.tmp + (int64_t)(P.elem_first + g * E) * NPTS;
        mbar_expect_tx(WsBars<C>::fullT(), (unsigned)(nplanes * nn * sizeof(double)));
        for (int v = w; v < NV; v += NW)
            bulk_g2s(lsmem + C::OFF_T + v * N, s0 + ndof * v, (unsigned)(nn * sizeof(double)), WsBars<C>::fullT(), pol);
    }
    const int g2 = g + 2 * ncta;
    if (g2 * E < P.elem_count) {
        const int ub = (i_ + 2) % 3;
        const int nn = min(E, P.elem_count - g2 * E) * NPTS;
        const double *s0 = P.u_in + (int64_t)(P.elem_first + g2 * E) * NPTS;
        mbar_expect_tx(WsBars<C>::fullU(ub), (unsigned)(nplanes * nn * sizeof(double)));
        for (int v = w; v < NV; v += NW)
            bulk_g2s(lsmem + C::OFF_U + ub * (NV * N) + v * N, s0 + ndof * v, (unsigned)(nn * sizeof(double)), WsBars<C>::fullU(ub), pol);
    }
}

template <class C, bool FAST>
__device__ __forceinline__ bool line_task(const KParams &P, int task, int64_t dof0, const volatile int *it)
{
    extern __shared__ __align__(16) double lsmem[];
    const double *const sA = lsmem + C::OFF_A + (*it & 1) * (C::NAUX * C::N);
    double *const sP = lsmem + C::OFF_P;
    constexpr int ND = C::ND, NP = C::NP, EQ = C::EQ, VOL = C::VOL, NV = C::NV;
    constexpr int NPTS = C::NPTS, NFP = C::NFP, NLINES = C::NLINES, NFACES = C::NFACES;
    constexpr int N = C::N, FNB = C::FNB;
    constexpr bool CART = C::CART, SPLIT = C::SPLIT, FOLD = C::FOLD;
    const int64_t ndof = P.ndof;
    bool redo = false;
        // indices of the line: element of the group, direction, line of that direction, first node
        // and node stride; pc: Cartesian split form, momentum components in the cyclic order that
        // puts the line's direction first (the flux code is then the same instruction stream for every d)
        int el, d, k, base, stride, pc[ND];
        double *pt;
        auto line_indices = [&](int t) {
            el = t / NLINES;
            const int r_ = t - el * NLINES;
            d = r_ / NFP; k = r_ - d * NFP;
            line_of<ND, NP>(d, k, base, stride);
            base += el * NPTS;
            pt = sP + (d * NV) * N;
#pragma unroll
            for (int c = 0; c < ND; c++) { const int s_ = d + c; pc[c] = FOLD ? (s_ >= ND ? s_ - ND : s_) : c; }
        };
        line_indices(task);
        auto var_of = [&](int v) { return (FOLD && EQ == EQ_EULER && v >= 1 && v <= ND) ? 1 + pc[v - 1] : v; };
        // Riemann fluxes at the line's two ends (surface_contribution!): Fn is the master-outward flux
        // in the master's face-dof order; the slave side sees it negated (sgL / sgR) and permuted
        const double *fL, *fR;
        double sgL, sgR;
        auto face_src = [&]() {
            const int i_ = *it, fbuf = i_ & 1;
            mbar_wait(WsBars<C>::fullF(fbuf), (unsigned)((i_ >> 1) & 1));      // the flux blocks of this group have landed
            const int2 *ec = reinterpret_cast<const int2 *>(lsmem + C::OFF_EC) + fbuf * (C::E * NFACES);
            const double *sFn = lsmem + C::OFF_F + fbuf * (C::E * NFACES * FNB);
            const int2 ecL = ec[el * NFACES + 2 * d], ecR = ec[el * NFACES + 2 * d + 1];
            const int iL = (ecL.y & 1) ? k : slave2master<ND, NP>(k, (ecL.y >> 1) & 7);
            const int iR = (ecR.y & 1) ? k : slave2master<ND, NP>(k, (ecR.y >> 1) & 7);
            fL = sFn + (el * NFACES + 2 * d) * FNB + iL;
            fR = sFn + (el * NFACES + 2 * d + 1) * FNB + iR;
            sgL = (ecL.y & 1) ? 1.0 : -1.0;
            sgR = (ecR.y & 1) ? 1.0 : -1.0;
        };

        double acc[NP][NV];
#pragma unroll
        for (int j = 0; j < NP; j++)
#pragma unroll
            for (int v = 0; v < NV; v++) acc[j][v] = 0.0;

        if constexpr (C::HYBRID) {
            double Qj[NP][NV];
#pragma unroll
            for (int j = 0; j < NP; j++)
#pragma unroll
                for (int v = 0; v < NV; v++) Qj[j][v] = sA[v * N + base + j * stride];
            // global node of the line's first node and the line's slot in the sub-grid tables
            const int64_t gnode0 = dof0 + base;
            const int64_t sub0 = (((dof0 / NPTS + el) * ND + d) * NFP + k) * (NP + 1);
            if constexpr (C::NB) {
                // nodes without boundaries: the face fluxes enter the sub-cell recursion itself
                face_src();
                double fnl[NV], fnr[NV];
#pragma unroll
                for (int v = 0; v < NV; v++) { fnl[v] = sgL * fL[v * NFP]; fnr[v] = sgR * fR[v * NFP]; }
                if (d == 0) hybrid_nb_line<C, 0>(P, Qj, fnl, fnr, gnode0, stride, sub0, acc);
                else if (ND >= 2 && d == 1) hybrid_nb_line<C, (ND >= 2 ? 1 : 0)>(P, Qj, fnl, fnr, gnode0, stride, sub0, acc);
                else if (ND >= 3) hybrid_nb_line<C, (ND >= 3 ? 2 : 0)>(P, Qj, fnl, fnr, gnode0, stride, sub0, acc);
            } else {
            if (d == 0) hybrid_line<C, 0>(P, Qj, gnode0, stride, sub0, acc);
            else if (ND >= 2 && d == 1) hybrid_line<C, (ND >= 2 ? 1 : 0)>(P, Qj, gnode0, stride, sub0, acc);
            else if (ND >= 3) hybrid_line<C, (ND >= 3 ? 2 : 0)>(P, Qj, gnode0, stride, sub0, acc);
            }
        } else if (!SPLIT) {
            // strong form: dQ[line] -= Ds * F~[line, d]
#pragma unroll
            for (int v = 0; v < NV; v++) {
                double f[NP];
#pragma unroll
                for (int j = 0; j < NP; j++) f[j] = sA[(d * NV + v) * N + base + j * stride];
#pragma unroll
                for (int i = 0; i < NP; i++)
#pragma unroll
                    for (int j = 0; j < NP; j++) acc[i][v] = fma(-P.Dvol[i + NP * j], f[j], acc[i][v]);
            }
        } else {
            // split form: dQ_i -= sum_j D#[i,j] F#(i,j), F# symmetric (OpDivergence.jl:248-282)
            double mt[CART ? 1 : NP][CART ? 1 : ND];
            if (!CART) {
#pragma unroll
                for (int j = 0; j < NP; j++)
#pragma unroll
                    for (int c = 0; c < ND; c++)
                        mt[j][c] = __ldg(P.metric + dof0 + base + j * stride + ndof * (c + ND * d));
            }
            if (EQ == EQ_EULER && VOL == VOL_SPLIT_CHA) {
                double r[NP], hv[NP][ND], b[NP];
#pragma unroll
                for (int j = 0; j < NP; j++) {
                    const int node = base + j * stride;
                    r[j] = sA[node];
#pragma unroll
                    for (int c = 0; c < ND; c++) hv[j][c] = sA[(1 + pc[c]) * N + node];
                    b[j] = sA[(ND + 1) * N + node];
                }
                redo = cha_pairs<C, FAST>(P, r, hv, b, mt, acc);
            } else if (EQ == EQ_EULER) {
                // StdAverage two-point flux (Equations/Euler.jl:385-472); node data (Q, v, p)
                double Qj[NP][NV], vj[NP][ND], pj[NP];
#pragma unroll
                for (int j = 0; j < NP; j++) {
                    const int node = base + j * stride;
                    Qj[j][0] = sA[node];
#pragma unroll
                    for (int c = 0; c < ND; c++) {
                        Qj[j][1 + c] = sA[(1 + pc[c]) * N + node];
                        vj[j][c] = sA[(NV + pc[c]) * N + node];
                    }
                    Qj[j][ND + 1] = sA[(ND + 1) * N + node];
                    pj[j] = sA[(NV + ND) * N + node];
                }
#pragma unroll
                for (int j = 0; j < NP; j++)
#pragma unroll
                    for (int l = j; l < NP; l++) {
                        if (l == j && !(P.diag_mask & (1 << j))) continue;
                        double n[ND], F[NV];
#pragma unroll
                        for (int c = 0; c < ND; c++) n[c] = CART ? (c == 0 ? 1.0 : 0.0) : 0.5 * (mt[j][c] + mt[l][c]);
                        tp_stdavg<ND>(Qj[j], vj[j], pj[j], Qj[l], vj[l], pj[l], n, F);
                        const double djl = P.Dvol[j + NP * l], dlj = P.Dvol[l + NP * j];
#pragma unroll
                        for (int v = 0; v < NV; v++) {
                            acc[j][v] = fma(-djl, F[v], acc[j][v]);
                            if (l != j) acc[l][v] = fma(-dlj, F[v], acc[l][v]);
                        }
                    }
            } else {
                // linear advection (Equations/LinearAdvection.jl:44-47)
                double q[NP];
#pragma unroll
                for (int j = 0; j < NP; j++) q[j] = sA[base + j * stride];
                const double ad = pick<ND>(P.fp.a, d);
#pragma unroll
                for (int j = 0; j < NP; j++)
#pragma unroll
                    for (int l = j; l < NP; l++) {
                        if (l == j && !(P.diag_mask & (1 << j))) continue;
                        double an = ad;
                        if (!CART) {
                            an = 0.0;
#pragma unroll
                            for (int c = 0; c < ND; c++) an += P.fp.a[c] * 0.5 * (mt[j][c] + mt[l][c]);
                        }
                        const double F = an * (q[j] + q[l]) * 0.5;
                        acc[j][0] = fma(-P.Dvol[j + NP * l], F, acc[j][0]);
                        if (l != j) acc[l][0] = fma(-P.Dvol[l + NP * j], F, acc[l][0]);
                    }
            }
        }

        // one line per thread and iteration (task == thread index): the indices are derived again
        // from a fresh read of the thread index instead of being kept (spilled) across the pair fluxes
        if constexpr (C::ONE_ROUND && !C::HYBRID && !C::NB) {
            int t_;
            asm volatile("mov.u32 %0, %%tid.x;" : "=r"(t_));
            line_indices(t_);
        }
        // lift of the two face fluxes (OpDivergence.jl:42-100); in FOLD mode the partial sum is
        // later multiplied by the metric factor of direction d, so the lift is pre-divided by it
        if (!(C::HYBRID && C::NB)) face_src();
        {
            const double rm = FOLD ? pick<ND>(P.rcmet, d) : 1.0;
            const double wl = sgL * rm, wr = sgR * rm;
            if (P.colloc) {
                // GLL: only the two end nodes of the line see the face fluxes
                const double w0 = P.dgl[0] * wl, w1 = P.dgr[NP - 1] * wr;
#pragma unroll
                for (int v = 0; v < NV; v++) {
                    acc[0][v] = fma(-w0, fL[var_of(v) * NFP], acc[0][v]);
                    acc[NP - 1][v] = fma(-w1, fR[var_of(v) * NFP], acc[NP - 1][v]);
                }
            } else if (C::HYBRID) {
                // hybrid operator on nodes without boundaries: hybrid_nb_line consumed the face fluxes
            } else if (SPLIT && EQ == EQ_EULER) {
                // split form on nodes without boundaries (Gauss): the surface term couples every
                // node of the line with the entropy-projected end states
                // (_splitdiv_nb_surface_contribution!, OpDivergence.jl:300-437); NB instances only (the host selects them for such nodes); these
                // never take the fast Chandrasekhar path, which has no surface code
                if constexpr (C::NB && !FAST) {
                    const int64_t sub0 = (((dof0 / NPTS + el) * ND + d) * NFP + k) * (NP + 1);
                    double fnl[NV], fnr[NV];
#pragma unroll
                    for (int v = 0; v < NV; v++) { fnl[v] = fL[var_of(v) * NFP]; fnr[v] = fR[var_of(v) * NFP]; }
                    splitdiv_nb_line<C>(P, sA, fnl, fnr, base, stride, pc, wl, wr, d, dof0 + base, sub0, acc);
                }
            } else {
#pragma unroll
                for (int j = 0; j < NP; j++) {
                    const double w0 = P.dgl[j] * wl, w1 = P.dgr[j] * wr;
#pragma unroll
                    for (int v = 0; v < NV; v++) {
                        acc[j][v] = fma(-w0, fL[var_of(v) * NFP], acc[j][v]);
                        acc[j][v] = fma(-w1, fR[var_of(v) * NFP], acc[j][v]);
                    }
                }
            }
        }
        // partial sums of direction d, momentum components back in physical order (warp-specialised
        // kernel: once the update warp has consumed the partial sums of the previous group)
        {
            const int i_ = *it;
            if (i_ > 0) {
                mbar_wait(WsBars<C>::freeP(), (unsigned)((i_ - 1) & 1));
            }
        }
#pragma unroll
        for (int j = 0; j < NP; j++) {
            const int node = base + j * stride;
#pragma unroll
            for (int v = 0; v < NV; v++) pt[var_of(v) * N + node] = acc[j][v];
        }
    return redo;
}

// exact redo of a line, kept out of line so that its register allocation (library log, calls) does
// not weigh on the fast path
template <class C>
__device__ __noinline__ void line_task_exact(const KParams &P, int task, int64_t dof0, const volatile int *it)
{
    line_task<C, false>(P, task, dof0, it);
}

// Node data of the line phase from the conservative state of one node (phase 1):
// Chandrasekhar (rho, v/2, beta); StdAverage (Q, v, p); strong form: the ND contravariant fluxes
// F~_d = sum_c F_c metric[c,d] (OpDivergence.jl:28-37); advection: q (split) or a~_d q (strong).
template <class C>
__device__ __forceinline__ void node_data(const KParams &P, const double (&Q)[C::NV], int64_t dof, double (&ax)[C::NAUX])
{
    constexpr int ND = C::ND, EQ = C::EQ, VOL = C::VOL, NV = C::NV;
    constexpr bool CART = C::CART, SPLIT = C::SPLIT;
    double met[(CART || SPLIT) ? 1 : ND * ND];
    if constexpr (C::HYBRID) {
        NodeAux<ND> A;
        node_aux<ND>(Q, P.fp.gamma, A);
        if (!(Q[0] > 0.0) || !(A.p > 0.0)) atomicOr(P.status, 1);
#pragma unroll
        for (int v = 0; v < NV; v++) ax[v] = Q[v];
        return;
    }
    if (!CART && !SPLIT) {
#pragma unroll
        for (int m = 0; m < ND * ND; m++) met[m] = __ldg(P.metric + dof + P.ndof * m);
    }
    if (EQ == EQ_EULER) {
        NodeAux<ND> A;
        node_aux<ND>(Q, P.fp.gamma, A);
        if (!(Q[0] > 0.0) || !(A.p > 0.0)) atomicOr(P.status, 1);
        if (VOL == VOL_SPLIT_CHA) {
            ax[0] = Q[0];
#pragma unroll
            for (int c = 0; c < ND; c++) ax[1 + c] = 0.5 * A.vel[c];
            ax[ND + 1] = A.beta;
        } else if (VOL == VOL_SPLIT_STD) {
#pragma unroll
            for (int v = 0; v < NV; v++) ax[v] = Q[v];
#pragma unroll
            for (int c = 0; c < ND; c++) ax[NV + c] = A.vel[c];
            ax[NV + ND] = A.p;
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
                for (int v = 0; v < NV; v++) ax[d * NV + v] = Ft[v];
            }
        }
    } else if (SPLIT) {
        ax[0] = Q[0];
    } else {
#pragma unroll
        for (int d = 0; d < ND; d++) {
            double an = 0.0;
#pragma unroll
            for (int c = 0; c < ND; c++)
                an += P.fp.a[c] * (CART ? (c == d ? P.cmet[d] : 0.0) : met[c + ND * d]);
            ax[d] = an * Q[0];
        }
    }
}

// Phase 1 of a group: node data of every node, RU nodes per thread at a time.
template <class C, int RU, int T>
__device__ __forceinline__ void phase1_nodes(const KParams &P, const double *U, double *sA, int tid, int nn, int64_t dof0)
{
    constexpr int NV = C::NV, N = C::N, NAUX = C::NAUX;
    for (int n0 = tid; n0 < nn; n0 += RU * T) {
        double Q[RU][NV], ax[RU][NAUX];
#pragma unroll
        for (int u = 0; u < RU; u++) {
            const int n = min(n0 + u * T, nn - 1);
#pragma unroll
            for (int v = 0; v < NV; v++) Q[u][v] = U[v * N + n];
        }
#pragma unroll
        for (int u = 0; u < RU; u++) node_data<C>(P, Q[u], dof0 + min(n0 + u * T, nn - 1), ax[u]);
#pragma unroll
        for (int u = 0; u < RU; u++) {
            const int n = n0 + u * T;
            if (n < nn) {
#pragma unroll
                for (int c = 0; c < NAUX; c++) sA[c * N + n] = ax[u][c];
            }
        }
    }
}

// Phase 3 of a group: sum of the directional partial sums, mass matrix, RK stage update, traces.
// Same arithmetic, operation by operation, as phase3_pairs (which path a group takes depends on the
// alignment of its planes, i.e. on the partition: the results must not).
template <class C, int RU, int T>
__device__ __forceinline__ void phase3_nodes(const KParams &P, const double *U, const double *sT, const double *sP,
                                             int tid, int nn, int64_t dof0, int g)
{
    constexpr int ND = C::ND, NP = C::NP, NV = C::NV, NPTS = C::NPTS, NFP = C::NFP, N = C::N, E = C::E;
    constexpr bool CART = C::CART, FOLD = C::FOLD;
    const int64_t ndof = P.ndof;
    const bool need_tmp = (P.mode == MODE_STAGE);
    const double cs = (P.mode == MODE_RHS) ? 1.0 : P.dt;
    double w[ND];
#pragma unroll
    for (int d = 0; d < ND; d++) w[d] = (FOLD ? P.cmet[d] : 1.0) * (CART ? P.crjac : 1.0) * cs;
    for (int n0 = tid; n0 < nn; n0 += RU * T) {
        double acc[RU][NV], tv[RU][NV], uv[RU][NV];
#pragma unroll
        for (int u = 0; u < RU; u++) {
            const int n = min(n0 + u * T, nn - 1);
            const double rj = CART ? 1.0 : fast_rcp(__ldg(P.jac + dof0 + n));
#pragma unroll
            for (int v = 0; v < NV; v++) {
                double s = 0.0;
                const bool src = P.source != nullptr;
                if (src) s = cs * __ldg(P.source + dof0 + n + ndof * v);
                if (CART) {
#pragma unroll
                    for (int d = 0; d < ND; d++) {
                        const double x = sP[(d * NV + v) * N + n];
                        s = (!src && d == 0) ? w[0] * x : fma(w[d], x, s);
                    }
                } else {
                    double q = 0.0;
#pragma unroll
                    for (int d = 0; d < ND; d++) q += sP[(d * NV + v) * N + n];
                    s = fma(q, rj * cs, s);
                }
                acc[u][v] = s;
                tv[u][v] = need_tmp ? sT[v * N + n] : 0.0;
                uv[u][v] = U[v * N + n];
            }
        }
#pragma unroll
        for (int u = 0; u < RU; u++) {
            const int n = n0 + u * T;
            if (n >= nn) break;
            const int64_t dof = dof0 + n;
            if (P.mode == MODE_RHS) {
#pragma unroll
                for (int v = 0; v < NV; v++) P.k_out[dof + ndof * v] = acc[u][v];
            } else {
                double un[NV];
#pragma unroll
                for (int v = 0; v < NV; v++) {
                    const double t = need_tmp ? fma(P.rkA, tv[u][v], acc[u][v]) : acc[u][v];
                    P.tmp[dof + ndof * v] = t;
                    un[v] = fma(P.rkB, t, uv[u][v]);
                    P.u_out[dof + ndof * v] = un[v];
                }
                // x-face traces of the new state for the next stage (collocated nodes only;
                // Gauss nodes are handled by emit_traces_kernel; no trace array: the face kernel
                // reads the node layers of u)
                if (P.colloc && P.tr_out) {
                    const int el = n / NPTS, node = n - el * NPTS;
                    int k, ii;
                    node_line<ND, NP>(node, 0, k, ii);
                    if (ii == 0 || ii == NP - 1) {
                        const int64_t e = P.elem_first + g * E + el;
                        double *dst = P.tr_out + (e * 2 + (ii == 0 ? 0 : 1)) * (NV * NFP) + k;
#pragma unroll
                        for (int v = 0; v < NV; v++) dst[v * NFP] = un[v];
                    }
                }
            }
        }
    }
}

// Phase 3 on PAIRS of adjacent nodes with 128-bit shared loads and global stores (half the
// instructions of phase3_nodes): needs an even node count and 16-byte aligned planes (the `wide`
// condition of the copies).  RP pairs per thread at a time, loads first.
// STAGE_TR: the new state is also written back over the old one in shared memory and the x-face
// traces are emitted afterwards by trace_pass (warp-specialised kernel).
// BULK (with STAGE_TR): tmp and the new state are ONLY written to shared memory (tmp in place over
// sT, the state over U); the caller sends the planes to global memory with TMA bulk stores -- the
// update warp is the critical path of the kernel and its 128-bit global stores with their 64-bit
// address arithmetic were a quarter of its instructions.
// Arithmetic: with w_d = metric_d / J * dt folded into three scalars,
//   dt*k = sum_d w_d * partial_d (+ dt * source),  tmp = A*tmp + dt*k,  u = u + B*tmp
// (5 fp64 instructions per value; MODE_RHS uses dt = 1 and stores k).
// MODE: the pass mode as a compile-time constant (the caller branches once, warp-uniformly) or -1 =
// read P.mode; SRC: 0 = no source term, 1 = read P.source (may be null).  The specialised copies
// carry no selects / predicate logic for the other modes: the update warp is the critical path of
// the kernel and a third of its instructions were such bookkeeping (profiles/r2_kernel_notes.md).
template <class C, int RP, int T, bool STAGE_TR = false, bool BULK = false, int MODE = -1, int SRC = 1>
__device__ __forceinline__ void phase3_pairs(const KParams &P, std::conditional_t<STAGE_TR, double, const double> *U,
                                             std::conditional_t<BULK, double, const double> *sT, const double *sP,
                                             int tid, int nn, int64_t dof0, int g)
{
    constexpr int ND = C::ND, NP = C::NP, NV = C::NV, NPTS = C::NPTS, NFP = C::NFP, N = C::N, E = C::E;
    constexpr bool CART = C::CART, FOLD = C::FOLD;
    static_assert(!BULK || STAGE_TR, "bulk stores go with the in-place update of the shared-memory copy");
    const int64_t ndof = P.ndof;
    const int mode = MODE >= 0 ? MODE : P.mode;
    const bool need_tmp = (mode == MODE_STAGE);
    const double cs = (mode == MODE_RHS) ? 1.0 : P.dt;
    double w[ND];
#pragma unroll
    for (int d = 0; d < ND; d++) w[d] = (FOLD ? P.cmet[d] : 1.0) * (CART ? P.crjac : 1.0) * cs;
    const int npairs = nn >> 1;
    for (int p0 = tid; p0 < npairs; p0 += RP * T) {
        double2 acc[RP][NV], tv[RP][NV], uv[RP][NV];
#pragma unroll
        for (int r = 0; r < RP; r++) {
            const int n = 2 * min(p0 + r * T, npairs - 1);
            // apply_sourceterm! (MultielementDiscontinuous.jl:139-146): tabulated source, after the
            // mass matrix; a warp-uniform branch that the default (no source) never takes
            double2 rj = make_double2(1.0, 1.0);
            if (!CART) {
                const double2 j = __ldg(reinterpret_cast<const double2 *>(P.jac + dof0 + n));
                rj = make_double2(fast_rcp(j.x), fast_rcp(j.y));
            }
#pragma unroll
            for (int v = 0; v < NV; v++) {
                double2 s = make_double2(0.0, 0.0);
                if (SRC && P.source) {
                    const double2 sv = __ldg(reinterpret_cast<const double2 *>(P.source + dof0 + n + ndof * v));
                    s = make_double2(cs * sv.x, cs * sv.y);
                }
                if (CART) {
#pragma unroll
                    for (int d = 0; d < ND; d++) {
                        const double2 x = *reinterpret_cast<const double2 *>(sP + (d * NV + v) * N + n);
                        if (!SRC && d == 0) { s.x = w[0] * x.x; s.y = w[0] * x.y; }
                        else { s.x = fma(w[d], x.x, s.x); s.y = fma(w[d], x.y, s.y); }
                    }
                } else {
                    double2 q = make_double2(0.0, 0.0);
#pragma unroll
                    for (int d = 0; d < ND; d++) {
                        const double2 x = *reinterpret_cast<const double2 *>(sP + (d * NV + v) * N + n);
                        q.x += x.x; q.y += x.y;
                    }
                    s.x = fma(q.x, rj.x * cs, s.x); s.y = fma(q.y, rj.y * cs, s.y);
                }
                acc[r][v] = s;
                // U is updated in place below (STAGE_TR): a lane without a pair of its own must not read
                // the clamped pair, which another lane is about to overwrite (racecheck)
                const bool own = STAGE_TR ? (p0 + r * T < npairs) : true;
                if constexpr (MODE >= 0) {      // compile-time mode: only what the mode reads
                    if (MODE == MODE_STAGE) tv[r][v] = *reinterpret_cast<const double2 *>(sT + v * N + n);
                    if (MODE != MODE_RHS) uv[r][v] = own ? *reinterpret_cast<const double2 *>(U + v * N + n) : make_double2(0.0, 0.0);
                } else {
                    tv[r][v] = need_tmp ? *reinterpret_cast<const double2 *>(sT + v * N + n) : make_double2(0.0, 0.0);
                    uv[r][v] = own ? *reinterpret_cast<const double2 *>(U + v * N + n) : make_double2(0.0, 0.0);
                }
            }
        }
#pragma unroll
        for (int r = 0; r < RP; r++) {
            const int n = 2 * (p0 + r * T);
            if (n >= nn) break;
            const int64_t dof = dof0 + n;
            if (mode == MODE_RHS) {
#pragma unroll
                for (int v = 0; v < NV; v++) __stcs(reinterpret_cast<double2 *>(P.k_out + dof + ndof * v), acc[r][v]);
            } else {
                double2 un[NV];
#pragma unroll
                for (int v = 0; v < NV; v++) {
                    double2 t = acc[r][v];
                    if (need_tmp) { t.x = fma(P.rkA, tv[r][v].x, t.x); t.y = fma(P.rkA, tv[r][v].y, t.y); }
                    un[v].x = fma(P.rkB, t.x, uv[r][v].x);
                    un[v].y = fma(P.rkB, t.y, uv[r][v].y);
                    if constexpr (BULK) {
                        *reinterpret_cast<double2 *>(sT + v * N + n) = t;
                    } else {
                        __stcs(reinterpret_cast<double2 *>(P.tmp + dof + ndof * v), t);
                        __stcs(reinterpret_cast<double2 *>(P.u_out + dof + ndof * v), un[v]);
                    }
                    if constexpr (STAGE_TR) *reinterpret_cast<double2 *>(U + v * N + n) = un[v];
                }
                if (!STAGE_TR && P.colloc && P.tr_out) {
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        const int m = n + h;
                        const int el = m / NPTS, node = m - el * NPTS;
                        int k, ii;
                        node_line<ND, NP>(node, 0, k, ii);
                        if (ii == 0 || ii == NP - 1) {
                            const int64_t e = P.elem_first + g * E + el;
                            double *dst = P.tr_out + (e * 2 + (ii == 0 ? 0 : 1)) * (NV * NFP) + k;
#pragma unroll
                            for (int v = 0; v < NV; v++) dst[v * NFP] = h ? un[v].y : un[v].x;
                        }
                    }
                }
            }
        }
    }
}

// x-face traces of the new state of a group from its copy in shared memory: one entry per
// (element, side, face dof), coalesced over the face dof.
template <class C, int T>
__device__ __forceinline__ void trace_pass(const KParams &P, const double *Unew, int tid, int nact, int g)
{
    constexpr int NP = C::NP, NV = C::NV, NPTS = C::NPTS, NFP = C::NFP, N = C::N, E = C::E;
    for (int t = tid; t < nact * 2 * NFP; t += T) {
        const int el = t / (2 * NFP), r = t - el * (2 * NFP);
        const int side = r / NFP, k = r - side * NFP;
        const int n = el * NPTS + NP * k + (side ? NP - 1 : 0);       // line_of(0, k): base NP*k, stride 1
        const int64_t e = P.elem_first + g * E + el;
        double *dst = P.tr_out + (e * 2 + side) * (NV * NFP) + k;
#pragma unroll
        for (int v = 0; v < NV; v++) __stcs(dst + v * NFP, Unew[v * N + n]);
    }
}

}  // namespace flou
