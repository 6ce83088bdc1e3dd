// Element kernel of the two-kernel RK stage, LINE-PER-THREAD formulation.
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
// Phases of a CTA (E consecutive elements, T threads):
//   1. node tasks:  load u_in, node primitives -> shared memory (NAUX planes per element)
//   2. line tasks:  (element, direction, line) -> partial sums, one plane set per direction
//   3. node tasks:  sum the ND partial sums, 1/jac, RK update, x-face traces of u_out
#pragma once
#include "stage_kernel.cuh"

namespace flou {

struct ETPick { int e, t; };

// elements per CTA / threads per CTA: maximise the lane utilisation of the line phase
constexpr ETPick pick_et(int nlines, int per_elem_doubles)
{
    const int ts[7] = {128, 160, 96, 192, 64, 224, 256};
    ETPick best{1, 64};
    int best_util = 0;                                    // in 1/1000
    for (int q = 0; q < 7; q++) {
        const int t = ts[q];
        int e = t / nlines;
        if (e < 1) e = 1;
        const int emax = (48 * 1024 / 8) / per_elem_doubles;     // <= 48 KB of shared memory per CTA
        if (e > emax) e = emax < 1 ? 1 : emax;
        if (e > 64) e = 64;
        const int rounds = (e * nlines + t - 1) / t;
        const int util = (1000 * e * nlines) / (rounds * t);
        if (util > best_util + 20) { best = ETPick{e, t}; best_util = util; }
    }
    return best;
}

template <int ND_, int NP_, int EQ_, int VOL_, bool CART_>
struct LCfg {
    static constexpr int ND = ND_, NP = NP_, EQ = EQ_, VOL = VOL_;
    static constexpr bool CART = CART_;
    static constexpr int NV = (EQ == EQ_ADV) ? 1 : ND + 2;
    static constexpr int NPTS = ipow_c(NP, ND);
    static constexpr int NFP = ipow_c(NP, ND - 1);
    static constexpr int NFACES = 2 * ND;
    static constexpr int NLINES = ND * NFP;
    static constexpr bool SPLIT = (VOL != VOL_STRONG);
    // Cartesian split form: the constant metric factor is applied when the partial sums are added
    static constexpr bool FOLD = CART && SPLIT;
    // node data in shared memory: Chandrasekhar (rho, v/2, beta); StdAverage (Q, v, p);
    // strong form: the ND contravariant fluxes; advection: q
    static constexpr int NAUX = (EQ == EQ_ADV) ? (SPLIT ? 1 : ND)
                              : (VOL == VOL_SPLIT_CHA ? ND + 2 : (VOL == VOL_SPLIT_STD ? NV + ND + 1 : ND * NV));
    static constexpr int NPART = ND * NV;
    static constexpr int PER_ELEM = (NAUX + NPART) * NPTS;
#if defined(FLOU_LINE_E) && defined(FLOU_LINE_T)
    static constexpr int E = FLOU_LINE_E, T = FLOU_LINE_T;
#else
    static constexpr int E = pick_et(NLINES, PER_ELEM).e, T = pick_et(NLINES, PER_ELEM).t;
#endif
    static constexpr size_t SMEM_BYTES = sizeof(double) * (size_t)E * PER_ELEM;
    // registers: a line task holds NP nodes and NP*NV accumulators
    static constexpr int MINB =
#ifdef FLOU_LINE_MINB
        FLOU_LINE_MINB;
#else
        (NP * (NAUX + NV) > 64) ? 1 : 2;
#endif
};

// ------------------------------------------------------------------ two-point fluxes of a line
// Chandrasekhar two-point flux (Equations/Euler.jl:474-536) along the (permuted) axis 0 with a
// unit metric; node data (rho, v/2, beta).  -|v1|^2/4 - |v2|^2/4 + |v_avg|^2 = 2 (v1/2).(v2/2).
template <int ND>
__device__ __forceinline__ void tp_cha_axis(double r1, const double *hv1, double b1,
                                            double r2, const double *hv2, double b2,
                                            double inv_gm1, double *F)
{
    const double rs = r1 + r2, bs = b1 + b2;
    const double irb = fast_rcp(rs * bs);
    const double irs = irb * bs, ibs = irb * rs;
    double Fr, Fb;
    logmean_F2(r1, r2, irs, b1, b2, ibs, Fr, Fb);
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
template <int ND>
__device__ __forceinline__ void tp_cha_n(double r1, const double *hv1, double b1,
                                         double r2, const double *hv2, double b2,
                                         double inv_gm1, const double *n, double *F)
{
    const double rs = r1 + r2, bs = b1 + b2;
    const double irb = fast_rcp(rs * bs);
    const double irs = irb * bs, ibs = irb * rs;
    double Fr, Fb;
    logmean_F2(r1, r2, irs, b1, b2, ibs, Fr, Fb);
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

// ------------------------------------------------------------------ the kernel
template <class C>
__global__ void __launch_bounds__(C::T, C::MINB)
line_kernel(const __grid_constant__ KParams P)
{
    constexpr int ND = C::ND, NP = C::NP, EQ = C::EQ, VOL = C::VOL, NV = C::NV;
    constexpr int NPTS = C::NPTS, NFP = C::NFP, NFACES = C::NFACES, NLINES = C::NLINES;
    constexpr int E = C::E, T = C::T, NAUX = C::NAUX;
    constexpr bool CART = C::CART, SPLIT = C::SPLIT, FOLD = C::FOLD;

    extern __shared__ double smem[];
    const int tid = threadIdx.x;
    const int g = blockIdx.x;
    const int nact = min(E, P.elem_count - g * E);
    const int64_t ndof = P.ndof;
    auto elem_of = [&](int idx) { return P.elem_list ? P.elem_list[idx] : P.elem_first + idx; };

    // The two face fluxes at the ends of a line (surface_contribution!): Fn is the master-outward
    // flux in the master's face-dof order.  With one line task per thread the connectivity is
    // requested before phase 1 and the fluxes before the barrier, so both latencies hide behind
    // phase 1 and the volume work; they are consumed last.
    constexpr bool ONE_ROUND = (E * NLINES <= T);
    auto get_ec = [&](int task, int2 &ecL, int2 &ecR) {
        const int el = task / NLINES, r_ = task - el * NLINES;
        const int d = r_ / NFP;
        const int e = elem_of(g * E + el);
        ecL = __ldg(P.econn + ((int64_t)e * NFACES + 2 * d));
        ecR = __ldg(P.econn + ((int64_t)e * NFACES + 2 * d + 1));
    };
    auto get_fn = [&](int task, int2 ecL, int2 ecR, double (&FnL)[NV], double (&FnR)[NV], double &sgL, double &sgR) {
        const int r_ = task % NLINES;
        const int d = r_ / NFP, k = r_ - d * NFP;
        const int iL = (ecL.y & 1) ? k : slave2master<ND, NP>(k, (ecL.y >> 1) & 7);
        const int iR = (ecR.y & 1) ? k : slave2master<ND, NP>(k, (ecR.y >> 1) & 7);
        const double *sL = P.Fn + (int64_t)ecL.x * (NV * NFP) + iL;
        const double *sR = P.Fn + (int64_t)ecR.x * (NV * NFP) + iR;
#pragma unroll
        for (int v = 0; v < NV; v++) {
            // FOLD mode handles the momentum components in cyclic order starting at d
            int vv = v;
            if (FOLD && EQ == EQ_EULER && v >= 1 && v <= ND) { const int s_ = d + v - 1; vv = 1 + (s_ >= ND ? s_ - ND : s_); }
            FnL[v] = __ldg(sL + vv * NFP);
            FnR[v] = __ldg(sR + vv * NFP);
        }
        sgL = (ecL.y & 1) ? 1.0 : -1.0;
        sgR = (ecR.y & 1) ? 1.0 : -1.0;
    };
    const bool has_task = ONE_ROUND && tid < nact * NLINES;
    int2 ec0L = make_int2(0, 0), ec0R = make_int2(0, 0);
    double Fn0L[NV], Fn0R[NV], sg0L = 1.0, sg0R = 1.0;
    if (has_task) get_ec(tid, ec0L, ec0R);

    // ---------------- phase 1: node primitives -> shared memory
    for (int n = tid; n < nact * NPTS; n += T) {
        const int el = n / NPTS, node = n - el * NPTS;
        const int e = elem_of(g * E + el);
        const int64_t dof = (int64_t)e * NPTS + node;
        double *ax = smem + (size_t)el * C::PER_ELEM;
        double Q[NV];
#pragma unroll
        for (int v = 0; v < NV; v++) Q[v] = __ldg(P.u_in + dof + ndof * v);
        double met[(CART || SPLIT) ? 1 : ND * ND];
        if (!CART && !SPLIT) {
#pragma unroll
            for (int m = 0; m < ND * ND; m++) met[m] = __ldg(P.metric + dof + ndof * m);
        }
        if (EQ == EQ_EULER) {
            NodeAux<ND> A;
            node_aux<ND>(Q, P.fp.gamma, A);
            if (!(Q[0] > 0.0) || !(A.p > 0.0)) atomicOr(P.status, 1);
            if (VOL == VOL_SPLIT_CHA) {
                ax[node] = Q[0];
#pragma unroll
                for (int c = 0; c < ND; c++) ax[(1 + c) * NPTS + node] = 0.5 * A.vel[c];
                ax[(ND + 1) * NPTS + node] = A.beta;
            } else if (VOL == VOL_SPLIT_STD) {
#pragma unroll
                for (int v = 0; v < NV; v++) ax[v * NPTS + node] = Q[v];
#pragma unroll
                for (int c = 0; c < ND; c++) ax[(NV + c) * NPTS + node] = A.vel[c];
                ax[(NV + ND) * NPTS + node] = A.p;
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
                    for (int v = 0; v < NV; v++) ax[(d * NV + v) * NPTS + node] = Ft[v];
                }
            }
        } else if (SPLIT) {
            ax[node] = Q[0];
        } else {
#pragma unroll
            for (int d = 0; d < ND; d++) {
                double an = 0.0;
#pragma unroll
                for (int c = 0; c < ND; c++)
                    an += P.fp.a[c] * (CART ? (c == d ? P.cmet[d] : 0.0) : met[c + ND * d]);
                ax[d * NPTS + node] = an * Q[0];
            }
        }
        // warm L2 for the CTA that occupies this SM slot one wave later
        if (P.prefetch_groups > 0 && (node & 15) == 0) {
            const int idx = (g + P.prefetch_groups) * E + el;
            if (idx < P.elem_count) {
                const int64_t pd = (int64_t)elem_of(idx) * NPTS + node;
#pragma unroll
                for (int v = 0; v < NV; v++) {
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(P.u_in + pd + ndof * v));
                    if (P.mode == MODE_STAGE) asm volatile("prefetch.global.L2 [%0];" ::"l"(P.tmp + pd + ndof * v));
                }
            }
        }
    }
    if (has_task) get_fn(tid, ec0L, ec0R, Fn0L, Fn0R, sg0L, sg0R);
    __syncthreads();

    // ---------------- phase 2: one tensor-product line per thread
    auto line_task = [&](int task, const double (&FnL)[NV], const double (&FnR)[NV], double sgL, double sgR) {
        const int el = task / NLINES, r_ = task - el * NLINES;
        const int d = r_ / NFP, k = r_ - d * NFP;
        const int e = elem_of(g * E + el);
        int base, stride;
        line_of<ND, NP>(d, k, base, stride);
        const double *ax = smem + (size_t)el * C::PER_ELEM;
        double *pt = smem + (size_t)el * C::PER_ELEM + (NAUX + d * NV) * NPTS;

        // Cartesian split form: momentum components are handled in the cyclic order that puts the
        // line's direction first, so the flux code is the same instruction stream for every d
        int pc[ND];
#pragma unroll
        for (int c = 0; c < ND; c++) { const int s = d + c; pc[c] = FOLD ? (s >= ND ? s - ND : s) : c; }
        auto var_of = [&](int v) { return (FOLD && EQ == EQ_EULER && v >= 1 && v <= ND) ? 1 + pc[v - 1] : v; };

        double acc[NP][NV];
#pragma unroll
        for (int j = 0; j < NP; j++)
#pragma unroll
            for (int v = 0; v < NV; v++) acc[j][v] = 0.0;

        if (!SPLIT) {
            // strong form: dQ[line] -= Ds * F~[line, d]
#pragma unroll
            for (int v = 0; v < NV; v++) {
                double f[NP];
#pragma unroll
                for (int j = 0; j < NP; j++) f[j] = ax[(d * NV + v) * NPTS + base + j * stride];
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
                        mt[j][c] = __ldg(P.metric + (int64_t)e * NPTS + base + j * stride + ndof * (c + ND * d));
            }
            if (EQ == EQ_EULER && VOL == VOL_SPLIT_CHA) {
                double r[NP], hv[NP][ND], b[NP];
#pragma unroll
                for (int j = 0; j < NP; j++) {
                    const int node = base + j * stride;
                    r[j] = ax[node];
#pragma unroll
                    for (int c = 0; c < ND; c++) hv[j][c] = ax[(1 + pc[c]) * NPTS + node];
                    b[j] = ax[(ND + 1) * NPTS + node];
                }
#pragma unroll
                for (int j = 0; j < NP; j++) {
                    if (P.diag_mask & (1 << j)) {
                        double n[ND], F[NV];
#pragma unroll
                        for (int c = 0; c < ND; c++) n[c] = CART ? (c == 0 ? 1.0 : 0.0) : mt[j][c];
                        diag_cha_n<ND>(r[j], hv[j], b[j], P.fp.inv_gm1, n, F);
                        const double djj = P.Dvol[j + NP * j];
#pragma unroll
                        for (int v = 0; v < NV; v++) acc[j][v] = fma(-djj, F[v], acc[j][v]);
                    }
#pragma unroll
                    for (int l = j + 1; l < NP; l++) {
                        double F[NV];
                        if (CART) {
                            tp_cha_axis<ND>(r[j], hv[j], b[j], r[l], hv[l], b[l], P.fp.inv_gm1, F);
                        } else {
                            double n[ND];
#pragma unroll
                            for (int c = 0; c < ND; c++) n[c] = 0.5 * (mt[j][c] + mt[l][c]);
                            tp_cha_n<ND>(r[j], hv[j], b[j], r[l], hv[l], b[l], P.fp.inv_gm1, n, F);
                        }
                        const double djl = P.Dvol[j + NP * l], dlj = P.Dvol[l + NP * j];
#pragma unroll
                        for (int v = 0; v < NV; v++) {
                            acc[j][v] = fma(-djl, F[v], acc[j][v]);
                            acc[l][v] = fma(-dlj, F[v], acc[l][v]);
                        }
                    }
                }
            } else if (EQ == EQ_EULER) {
                // StdAverage two-point flux (Equations/Euler.jl:385-472); node data (Q, v, p)
                double Qj[NP][NV], vj[NP][ND], pj[NP];
#pragma unroll
                for (int j = 0; j < NP; j++) {
                    const int node = base + j * stride;
                    Qj[j][0] = ax[node];
#pragma unroll
                    for (int c = 0; c < ND; c++) {
                        Qj[j][1 + c] = ax[(1 + pc[c]) * NPTS + node];
                        vj[j][c] = ax[(NV + pc[c]) * NPTS + node];
                    }
                    Qj[j][ND + 1] = ax[(ND + 1) * NPTS + node];
                    pj[j] = ax[(NV + ND) * NPTS + node];
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
                for (int j = 0; j < NP; j++) q[j] = ax[base + j * stride];
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

        // lift of the two face fluxes (OpDivergence.jl:42-100); in FOLD mode the partial sum is
        // later multiplied by the metric factor of direction d, so the lift is pre-divided by it
        {
            const double rm = FOLD ? pick<ND>(P.rcmet, d) : 1.0;
            const double wl = sgL * rm, wr = sgR * rm;
#pragma unroll
            for (int j = 0; j < NP; j++) {
                if (P.dgl[j] != 0.0) {
                    const double w = P.dgl[j] * wl;
#pragma unroll
                    for (int v = 0; v < NV; v++) acc[j][v] = fma(-w, FnL[v], acc[j][v]);
                }
                if (P.dgr[j] != 0.0) {
                    const double w = P.dgr[j] * wr;
#pragma unroll
                    for (int v = 0; v < NV; v++) acc[j][v] = fma(-w, FnR[v], acc[j][v]);
                }
            }
        }
        // partial sums of direction d, momentum components back in physical order
#pragma unroll
        for (int j = 0; j < NP; j++) {
            const int node = base + j * stride;
#pragma unroll
            for (int v = 0; v < NV; v++) pt[var_of(v) * NPTS + node] = acc[j][v];
        }
    };
    if (ONE_ROUND) {
        if (has_task) line_task(tid, Fn0L, Fn0R, sg0L, sg0R);
    } else {
        for (int task = tid; task < nact * NLINES; task += T) {
            int2 ecL, ecR;
            double FnL[NV], FnR[NV], sgL, sgR;
            get_ec(task, ecL, ecR);
            get_fn(task, ecL, ecR, FnL, FnR, sgL, sgR);
            line_task(task, FnL, FnR, sgL, sgR);
        }
    }
    __syncthreads();

    // ---------------- phase 3: sum the directions, mass matrix, RK stage update
    for (int n = tid; n < nact * NPTS; n += T) {
        const int el = n / NPTS, node = n - el * NPTS;
        const int e = elem_of(g * E + el);
        const int64_t dof = (int64_t)e * NPTS + node;
        const double *pt = smem + (size_t)el * C::PER_ELEM + NAUX * NPTS;
        double acc[NV];
#pragma unroll
        for (int v = 0; v < NV; v++) {
            double s = 0.0;
#pragma unroll
            for (int d = 0; d < ND; d++) {
                const double x = pt[(d * NV + v) * NPTS + node];
                s = FOLD ? fma(P.cmet[d], x, s) : s + x;
            }
            acc[v] = s;
        }
        const double rjac = CART ? P.crjac : fast_rcp(__ldg(P.jac + dof));
        if (P.mode == MODE_RHS) {
#pragma unroll
            for (int v = 0; v < NV; v++) P.k_out[dof + ndof * v] = acc[v] * rjac;
        } else {
            // every load before the first store: tmp and u_out may alias as far as the compiler
            // knows, and five serialised DRAM round trips were 32 % of the kernel's stall samples
            double un[NV], tv[NV];
#pragma unroll
            for (int v = 0; v < NV; v++) {
                un[v] = __ldg(P.u_in + dof + ndof * v);
                tv[v] = (P.mode == MODE_STAGE_FIRST) ? 0.0 : __ldcs(P.tmp + dof + ndof * v);
            }
#pragma unroll
            for (int v = 0; v < NV; v++) {
                const double kv = acc[v] * rjac;
                const double t = (P.mode == MODE_STAGE_FIRST) ? P.dt * kv : fma(P.dt, kv, P.rkA * tv[v]);
                un[v] = fma(P.rkB, t, un[v]);
                tv[v] = t;
            }
#pragma unroll
            for (int v = 0; v < NV; v++) {
                P.tmp[dof + ndof * v] = tv[v];
                P.u_out[dof + ndof * v] = un[v];
            }
            // x-face traces of the new state for the next stage (collocated nodes only; Gauss
            // nodes are handled by emit_traces_kernel)
            if (P.colloc) {
                int k, ii;
                node_line<ND, NP>(node, 0, k, ii);
                if (ii == 0 || ii == NP - 1) {
                    double *dst = P.tr_out + ((int64_t)e * 2 + (ii == 0 ? 0 : 1)) * (NV * NFP) + k;
#pragma unroll
                    for (int v = 0; v < NV; v++) dst[v * NFP] = un[v];
                }
            }
        }
    }
}

}  // namespace flou
