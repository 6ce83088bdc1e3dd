// Pointwise physics and numerical fluxes on the device (fp64).
//
// What each routine computes is fixed by Flou.jl (paths relative to the reference root):
//   logarithmic_mean                       src/FlouCommon/Utilities.jl:34-44
//   pressure / cons2prim / entropy vars    src/FlouCommon/Euler.jl:142-155, 237-253, 273-307
//   volumeflux                             src/FlouCommon/Euler.jl:54-114, LinearAdvection.jl:42-44
//   rotate2face / rotate2phys              src/FlouSpatial/Equations/Euler.jl:16-64
//   numericalflux StdAverage/LxF/Chandrasekhar/ScalarDissipation/MatrixDissipation
//                                          src/FlouSpatial/Equations/Euler.jl:99-380
//   twopointflux StdAverage/Chandrasekhar  src/FlouSpatial/Equations/Euler.jl:385-536
//   linear advection fluxes                src/FlouSpatial/Equations/LinearAdvection.jl:27-47
// How it is computed is ours: per-node primitives are computed once and reused for every
// node pair, divisions are shared, and the Roe-type dissipation R|L|T R'(Wl-Wr) is applied
// as a sum over eigenvectors instead of forming 5x5 products.  These rewrites change
// results at round-off level only (parity bar: 1e-12 relative on the RHS).
// Reference quirks kept on purpose: `wl^2` twice in the 3-D
// ChandrasekharAverage *numerical* flux (Euler.jl:216) and the left momenta taken from Qr
// in 3-D ScalarDissipation (Euler.jl:257).
#pragma once
#include "flou_b200.h"

namespace flou {

enum : int { EQ_ADV = FLOU_B200_EQ_LINEAR_ADVECTION, EQ_EULER = FLOU_B200_EQ_EULER };
// VOL_HYBRID: HybridDivOperator (telescopic split form blended with sub-cell finite volumes,
// OpDivergence.jl:452-612); line kernel only, two-point flux selected at run time (KParams::tpflux)
enum : int { VOL_STRONG = 0, VOL_SPLIT_STD = 1, VOL_SPLIT_CHA = 2, VOL_HYBRID = 3 };
enum : int { FX_STD = FLOU_B200_FLUX_STDAVERAGE, FX_LXF = FLOU_B200_FLUX_LXF,
             FX_CHA = FLOU_B200_FLUX_CHANDRASEKHAR, FX_SCA = FLOU_B200_FLUX_SCALARDISSIPATION,
             FX_MAT = FLOU_B200_FLUX_MATRIXDISSIPATION };

// 1/x for normal, finite x (states are O(1)): MUFU.RCP64H seed + one cubic (Halley) step, no
// special-case slow path.  Relative error ~1e-16 (checked in tests/test_parity_gpu.py).
__device__ __forceinline__ double fast_rcp(double x)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    // third-order step: r (1 + e + e^2), e = 1 - x r; a 2^-20 seed gives 2^-60
    const double e = fma(-x, r, 1.0);
    const double t = fma(e, e, e);
    return fma(r, t, r);
}

// rare branch of the logarithmic mean (|f| >= 0.1, i.e. a jump ratio above ~1.22): kept out
// of line so its log/division code is not replicated at every flux site
static __device__ __noinline__ double logmean_F_slow(double al, double ar, double f)
{
    return log(al / ar) / (2.0 * f);
}

// F = log(xi)/(2f) with xi = al/ar and f = (al-ar)/(al+ar), which is algebraically the
// reference's (xi-1)/(xi+1); same u < 0.01 threshold and 3-term series (Utilities.jl:34-44).
__device__ __forceinline__ double logmean_F(double al, double ar, double rsum)
{
    const double f = (al - ar) * rsum;
    const double u = f * f;
    double F = 1.0 + u * (1.0 / 3.0 + u * (1.0 / 5.0 + u * (1.0 / 7.0)));
    if (u >= 0.01) F = logmean_F_slow(al, ar, f);
    return F;
}

static // Both log-mean factors of a flux with ONE branch on the common (smooth) path.
__device__ __forceinline__ void logmean_F2(double a1, double a2, double ia, double b1, double b2,
                                           double ib, double &Fa, double &Fb)
{
    const double fa = (a1 - a2) * ia, fb = (b1 - b2) * ib;
    const double ua = fa * fa, ub = fb * fb;
    Fa = 1.0 + ua * (1.0 / 3.0 + ua * (1.0 / 5.0 + ua * (1.0 / 7.0)));
    Fb = 1.0 + ub * (1.0 / 3.0 + ub * (1.0 / 5.0 + ub * (1.0 / 7.0)));
    if (ua >= 0.01 || ub >= 0.01) {
        if (ua >= 0.01) Fa = logmean_F_slow(a1, a2, fa);
        if (ub >= 0.01) Fb = logmean_F_slow(b1, b2, fb);
    }
}

// Branch-free variant for the line kernel: series only.  `umax` accumulates (as an integer
// maximum of the high words, exact for non-negative doubles and one ALU instruction per value
// instead of a DSETP/FSEL/SEL triple) the largest u = f^2 met so far; series_out_of_range(umax)
// says whether some argument pair lies outside the series range, in which case the caller redoes
// its work with logmean_F2.  The high-word test is conservative by < 2^-20 relative: values that
// close below the threshold take the exact path, which applies the reference's own test.
__device__ __forceinline__ void logmean_F2_series(double a1, double a2, double ia, double b1, double b2,
                                                  double ib, double &Fa, double &Fb, int &umax)
{
    const double fa = (a1 - a2) * ia, fb = (b1 - b2) * ib;
    const double ua = fa * fa, ub = fb * fb;
    Fa = 1.0 + ua * (1.0 / 3.0 + ua * (1.0 / 5.0 + ua * (1.0 / 7.0)));
    Fb = 1.0 + ub * (1.0 / 3.0 + ub * (1.0 / 5.0 + ub * (1.0 / 7.0)));
#ifdef FLOU_REDO_FP      // A/B: the earlier floating-point test (DSETP/FSEL/SEL per value)
    if ((ua >= 0.01) | (ub >= 0.01)) umax = 0x7FFFFFFF;
#else
    umax = max(umax, max(__double2hiint(ua), __double2hiint(ub)));
#endif
}
__device__ __forceinline__ bool series_out_of_range(int umax)
{
    return umax >= 0x3F847AE1;      // high word of 0.01 = 0x3F847AE147AE147B
}

static __device__ __noinline__ double log_ratio_slow(double al, double ar) { return log(al / ar); }

// log(al/ar) = 2 atanh(f) from f = (al-ar)/(al+ar): 8-term series below u = f^2 < 0.01
// (truncation < 1e-17 relative), library log otherwise.
__device__ __forceinline__ double log_ratio(double al, double ar, double rsum)
{
    const double f = (al - ar) * rsum;
    const double u = f * f;
    double G = 1.0 + u * (1.0 / 3.0 + u * (1.0 / 5.0 + u * (1.0 / 7.0 + u * (1.0 / 9.0
             + u * (1.0 / 11.0 + u * (1.0 / 13.0 + u * (1.0 / 15.0)))))));
    double r = 2.0 * f * G;
    if (u >= 0.01) r = log_ratio_slow(al, ar);
    return r;
}

template <int ND>
__device__ __forceinline__ double pressure(const double *Q, double gamma)
{
    double m2 = Q[1] * Q[1];
#pragma unroll
    for (int d = 1; d < ND; d++) m2 += Q[1 + d] * Q[1 + d];
    return (gamma - 1.0) * (Q[ND + 1] - m2 / (2.0 * Q[0]));
}

// Node primitives: vel[ND], p, beta = rho/(2p).
template <int ND>
struct NodeAux {
    double vel[ND];
    double p;
    double beta;
};

template <int ND>
__device__ __forceinline__ void node_aux(const double *Q, double gamma, NodeAux<ND> &A)
{
    const double ir = fast_rcp(Q[0]);
    double m2 = 0.0;
#pragma unroll
    for (int d = 0; d < ND; d++) {
        A.vel[d] = Q[1 + d] * ir;
        m2 += Q[1 + d] * Q[1 + d];
    }
    A.p = (gamma - 1.0) * (Q[ND + 1] - m2 * (0.5 * ir));
    A.beta = 0.5 * Q[0] * fast_rcp(A.p);
}

// Physical flux in direction c of a node (Euler): F_c = (m_c, m_c*vel + p e_c, (E+p) vel_c)
template <int ND>
__device__ __forceinline__ void euler_flux_dir(const double *Q, const double *vel, double p,
                                               int c, double *F)
{
    const double mc = Q[1 + c];
    F[0] = mc;
#pragma unroll
    for (int d = 0; d < ND; d++) F[1 + d] = mc * vel[d] + (d == c ? p : 0.0);
    F[ND + 1] = (Q[ND + 1] + p) * vel[c];
}

// Two-point Chandrasekhar flux contracted with the averaged metric vector n[ND].
// Node data: rho, half velocities hv = v/2, q = |v|^2, beta = rho/(2p); inv_gm1 = 1/(gamma-1).
// Identities used (beta_ln = bs/(2 Fb), rho_ln = rs/(2 Fr), p_hat = rs/(2 bs)):
//   1/(2 beta_ln (g-1)) = Fb/(bs (g-1)),      p_hat/rho_ln = Fr/bs.
template <int ND>
__device__ __forceinline__ void tp_chandrasekhar(double r1, const double *hv1, double q1, double b1,
                                                 double r2, const double *hv2, double q2, double b2,
                                                 double inv_gm1, const double *n, double *F)
{
    const double rs = r1 + r2, bs = b1 + b2;
    const double irb = fast_rcp(rs * bs);        // one reciprocal for both sums
    const double irs = irb * bs, ibs = irb * rs;
    double Fr, Fb;
    logmean_F2(r1, r2, irs, b1, b2, ibs, Fr, Fb);
    const double rho = 0.5 * rs * fast_rcp(Fr);  // logarithmic_mean(rho1, rho2)
    const double p = rs * ibs * 0.5;             // (rho1+rho2)/(2(beta1+beta2))
    double vavg[ND], qa = 0.0, vn = 0.0;
#pragma unroll
    for (int d = 0; d < ND; d++) {
        vavg[d] = hv1[d] + hv2[d];
        qa = fma(vavg[d], vavg[d], qa);
        vn = fma(vavg[d], n[d], vn);
    }
    const double h = fma(fma(Fb, inv_gm1, Fr), ibs, fma(-0.25, q1 + q2, qa));
    const double mdot = rho * vn;                 // rho * (v . n)
    F[0] = mdot;
#pragma unroll
    for (int d = 0; d < ND; d++) F[1 + d] = fma(mdot, vavg[d], p * n[d]);
    F[ND + 1] = mdot * h;
}

// Two-point StdAverage flux contracted with n (Euler.jl:385-472).
template <int ND>
__device__ __forceinline__ void tp_stdavg(const double *Q1, const double *v1, double p1,
                                          const double *Q2, const double *v2, double p2,
                                          const double *n, double *F)
{
    double f0 = 0.0, fe = 0.0, fm[ND];
#pragma unroll
    for (int d = 0; d < ND; d++) fm[d] = 0.0;
#pragma unroll
    for (int c = 0; c < ND; c++) {
        f0 += 0.5 * (Q1[1 + c] + Q2[1 + c]) * n[c];
#pragma unroll
        for (int d = 0; d < ND; d++) {
            double t = Q1[1 + d] * v1[c] + Q2[1 + d] * v2[c];
            if (d == c) t += p1 + p2;
            fm[d] += 0.5 * t * n[c];
        }
        fe += 0.5 * ((Q1[ND + 1] + p1) * v1[c] + (Q2[ND + 1] + p2) * v2[c]) * n[c];
    }
    F[0] = f0;
#pragma unroll
    for (int d = 0; d < ND; d++) F[1 + d] = fm[d];
    F[ND + 1] = fe;
}

// ------------------------------------------------------------------ surface fluxes
struct FluxParams {
    double gamma, intensity;
    double gm1, inv_gm1;      // gamma-1, 1/(gamma-1)
    double inv_gamma;         // 1/gamma
    double a[3];
    int numflux, numflux_avg;
};

// Rotated-frame primitives of one side of a face
template <int ND>
struct SidePrim {
    double rho, vel[ND], p;
};

template <int ND>
__device__ __forceinline__ void side_prim(const double *Q, double gamma, SidePrim<ND> &S)
{
    const double ir = fast_rcp(Q[0]);
    double m2 = 0.0;
    S.rho = Q[0];
#pragma unroll
    for (int d = 0; d < ND; d++) {
        S.vel[d] = Q[1 + d] * ir;
        m2 += Q[1 + d] * Q[1 + d];
    }
    S.p = (gamma - 1.0) * (Q[ND + 1] - m2 * (0.5 * ir));
}

template <int ND>
__device__ __forceinline__ void nf_stdavg(const double *Ql, const double *Qr,
                                          const SidePrim<ND> &L, const SidePrim<ND> &R,
                                          double *F)
{
    const double ul = L.vel[0], ur = R.vel[0];
    F[0] = 0.5 * (Ql[1] + Qr[1]);
    F[1] = 0.5 * (Ql[1] * ul + L.p + Qr[1] * ur + R.p);
#pragma unroll
    for (int d = 1; d < ND; d++) F[1 + d] = 0.5 * (Ql[1 + d] * ul + Qr[1 + d] * ur);
    F[ND + 1] = 0.5 * ((Ql[ND + 1] + L.p) * ul + (Qr[ND + 1] + R.p) * ur);
}

// Chandrasekhar averages shared by the EC flux and the dissipation operators
template <int ND>
struct ChaAvg {
    double rho, p, beta_inv2;   // rho_ln, p_hat, 1/(2 beta_ln)
    double p_over_rho;          // p_hat / rho_ln
    double irs;                 // 1/(rho_l + rho_r)
    double v[ND];
};

template <int ND>
__device__ __forceinline__ void cha_avg(const SidePrim<ND> &L, const SidePrim<ND> &R,
                                        double bl, double br, ChaAvg<ND> &A)
{
    const double rs = L.rho + R.rho, bs = bl + br;
    const double irb = fast_rcp(rs * bs);
    const double irs = irb * bs, ibs = irb * rs;
    double Fr, Fb;
    logmean_F2(L.rho, R.rho, irs, bl, br, ibs, Fr, Fb);
    A.irs = irs;
    A.rho = 0.5 * rs * fast_rcp(Fr);
    A.p = rs * ibs * 0.5;
    A.beta_inv2 = Fb * ibs;          // 1/(2 beta_ln) = Fb/(bl+br)
    A.p_over_rho = Fr * ibs;
#pragma unroll
    for (int d = 0; d < ND; d++) A.v[d] = 0.5 * (L.vel[d] + R.vel[d]);
}

template <int ND>
__device__ __forceinline__ void nf_chandrasekhar(const SidePrim<ND> &L, const SidePrim<ND> &R,
                                                 const ChaAvg<ND> &A, double inv_gm1, double *F)
{
    double ql = 0.0, qr = 0.0, qa = 0.0;
#pragma unroll
    for (int d = 0; d < ND; d++) {
        ql += L.vel[d] * L.vel[d];
        qr += R.vel[d] * R.vel[d];
        qa += A.v[d] * A.v[d];
    }
    if (ND == 3) {   // reference quirk, Euler.jl:216: (... + ur^2 + vr^2 + wl^2)
        qr = R.vel[0] * R.vel[0] + R.vel[1] * R.vel[1] + L.vel[ND - 1] * L.vel[ND - 1];
    }
    const double h = A.beta_inv2 * inv_gm1 - 0.25 * (ql + qr) + A.p_over_rho + qa;
    const double mdot = A.rho * A.v[0];
    F[0] = mdot;
    F[1] = mdot * A.v[0] + A.p;
#pragma unroll
    for (int d = 1; d < ND; d++) F[1 + d] = mdot * A.v[d];
    F[ND + 1] = mdot * h;
}

// Full Euler surface flux in the rotated frame (normal = component 1).
template <int ND>
__device__ __forceinline__ void euler_numflux(const FluxParams &fp, const double *Ql,
                                              const double *Qr, double *F)
{
    constexpr int NV = ND + 2;
    const double g = fp.gamma;
    SidePrim<ND> L, R;
    side_prim<ND>(Ql, g, L);
    side_prim<ND>(Qr, g, R);
    const int kind = fp.numflux;
    const int avg = (kind == FX_STD || kind == FX_CHA) ? kind : fp.numflux_avg;
    const double ipl = fast_rcp(L.p), ipr = fast_rcp(R.p);
    const double bl = 0.5 * L.rho * ipl, br = 0.5 * R.rho * ipr;
    ChaAvg<ND> A;
    const bool need_cha = (avg == FX_CHA) || kind == FX_SCA || kind == FX_MAT;
    if (need_cha) cha_avg<ND>(L, R, bl, br, A);
    if (avg == FX_CHA) nf_chandrasekhar<ND>(L, R, A, fp.inv_gm1, F);
    else nf_stdavg<ND>(Ql, Qr, L, R, F);
    if (kind == FX_STD || kind == FX_CHA) return;

    if (kind == FX_LXF) {
        const double al = sqrt(g * L.p / L.rho), ar = sqrt(g * R.p / R.rho);
        const double lam = fmax(fabs(L.vel[0]) + al, fabs(R.vel[0]) + ar);
        const double c = 0.5 * lam * fp.intensity;
#pragma unroll
        for (int v = 0; v < NV; v++) F[v] = fma(c, Ql[v] - Qr[v], F[v]);
        return;
    }
    if (kind == FX_SCA) {
        const double al = sqrt(g * L.p / L.rho), ar = sqrt(g * R.p / R.rho);
        const double lam = fmax(fabs(L.vel[0]) + al, fabs(R.vel[0]) + ar);
        const double rho = 0.5 * (L.rho + R.rho);
        const double dr = R.rho - L.rho;
        const double gm1 = fp.gm1;
        double dot_lr = 0.0, jump = 0.0;
#pragma unroll
        for (int d = 0; d < ND; d++) {
            dot_lr += L.vel[d] * R.vel[d];
            jump += A.v[d] * (R.vel[d] - L.vel[d]);
        }
        double Dv[NV];
        Dv[0] = dr;
#pragma unroll
        for (int d = 0; d < ND; d++) {
            // reference quirk (3-D only, Euler.jl:257): left momenta read from Qr
            const double ml = (ND == 3) ? Qr[1 + d] : Ql[1 + d];
            Dv[1 + d] = Qr[1 + d] - ml;
        }
        // 1/beta_ln/(g-1) = 2*beta_inv2/(g-1)
        Dv[ND + 1] = (2.0 * A.beta_inv2 / gm1 + dot_lr) * dr * 0.5
                   + rho * (jump + (1.0 / br - 1.0 / bl) / (2.0 * gm1));
        const double c = 0.5 * lam * fp.intensity;
#pragma unroll
        for (int v = 0; v < NV; v++) F[v] -= c * Dv[v];
        return;
    }
    // MatrixDissipation (Euler.jl:303-380): F += R |Lambda| T R' (Wl - Wr) / 2 * intensity
    {
        double ql = 0.0, qr = 0.0, qa = 0.0;
#pragma unroll
        for (int d = 0; d < ND; d++) {
            ql += L.vel[d] * L.vel[d];
            qr += R.vel[d] * R.vel[d];
            qa += A.v[d] * A.v[d];
        }
        const double v2 = 2.0 * qa - 0.5 * (ql + qr);
        // a^2 = g p_hat / rho_ln = g * (p_hat/rho_ln)
        const double a = sqrt(g * A.p_over_rho);
        const double h = g * A.beta_inv2 * fp.inv_gm1 + 0.5 * v2;
        const double u = A.v[0];
        // entropy-variable jump  W = ((g-s)/(g-1) - rho|v|^2/(2p), rho v/p, -rho/p) with
        // s_l - s_r = log(p_l/p_r) - g log(rho_l/rho_r) evaluated as log-ratios
        const double ds = log_ratio(L.p, R.p, fast_rcp(L.p + R.p)) - g * log_ratio(L.rho, R.rho, A.irs);
        double dW[NV];
        dW[0] = -ds * fp.inv_gm1 - (bl * ql - br * qr);
#pragma unroll
        for (int d = 0; d < ND; d++) dW[1 + d] = 2.0 * (bl * L.vel[d] - br * R.vel[d]);
        dW[ND + 1] = -2.0 * (bl - br);
        const double c = 0.5 * fp.intensity;
        // acoustic eigenvectors (1, u -+ a, v, w, h -+ u a)
        {
            double dm = dW[0] + (u - a) * dW[1] + (h - u * a) * dW[ND + 1];
            double dp = dW[0] + (u + a) * dW[1] + (h + u * a) * dW[ND + 1];
#pragma unroll
            for (int d = 1; d < ND; d++) {
                dm += A.v[d] * dW[1 + d];
                dp += A.v[d] * dW[1 + d];
            }
            const double t = 0.5 * A.rho * fp.inv_gamma;
            dm *= fabs(u - a) * t * c;
            dp *= fabs(u + a) * t * c;
            F[0] += dm + dp;
            F[1] += dm * (u - a) + dp * (u + a);
#pragma unroll
            for (int d = 1; d < ND; d++) F[1 + d] += (dm + dp) * A.v[d];
            F[ND + 1] += dm * (h - u * a) + dp * (h + u * a);
        }
        // entropy eigenvector (1, u, v, w, v2/2)
        {
            double de = dW[0] + u * dW[1] + 0.5 * v2 * dW[ND + 1];
#pragma unroll
            for (int d = 1; d < ND; d++) de += A.v[d] * dW[1 + d];
            de *= fabs(u) * (fp.gm1 * A.rho * fp.inv_gamma) * c;
            F[0] += de;
#pragma unroll
            for (int d = 0; d < ND; d++) F[1 + d] += de * A.v[d];
            F[ND + 1] += de * (0.5 * v2);
        }
        // shear eigenvectors e_{1+d} + v_d e_{NV}   (T = p)
#pragma unroll
        for (int d = 1; d < ND; d++) {
            double ds = dW[1 + d] + A.v[d] * dW[ND + 1];
            ds *= fabs(u) * A.p * c;
            F[1 + d] += ds;
            F[ND + 1] += ds * A.v[d];
        }
    }
}

}  // namespace flou
