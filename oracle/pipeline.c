/* ORACLE -- test infrastructure only.  Never linked, imported or executed by the product
 * path (flou.jl_b200/); only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may use it, and only as the checker / CPU baseline.
 *
 * CPU restatement (fp64) of Flou.jl's DGSEM right-hand side, sweep by sweep and in the
 * reference's order of operations, plus the 2N low-storage RK recurrence that
 * OrdinaryDiffEq v6.49.1 (third-party, not under /root/reference) applies around it.
 *
 *   rhs!                     src/FlouSpatial/Equations/Hyperbolic.jl:31-69
 *   project2faces!           src/FlouSpatial/Interfaces.jl:51-109
 *   _volumeflux!             src/FlouSpatial/Equations/OpDivergence.jl:28-37
 *   StrongDivOperator        OpDivergence.jl:105-173
 *   SplitDivOperator         OpDivergence.jl:184-299
 *   SplitDivOperator, nodes  OpDivergence.jl:284-437 (_splitdiv_nb_surface_contribution!: Gauss
 *     without boundaries     nodes, entropy-projected end states; Cartesian sub-grid frames)
 *   HybridDivOperator, nodes OpDivergence.jl:629-779 (_hybrid_nb_surface_contribution!: everything
 *     without boundaries     as a surface contribution)
 *   HybridDivOperator        OpDivergence.jl:452-612, 781-801 (GLL nodes, Cartesian sub-grid
 *                            frames PhysicalRegions.jl:72-148); ORACLE ONLY -- it exists to
 *                            pin the 2-D machinery against the reference's Shockwave2D KAT
 *   applyBCs!                Interfaces.jl:25-49
 *   interface_fluxes!        Interfaces.jl:111-136
 *   _surface_contribution!   OpDivergence.jl:42-100
 *   apply_massmatrix!        src/FlouSpatial/MultielementDiscontinuous.jl:132-137
 *   pointwise physics        src/FlouCommon/Euler.jl:54-307, LinearAdvection.jl:42-44,
 *                            Utilities.jl:34-44
 *   fluxes / rotations / BCs src/FlouSpatial/Equations/Euler.jl:16-536,
 *                            Equations/LinearAdvection.jl:16-47
 *   master2slave             StdRegions/StdSegment.jl:161-167, StdQuad.jl:161-181
 *   tpdofs (line order)      StdRegions/StdQuad.jl:116-124, StdHex.jl:135-145
 *   contravariant            src/FlouSpatial/PhysicalRegions.jl:1001-1033
 *   LSRK-2N update           OrdinaryDiffEq LowStorageRK2N perform_step! (restated from
 *                            the published 2N recurrence; call site FlouTime.jl:34-38)
 *
 * Reference quirks reproduced on purpose: `wl^2` used twice in
 * the 3-D ChandrasekharAverage numerical flux (Euler.jl:216); ScalarDissipation 3-D reads
 * the left momenta from Qr (Euler.jl:257); logarithmic_mean threshold/series
 * (Utilities.jl:34-44); true division by jac in the mass solve.
 *
 * Compile with -ffp-contract=off: Julia does not contract a*b+c outside @muladd.
 * Threading: OpenMP `parallel for` over elements / faces exactly where the reference
 * has `@flouthreads` (Polyester @batch).
 */
#include <math.h>
#include <omp.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define MAXV 5
#define MAXNP 16

enum { EQ_ADVECTION = 0, EQ_EULER = 1 };
enum { OP_STRONG = 0, OP_SPLIT = 1, OP_HYBRID = 2 };
enum { FLUX_STDAVG = 0, FLUX_LXF = 1, FLUX_CHANDRASEKHAR = 2, FLUX_SCALARDISS = 3,
       FLUX_MATRIXDISS = 4 };
enum { BC_INFLOW = 0, BC_OUTFLOW = 1, BC_SLIP = 2, BC_TABLE = 3 };

typedef struct {
    int32_t nd, nv, np;
    int32_t equation;         /* EQ_* */
    int32_t op;               /* OP_* */
    int32_t tpflux;           /* two-point flux of the split form: STDAVG or CHANDRASEKHAR */
    int32_t numflux;          /* surface flux */
    int32_t numflux_avg;      /* .avg of LxF / ScalarDissipation / MatrixDissipation */
    double intensity;
    double gamma;
    double a[3];
    int64_t ne, nf;
    /* connectivity, reference layout, 1-based ids */
    const int64_t *faceinds;  /* ne * 2nd */
    const int64_t *facepos;   /* ne * 2nd */
    const int64_t *eleminds;  /* nf * 2  (0 = none) */
    const int64_t *elempos;   /* nf * 2 */
    const uint8_t *orientation; /* nf */
    /* 1-D operators, column-major np x np like Julia */
    const double *D, *Ds, *Dsharp, *lm, *lp, *dgl, *dgr;
    /* geometry */
    const double *jac;        /* ndof */
    const double *metric;     /* ndof * nd * nd, [c + nd*d] = Ja^d_c */
    const double *fjac;       /* nf*nfp */
    const double *frames;     /* nf*nfp * 3 * nd : n, t, b */
    /* boundary conditions */
    int32_t nbound;
    const int32_t *bc_kind;   /* nbound */
    const int64_t *bc_offsets;/* nbound+1, into bc_faces */
    const int64_t *bc_faces;  /* 1-based face ids */
    const double *bc_state;   /* INFLOW: nbound * nv (row per boundary) */
    const double *bc_table;   /* TABLE : (total bd faces * nfp) * nv, row-major per face node,
                                 indexed by position in bc_faces */
    /* work arrays (allocated by the caller) */
    double *Qf[2];            /* each (nf*nfp) * nv column-major */
    double *Fn[2];
    /* HybridDivOperator only */
    double blend;             /* op.blend */
    const double *w1d;        /* 1-D weights of the standard region */
    double sub_jac[3];        /* Cartesian sub-grid face Jacobians by direction (kept for reference) */
    /* geometry.subgrids by (element, direction, line, position along the line):
     * sub_frames[(((e*nd + d)*nlines + k)*(np+1) + ii)*3*nd + r*nd + c], rows r = n, t, b;
     * sub_fjac[((e*nd + d)*nlines + k)*(np+1) + ii]   (PhysicalRegions.jl:72-292) */
    const double *sub_frames;
    const double *sub_fjac;
    /* std |> basis |> hasboundaries (GLL/CGL: 1, GL: 0): selects the surface term of the split form */
    int32_t hasboundaries;
    /* TEST DEVICE, not the reference's behaviour: face traces of the Gauss-node split form from the
     * entropy-projected end states instead of l'Q.  With it the scheme is exactly entropy
     * conservative, which is how tests/test_oracle_properties.py pins the restatement of the
     * projection terms; the reference interpolates the conservative variables (Interfaces.jl:93-109). */
    int32_t proj_traces;
} oracle_problem;

static inline int ipow(int b, int e) { int r = 1; while (e-- > 0) r *= b; return r; }

/* ---------------------------------------------------------------- pointwise physics */
static inline double logarithmic_mean(double al, double ar)
{   /* Utilities.jl:34-44 */
    double xi = al / ar;
    double f = (xi - 1) / (xi + 1);
    double u = f * f;
    double F;
    if (u < 0.01) F = 1 + u / 3 + u * u / 5 + u * u * u / 7;
    else F = log(xi) / (2 * f);
    return (al + ar) / (2 * F);
}

static inline double pressure(const double *Q, int nd, double g)
{   /* FlouCommon/Euler.jl:142-155 */
    double m2 = 0;
    for (int d = 0; d < nd; d++) m2 += Q[1 + d] * Q[1 + d];
    return (g - 1) * (Q[nd + 1] - m2 / (2 * Q[0]));
}

/* P = (rho, u.., p)   FlouCommon/Euler.jl:237-253 */
static inline void cons2prim(const double *Q, int nd, double g, double *P)
{
    P[0] = Q[0];
    for (int d = 0; d < nd; d++) P[1 + d] = Q[1 + d] / Q[0];
    P[nd + 1] = pressure(Q, nd, g);
}

static inline double soundvelocity(double rho, double p, double g) { return sqrt(g * p / rho); }

/* FlouCommon/Euler.jl:273-307 */
static inline void cons2entropy(const double *Q, int nd, double g, double *W)
{
    double rho = Q[0];
    double p = pressure(Q, nd, g);
    double s = log(p) - g * log(rho);
    double m2 = 0;
    for (int d = 0; d < nd; d++) m2 += Q[1 + d] * Q[1 + d];
    W[0] = (g - s) / (g - 1) - m2 / rho / (2 * p);
    for (int d = 0; d < nd; d++) W[1 + d] = Q[1 + d] / p;
    W[nd + 1] = -rho / p;
}

/* vars_entropy2cons = vars_prim2cons(vars_entropy2prim(W))   FlouCommon/Euler.jl:309-339, 255-271 */
static inline void entropy2cons(const double *W, int nd, double g, double *Q)
{
    double wn = W[nd + 1], vel[3], q = 0;
    for (int d = 0; d < nd; d++) { vel[d] = -W[1 + d] / wn; q += vel[d] * vel[d]; }
    double s = g - (g - 1) * (W[0] - wn * q / 2);
    double p = pow(pow(-wn, g) * exp(s), 1 / (1 - g));
    double rho = -p * wn;
    Q[0] = rho;
    for (int d = 0; d < nd; d++) Q[1 + d] = rho * vel[d];
    Q[nd + 1] = p / (g - 1) + rho * q / 2;
}

/* volumeflux: F[c][v], c = physical direction   FlouCommon/Euler.jl:54-114 */
static inline void volumeflux(const oracle_problem *P, const double *Q, double F[3][MAXV])
{
    int nd = P->nd;
    if (P->equation == EQ_ADVECTION) {
        for (int c = 0; c < nd; c++) F[c][0] = P->a[c] * Q[0];
        return;
    }
    double rho = Q[0], vel[3];
    for (int d = 0; d < nd; d++) vel[d] = Q[1 + d] / rho;
    double p = pressure(Q, nd, P->gamma);
    double rhoe = Q[nd + 1];
    for (int c = 0; c < nd; c++) {
        F[c][0] = Q[1 + c];
        for (int d = 0; d < nd; d++) {
            F[c][1 + d] = Q[1 + c] * vel[d];
            if (d == c) F[c][1 + d] = Q[1 + c] * vel[d] + p;
        }
        F[c][nd + 1] = (rhoe + p) * vel[c];
    }
}

/* ---------------------------------------------------------------- rotations */
static inline void rotate2face(const oracle_problem *P, const double *Q, const double *fr, double *R)
{   /* Equations/Euler.jl:16-52 ; LinearAdvection.jl:16-18 */
    int nd = P->nd;
    if (P->equation == EQ_ADVECTION) { R[0] = Q[0]; return; }
    const double *n = fr, *t = fr + nd, *b = fr + 2 * nd;
    R[0] = Q[0];
    if (nd == 1) { R[1] = Q[1] * n[0]; R[2] = Q[2]; }
    else if (nd == 2) {
        R[1] = Q[1] * n[0] + Q[2] * n[1];
        R[2] = Q[1] * t[0] + Q[2] * t[1];
        R[3] = Q[3];
    } else {
        R[1] = Q[1] * n[0] + Q[2] * n[1] + Q[3] * n[2];
        R[2] = Q[1] * t[0] + Q[2] * t[1] + Q[3] * t[2];
        R[3] = Q[1] * b[0] + Q[2] * b[1] + Q[3] * b[2];
        R[4] = Q[4];
    }
}

static inline void rotate2phys(const oracle_problem *P, const double *R, const double *fr, double *Q)
{   /* Equations/Euler.jl:20-64 ; LinearAdvection.jl:20-22 */
    int nd = P->nd;
    if (P->equation == EQ_ADVECTION) { Q[0] = R[0]; return; }
    const double *n = fr, *t = fr + nd, *b = fr + 2 * nd;
    Q[0] = R[0];
    if (nd == 1) { Q[1] = R[1] * n[0]; Q[2] = R[2]; }
    else if (nd == 2) {
        Q[1] = R[1] * n[0] + R[2] * t[0];
        Q[2] = R[1] * n[1] + R[2] * t[1];
        Q[3] = R[3];
    } else {
        Q[1] = R[1] * n[0] + R[2] * t[0] + R[3] * b[0];
        Q[2] = R[1] * n[1] + R[2] * t[1] + R[3] * b[1];
        Q[3] = R[1] * n[2] + R[2] * t[2] + R[3] * b[2];
        Q[4] = R[4];
    }
}

/* ---------------------------------------------------------------- numerical fluxes */
static void numflux_euler_stdavg(const double *Ql, const double *Qr, int nd, double g, double *F)
{   /* Equations/Euler.jl:99-136 */
    double Pl[MAXV], Pr[MAXV];
    cons2prim(Ql, nd, g, Pl); cons2prim(Qr, nd, g, Pr);
    double ul = Pl[1], ur = Pr[1], pl = Pl[nd + 1], pr = Pr[nd + 1];
    F[0] = (Ql[1] + Qr[1]) / 2;
    F[1] = (Ql[1] * ul + pl + Qr[1] * ur + pr) / 2;
    for (int d = 1; d < nd; d++) F[1 + d] = (Ql[1 + d] * ul + Qr[1 + d] * ur) / 2;
    F[nd + 1] = ((Ql[nd + 1] + pl) * ul + (Qr[nd + 1] + pr) * ur) / 2;
}

static void numflux_euler_chandrasekhar(const double *Ql, const double *Qr, int nd, double g, double *F)
{   /* Equations/Euler.jl:167-226 (incl. the wl^2-twice quirk of line 216) */
    double Pl[MAXV], Pr[MAXV];
    cons2prim(Ql, nd, g, Pl); cons2prim(Qr, nd, g, Pr);
    double rl = Pl[0], rr = Pr[0], pl = Pl[nd + 1], pr = Pr[nd + 1];
    double ul = Pl[1], ur = Pr[1];
    double u = (ul + ur) / 2, v = 0, w = 0;
    double vl = 0, vr = 0, wl = 0, wr = 0;
    if (nd >= 2) { vl = Pl[2]; vr = Pr[2]; v = (vl + vr) / 2; }
    if (nd >= 3) { wl = Pl[3]; wr = Pr[3]; w = (wl + wr) / 2; }
    double bl = rl / (2 * pl), br = rr / (2 * pr);
    double rho = logarithmic_mean(rl, rr);
    double p = (rl + rr) / (2 * (bl + br));
    double beta = logarithmic_mean(bl, br);
    (void)wr;
    if (nd == 1) {
        double h = 1 / (2 * beta * (g - 1)) - (ul * ul + ur * ur) / 4 + p / rho + u * u;
        F[0] = rho * u; F[1] = rho * (u * u) + p; F[2] = rho * u * h;
    } else if (nd == 2) {
        double h = 1 / (2 * beta * (g - 1)) - (ul * ul + vl * vl + ur * ur + vr * vr) / 4
                   + p / rho + u * u + v * v;
        F[0] = rho * u; F[1] = rho * (u * u) + p; F[2] = rho * u * v; F[3] = rho * u * h;
    } else {
        double h = 1 / (2 * beta * (g - 1))
                   - (ul * ul + vl * vl + wl * wl + ur * ur + vr * vr + wl * wl) / 4
                   + p / rho + u * u + v * v + w * w;
        F[0] = rho * u; F[1] = rho * (u * u) + p; F[2] = rho * u * v; F[3] = rho * u * w;
        F[4] = rho * u * h;
    }
}

static void numflux_euler_avg(int kind, const double *Ql, const double *Qr, int nd, double g, double *F)
{
    if (kind == FLUX_CHANDRASEKHAR) numflux_euler_chandrasekhar(Ql, Qr, nd, g, F);
    else numflux_euler_stdavg(Ql, Qr, nd, g, F);
}

static void numflux_euler(const oracle_problem *P, const double *Ql, const double *Qr, double *F)
{
    int nd = P->nd, nv = P->nv;
    double g = P->gamma;
    switch (P->numflux) {
    case FLUX_STDAVG: numflux_euler_stdavg(Ql, Qr, nd, g, F); return;
    case FLUX_CHANDRASEKHAR: numflux_euler_chandrasekhar(Ql, Qr, nd, g, F); return;
    case FLUX_LXF: {   /* Equations/Euler.jl:138-165 */
        numflux_euler_avg(P->numflux_avg, Ql, Qr, nd, g, F);
        double Pl[MAXV], Pr[MAXV];
        cons2prim(Ql, nd, g, Pl); cons2prim(Qr, nd, g, Pr);
        double al = soundvelocity(Pl[0], Pl[nd + 1], g);
        double ar = soundvelocity(Pr[0], Pr[nd + 1], g);
        double lam = fmax(fabs(Pl[1]) + al, fabs(Pr[1]) + ar);
        for (int v = 0; v < nv; v++) F[v] = F[v] + lam * (Ql[v] - Qr[v]) / 2 * P->intensity;
        return;
    }
    case FLUX_SCALARDISS: {   /* Equations/Euler.jl:228-301 (incl. the Qr quirk of line 257) */
        double Pl[MAXV], Pr[MAXV];
        cons2prim(Ql, nd, g, Pl); cons2prim(Qr, nd, g, Pr);
        double rl = Pl[0], rr = Pr[0], pl = Pl[nd + 1], pr = Pr[nd + 1];
        double ul = Pl[1], ur = Pr[1], vl = 0, vr = 0, wl = 0, wr = 0;
        double u = (ul + ur) / 2, v = 0, w = 0;
        double ml[3], mr[3];
        for (int d = 0; d < nd; d++) { ml[d] = Ql[1 + d]; mr[d] = Qr[1 + d]; }
        if (nd >= 2) { vl = Pl[2]; vr = Pr[2]; v = (vl + vr) / 2; }
        if (nd == 3) {
            wl = Pl[3]; wr = Pr[3]; w = (wl + wr) / 2;
            for (int d = 0; d < 3; d++) ml[d] = Qr[1 + d];
        }
        double rho = (rl + rr) / 2;
        double bl = rl / (2 * pl), br = rr / (2 * pr);
        double beta = logarithmic_mean(bl, br);
        double al = soundvelocity(rl, pl, g), ar = soundvelocity(rr, pr, g);
        numflux_euler_avg(P->numflux_avg, Ql, Qr, nd, g, F);
        double lam = fmax(fabs(ul) + al, fabs(ur) + ar);
        double Dv[MAXV];
        Dv[0] = rr - rl;
        for (int d = 0; d < nd; d++) Dv[1 + d] = mr[d] - ml[d];
        double gm1 = g - 1;
        if (nd == 1)
            Dv[2] = (1 / beta / gm1 + ul * ur) * (rr - rl) / 2
                  + rho * (u * (ur - ul) + (1 / br - 1 / bl) / (2 * gm1));
        else if (nd == 2)
            Dv[3] = (1 / beta / gm1 + ul * ur + vl * vr) * (rr - rl) / 2
                  + rho * (u * (ur - ul) + v * (vr - vl) + (1 / br - 1 / bl) / (2 * gm1));
        else
            Dv[4] = (1 / beta / gm1 + ul * ur + vl * vr + wl * wr) * (rr - rl) / 2
                  + rho * (u * (ur - ul) + v * (vr - vl) + w * (wr - wl)
                           + (1 / br - 1 / bl) / (2 * gm1));
        for (int k = 0; k < nv; k++) F[k] = F[k] - lam / 2 * Dv[k] * P->intensity;
        return;
    }
    case FLUX_MATRIXDISS: {   /* Equations/Euler.jl:303-380 */
        double Pl[MAXV], Pr[MAXV];
        cons2prim(Ql, nd, g, Pl); cons2prim(Qr, nd, g, Pr);
        double rl = Pl[0], rr = Pr[0], pl = Pl[nd + 1], pr = Pr[nd + 1];
        double ul = Pl[1], ur = Pr[1], vl = 0, vr = 0, wl = 0, wr = 0;
        double u = (ul + ur) / 2, v = 0, w = 0, v2;
        if (nd == 1) v2 = 2 * (u * u) - (ul * ul + ur * ur) / 2;
        else if (nd == 2) {
            vl = Pl[2]; vr = Pr[2]; v = (vl + vr) / 2;
            v2 = 2 * (u * u + v * v) - (ul * ul + vl * vl + ur * ur + vr * vr) / 2;
        } else {
            vl = Pl[2]; vr = Pr[2]; v = (vl + vr) / 2;
            wl = Pl[3]; wr = Pr[3]; w = (wl + wr) / 2;
            v2 = 2 * (u * u + v * v + w * w)
               - (ul * ul + vl * vl + wl * wl + ur * ur + vr * vr + wr * wr) / 2;
        }
        double bl = rl / (2 * pl), br = rr / (2 * pr);
        double rho = logarithmic_mean(rl, rr);
        double p = (rl + rr) / (2 * (bl + br));
        double beta = logarithmic_mean(bl, br);
        double a = soundvelocity(rho, p, g);
        double h = g / (2 * beta) / (g - 1) + v2 / 2;
        numflux_euler_avg(P->numflux_avg, Ql, Qr, nd, g, F);
        double Wl[MAXV], Wr[MAXV];
        cons2entropy(Ql, nd, g, Wl); cons2entropy(Qr, nd, g, Wr);
        /* R columns (SMatrix literal is column-major): R[row][col] */
        double R[MAXV][MAXV], Lam[MAXV], T[MAXV];
        memset(R, 0, sizeof R);
        if (nd == 1) {
            Lam[0] = fabs(u - a); Lam[1] = fabs(u); Lam[2] = fabs(u + a);
            T[0] = rho / (2 * g); T[1] = (g - 1) * rho / g; T[2] = rho / (2 * g);
            R[0][0] = 1; R[1][0] = u - a; R[2][0] = h - u * a;
            R[0][1] = 1; R[1][1] = u;     R[2][1] = v2 / 2;
            R[0][2] = 1; R[1][2] = u + a; R[2][2] = h + u * a;
        } else if (nd == 2) {
            Lam[0] = fabs(u - a); Lam[1] = fabs(u); Lam[2] = fabs(u); Lam[3] = fabs(u + a);
            T[0] = rho / (2 * g); T[1] = (g - 1) * rho / g; T[2] = p; T[3] = rho / (2 * g);
            R[0][0] = 1; R[1][0] = u - a; R[2][0] = v; R[3][0] = h - u * a;
            R[0][1] = 1; R[1][1] = u;     R[2][1] = v; R[3][1] = v2 / 2;
            R[0][2] = 0; R[1][2] = 0;     R[2][2] = 1; R[3][2] = v;
            R[0][3] = 1; R[1][3] = u + a; R[2][3] = v; R[3][3] = h + u * a;
        } else {
            Lam[0] = fabs(u - a); Lam[1] = fabs(u); Lam[2] = fabs(u); Lam[3] = fabs(u);
            Lam[4] = fabs(u + a);
            T[0] = rho / (2 * g); T[1] = (g - 1) * rho / g; T[2] = p; T[3] = p;
            T[4] = rho / (2 * g);
            R[0][0] = 1; R[1][0] = u - a; R[2][0] = v; R[3][0] = w; R[4][0] = h - u * a;
            R[0][1] = 1; R[1][1] = u;     R[2][1] = v; R[3][1] = w; R[4][1] = v2 / 2;
            R[0][2] = 0; R[1][2] = 0;     R[2][2] = 1; R[3][2] = 0; R[4][2] = v;
            R[0][3] = 0; R[1][3] = 0;     R[2][3] = 0; R[3][3] = 1; R[4][3] = w;
            R[0][4] = 1; R[1][4] = u + a; R[2][4] = v; R[3][4] = w; R[4][4] = h + u * a;
        }
        /* Fn + R*Lam*T*R'*(Wl-Wr)/2*intensity, evaluated left to right like Julia:
           M = ((R*Lam)*T)*R' ; then M*(Wl-Wr) */
        double M1[MAXV][MAXV], M[MAXV][MAXV];
        for (int i = 0; i < nv; i++)
            for (int j = 0; j < nv; j++) M1[i][j] = R[i][j] * Lam[j] * T[j];
        for (int i = 0; i < nv; i++)
            for (int j = 0; j < nv; j++) {
                double s = 0;
                for (int k = 0; k < nv; k++) s += M1[i][k] * R[j][k];
                M[i][j] = s;
            }
        for (int i = 0; i < nv; i++) {
            double s = 0;
            for (int j = 0; j < nv; j++) s += M[i][j] * (Wl[j] - Wr[j]);
            F[i] = F[i] + s / 2 * P->intensity;
        }
        return;
    }
    }
}

static void numericalflux(const oracle_problem *P, const double *Ql, const double *Qr,
                          const double *n, double *F)
{
    if (P->equation == EQ_ADVECTION) {   /* Equations/LinearAdvection.jl:27-39 */
        double an = 0;
        for (int d = 0; d < P->nd; d++) an += P->a[d] * n[d];
        double Fa = an * (Ql[0] + Qr[0]) / 2;
        if (P->numflux == FLUX_LXF) Fa = Fa + fabs(an) * (Ql[0] - Qr[0]) / 2 * P->intensity;
        F[0] = Fa;
        return;
    }
    numflux_euler(P, Ql, Qr, F);
}

/* ---------------------------------------------------------------- two-point fluxes */
static void twopointflux(const oracle_problem *P, const double *Q1, const double *Q2,
                         const double *Ja1, const double *Ja2, double *F)
{
    int nd = P->nd;
    double g = P->gamma;
    double n[3];
    for (int c = 0; c < nd; c++) n[c] = (Ja1[c] + Ja2[c]) / 2;
    if (P->equation == EQ_ADVECTION) {   /* Equations/LinearAdvection.jl:44-47 */
        double an = 0;
        for (int d = 0; d < nd; d++) an += P->a[d] * n[d];
        F[0] = an * (Q1[0] + Q2[0]) / 2;
        return;
    }
    double P1[MAXV], P2[MAXV];
    cons2prim(Q1, nd, g, P1); cons2prim(Q2, nd, g, P2);
    if (P->tpflux == FLUX_STDAVG) {   /* Equations/Euler.jl:385-472 */
        double p1 = P1[nd + 1], p2 = P2[nd + 1];
        double f[MAXV][3];
        for (int c = 0; c < nd; c++) {
            f[0][c] = (Q1[1 + c] + Q2[1 + c]) / 2;
            for (int d = 0; d < nd; d++) {
                if (d == c)
                    f[1 + d][c] = (Q1[1 + d] * P1[1 + c] + p1 + Q2[1 + d] * P2[1 + c] + p2) / 2;
                else
                    f[1 + d][c] = (Q1[1 + d] * P1[1 + c] + Q2[1 + d] * P2[1 + c]) / 2;
            }
            f[nd + 1][c] = ((Q1[nd + 1] + p1) * P1[1 + c] + (Q2[nd + 1] + p2) * P2[1 + c]) / 2;
        }
        for (int v = 0; v < nd + 2; v++) {
            double s = f[v][0] * n[0];
            for (int c = 1; c < nd; c++) s += f[v][c] * n[c];
            F[v] = s;
        }
        return;
    }
    /* ChandrasekharAverage   Equations/Euler.jl:474-536 */
    double r1 = P1[0], r2 = P2[0], p1 = P1[nd + 1], p2 = P2[nd + 1];
    double u1 = P1[1], u2 = P2[1], v1 = 0, v2 = 0, w1 = 0, w2 = 0;
    double u = (u1 + u2) / 2, v = 0, w = 0;
    if (nd >= 2) { v1 = P1[2]; v2 = P2[2]; v = (v1 + v2) / 2; }
    if (nd >= 3) { w1 = P1[3]; w2 = P2[3]; w = (w1 + w2) / 2; }
    double b1 = r1 / (2 * p1), b2 = r2 / (2 * p2);
    double rho = logarithmic_mean(r1, r2);
    double p = (r1 + r2) / (2 * (b1 + b2));
    double beta = logarithmic_mean(b1, b2);
    if (nd == 1) {
        double h = 1 / (2 * beta * (g - 1)) - (u1 * u1 + u2 * u2) / 4 + p / rho + u * u;
        F[0] = (rho * u) * n[0];
        F[1] = (rho * (u * u) + p) * n[0];
        F[2] = (rho * u * h) * n[0];
    } else if (nd == 2) {
        double h = 1 / (2 * beta * (g - 1)) - (u1 * u1 + v1 * v1 + u2 * u2 + v2 * v2) / 4
                   + p / rho + u * u + v * v;
        F[0] = (rho * u) * n[0] + (rho * v) * n[1];
        F[1] = (rho * (u * u) + p) * n[0] + (rho * u * v) * n[1];
        F[2] = (rho * u * v) * n[0] + (rho * (v * v) + p) * n[1];
        F[3] = (rho * u * h) * n[0] + (rho * v * h) * n[1];
    } else {
        double h = 1 / (2 * beta * (g - 1))
                   - (u1 * u1 + v1 * v1 + w1 * w1 + u2 * u2 + v2 * v2 + w2 * w2) / 4
                   + p / rho + u * u + v * v + w * w;
        F[0] = (rho * u) * n[0] + (rho * v) * n[1] + (rho * w) * n[2];
        F[1] = (rho * (u * u) + p) * n[0] + (rho * u * v) * n[1] + (rho * u * w) * n[2];
        F[2] = (rho * u * v) * n[0] + (rho * (v * v) + p) * n[1] + (rho * v * w) * n[2];
        F[3] = (rho * u * w) * n[0] + (rho * v * w) * n[1] + (rho * (w * w) + p) * n[2];
        F[4] = (rho * u * h) * n[0] + (rho * v * h) * n[1] + (rho * w * h) * n[2];
    }
}

/* ---------------------------------------------------------------- index helpers */
/* first node and stride of line k (0-based) of direction d (0-based): tpdofs order */
static inline void line_of(int nd, int np, int d, int k, int *base, int *stride)
{
    if (nd == 1) { *base = 0; *stride = 1; return; }
    if (nd == 2) {
        if (d == 0) { *base = np * k; *stride = 1; }
        else { *base = k; *stride = np; }
        return;
    }
    if (d == 0) { *base = np * k; *stride = 1; }
    else if (d == 1) { *base = (k % np) + np * np * (k / np); *stride = np; }
    else { *base = k; *stride = np * np; }
}

/* master2slave, 0-based face dof i -> slave face dof j */
static inline int master2slave(int nd, int np, int i, int o)
{
    if (nd <= 1) return i;
    if (nd == 2) return o == 0 ? i : np - 1 - i;   /* StdSegment.jl:161-167 */
    /* StdQuad.jl:161-181 ; m = (m1, m2) 1-based, li[a, b] = a + np*(b-1) */
    int m1 = i % np + 1, m2 = i / np + 1, a, b;
    switch (o) {
    case 0: return i;
    case 1: a = m2; b = np - m1 + 1; break;
    case 2: a = np - m1 + 1; b = np - m2 + 1; break;
    case 3: a = np - m2 + 1; b = m1; break;
    case 4: a = m2; b = m1; break;
    case 5: a = np - m1 + 1; b = m2; break;
    case 6: a = np - m2 + 1; b = np - m1 + 1; break;
    default: a = m1; b = np - m2 + 1; break;
    }
    return (a - 1) + np * (b - 1);
}

/* ---------------------------------------------------------------- the RHS */
void oracle_rhs(const oracle_problem *P, const double *Q, double *dQ, double time)
{
    (void)time;
    const int nd = P->nd, nv = P->nv, np = P->np;
    const int npts = ipow(np, nd), nfp = ipow(np, nd - 1), nlines = nfp;
    const int64_t ne = P->ne, nf = P->nf;
    const int64_t ndof = ne * npts, nfd = nf * nfp;
    double *Qf0 = P->Qf[0], *Qf1 = P->Qf[1], *Fn0 = P->Fn[0], *Fn1 = P->Fn[1];

    /* fill!(dQ, 0)   Hyperbolic.jl:45 */
    memset(dQ, 0, sizeof(double) * ndof * nv);

    /* project2faces!   Interfaces.jl:51-109 */
    #pragma omp parallel for schedule(static)
    for (int64_t e = 0; e < ne; e++) {
        const int64_t *faces = P->faceinds + e * 2 * nd, *sides = P->facepos + e * 2 * nd;
        for (int d = 0; d < nd; d++) {
            double *QL = sides[2 * d] == 1 ? Qf0 : Qf1;
            double *QR = sides[2 * d + 1] == 1 ? Qf0 : Qf1;
            int64_t fl = (faces[2 * d] - 1) * nfp, fr = (faces[2 * d + 1] - 1) * nfp;
            for (int k = 0; k < nlines; k++) {
                int base, stride;
                line_of(nd, np, d, k, &base, &stride);
                for (int v = 0; v < nv; v++) {
                    double sl = 0, sr = 0;
                    for (int ii = 0; ii < np; ii++)
                        sl += P->lm[ii] * Q[e * npts + base + ii * stride + ndof * v];
                    for (int ii = 0; ii < np; ii++)
                        sr += P->lp[ii] * Q[e * npts + base + ii * stride + ndof * v];
                    QL[fl + k + nfd * v] = sl;
                    QR[fr + k + nfd * v] = sr;
                }
                if (P->proj_traces && P->equation == EQ_EULER) {
                    double Wl[MAXV], Wr[MAXV], Qa[MAXV], Qb[MAXV];
                    for (int v = 0; v < nv; v++) { Wl[v] = 0; Wr[v] = 0; }
                    for (int ii = 0; ii < np; ii++) {
                        double Qi[MAXV], Wi[MAXV];
                        for (int v = 0; v < nv; v++) Qi[v] = Q[e * npts + base + ii * stride + ndof * v];
                        cons2entropy(Qi, nd, P->gamma, Wi);
                        for (int v = 0; v < nv; v++) { Wl[v] += P->lm[ii] * Wi[v]; Wr[v] += P->lp[ii] * Wi[v]; }
                    }
                    entropy2cons(Wl, nd, P->gamma, Qa);
                    entropy2cons(Wr, nd, P->gamma, Qb);
                    for (int v = 0; v < nv; v++) { QL[fl + k + nfd * v] = Qa[v]; QR[fr + k + nfd * v] = Qb[v]; }
                }
            }
        }
    }

    /* volume_contribution!   Hyperbolic.jl:71-77 -> OpDivergence.jl:109-160 / 200-282 */
    #pragma omp parallel
    {
        double *Ft = (double *)malloc(sizeof(double) * npts * nv * nd);       /* F~[i][v][d] */
        double *Fs = (double *)malloc(sizeof(double) * np * nv * npts);       /* F#[a][v][node] */
        #pragma omp for schedule(static)
        for (int64_t e = 0; e < ne; e++) {
            const double *Ja = P->metric + e * npts * nd * nd;
            /* _volumeflux! */
            for (int i = 0; i < npts; i++) {
                double Qi[MAXV], F[3][MAXV];
                for (int v = 0; v < nv; v++) Qi[v] = Q[e * npts + i + ndof * v];
                volumeflux(P, Qi, F);
                const double *M = Ja + i * nd * nd;
                for (int d = 0; d < nd; d++)
                    for (int v = 0; v < nv; v++) {
                        double s = F[0][v] * M[0 + nd * d];
                        for (int c = 1; c < nd; c++) s += F[c][v] * M[c + nd * d];
                        Ft[(i * nv + v) * nd + d] = s;
                    }
            }
            for (int d = 0; d < nd; d++) {
                if (P->op == OP_STRONG) {
                    /* mul!(dQ[line], Ds, F~[line, d], -1, 1) */
                    for (int k = 0; k < nlines; k++) {
                        int base, stride;
                        line_of(nd, np, d, k, &base, &stride);
                        for (int v = 0; v < nv; v++)
                            for (int ii = 0; ii < np; ii++) {
                                double s = 0;
                                for (int jj = 0; jj < np; jj++)
                                    s += P->Ds[ii + np * jj]
                                       * Ft[((base + jj * stride) * nv + v) * nd + d];
                                dQ[e * npts + base + ii * stride + ndof * v] -= s;
                            }
                    }
                } else if (P->op == OP_HYBRID) {
                    /* _vol_hybrid_tensorproduct!   OpDivergence.jl:554-612 (fvflux = numflux); on
                     * nodes without boundaries everything is a surface contribution (:478-492) */
                    for (int k = 0; k < nlines && P->hasboundaries; k++) {
                        int base, stride;
                        line_of(nd, np, d, k, &base, &stride);
                        double Fb[MAXNP + 1][MAXV];
                        memset(Fb, 0, sizeof Fb);
                        /* sub-grid frames / Jacobians of this line: PhysicalRegions.jl:72-292 */
                        const int64_t sg0 = (((int64_t)e * nd + d) * nlines + k) * (np + 1);
                        for (int ii = 1; ii < np; ii++) {          /* Julia ii = 2..npts */
                            for (int ik = ii; ik < np; ik++) {
                                int kk = base + ik * stride;
                                for (int il = 0; il < ii; il++) {
                                    int l = base + il * stride;
                                    double Ql[MAXV], Qk[MAXV], F[MAXV];
                                    for (int v = 0; v < nv; v++) {
                                        Ql[v] = Q[e * npts + l + ndof * v];
                                        Qk[v] = Q[e * npts + kk + ndof * v];
                                    }
                                    twopointflux(P, Ql, Qk, Ja + l * nd * nd + nd * d,
                                                 Ja + kk * nd * nd + nd * d, F);
                                    for (int v = 0; v < nv; v++)
                                        Fb[ii][v] += 2 * P->w1d[il] * P->D[il + np * ik] * F[v];
                                }
                            }
                            int i = base + ii * stride, il = base + (ii - 1) * stride;
                            double Qa[MAXV], Qb[MAXV], Qln[MAXV], Qrn[MAXV], Fn_[MAXV], Fv[MAXV];
                            for (int v = 0; v < nv; v++) {
                                Qa[v] = Q[e * npts + il + ndof * v];
                                Qb[v] = Q[e * npts + i + ndof * v];
                            }
                            const double *fr = P->sub_frames + (sg0 + ii) * 3 * nd;
                            rotate2face(P, Qa, fr, Qln);
                            rotate2face(P, Qb, fr, Qrn);
                            numericalflux(P, Qln, Qrn, fr, Fn_);
                            rotate2phys(P, Fn_, fr, Fv);
                            for (int v = 0; v < nv; v++) Fv[v] *= P->sub_fjac[sg0 + ii];
                            double Wl[MAXV], Wr[MAXV], b = 0;
                            cons2entropy(Qa, nd, P->gamma, Wl);
                            cons2entropy(Qb, nd, P->gamma, Wr);
                            for (int v = 0; v < nv; v++) b += (Wr[v] - Wl[v]) * (Fb[ii][v] - Fv[v]);
                            double delta = sqrt(b * b + P->blend);      /* _hybrid_compute_delta */
                            delta = (delta - b) / delta;
                            delta = fmax(delta, 0.5);
                            for (int v = 0; v < nv; v++)
                                Fb[ii][v] = (1 - delta) * Fv[v] + delta * Fb[ii][v];
                        }
                        for (int ii = 0; ii < np; ii++)
                            for (int v = 0; v < nv; v++)
                                dQ[e * npts + base + ii * stride + ndof * v] +=
                                    (Fb[ii][v] - Fb[ii + 1][v]) / P->w1d[ii];
                    }
                } else {
                    /* _flux_splitdiv_tensorproduct!   OpDivergence.jl:248-271 */
                    for (int k = 0; k < nlines; k++) {
                        int base, stride;
                        line_of(nd, np, d, k, &base, &stride);
                        for (int ii = 0; ii < np; ii++) {
                            int i = base + ii * stride;
                            for (int v = 0; v < nv; v++)
                                Fs[(ii * nv + v) * npts + i] = Ft[(i * nv + v) * nd + d];
                            for (int il = ii + 1; il < np; il++) {
                                int l = base + il * stride;
                                double Qi[MAXV], Ql[MAXV], F[MAXV];
                                for (int v = 0; v < nv; v++) {
                                    Qi[v] = Q[e * npts + i + ndof * v];
                                    Ql[v] = Q[e * npts + l + ndof * v];
                                }
                                twopointflux(P, Qi, Ql, Ja + i * nd * nd + nd * d,
                                             Ja + l * nd * nd + nd * d, F);
                                for (int v = 0; v < nv; v++) {
                                    Fs[(il * nv + v) * npts + i] = F[v];
                                    Fs[(ii * nv + v) * npts + l] = F[v];
                                }
                            }
                        }
                    }
                    /* _vol_splitdiv_tensorproduct!   OpDivergence.jl:273-282 */
                    for (int k = 0; k < nlines; k++) {
                        int base, stride;
                        line_of(nd, np, d, k, &base, &stride);
                        for (int ij = 0; ij < np; ij++) {
                            int j = base + ij * stride;
                            for (int ii = 0; ii < np; ii++) {
                                int i = base + ii * stride;
                                for (int v = 0; v < nv; v++)
                                    dQ[e * npts + i + ndof * v] -=
                                        P->Dsharp[ii + np * ij] * Fs[(ii * nv + v) * npts + j];
                            }
                        }
                    }
                }
            }
        }
        free(Ft); free(Fs);
    }

    /* applyBCs!   Interfaces.jl:25-49 */
    for (int ib = 0; ib < P->nbound; ib++) {
        int kind = P->bc_kind[ib];
        #pragma omp parallel for schedule(static)
        for (int64_t m = P->bc_offsets[ib]; m < P->bc_offsets[ib + 1]; m++) {
            int64_t f = P->bc_faces[m] - 1;
            for (int i = 0; i < nfp; i++) {
                double Qi[MAXV], Qe[MAXV];
                for (int v = 0; v < nv; v++) Qi[v] = Qf0[f * nfp + i + nfd * v];
                if (kind == BC_INFLOW) {
                    for (int v = 0; v < nv; v++) Qe[v] = P->bc_state[ib * nv + v];
                } else if (kind == BC_OUTFLOW) {
                    for (int v = 0; v < nv; v++) Qe[v] = Qi[v];
                } else if (kind == BC_SLIP) {   /* Equations/Euler.jl:88-94 */
                    double R[MAXV];
                    const double *fr = P->frames + (f * nfp + i) * 3 * nd;
                    rotate2face(P, Qi, fr, R);
                    R[1] = -R[1];
                    rotate2phys(P, R, fr, Qe);
                } else {
                    for (int v = 0; v < nv; v++) Qe[v] = P->bc_table[(m * nfp + i) * nv + v];
                }
                for (int v = 0; v < nv; v++) Qf1[f * nfp + i + nfd * v] = Qe[v];
            }
        }
    }

    /* interface_fluxes!   Interfaces.jl:111-136 */
    #pragma omp parallel for schedule(static)
    for (int64_t f = 0; f < nf; f++) {
        int o = P->orientation[f];
        for (int i = 0; i < nfp; i++) {
            int j = master2slave(nd, np, i, o);
            const double *fr = P->frames + (f * nfp + i) * 3 * nd;
            double Ql[MAXV], Qr[MAXV], Qln[MAXV], Qrn[MAXV], Fni[MAXV], Fp[MAXV];
            for (int v = 0; v < nv; v++) {
                Ql[v] = Qf0[f * nfp + i + nfd * v];
                Qr[v] = Qf1[f * nfp + j + nfd * v];
            }
            rotate2face(P, Ql, fr, Qln);
            rotate2face(P, Qr, fr, Qrn);
            numericalflux(P, Qln, Qrn, fr, Fni);
            rotate2phys(P, Fni, fr, Fp);
            for (int v = 0; v < nv; v++) {
                double x = Fp[v] * P->fjac[f * nfp + i];
                Fn0[f * nfp + i + nfd * v] = x;
                Fn1[f * nfp + j + nfd * v] = -x;
            }
        }
    }

    /* surface_contribution!   OpDivergence.jl:42-100; split form on nodes without boundaries:
     * _splitdiv_nb_surface_contribution!   OpDivergence.jl:300-437 */
    const int split_nb = (P->op == OP_SPLIT && !P->hasboundaries);
    const int hybrid_nb = (P->op == OP_HYBRID && !P->hasboundaries);
    #pragma omp parallel for schedule(static)
    for (int64_t e = 0; e < ne; e++) {
        const int64_t *faces = P->faceinds + e * 2 * nd, *sides = P->facepos + e * 2 * nd;
        for (int d = 0; d < nd; d++) {
            const double *FL = sides[2 * d] == 1 ? Fn0 : Fn1;
            const double *FR = sides[2 * d + 1] == 1 ? Fn0 : Fn1;
            int64_t fl = (faces[2 * d] - 1) * nfp, fr = (faces[2 * d + 1] - 1) * nfp;
            for (int k = 0; k < nlines && split_nb; k++) {
                /* _flux_splitdiv_nb_tensorproduct! + _surf_splitdiv_nb_tensorproduct! for one row */
                int base, stride;
                line_of(nd, np, d, k, &base, &stride);
                const double *Ja = P->metric + (int64_t)e * npts * nd * nd;
                double nl[3] = {0, 0, 0}, nr[3] = {0, 0, 0}, Wl[MAXV], Wr[MAXV], Qa[MAXV], Qb[MAXV];
                double Fl[MAXNP][MAXV], Fr[MAXNP][MAXV], lFl[MAXV], rFr[MAXV];
                const int64_t sg0 = (((int64_t)e * nd + d) * nlines + k) * (np + 1);
                for (int c = 0; c < nd; c++) {            /* frames[dir][i1].n * Js[dir][i1], i2 */
                    nl[c] = P->sub_frames[sg0 * 3 * nd + c] * P->sub_fjac[sg0];
                    nr[c] = P->sub_frames[(sg0 + np) * 3 * nd + c] * P->sub_fjac[sg0 + np];
                }
                for (int v = 0; v < nv; v++) { Wl[v] = 0; Wr[v] = 0; lFl[v] = 0; rFr[v] = 0; }
                for (int ii = 0; ii < np; ii++) {
                    double Qi[MAXV], Wi[MAXV];
                    for (int v = 0; v < nv; v++) Qi[v] = Q[e * npts + base + ii * stride + ndof * v];
                    cons2entropy(Qi, nd, P->gamma, Wi);
                    for (int v = 0; v < nv; v++) { Wl[v] += P->lm[ii] * Wi[v]; Wr[v] += P->lp[ii] * Wi[v]; }
                }
                entropy2cons(Wl, nd, P->gamma, Qa);
                entropy2cons(Wr, nd, P->gamma, Qb);
                for (int ii = 0; ii < np; ii++) {
                    int i = base + ii * stride;
                    double Qi[MAXV];
                    for (int v = 0; v < nv; v++) Qi[v] = Q[e * npts + i + ndof * v];
                    twopointflux(P, Qi, Qa, Ja + i * nd * nd + nd * d, nl, Fl[ii]);
                    twopointflux(P, Qi, Qb, Ja + i * nd * nd + nd * d, nr, Fr[ii]);
                    for (int v = 0; v < nv; v++) { lFl[v] += P->lm[ii] * Fl[ii][v]; rFr[v] += P->lp[ii] * Fr[ii][v]; }
                }
                for (int ii = 0; ii < np; ii++)
                    for (int v = 0; v < nv; v++) {
                        double a = Fl[ii][v] - (lFl[v] + FL[fl + k + nfd * v]);
                        double b = Fr[ii][v] - (rFr[v] - FR[fr + k + nfd * v]);
                        dQ[e * npts + base + ii * stride + ndof * v] += P->dgl[ii] * a - P->dgr[ii] * b;
                    }
            }
            for (int k = 0; k < nlines && hybrid_nb; k++) {
                /* _hybrid_nb_surface_contribution!   OpDivergence.jl:629-779 for one row:
                 * _flux_splitdiv_tensorproduct! (F#), _flux_splitdiv_nb_tensorproduct! (Fl, Fr),
                 * _surf_hybrid_nb_tensorproduct! (sub-cell fluxes, FV blending, differencing) */
                int base, stride;
                line_of(nd, np, d, k, &base, &stride);
                const double *Ja = P->metric + (int64_t)e * npts * nd * nd;
                const int64_t sg0 = (((int64_t)e * nd + d) * nlines + k) * (np + 1);
                double nl[3] = {0, 0, 0}, nr[3] = {0, 0, 0}, Wl[MAXV], Wr[MAXV], Qa[MAXV], Qb[MAXV];
                double Qn[MAXNP][MAXV], Wn[MAXNP][MAXV], Fsh[MAXNP][MAXNP][MAXV];
                double Fl[MAXNP][MAXV], Fr[MAXNP][MAXV], lFl[MAXV], rFr[MAXV], Fb[MAXNP + 1][MAXV];
                for (int c = 0; c < nd; c++) {
                    nl[c] = P->sub_frames[sg0 * 3 * nd + c] * P->sub_fjac[sg0];
                    nr[c] = P->sub_frames[(sg0 + np) * 3 * nd + c] * P->sub_fjac[sg0 + np];
                }
                for (int v = 0; v < nv; v++) { Wl[v] = 0; Wr[v] = 0; lFl[v] = 0; rFr[v] = 0; }
                for (int ii = 0; ii < np; ii++) {
                    for (int v = 0; v < nv; v++) Qn[ii][v] = Q[e * npts + base + ii * stride + ndof * v];
                    cons2entropy(Qn[ii], nd, P->gamma, Wn[ii]);
                    for (int v = 0; v < nv; v++) { Wl[v] += P->lm[ii] * Wn[ii][v]; Wr[v] += P->lp[ii] * Wn[ii][v]; }
                }
                /* F#: diagonal = contravariant flux, off-diagonal = two-point flux (symmetric) */
                for (int ii = 0; ii < np; ii++) {
                    int i = base + ii * stride;
                    double F[3][MAXV];
                    volumeflux(P, Qn[ii], F);
                    const double *M = Ja + i * nd * nd;
                    for (int v = 0; v < nv; v++) {
                        double t = F[0][v] * M[0 + nd * d];
                        for (int c = 1; c < nd; c++) t += F[c][v] * M[c + nd * d];
                        Fsh[ii][ii][v] = t;
                    }
                    for (int il = ii + 1; il < np; il++) {
                        int l = base + il * stride;
                        double F2[MAXV];
                        twopointflux(P, Qn[ii], Qn[il], Ja + i * nd * nd + nd * d, Ja + l * nd * nd + nd * d, F2);
                        for (int v = 0; v < nv; v++) { Fsh[ii][il][v] = F2[v]; Fsh[il][ii][v] = F2[v]; }
                    }
                }
                entropy2cons(Wl, nd, P->gamma, Qa);
                entropy2cons(Wr, nd, P->gamma, Qb);
                for (int ii = 0; ii < np; ii++) {
                    int i = base + ii * stride;
                    twopointflux(P, Qn[ii], Qa, Ja + i * nd * nd + nd * d, nl, Fl[ii]);
                    twopointflux(P, Qn[ii], Qb, Ja + i * nd * nd + nd * d, nr, Fr[ii]);
                    for (int v = 0; v < nv; v++) { lFl[v] += P->lm[ii] * Fl[ii][v]; rFr[v] += P->lp[ii] * Fr[ii][v]; }
                }
                for (int ii = 0; ii < np; ii++)
                    for (int v = 0; v < nv; v++) {
                        Fl[ii][v] -= lFl[v] + FL[fl + k + nfd * v];
                        Fr[ii][v] -= rFr[v] - FR[fr + k + nfd * v];
                    }
                for (int v = 0; v < nv; v++) { Fb[0][v] = -FL[fl + k + nfd * v]; Fb[np][v] = FR[fr + k + nfd * v]; }
                for (int ii = 0; ii < np - 1; ii++) {
                    for (int v = 0; v < nv; v++) {
                        double t = 0;
                        for (int jj = 0; jj < np; jj++) t += P->Dsharp[ii + np * jj] * Fsh[jj][ii][v];
                        Fb[ii + 1][v] = Fb[ii][v] + t * P->w1d[ii] - P->lm[ii] * Fl[ii][v] + P->lp[ii] * Fr[ii][v];
                    }
                    const double *fr_ = P->sub_frames + (sg0 + ii + 1) * 3 * nd;
                    double Qln[MAXV], Qrn[MAXV], Fn_[MAXV], Fv[MAXV], b = 0;
                    rotate2face(P, Qn[ii], fr_, Qln);
                    rotate2face(P, Qn[ii + 1], fr_, Qrn);
                    numericalflux(P, Qln, Qrn, fr_, Fn_);
                    rotate2phys(P, Fn_, fr_, Fv);
                    for (int v = 0; v < nv; v++) Fv[v] *= P->sub_fjac[sg0 + ii + 1];
                    for (int v = 0; v < nv; v++) b += (Wn[ii + 1][v] - Wn[ii][v]) * (Fb[ii + 1][v] - Fv[v]);
                    double delta = sqrt(b * b + P->blend);
                    delta = (delta - b) / delta;
                    delta = fmax(delta, 0.5);
                    for (int v = 0; v < nv; v++) Fb[ii + 1][v] = (1 - delta) * Fv[v] + delta * Fb[ii + 1][v];
                }
                for (int ii = 0; ii < np; ii++)
                    for (int v = 0; v < nv; v++)
                        dQ[e * npts + base + ii * stride + ndof * v] += (Fb[ii][v] - Fb[ii + 1][v]) / P->w1d[ii];
            }
            for (int k = 0; k < nlines && !split_nb && !hybrid_nb; k++) {
                int base, stride;
                line_of(nd, np, d, k, &base, &stride);
                for (int ii = 0; ii < np; ii++)
                    for (int v = 0; v < nv; v++)
                        dQ[e * npts + base + ii * stride + ndof * v] -=
                            P->dgl[ii] * FL[fl + k + nfd * v] + P->dgr[ii] * FR[fr + k + nfd * v];
            }
        }
    }

    /* apply_massmatrix!   MultielementDiscontinuous.jl:132-137 (Pover = I -> Diagonal ldiv!) */
    #pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < ndof; i++)
        for (int v = 0; v < nv; v++) dQ[i + ndof * v] = dQ[i + ndof * v] / P->jac[i];

    /* apply_sourceterm!: default closure does nothing (MultielementDiscontinuous.jl:75-79) */
}

/* 2N low-storage RK (OrdinaryDiffEq LowStorageRK2N, williamson_condition=false):
 *   stage 1 : k = f(u, t);            tmp = dt*k;             u = u + B1*tmp
 *   stage s : k = f(u, t + c_s dt);   tmp = A_s*tmp + dt*k;   u = u + B_s*tmp
 * A[0] is ignored (stage 1), A[s], B[s], c[s] for s = 0..nstages-1. `@muladd` in the
 * third-party source allows FMA contraction; fma() is used here. */
void oracle_lsrk2n(const oracle_problem *P, double *u, double *k, double *tmp,
                   int nstages, const double *A, const double *B, const double *c,
                   double dt, double t0, int64_t nsteps)
{
    const int npts = ipow(P->np, P->nd);
    const int64_t n = P->ne * npts * P->nv;
    double t = t0;
    for (int64_t it = 0; it < nsteps; it++) {
        for (int s = 0; s < nstages; s++) {
            oracle_rhs(P, u, k, t + c[s] * dt);
            if (s == 0) {
                #pragma omp parallel for schedule(static)
                for (int64_t i = 0; i < n; i++) { tmp[i] = dt * k[i]; u[i] = fma(B[0], tmp[i], u[i]); }
            } else {
                const double a = A[s], b = B[s];
                #pragma omp parallel for schedule(static)
                for (int64_t i = 0; i < n; i++) {
                    tmp[i] = fma(dt, k[i], a * tmp[i]);
                    u[i] = fma(b, tmp[i], u[i]);
                }
            }
        }
        t = t0 + (double)(it + 1) * dt;
    }
}

/* get_max_dt(q, disc, eq, cfl)   MultielementDiscontinuous.jl:162-178 (serial loop) with
 * FlouCommon/Euler.jl:116-135 and LinearAdvection.jl:46-48; `volume` = per-element sum(J*w),
 * PhysicalRegions.jl:403-405. */
double oracle_max_dt(const oracle_problem *P, const double *Q, const double *volume, double cfl)
{
    const int nd = P->nd, npts = ipow(P->np, nd);
    const int64_t ndof = P->ne * npts;
    double dt = INFINITY;
    for (int64_t e = 0; e < P->ne; e++) {
        double dx = volume[e] / npts;
        dx = nd == 1 ? dx : (nd == 2 ? sqrt(dx) : cbrt(dx));
        for (int i = 0; i < npts; i++) {
            double v;
            if (P->equation == EQ_ADVECTION) {
                double a2 = 0;
                for (int d = 0; d < nd; d++) a2 += P->a[d] * P->a[d];
                v = cfl * dx / sqrt(a2);
            } else {
                double Qi[MAXV];
                for (int k = 0; k < P->nv; k++) Qi[k] = Q[e * npts + i + ndof * k];
                double c = sqrt(P->gamma * pressure(Qi, nd, P->gamma) / Qi[0]);
                double s2 = 0;
                for (int d = 0; d < nd; d++) { double w = Qi[1 + d] / Qi[0]; s2 += w * w; }
                v = cfl * dx / ((nd == 1 ? fabs(Qi[1] / Qi[0]) : sqrt(s2)) + c);
            }
            dt = fmin(dt, v);
        }
    }
    return dt;
}

int oracle_sizeof_problem(void) { return (int)sizeof(oracle_problem); }

/* Thread count of the `parallel for` sweeps: set explicitly by bench.py (torchrun exports
 * OMP_NUM_THREADS=1 to its workers, which would silently make the CPU baseline single-threaded)
 * and read back for the `cores` field of the bench line. */
void oracle_set_threads(int n) { if (n > 0) omp_set_num_threads(n); }
int oracle_max_threads(void) { return omp_get_max_threads(); }
