"""ORACLE (test infrastructure, never shipped, never on the product path).

CPU restatement of Flou.jl's 1-D nodal operators, following the reference's own
construction route (monomial Lagrange polynomials, not analytic formulas):

* nodes                   src/FlouSpatial/StdRegions/ApproximationBases.jl:135-152
                          (FastGaussQuadrature v0.5.0 `gausslegendre/gausslobatto/
                          gausschebyshev`; third-party, absent from /root/reference --
                          restated as correctly-rounded roots computed with mpmath)
* Lagrange polynomials    ApproximationBases.jl:154-165 (`Polynomials.fit(xi, e_i)`,
                          Polynomials v3.2.7: square Vandermonde `vand \\ y`, i.e. LU)
* D[i,j] = l_j'(xi_i)     ApproximationBases.jl:85-91
* w_i = int l_i           ApproximationBases.jl:179-186
* l(-1), l(+1)            StdSegment.jl:76-80
* dg = (l-/w, l+/w)       Reconstruction.jl:121-134 + sign flip StdSegment.jl:84-85
* B, Ds = D-B, Dsharp = 2D-B   StdSegment.jl:87-89
"""
import numpy as np
import mpmath as mp


def _legendre_and_derivs(n, x):
    """P_n(x), P_n'(x) by the three-term recurrence (mpmath precision)."""
    p0, p1 = mp.mpf(1), x
    if n == 0:
        return p0, mp.mpf(0)
    for k in range(2, n + 1):
        p0, p1 = p1, ((2 * k - 1) * x * p1 - (k - 1) * p0) / k
    dp = n * (x * p1 - p0) / (x * x - 1)
    return p1, dp


def gausslegendre_nodes(n):
    """Roots of P_n, ascending (FastGaussQuadrature.gausslegendre nodes)."""
    mp.mp.dps = 50
    xs = []
    for i in range(1, n + 1):
        x = mp.cos(mp.pi * (i - mp.mpf(1) / 4) / (n + mp.mpf(1) / 2))
        for _ in range(100):
            p, dp = _legendre_and_derivs(n, x)
            dx = p / dp
            x -= dx
            if abs(dx) < mp.mpf(10) ** (-45):
                break
        xs.append(x)
    xs.sort()
    return np.array([float(x) for x in xs])


def gausslobatto_nodes(n):
    """+-1 and the roots of P'_{n-1}, ascending (FastGaussQuadrature.gausslobatto)."""
    mp.mp.dps = 50
    if n < 2:
        raise ValueError("Gauss-Lobatto needs n >= 2")
    m = n - 1
    xs = [mp.mpf(-1), mp.mpf(1)]
    # interior nodes: roots of P'_m ; Newton on q(x) = P'_m(x) with
    # q'(x) = (2x P'_m - m(m+1) P_m)/(1-x^2)
    for i in range(1, m):
        x = mp.cos(mp.pi * i / m)
        for _ in range(200):
            p, dp = _legendre_and_derivs(m, x)
            d2p = (2 * x * dp - m * (m + 1) * p) / (1 - x * x)
            dx = dp / d2p
            x -= dx
            if abs(dx) < mp.mpf(10) ** (-45):
                break
        xs.append(x)
    xs.sort()
    out = np.array([float(x) for x in xs])
    if n % 2 == 1:
        out[n // 2] = 0.0
    return out


def gausschebyshev_nodes(n):
    mp.mp.dps = 50
    xs = sorted(mp.cos((2 * i - 1) * mp.pi / (2 * n)) for i in range(1, n + 1))
    return np.array([float(x) for x in xs])


def nodes_from_name(n, name):
    """ApproximationBases.jl:135-152 -> (xi, hasboundaries)."""
    if name in ("GL", "Gauss"):
        return gausslegendre_nodes(n), False
    if name in ("GLL", "GaussLobatto"):
        return gausslobatto_nodes(n), True
    if name in ("CGL", "ChebyshevGauss"):
        return gausschebyshev_nodes(n), False
    raise ValueError(f"Nodes of type {name} cannot be used in Lagrange bases.")


def _horner(c, x):
    r = 0.0
    for a in c[::-1]:
        r = r * x + a
    return r


def lagrange_monomials(xi):
    """Polynomials.fit(xi, e_i): coefficient rows, ascending powers."""
    n = len(xi)
    V = np.vander(xi, n, increasing=True)
    return [np.linalg.solve(V, np.eye(n)[:, i]) for i in range(n)]


def operators_1d(nodetype, n):
    """All 1-D tables the hot path consumes, built the reference's way."""
    xi, bounds = nodes_from_name(n, nodetype)
    polys = lagrange_monomials(xi)
    dpolys = [np.array([k * c[k] for k in range(1, n)]) if n > 1 else np.zeros(1)
              for c in polys]
    ipolys = [np.concatenate(([0.0], [c[k] / (k + 1) for k in range(n)])) for c in polys]
    w = np.array([_horner(ic, 1.0) - _horner(ic, -1.0) for ic in ipolys])
    D = np.array([[_horner(dpolys[j], xi[i]) for j in range(n)] for i in range(n)])
    lm = np.array([_horner(polys[j], -1.0) for j in range(n)])
    lp = np.array([_horner(polys[j], +1.0) for j in range(n)])
    # DGSEMrec gives (-l-/w, +l+/w); StdSegment flips the first one
    dgl = -(-lm / w)
    dgr = lp / w
    B = np.outer(dgr, lp) - np.outer(dgl, lm)
    Ds = D - B
    Dsharp = 2 * D - B
    return dict(xi=xi, w=w, D=D, Ds=Ds, Dsharp=Dsharp, lm=lm, lp=lp, dgl=dgl, dgr=dgr,
                hasboundaries=bounds, np=n)
