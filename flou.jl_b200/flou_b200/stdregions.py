"""Host-side mirror of Flou.jl's approximation bases and standard regions.

  LagrangeBasis(nodetype, n)       src/FlouSpatial/StdRegions/ApproximationBases.jl:60-83
  DGSEMrec(basis)                  src/FlouSpatial/StdRegions/Reconstruction.jl:121-134
  StdSegment / StdQuad / StdHex    src/FlouSpatial/StdRegions/StdSegment.jl:34-119,
                                   StdQuad.jl:34-96, StdHex.jl:36-113

The 1-D tables (D, Ds, D♯, l, ∂g, ω) are produced along the reference's route -- monomial
Lagrange polynomials from a Vandermonde solve, evaluated by Horner -- because ω and l(±1)
obtained that way differ from closed forms at round-off level and the RHS parity bar is
1e-12.  In a Julia deployment these arrays come straight from Flou's own `std` object
(INTEGRATION.md); this module exists so the Python harness can drive the same C ABI.
"""
import numpy as np


def _legendre(n, x):
    """P_n and P_n' at x (extended precision where the platform has it)."""
    p0 = np.ones_like(x)
    if n == 0:
        return p0, np.zeros_like(x)
    p1 = x.copy()
    for k in range(2, n + 1):
        p0, p1 = p1, ((2 * k - 1) * x * p1 - (k - 1) * p0) / k
    return p1, n * (x * p1 - p0) / (x * x - 1)


def _gausslegendre(n):
    i = np.arange(1, n + 1, dtype=np.longdouble)
    x = np.cos(np.pi * (i - 0.25) / (n + 0.5)).astype(np.longdouble)
    for _ in range(100):
        p, dp = _legendre(n, x)
        dx = p / dp
        x = x - dx
        if np.max(np.abs(dx)) < 1e-19:
            break
    return np.sort(x).astype(np.float64)


def _gausslobatto(n):
    if n < 2:
        raise ValueError("Gauss-Lobatto nodes need n >= 2")
    m = n - 1
    out = np.empty(n, dtype=np.longdouble)
    out[0], out[-1] = -1.0, 1.0
    if n > 2:
        i = np.arange(1, m, dtype=np.longdouble)
        x = np.cos(np.pi * i / m).astype(np.longdouble)
        for _ in range(200):
            p, dp = _legendre(m, x)
            d2p = (2 * x * dp - m * (m + 1) * p) / (1 - x * x)
            dx = dp / d2p
            x = x - dx
            if np.max(np.abs(dx)) < 1e-19:
                break
        out[1:-1] = np.sort(x)
    out = out.astype(np.float64)
    if n % 2 == 1:
        out[n // 2] = 0.0
    return out


def _gausschebyshev(n):
    i = np.arange(1, n + 1, dtype=np.longdouble)
    return np.sort(np.cos((2 * i - 1) * np.pi / (2 * n))).astype(np.float64)


def _horner(c, x):
    r = 0.0
    for a in c[::-1]:
        r = r * x + a
    return r


class LagrangeBasis:
    """LagrangeBasis(nodetype, nnodes): nodetype in {:GL, :GLL, :CGL} (strings here)."""

    def __init__(self, nodetype, nnodes):
        name = str(nodetype).lstrip(":")
        if name in ("GL", "Gauss"):
            xi, self.hasboundaries, self.nname = _gausslegendre(nnodes), False, "Gauss"
        elif name in ("GLL", "GaussLobatto"):
            xi, self.hasboundaries, self.nname = _gausslobatto(nnodes), True, "Gauss-Lobatto"
        elif name in ("CGL", "ChebyshevGauss"):
            xi, self.hasboundaries, self.nname = _gausschebyshev(nnodes), False, "Chebyshev-Gauss"
        else:
            raise ValueError(f"Nodes of type {nodetype} cannot be used in Lagrange bases.")
        self.bname = "Lagrange"
        self.xi = xi
        n = nnodes
        V = np.vander(xi, n, increasing=True)
        # Polynomials.fit(xi, e_i): monomial coefficients of each Lagrange polynomial
        self.polys = [np.linalg.solve(V, np.eye(n)[:, i]) for i in range(n)]
        ints = [np.concatenate(([0.0], c / np.arange(1, n + 1))) for c in self.polys]
        self.w = np.array([_horner(c, 1.0) - _horner(c, -1.0) for c in ints])

    def nnodes(self):
        return len(self.xi)

    def interp_matrix(self, x):
        return np.array([[_horner(c, xi) for c in self.polys] for xi in x])

    def derivative_matrix(self, x):
        n = len(self.xi)
        ders = [c[1:] * np.arange(1, n) if n > 1 else np.zeros(1) for c in self.polys]
        return np.array([[_horner(c, xi) for c in ders] for xi in x])


class DGSEMrec:
    def __init__(self, basis):
        self.basis = basis

    def reconstruction(self):
        P = self.basis.interp_matrix([-1.0, 1.0])
        return (-P[0] / self.basis.w, +P[1] / self.basis.w)


class _StdRegion:
    nd = 0

    def __init__(self, solbasis, rec, nvars=1, nequispaced=None):
        if rec.basis is not solbasis:
            raise ValueError("over-integration (solbasis != flux basis) is outside the B200 hot path")
        b = rec.basis
        self.basis, self.solbasis, self.reconstruction = b, solbasis, rec
        self.np = b.nnodes()
        self.xi1d, self.w1d = b.xi, b.w
        lm, lp = b.interp_matrix([-1.0])[0], b.interp_matrix([1.0])[0]
        self.l = (lm, lp)
        self.D = b.derivative_matrix(b.xi)
        g = rec.reconstruction()
        self.dg = (-g[0], g[1])                       # StdSegment.jl:84-85
        B = np.outer(self.dg[1], lp) - np.outer(self.dg[0], lm)
        self.Ds = self.D - B
        self.Dsharp = 2 * self.D - B
        n, nd = self.np, self.nd
        grids = np.meshgrid(*([b.xi] * nd), indexing="ij")
        self.xi = np.stack([g.reshape(-1, order="F") for g in grids], axis=1)   # x fastest
        wg = np.meshgrid(*([b.w] * nd), indexing="ij")
        w = np.ones(n ** nd)
        for g_ in wg:
            w = w * g_.reshape(-1, order="F")
        self.w = w
        # equispaced nodes for output (StdSegment.jl:40,55-58): as many as solution nodes by default;
        # quads and hexes use the Kronecker products of the 1-D matrix (StdQuad.jl:52-53, StdHex.jl:57-61)
        ne_ = n if nequispaced is None else int(nequispaced)
        self.xe1d = np.linspace(-1.0, 1.0, ne_) if ne_ > 1 else np.zeros(1)
        self.node2eq1d = np.ascontiguousarray(b.interp_matrix(self.xe1d))       # [neq][np]
        ge = np.meshgrid(*([self.xe1d] * nd), indexing="ij")
        self.xe = np.stack([g.reshape(-1, order="F") for g in ge], axis=1)      # x fastest

    def nequispaced(self):
        return len(self.xe)

    def ndofs(self):
        return self.np ** self.nd

    def nfacedofs(self):
        return self.np ** (self.nd - 1)


class StdSegment(_StdRegion):
    nd = 1


class StdQuad(_StdRegion):
    nd = 2


class StdHex(_StdRegion):
    nd = 3
