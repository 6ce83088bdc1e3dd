"""Host-side mirror of Flou.jl's equation, numerical-flux, operator and BC types.

Same names and argument meaning as the reference so user scripts read alike:
  LinearAdvection(a...)            src/FlouCommon/LinearAdvection.jl:16-26
  EulerEquation{ND}(gamma)         src/FlouCommon/Euler.jl:16-24
  StdAverage, LxF                  src/FlouSpatial/Interfaces.jl:16-23
  ChandrasekharAverage, ScalarDissipation, MatrixDissipation
                                   src/FlouSpatial/Equations/Euler.jl:167,228-231,303-306
  StrongDivOperator, SplitDivOperator   src/FlouSpatial/Equations/OpDivergence.jl:105,184-194
  HybridDivOperator                src/FlouSpatial/Equations/OpDivergence.jl:452-477
  EulerInflowBC/OutflowBC/SlipBC   src/FlouSpatial/Equations/Euler.jl:69-94
  GenericBC                        src/FlouSpatial/FlouSpatial.jl:85-91
These objects only carry parameters; all arithmetic happens in the CUDA library.
"""
from dataclasses import dataclass
from typing import Callable, Optional, Sequence

import numpy as np

from . import _lib as L


# ------------------------------------------------------------------ equations
class LinearAdvection:
    def __init__(self, *velocity):
        if not 1 <= len(velocity) <= 3:
            raise ValueError("Linear advection is implemented in 1D, 2D and 3D.")
        self.a = tuple(float(v) for v in velocity)
        self.nd = len(velocity)
        self.nv = 1
        self.kind = L.EQ_LINEAR_ADVECTION

    def variablenames(self, unicode=False):
        return ("u",)


class EulerEquation:
    def __init__(self, nd, gamma):
        if not 1 <= nd <= 3:
            raise ValueError("The Euler equations are only implemented in 1D, 2D and 3D.")
        self.nd = int(nd)
        self.nv = self.nd + 2
        self.gamma = float(gamma)
        self.kind = L.EQ_EULER

    def variablenames(self, unicode=False):
        names = (("ρ", "ρu", "ρv", "ρw") if unicode else ("rho", "rhou", "rhov", "rhow"))[:self.nd + 1]
        return names + (("ρe",) if unicode else ("rhoe",))


def nvariables(eq):
    return eq.nv


def spatialdim(eq):
    return eq.nd


def vars_prim2cons(P, eq):
    """src/FlouCommon/Euler.jl:255-271 (host helper for initial/boundary data)."""
    P = np.asarray(P, dtype=np.float64)
    nd = eq.nd
    rho, vel, p = P[0], P[1:1 + nd], P[nd + 1]
    rhoe = p / (eq.gamma - 1) + rho * float(np.sum(vel * vel)) / 2
    return np.concatenate(([rho], rho * vel, [rhoe]))


def soundvelocity(rho, p, eq):
    return float(np.sqrt(eq.gamma * p / rho))


def normal_shockwave(rho0, u0, p0, eq):
    """src/FlouCommon/Euler.jl:337-358."""
    g = eq.gamma
    a = soundvelocity(rho0, p0, eq)
    M0 = u0 / a
    rho1 = rho0 * M0 ** 2 * (g + 1) / ((g - 1) * M0 ** 2 + 2)
    p1 = p0 * (2 * g * M0 ** 2 - (g - 1)) / (g + 1)
    M1 = np.sqrt(((g - 1) * M0 ** 2 + 2) / (2 * g * M0 ** 2 - (g - 1)))
    return rho1, M1 * soundvelocity(rho1, p1, eq), p1


def gaussian_bump(*args):
    """src/FlouCommon/Utilities.jl:16-32: (x, x0, sx, h), (x, y, x0, y0, sx, sy, h) or 3-D."""
    n = (len(args) - 1) // 3
    x, x0, s, h = args[:n], args[n:2 * n], args[2 * n:3 * n], args[-1]
    e = 0.0
    for xi, x0i, si in zip(x, x0, s):
        e = e - (xi - x0i) ** 2 / (2 * si ** 2)
    return h * np.exp(e)


# ------------------------------------------------------------------ numerical fluxes
@dataclass(frozen=True)
class StdAverage:
    kind: int = L.FLUX_STDAVERAGE


@dataclass(frozen=True)
class ChandrasekharAverage:
    kind: int = L.FLUX_CHANDRASEKHAR


@dataclass(frozen=True)
class LxF:
    avg: object
    intensity: float
    kind: int = L.FLUX_LXF


@dataclass(frozen=True)
class ScalarDissipation:
    avg: object
    intensity: float
    kind: int = L.FLUX_SCALARDISSIPATION


@dataclass(frozen=True)
class MatrixDissipation:
    avg: object
    intensity: float
    kind: int = L.FLUX_MATRIXDISSIPATION


# ------------------------------------------------------------------ divergence operators
class StrongDivOperator:
    def __init__(self, numflux):
        self.numflux = numflux
        self.tpflux = None
        self.kind = L.OP_STRONG


class SplitDivOperator:
    """SplitDivOperator([tpflux=numflux.avg], numflux)  (OpDivergence.jl:184-194)."""

    def __init__(self, *args):
        if len(args) == 1:
            (numflux,) = args
            tpflux = numflux.avg
        elif len(args) == 2:
            tpflux, numflux = args
        else:
            raise TypeError("SplitDivOperator([tpflux], numflux)")
        if not isinstance(tpflux, (StdAverage, ChandrasekharAverage)):
            raise ValueError("the two-point flux must be StdAverage or ChandrasekharAverage")
        self.tpflux, self.numflux = tpflux, numflux
        self.kind = L.OP_SPLIT


class HybridDivOperator:
    """HybridDivOperator([tpflux=numflux.avg], numflux, blend)  (OpDivergence.jl:452-477):
    split form in telescopic (flux-differencing) form, blended sub-cell interface by sub-cell
    interface with finite-volume fluxes.  As in both of the reference's convenience constructors
    `fvflux = numflux`; a different `fvflux` (raw four-field constructor) is not offered."""

    def __init__(self, *args):
        if len(args) == 2:
            numflux, blend = args
            if not hasattr(numflux, "avg"):
                raise TypeError("HybridDivOperator(numflux, blend) needs a flux with an `.avg` "
                                "(the reference reads numflux.avg)")
            tpflux = numflux.avg
        elif len(args) == 3:
            tpflux, numflux, blend = args
        else:
            raise TypeError("HybridDivOperator([tpflux], numflux, blend)")
        if not isinstance(tpflux, (StdAverage, ChandrasekharAverage)):
            raise ValueError("the two-point flux must be StdAverage or ChandrasekharAverage")
        self.tpflux, self.fvflux, self.numflux = tpflux, numflux, numflux
        self.blend = float(blend)
        self.kind = L.OP_HYBRID


# ------------------------------------------------------------------ boundary conditions
class EulerInflowBC:
    def __init__(self, Qext: Sequence[float]):
        if not 3 <= len(Qext) <= 5:
            raise ValueError("`Qext` must have a length of 3, 4 or 5.")
        self.Qext = np.asarray(Qext, dtype=np.float64)
        self.kind = L.BC_INFLOW


class EulerOutflowBC:
    kind = L.BC_OUTFLOW


class EulerSlipBC:
    kind = L.BC_SLIP


class GenericBC:
    """GenericBC(Qext) with `Qext(Qin, x, frame, time, eq)` (FlouSpatial.jl:85-91, called from
    applyBC!, Interfaces.jl:44-48).

    The closure is host code, so the exterior state is TABULATED per boundary-face node and the
    device reads the table.  A closure that depends on the position `x` alone (every use in the
    reference's tests, test/tests.jl:103-113,152-159) is tabulated once when the discretisation is
    built.  One that reads `Qin`, `frame` or `time` is detected at construction (`dynamic=None`)
    or declared (`dynamic=True`) and re-tabulated by the host before every RK stage from the
    interior traces of the device-resident state; `frame` has fields n, t, b.  `dynamic=False`
    insists on the static table and raises if the closure reads anything else.
    """
    kind = L.BC_TABLE

    def __init__(self, Qext: Callable, dynamic=None):
        self.Qext = Qext
        self.dynamic = dynamic

    def tabulate(self, x, eq):
        return np.asarray(self.Qext(_Forbidden("Qin"), x, _Forbidden("frame"), _Forbidden("time"), eq),
                          dtype=np.float64)

    def evaluate(self, Qin, x, frame, time, eq):
        return np.asarray(self.Qext(Qin, x, frame, time, eq), dtype=np.float64)


class Frame:
    """geometry.faces.frames[i]: unit normal, tangent and bi-tangent of a face node."""
    __slots__ = ("n", "t", "b")

    def __init__(self, n, t, b):
        self.n, self.t, self.b = n, t, b


class _DependsOn(ValueError):
    pass


class _Forbidden:
    def __init__(self, what):
        self._what = what

    def _raise(self, *a, **k):
        raise _DependsOn(
            f"GenericBC closure reads `{self._what}`: not a position-only boundary condition "
            f"(pass dynamic=True, or drop dynamic=False, to have it re-tabulated every stage)")

    __getitem__ = __iter__ = __float__ = __add__ = __radd__ = __mul__ = __rmul__ = _raise
    __sub__ = __rsub__ = __neg__ = __lt__ = __gt__ = __le__ = __ge__ = __len__ = _raise
    __truediv__ = __rtruediv__ = __array__ = __getattr__ = _raise


class Source:
    """Source term of `MultielementDisc(mesh, std, eq, operators, bcs, source)`
    (MultielementDiscontinuous.jl:29-36, applied by apply_sourceterm!, :139-146, after the mass
    matrix).  `f(Q_i, x_i, t)` returns the increment of dQ at one node (length nv) -- the
    reference's closure receives `dQ.dofs[i]` to add it to; with `vectorized=True` it is called
    once with (Q (n, nv), x (n, nd), t) and returns (n, nv).  The source is tabulated per node by
    the host and added on the device; `state` / `time` say what it depends on: a position-only
    source (both False) is tabulated once and costs nothing afterwards, the others are
    re-tabulated before every RK stage (`state=True` downloads the state for it)."""

    def __init__(self, f, state=True, time=True, vectorized=False):
        self.f, self.state, self.time, self.vectorized = f, bool(state), bool(time), bool(vectorized)

    @property
    def dynamic(self):
        return self.state or self.time

    def tabulate(self, Q, x, t, nv):
        if self.vectorized:
            return np.asfortranarray(np.asarray(self.f(Q, x, t), dtype=np.float64).reshape(x.shape[0], nv))
        out = np.zeros((x.shape[0], nv), order="F")
        for i in range(x.shape[0]):
            r = self.f(None if Q is None else Q[i], x[i], t)
            if r is not None:
                out[i] = r
        return out
