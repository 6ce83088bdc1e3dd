"""ORACLE -- test infrastructure only.

CPU restatement of Flou.jl's DGSEM `rhs!` + 2N low-storage RK loop (see pipeline.c for the
file:line map).  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline`
/ `--impl reference` legs may import this package; the product (`flou.jl_b200/`) never
does, and fails loudly when its CUDA library is missing.

Parity pinning: checked against ALL of the reference's own known-answer tests
(`test/runtests.jl:24-44`: SodTube1D min/max and Shockwave2D max to rtol 1e-7 -- reproduced to
~1e-13 -- and the Advection1D/2D periodic returns) in `tests/test_oracle_kat.py`.  Third-party pieces restated from published algorithms:
OrdinaryDiffEq v6.49.1 `ORK256` / `CarpenterKennedy2N54` 2N tableaus (ORK256 pinned through
the Sod KAT; CarpenterKennedy2N54 parity unpinned), FastGaussQuadrature v0.5.0 nodes,
Polynomials v3.2.7 `fit`.  Nothing in the reference's tests pins 3-D or unstructured
results: for those this restatement is the only ground truth.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from . import connectivity, geometry, operators

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

EQ_ADVECTION, EQ_EULER = 0, 1
OP_STRONG, OP_SPLIT, OP_HYBRID = 0, 1, 2
FLUX_STDAVG, FLUX_LXF, FLUX_CHANDRASEKHAR, FLUX_SCALARDISS, FLUX_MATRIXDISS = range(5)
BC_INFLOW, BC_OUTFLOW, BC_SLIP, BC_TABLE = range(4)

# OrdinaryDiffEq v6.49.1 low-storage 2N tableaus (third-party; call site FlouTime.jl:34-38,
# solver objects test/tests.jl:19,56,93,139).  A[0] is unused (first stage).
ORK256 = dict(
    A=[0.0, -1.0, -1.55798, -1.0, -0.45031],
    B=[0.2, 0.83204, 0.6, 0.35394, 0.2],
    c=[0.0, 0.2, 0.2, 0.8, 0.8],
)
CARPENTER_KENNEDY_2N54 = dict(   # Carpenter & Kennedy (1994), NASA TM-109112
    A=[0.0, -567301805773 / 1357537059087, -2404267990393 / 2016746695238,
       -3550918686646 / 2091501179385, -1275806237668 / 842570457699],
    B=[1432997174477 / 9575080441755, 5161836677717 / 13612068292357,
       1720146321549 / 2090206949498, 3134564353537 / 4481467310338,
       2277821191437 / 14882151754819],
    c=[0.0, 1432997174477 / 9575080441755, 2526269341429 / 6820363962896,
       2006345519317 / 3224310063776, 2802321613138 / 2924317926251],
)


class _Problem(C.Structure):
    _fields_ = [
        ("nd", C.c_int32), ("nv", C.c_int32), ("np", C.c_int32),
        ("equation", C.c_int32), ("op", C.c_int32), ("tpflux", C.c_int32),
        ("numflux", C.c_int32), ("numflux_avg", C.c_int32),
        ("intensity", C.c_double), ("gamma", C.c_double), ("a", C.c_double * 3),
        ("ne", C.c_int64), ("nf", C.c_int64),
        ("faceinds", C.c_void_p), ("facepos", C.c_void_p),
        ("eleminds", C.c_void_p), ("elempos", C.c_void_p), ("orientation", C.c_void_p),
        ("D", C.c_void_p), ("Ds", C.c_void_p), ("Dsharp", C.c_void_p),
        ("lm", C.c_void_p), ("lp", C.c_void_p), ("dgl", C.c_void_p), ("dgr", C.c_void_p),
        ("jac", C.c_void_p), ("metric", C.c_void_p), ("fjac", C.c_void_p),
        ("frames", C.c_void_p),
        ("nbound", C.c_int32), ("bc_kind", C.c_void_p), ("bc_offsets", C.c_void_p),
        ("bc_faces", C.c_void_p), ("bc_state", C.c_void_p), ("bc_table", C.c_void_p),
        ("Qf", C.c_void_p * 2), ("Fn", C.c_void_p * 2),
        ("blend", C.c_double), ("w1d", C.c_void_p), ("sub_jac", C.c_double * 3),
        ("sub_frames", C.c_void_p), ("sub_fjac", C.c_void_p),
        ("hasboundaries", C.c_int32), ("proj_traces", C.c_int32),
    ]


def build(force=False):
    """Compile pipeline.c -> oracle/_build/liboracle.so (gcc, OpenMP, no FP contraction)."""
    out_dir = os.path.join(_HERE, "_build")
    so = os.path.join(out_dir, "liboracle.so")
    src = os.path.join(_HERE, "pipeline.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        os.makedirs(out_dir, exist_ok=True)
        subprocess.check_call(
            ["gcc", "-O2", "-march=x86-64-v3", "-ffp-contract=off", "-fopenmp", "-fPIC",
             "-shared", "-o", so, src, "-lm"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.oracle_rhs.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double]
        _LIB.oracle_rhs.restype = None
        _LIB.oracle_lsrk2n.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_double, C.c_double, C.c_int64]
        _LIB.oracle_lsrk2n.restype = None
        _LIB.oracle_max_dt.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double]
        _LIB.oracle_max_dt.restype = C.c_double
        _LIB.oracle_set_threads.argtypes = [C.c_int]
        _LIB.oracle_set_threads.restype = None
        _LIB.oracle_max_threads.restype = C.c_int
        assert _LIB.oracle_sizeof_problem() == C.sizeof(_Problem)
    return _LIB


def set_threads(n):
    """OpenMP threads of the sweeps (bench.py: all host cores, whatever OMP_NUM_THREADS says)."""
    lib().oracle_set_threads(int(n))


def max_threads():
    return int(lib().oracle_max_threads())


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class Problem:
    """A fully assembled discretisation (the oracle's MultielementDisc + EquationConfig,
    MultielementDiscontinuous.jl:29-92)."""

    def __init__(self, mesh, nodetype, npn, equation, op, numflux, *, tpflux=None,
                 numflux_avg=FLUX_STDAVG, intensity=1.0, gamma=1.4, a=(0.0, 0.0, 0.0),
                 bcs=(), cartesian=True, blend=0.0):
        self.mesh = mesh
        nd = mesh.nd
        self.nd, self.np = nd, npn
        self.nv = 1 if equation == EQ_ADVECTION else nd + 2
        self.ops = operators.operators_1d(nodetype, npn)
        self.npts = npn ** nd
        self.nfp = npn ** (nd - 1)
        self.ne, self.nf = mesh.nelements, mesh.nfaces
        self.ndof = self.ne * self.npts
        self.coords, self.jac, self.metric = geometry.element_geometry(
            mesh, self.ops["xi"], cartesian)
        # face coordinates are only needed to tabulate GenericBC closures
        want = any(isinstance(b, tuple) and b[0] == BC_TABLE
                   for b in (bcs.values() if isinstance(bcs, dict) else bcs))
        self.fcoords, self.fjac, self.frames = geometry.face_geometry(
            mesh, self.ops["xi"], cartesian, want_coords=want or not cartesian)
        self.weights = geometry.tensor_weights(self.ops["w"], nd)
        if tpflux is None:   # SplitDivOperator(numflux) -> tpflux = numflux.avg
            tpflux = numflux_avg if numflux in (FLUX_LXF, FLUX_SCALARDISS, FLUX_MATRIXDISS) \
                else numflux
        k = self._keep = {}
        k["faceinds"] = np.ascontiguousarray(mesh.faceinds, dtype=np.int64)
        k["facepos"] = np.ascontiguousarray(mesh.facepos, dtype=np.int64)
        k["eleminds"] = np.ascontiguousarray(mesh.eleminds, dtype=np.int64)
        k["elempos"] = np.ascontiguousarray(mesh.elempos, dtype=np.int64)
        k["orientation"] = np.ascontiguousarray(mesh.orientation, dtype=np.uint8)
        for name in ("D", "Ds", "Dsharp"):
            k[name] = np.asfortranarray(self.ops[name]).ravel(order="F").copy()
        for name in ("lm", "lp", "dgl", "dgr"):
            k[name] = np.ascontiguousarray(self.ops[name])
        k["jac"] = self.jac
        # metric[i, c, d] -> flat [c + nd*d]
        k["metric"] = np.ascontiguousarray(self.metric.transpose(0, 2, 1))
        k["fjac"] = self.fjac
        k["frames"] = np.ascontiguousarray(self.frames)
        # boundary conditions, ordered like mesh.bdfaces (MultielementDiscontinuous.jl:45-51)
        nb = len(mesh.bdfaces)
        if len(bcs) != nb:
            raise ValueError("The number of BCs does not match the number of boundaries.")
        kinds, offs, faces, state, table = [], [0], [], np.zeros((max(nb, 1), self.nv)), []
        if isinstance(bcs, dict):
            ordered = [None] * nb
            for key, val in bcs.items():
                j = mesh.bdnames.index(key) + 1
                ordered[mesh.bdmap[j] - 1] = val
            bcs = ordered
        for ib, bc in enumerate(bcs):
            kind, param = bc if isinstance(bc, tuple) else (bc, None)
            kinds.append(kind)
            for f in mesh.bdfaces[ib]:
                faces.append(f)
                for i in range(self.nfp):
                    if kind == BC_TABLE:
                        table.append(np.asarray(param(self.fcoords[(f - 1) * self.nfp + i]),
                                                dtype=float))
                    else:
                        table.append(np.zeros(self.nv))
            offs.append(len(faces))
            if kind == BC_INFLOW:
                state[ib] = param
        k["bc_kind"] = np.array(kinds + [0], dtype=np.int32)
        k["bc_offsets"] = np.array(offs, dtype=np.int64)
        k["bc_faces"] = np.array(faces + [0], dtype=np.int64)
        k["bc_state"] = np.ascontiguousarray(state)
        k["bc_table"] = np.ascontiguousarray(np.array(table + [np.zeros(self.nv)]))
        nfd = self.nf * self.nfp
        for name in ("Qf0", "Qf1", "Fn0", "Fn1"):
            k[name] = np.zeros(nfd * self.nv)
        p = self.c = _Problem()
        p.nd, p.nv, p.np = nd, self.nv, npn
        p.equation, p.op, p.tpflux = equation, op, tpflux
        p.numflux, p.numflux_avg = numflux, numflux_avg
        p.intensity, p.gamma = intensity, gamma
        for d in range(3):
            p.a[d] = a[d] if d < len(a) else 0.0
        p.ne, p.nf, p.nbound = self.ne, self.nf, nb
        for name in ("faceinds", "facepos", "eleminds", "elempos", "orientation", "D", "Ds",
                     "Dsharp", "lm", "lp", "dgl", "dgr", "jac", "metric", "fjac", "frames",
                     "bc_kind", "bc_offsets", "bc_faces", "bc_state", "bc_table"):
            setattr(p, name, _ptr(k[name]))
        # HybridDivOperator (oracle only): 1-D weights, Cartesian sub-grid face Jacobians
        k["w1d"] = np.ascontiguousarray(self.ops["w"])
        p.w1d, p.blend = _ptr(k["w1d"]), float(blend)
        p.hasboundaries = 1 if self.ops["hasboundaries"] else 0
        split_nb = op == OP_SPLIT and not self.ops["hasboundaries"]
        if split_nb and equation != EQ_EULER:
            raise ValueError("the split form on Gauss nodes (entropy-projected surface term) needs the "
                             "Euler equations")
        if op == OP_HYBRID and equation != EQ_EULER:
            raise ValueError("HybridDivOperator needs entropy variables: Euler equations only")
        if op == OP_HYBRID or split_nb:
            # geometry.subgrids (PhysicalRegions.jl:28-292): Cartesian constants or, on general
            # meshes, frames / Jacobians from the mapping at the complementary-grid points
            fr, fj = geometry.subgrid_geometry(mesh, self.ops["xi"], self.ops["w"], cartesian)
            k["sub_frames"], k["sub_fjac"] = np.ascontiguousarray(fr), np.ascontiguousarray(fj)
            p.sub_frames, p.sub_fjac = _ptr(k["sub_frames"]), _ptr(k["sub_fjac"])
            if cartesian:
                dx = mesh.dx
                sub = [1.0] if nd == 1 else [dx[1] / 2, dx[0] / 2] if nd == 2 else \
                    [dx[1] * dx[2] / 4, dx[0] * dx[2] / 4, dx[0] * dx[1] / 4]
                for d_, v_ in enumerate(sub):
                    p.sub_jac[d_] = v_
        p.Qf[0], p.Qf[1] = _ptr(k["Qf0"]), _ptr(k["Qf1"])
        p.Fn[0], p.Fn[1] = _ptr(k["Fn0"]), _ptr(k["Fn1"])

    # state arrays are (ndof, nv) Fortran-ordered, i.e. Julia's Matrix(ndofs, NV)
    def new_state(self):
        return np.zeros((self.ndof, self.nv), order="F")

    def rhs(self, Q, t=0.0):
        Q = np.asfortranarray(Q, dtype=np.float64)
        dQ = np.zeros_like(Q, order="F")
        lib().oracle_rhs(C.byref(self.c), _ptr(Q), _ptr(dQ), t)
        return dQ

    def max_dt(self, Q, cfl):
        """get_max_dt(q, disc, equation, cfl) (MultielementDiscontinuous.jl:162-178)."""
        Q = np.asfortranarray(Q, dtype=np.float64)
        vol = np.ascontiguousarray((self.jac * np.tile(self.weights, self.ne)).reshape(self.ne, -1).sum(axis=1))
        return float(lib().oracle_max_dt(C.byref(self.c), _ptr(Q), _ptr(vol), float(cfl)))

    # ---- monitors and limiter; plain numpy, element by element in the
    # reference's order.  No reference test evaluates them: parity unpinned, checked by properties
    # (tests/test_oracle_properties.py).
    def _jw(self):
        return (self.jac * np.tile(self.weights, self.ne)).reshape(self.ne, self.npts)

    def _pressure(self, Q):
        """pressure(Q, eq) (FlouCommon/Euler.jl:142-155)."""
        g, nd = self.c.gamma, self.nd
        m2 = np.sum(Q[..., 1:1 + nd] ** 2, axis=-1)
        return (g - 1) * (Q[..., nd + 1] - m2 / (2 * Q[..., 0]))

    def monitor(self, Q, name):
        """kinetic_energy_monitor / entropy_monitor (src/FlouSpatial/Equations/Euler.jl:559-593):
        s += integrate(f.(Qe.dofs), geom_e) over the elements, integrate = Jw' * f
        (PhysicalRegions.jl:366-368)."""
        Q = np.asarray(Q, dtype=np.float64)
        g, nd = self.c.gamma, self.nd
        Jw = self._jw()
        Qe = Q.reshape(self.ne, self.npts, self.nv)
        s = 0.0
        for e in range(self.ne):
            q = Qe[e]
            if name in ("kinetic_energy", "energy"):
                f = np.sum(q[:, 1:1 + nd] ** 2, axis=1) / (2 * q[:, 0])         # Euler.jl:162-175
            elif name == "entropy":
                ent = np.log(self._pressure(q)) - g * np.log(q[:, 0])           # entropy :202-206
                f = -q[:, 0] * ent / (g - 1)                                     # math_entropy :213-217
            else:
                raise ValueError(f"Unknown monitor '{name}'.")
            s += float(Jw[e] @ f)
        return s

    def zhang_shu(self, Q, minval):
        """zhang_shu_limiter (src/FlouSpatial/Equations/Euler.jl:616-660), returns the limited
        copy.  Element mean = integrate(Qe.dofs, geom)/geom.volume, volume = sum(Jw)
        (PhysicalRegions.jl:402-405)."""
        Q = np.array(Q, dtype=np.float64, order="F", copy=True)
        Jw = self._jw()
        with np.errstate(divide="ignore", invalid="ignore"):
            for e in range(self.ne):
                rows = slice(e * self.npts, (e + 1) * self.npts)
                q = Q[rows]                              # view: (npts, nv)
                vol = float(np.sum(Jw[e]))
                qbar = (Jw[e] @ q) / vol
                m = min(minval, qbar[0])
                rmin = float(np.min(q[:, 0]))
                theta = abs(np.float64(qbar[0] - m) / np.float64(qbar[0] - rmin))
                if theta <= 1.0:
                    q[:, 0] = theta * (q[:, 0] - qbar[0]) + qbar[0]
                p = self._pressure(q)
                pbar = float(Jw[e] @ p) / vol
                m = min(minval, pbar)
                pmin = float(np.min(p))
                theta = abs(np.float64(pbar - m) / np.float64(pbar - pmin))
                if theta <= 1.0:
                    q[:, :] = theta * (q - qbar) + qbar
        return Q

    def lsrk2n_limited(self, Q, tableau, dt, nsteps, minval):
        """The 2N recurrence with `stage_limiter!` = Zhang-Shu after every stage
        (OrdinaryDiffEq LowStorageRK2N; usage examples/src/3D_Euler.jl:76-80)."""
        u = np.array(Q, dtype=np.float64, order="F", copy=True)
        tmp = np.zeros_like(u, order="F")
        A, B = tableau["A"], tableau["B"]
        for _ in range(nsteps):
            for s in range(len(B)):
                k = self.rhs(u)
                tmp = dt * k if s == 0 else A[s] * tmp + dt * k
                u = self.zhang_shu(u + B[s] * tmp, minval)
        return u

    def lsrk2n(self, Q, tableau, dt, nsteps, t0=0.0):
        u = np.array(Q, dtype=np.float64, order="F", copy=True)
        k = np.zeros_like(u, order="F")
        tmp = np.zeros_like(u, order="F")
        A = np.array(tableau["A"], dtype=np.float64)
        B = np.array(tableau["B"], dtype=np.float64)
        c = np.array(tableau["c"], dtype=np.float64)
        lib().oracle_lsrk2n(C.byref(self.c), _ptr(u), _ptr(k), _ptr(tmp), len(B),
                            _ptr(A), _ptr(B), _ptr(c), dt, t0, nsteps)
        return u


def vars_prim2cons(P, gamma):
    """FlouCommon/Euler.jl:255-271 (+ energy :180-196)."""
    P = np.asarray(P, dtype=float)
    nd = len(P) - 2
    rho, vel, p = P[0], P[1:1 + nd], P[-1]
    rhoe = p / (gamma - 1) + rho * float(np.sum(vel ** 2)) / 2
    return np.concatenate(([rho], rho * vel, [rhoe]))


def gaussian_bump(x, x0, s, h):
    """FlouCommon/Utilities.jl:16-32 (any dimension)."""
    x, x0, s = (np.atleast_1d(np.asarray(v, dtype=float)) for v in (x, x0, s))
    return h * np.exp(-np.sum((x - x0) ** 2 / (2 * s ** 2)))
