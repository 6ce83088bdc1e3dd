"""Shared builders: the same case assembled twice, once for the oracle (CPU restatement)
and once for the product (Python host mirror -> C ABI -> CUDA)."""
import numpy as np

import oracle as O
from oracle import connectivity as ocn

SEED = 20230917   # reproducible random admissible state

_FLUX_O = {"std": O.FLUX_STDAVG, "lxf": O.FLUX_LXF, "cha": O.FLUX_CHANDRASEKHAR,
           "sca": O.FLUX_SCALARDISS, "mat": O.FLUX_MATRIXDISS}


def box(nd, scale=1.0):
    return [0.0] * nd, [scale * (1.0 + 0.5 * d) for d in range(nd)]


def random_state(ndof, nd, eq, gamma=1.4, seed=SEED, amp=0.5):
    """rho, p in U[1-amp, 1+amp], velocities in U[-amp, amp] (amp = 0.5: SURVEY.md 8(d); use a
    smaller amp with Gauss nodes, whose face extrapolation of white noise overshoots)."""
    rng = np.random.default_rng(seed)
    if eq == "adv":
        return np.asfortranarray(rng.uniform(-1.0, 1.0, size=(ndof, 1)))
    rho = rng.uniform(1 - amp, 1 + amp, ndof)
    vel = rng.uniform(-amp, amp, (ndof, nd))
    p = rng.uniform(1 - amp, 1 + amp, ndof)
    Q = np.zeros((ndof, nd + 2), order="F")
    Q[:, 0] = rho
    for d in range(nd):
        Q[:, 1 + d] = rho * vel[:, d]
    Q[:, nd + 1] = p / (gamma - 1) + 0.5 * rho * np.sum(vel ** 2, axis=1)
    return Q


def smooth_state(coords, nd, eq, gamma=1.4, waves=1):
    """Smooth periodic field on the unit box of `box`; `waves` (int or one per direction):
    wavelengths per box length."""
    x = coords
    s = np.ones(len(x))
    w = [waves] * nd if np.isscalar(waves) else list(waves)
    for d in range(nd):
        s = s * np.sin(2 * np.pi * w[d] * x[:, d] / (1.0 + 0.5 * d) + 0.3 * d)
    if eq == "adv":
        return np.asfortranarray((1.0 + 0.5 * s)[:, None])
    rho = 1.0 + 0.2 * s
    vel = np.stack([0.3 * np.cos(2 * np.pi * w[d] * x[:, d] / (1.0 + 0.5 * d)) * (1 + 0.1 * s)
                    for d in range(nd)], axis=1)
    p = 1.0 + 0.1 * s
    Q = np.zeros((len(x), nd + 2), order="F")
    Q[:, 0] = rho
    for d in range(nd):
        Q[:, 1 + d] = rho * vel[:, d]
    Q[:, nd + 1] = p / (gamma - 1) + 0.5 * rho * np.sum(vel ** 2, axis=1)
    return Q


def perturb(nodes, amp, seed=7):
    rng = np.random.default_rng(seed)
    return nodes + amp * rng.uniform(-1.0, 1.0, nodes.shape)


class Case:
    """One discretisation described independently of either implementation."""

    def __init__(self, nd, n, npn, nodes="GLL", eq="euler", op="split", tp=None, nf="mat",
                 avg="cha", intensity=1.0, periodic="all", bcs=None, general=False,
                 perturb_amp=0.0, gamma=1.4, a=(2.0, -1.0, 0.5), blend=1.0, box_scale=1.0):
        self.blend = blend
        self.box_scale = box_scale      # domain = box_scale x the unit box (same dx on a mesh box_scale x finer)
        self.nd, self.n, self.np, self.nodes = nd, tuple(n), npn, nodes
        self.eq, self.op, self.tp, self.nf, self.avg = eq, op, tp, nf, avg
        self.intensity, self.gamma, self.a = intensity, gamma, tuple(a[:nd])
        if periodic == "all":
            periodic = [(str(2 * d + 1), str(2 * d + 2)) for d in range(nd)]
        self.periodic = list(periodic or [])
        self.bcs = bcs or {}
        self.general = general or perturb_amp > 0
        self.perturb_amp = perturb_amp

    @property
    def amp(self):
        return 0.5 if self.nodes == "GLL" else 0.15

    def __repr__(self):
        return (f"{self.nd}D n={self.n} np={self.np} {self.nodes} {self.eq} {self.op}"
                f"{'(blend=%g)' % self.blend if self.op == 'hybrid' else ''}"
                f"{'/' + self.tp if self.tp else ''} {self.nf}({self.avg}) "
                f"{'general' if self.general else 'cart'} per={len(self.periodic)}")

    # ---------------------------------------------------------------- oracle side
    def oracle(self):
        start, finish = box(self.nd, self.box_scale)
        mesh = ocn.cartesian_mesh(start, finish, self.n)
        if self.perturb_amp > 0:
            h = min(mesh.dx)
            mesh.nodes = perturb(mesh.nodes, self.perturb_amp * h)
        ocn.apply_periodic_bcs(mesh, *self.periodic)
        g = self.gamma
        bcs = {}
        for name, (kind, param) in self.bcs.items():
            if kind == "inflow":
                bcs[name] = (O.BC_INFLOW, np.asarray(param, dtype=float))
            elif kind == "outflow":
                bcs[name] = (O.BC_OUTFLOW, None)
            elif kind == "slip":
                bcs[name] = (O.BC_SLIP, None)
            else:
                bcs[name] = (O.BC_TABLE, param)
        return O.Problem(
            mesh, self.nodes, self.np,
            O.EQ_ADVECTION if self.eq == "adv" else O.EQ_EULER,
            {"strong": O.OP_STRONG, "split": O.OP_SPLIT, "hybrid": O.OP_HYBRID}[self.op],
            _FLUX_O[self.nf], tpflux=_FLUX_O[self.tp] if self.tp else None,
            numflux_avg=_FLUX_O[self.avg], intensity=self.intensity, gamma=g, a=self.a,
            bcs=bcs, cartesian=not self.general, blend=self.blend if self.op == "hybrid" else 0.0)

    # ---------------------------------------------------------------- product side
    def product(self, rank=0, nranks=1, device=0, use_graph=True, create=True, kernel="line"):
        """kernel="line": the production path of meshes that fill a GPU (two-kernel stage,
        line-per-thread element kernel) even on these small test meshes; "auto" lets the library
        pick (fused single-kernel stage below one element group per SM)."""
        import flou_b200 as F
        start, finish = box(self.nd, self.box_scale)
        mesh = F.CartesianMesh(self.nd, start, finish, self.n)
        if self.perturb_amp > 0:
            h = min(mesh.dx)
            mesh.nodes = perturb(mesh.nodes, self.perturb_amp * h)
        mesh.apply_periodicBCs(*self.periodic)
        eq = F.LinearAdvection(*self.a) if self.eq == "adv" else F.EulerEquation(self.nd, self.gamma)
        basis = F.LagrangeBasis(self.nodes, self.np)
        rec = F.DGSEMrec(basis)
        std = {1: F.StdSegment, 2: F.StdQuad, 3: F.StdHex}[self.nd](basis, rec, eq.nv)
        avg = {"std": F.StdAverage(), "cha": F.ChandrasekharAverage()}[self.avg]
        nf = {"std": F.StdAverage(), "cha": F.ChandrasekharAverage(),
              "lxf": F.LxF(avg, self.intensity), "sca": F.ScalarDissipation(avg, self.intensity),
              "mat": F.MatrixDissipation(avg, self.intensity)}[self.nf]
        if self.op == "strong":
            op = F.StrongDivOperator(nf)
        elif self.op == "hybrid":
            tp = {"std": F.StdAverage(), "cha": F.ChandrasekharAverage()}[self.tp or self.avg]
            op = F.HybridDivOperator(tp, nf, self.blend)
        elif self.tp:
            op = F.SplitDivOperator({"std": F.StdAverage(), "cha": F.ChandrasekharAverage()}[self.tp], nf)
        elif self.nf in ("std", "cha"):
            op = F.SplitDivOperator(nf, nf)
        else:
            op = F.SplitDivOperator(nf)
        bcs = {}
        for name, (kind, param) in self.bcs.items():
            if kind == "inflow":
                bcs[name] = F.EulerInflowBC(param)
            elif kind == "outflow":
                bcs[name] = F.EulerOutflowBC()
            elif kind == "slip":
                bcs[name] = F.EulerSlipBC()
            else:
                bcs[name] = F.GenericBC(lambda Qin, x, frame, t, eq_, f=param: f(x))
        disc = F.MultielementDisc(mesh, std, eq, op, bcs, rank=rank, nranks=nranks, device=device,
                                  geometry="general" if self.general else None,
                                  use_graph=use_graph, create=create, kernel=kernel)
        return disc, eq


def relerr(a, b):
    """inf-norm error relative to max|b|."""
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def troubled_state(ndof, nd, gamma=1.4, seed=SEED, frac=0.05):
    """A random admissible state in which a few nodes are pushed below the Zhang-Shu floor:
    tiny (some negative) densities, and energies that give tiny or negative pressures."""
    rng = np.random.default_rng(seed + 1)
    Q = random_state(ndof, nd, "euler", gamma=gamma, seed=seed, amp=0.3)
    bad_r = rng.random(ndof) < frac
    Q[bad_r, 0] = rng.uniform(-2e-3, 5e-3, int(bad_r.sum()))
    bad_p = rng.random(ndof) < frac
    kin = 0.5 * np.sum(Q[:, 1:1 + nd] ** 2, axis=1) / Q[:, 0]
    Q[bad_p, nd + 1] = kin[bad_p] + rng.uniform(-1e-3, 5e-3, int(bad_p.sum())) / (gamma - 1)
    return np.asfortranarray(Q)
