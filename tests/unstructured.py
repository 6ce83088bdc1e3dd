"""Helpers for the unstructured-mesh (config 5) tests: a synthetic quad mesh written as an MSH
4.1 file (perturbed vertices, randomly rotated element node lists so that both face
orientations and every master/slave position combination occur), and the paired
oracle/product builders."""
import os

import numpy as np

import oracle as O
from oracle import gmshmesh as ogm

REF_MESHES = "/root/reference/test/meshes"      # present in the build container only


def synthetic_raw(nx=5, ny=4, seed=3, amp=0.18):
    import flou_b200 as F
    rng = np.random.default_rng(seed)
    xs, ys = np.linspace(0, 2.0, nx + 1), np.linspace(0, 1.0, ny + 1)
    nodes = np.array([[x, y] for y in ys for x in xs])
    hx, hy = xs[1] - xs[0], ys[1] - ys[0]
    for j in range(1, ny):
        for i in range(1, nx):
            nodes[j * (nx + 1) + i] += amp * np.array([hx, hy]) * rng.uniform(-1, 1, 2)
    nid = lambda i, j: j * (nx + 1) + i + 1
    quads = []
    for j in range(ny):
        for i in range(nx):
            q = [nid(i, j), nid(i + 1, j), nid(i + 1, j + 1), nid(i, j + 1)]
            s = int(rng.integers(0, 4))
            quads.append(q[s:] + q[:s])                 # cyclic shift keeps the orientation
    lines, ent = [], []
    for i in range(nx):
        lines.append([nid(i, 0), nid(i + 1, 0)]); ent.append(1)          # Bottom
    for j in range(ny):
        lines.append([nid(nx, j), nid(nx, j + 1)]); ent.append(2)        # Right
    for i in range(nx, 0, -1):
        lines.append([nid(i, ny), nid(i - 1, ny)]); ent.append(3)        # Top
    for j in range(ny, 0, -1):
        lines.append([nid(0, j), nid(0, j - 1)]); ent.append(4)          # Left
    groups = [("Bottom", [1]), ("Right", [2]), ("Top", [3]), ("Left", [4])]
    return F.RawMesh(nodes, quads, lines, np.arange(1, len(lines) + 1), ent, groups)


def euler_bcs(names, inflow=("Left",), outflow=("Right",), Qinf=(1.0, 0.45, 0.05, 2.8)):
    """wall/farfield boundary conditions: slip everywhere except the named in/outflow groups."""
    spec = {}
    for n in names:
        spec[n] = ("inflow", list(Qinf)) if n in inflow else ("outflow", None) if n in outflow else ("slip", None)
    return spec


def build_pair(mshfile, npn, bcs, nf="mat", avg="cha", op="split", nodes="GLL", create=True,
               refinement=1, rank=0, nranks=1, blend=0.5):
    import flou_b200 as F
    omesh = ogm.unstructured_mesh_2d(mshfile)
    obcs = {}
    for name, (kind, param) in bcs.items():
        obcs[name] = {"inflow": (O.BC_INFLOW, np.asarray(param, dtype=float) if param is not None else None),
                      "outflow": (O.BC_OUTFLOW, None), "slip": (O.BC_SLIP, None)}[kind]
    FL = {"std": O.FLUX_STDAVG, "lxf": O.FLUX_LXF, "cha": O.FLUX_CHANDRASEKHAR,
          "sca": O.FLUX_SCALARDISS, "mat": O.FLUX_MATRIXDISS}
    orc = O.Problem(omesh, nodes, npn, O.EQ_EULER,
                    {"strong": O.OP_STRONG, "split": O.OP_SPLIT, "hybrid": O.OP_HYBRID}[op],
                    FL[nf], numflux_avg=FL[avg], intensity=1.0, gamma=1.4, bcs=obcs, cartesian=False,
                    blend=blend if op == "hybrid" else 0.0)
    mesh = F.UnstructuredMesh(2, mshfile, refinement=refinement)
    eq = F.EulerEquation(2, 1.4)
    basis = F.LagrangeBasis(nodes, npn)
    std = F.StdQuad(basis, F.DGSEMrec(basis), eq.nv)
    a = {"std": F.StdAverage(), "cha": F.ChandrasekharAverage()}[avg]
    numflux = {"std": F.StdAverage(), "cha": F.ChandrasekharAverage(), "lxf": F.LxF(a, 1.0),
               "sca": F.ScalarDissipation(a, 1.0), "mat": F.MatrixDissipation(a, 1.0)}[nf]
    oper = (F.StrongDivOperator(numflux) if op == "strong" else
            F.HybridDivOperator(numflux, blend) if op == "hybrid" else F.SplitDivOperator(numflux))
    pb = {}
    for name, (kind, param) in bcs.items():
        pb[name] = {"inflow": lambda p=param: F.EulerInflowBC(p), "outflow": F.EulerOutflowBC,
                    "slip": F.EulerSlipBC}[kind]()
    # kernel="line": the production path of large meshes also on these small ones (see common.py)
    disc = F.MultielementDisc(mesh, std, eq, oper, pb, create=create, rank=rank, nranks=nranks, kernel="line")
    return orc, disc, eq
