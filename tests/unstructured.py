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


# the 24 proper rotations of the hexahedron as permutations of gmsh's vertex numbering
_HEX_REF = np.array([[-1, -1, -1], [1, -1, -1], [1, 1, -1], [-1, 1, -1],
                     [-1, -1, 1], [1, -1, 1], [1, 1, 1], [-1, 1, 1]])


def hex_rotations():
    import itertools
    out = []
    for perm in itertools.permutations(range(3)):
        for signs in itertools.product((1, -1), repeat=3):
            R = np.zeros((3, 3), dtype=int)
            for r, (c, s) in enumerate(zip(perm, signs)):
                R[r, c] = s
            if round(np.linalg.det(R)) != 1:
                continue
            rot = _HEX_REF @ R.T                    # where the new local vertex j sits in the old frame
            out.append([int(np.nonzero((_HEX_REF == v).all(axis=1))[0][0]) for v in rot])
    return out


def synthetic_hex_raw(nx=3, ny=2, nz=3, seed=11, amp=0.12):
    """A block of nx x ny x nz hexahedra with perturbed interior vertices; every element's vertex
    list is one of the 24 rotations of the hexahedron, so that neighbouring elements meet with
    all kinds of local-face pairs and face orientations (GmshMesh.jl:343-390: codes 0..7).
    Boundary quads: six surface entities / physical groups."""
    import flou_b200 as F
    rng = np.random.default_rng(seed)
    xs, ys, zs = np.linspace(0, 1.5, nx + 1), np.linspace(0, 1.0, ny + 1), np.linspace(0, 1.2, nz + 1)
    nid = lambda i, j, k: (k * (ny + 1) + j) * (nx + 1) + i + 1
    nodes = np.array([[x, y, z] for z in zs for y in ys for x in xs])
    h = np.array([xs[1] - xs[0], ys[1] - ys[0], zs[1] - zs[0]])
    for k in range(1, nz):
        for j in range(1, ny):
            for i in range(1, nx):
                nodes[nid(i, j, k) - 1] += amp * h * rng.uniform(-1, 1, 3)
    rots = hex_rotations()
    hexes = []
    for k in range(nz):
        for j in range(ny):
            for i in range(nx):
                v = [nid(i, j, k), nid(i + 1, j, k), nid(i + 1, j + 1, k), nid(i, j + 1, k),
                     nid(i, j, k + 1), nid(i + 1, j, k + 1), nid(i + 1, j + 1, k + 1), nid(i, j + 1, k + 1)]
                r = rots[int(rng.integers(0, 24))]
                hexes.append([v[q] for q in r])
    quads, ent = [], []
    for k in range(nz):
        for j in range(ny):
            quads.append([nid(0, j, k), nid(0, j + 1, k), nid(0, j + 1, k + 1), nid(0, j, k + 1)]); ent.append(1)
    for k in range(nz):
        for j in range(ny):
            quads.append([nid(nx, j, k), nid(nx, j + 1, k), nid(nx, j + 1, k + 1), nid(nx, j, k + 1)]); ent.append(2)
    for k in range(nz):
        for i in range(nx):
            quads.append([nid(i, 0, k), nid(i + 1, 0, k), nid(i + 1, 0, k + 1), nid(i, 0, k + 1)]); ent.append(3)
    for k in range(nz):
        for i in range(nx):
            quads.append([nid(i, ny, k), nid(i + 1, ny, k), nid(i + 1, ny, k + 1), nid(i, ny, k + 1)]); ent.append(4)
    for j in range(ny):
        for i in range(nx):
            quads.append([nid(i, j, 0), nid(i + 1, j, 0), nid(i + 1, j + 1, 0), nid(i, j + 1, 0)]); ent.append(5)
    for j in range(ny):
        for i in range(nx):
            quads.append([nid(i, j, nz), nid(i + 1, j, nz), nid(i + 1, j + 1, nz), nid(i, j + 1, nz)]); ent.append(6)
    groups = [("Xm", [1]), ("Xp", [2]), ("Ym", [3]), ("Yp", [4]), ("Zm", [5]), ("Zp", [6])]
    return F.RawHexMesh(nodes, hexes, quads, ent, groups)


def build_pair_3d(raw, npn, bcs, nf="mat", avg="cha", op="split", nodes="GLL", create=True, blend=0.5):
    """Oracle (literal _facemap_3d restatement) and product (closed-form tables) on a RawHexMesh."""
    import flou_b200 as F
    omesh = ogm.unstructured_mesh_3d(raw.nodes, raw.hexes, raw.quads, raw.quad_entity, raw.groups)
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
    mesh = F.UnstructuredMesh(3, raw)
    eq = F.EulerEquation(3, 1.4)
    basis = F.LagrangeBasis(nodes, npn)
    std = F.StdHex(basis, F.DGSEMrec(basis), eq.nv)
    a = {"std": F.StdAverage(), "cha": F.ChandrasekharAverage()}[avg]
    numflux = {"std": F.StdAverage(), "cha": F.ChandrasekharAverage(), "lxf": F.LxF(a, 1.0),
               "sca": F.ScalarDissipation(a, 1.0), "mat": F.MatrixDissipation(a, 1.0)}[nf]
    oper = (F.StrongDivOperator(numflux) if op == "strong" else
            F.HybridDivOperator(numflux, blend) if op == "hybrid" else F.SplitDivOperator(numflux))
    pb = {}
    for name, (kind, param) in bcs.items():
        pb[name] = {"inflow": lambda p=param: F.EulerInflowBC(p), "outflow": F.EulerOutflowBC,
                    "slip": F.EulerSlipBC}[kind]()
    disc = F.MultielementDisc(mesh, std, eq, oper, pb, create=create, kernel="line")
    return orc, disc, eq


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
