"""TEST INFRASTRUCTURE (oracle): CPU restatement of the reference's VTKHDF snapshot content.

  open_for_write!(file, disc)      src/FlouSpatial/IO.jl:16-76     -> mesh_datasets()
  pointdata2VTKHDF(Q, disc)        src/FlouSpatial/IO.jl:78-97     -> pointdata()
  project2equispaced!              src/FlouSpatial/StdRegions/StdRegions.jl:232-238 (mul! by node2eq)
  node2eq, ξe                      StdSegment.jl:55-58 (interp_matrix at range(-1, 1, n)),
                                   StdQuad.jl:52-53, StdHex.jl:57-61 (kron of the segment's matrix)
  vtk_type / vtk_connectivities    StdSegment.jl:169-175, StdQuad.jl:183-195, StdHex.jl:172-199
  _save_function                   src/FlouTime/FlouTime.jl:67-81

Written with plain loops in the reference's order (element by element, dense Kronecker matrix);
only tests/ may import it.  Parity unpinned against HDF5.jl itself: no HDF5 library and no Julia
exist in this image, so what is compared is the CONTENT of every dataset and attribute the
reference writes (names, shapes, element types, values), read back from the product's file by
tests/hdf5_reader.py.
"""
import numpy as np

from . import geometry, operators


def equispaced_1d(nodetype, npn, nequispaced=None):
    """(ξe, node2eq) of the segment: node2eq[i, j] = l_j(ξe_i)."""
    n = npn if nequispaced is None else nequispaced
    xi, _ = operators.nodes_from_name(npn, nodetype)
    polys = operators.lagrange_monomials(xi)
    xe = np.linspace(-1.0, 1.0, n) if n > 1 else np.zeros(1)
    M = np.array([[operators._horner(polys[j], x) for j in range(npn)] for x in xe])
    return xe, M


def node2eq(nd, M):
    """kron(M, M, M): Julia's kron has the LAST factor on the fastest index, and so has numpy's."""
    out = M
    for _ in range(nd - 1):
        out = np.kron(M, out)
    return out


def vtk_connectivities(nd, n):
    def li(*idx):       # 1-based Cartesian index -> 1-based linear index, first index fastest
        r, stride = 0, 1
        for i in idx:
            r += (i - 1) * stride
            stride *= n
        return r + 1
    mid = range(2, n)
    if nd == 1:
        conns = [1, n] + list(mid)
    elif nd == 2:
        conns = [li(1, 1), li(n, 1), li(n, n), li(1, n)]
        conns += [li(i, 1) for i in mid] + [li(n, j) for j in mid]
        conns += [li(i, n) for i in mid] + [li(1, j) for j in mid]
        conns += [li(i, j) for j in mid for i in mid]
    else:
        conns = [li(1, 1, 1), li(n, 1, 1), li(n, n, 1), li(1, n, 1),
                 li(1, 1, n), li(n, 1, n), li(n, n, n), li(1, n, n)]
        for k in (1, n):
            conns += [li(i, 1, k) for i in mid] + [li(n, j, k) for j in mid]
            conns += [li(i, n, k) for i in mid] + [li(1, j, k) for j in mid]
        conns += [li(1, 1, k) for k in mid] + [li(n, 1, k) for k in mid]
        conns += [li(n, n, k) for k in mid] + [li(1, n, k) for k in mid]
        conns += [li(1, j, k) for k in mid for j in mid] + [li(n, j, k) for k in mid for j in mid]
        conns += [li(i, 1, k) for k in mid for i in mid] + [li(i, n, k) for k in mid for i in mid]
        conns += [li(i, j, 1) for j in mid for i in mid] + [li(i, j, n) for j in mid for i in mid]
        conns += [li(i, j, k) for k in mid for j in mid for i in mid]
    return [c - 1 for c in conns]


def mesh_datasets(problem, nodetype, nequispaced=None, regions=None):
    """{dataset path: array} of everything open_for_write! puts under /VTKHDF, plus the attributes."""
    mesh, nd, npn = problem.mesh, problem.nd, problem.np
    xe1, _ = equispaced_1d(nodetype, npn, nequispaced)
    xe = geometry.tensor_nodes(xe1, nd)
    neq = len(xe)
    points, conn, offsets, types, regs = [], [], [0], [], []
    vtype = {1: 68, 2: 70, 3: 72}[nd]
    base = vtk_connectivities(nd, npn)
    for ie in range(mesh.nelements):
        nodes = [np.asarray(mesh.nodes[i - 1], dtype=float) for i in mesh.enodes[ie]]
        for x in xe:
            points.extend(geometry.phys_coords(x, nodes))
            points.extend([0.0] * (3 - nd))
        conn.extend(c + offsets[-1] for c in base)
        offsets.append(neq + offsets[-1])
        types.append(vtype)
        regs.append(1 if regions is None else regions[ie])
    pts = np.array(points).reshape(-1, 3)       # Julia (3, N) column-major = HDF5 shape (N, 3)
    return {
        "/VTKHDF/NumberOfPoints": np.array([pts.shape[0]], dtype=np.int64),
        "/VTKHDF/Points": pts,
        "/VTKHDF/NumberOfConnectivityIds": np.array([len(conn)], dtype=np.int64),
        "/VTKHDF/Connectivity": np.array(conn, dtype=np.int64),
        "/VTKHDF/NumberOfCells": np.array([len(types)], dtype=np.int64),
        "/VTKHDF/Types": np.array(types, dtype=np.uint8),
        "/VTKHDF/Offsets": np.array(offsets, dtype=np.int64),
        "/VTKHDF/CellData/Region": np.array(regs, dtype=np.int64),
    }, {"Version": np.array([1, 0], dtype=np.int64), "Type": "UnstructuredGrid"}


def pointdata(problem, nodetype, Q, nequispaced=None):
    """One vector per variable, elements appended in order (IO.jl:86-96)."""
    nd, npn, npts = problem.nd, problem.np, problem.npts
    _, M1 = equispaced_1d(nodetype, npn, nequispaced)
    M = node2eq(nd, M1)
    Q = np.asarray(Q).reshape(problem.ndof, -1, order="F")
    out = [[] for _ in range(Q.shape[1])]
    for ie in range(problem.ne):
        for iv in range(Q.shape[1]):
            out[iv].extend(M @ Q[ie * npts:(ie + 1) * npts, iv])
    return [np.array(v) for v in out]


def variablenames(equation_is_euler, nd):
    if not equation_is_euler:
        return ("u",)
    return (("rho", "rhou", "rhoe"), ("rho", "rhou", "rhov", "rhoe"), ("rho", "rhou", "rhov", "rhow", "rhoe"))[nd - 1]
