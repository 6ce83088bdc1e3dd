"""ORACLE (test infrastructure, never shipped, never on the product path).

Restatement of Flou.jl's element metrics and face frames.

* linear mappings (segment/quad/hex)      src/FlouCommon/Mesh.jl:363-487
* Cartesian element metric                src/FlouSpatial/PhysicalRegions.jl:370-408
* unstructured element metric             PhysicalRegions.jl:438-472
* Cartesian face frames 1/2/3-D           PhysicalRegions.jl:541-696
* unstructured face frames 2-D / 3-D      PhysicalRegions.jl:797-871 / 873-971
* sub-grid points, frames, Jacobians      StdSegment.jl:60-74, StdQuad.jl:56-58,126-137,
                                          PhysicalRegions.jl:72-148 (Cartesian), 179-292 (general)
* tensor-product node order               StdQuad.jl:46-50, StdHex.jl:48-54 (x fastest)

Arrays returned (all float64, C-contiguous):
  coords  (N_dof, nd)          jac (N_dof,)          metric (N_dof, nd, nd) with
  metric[i, c, d] = Ja^d_c (the reference's SMatrix element [c, d]);
  fcoords (N_f*nfp, nd)        fjac (N_f*nfp,)       frames (N_f*nfp, 3, nd)  rows n, t, b.
"""
import numpy as np


def tensor_nodes(xi, nd):
    n = len(xi)
    if nd == 1:
        return xi.reshape(-1, 1).copy()
    if nd == 2:
        return np.array([[xi[i], xi[j]] for j in range(n) for i in range(n)])
    return np.array([[xi[i], xi[j], xi[k]] for k in range(n) for j in range(n) for i in range(n)])


def tensor_weights(w, nd):
    n = len(w)
    if nd == 1:
        return w.copy()
    if nd == 2:
        return np.array([w[i] * w[j] for j in range(n) for i in range(n)])
    return np.array([w[i] * w[j] * w[k] for k in range(n) for j in range(n) for i in range(n)])


# ---------------------------------------------------------------- mappings (Mesh.jl)
def phys_coords(xi, nodes):
    nv = len(nodes)
    if nv == 1:
        return np.array(nodes[0], dtype=float)
    if nv == 2:
        r = (xi[0] + 1) / 2
        return nodes[0] * (1 - r) + nodes[1] * r
    if nv == 4:
        r, s = (xi[0] + 1) / 2, (xi[1] + 1) / 2
        return (nodes[0] * (1 - r) * (1 - s) + nodes[1] * r * (1 - s)
                + nodes[2] * r * s + nodes[3] * (1 - r) * s)
    r, s, t = (xi[0] + 1) / 2, (xi[1] + 1) / 2, (xi[2] + 1) / 2
    return (nodes[0] * (1 - r) * (1 - s) * (1 - t) + nodes[1] * r * (1 - s) * (1 - t)
            + nodes[2] * r * s * (1 - t) + nodes[3] * (1 - r) * s * (1 - t)
            + nodes[4] * (1 - r) * (1 - s) * t + nodes[5] * r * (1 - s) * t
            + nodes[6] * r * s * t + nodes[7] * (1 - r) * s * t)


def map_basis(xi, nodes):
    nv = len(nodes)
    if nv == 2:
        return ((nodes[1] - nodes[0]) / 2,)
    if nv == 4:
        r, s = (xi[0] + 1) / 2, (xi[1] + 1) / 2
        dxi = (nodes[1] - nodes[0]) / 2 * (1 - s) + (nodes[2] - nodes[3]) / 2 * s
        deta = (nodes[3] - nodes[0]) / 2 * (1 - r) + (nodes[2] - nodes[1]) / 2 * r
        return (dxi, deta)
    r, s, t = (xi[0] + 1) / 2, (xi[1] + 1) / 2, (xi[2] + 1) / 2
    dxi = ((1 - t) * ((nodes[1] - nodes[0]) / 2 * (1 - s) + (nodes[2] - nodes[3]) / 2 * s)
           + t * ((nodes[5] - nodes[4]) / 2 * (1 - s) + (nodes[6] - nodes[7]) / 2 * s))
    deta = ((1 - r) * ((nodes[3] - nodes[0]) / 2 * (1 - t) + (nodes[7] - nodes[4]) / 2 * t)
            + r * ((nodes[2] - nodes[1]) / 2 * (1 - t) + (nodes[6] - nodes[5]) / 2 * t))
    dzeta = ((1 - s) * ((nodes[4] - nodes[0]) / 2 * (1 - r) + (nodes[5] - nodes[1]) / 2 * r)
             + s * ((nodes[7] - nodes[3]) / 2 * (1 - r) + (nodes[6] - nodes[2]) / 2 * r))
    return (dxi, deta, dzeta)


def map_dual_basis(main):
    nd = len(main)
    if nd == 1:
        return (np.array([1.0]),)
    if nd == 2:
        return (np.array([main[1][1], -main[1][0]]), np.array([-main[0][1], main[0][0]]))
    return (np.cross(main[1], main[2]), np.cross(main[2], main[0]), np.cross(main[0], main[1]))


def map_jacobian(main):
    nd = len(main)
    if nd == 1:
        return main[0][0]
    if nd == 2:
        return main[0][0] * main[1][1] - main[0][1] * main[1][0]
    return float(np.dot(main[0], np.cross(main[1], main[2])))


# ------------------------------------------------------------------- elements
def element_geometry(mesh, xi1d, cartesian=True):
    nd = mesh.nd
    xi = tensor_nodes(xi1d, nd)
    npts = len(xi)
    ne = mesh.nelements
    coords = np.zeros((ne * npts, nd))
    jac = np.zeros(ne * npts)
    metric = np.zeros((ne * npts, nd, nd))
    # shape-function weights of every reference node (same formula as phys_coords)
    nvert = 2 ** nd
    shape = np.zeros((npts, nvert))
    for i in range(npts):
        for v in range(nvert):
            unit = [np.zeros(1) for _ in range(nvert)]
            unit[v] = np.ones(1)
            shape[i, v] = phys_coords(xi[i], unit)[0]
    enodes = np.asarray(mesh.enodes, dtype=np.int64) - 1
    verts = np.asarray(mesh.nodes)[enodes]                    # (ne, nvert, nd)
    coords[:] = np.einsum("pv,evd->epd", shape, verts).reshape(ne * npts, nd)
    for e in range(ne):
        nodes = [mesh.nodes[i - 1] for i in mesh.enodes[e]] if not cartesian else None
        if cartesian:
            dx = mesh.dx
            jac[e * npts:(e + 1) * npts] = np.prod(dx) / 2 ** nd
            if nd == 1:
                m = np.array([[1.0]])
            elif nd == 2:
                m = np.array([[dx[1] / 2, 0.0], [0.0, dx[0] / 2]])
            else:
                m = np.diag([dx[1] * dx[2] / 4, dx[0] * dx[2] / 4, dx[0] * dx[1] / 4])
            metric[e * npts:(e + 1) * npts] = m
        else:
            for i in range(npts):
                main = map_basis(xi[i], nodes)
                dual = map_dual_basis(main)
                j = map_jacobian(main)
                if nd == 3:
                    if not j > 0:
                        raise ValueError(f"Found a negative Jacobian in element {e + 1}.")
                else:
                    j = abs(j)
                jac[e * npts + i] = j
                for d in range(nd):
                    metric[e * npts + i, :, d] = dual[d]
    return coords, jac, metric


# ---------------------------------------------------------------------- faces
_CART_FRAMES_2D = {
    1: ([-1, 0], [0, -1]), 2: ([1, 0], [0, 1]), 3: ([0, -1], [1, 0]), 4: ([0, 1], [-1, 0]),
}
_CART_FRAMES_3D = {
    1: ([-1, 0, 0], [0, -1, 0], [0, 0, 1]), 2: ([1, 0, 0], [0, 1, 0], [0, 0, 1]),
    3: ([0, -1, 0], [0, 0, -1], [1, 0, 0]), 4: ([0, 1, 0], [0, 0, 1], [1, 0, 0]),
    5: ([0, 0, -1], [-1, 0, 0], [0, 1, 0]), 6: ([0, 0, 1], [1, 0, 0], [0, 1, 0]),
}


def _normalize(v):
    return v / np.sqrt(np.dot(v, v))


def _face_ref_point(pos, xif, nd):
    """Reference coordinates of a face node inside the master element."""
    d = (pos - 1) // 2
    s = -1.0 if pos % 2 == 1 else 1.0
    out = np.zeros(nd)
    out[d] = s
    rest = [c for c in range(nd) if c != d]
    for c, v in zip(rest, xif):
        out[c] = v
    return out


def face_geometry(mesh, xi1d, cartesian=True, want_coords=True):
    nd = mesh.nd
    xif = tensor_nodes(xi1d, nd - 1) if nd > 1 else np.zeros((1, 1))
    nfp = len(xif) if nd > 1 else 1
    nf = mesh.nfaces
    fcoords = np.zeros((nf * nfp, nd))
    fjac = np.zeros(nf * nfp)
    frames = np.zeros((nf * nfp, 3, nd))
    for f in range(nf):
        pos = mesh.elempos[f][0]
        ielem = mesh.eleminds[f][0]
        sl = slice(f * nfp, (f + 1) * nfp)
        if cartesian:
            dx = mesh.dx
            if nd == 1:
                frames[sl, 0, 0] = -1.0 if pos == 1 else 1.0
                fjac[sl] = 1.0
            elif nd == 2:
                n, t = _CART_FRAMES_2D[pos]
                frames[sl, 0], frames[sl, 1] = n, t
                fjac[sl] = dx[1] / 2 if pos <= 2 else dx[0] / 2
            else:
                n, t, b = _CART_FRAMES_3D[pos]
                frames[sl, 0], frames[sl, 1], frames[sl, 2] = n, t, b
                fjac[sl] = (dx[1] * dx[2] / 4 if pos <= 2 else
                            dx[0] * dx[2] / 4 if pos <= 4 else dx[0] * dx[1] / 4)
            # face coordinates through the face's own vertex list (Mesh.jl:338-343);
            # on a Cartesian mesh they coincide with the master element's face nodes
            if want_coords:
                nodes = [mesh.nodes[i - 1] for i in mesh.enodes[ielem - 1]]
                for i in range(nfp):
                    fcoords[f * nfp + i] = phys_coords(_face_ref_point(pos, xif[i], nd), nodes)
        else:
            nodes = [mesh.nodes[i - 1] for i in mesh.enodes[ielem - 1]]
            d = (pos - 1) // 2
            sgn = -1.0 if pos % 2 == 1 else 1.0
            for i in range(nfp):
                xi = _face_ref_point(pos, xif[i], nd)
                fcoords[f * nfp + i] = phys_coords(xi, nodes)
                main = map_basis(xi, nodes)
                dual = map_dual_basis(main)
                if nd == 2:
                    s = np.sign(map_jacobian(main))
                    n = sgn * dual[d] * s
                    # pos1: -main[2]; pos2: +main[2]; pos3: +main[1]; pos4: -main[1]
                    tsign = {1: -1.0, 2: 1.0, 3: 1.0, 4: -1.0}[pos]
                    t = tsign * _normalize(main[1 - d]) * s
                    b = np.zeros(2)
                else:
                    n = sgn * dual[d]
                    tdir = {0: 1, 1: 2, 2: 0}[d]
                    t = sgn * _normalize(main[tdir])
                    b = _normalize(np.cross(n, t))
                j = np.sqrt(np.dot(n, n))
                fjac[f * nfp + i] = j
                frames[f * nfp + i, 0] = n / j
                frames[f * nfp + i, 1] = t
                frames[f * nfp + i, 2] = b
    return fcoords, fjac, frames


# ------------------------------------------------------------------- sub-grid
def subgrid_points_1d(w):
    """Complementary-grid points of one direction (StdSegment.jl:60-74): cumulative sums of the
    weights from both ends, averaged, ends forced to -1 / +1."""
    n = len(w)
    c1 = np.zeros(n + 1)
    c1[0] = -1.0
    for i in range(n):
        c1[i + 1] = c1[i] + w[i]
    c2 = np.zeros(n + 1)
    c2[n] = 1.0
    for i in range(n - 1, -1, -1):
        c2[i] = c2[i + 1] - w[i]
    c = (c1 + c2) / 2
    c[0], c[n] = -1.0, 1.0
    return c


def _line_base_xi(nd, n, d, k, xi1d):
    """Reference coordinates (other than direction d) of tensor-product line k of direction d
    (tpdofs order: StdQuad.jl:116-124, StdHex.jl:135-145)."""
    out = np.zeros(nd)
    if nd == 2:
        out[1 - d] = xi1d[k]
    elif nd == 3:
        a, b = k % n, k // n
        if d == 0:
            out[1], out[2] = xi1d[a], xi1d[b]
        elif d == 1:
            out[0], out[2] = xi1d[a], xi1d[b]
        else:
            out[0], out[1] = xi1d[a], xi1d[b]
    return out


def subgrid_geometry(mesh, xi1d, w1d, cartesian=True):
    """geometry.subgrids[e].frames[dir][is], .jac[dir][is] re-indexed by (element, direction,
    line, position along the line): frames (ne, nd, nlines, np+1, 3, nd) rows n, t, b and
    jac (ne, nd, nlines, np+1).  Sub-grid point `ii` of a line sits at xi_c[ii] along the line's
    direction and at the line's own nodes otherwise (tpdofs_subgrid)."""
    nd, n = mesh.nd, len(xi1d)
    nlines = n ** (nd - 1)
    ne = mesh.nelements
    xic = subgrid_points_1d(w1d)
    frames = np.zeros((ne, nd, nlines, n + 1, 3, nd))
    jac = np.zeros((ne, nd, nlines, n + 1))
    for e in range(ne):
        if cartesian:
            dx = mesh.dx
            for d in range(nd):
                if nd == 1:
                    frames[e, d, :, :, 0, 0] = 1.0
                    jac[e, d] = 1.0
                elif nd == 2:
                    frames[e, d, :, :, 0, d] = 1.0
                    frames[e, d, :, :, 1, 1 - d] = 1.0 if d == 0 else -1.0
                    jac[e, d] = dx[1 - d] / 2
                else:
                    frames[e, d, :, :, 0, d] = 1.0
                    frames[e, d, :, :, 1, (d + 1) % 3] = 1.0
                    frames[e, d, :, :, 2, (d + 2) % 3] = 1.0
                    jac[e, d] = np.prod(dx) / dx[d] / 4
            continue
        nodes = [mesh.nodes[i - 1] for i in mesh.enodes[e]]
        for d in range(nd):
            for k in range(nlines):
                base = _line_base_xi(nd, n, d, k, xi1d)
                for ii in range(n + 1):
                    xi = base.copy()
                    xi[d] = xic[ii]
                    main = map_basis(xi, nodes)
                    dual = map_dual_basis(main)
                    if nd == 1:
                        s = np.sign(map_jacobian(main))
                        frames[e, d, k, ii, 0] = s * dual[0]
                        jac[e, d, k, ii] = 1.0
                        continue
                    if nd == 2:
                        s = np.sign(map_jacobian(main))
                        nvec = s * dual[d]
                        # vertical faces: t = s normalize(main[2]); horizontal: t = -s normalize(main[1])
                        t = s * _normalize(main[1]) if d == 0 else -s * _normalize(main[0])
                        b = np.zeros(2)
                    else:
                        nvec = dual[d]
                        t = _normalize(main[(d + 1) % 3])
                        b = None
                    j = np.sqrt(np.dot(nvec, nvec))
                    nvec = nvec / j
                    if nd == 3:
                        b = _normalize(np.cross(nvec, t))
                    frames[e, d, k, ii, 0], frames[e, d, k, ii, 1], frames[e, d, k, ii, 2] = nvec, t, b
                    jac[e, d, k, ii] = j
    return frames, jac
