"""Host-side geometry tables (vectorised numpy).

  linear segment / bilinear quad / trilinear hex mappings    src/FlouCommon/Mesh.jl:363-487
  per-node jac + metric (unstructured)                       src/FlouSpatial/PhysicalRegions.jl:438-472
  face frames + face jac from the MASTER element             PhysicalRegions.jl:797-871 (2-D), 873-971 (3-D)
Cartesian meshes do not need these tables on the device (constants derived from dx inside
the library, PhysicalRegions.jl:370-408, 541-696); `element_coords` is still used for
initial conditions and GenericBC tabulation.
"""
import numpy as np


def _shape(xi, nd):
    """Vertex weights of the (bi/tri)linear map at reference points xi (npts, nd) -> (npts, 2^nd)."""
    r = (xi + 1) / 2
    if nd == 1:
        return np.stack([1 - r[:, 0], r[:, 0]], axis=1)
    if nd == 2:
        a, b = r[:, 0], r[:, 1]
        return np.stack([(1 - a) * (1 - b), a * (1 - b), a * b, (1 - a) * b], axis=1)
    a, b, c = r[:, 0], r[:, 1], r[:, 2]
    return np.stack([(1 - a) * (1 - b) * (1 - c), a * (1 - b) * (1 - c), a * b * (1 - c),
                     (1 - a) * b * (1 - c), (1 - a) * (1 - b) * c, a * (1 - b) * c,
                     a * b * c, (1 - a) * b * c], axis=1)


def element_coords(verts, xi):
    """Physical coordinates of every node: verts (ne, 2^nd, nd), xi (npts, nd) -> (ne*npts, nd)."""
    nd = xi.shape[1]
    N = _shape(xi, nd)
    x = np.einsum("pv,evd->epd", N, verts)
    return x.reshape(-1, nd)


def _main_basis(verts, xi):
    """Covariant basis vectors dx/dxi_d at reference points: (ne, npts, nd[d], nd[c])."""
    nd = xi.shape[1]
    r = (xi + 1) / 2
    V = verts
    if nd == 1:
        d0 = (V[:, 1] - V[:, 0]) / 2
        return np.broadcast_to(d0[:, None, None, :], (V.shape[0], xi.shape[0], 1, 1)).copy()
    if nd == 2:
        a, b = r[None, :, 0, None], r[None, :, 1, None]
        e = lambda i: V[:, None, i, :]
        dxi = (e(1) - e(0)) / 2 * (1 - b) + (e(2) - e(3)) / 2 * b
        deta = (e(3) - e(0)) / 2 * (1 - a) + (e(2) - e(1)) / 2 * a
        return np.stack([dxi, deta], axis=2)
    a, b, c = r[None, :, 0, None], r[None, :, 1, None], r[None, :, 2, None]
    e = lambda i: V[:, None, i, :]
    dxi = ((1 - c) * ((e(1) - e(0)) / 2 * (1 - b) + (e(2) - e(3)) / 2 * b)
           + c * ((e(5) - e(4)) / 2 * (1 - b) + (e(6) - e(7)) / 2 * b))
    deta = ((1 - a) * ((e(3) - e(0)) / 2 * (1 - c) + (e(7) - e(4)) / 2 * c)
            + a * ((e(2) - e(1)) / 2 * (1 - c) + (e(6) - e(5)) / 2 * c))
    dzeta = ((1 - b) * ((e(4) - e(0)) / 2 * (1 - a) + (e(5) - e(1)) / 2 * a)
             + b * ((e(7) - e(3)) / 2 * (1 - a) + (e(6) - e(2)) / 2 * a))
    return np.stack([dxi, deta, dzeta], axis=2)


def _dual_and_jac(main):
    nd = main.shape[2]
    if nd == 1:
        return np.ones_like(main), main[..., 0, 0]
    if nd == 2:
        m1, m2 = main[..., 0, :], main[..., 1, :]
        d1 = np.stack([m2[..., 1], -m2[..., 0]], axis=-1)
        d2 = np.stack([-m1[..., 1], m1[..., 0]], axis=-1)
        jac = m1[..., 0] * m2[..., 1] - m1[..., 1] * m2[..., 0]
        return np.stack([d1, d2], axis=2), jac
    m1, m2, m3 = main[..., 0, :], main[..., 1, :], main[..., 2, :]
    d1, d2, d3 = np.cross(m2, m3), np.cross(m3, m1), np.cross(m1, m2)
    jac = np.einsum("...c,...c->...", m1, d1)
    return np.stack([d1, d2, d3], axis=2), jac


def general_element_geometry(verts, xi):
    """jac (ne*npts,), metric (ne*npts, nd*nd) with [c + nd*d] = Ja^d_c."""
    ne, nd = verts.shape[0], xi.shape[1]
    main = _main_basis(verts, xi)
    dual, jac = _dual_and_jac(main)
    if nd == 3:
        if not np.all(jac > 0):
            bad = int(np.argwhere(~(jac > 0))[0][0]) + 1
            raise ArithmeticError(f"Found a negative Jacobian in element {bad}.")
    else:
        jac = np.abs(jac)
    metric = dual.reshape(ne * xi.shape[0], nd * nd)     # [d][c] flattened -> index d*nd + c
    return jac.reshape(-1), np.ascontiguousarray(metric)


def _face_ref_points(pos, xif, nd):
    d = (pos - 1) // 2
    s = -1.0 if pos % 2 == 1 else 1.0
    out = np.zeros((xif.shape[0], nd))
    out[:, d] = s
    rest = [c for c in range(nd) if c != d]
    for j, c in enumerate(rest):
        out[:, c] = xif[:, j]
    return out


def general_face_geometry(verts, eleminds, elempos, xif, nd):
    """Face coordinates, jac and frames from the master element.

    Returns fcoords (nf*nfp, nd), fjac (nf*nfp,), frames (nf*nfp, 3*nd) rows n, t, b."""
    nf = eleminds.shape[0]
    nfp = xif.shape[0] if nd > 1 else 1
    fcoords = np.zeros((nf, nfp, nd))
    fjac = np.zeros((nf, nfp))
    frames = np.zeros((nf, nfp, 3, nd))
    master = eleminds[:, 0] - 1
    for pos in range(1, 2 * nd + 1):
        sel = np.nonzero(elempos[:, 0] == pos)[0]
        if sel.size == 0:
            continue
        xi = _face_ref_points(pos, xif if nd > 1 else np.zeros((1, 0)), nd)
        V = verts[master[sel]]
        fcoords[sel] = np.einsum("pv,evd->epd", _shape(xi, nd), V)
        main = _main_basis(V, xi)
        dual, jac = _dual_and_jac(main)
        d = (pos - 1) // 2
        sgn = -1.0 if pos % 2 == 1 else 1.0
        if nd == 1:
            # PhysicalRegions.jl:760-795: n = -+1, jac = 1
            frames[sel, :, 0, 0] = sgn
            fjac[sel] = 1.0
            continue
        if nd == 2:
            s = np.sign(jac)[..., None]
            n = sgn * dual[..., d, :] * s
            tsign = {1: -1.0, 2: 1.0, 3: 1.0, 4: -1.0}[pos]
            tv = main[..., 1 - d, :]
            t = tsign * tv / np.linalg.norm(tv, axis=-1, keepdims=True) * s
            b = np.zeros_like(n)
        else:
            n = sgn * dual[..., d, :]
            tv = main[..., {0: 1, 1: 2, 2: 0}[d], :]
            t = sgn * tv / np.linalg.norm(tv, axis=-1, keepdims=True)
            b = np.cross(n, t)
            b = b / np.linalg.norm(b, axis=-1, keepdims=True)
        j = np.linalg.norm(n, axis=-1)
        fjac[sel] = j
        frames[sel, :, 0] = n / j[..., None]
        frames[sel, :, 1] = t
        frames[sel, :, 2] = b
    return (fcoords.reshape(-1, nd), fjac.reshape(-1),
            np.ascontiguousarray(frames.reshape(nf * nfp, 3 * nd)))


# ------------------------------------------------------------------ sub-grid (HybridDivOperator, Gauss-node split form)
def subgrid_points_1d(w):
    """Complementary-grid points of one direction (StdSegment.jl:60-74): cumulative sums of the
    weights from either end, averaged; end points exactly -1 and +1."""
    w = np.asarray(w, dtype=np.float64)
    n = w.size
    c1 = np.concatenate(([-1.0], -1.0 + np.cumsum(w)))
    c2 = np.concatenate((1.0 - np.cumsum(w[::-1])[::-1], [1.0]))
    c = (c1 + c2) / 2
    c[0], c[n] = -1.0, 1.0
    return c


def general_subgrid_geometry(verts, xi1d, w1d):
    """geometry.subgrids (PhysicalRegions.jl:179-292) of unstructured elements, indexed by
    (element, direction, tensor-product line, position along the line):
      frames (ne, nd, nlines, np+1, 3*nd) rows n, t, b;   jac (ne, nd, nlines, np+1).
    Point `ii` of line k of direction d sits at xi_c[ii] along d and at the line's own node
    coordinates otherwise (tpdofs_subgrid, StdQuad.jl:126-137)."""
    ne, nd = verts.shape[0], verts.shape[2]
    n = len(xi1d)
    nlines = n ** (nd - 1)
    xic = subgrid_points_1d(w1d)
    frames = np.zeros((ne, nd, nlines, n + 1, 3, nd))
    fjac = np.zeros((ne, nd, nlines, n + 1))
    for d in range(nd):
        # reference points (nlines*(n+1), nd), line-major
        pts = np.zeros((nlines, n + 1, nd))
        pts[:, :, d] = xic[None, :]
        if nd == 2:
            pts[:, :, 1 - d] = np.asarray(xi1d)[:, None]
        elif nd == 3:
            k = np.arange(nlines)
            a, b = np.asarray(xi1d)[k % n], np.asarray(xi1d)[k // n]
            o1, o2 = [c for c in range(3) if c != d]
            pts[:, :, o1], pts[:, :, o2] = a[:, None], b[:, None]
        main = _main_basis(verts, pts.reshape(-1, nd))            # (ne, P, nd, nd)
        dual, jac = _dual_and_jac(main)
        if nd == 1:
            frames[:, 0, :, :, 0, 0] = (np.sign(jac) * dual[..., 0, 0]).reshape(ne, nlines, n + 1)
            fjac[:, 0] = 1.0
            continue
        if nd == 2:
            s = np.sign(jac)[..., None]
            nv = s * dual[..., d, :]
            tv = main[..., 1, :] if d == 0 else -main[..., 0, :]
            t = s * tv / np.linalg.norm(tv, axis=-1, keepdims=True)
            b = np.zeros_like(nv)
        else:
            nv = dual[..., d, :]
            tv = main[..., (d + 1) % 3, :]
            t = tv / np.linalg.norm(tv, axis=-1, keepdims=True)
        j = np.linalg.norm(nv, axis=-1)
        nv = nv / j[..., None]
        if nd == 3:
            b = np.cross(nv, t)
            b = b / np.linalg.norm(b, axis=-1, keepdims=True)
        shp = (ne, nlines, n + 1)
        frames[:, d, :, :, 0] = nv.reshape(shp + (nd,))
        frames[:, d, :, :, 1] = t.reshape(shp + (nd,))
        frames[:, d, :, :, 2] = b.reshape(shp + (nd,))
        fjac[:, d] = j.reshape(shp)
    return np.ascontiguousarray(frames.reshape(ne, nd, nlines, n + 1, 3 * nd)), np.ascontiguousarray(fjac)
