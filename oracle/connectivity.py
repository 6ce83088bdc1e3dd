"""ORACLE (test infrastructure, never shipped, never on the product path).

Literal restatement (plain Python loops, small meshes only) of Flou.jl's Cartesian mesh
connectivity and of the periodic-boundary face merge.  All ids are 1-based exactly as in
the reference; `0` marks "no element" on a boundary face.

* element connectivity   src/FlouCommon/CartesianMesh.jl:152-236
* face connectivity      src/FlouCommon/CartesianMesh.jl:238-500
* periodic merge         src/FlouCommon/Mesh.jl:236-322, CartesianMesh.jl:127-150
"""
import numpy as np


class Mesh:
    """Plain container mirroring the fields of CartesianMesh (CartesianMesh.jl:15-27)."""

    def __init__(self, nd, nxyz, dx, nodes, faceinds, facepos, eleminds, elempos,
                 orientation, intfaces, bdfaces, bdnames, bdmap):
        self.nd = nd
        self.nxyz = tuple(nxyz)
        self.dx = tuple(dx)
        self.nodes = nodes
        self.faceinds = faceinds      # list of lists, per element, 1-based face ids
        self.facepos = facepos        # list of lists, per element, 1 (master) / 2 (slave)
        self.eleminds = eleminds      # list of [master, slave] (0 = none)
        self.elempos = elempos        # list of [pos in master, pos in slave]
        self.orientation = orientation
        self.intfaces = intfaces
        self.bdfaces = bdfaces        # list of lists (per boundary)
        self.bdnames = bdnames
        self.bdmap = bdmap            # dict: original boundary id -> current index (1-based)
        self.periodic = {}

    @property
    def nelements(self):
        return len(self.faceinds)

    @property
    def nfaces(self):
        return len(self.eleminds)


def _elements(nd, n):
    faceinds, facepos, enodes = [], [], []
    if nd == 1:
        (nx,) = n
        for i in range(1, nx + 1):
            enodes.append([i, i + 1])
            faceinds.append([i, i + 1])
            facepos.append([1 if i == 1 else 2, 1])
    elif nd == 2:
        nx, ny = n
        npx = nx + 1
        for j in range(1, ny + 1):
            for i in range(1, nx + 1):
                enodes.append([(j-1)*npx + i, (j-1)*npx + i + 1,
                               (j-1)*npx + i + 1 + npx, (j-1)*npx + i + npx])
                faceinds.append([(j-1)*npx + i, (j-1)*npx + i + 1,
                                 npx*ny + (j-1)*nx + i, npx*ny + (j-1)*nx + i + nx])
                facepos.append([1 if i == 1 else 2, 1, 1 if j == 1 else 2, 1])
    else:
        nx, ny, nz = n
        npx, npy = nx + 1, ny + 1
        for k in range(1, nz + 1):
            for j in range(1, ny + 1):
                for i in range(1, nx + 1):
                    b = (k-1)*npx*npy + (j-1)*npx + i
                    enodes.append([b, b + 1, b + 1 + npx, b + npx,
                                   npx*npy + b, npx*npy + b + 1,
                                   npx*npy + b + 1 + npx, npx*npy + b + npx])
                    faceinds.append([
                        (k-1)*npx*ny + (j-1)*npx + i,
                        (k-1)*npx*ny + (j-1)*npx + i + 1,
                        npx*ny*nz + (k-1)*nx*npy + (j-1)*nx + i,
                        npx*ny*nz + (k-1)*nx*npy + (j-1)*nx + i + nx,
                        npx*ny*nz + nx*npy*nz + (k-1)*nx*ny + (j-1)*nx + i,
                        npx*ny*nz + nx*npy*nz + (k-1)*nx*ny + (j-1)*nx + i + nx*ny,
                    ])
                    facepos.append([1 if i == 1 else 2, 1, 1 if j == 1 else 2, 1,
                                    1 if k == 1 else 2, 1])
    return enodes, faceinds, facepos


def _faces(nd, n):
    if nd == 1:
        (nx,) = n
        nfaces = nx + 1
        eleminds = [None] * nfaces
        elempos = [None] * nfaces
        intfaces, bdfaces = [], [[], []]
        for f in range(1, nfaces + 1):
            if f == 1:
                bdfaces[0].append(f); eleminds[f-1] = [1, 0]; elempos[f-1] = [1, 0]
            elif f == nfaces:
                bdfaces[1].append(f); eleminds[f-1] = [nx, 0]; elempos[f-1] = [2, 0]
            else:
                intfaces.append(f); eleminds[f-1] = [f - 1, f]; elempos[f-1] = [2, 1]
        return intfaces, bdfaces, eleminds, elempos
    if nd == 2:
        nx, ny = n
        npx, npy = nx + 1, ny + 1
        nfaces = npx*ny + npy*nx
        eleminds = [None] * nfaces
        elempos = [None] * nfaces
        intfaces, bdfaces = [], [[] for _ in range(4)]
        for j in range(1, ny + 1):
            for i in range(1, npx + 1):
                f = (j-1)*npx + i
                if i == 1:
                    bdfaces[0].append(f); eleminds[f-1] = [(j-1)*nx + 1, 0]; elempos[f-1] = [1, 0]
                elif i == npx:
                    bdfaces[1].append(f); eleminds[f-1] = [(j-1)*nx + nx, 0]; elempos[f-1] = [2, 0]
                else:
                    intfaces.append(f)
                    eleminds[f-1] = [(j-1)*nx + i - 1, (j-1)*nx + i]; elempos[f-1] = [2, 1]
        for j in range(1, npy + 1):
            for i in range(1, nx + 1):
                f = npx*ny + (j-1)*nx + i
                if j == 1:
                    bdfaces[2].append(f); eleminds[f-1] = [i, 0]; elempos[f-1] = [3, 0]
                elif j == npy:
                    bdfaces[3].append(f); eleminds[f-1] = [(ny-1)*nx + i, 0]; elempos[f-1] = [4, 0]
                else:
                    intfaces.append(f)
                    eleminds[f-1] = [(j-2)*nx + i, (j-2)*nx + i + nx]; elempos[f-1] = [4, 3]
        return intfaces, bdfaces, eleminds, elempos
    nx, ny, nz = n
    npx, npy, npz = nx + 1, ny + 1, nz + 1
    nfaces = npx*ny*nz + npy*nx*nz + npz*nx*ny
    eleminds = [None] * nfaces
    elempos = [None] * nfaces
    intfaces, bdfaces = [], [[] for _ in range(6)]
    for k in range(1, nz + 1):
        for j in range(1, ny + 1):
            for i in range(1, npx + 1):
                f = (k-1)*npx*ny + (j-1)*npx + i
                if i == 1:
                    bdfaces[0].append(f)
                    eleminds[f-1] = [(k-1)*nx*ny + (j-1)*nx + 1, 0]; elempos[f-1] = [1, 0]
                elif i == npx:
                    bdfaces[1].append(f)
                    eleminds[f-1] = [(k-1)*nx*ny + (j-1)*nx + nx, 0]; elempos[f-1] = [2, 0]
                else:
                    intfaces.append(f)
                    eleminds[f-1] = [(k-1)*nx*ny + (j-1)*nx + i - 1, (k-1)*nx*ny + (j-1)*nx + i]
                    elempos[f-1] = [2, 1]
    for k in range(1, nz + 1):
        for j in range(1, npy + 1):
            for i in range(1, nx + 1):
                f = npx*ny*nz + (k-1)*nx*npy + (j-1)*nx + i
                if j == 1:
                    bdfaces[2].append(f)
                    eleminds[f-1] = [(k-1)*nx*ny + i, 0]; elempos[f-1] = [3, 0]
                elif j == npy:
                    bdfaces[3].append(f)
                    eleminds[f-1] = [(k-1)*nx*ny + (ny-1)*nx + i, 0]; elempos[f-1] = [4, 0]
                else:
                    intfaces.append(f)
                    eleminds[f-1] = [(k-1)*nx*ny + (j-2)*nx + i, (k-1)*nx*ny + (j-2)*nx + i + nx]
                    elempos[f-1] = [4, 3]
    for k in range(1, npz + 1):
        for j in range(1, ny + 1):
            for i in range(1, nx + 1):
                f = npx*ny*nz + npy*nx*nz + (k-1)*nx*ny + (j-1)*nx + i
                if k == 1:
                    bdfaces[4].append(f)
                    eleminds[f-1] = [(j-1)*nx + i, 0]; elempos[f-1] = [5, 0]
                elif k == npz:
                    bdfaces[5].append(f)
                    eleminds[f-1] = [(nz-1)*nx*ny + (j-1)*nx + i, 0]; elempos[f-1] = [6, 0]
                else:
                    intfaces.append(f)
                    eleminds[f-1] = [(k-2)*nx*ny + (j-1)*nx + i, (k-2)*nx*ny + (j-1)*nx + i + nx*ny]
                    elempos[f-1] = [6, 5]
    return intfaces, bdfaces, eleminds, elempos


def cartesian_mesh(start, finish, nxyz):
    """CartesianMesh{ND,Float64}(start, finish, nxyz)  (CartesianMesh.jl:38-97)."""
    start = np.atleast_1d(np.asarray(start, dtype=float))
    finish = np.atleast_1d(np.asarray(finish, dtype=float))
    nxyz = [int(v) for v in np.atleast_1d(nxyz)]
    nd = len(nxyz)
    if not (1 <= nd <= 3) or len(start) != nd or len(finish) != nd:
        raise ValueError("The mesh can only have 1, 2 or 3 dimensions.")
    if not np.all(start < finish):
        raise ValueError("All components of `start` must be lower than those of `finish`.")
    # Julia `range(a, b, n)` : a + (i-1)*(b-a)/(n-1) evaluated in twice-precision;
    # np.linspace agrees to the last bit or one ulp.
    xyz = [np.linspace(start[d], finish[d], nxyz[d] + 1) for d in range(nd)]
    grids = np.meshgrid(*xyz, indexing="ij")
    nodes = np.stack([g.reshape(-1, order="F") for g in grids], axis=1)
    enodes, faceinds, facepos = _elements(nd, nxyz)
    intfaces, bdfaces, eleminds, elempos = _faces(nd, nxyz)
    dx = tuple((finish - start) / np.array(nxyz))
    mesh = Mesh(nd, nxyz, dx, nodes, faceinds, facepos, eleminds, elempos,
                [0] * len(eleminds), intfaces, bdfaces,
                [str(i) for i in range(1, 2 * nd + 1)],
                {i: i for i in range(1, 2 * nd + 1)})
    mesh.enodes = enodes
    return mesh


def apply_periodic_bcs(mesh, *pairs):
    """apply_periodicBCs!(mesh, "1"=>"2", ...)  (CartesianMesh.jl:127-150 -> Mesh.jl:236-322)."""
    bcs = {}
    for a, b in pairs:
        bd1, bd2 = int(a), int(b)
        nb = len(mesh.bdfaces)
        if not (1 <= bd1 <= nb and 1 <= bd2 <= nb):
            raise ValueError(f"Boundary IDs must be between 1 and {nb}.")
        if not (bd2 % 2 == 0 and bd1 + 1 == bd2):
            raise ValueError(f"Boundaries {a} and {b} cannot be made periodic.")
        bcs[bd1] = bd2
    faces2del = []
    for bd1, bd2 in bcs.items():
        if bd1 not in mesh.bdmap or bd2 not in mesh.bdmap:
            raise ValueError(f"Boundaries {bd1} and {bd2} cannot be periodic.")
        first, second = mesh.bdmap[bd1], mesh.bdmap[bd2]
        for if1, if2 in zip(mesh.bdfaces[first - 1], mesh.bdfaces[second - 1]):
            elmind = mesh.eleminds[if2 - 1][0]
            elmpos = mesh.elempos[if2 - 1][0]
            mesh.eleminds[if1 - 1][1] = elmind
            mesh.elempos[if1 - 1][1] = elmpos
            mesh.faceinds[elmind - 1][elmpos - 1] = if1
            mesh.facepos[elmind - 1][elmpos - 1] = 2
        mesh.intfaces.extend(mesh.bdfaces[first - 1])
        faces2del.extend(mesh.bdfaces[second - 1])
        for idx in sorted((first, second), reverse=True):
            del mesh.bdfaces[idx - 1]
        del mesh.bdmap[bd1]
        del mesh.bdmap[bd2]
        for i in list(mesh.bdmap):
            if mesh.bdmap[i] > second:
                mesh.bdmap[i] -= 1
            if mesh.bdmap[i] > first:
                mesh.bdmap[i] -= 1
        mesh.periodic[bd1] = bd2
    mesh.intfaces.sort()
    faces2del.sort()
    f2d = np.array(faces2del, dtype=np.int64)

    def shift(f):
        # number of deleted faces with id < f  (Mesh.jl:283-304)
        return int(np.searchsorted(f2d, f, side="right"))

    mesh.intfaces = [f - shift(f) for f in mesh.intfaces]
    mesh.bdfaces = [[f - shift(f) for f in bd] for bd in mesh.bdfaces]
    dele = set(faces2del)
    keep = [i for i in range(len(mesh.eleminds)) if (i + 1) not in dele]
    mesh.eleminds = [mesh.eleminds[i] for i in keep]
    mesh.elempos = [mesh.elempos[i] for i in keep]
    mesh.orientation = [mesh.orientation[i] for i in keep]
    for iface in range(1, len(mesh.eleminds) + 1):
        e, p = mesh.eleminds[iface - 1][0], mesh.elempos[iface - 1][0]
        mesh.faceinds[e - 1][p - 1] = iface
        e = mesh.eleminds[iface - 1][1]
        if e != 0:
            p = mesh.elempos[iface - 1][1]
            mesh.faceinds[e - 1][p - 1] = iface
    return mesh
