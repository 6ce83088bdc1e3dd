"""Host-side mirror of Flou.jl's meshes: ids, master/slave sides, face positions and
orientations are produced with the reference's numbering (1-based), vectorised in numpy so
that 128^3-element meshes build in seconds.

  CartesianMesh{ND,RT}(start, finish, nxyz)   src/FlouCommon/CartesianMesh.jl:38-97
  element / face connectivity                  CartesianMesh.jl:152-236, 238-500
  apply_periodicBCs!                           CartesianMesh.jl:127-150 -> Mesh.jl:236-322
"""
import numpy as np


class CartesianMesh:
    """CartesianMesh(nd, start, finish, nxyz)  ==  CartesianMesh{nd,Float64}(start, finish, nxyz)."""

    cartesian = True

    def __init__(self, nd, start, finish, nxyz):
        start = np.atleast_1d(np.asarray(start, dtype=np.float64))
        finish = np.atleast_1d(np.asarray(finish, dtype=np.float64))
        n = [int(v) for v in np.atleast_1d(nxyz)]
        if not 1 <= nd <= 3:
            raise ValueError("The mesh can only have 1, 2 or 3 dimensions.")
        if len(start) != nd:
            raise ValueError(f"The `start` point must have {nd} coordinates.")
        if len(finish) != nd:
            raise ValueError(f"The `finish` point must have {nd} coordinates.")
        if len(n) != nd:
            raise ValueError("The number of elements in all directions must be given.")
        if not np.all(start < finish):
            raise ValueError("All components of `start` must be lower than those of `finish`.")
        self.nd = nd
        self.nelements_dir = tuple(n)
        self.dx = tuple((finish - start) / np.array(n, dtype=np.float64))
        self.start, self.finish = start, finish
        self.xyz = [np.linspace(start[d], finish[d], n[d] + 1) for d in range(nd)]
        self._build(n)
        self.bdnames = [str(i) for i in range(1, 2 * nd + 1)]
        self.bdmap = {i: i for i in range(1, 2 * nd + 1)}
        self.periodic = {}

    # ------------------------------------------------------------------ connectivity
    def _build(self, n):
        nd = self.nd
        nx = n[0]
        ny = n[1] if nd > 1 else 1
        nz = n[2] if nd > 2 else 1
        npx, npy, npz = nx + 1, ny + 1, nz + 1
        ne = nx * ny * nz
        I, J, K = np.meshgrid(np.arange(1, nx + 1), np.arange(1, ny + 1), np.arange(1, nz + 1),
                              indexing="ij")
        i = I.reshape(-1, order="F").astype(np.int64)
        j = J.reshape(-1, order="F").astype(np.int64)
        k = K.reshape(-1, order="F").astype(np.int64)
        faceinds = np.empty((ne, 2 * nd), dtype=np.int64)
        facepos = np.ones((ne, 2 * nd), dtype=np.int64)
        if nd == 1:
            faceinds[:, 0], faceinds[:, 1] = i, i + 1
            nfx, nfy, nfz = npx, 0, 0
        elif nd == 2:
            faceinds[:, 0] = (j - 1) * npx + i
            faceinds[:, 1] = faceinds[:, 0] + 1
            faceinds[:, 2] = npx * ny + (j - 1) * nx + i
            faceinds[:, 3] = faceinds[:, 2] + nx
            nfx, nfy, nfz = npx * ny, npy * nx, 0
        else:
            faceinds[:, 0] = (k - 1) * npx * ny + (j - 1) * npx + i
            faceinds[:, 1] = faceinds[:, 0] + 1
            faceinds[:, 2] = npx * ny * nz + (k - 1) * nx * npy + (j - 1) * nx + i
            faceinds[:, 3] = faceinds[:, 2] + nx
            faceinds[:, 4] = npx * ny * nz + nx * npy * nz + (k - 1) * nx * ny + (j - 1) * nx + i
            faceinds[:, 5] = faceinds[:, 4] + nx * ny
            nfx, nfy, nfz = npx * ny * nz, npy * nx * nz, npz * nx * ny
        facepos[:, 0] = np.where(i == 1, 1, 2)
        if nd > 1:
            facepos[:, 2] = np.where(j == 1, 1, 2)
        if nd > 2:
            facepos[:, 4] = np.where(k == 1, 1, 2)
        nf = nfx + nfy + nfz
        eleminds = np.zeros((nf, 2), dtype=np.int64)
        elempos = np.zeros((nf, 2), dtype=np.int64)
        elem = np.arange(1, ne + 1, dtype=np.int64)
        # each element is the master (side 1) of its high face and, on the low boundary, of
        # its low face; it is the slave (side 2) of an interior low face
        for d in range(nd):
            lo, hi = faceinds[:, 2 * d] - 1, faceinds[:, 2 * d + 1] - 1
            eleminds[hi, 0] = elem
            elempos[hi, 0] = 2 * d + 2
            first = facepos[:, 2 * d] == 1
            eleminds[lo[first], 0] = elem[first]
            elempos[lo[first], 0] = 2 * d + 1
            eleminds[lo[~first], 1] = elem[~first]
            elempos[lo[~first], 1] = 2 * d + 1
        self.faceinds, self.facepos = faceinds, facepos
        self.eleminds, self.elempos = eleminds, elempos
        self.orientation = np.zeros(nf, dtype=np.uint8)
        # boundary lists in -x, +x, -y, +y, -z, +z order, ascending face id (loop order)
        idx = (i, j, k)
        nn = (nx, ny, nz)
        bdfaces = []
        for d in range(nd):
            bdfaces.append(np.sort(faceinds[idx[d] == 1, 2 * d]))
            bdfaces.append(np.sort(faceinds[idx[d] == nn[d], 2 * d + 1]))
        self.bdfaces = bdfaces
        isbd = np.zeros(nf + 1, dtype=bool)
        for b in bdfaces:
            isbd[b] = True
        self.intfaces = np.nonzero(~isbd[1:])[0].astype(np.int64) + 1

    @property
    def nelements(self):
        return self.faceinds.shape[0]

    @property
    def nfaces(self):
        return self.eleminds.shape[0]

    def nboundaries(self):
        return len(self.bdfaces)

    # ------------------------------------------------------------------ periodic merge
    def apply_periodicBCs(self, *pairs):
        """apply_periodicBCs!(mesh, "1" => "2", ...) with pairs given as ("1", "2")."""
        bcs = {}
        nb = self.nboundaries()
        for a, b in pairs:
            try:
                bd1, bd2 = int(a), int(b)
            except (TypeError, ValueError):
                raise ValueError("Boundary IDs must be integers.")
            if not (1 <= bd1 <= nb and 1 <= bd2 <= nb):
                raise ValueError(f"Boundary IDs must be between 1 and {nb}.")
            if not (bd2 % 2 == 0 and bd1 + 1 == bd2):
                raise ValueError(f"Boundaries {a} and {b} cannot be made periodic.")
            bcs[bd1] = bd2
        faces2del = []
        intfaces = [self.intfaces]
        for bd1, bd2 in bcs.items():
            if bd1 not in self.bdmap or bd2 not in self.bdmap:
                raise ValueError(f"Boundaries {bd1} and {bd2} cannot be periodic.")
            first, second = self.bdmap[bd1], self.bdmap[bd2]
            f1, f2 = self.bdfaces[first - 1], self.bdfaces[second - 1]
            elmind = self.eleminds[f2 - 1, 0]
            elmpos = self.elempos[f2 - 1, 0]
            self.eleminds[f1 - 1, 1] = elmind
            self.elempos[f1 - 1, 1] = elmpos
            self.faceinds[elmind - 1, elmpos - 1] = f1
            self.facepos[elmind - 1, elmpos - 1] = 2
            intfaces.append(f1)
            faces2del.append(f2)
            for idx in sorted((first, second), reverse=True):
                del self.bdfaces[idx - 1]
            del self.bdmap[bd1]
            del self.bdmap[bd2]
            for key in list(self.bdmap):
                if self.bdmap[key] > second:
                    self.bdmap[key] -= 1
                if self.bdmap[key] > first:
                    self.bdmap[key] -= 1
            self.periodic[bd1] = bd2
        if not faces2del:
            return self
        f2d = np.sort(np.concatenate(faces2del))
        nf_old = self.nfaces
        keep = np.ones(nf_old, dtype=bool)
        keep[f2d - 1] = False
        # new id of a surviving face = old id - #(deleted ids below it)   (Mesh.jl:283-304)
        newid = np.zeros(nf_old + 1, dtype=np.int64)
        newid[1:][keep] = np.arange(1, int(keep.sum()) + 1)
        self.intfaces = np.sort(newid[np.concatenate(intfaces)])
        self.bdfaces = [newid[b] for b in self.bdfaces]
        self.eleminds = self.eleminds[keep]
        self.elempos = self.elempos[keep]
        self.orientation = self.orientation[keep]
        self.faceinds = newid[self.faceinds]
        return self

    # ------------------------------------------------------------------ geometry helpers
    @property
    def nodes(self):
        """mesh.nodes: (nvertices, nd), x fastest (CartesianMesh.jl:56-72)."""
        if getattr(self, "_nodes", None) is None:
            grids = np.meshgrid(*self.xyz, indexing="ij")
            self._nodes = np.stack([g.reshape(-1, order="F") for g in grids], axis=1)
        return self._nodes

    @nodes.setter
    def nodes(self, value):
        self._nodes = np.ascontiguousarray(value, dtype=np.float64)

    @property
    def nodeinds(self):
        """mesh.elements.nodeinds: (ne, 2^nd) 1-based vertex ids (CartesianMesh.jl:157,175-215)."""
        nd = self.nd
        n = self.nelements_dir
        npx = n[0] + 1
        npy = n[1] + 1 if nd > 1 else 1
        grids = np.meshgrid(*[np.arange(m, dtype=np.int64) for m in n], indexing="ij")
        ijk = [g.reshape(-1, order="F") for g in grids]
        base = ijk[0] + 1
        if nd > 1:
            base = base + ijk[1] * npx
        if nd > 2:
            base = base + ijk[2] * npx * npy
        if nd == 1:
            offs = [0, 1]
        elif nd == 2:
            offs = [0, 1, 1 + npx, npx]
        else:
            offs = [0, 1, 1 + npx, npx]
            offs = offs + [o + npx * npy for o in offs]
        return np.stack([base + o for o in offs], axis=1)

    def element_vertices(self):
        """(ne, 2^nd, nd) vertex coordinates in the reference's element node order."""
        return self.nodes[self.nodeinds - 1]


def apply_periodicBCs(mesh, *pairs):
    return mesh.apply_periodicBCs(*pairs)


def partition_offsets(ne, nranks):
    """Contiguous ranges of the reference's global element order:
    rank r owns [offsets[r], offsets[r+1])."""
    base, rem = divmod(int(ne), int(nranks))
    sizes = [base + (1 if r < rem else 0) for r in range(nranks)]
    return np.concatenate(([0], np.cumsum(sizes))).astype(np.int64)
