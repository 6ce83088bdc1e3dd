"""VTKHDF snapshots: the reference's FlouBiz module and its FlouSpatial methods.

  FlouFile, open_for_write, add_fielddata!, add_celldata!,        src/FlouBiz/FlouBiz.jl:25-113
  add_pointdata!, add_solution!, close_file!
  open_for_write!(file, disc): mesh datasets of the VTKHDF group   src/FlouSpatial/IO.jl:16-76
  pointdata2VTKHDF(Q, disc): projection to equispaced nodes        src/FlouSpatial/IO.jl:78-97
  vtk_type / vtk_connectivities                                    src/FlouSpatial/StdRegions/StdSegment.jl:169-175,
                                                                   StdQuad.jl:183-195, StdHex.jl:172-199
  get_save_callback                                                src/FlouTime/FlouTime.jl:67-90

The file is a VTKHDF "UnstructuredGrid" (version 1.0): one Lagrange cell (VTK types 68 / 70 / 72)
per element with its equispaced nodes as points.  The projection of the state runs on the device
(flou_b200_project_equispaced); the mesh datasets are host tables built once per file, as in the
reference.  The container is written by flou_b200.hdf5min (no HDF5 library in this image).
A partitioned discretisation writes the elements its rank owns (one valid file per rank).
"""
import numpy as np

from . import _lib as L
from . import hdf5min
from .disc import _ptr, _state
from .geometry import element_coords
from .monitors import _Callback


class FlouFile:
    def __init__(self, name, handler):
        self.name, self.handler = name, handler


def vtk_type(std):
    return np.uint8({1: 68, 2: 70, 3: 72}[std.nd])


def vtk_connectivities(std):
    """0-based point order of a VTK Lagrange cell over the element's x-fastest node grid: corners,
    edges, (faces,) interior, each in the reference's order; n = solution nodes per direction."""
    n, nd = std.np, std.nd
    li = np.arange(n ** nd).reshape((n,) * nd, order="F")       # li[i, j, k], x fastest
    m = slice(1, n - 1)
    e = n - 1
    if nd == 1:
        parts = [[li[0], li[e]], li[m]]
    elif nd == 2:
        parts = [[li[0, 0], li[e, 0], li[e, e], li[0, e]],
                 li[m, 0], li[e, m], li[m, e], li[0, m],
                 li[m, m].reshape(-1, order="F")]
    else:
        parts = [[li[0, 0, 0], li[e, 0, 0], li[e, e, 0], li[0, e, 0],
                  li[0, 0, e], li[e, 0, e], li[e, e, e], li[0, e, e]],
                 li[m, 0, 0], li[e, m, 0], li[m, e, 0], li[0, m, 0],
                 li[m, 0, e], li[e, m, e], li[m, e, e], li[0, m, e],
                 li[0, 0, m], li[e, 0, m], li[e, e, m], li[0, e, m]]
        parts += [a.reshape(-1, order="F") for a in
                  (li[0, m, m], li[e, m, m], li[m, 0, m], li[m, e, m], li[m, m, 0], li[m, m, e])]
        parts.append(li[m, m, m].reshape(-1, order="F"))
    return np.concatenate([np.asarray(p, dtype=np.int64).reshape(-1) for p in parts])


def _local_elements(disc):
    return disc.elem_begin, disc.elem_end


def open_for_write(filename, disc):
    """Create (or rewrite) a VTKHDF file, returning a handler that can be used to add data."""
    fh = hdf5min.File(filename)
    std, mesh = disc.std, disc.mesh
    root = fh.create_group("/VTKHDF")
    root.attrs["Version"] = np.array([1, 0], dtype=np.int64)
    root.attrs["Type"] = "UnstructuredGrid"
    e0, e1 = _local_elements(disc)
    ne, neq = e1 - e0, std.nequispaced()
    verts = mesh.element_vertices()[e0:e1]
    pts = element_coords(verts, std.xe)                         # (ne*neq, nd)
    points = np.zeros((ne * neq, 3))
    points[:, :disc.nd] = pts
    conn1 = vtk_connectivities(std)
    # the reference offsets every cell's connectivity by nequispaced(std) per element
    conn = (conn1[None, :] + neq * np.arange(ne, dtype=np.int64)[:, None]).reshape(-1)
    offsets = neq * np.arange(ne + 1, dtype=np.int64)
    regions = getattr(mesh, "regionmap", None)
    regions = np.ones(ne, dtype=np.int64) if regions is None else np.asarray(regions, dtype=np.int64)[e0:e1]
    fh.write("/VTKHDF/NumberOfPoints", np.array([points.shape[0]], dtype=np.int64))
    fh.write("/VTKHDF/Points", points)
    fh.write("/VTKHDF/NumberOfConnectivityIds", np.array([conn.size], dtype=np.int64))
    fh.write("/VTKHDF/Connectivity", conn)
    fh.write("/VTKHDF/NumberOfCells", np.array([ne], dtype=np.int64))
    fh.write("/VTKHDF/Types", np.full(ne, vtk_type(std), dtype=np.uint8))
    fh.write("/VTKHDF/Offsets", offsets)
    fh.write("/VTKHDF/CellData/Region", regions)
    return FlouFile(filename, fh)


def add_fielddata(file, data, name):
    file.handler.write("/VTKHDF/FieldData/" + name, np.atleast_1d(data))


def add_celldata(file, data, name):
    file.handler.write("/VTKHDF/CellData/" + name, np.asarray(data))


def pointdata2VTKHDF(Q, disc):
    """One vector per variable: the state at the equispaced nodes of every (local) element.
    `Q=None` projects the device-resident state (what the save callback does)."""
    std = disc.std
    npoints = (disc.elem_end - disc.elem_begin) * std.nequispaced()
    out = np.empty((npoints, disc.nv), order="F")
    q = None if Q is None else _ptr(_state(Q, disc.ndofs, disc.nv))
    M = np.ascontiguousarray(std.node2eq1d, dtype=np.float64)
    L.check(L.lib().flou_b200_project_equispaced(disc.handle, q, int(M.shape[0]), _ptr(M), _ptr(out)))
    return [out[:, v] for v in range(disc.nv)]


def add_pointdata(file, data, disc, name):
    file.handler.write("/VTKHDF/PointData/" + name, pointdata2VTKHDF(data, disc)[0])


def add_solution(file, sol, disc, equation):
    for name, data in zip(equation.variablenames(), pointdata2VTKHDF(sol, disc)):
        file.handler.write("/VTKHDF/PointData/" + name, data)


def close_file(file):
    file.handler.close()


class SaveCallback(_Callback):
    """get_save_callback(basename; iter): `basename_%010d.hdf` with the time as field data and the
    solution as point data, also once before the first step (`initialize`)."""
    initialize = True

    def __init__(self, basename, iter=True):
        super().__init__(iter)
        self.basename, self.files = basename, []

    def affect(self, integ):
        disc = integ.disc
        part = f".rank{disc.rank}" if disc.nranks > 1 else ""
        filename = "%s_%010d%s.hdf" % (self.basename, integ.iter, part)
        file = open_for_write(filename, disc)
        add_fielddata(file, [integ.t], "Time")
        add_solution(file, None, disc, integ.equation)
        close_file(file)
        self.files.append(filename)
        print("Saved solution at t=%.7g in `%s`" % (integ.t, filename))


def get_save_callback(basename, *, iter=True):
    return SaveCallback(basename, iter)
