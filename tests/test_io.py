"""VTKHDF snapshot path (second half): FlouBiz.jl:25-113, FlouSpatial/IO.jl:16-97,
FlouTime.jl:67-90.

CPU: the HDF5 container (flou_b200.hdf5min) read back by the independent parser tests/hdf5_reader.py;
known-answer VTK Lagrange cell orderings derived by hand from the reference's formulas; the mesh
datasets of open_for_write against the oracle restatement (oracle/io.py), dataset by dataset.
GPU: the projection to equispaced nodes (device kernel) against the oracle's dense Kronecker
product, and the save callback's files read back and compared with the oracle's content.
"""
import os
import struct

import numpy as np
import pytest

from common import Case, random_state, relerr, smooth_state
from hdf5_reader import Reader

EC = dict(nodes="GLL", eq="euler", op="split", nf="mat", avg="cha")
CASES = [
    Case(1, (7,), 4, nodes="GLL", eq="adv", op="strong", nf="lxf", avg="std"),
    Case(2, (5, 4), 4, **EC),
    Case(2, (4, 3), 5, nodes="GL", eq="euler", op="strong", nf="lxf", avg="std"),
    Case(3, (3, 2, 3), 3, **EC),
    Case(3, (3, 3, 2), 4, perturb_amp=0.08, periodic=[], bcs={str(i): ("slip", None) for i in range(1, 7)}, **EC),
]


# ----------------------------------------------------------------------------- container
def test_hdf5_container_round_trip(tmp_path):
    from flou_b200 import hdf5min
    rng = np.random.default_rng(5)
    path = str(tmp_path / "t.hdf")
    f = hdf5min.File(path)
    g = f.create_group("/VTKHDF")
    g.attrs["Version"] = np.array([1, 0])
    g.attrs["Type"] = "UnstructuredGrid"
    want = {
        "/VTKHDF/Points": rng.random((11, 3)),
        "/VTKHDF/Types": np.full(4, 72, np.uint8),
        "/VTKHDF/Offsets": np.arange(5, dtype=np.int64) * 27,
        "/VTKHDF/CellData/Region": np.ones(4, np.int64),
        "/VTKHDF/FieldData/Time": np.array([0.125]),
        "/VTKHDF/f32": rng.random(5).astype(np.float32),
        "/VTKHDF/i32": np.arange(-3, 4, dtype=np.int32),
        "/VTKHDF/empty": np.zeros(0),
    }
    # more links than one symbol-table node holds (2 x leaf K = 32): several nodes under one B-tree
    for i in range(75):
        want["/VTKHDF/PointData/var%03d" % i] = rng.random(6)
    for k, v in want.items():
        f.write(k, v)
    with pytest.raises(ValueError):
        f.write("/VTKHDF/Points", np.zeros(3))          # HDF5.jl: the name already exists
    f.close()
    r = Reader(path)
    got = r.tree()
    assert sorted(got) == sorted(want)
    for k, v in want.items():
        assert got[k].dtype == v.dtype and got[k].shape == v.shape and np.array_equal(got[k], v), k
    a = r.get("/VTKHDF").attrs
    assert a["Type"] == "UnstructuredGrid" and a["Version"].dtype == np.int64 and list(a["Version"]) == [1, 0]
    # fixed points of the format: signature, superblock version 0, 8-byte offsets, EOF address = size
    raw = open(path, "rb").read()
    assert raw[:8] == b"\x89HDF\r\n\x1a\n" and raw[8] == 0 and raw[13] == 8 and raw[14] == 8
    assert struct.unpack("<Q", raw[40:48])[0] == len(raw) == os.path.getsize(path)


def test_hdf5_writer_rejects_what_it_cannot_store(tmp_path):
    from flou_b200 import hdf5min
    f = hdf5min.File(str(tmp_path / "t.hdf"))
    with pytest.raises(TypeError):
        f.write("/a", np.array(["x", "y"]))
    with pytest.raises(TypeError):
        f.write("/b", np.zeros(3, dtype=np.complex128))
    f.write("/g/d", np.zeros(2))
    with pytest.raises(ValueError):
        f.write("/g/d/e", np.zeros(2))                  # a dataset is not a group
    f.close()
    with pytest.raises(ValueError):
        f.write("/late", np.zeros(1))


# ----------------------------------------------------------------------------- cell ordering
def test_vtk_connectivities_known_answers():
    """Hand-derived from StdSegment.jl:171-175, StdQuad.jl:185-195, StdHex.jl:174-199 on the
    x-fastest node grid (0-based)."""
    import oracle.io as OIO
    assert OIO.vtk_connectivities(1, 4) == [0, 3, 1, 2]
    assert OIO.vtk_connectivities(2, 2) == [0, 1, 3, 2]
    assert OIO.vtk_connectivities(2, 3) == [0, 2, 8, 6, 1, 5, 7, 3, 4]
    assert OIO.vtk_connectivities(2, 4) == [0, 3, 15, 12, 1, 2, 7, 11, 13, 14, 4, 8, 5, 6, 9, 10]
    assert OIO.vtk_connectivities(3, 2) == [0, 1, 3, 2, 4, 5, 7, 6]
    assert OIO.vtk_connectivities(3, 3) == [0, 2, 8, 6, 18, 20, 26, 24,            # corners
                                            1, 5, 7, 3, 19, 23, 25, 21, 9, 11, 17, 15,   # edges
                                            12, 14, 10, 16, 4, 22,                  # faces
                                            13]                                     # interior


@pytest.mark.parametrize("nd", [1, 2, 3])
@pytest.mark.parametrize("n", [2, 3, 4, 5, 6, 8])
def test_vtk_connectivities_product_matches_oracle(nd, n):
    import flou_b200 as F
    import oracle.io as OIO
    basis = F.LagrangeBasis("GLL", n)
    std = {1: F.StdSegment, 2: F.StdQuad, 3: F.StdHex}[nd](basis, F.DGSEMrec(basis), 1)
    conn = F.vtk_connectivities(std)
    assert conn.dtype == np.int64 and list(conn) == OIO.vtk_connectivities(nd, n)
    assert sorted(conn) == list(range(n ** nd))         # a permutation of the element's nodes
    assert F.vtk_type(std) == {1: 68, 2: 70, 3: 72}[nd] and F.vtk_type(std).dtype == np.uint8


def test_equispaced_nodes_and_matrix_match_oracle():
    import flou_b200 as F
    import oracle.io as OIO
    for nodes, n, neq in (("GLL", 4, None), ("GL", 5, None), ("GLL", 3, 7), ("GL", 2, 1)):
        basis = F.LagrangeBasis(nodes, n)
        std = F.StdQuad(basis, F.DGSEMrec(basis), 1, nequispaced=neq)
        xe, M = OIO.equispaced_1d(nodes, n, neq)
        assert np.array_equal(std.xe1d, xe) and np.allclose(std.node2eq1d, M, rtol=0, atol=1e-14)
        assert std.nequispaced() == len(xe) ** 2
        # interpolation of the constant: rows sum to one
        assert np.allclose(std.node2eq1d.sum(axis=1), 1.0, atol=1e-13)


# ----------------------------------------------------------------------------- mesh datasets
@pytest.mark.parametrize("case", CASES, ids=lambda c: repr(c))
def test_open_for_write_mesh_datasets_match_oracle(case, tmp_path):
    import flou_b200 as F
    import oracle.io as OIO
    orc = case.oracle()
    disc, eq = case.product(create=False)
    path = str(tmp_path / "mesh.hdf")
    file = F.open_for_write(path, disc)
    assert isinstance(file, F.FlouFile) and file.name == path
    F.add_celldata(file, np.arange(orc.ne, dtype=np.float64), "Indicator")
    F.add_fielddata(file, [0.75], "Time")
    F.close_file(file)
    want, attrs = OIO.mesh_datasets(orc, case.nodes)
    want["/VTKHDF/CellData/Indicator"] = np.arange(orc.ne, dtype=np.float64)
    want["/VTKHDF/FieldData/Time"] = np.array([0.75])
    r = Reader(path)
    got = r.tree()
    assert sorted(got) == sorted(want)
    for k, v in want.items():
        assert got[k].dtype == v.dtype and got[k].shape == v.shape, k
        if v.dtype.kind == "f":
            assert np.allclose(got[k], v, rtol=0, atol=1e-14), k
        else:
            assert np.array_equal(got[k], v), k
    a = r.get("/VTKHDF").attrs
    assert a["Type"] == attrs["Type"] and np.array_equal(a["Version"], attrs["Version"])
    # every cell's connectivity addresses its own block of points
    off, conn = got["/VTKHDF/Offsets"], got["/VTKHDF/Connectivity"]
    for e in range(orc.ne):
        assert sorted(conn[off[e]:off[e + 1]]) == list(range(off[e], off[e + 1]))


def test_partitioned_disc_writes_its_own_elements(tmp_path):
    import flou_b200 as F
    import oracle.io as OIO
    case = CASES[1]
    orc = case.oracle()
    want, _ = OIO.mesh_datasets(orc, case.nodes)
    neq = case.np ** case.nd
    pieces = []
    for rank in range(3):
        disc, _ = case.product(rank=rank, nranks=3, create=False)
        path = str(tmp_path / f"p{rank}.hdf")
        F.close_file(F.open_for_write(path, disc))
        got = Reader(path).tree()
        ne = disc.elem_end - disc.elem_begin
        assert got["/VTKHDF/NumberOfCells"][0] == ne and got["/VTKHDF/Offsets"][-1] == ne * neq
        pieces.append(got["/VTKHDF/Points"])
    assert np.allclose(np.concatenate(pieces), want["/VTKHDF/Points"], rtol=0, atol=1e-14)


# ----------------------------------------------------------------------------- device projection
@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=lambda c: repr(c))
def test_pointdata2vtkhdf_matches_oracle(gpu, case):
    import flou_b200 as F
    import oracle.io as OIO
    orc = case.oracle()
    disc, eq = case.product()
    Q = np.asfortranarray(0.9 * smooth_state(orc.coords, case.nd, case.eq)
                          + 0.1 * random_state(orc.ndof, case.nd, case.eq, amp=0.3))
    got = F.pointdata2VTKHDF(Q, disc)
    want = OIO.pointdata(orc, case.nodes, Q)
    assert len(got) == len(want) == disc.nv
    for g, w in zip(got, want):
        assert g.shape == w.shape and relerr(g, w) <= 1e-13
    # Q = None projects the device-resident state
    disc.upload(Q)
    for g, w in zip(F.pointdata2VTKHDF(None, disc), want):
        assert relerr(g, w) <= 1e-13
    if case.nodes == "GLL":
        # Lobatto nodes contain the element's end points: the corner values are reproduced
        assert relerr(got[0][::case.np ** case.nd], np.asarray(Q).reshape(orc.ndof, -1, order="F")[::orc.npts, 0]) <= 1e-13
    disc.close()


@pytest.mark.gpu
def test_projection_with_more_equispaced_nodes_than_solution_nodes(gpu):
    import flou_b200 as F
    import oracle.io as OIO
    case = Case(2, (6, 5), 4, **EC)
    orc = case.oracle()
    mesh = F.CartesianMesh(2, (0.0, 0.0), (2.0, 2.0), case.n)      # geometry is irrelevant to the projection
    mesh.apply_periodicBCs(*case.periodic)
    eq = F.EulerEquation(2, case.gamma)
    basis = F.LagrangeBasis("GLL", 4)
    std = F.StdQuad(basis, F.DGSEMrec(basis), eq.nv, nequispaced=9)
    disc = F.MultielementDisc(mesh, std, eq, F.SplitDivOperator(F.MatrixDissipation(F.ChandrasekharAverage(), 1.0)), {})
    Q = random_state(orc.ndof, 2, "euler", amp=0.3)
    got = F.pointdata2VTKHDF(Q, disc)
    want = OIO.pointdata(orc, "GLL", Q, nequispaced=9)
    for g, w in zip(got, want):
        assert g.shape == (orc.ne * 81,) and relerr(g, w) <= 1e-13
    disc.close()


@pytest.mark.gpu
def test_save_callback_writes_the_reference_files(gpu, tmp_path):
    """get_save_callback(basename; iter) through timeintegrate: a file before the first step
    (`initialize`) and one after every selected step, named basename_%010d.hdf, holding the mesh, the
    time as field data and the solution per variable name (FlouTime.jl:67-90)."""
    import flou_b200 as F
    import oracle as O
    import oracle.io as OIO
    case = Case(3, (4, 3, 3), 4, **EC)
    orc = case.oracle()
    disc, eq = case.product()
    Q0 = np.asfortranarray(0.9 * smooth_state(orc.coords, 3, "euler") + 0.1 * random_state(orc.ndof, 3, "euler", amp=0.3))
    dt, base = 1e-4, str(tmp_path / "tgv")
    cb = F.get_save_callback(base, iter=(2, 3))
    sol, _ = F.timeintegrate(Q0.copy(order="F"), disc, eq, F.ORK256(), 3 * dt, dt=dt, callback=F.make_callback_list(cb))
    assert sol is not None
    names = sorted(os.listdir(tmp_path))
    assert names == ["tgv_0000000000.hdf", "tgv_0000000002.hdf", "tgv_0000000003.hdf"]
    assert [os.path.basename(f) for f in cb.files] == names
    mesh_want, attrs = OIO.mesh_datasets(orc, case.nodes)
    for name, nsteps in zip(names, (0, 2, 3)):
        r = Reader(str(tmp_path / name))
        got = r.tree()
        ref = Q0 if nsteps == 0 else orc.lsrk2n(Q0, O.ORK256, dt, nsteps)
        want = dict(mesh_want)
        want["/VTKHDF/FieldData/Time"] = np.array([nsteps * dt])
        for vname, vec in zip(OIO.variablenames(True, 3), OIO.pointdata(orc, case.nodes, ref)):
            want["/VTKHDF/PointData/" + vname] = vec
        assert sorted(got) == sorted(want)
        for k, v in want.items():
            assert got[k].dtype == v.dtype and got[k].shape == v.shape, k
            if k.startswith("/VTKHDF/PointData/"):
                assert relerr(got[k], v) <= 1e-10, k
            elif v.dtype.kind == "f":
                assert np.allclose(got[k], v, rtol=1e-13, atol=1e-14), k
            else:
                assert np.array_equal(got[k], v), k
        assert r.get("/VTKHDF").attrs["Type"] == "UnstructuredGrid"
    disc.close()


def test_container_opens_with_libhdf5_where_available(tmp_path):
    """Not runnable in the build image (no HDF5 library anywhere: the container is pinned only by
    tests/hdf5_reader.py); wherever h5py exists this reads the same file through libhdf5."""
    h5py = pytest.importorskip("h5py")
    import flou_b200 as F
    case = CASES[1]
    disc, _ = case.product(create=False)
    path = str(tmp_path / "mesh.hdf")
    file = F.open_for_write(path, disc)
    F.add_fielddata(file, [0.5], "Time")
    F.close_file(file)
    want = Reader(path).tree()
    with h5py.File(path, "r") as f:
        assert f["VTKHDF"].attrs["Type"] in (b"UnstructuredGrid", "UnstructuredGrid")
        assert list(f["VTKHDF"].attrs["Version"]) == [1, 0]
        for k, v in want.items():
            assert np.array_equal(f[k][...], v), k
