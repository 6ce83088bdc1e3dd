"""Connectivity must be BIT-EXACT with the reference: the product's vectorised builder vs the
oracle's literal restatement of CartesianMesh.jl / Mesh.jl, plus the closed forms of
SURVEY.md 10.B for fully periodic meshes.  CPU only."""
import numpy as np
import pytest

import flou_b200 as F
from oracle import connectivity as cn

MESHES = [
    (1, (7,), [("1", "2")]), (1, (5,), []),
    (2, (5, 3), [("1", "2"), ("3", "4")]), (2, (4, 3), [("3", "4")]), (2, (4, 3), [("1", "2")]),
    (2, (4, 3), []), (2, (1, 1), [("1", "2"), ("3", "4")]),
    (3, (4, 3, 2), [("1", "2"), ("3", "4"), ("5", "6")]), (3, (3, 2, 4), [("3", "4")]),
    (3, (2, 2, 2), [("1", "2"), ("5", "6")]), (3, (2, 3, 2), []),
    (3, (1, 1, 1), [("1", "2"), ("3", "4"), ("5", "6")]), (3, (5, 1, 2), [("5", "6"), ("1", "2")]),
]


def _pair(nd, n, per):
    lo, hi = [0.0] * nd, [1.0 + d for d in range(nd)]
    m = F.CartesianMesh(nd, lo, hi, n)
    m.apply_periodicBCs(*per)
    o = cn.cartesian_mesh(lo, hi, n)
    cn.apply_periodic_bcs(o, *per)
    return m, o


@pytest.mark.parametrize("nd,n,per", MESHES, ids=lambda v: str(v))
def test_product_matches_literal_restatement(nd, n, per):
    m, o = _pair(nd, n, per)
    assert np.array_equal(m.faceinds, np.array(o.faceinds))
    assert np.array_equal(m.facepos, np.array(o.facepos))
    assert np.array_equal(m.eleminds, np.array(o.eleminds))
    assert np.array_equal(m.elempos, np.array(o.elempos))
    assert np.array_equal(m.orientation, np.array(o.orientation, dtype=np.uint8))
    assert np.array_equal(m.intfaces, np.array(o.intfaces))
    assert len(m.bdfaces) == len(o.bdfaces)
    for a, b in zip(m.bdfaces, o.bdfaces):
        assert np.array_equal(a, np.array(b))
    assert m.bdmap == o.bdmap and m.periodic == o.periodic
    assert np.array_equal(m.nodeinds, np.array(o.enodes))
    assert np.array_equal(m.nodes, o.nodes)
    assert m.dx == o.dx


@pytest.mark.parametrize("n", [(5, 3), (4, 3, 2), (3, 3, 3), (2, 5, 4)], ids=str)
def test_closed_forms_fully_periodic(n):
    """low-d face of element e has id d*N + e; seam faces keep the first
    element as master with elempos 2d-1; orientation 0."""
    nd = len(n)
    per = [(str(2 * d + 1), str(2 * d + 2)) for d in range(nd)]
    m, _ = _pair(nd, n, per)
    N = int(np.prod(n))
    assert m.nfaces == nd * N and len(m.bdfaces) == 0
    assert np.array_equal(m.intfaces, np.arange(1, nd * N + 1))
    idx = np.unravel_index(np.arange(N), n, order="F")
    e = np.arange(1, N + 1)
    for d in range(nd):
        assert np.array_equal(m.faceinds[:, 2 * d], d * N + e)
        up = list(idx)
        up[d] = (idx[d] + 1) % n[d]
        enext = np.ravel_multi_index(up, n, order="F") + 1
        assert np.array_equal(m.faceinds[:, 2 * d + 1], d * N + enext)
        first, last = idx[d] == 0, idx[d] == n[d] - 1
        assert np.array_equal(m.facepos[:, 2 * d], np.where(first, 1, 2))
        assert np.array_equal(m.facepos[:, 2 * d + 1], np.where(last, 2, 1))
        f = d * N + e - 1                       # low-d face of every element
        # seam: master is the element itself on its low side, slave the wrap-around element
        assert np.all(m.eleminds[f[first], 0] == e[first])
        assert np.all(m.elempos[f[first], 0] == 2 * d + 1)
        assert np.all(m.elempos[f[first], 1] == 2 * d + 2)
        # interior: master is the lower neighbour through its high side
        assert np.all(m.eleminds[f[~first], 1] == e[~first])
        assert np.all(m.elempos[f[~first], 0] == 2 * d + 2)
        assert np.all(m.elempos[f[~first], 1] == 2 * d + 1)
    assert not m.orientation.any()


def test_partition_offsets_are_contiguous_ranges():
    for ne, p in [(10, 3), (32768, 8), (7, 7), (5, 1)]:
        off = F.partition_offsets(ne, p)
        assert off[0] == 0 and off[-1] == ne and np.all(np.diff(off) >= ne // p)
        assert np.all(np.diff(off) <= ne // p + 1)


def test_mesh_argument_errors():
    with pytest.raises(ValueError):
        F.CartesianMesh(2, (0, 0), (1, 1), (3,))
    with pytest.raises(ValueError):
        F.CartesianMesh(2, (0, 1), (1, 1), (3, 3))
    with pytest.raises(ValueError):
        F.CartesianMesh(4, (0,) * 4, (1,) * 4, (1,) * 4)
    m = F.CartesianMesh(2, (0, 0), (1, 1), (3, 3))
    with pytest.raises(ValueError):
        m.apply_periodicBCs(("1", "3"))
    with pytest.raises(ValueError):
        m.apply_periodicBCs(("a", "b"))
    with pytest.raises(ValueError):
        m.apply_periodicBCs(("5", "6"))
