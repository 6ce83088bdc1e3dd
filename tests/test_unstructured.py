"""Config 5: 2-D Euler on unstructured quad meshes (MSH 4.1) with wall / farfield BCs.
CPU: the product's reader + gmsh numbering rule vs the oracle's literal restatement of
GmshMesh.jl, on the reference's own fixtures (when /root/reference is mounted) and on a synthetic
mesh.  GPU: RHS and N-step parity on those meshes at p=5."""
import os

import numpy as np
import pytest

import flou_b200 as F
from common import random_state, relerr
from oracle import gmshmesh as ogm
from unstructured import REF_MESHES, build_pair, euler_bcs, synthetic_raw


def _same_topology(m, o):
    assert np.array_equal(m.faceinds, np.array(o.faceinds))
    assert np.array_equal(m.facepos, np.array(o.facepos))
    assert np.array_equal(m.eleminds, np.array(o.eleminds))
    assert np.array_equal(m.elempos, np.array(o.elempos))
    assert np.array_equal(m.orientation, np.array(o.orientation, dtype=np.uint8))
    assert np.array_equal(m.intfaces, np.array(o.intfaces))
    assert m.bdnames == o.bdnames and len(m.bdfaces) == len(o.bdfaces)
    for a, b in zip(m.bdfaces, o.bdfaces):
        assert np.array_equal(a, np.array(b))
    assert np.array_equal(m.nodeinds, np.array(o.enodes)) and np.array_equal(m.nodes, o.nodes)


@pytest.fixture(scope="module")
def synthetic_msh(tmp_path_factory):
    path = str(tmp_path_factory.mktemp("msh") / "synthetic.msh")
    F.write_msh(synthetic_raw(), path)
    return path


def test_synthetic_mesh_topology_matches_oracle(synthetic_msh):
    m, o = F.UnstructuredMesh(2, synthetic_msh), ogm.unstructured_mesh_2d(synthetic_msh)
    _same_topology(m, o)
    assert m.nelements == 20 and m.nfaces == (4 * 20 + 18) // 2
    assert 0 < int(m.orientation.sum()) < m.nfaces            # both orientations occur
    assert set(np.unique(m.elempos[m.eleminds[:, 1] != 0])) == {1, 2, 3, 4}
    # Flou's invariant: a boundary face id is its line-element tag
    assert np.array_equal(np.sort(np.concatenate(m.bdfaces)), np.arange(1, 19))


@pytest.mark.parametrize("name,ne,nf", [("2D_cylinder", 73, 165), ("2D_wedge_wing", 2700, 5520)])
def test_reference_fixture_topology(name, ne, nf):
    path = os.path.join(REF_MESHES, name + ".msh")
    if not os.path.exists(path):
        pytest.skip("reference fixtures are only mounted in the build container")
    m, o = F.UnstructuredMesh(2, path), ogm.unstructured_mesh_2d(path)
    _same_topology(m, o)
    assert (m.nelements, m.nfaces) == (ne, nf)                # N_f = (4 N_quads + N_lines)/2


def test_refinement_keeps_the_boundary_tag_rule(synthetic_msh, tmp_path):
    raw = F.refine(F.read_msh(synthetic_msh), 3)
    assert len(raw.quads) == 9 * 20 and len(raw.lines) == 3 * 18
    path = str(tmp_path / "refined.msh")
    F.write_msh(raw, path)
    m, o = F.UnstructuredMesh(2, path), ogm.unstructured_mesh_2d(path)
    _same_topology(m, o)
    m2 = F.UnstructuredMesh(2, synthetic_msh, refinement=3)
    assert np.array_equal(m2.faceinds, m.faceinds) and np.allclose(m2.nodes, m.nodes)
    # refined mesh covers the same area
    v = m.element_vertices()
    area = 0.5 * np.abs(np.sum(v[:, :, 0] * np.roll(v[:, :, 1], -1, axis=1)
                               - np.roll(v[:, :, 0], -1, axis=1) * v[:, :, 1], axis=1)).sum()
    assert abs(area - 2.0) < 1e-12


def test_reader_rejects_what_flou_cannot_use(tmp_path, synthetic_msh):
    raw = F.read_msh(synthetic_msh)
    bad = F.RawMesh(raw.nodes, raw.quads, raw.lines, raw.line_tags + 5, raw.line_entity, raw.groups)
    with pytest.raises(ValueError):
        F.UnstructuredMesh(2, bad)
    with pytest.raises(ValueError):
        F.UnstructuredMesh(3, synthetic_msh)


@pytest.mark.gpu
@pytest.mark.parametrize("npn,nf,avg,op", [(6, "mat", "cha", "split"), (4, "lxf", "std", "strong"),
                                           (5, "sca", "cha", "split")])
def test_rhs_on_synthetic_unstructured_mesh(gpu, synthetic_msh, npn, nf, avg, op):
    names = ["Bottom", "Right", "Top", "Left"]
    orc, disc, eq = build_pair(synthetic_msh, npn, euler_bcs(names), nf=nf, avg=avg, op=op)
    Q = random_state(orc.ndof, 2, "euler")
    dQ = disc.new_state()
    F.rhs(dQ, Q, F.EquationConfig(disc, eq), 0.0)
    assert relerr(dQ, orc.rhs(Q)) <= 1e-12
    disc.close()


@pytest.mark.gpu
@pytest.mark.parametrize("npn,op,nodes", [(5, "hybrid", "GLL"), (4, "split", "GL")])
def test_subgrid_operators_on_synthetic_unstructured_mesh(gpu, synthetic_msh, npn, op, nodes):
    """Hybrid operator and Gauss-node split form on the unstructured mesh (rotated element node
    lists, both face orientations, slip walls / inflow / outflow): curved sub-grid frames from the
    mapping (PhysicalRegions.jl:179-292) on every element."""
    names = ["Bottom", "Right", "Top", "Left"]
    orc, disc, eq = build_pair(synthetic_msh, npn, euler_bcs(names), nf="mat", avg="cha", op=op, nodes=nodes)
    Q = random_state(orc.ndof, 2, "euler", amp=0.5 if nodes == "GLL" else 0.15)
    dQ = disc.new_state()
    F.rhs(dQ, Q, F.EquationConfig(disc, eq), 0.0)
    assert relerr(dQ, orc.rhs(Q)) <= 1e-12
    disc.close()


@pytest.mark.gpu
def test_config5_state_after_n_steps(gpu, synthetic_msh, tmp_path):
    """p=5 (np=6), EC split form + matrix dissipation, slip walls + inflow/outflow, refined 2x2."""
    import oracle as O
    path = str(tmp_path / "refined.msh")
    F.write_msh(F.refine(F.read_msh(synthetic_msh), 2), path)
    orc, disc, eq = build_pair(path, 6, euler_bcs(["Bottom", "Right", "Top", "Left"]))
    Q = np.asfortranarray(np.tile(np.array([1.0, 0.45, 0.05, 2.8]), (orc.ndof, 1)))
    Q += 0.02 * random_state(orc.ndof, 2, "euler", amp=0.3)
    ref = orc.lsrk2n(Q, O.ORK256, 2e-4, 20)
    u = Q.copy(order="F")
    sol, _ = F.timeintegrate(u, disc, eq, F.ORK256(), 20 * 2e-4, dt=2e-4)
    assert sol is not None and relerr(sol.u[-1], ref) <= 1e-10
    disc.close()


def _golden(name):
    return np.load(os.path.join(os.path.dirname(__file__), "golden", name), allow_pickle=False)


def _cylinder_raw(g):
    groups = [(str(n), [int(v) for v in str(e).split(",")])
              for n, e in zip(g["group_names"], g["group_entities"])]
    return F.RawMesh(g["nodes"], g["quads"], g["lines"], g["line_tags"], g["line_entity"], groups)


def test_golden_cylinder_mesh_tables_are_consistent():
    g = _golden("cylinder_p5.npz")
    m = F.UnstructuredMesh(2, _cylinder_raw(g))
    assert (m.nelements, m.nfaces) == (73, 165) and int(m.orientation.sum()) == 60
    assert m.bdnames == ["Bottom", "Right", "Top", "Left", "Hole"]
    assert g["Q"].shape == (73 * 36, 4)


@pytest.mark.gpu
def test_config5_reference_cylinder_mesh_against_golden_and_oracle(gpu, tmp_path):
    """The reference's own 2D_cylinder mesh (committed as parsed tables), p=5, wall/farfield BCs."""
    g = _golden("cylinder_p5.npz")
    raw = _cylinder_raw(g)
    path = str(tmp_path / "cylinder.msh")
    F.write_msh(raw, path)
    orc, disc, eq = build_pair(path, 6, euler_bcs([n for n, _ in raw.groups]))
    Q = np.asfortranarray(g["Q"])
    dQ = disc.new_state()
    F.rhs(dQ, Q, F.EquationConfig(disc, eq), 0.0)
    assert relerr(dQ, g["dQ"]) <= 1e-12            # committed golden vector
    assert relerr(dQ, orc.rhs(Q)) <= 1e-12         # live oracle
    disc.close()
