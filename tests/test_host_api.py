"""Host mirror of the reference's operator/plugin interface: same names, argument meaning and
error behaviour (ArgumentError -> ValueError).  CPU only (create=False: no device handle)."""
import numpy as np
import pytest

import flou_b200 as F
from flou_b200 import _lib as L


def _std(nd, n=4, nodes="GLL", nv=None):
    b = F.LagrangeBasis(nodes, n)
    return {1: F.StdSegment, 2: F.StdQuad, 3: F.StdHex}[nd](b, F.DGSEMrec(b), nv or nd + 2)


def test_equation_types():
    assert F.nvariables(F.LinearAdvection(2.0, -1.0)) == 1 and F.spatialdim(F.LinearAdvection(1.0)) == 1
    assert F.nvariables(F.EulerEquation(3, 1.4)) == 5
    assert F.EulerEquation(2, 1.4).variablenames() == ("rho", "rhou", "rhov", "rhoe")
    with pytest.raises(ValueError):
        F.EulerEquation(4, 1.4)
    with pytest.raises(ValueError):
        F.LinearAdvection()
    with pytest.raises(ValueError):
        F.EulerInflowBC([1.0, 2.0])


def test_split_operator_default_two_point_flux():
    nf = F.MatrixDissipation(F.ChandrasekharAverage(), 1.0)
    op = F.SplitDivOperator(nf)                  # tpflux = numflux.avg (OpDivergence.jl:192-194)
    assert isinstance(op.tpflux, F.ChandrasekharAverage) and op.numflux is nf
    op2 = F.SplitDivOperator(F.StdAverage(), nf)
    assert isinstance(op2.tpflux, F.StdAverage)
    with pytest.raises(ValueError):
        F.SplitDivOperator(F.LxF(F.StdAverage(), 1.0), nf)


def test_descriptor_carries_the_operator_choice():
    mesh = F.CartesianMesh(2, (0, 0), (1, 1), (3, 3)).apply_periodicBCs(("1", "2"), ("3", "4"))
    eq = F.EulerEquation(2, 1.4)
    d = F.MultielementDisc(mesh, _std(2), eq, F.SplitDivOperator(F.MatrixDissipation(
        F.ChandrasekharAverage(), 0.7)), {}, create=False)._desc
    assert (d.divop, d.tpflux, d.numflux, d.numflux_avg) == (
        L.OP_SPLIT, L.FLUX_CHANDRASEKHAR, L.FLUX_MATRIXDISSIPATION, L.FLUX_CHANDRASEKHAR)
    assert d.intensity == 0.7 and d.gamma == 1.4 and d.geometry == L.GEOM_CARTESIAN
    d = F.MultielementDisc(mesh, _std(2, nodes="GL"), eq, F.StrongDivOperator(
        F.LxF(F.StdAverage(), 1.0)), {}, create=False)._desc
    assert (d.divop, d.numflux, d.numflux_avg) == (L.OP_STRONG, L.FLUX_LXF, L.FLUX_STDAVERAGE)


def test_boundary_conditions_are_ordered_by_bdmap():
    mesh = F.CartesianMesh(2, (0, 0), (1, 1), (3, 2)).apply_periodicBCs(("1", "2"))
    eq = F.EulerEquation(2, 1.4)
    bcs = {"4": F.EulerSlipBC(), "3": F.EulerInflowBC([1.0, 0.1, 0.0, 2.5])}
    disc = F.MultielementDisc(mesh, _std(2), eq, F.StrongDivOperator(F.LxF(F.StdAverage(), 1.0)),
                              bcs, create=False)
    assert [bc.kind for bc in disc.bcs] == [L.BC_INFLOW, L.BC_SLIP]
    assert np.array_equal(disc._keep["bc_state"][0], [1.0, 0.1, 0.0, 2.5])
    with pytest.raises(ValueError):        # "The number of BCs does not match ..."
        F.MultielementDisc(mesh, _std(2), eq, F.StrongDivOperator(F.StdAverage()), {}, create=False)


def test_generic_bc_is_tabulated_at_face_nodes():
    mesh = F.CartesianMesh(1, 0, 1, 4)
    eq = F.EulerEquation(1, 1.4)

    def qext(_, x, __, ___, e):
        return F.vars_prim2cons((1.0, 0.0, 100.0) if x[0] < 0.5 else (0.125, 0.0, 10.0), e)
    disc = F.MultielementDisc(mesh, _std(1), eq, F.SplitDivOperator(F.MatrixDissipation(
        F.ChandrasekharAverage(), 1.0)), {"1": F.GenericBC(qext), "2": F.GenericBC(qext)},
        create=False)
    table = disc._keep["bc_table"]
    assert np.allclose(table[0], [1.0, 0.0, 250.0]) and np.allclose(table[1], [0.125, 0.0, 25.0])

    def state_dependent(Qin, x, frame, t, e):
        return 2 * Qin[0]
    # a closure that reads Qin / frame / time is detected and re-tabulated every stage ...
    dyn = F.MultielementDisc(mesh, _std(1), eq, F.StrongDivOperator(F.StdAverage()),
                             {"1": F.GenericBC(state_dependent), "2": F.GenericBC(qext)}, create=False)
    assert [ib for ib, _ in dyn._dynamic_bcs] == [0] and dyn.has_dynamic
    assert not disc.has_dynamic
    # ... unless the caller insists on the static table
    with pytest.raises(ValueError):
        F.MultielementDisc(mesh, _std(1), eq, F.StrongDivOperator(F.StdAverage()),
                           {"1": F.GenericBC(state_dependent, dynamic=False), "2": F.GenericBC(qext)}, create=False)


def test_out_of_scope_features_raise_instead_of_falling_back():
    mesh = F.CartesianMesh(2, (0, 0), (1, 1), (2, 2)).apply_periodicBCs(("1", "2"), ("3", "4"))
    eq = F.EulerEquation(2, 1.4)
    # Gauss-node split form (row f4): Euler only -- the reference has no entropy variables for advection
    F.MultielementDisc(mesh, _std(2, nodes="GL"), eq, F.SplitDivOperator(
        F.MatrixDissipation(F.ChandrasekharAverage(), 1.0)), {}, create=False)
    with pytest.raises(ValueError):
        F.MultielementDisc(mesh, _std(2, nodes="GL", nv=1), F.LinearAdvection(1.0, 0.5),
                           F.SplitDivOperator(F.StdAverage(), F.LxF(F.StdAverage(), 1.0)), {}, create=False)
    # source terms (row a13): a plain callable is re-tabulated every stage, Source(...) says more
    d1 = F.MultielementDisc(mesh, _std(2), eq, F.StrongDivOperator(F.StdAverage()), {},
                            source=lambda Q, x, t: None, create=False)
    assert d1.has_dynamic and d1.source.state and d1.source.time
    d2 = F.MultielementDisc(mesh, _std(2), eq, F.StrongDivOperator(F.StdAverage()), {},
                            source=F.Source(lambda Q, x, t: [0.0, 1.0, 0.0, 0.0], state=False, time=False),
                            create=False)
    assert not d2.has_dynamic
    with pytest.raises(ValueError):
        F.MultielementDisc(mesh, _std(2), eq, F.StrongDivOperator(F.StdAverage()), {}, source=3.0, create=False)
    with pytest.raises(ValueError):
        F.ORK256(williamson_condition=True)
    with pytest.raises(ValueError):
        F.MultielementDisc(F.CartesianMesh(1, 0, 1, 3), _std(2), eq,
                           F.StrongDivOperator(F.StdAverage()), {}, create=False)


def test_rk_tableaus():
    assert F.ORK256().nstages == 5 and F.CarpenterKennedy2N54().nstages == 5
    assert F.ORK256.B[0] == 0.2 and F.ORK256.A[2] == -1.55798
    assert abs(sum(F.CarpenterKennedy2N54.c) - 2.1) < 0.2


def test_helpers_match_reference_formulas():
    eq = F.EulerEquation(2, 1.4)
    Q = F.vars_prim2cons((1.2, 0.5, -0.25, 2.0), eq)
    assert np.allclose(Q, [1.2, 0.6, -0.3, 2.0 / 0.4 + 0.5 * 1.2 * (0.25 + 0.0625)])
    assert abs(F.gaussian_bump(0.5, 0.5, 0.1, 2.0) - 2.0) < 1e-15
    assert abs(F.gaussian_bump(0.6, 0.5, 0.5, 0.5, 0.1, 0.1, 1.0) - np.exp(-0.5)) < 1e-15
    r1, u1, p1 = F.normal_shockwave(1.0, 2.0 * F.soundvelocity(1.0, 1.0, eq), 1.0, eq)
    assert abs(r1 - 2.4 * 4 / (0.4 * 4 + 2)) < 1e-14 and abs(p1 - 4.5) < 1e-14


def test_coords_follow_the_reference_node_order():
    mesh = F.CartesianMesh(2, (0, 0), (2, 1), (2, 1))
    disc = F.MultielementDisc(mesh, _std(2, n=3), F.EulerEquation(2, 1.4),
                              F.StrongDivOperator(F.StdAverage()),
                              {str(i): F.EulerOutflowBC() for i in range(1, 5)}, create=False)
    x = disc.coords()
    assert x.shape == (18, 2)
    assert np.allclose(x[:3, 0], [0.0, 0.5, 1.0]) and np.allclose(x[:3, 1], 0.0)   # x fastest
    assert np.allclose(x[9:12, 0], [1.0, 1.5, 2.0])                                 # element 2


def test_hybrid_operator_constructors_and_descriptor():
    """HybridDivOperator([tpflux], numflux, blend) (OpDivergence.jl:452-477): tpflux defaults to
    numflux.avg, fvflux = numflux; refused for linear advection."""
    nf = F.MatrixDissipation(F.ChandrasekharAverage(), 1.0)
    op = F.HybridDivOperator(nf, 0.25)
    assert isinstance(op.tpflux, F.ChandrasekharAverage) and op.fvflux is nf and op.numflux is nf
    assert isinstance(F.HybridDivOperator(F.StdAverage(), nf, 1.0).tpflux, F.StdAverage)
    with pytest.raises(TypeError):
        F.HybridDivOperator(F.StdAverage(), 1.0)          # no `.avg` to read the two-point flux from
    with pytest.raises(ValueError):
        F.HybridDivOperator(F.LxF(F.StdAverage(), 1.0), nf, 1.0)
    mesh = F.CartesianMesh(2, (0, 0), (1, 1), (3, 3)).apply_periodicBCs(("1", "2"), ("3", "4"))
    eq = F.EulerEquation(2, 1.4)
    d = F.MultielementDisc(mesh, _std(2), eq, op, {}, create=False)._desc
    assert (d.divop, d.tpflux, d.numflux, d.blend) == (
        L.OP_HYBRID, L.FLUX_CHANDRASEKHAR, L.FLUX_MATRIXDISSIPATION, 0.25)
    F.MultielementDisc(mesh, _std(2, nodes="GL"), eq, op, {}, create=False)      # Gauss nodes: all-surface form
    with pytest.raises(ValueError):
        F.MultielementDisc(mesh, _std(2, nv=1), F.LinearAdvection(1.0, 1.0), op, {}, create=False)


def test_last_step_is_shortened_to_land_on_tfinal():
    """OrdinaryDiffEq with adaptive=false takes full steps of dt and shortens the last one (tstops);
    an integer number of steps stays one fused call."""
    from flou_b200.time import _split_steps
    assert _split_steps(0.0, 1.0, 0.25) == (4, 0.0)
    assert _split_steps(0.0, 180 * 1e-4, 1e-4) == (180, 0.0)          # Sod KAT: no round-off remainder
    n, last = _split_steps(0.0, 1.0, 0.3)
    assert n == 3 and abs(last - 0.1) < 1e-15
    n, last = _split_steps(0.5, 0.55, 0.1)
    assert n == 0 and abs(last - 0.05) < 1e-15


def test_source_tabulation_scalar_and_vectorised():
    x = np.array([[0.0, 1.0], [0.5, 2.0], [1.0, 3.0]])
    Q = np.arange(12.0).reshape(3, 4)
    f = lambda q, xi, t: [xi[0], q[1], t, 0.0]
    fv = lambda q, xx, t: np.stack([xx[:, 0], q[:, 1], np.full(len(xx), t), np.zeros(len(xx))], axis=1)
    a = F.Source(f).tabulate(Q, x, 2.0, 4)
    b = F.Source(fv, vectorized=True).tabulate(Q, x, 2.0, 4)
    assert a.flags.f_contiguous and np.array_equal(a, b)
    assert np.array_equal(a[:, 0], x[:, 0]) and np.array_equal(a[:, 2], [2.0] * 3)
    assert F.Source(f, state=False, time=False).dynamic is False and F.Source(f, state=False).dynamic is True
    # a closure that returns nothing (the reference's default source) adds nothing
    assert not F.Source(lambda q, xi, t: None).tabulate(Q, x, 0.0, 4).any()
