"""GPU parity: the CUDA path through the C ABI vs the CPU oracle on the same seeded inputs.
Bar (BASELINE.json north_star): fp64 RHS within 1e-12 relative; connectivity bit-exact."""
import numpy as np
import pytest

from common import Case, random_state, relerr, smooth_state

RHS_TOL = 1e-12

CASES = [
    # config 1 family: linear advection, strong form, LxF, GL and GLL nodes
    Case(2, (6, 5), 4, nodes="GL", eq="adv", op="strong", nf="lxf", avg="std"),
    Case(2, (6, 5), 4, nodes="GLL", eq="adv", op="strong", nf="lxf", avg="std"),
    Case(1, (9,), 5, nodes="GL", eq="adv", op="strong", nf="lxf", avg="std"),
    Case(3, (3, 4, 2), 3, nodes="GL", eq="adv", op="strong", nf="std", avg="std"),
    Case(2, (4, 4), 4, nodes="GLL", eq="adv", op="split", nf="lxf", avg="std"),
    # config 2/3 family: Euler, EC split form + matrix dissipation
    Case(1, (10,), 4, nodes="GLL", eq="euler", op="split", nf="mat", avg="cha"),
    Case(2, (5, 4), 5, nodes="GLL", eq="euler", op="split", nf="mat", avg="cha"),
    Case(3, (3, 3, 4), 4, nodes="GLL", eq="euler", op="split", nf="mat", avg="cha"),
    Case(3, (2, 3, 2), 5, nodes="GLL", eq="euler", op="split", nf="mat", avg="cha"),
    # every other flux / operator combination the reference offers
    Case(2, (4, 3), 4, nodes="GLL", eq="euler", op="strong", nf="lxf", avg="std"),
    Case(2, (4, 3), 4, nodes="GL", eq="euler", op="strong", nf="lxf", avg="cha"),
    Case(3, (2, 2, 3), 3, nodes="GL", eq="euler", op="strong", nf="std", avg="std"),
    Case(2, (4, 3), 4, nodes="GLL", eq="euler", op="split", nf="std", avg="std"),
    Case(2, (4, 3), 4, nodes="GLL", eq="euler", op="split", nf="cha", avg="cha"),
    Case(3, (2, 2, 2), 4, nodes="GLL", eq="euler", op="split", nf="cha", avg="cha"),
    Case(2, (4, 3), 4, nodes="GLL", eq="euler", op="split", nf="sca", avg="cha"),
    Case(3, (2, 3, 2), 3, nodes="GLL", eq="euler", op="split", nf="sca", avg="std"),
    Case(1, (8,), 6, nodes="GLL", eq="euler", op="split", nf="sca", avg="cha"),
    Case(2, (3, 3), 6, nodes="GLL", eq="euler", op="split", tp="std", nf="mat", avg="cha"),
    Case(3, (2, 2, 2), 4, nodes="GLL", eq="euler", op="split", nf="lxf", avg="cha"),
    Case(2, (3, 4), 4, nodes="CGL", eq="euler", op="strong", nf="mat", avg="std"),
]


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=repr)
@pytest.mark.parametrize("state", ["random", "smooth"])
def test_rhs_matches_oracle(gpu, case, state):
    import flou_b200 as F
    orc = case.oracle()
    disc, eq = case.product()
    Q = (random_state(orc.ndof, case.nd, case.eq, amp=case.amp) if state == "random"
         else smooth_state(orc.coords, case.nd, case.eq))
    ref = orc.rhs(Q)
    dQ = disc.new_state()
    F.rhs(dQ, Q, F.EquationConfig(disc, eq), 0.0)
    assert np.all(np.isfinite(dQ))
    assert relerr(dQ, ref) <= RHS_TOL
    disc.close()


@pytest.mark.gpu
@pytest.mark.parametrize("kernel", ["auto", "node", "fused"])
@pytest.mark.parametrize("case", [CASES[1], CASES[6], CASES[8], CASES[11], CASES[16]], ids=repr)
def test_rhs_other_kernel_paths(gpu, case, kernel):
    """The parity cases above run the production path of large meshes (kernel="line"); the
    library's own choice on these small meshes (fused single-kernel stage) and the node-per-thread
    element kernel kept for A/B timing must agree with the oracle as well."""
    import flou_b200 as F
    orc = case.oracle()
    disc, eq = case.product(kernel=kernel)
    Q = random_state(orc.ndof, case.nd, case.eq, amp=case.amp)
    dQ = disc.new_state()
    F.rhs(dQ, Q, F.EquationConfig(disc, eq), 0.0)
    assert relerr(dQ, orc.rhs(Q)) <= RHS_TOL
    u = Q.copy(order="F")
    sol, _ = F.timeintegrate(u, disc, eq, F.ORK256(), 3e-4, dt=1e-4)
    import oracle as O
    assert relerr(sol.u[-1], orc.lsrk2n(Q, O.ORK256, 1e-4, 3)) <= 1e-10
    disc.close()


BC_CASES = [
    Case(2, (5, 4), 4, eq="euler", op="split", nf="mat", avg="cha", periodic=[("3", "4")],
         bcs={"1": ("inflow", [1.1, 0.33, 0.02, 2.6]), "2": ("outflow", None)}),
    Case(2, (4, 4), 4, eq="euler", op="split", nf="mat", avg="cha", periodic=[],
         bcs={"1": ("slip", None), "2": ("slip", None), "3": ("slip", None), "4": ("slip", None)}),
    Case(3, (3, 2, 2), 3, eq="euler", op="split", nf="mat", avg="cha", periodic=[("5", "6")],
         bcs={"1": ("inflow", [1.0, 0.4, 0.0, 0.1, 2.7]), "2": ("outflow", None),
              "3": ("slip", None), "4": ("slip", None)}),
    Case(1, (12,), 4, eq="euler", op="split", nf="mat", avg="cha", periodic=[],
         bcs={"1": ("table", lambda x: np.array([1.0, 0.0, 250.0]) if x[0] < 0.5
                    else np.array([0.125, 0.0, 25.0])),
              "2": ("table", lambda x: np.array([1.0, 0.0, 250.0]) if x[0] < 0.5
                    else np.array([0.125, 0.0, 25.0]))}),
    Case(2, (4, 3), 3, nodes="GL", eq="adv", op="strong", nf="lxf", avg="std", periodic=[("1", "2")],
         bcs={"3": ("table", lambda x: np.array([np.sin(3 * x[0])])), "4": ("outflow", None)}),
]


@pytest.mark.gpu
@pytest.mark.parametrize("case", BC_CASES, ids=repr)
def test_rhs_with_boundary_conditions(gpu, xtrace, case):
    import flou_b200 as F
    orc = case.oracle()
    disc, eq = case.product()
    Q = random_state(orc.ndof, case.nd, case.eq, amp=case.amp)
    ref = orc.rhs(Q)
    dQ = disc.new_state()
    F.rhs(dQ, Q, F.EquationConfig(disc, eq), 0.0)
    assert relerr(dQ, ref) <= RHS_TOL
    disc.close()


GENERAL_CASES = [
    # Cartesian mesh pushed through the per-node metric path must agree with both
    Case(2, (4, 3), 4, eq="euler", op="split", nf="mat", avg="cha", general=True),
    Case(3, (2, 3, 2), 3, eq="euler", op="split", nf="mat", avg="cha", general=True),
    # curved (vertex-perturbed) meshes, wall/inflow/outflow boundaries
    Case(2, (5, 4), 4, eq="euler", op="split", nf="mat", avg="cha", periodic=[], perturb_amp=0.15,
         bcs={"1": ("inflow", [1.1, 0.33, 0.02, 2.6]), "2": ("outflow", None),
              "3": ("slip", None), "4": ("slip", None)}),
    Case(2, (4, 4), 6, eq="euler", op="split", nf="mat", avg="cha", periodic=[], perturb_amp=0.1,
         bcs={"1": ("slip", None), "2": ("slip", None), "3": ("slip", None), "4": ("slip", None)}),
    Case(3, (3, 2, 2), 4, eq="euler", op="split", nf="mat", avg="cha", periodic=[], perturb_amp=0.1,
         bcs={"1": ("inflow", [1.0, 0.4, 0.0, 0.1, 2.7]), "2": ("outflow", None),
              "3": ("slip", None), "4": ("slip", None), "5": ("slip", None), "6": ("slip", None)}),
    Case(2, (4, 3), 4, nodes="GL", eq="euler", op="strong", nf="lxf", avg="std", periodic=[],
         perturb_amp=0.1,
         bcs={"1": ("slip", None), "2": ("outflow", None), "3": ("slip", None), "4": ("slip", None)}),
    Case(2, (4, 3), 3, nodes="GL", eq="adv", op="strong", nf="lxf", avg="std", periodic=[],
         perturb_amp=0.1, bcs={str(i): ("outflow", None) for i in range(1, 5)}),
]


@pytest.mark.gpu
@pytest.mark.parametrize("case", GENERAL_CASES, ids=repr)
def test_rhs_general_geometry(gpu, xtrace, case):
    import flou_b200 as F
    orc = case.oracle()
    disc, eq = case.product()
    Q = random_state(orc.ndof, case.nd, case.eq, amp=case.amp)
    ref = orc.rhs(Q)
    dQ = disc.new_state()
    F.rhs(dQ, Q, F.EquationConfig(disc, eq), 0.0)
    assert relerr(dQ, ref) <= RHS_TOL
    disc.close()


@pytest.mark.gpu
@pytest.mark.parametrize("case,dt,nsteps", [
    (Case(2, (6, 5), 4, nodes="GL", eq="adv", op="strong", nf="lxf", avg="std"), 1e-3, 40),
    (Case(2, (5, 4), 5, eq="euler", op="split", nf="mat", avg="cha"), 1e-3, 40),
    (Case(3, (3, 3, 3), 4, eq="euler", op="split", nf="mat", avg="cha"), 1e-3, 25),
], ids=lambda v: repr(v) if isinstance(v, Case) else None)
@pytest.mark.parametrize("solver", ["ORK256", "CarpenterKennedy2N54"])
def test_state_after_n_steps(gpu, case, dt, nsteps, solver):
    """State after N RK steps: stated tolerance 1e-10 relative."""
    import flou_b200 as F
    import oracle as O
    orc = case.oracle()
    disc, eq = case.product()
    Q = smooth_state(orc.coords, case.nd, case.eq)
    tab = O.ORK256 if solver == "ORK256" else O.CARPENTER_KENNEDY_2N54
    ref = orc.lsrk2n(Q, tab, dt, nsteps)
    u = Q.copy(order="F")
    sol, _ = F.timeintegrate(u, disc, eq, getattr(F, solver)(williamson_condition=False),
                             nsteps * dt, dt=dt, adaptive=False, alias_u0=True)
    assert sol is not None
    assert relerr(sol.u[-1], ref) <= 1e-10
    assert relerr(sol.u[0], Q) == 0.0
    # graph replay and direct launches agree bitwise
    disc2, eq2 = case.product(use_graph=False)
    u2 = Q.copy(order="F")
    F.timeintegrate(u2, disc2, eq2, getattr(F, solver)(), nsteps * dt, dt=dt)
    assert np.array_equal(u2, u)
    disc.close(); disc2.close()


@pytest.mark.gpu
def test_sod_tube_kat_on_gpu(gpu):
    """The reference's own KAT (test/runtests.jl:35-39, setup test/tests.jl:90-134) through
    the CUDA path: minimum/maximum of u(tf) to rtol 1e-7."""
    import flou_b200 as F
    eq = F.EulerEquation(1, 1.4)
    basis = F.LagrangeBasis("GLL", 4)
    std = F.StdSegment(basis, F.DGSEMrec(basis), eq.nv)
    mesh = F.CartesianMesh(1, 0, 1, 20)

    def Qext(_, x, __, ___, eq_):
        P = (1.0, 0.0, 100.0) if x[0] < 0.5 else (0.125, 0.0, 10.0)
        return F.vars_prim2cons(P, eq_)
    bcs = {"1": F.GenericBC(Qext), "2": F.GenericBC(Qext)}
    op = F.SplitDivOperator(F.MatrixDissipation(F.ChandrasekharAverage(), 1.0))
    dg = F.MultielementDisc(mesh, std, eq, op, bcs)
    Q = dg.new_state()
    for i, x in enumerate(dg.coords()):
        Q[i] = Qext((), x, (), 0.0, eq)
    sol, _ = F.timeintegrate(Q, dg, eq, F.ORK256(williamson_condition=False), 0.018,
                             dt=1e-4, adaptive=False, alias_u0=True)
    assert abs(sol.u[-1].min() / -0.1662939230897596 - 1) <= 1e-7
    assert abs(sol.u[-1].max() / 254.90504152149907 - 1) <= 1e-7
    dg.close()


# ------------------------------------------------------------------ row f4: split form on Gauss nodes
GAUSS_SPLIT_CASES = [
    Case(1, (9,), 4, nodes="GL", op="split", nf="mat", avg="cha"),
    Case(2, (4, 5), 4, nodes="GL", op="split", nf="cha", avg="cha"),
    Case(2, (3, 4), 5, nodes="GL", op="split", tp="std", nf="lxf", avg="std"),
    Case(3, (2, 3, 2), 3, nodes="GL", op="split", nf="mat", avg="cha"),
    Case(3, (2, 2, 2), 4, nodes="GL", op="split", tp="std", nf="sca", avg="cha"),
    Case(2, (4, 3), 4, nodes="GL", op="split", nf="mat", avg="cha", periodic=[("3", "4")],
         bcs={"1": ("inflow", [1.1, 0.33, 0.02, 2.6]), "2": ("outflow", None)}),
    # curved sub-grids (PhysicalRegions.jl:179-292): vertex-perturbed meshes
    Case(2, (4, 3), 4, nodes="GL", op="split", nf="mat", avg="cha", perturb_amp=0.08),
    Case(2, (3, 3), 5, nodes="GL", op="split", tp="std", nf="lxf", avg="std", perturb_amp=0.1),
    Case(3, (2, 2, 2), 3, nodes="GL", op="split", nf="mat", avg="cha", perturb_amp=0.06),
]


@pytest.mark.gpu
@pytest.mark.parametrize("case", GAUSS_SPLIT_CASES, ids=repr)
@pytest.mark.parametrize("state", ["random", "smooth"])
def test_gauss_node_split_form_matches_oracle(gpu, case, state):
    """SplitDivOperator on Gauss nodes: the entropy-projected surface term
    (_splitdiv_nb_surface_contribution!, OpDivergence.jl:300-437) in the line kernel vs the oracle
    (pinned by its exact entropy balance, tests/test_oracle_properties.py)."""
    import flou_b200 as F
    import oracle as O
    orc = case.oracle()
    disc, eq = case.product()
    Q = (random_state(orc.ndof, case.nd, case.eq, amp=case.amp) if state == "random"
         else smooth_state(orc.coords, case.nd, case.eq))
    dQ = disc.new_state()
    F.rhs(dQ, Q, F.EquationConfig(disc, eq), 0.0)
    assert np.all(np.isfinite(dQ))
    assert relerr(dQ, orc.rhs(Q)) <= RHS_TOL
    u = Q.copy(order="F")
    sol, _ = F.timeintegrate(u, disc, eq, F.ORK256(), 4e-4, dt=1e-4)
    assert relerr(sol.u[-1], orc.lsrk2n(Q, O.ORK256, 1e-4, 4)) <= 1e-10
    disc.close()


@pytest.mark.gpu
def test_gauss_node_split_form_unsupported_combinations_raise(gpu):
    with pytest.raises(ValueError):
        Case(2, (3, 3), 4, nodes="GL", op="split").product(kernel="fused")
    with pytest.raises(ValueError):
        Case(2, (3, 3), 4, nodes="GL", op="split").product(kernel="node")


# ------------------------------------------------------------------ row f2: HybridDivOperator
HYBRID_CASES = [
    Case(1, (10,), 4, op="hybrid", nf="mat", avg="cha", blend=1.0),
    Case(2, (5, 4), 6, op="hybrid", nf="mat", avg="cha", blend=1.0),          # Shockwave2D's operator
    Case(2, (4, 5), 4, op="hybrid", tp="std", nf="lxf", avg="std", blend=1e-3),
    Case(2, (4, 3), 5, op="hybrid", tp="cha", nf="sca", avg="cha", blend=0.05),
    Case(3, (3, 2, 3), 4, op="hybrid", nf="mat", avg="cha", blend=1.0),
    Case(3, (2, 3, 2), 3, op="hybrid", tp="std", nf="cha", avg="cha", blend=0.2),
    Case(2, (5, 4), 4, op="hybrid", nf="mat", avg="cha", blend=1.0, periodic=[("3", "4")],
         bcs={"1": ("inflow", [1.1, 0.33, 0.02, 2.6]), "2": ("outflow", None)}),
    # curved sub-grids (PhysicalRegions.jl:179-292): vertex-perturbed meshes
    Case(2, (4, 3), 4, op="hybrid", nf="mat", avg="cha", blend=0.3, perturb_amp=0.08),
    Case(2, (3, 4), 5, op="hybrid", tp="std", nf="lxf", avg="std", blend=1.0, perturb_amp=0.1),
    Case(3, (2, 2, 2), 3, op="hybrid", nf="mat", avg="cha", blend=0.5, perturb_amp=0.06),
    # Gauss nodes: everything as a surface contribution (_hybrid_nb_surface_contribution!,
    # OpDivergence.jl:629-779)
    Case(1, (9,), 4, nodes="GL", op="hybrid", nf="mat", avg="cha", blend=1.0),
    Case(2, (4, 5), 4, nodes="GL", op="hybrid", nf="mat", avg="cha", blend=0.3),
    Case(2, (3, 4), 5, nodes="GL", op="hybrid", tp="std", nf="lxf", avg="std", blend=1.0),
    Case(3, (2, 2, 2), 3, nodes="GL", op="hybrid", nf="mat", avg="cha", blend=0.5),
    Case(2, (4, 3), 4, nodes="GL", op="hybrid", nf="mat", avg="cha", blend=0.5, perturb_amp=0.08),
    Case(3, (2, 2, 2), 3, nodes="GL", op="hybrid", tp="std", nf="sca", avg="cha", blend=1.0, perturb_amp=0.06),
]


@pytest.mark.gpu
@pytest.mark.parametrize("case", HYBRID_CASES, ids=repr)
@pytest.mark.parametrize("state", ["random", "smooth"])
def test_hybrid_rhs_matches_oracle(gpu, xtrace, case, state):
    """HybridDivOperator (OpDivergence.jl:452-612) on the device vs the oracle restatement that
    reproduces the reference's Shockwave2D value; RHS and the state after a few RK steps."""
    import flou_b200 as F
    import oracle as O
    orc = case.oracle()
    disc, eq = case.product()
    Q = (random_state(orc.ndof, case.nd, case.eq, amp=case.amp) if state == "random"
         else smooth_state(orc.coords, case.nd, case.eq))
    dQ = disc.new_state()
    F.rhs(dQ, Q, F.EquationConfig(disc, eq), 0.0)
    assert np.all(np.isfinite(dQ))
    assert relerr(dQ, orc.rhs(Q)) <= RHS_TOL
    u = Q.copy(order="F")
    sol, _ = F.timeintegrate(u, disc, eq, F.ORK256(), 4e-4, dt=1e-4)
    assert relerr(sol.u[-1], orc.lsrk2n(Q, O.ORK256, 1e-4, 4)) <= 1e-10
    disc.close()


@pytest.mark.gpu
def test_shockwave_2d_kat_on_gpu(gpu):
    """The reference's only 2-D known-answer test (test/runtests.jl:40-44, setup
    test/tests.jl:136-185) through the CUDA path: 11x3 elements, GLL(6),
    HybridDivOperator(MatrixDissipation(ChandrasekharAverage(), 1.0), 1.0), y-periodic, GenericBC
    in x, ORK256, dt = 1e-2, tf = 1: maximum(u(tf)) to the reference's rtol 1e-7."""
    import flou_b200 as F
    eq = F.EulerEquation(2, 1.4)
    basis = F.LagrangeBasis("GLL", 6)
    std = F.StdQuad(basis, F.DGSEMrec(basis), eq.nv)
    mesh = F.CartesianMesh(2, (-1, 0), (1, 1), (11, 3))
    mesh.apply_periodicBCs(("3", "4"))
    rho0, M0, p0 = 1.0, 2.0, 1.0
    u0 = M0 * F.soundvelocity(rho0, p0, eq)
    rho1, u1, p1 = F.normal_shockwave(rho0, u0, p0, eq)
    Q0 = F.vars_prim2cons((rho0, u0, 0.0, p0), eq)
    Q1 = F.vars_prim2cons((rho1, u1, 0.0, p1), eq)

    def Qext(_, xy, __, ___, ____):
        return Q0 if xy[0] < 0 else Q1
    bcs = {"1": F.GenericBC(Qext), "2": F.GenericBC(Qext)}
    op = F.HybridDivOperator(F.MatrixDissipation(F.ChandrasekharAverage(), 1.0), 1.0)
    dg = F.MultielementDisc(mesh, std, eq, op, bcs)
    Q = dg.new_state()
    for i, xy in enumerate(dg.coords()):
        Q[i] = Qext((), xy, (), 0.0, eq)
    sol, _ = F.timeintegrate(Q, dg, eq, F.ORK256(williamson_condition=False), 1.0,
                             dt=1e-2, adaptive=False, alias_u0=True)
    assert abs(sol.u[-1].max() / 12.977466260673845 - 1) <= 1e-7
    assert abs(sol.u[-1].min()) < 1e-9
    dg.close()


@pytest.mark.gpu
def test_hybrid_unsupported_combinations_raise(gpu):
    """No silent fallback: the fused / node kernels are refused for the hybrid operator."""
    import flou_b200 as F
    with pytest.raises(ValueError):
        Case(2, (3, 3), 4, op="hybrid").product(kernel="node")
    with pytest.raises(ValueError):
        Case(2, (3, 3), 4, op="hybrid").product(kernel="fused")


@pytest.mark.gpu
@pytest.mark.parametrize("case", [
    Case(1, (9,), 4), Case(2, (5, 4), 5), Case(3, (3, 2, 3), 4), Case(3, (2, 2, 2), 5, perturb_amp=0.08),
    Case(2, (4, 3), 6, perturb_amp=0.1),
    Case(2, (6, 5), 4, nodes="GL", eq="adv", op="strong", nf="lxf", avg="std"),
], ids=repr)
def test_max_dt_matches_oracle(gpu, case):
    """Row f1: get_max_dt as a device reduction (host state and device-resident state)."""
    import flou_b200 as F
    orc = case.oracle()
    disc, eq = case.product()
    Q = random_state(orc.ndof, case.nd, case.eq, amp=case.amp)
    want = orc.max_dt(Q, 0.4)
    got = F.get_max_dt(Q, disc, eq, 0.4)
    assert abs(got / want - 1) <= 1e-14
    disc.upload(Q)
    assert F.get_max_dt(None, disc, eq, 0.4) == got
    disc.close()


@pytest.mark.gpu
def test_domain_error_is_reported(gpu):
    import flou_b200 as F
    case = Case(2, (3, 3), 4, eq="euler", op="split", nf="mat", avg="cha")
    disc, eq = case.product()
    Q = random_state(disc.ndofs, 2, "euler")
    Q[5, 0] = -1.0                      # negative density
    sol, _ = F.timeintegrate(Q, disc, eq, F.ORK256(), 1e-3, dt=1e-3)
    assert sol is None                  # FlouTime.jl:39-51 returns nothing after "crashed"
    disc.close()


@pytest.mark.gpu
def test_golden_cart3d_fixture(gpu):
    """Committed golden vectors (tests/golden/make_golden.py): RHS and state after 5 steps."""
    import os
    import flou_b200 as F
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "cart3d_p3.npz"))
    disc, eq = Case(3, (3, 3, 3), 4).product()
    Q = np.asfortranarray(g["Q"])
    dQ = disc.new_state()
    F.rhs(dQ, Q, F.EquationConfig(disc, eq), 0.0)
    assert relerr(dQ, g["dQ"]) <= RHS_TOL
    u = Q.copy(order="F")
    sol, _ = F.timeintegrate(u, disc, eq, F.ORK256(), 5e-3, dt=1e-3)
    assert relerr(sol.u[-1], g["u5"]) <= 1e-10
    disc.close()
