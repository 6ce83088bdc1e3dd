"""Rows a13 / a9 beyond the defaults: the source-term hook (apply_sourceterm!,
MultielementDiscontinuous.jl:139-146) and GenericBC closures that read Qin / frame / time
(Interfaces.jl:44-48, FlouSpatial.jl:85-91).  Both are host closures in the reference; here the
host tabulates them (per node / per boundary-face node, re-tabulated before every stage when they
depend on the state or the time) and the device adds the table / reads the table.

Reference values: the oracle's RHS with the same tables -- the source is additive after the mass
matrix, the exterior states are written into the oracle's own bc_table from the oracle's own
interior traces (Qf, master side) -- and a plain numpy 2N loop around it."""
import numpy as np
import pytest

from common import Case, random_state, relerr, smooth_state

RHS_TOL = 1e-12


def _lsrk(rhs, Q, tab, dt, nsteps, t0=0.0):
    """tmp = A_s tmp + dt k(u, t + c_s dt); u += B_s tmp (LowStorageRK2N, FlouTime.jl:34-38)."""
    A, B, c = tab
    u, tmp = Q.copy(order="F"), np.zeros_like(Q)
    for n in range(nsteps):
        t = t0 + n * dt
        for s in range(len(B)):
            tmp = A[s] * tmp + dt * rhs(u, t + c[s] * dt)
            u = u + B[s] * tmp
    return u


def _source(Q, x, t):
    """Depends on the state, the position and the time (a damping + a travelling body force)."""
    S = np.zeros_like(Q)
    S[0] = 0.05 * np.sin(3 * x[0] + 2 * t)
    S[1] = -0.3 * Q[1] + 0.2 * Q[0] * np.cos(t + x[-1])
    S[-1] = 0.1 * Q[0] * x[0] - 0.05 * t * Q[-1]
    return S


def _source_vec(Q, x, t):
    S = np.zeros_like(Q)
    S[:, 0] = 0.05 * np.sin(3 * x[:, 0] + 2 * t)
    S[:, 1] = -0.3 * Q[:, 1] + 0.2 * Q[:, 0] * np.cos(t + x[:, -1])
    S[:, -1] = 0.1 * Q[:, 0] * x[:, 0] - 0.05 * t * Q[:, -1]
    return S


SRC_CASES = [Case(2, (5, 4), 4), Case(3, (3, 2, 3), 4), Case(3, (12, 12, 6), 5),
             Case(2, (4, 3), 4, perturb_amp=0.1)]


@pytest.mark.gpu
@pytest.mark.parametrize("case", SRC_CASES, ids=repr)
def test_state_and_time_dependent_source(gpu, case):
    import flou_b200 as F
    orc = case.oracle()
    disc, eq = case.product(create=False)
    # same discretisation with the source (vectorised closure; the scalar form is checked below)
    disc = F.MultielementDisc(disc.mesh, disc.std, eq, disc.operators[0], {},
                              source=F.Source(_source_vec, vectorized=True), kernel="line",
                              geometry="general" if case.general else None)
    assert disc.has_dynamic
    Q = np.asfortranarray(0.9 * smooth_state(orc.coords, case.nd, "euler") + 0.1 * random_state(orc.ndof, case.nd, "euler", amp=0.3))
    t = 0.37
    ref = orc.rhs(Q) + _source_vec(Q, orc.coords, t)
    dQ = disc.new_state()
    F.rhs(dQ, Q, F.EquationConfig(disc, eq), t)
    assert relerr(dQ, ref) <= RHS_TOL
    assert relerr(dQ, orc.rhs(Q)) > 1e-4                      # the source is really there
    dt, n = 1e-4, 3
    want = _lsrk(lambda u, ts: orc.rhs(u) + _source_vec(u, orc.coords, ts), Q, _tab(), dt, n, t0=0.2)
    u = Q.copy(order="F")
    sol, _ = F.timeintegrate(u, disc, eq, F.ORK256(), 0.2 + n * dt, dt=dt, t0=0.2)
    assert sol is not None and relerr(sol.u[-1], want) <= 1e-10
    disc.close()


def _tab():
    import flou_b200 as F
    return (np.array(F.ORK256.A), np.array(F.ORK256.B), np.array(F.ORK256.c))


@pytest.mark.gpu
def test_position_only_source_stays_on_the_fast_path(gpu):
    """Source(f, state=False, time=False): tabulated once, the RK loop keeps its CUDA graph."""
    import flou_b200 as F
    case = Case(3, (4, 3, 3), 4)
    orc = case.oracle()
    base, eq = case.product(create=False)
    f = lambda Q, x, t: [0.0, 0.4 * np.sin(2 * x[0]), 0.0, -0.2, 0.1 * x[1]]      # scalar closure
    disc = F.MultielementDisc(base.mesh, base.std, eq, base.operators[0], {},
                              source=F.Source(f, state=False, time=False), kernel="line")
    assert not disc.has_dynamic
    S = np.array([f(None, x, 0.0) for x in orc.coords])
    Q = smooth_state(orc.coords, 3, "euler")
    dQ = disc.new_state()
    F.rhs(dQ, Q, F.EquationConfig(disc, eq), 0.0)
    assert relerr(dQ, orc.rhs(Q) + S) <= RHS_TOL
    dt, n = 1e-4, 6                                            # >= 4 steps: graph replay
    want = _lsrk(lambda u, ts: orc.rhs(u) + S, Q, _tab(), dt, n)
    u = Q.copy(order="F")
    sol, _ = F.timeintegrate(u, disc, eq, F.ORK256(), n * dt, dt=dt)
    assert sol is not None and relerr(sol.u[-1], want) <= 1e-10
    # a plain callable works too (treated as depending on everything)
    disc2 = F.MultielementDisc(base.mesh, base.std, eq, base.operators[0], {}, source=lambda Q, x, t: _source(Q, x, t),
                               kernel="line")
    dQ2 = disc2.new_state()
    F.rhs(dQ2, Q, F.EquationConfig(disc2, eq), 0.1)
    assert relerr(dQ2, orc.rhs(Q) + _source_vec(Q, orc.coords, 0.1)) <= RHS_TOL
    disc.close(); disc2.close()


# ------------------------------------------------------------------------------------ GenericBC
def _bc_dynamic(Qin, x, frame, t, eq):
    """Exterior state from the interior one, the face normal and the time: a partially reflecting
    wall whose reflection coefficient oscillates in time, plus a position-dependent density."""
    Qin = np.asarray(Qin, dtype=float)
    nd = len(x)
    n = np.asarray(frame.n)
    m = Qin[1:1 + nd]
    alpha = 1.0 + 0.5 * np.sin(3.0 * t)
    Qe = Qin.copy()
    Qe[1:1 + nd] = m - alpha * np.dot(m, n) * n
    Qe[0] = Qin[0] * (1.0 + 0.05 * np.cos(x[0] + t))
    return Qe


def _oracle_rhs_with_dynamic_bc(orc, Q, t, names, eq):
    """Write the closure's exterior states into the oracle's bc_table from the oracle's OWN interior
    traces (Qf side 1, filled by a first pass) and evaluate the RHS again."""
    orc.rhs(Q)
    nfp, nv = orc.nfp, orc.nv
    Qf = orc._keep["Qf0"].reshape(nv, orc.nf * nfp).T           # (face dof, var), column-major in C
    table = orc._keep["bc_table"]
    offs, faces = orc._keep["bc_offsets"], orc._keep["bc_faces"]
    from flou_b200 import Frame
    for ib, name in enumerate(orc.mesh.bdnames):
        if name not in names:
            continue
        for m in range(int(offs[ib]), int(offs[ib + 1])):
            f = int(faces[m]) - 1
            for i in range(nfp):
                r = f * nfp + i
                fr = orc.frames[r]
                table[m * nfp + i] = _bc_dynamic(Qf[r], orc.fcoords[r], Frame(fr[0], fr[1], fr[2]), t, eq)
    return orc.rhs(Q)


BC_CASES = [
    Case(2, (5, 4), 4, periodic=[("3", "4")],
         bcs={"1": ("table", lambda x: np.zeros(4)), "2": ("outflow", None)}),
    Case(3, (3, 3, 2), 3, periodic=[("5", "6")],
         bcs={"1": ("table", lambda x: np.zeros(5)), "2": ("table", lambda x: np.zeros(5)),
              "3": ("slip", None), "4": ("slip", None)}),
    Case(2, (4, 4), 5, periodic=[], perturb_amp=0.12,
         bcs={"1": ("table", lambda x: np.zeros(4)), "2": ("slip", None), "3": ("table", lambda x: np.zeros(4)),
              "4": ("outflow", None)}),
]


@pytest.mark.gpu
@pytest.mark.parametrize("case", BC_CASES, ids=repr)
def test_generic_bc_reading_state_frame_and_time(gpu, case):
    import flou_b200 as F
    orc = case.oracle()
    base, eq = case.product(create=False)
    dyn_names = [n for n, (kind, _) in case.bcs.items() if kind == "table"]
    bcs = {}
    for name, (kind, param) in case.bcs.items():
        bcs[name] = (F.GenericBC(_bc_dynamic) if kind == "table" else
                     F.EulerOutflowBC() if kind == "outflow" else F.EulerSlipBC())
    disc = F.MultielementDisc(base.mesh, base.std, eq, base.operators[0], bcs, kernel="line",
                              geometry="general" if case.general else None)
    assert disc.has_dynamic and len(disc._dynamic_bcs) == len(dyn_names)
    Q = np.asfortranarray(0.9 * smooth_state(orc.coords, case.nd, "euler") + 0.1 * random_state(orc.ndof, case.nd, "euler", amp=0.3))
    t = 0.21
    ref = _oracle_rhs_with_dynamic_bc(orc, Q, t, dyn_names, eq)
    dQ = disc.new_state()
    F.rhs(dQ, Q, F.EquationConfig(disc, eq), t)
    assert relerr(dQ, ref) <= RHS_TOL
    dt, n = 2e-4, 3
    want = _lsrk(lambda u, ts: _oracle_rhs_with_dynamic_bc(orc, u, ts, dyn_names, eq), Q, _tab(), dt, n, t0=0.1)
    u = Q.copy(order="F")
    sol, _ = F.timeintegrate(u, disc, eq, F.ORK256(), 0.1 + n * dt, dt=dt, t0=0.1)
    assert sol is not None and relerr(sol.u[-1], want) <= 1e-10
    disc.close()
