"""Row f3 on the GPU: monitors, Zhang-Shu limiter, stage limiter inside the RK loop and the
device-side callbacks, against the oracle restatement (numpy, element by element in the
reference's order).  No reference test pins these (parity unpinned): the oracle is
cross-checked by properties in tests/test_oracle_properties.py."""
import numpy as np
import pytest

from common import Case, random_state, relerr, troubled_state

CASES = [
    Case(1, (10,), 4),
    Case(2, (5, 4), 5),
    Case(3, (3, 2, 3), 4),
    Case(3, (2, 2, 2), 5, perturb_amp=0.08),
    Case(2, (4, 3), 6, perturb_amp=0.1),
]


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=repr)
def test_monitors_match_oracle(gpu, case):
    """get_monitor(:kinetic_energy | :entropy) (Equations/Euler.jl:541-593), host state and
    device-resident state; tolerance 1e-12 relative (summation order differs)."""
    import flou_b200 as F
    orc = case.oracle()
    disc, eq = case.product()
    assert F.list_monitors(disc, eq) == ("kinetic_energy", "entropy")
    Q = random_state(orc.ndof, case.nd, "euler")
    # an explicit host state is staged in a scratch buffer: the device-resident state of an ongoing
    # run (here: another state, uploaded first) is not replaced by the query
    resident = random_state(orc.ndof, case.nd, "euler", seed=99, amp=0.2)
    disc.upload(resident)
    for name in ("kinetic_energy", "entropy", "energy"):
        mon = F.get_monitor(disc, eq, name)
        want = orc.monitor(Q, name)
        got = mon(Q, disc, eq)
        assert abs(got / want - 1) <= 1e-12
        assert abs(mon(None, disc, eq) / orc.monitor(resident, name) - 1) <= 1e-12
        assert F.get_max_dt(Q, disc, eq, 0.5) > 0.0
    assert np.array_equal(disc.download(), resident)
    disc.upload(Q)
    for name in ("kinetic_energy", "entropy"):
        mon = F.get_monitor(disc, eq, name)
        assert mon(None, disc, eq) == mon(Q, disc, eq)      # same kernels on the resident copy: bitwise
    with pytest.raises(ValueError):
        F.get_monitor(disc, eq, "enstrophy")
    disc.close()


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=repr)
def test_zhang_shu_matches_oracle(gpu, case):
    """get_limiter(:zhang_shu, minval) (Equations/Euler.jl:597-660): limited state to 1e-12
    relative, identity (bitwise) on a state above the floor."""
    import flou_b200 as F
    orc = case.oracle()
    disc, eq = case.product()
    assert F.list_limiters(disc, eq) == ("zhang_shu",)
    lim = F.get_limiter(disc, eq, "zhang_shu", 1e-2)
    good = random_state(orc.ndof, case.nd, "euler", amp=0.3)
    out = good.copy(order="F")
    lim(out, disc, eq)
    assert np.array_equal(out, good)
    Q = troubled_state(orc.ndof, case.nd)
    want = orc.zhang_shu(Q, 1e-2)
    got = Q.copy(order="F")
    lim(got, disc, eq)
    assert not np.array_equal(got, Q)
    assert relerr(got, want) <= 1e-12
    # device-resident variant
    disc.upload(Q)
    lim(None, disc, eq)
    assert np.array_equal(disc.download(), got)
    with pytest.raises(ValueError):
        F.get_limiter(disc, eq, "zhang_shu")         # "The minimum value must be specified ..."
    disc.close()


@pytest.mark.gpu
@pytest.mark.parametrize("case", [CASES[1], CASES[2], CASES[4]], ids=repr)
@pytest.mark.parametrize("nsteps", [3, 6])
def test_stage_limiter_in_the_rk_loop(gpu, case, nsteps):
    """ORK256(stage_limiter! = get_limiter_callback(dg, eq, :zhang_shu, minval)) as in
    examples/src/3D_Euler.jl:76-80: limiter after every stage, on the device (direct launches for
    3 steps, CUDA-graph replay for 6)."""
    import flou_b200 as F
    import oracle as O
    orc = case.oracle()
    disc, eq = case.product()
    # mild state + a floor that bites: density/pressure around 1 +- 0.3, floor 0.9
    Q = random_state(orc.ndof, case.nd, "euler", amp=0.3)
    minval, dt = 0.9, 1e-4
    want = orc.lsrk2n_limited(Q, O.ORK256, dt, nsteps, minval)
    assert relerr(want, orc.lsrk2n(Q, O.ORK256, dt, nsteps)) > 1e-6      # the limiter is active
    solver = F.ORK256(williamson_condition=False,
                      stage_limiter=F.get_limiter_callback(disc, eq, "zhang_shu", minval))
    u = Q.copy(order="F")
    sol, _ = F.timeintegrate(u, disc, eq, solver, nsteps * dt, dt=dt)
    assert relerr(sol.u[-1], want) <= 1e-10
    # and the limiter is off again for a plain solver on the same handle
    u = Q.copy(order="F")
    sol, _ = F.timeintegrate(u, disc, eq, F.ORK256(), nsteps * dt, dt=dt)
    assert relerr(sol.u[-1], orc.lsrk2n(Q, O.ORK256, dt, nsteps)) <= 1e-10
    disc.close()


@pytest.mark.gpu
def test_monitor_and_cfl_callbacks(gpu):
    """get_monitor_callback + get_cfl_callback (FlouTime.jl:92-134) evaluated on the device between
    steps: the CFL callback sets dt before the first step and after every step, the monitor
    records (t, iter, value) after every selected step, the last step lands on tfinal."""
    import flou_b200 as F
    import oracle as O
    case = Case(2, (5, 4), 4)
    orc = case.oracle()
    disc, eq = case.product()
    Q = random_state(orc.ndof, case.nd, "euler", amp=0.2)
    cfl, tf = 0.01, 2.5e-3
    mcb, mout = F.get_monitor_callback(float, float, disc, eq, "entropy")
    kcb, kout = F.get_monitor_callback(float, float, disc, eq, "kinetic_energy", iter=range(2, 100, 2))
    cb = F.make_callback_list(F.get_cfl_callback(cfl, 1e-3), mcb, kcb)
    u = Q.copy(order="F")
    sol, _ = F.timeintegrate(u, disc, eq, F.ORK256(), tf, dt=1.0, callback=cb)
    # the same loop on the oracle
    uo, t, times, ent, kin = Q.copy(order="F"), 0.0, [], [], []
    dt = min(orc.max_dt(uo, cfl), 1e-3)
    it = 0
    while t < tf - 1e-14:
        h = min(dt, tf - t)
        uo = orc.lsrk2n(uo, O.ORK256, h, 1)
        t += h
        it += 1
        dt = min(orc.max_dt(uo, cfl), 1e-3)
        times.append(t)
        ent.append(orc.monitor(uo, "entropy"))
        if it % 2 == 0:
            kin.append(orc.monitor(uo, "kinetic_energy"))
    assert sol.iterations == it and abs(sol.t[-1] - tf) <= 1e-15
    assert mout.iter == list(range(1, it + 1)) and np.allclose(mout.time, times, rtol=1e-12, atol=0)
    assert np.allclose(mout.value, ent, rtol=1e-11, atol=0)
    assert kout.iter == list(range(2, it + 1, 2)) and np.allclose(kout.value, kin, rtol=1e-11, atol=0)
    assert relerr(sol.u[-1], uo) <= 1e-10
    disc.close()
