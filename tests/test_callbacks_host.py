"""Host logic of the step-by-step loop with device-side callbacks (flou_b200/time.py,
flou_b200/monitors.py; FlouTime.jl:56-152) on CPU: the C library is replaced by a recording stub,
so only the control flow is under test -- which callbacks fire at which step, what dt each step
gets, that the last step lands on tfinal and that the stage limiter is switched on and off."""
import ctypes as C

import numpy as np
import pytest

import flou_b200 as F
from flou_b200 import _lib as L


class _StubLib:
    """Implements the entry points the loop calls; records them."""

    def __init__(self, max_dts):
        self.calls, self.max_dts, self.t = [], list(max_dts), 0.0

    def flou_b200_set_stage_limiter(self, h, enable, minval):
        self.calls.append(("limiter", int(enable), float(minval)))
        return 0

    def flou_b200_upload_state(self, h, q):
        self.calls.append(("upload",))
        return 0

    def flou_b200_download_state(self, h, q):
        self.calls.append(("download",))
        return 0

    def flou_b200_lsrk2n_advance(self, h, nstages, A, B, c, dt, t0, nsteps):
        self.calls.append(("advance", float(dt), float(t0), int(nsteps), int(nstages)))
        self.t = t0 + dt * nsteps
        return 0

    def flou_b200_status(self, h, flags):
        flags._obj.value = 0
        return 0

    def flou_b200_max_dt(self, h, q, cfl, out):
        out._obj.value = self.max_dts.pop(0) * cfl
        self.calls.append(("max_dt",))
        return 0

    def flou_b200_monitor(self, h, kind, q, out):
        out._obj.value = 100.0 * kind + self.t          # a value that identifies (kind, time)
        self.calls.append(("monitor", int(kind)))
        return 0

    def flou_b200_destroy(self, h):
        return 0


@pytest.fixture
def stub(monkeypatch):
    def make(max_dts=()):
        lib = _StubLib(max_dts)
        monkeypatch.setattr(L, "lib", lambda: lib)
        return lib
    return make


def _disc():
    mesh = F.CartesianMesh(2, (0, 0), (1, 1), (2, 2)).apply_periodicBCs(("1", "2"), ("3", "4"))
    eq = F.EulerEquation(2, 1.4)
    b = F.LagrangeBasis("GLL", 3)
    std = F.StdQuad(b, F.DGSEMrec(b), 4)
    op = F.SplitDivOperator(F.MatrixDissipation(F.ChandrasekharAverage(), 1.0))
    d = F.MultielementDisc(mesh, std, eq, op, {}, create=False)
    d._h = C.c_void_p(1)            # pretend a device handle exists; every call goes to the stub
    return d, eq


def test_monitor_callbacks_fire_on_their_steps_and_the_last_step_lands_on_tfinal(stub):
    lib = stub()
    disc, eq = _disc()
    mcb, mout = F.get_monitor_callback(float, float, disc, eq, "entropy")
    kcb, kout = F.get_monitor_callback(float, float, disc, eq, "kinetic_energy", iter=(2, 3))
    Q = disc.new_state()
    sol, _ = F.timeintegrate(Q, disc, eq, F.ORK256(), 0.35, dt=0.1, callback=F.make_callback_list(mcb, kcb))
    steps = [c for c in lib.calls if c[0] == "advance"]
    assert [round(c[1], 12) for c in steps] == [0.1, 0.1, 0.1, 0.05]          # last step shortened
    assert all(c[3] == 1 and c[4] == 5 for c in steps)
    assert sol.iterations == 4 and abs(sol.t[-1] - 0.35) < 1e-15
    assert mout.iter == [1, 2, 3, 4] and np.allclose(mout.time, [0.1, 0.2, 0.3, 0.35])
    assert np.allclose(mout.value, [100.0 + t for t in (0.1, 0.2, 0.3, 0.35)])   # entropy = kind 1
    assert kout.iter == [2, 3] and np.allclose(kout.value, [0.2, 0.3])            # kinetic energy = kind 0
    assert lib.calls[0] == ("limiter", 0, 0.0) and lib.calls[1] == ("upload",)
    assert lib.calls[-1] == ("download",) or lib.calls[-2] == ("download",)
    disc._h = C.c_void_p()


def test_cfl_callback_sets_dt_before_the_first_step_and_after_every_step(stub):
    lib = stub(max_dts=[1.0, 2.0, 0.5, 4.0, 4.0, 4.0])
    disc, eq = _disc()
    cb = F.make_callback_list(F.get_cfl_callback(0.1, 0.15))                  # dt = min(0.1*max_dt, 0.15)
    Q = disc.new_state()
    sol, _ = F.timeintegrate(Q, disc, eq, F.ORK256(), 0.45, dt=123.0, callback=cb)
    steps = [round(c[1], 12) for c in lib.calls if c[0] == "advance"]
    assert steps == [0.1, 0.15, 0.05, 0.15]          # initialize, then after steps 1..3 (0.2 capped to 0.15)
    assert abs(sol.t[-1] - 0.45) < 1e-15
    disc._h = C.c_void_p()


def test_stage_limiter_is_switched_on_for_the_call_and_off_for_the_next(stub):
    lib = stub()
    disc, eq = _disc()
    lim = F.get_limiter_callback(disc, eq, "zhang_shu", 1e-10)
    F.advance(disc, F.ORK256(stage_limiter=lim), 1e-3, 3)
    F.advance(disc, F.ORK256(), 1e-3, 3)
    assert [c for c in lib.calls if c[0] == "limiter"] == [("limiter", 1, 1e-10), ("limiter", 0, 0.0)]
    with pytest.raises(ValueError):
        F.ORK256(stage_limiter=lambda *a: None)          # only limiters from get_limiter_callback
    with pytest.raises(ValueError):
        F.get_limiter_callback(disc, eq, "zhang_shu")     # minimum value missing
    with pytest.raises(ValueError):
        F.timeintegrate(disc.new_state(), disc, eq, F.ORK256(), 0.1, dt=0.1, callback=[object()])
    disc._h = C.c_void_p()
