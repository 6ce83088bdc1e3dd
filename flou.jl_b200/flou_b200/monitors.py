"""Monitors, limiters and the callbacks built from them, evaluated on the device.

  list_monitors / get_monitor        src/FlouSpatial/Equations/Euler.jl:541-593
  list_limiters / get_limiter        src/FlouSpatial/Equations/Euler.jl:597-660 (zhang_shu_limiter)
  get_cfl_callback                   src/FlouTime/FlouTime.jl:92-105
  get_monitor_callback, MonitorOutput   src/FlouTime/FlouTime.jl:107-134
  get_limiter_callback               src/FlouTime/FlouTime.jl:136-148
  make_callback_list                 src/FlouTime/FlouTime.jl:150-152

The reference evaluates all of these on the host from `integrator.u`; here the state stays on
the device and only the scalar result crosses PCIe.
"""
import ctypes as C

from . import _lib as L
from .disc import _ptr, _state

_MONITORS = {"kinetic_energy": L.MONITOR_KINETIC_ENERGY, "energy": L.MONITOR_KINETIC_ENERGY,
             "entropy": L.MONITOR_ENTROPY}


def list_monitors(disc, equation):
    return ("kinetic_energy", "entropy") if equation.kind == L.EQ_EULER else ()


def get_monitor(disc, equation, name, p=None):
    """Returns `(Q, disc, equation) -> value`; `Q=None` reads the device-resident state.  The
    reference lists `:kinetic_energy` but dispatches on `:energy` (Euler.jl:541-557): both
    names are accepted."""
    if equation.kind != L.EQ_EULER or name not in _MONITORS:
        raise ValueError(f"Unknown monitor '{name}'.")
    kind = _MONITORS[name]

    def monitor(Q, disc_, equation_=None):
        v = C.c_double(0.0)
        q = None if Q is None else _ptr(_state(Q, disc_.ndofs, disc_.nv))
        L.check(L.lib().flou_b200_monitor(disc_.handle, kind, q, C.byref(v)))
        return v.value
    return monitor


def list_limiters(disc, equation):
    return ("zhang_shu",) if equation.kind == L.EQ_EULER else ()


def get_limiter(disc, equation, name, p=None):
    """Returns `(Q, disc, equation) -> None` limiting Q in place (`Q=None`: the device state)."""
    if equation.kind != L.EQ_EULER or name != "zhang_shu":
        raise ValueError(f"Unknown limiter '{name}'.")
    if p is None:
        raise ValueError("The minimum value must be specified when using the limiter of Zhang & Shu.")
    minval = float(p)

    def limiter(Q, disc_, equation_=None):
        q = None if Q is None else _ptr(_state(Q, disc_.ndofs, disc_.nv, writable=True))
        L.check(L.lib().flou_b200_zhang_shu(disc_.handle, q, minval))
    limiter.minval = minval
    limiter.name = name
    return limiter


class StageLimiter:
    """What `get_limiter_callback` returns: pass it as `ORK256(stage_limiter=...)`; the RK loop
    then applies the limiter on the device after every stage."""

    def __init__(self, disc, equation, name, p):
        self.limiter = get_limiter(disc, equation, name, p)
        self.minval = self.limiter.minval


def get_limiter_callback(disc, equation, name, p=None):
    return StageLimiter(disc, equation, name, p)


class _Callback:
    """DiscreteCallback(condition, affect; initialize): `iter=True` fires after every step, an
    iterable of step numbers restricts it (FlouTime.jl:59-65)."""
    initialize = False

    def __init__(self, iter=True):
        self.iter = iter

    def selected(self, it):
        return True if self.iter is True else it in self.iter


class MonitorOutput:
    def __init__(self):
        self.time, self.iter, self.value = [], [], []


class MonitorCallback(_Callback):
    def __init__(self, disc, equation, name, p=None, iter=True):
        super().__init__(iter)
        self.func = get_monitor(disc, equation, name, p)
        self.output = MonitorOutput()

    def affect(self, integ):
        self.output.time.append(integ.t)
        self.output.iter.append(integ.iter)
        self.output.value.append(self.func(None, integ.disc, integ.equation))


def get_monitor_callback(timetype, valuetype, disc, equation, name, p=None, *, iter=True):
    cb = MonitorCallback(disc, equation, name, p, iter)
    return cb, cb.output


class CFLCallback(_Callback):
    initialize = True       # the reference's `initialize` calls `affect` before the first step

    def __init__(self, cfl, dtmax=float("inf"), iter=True):
        super().__init__(iter)
        self.cfl, self.dtmax = float(cfl), float(dtmax)

    def affect(self, integ):
        integ.dt = min(integ.disc.get_max_dt(self.cfl, None), self.dtmax)


def get_cfl_callback(cfl, dtmax=float("inf"), *, iter=True):
    return CFLCallback(cfl, dtmax, iter)


def make_callback_list(*callbacks):
    return list(callbacks)
