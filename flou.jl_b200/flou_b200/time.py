"""FlouTime on the B200 library.

  timeintegrate(Q0, disc, equation, solver, tfinal; adaptive=false, dt, alias_u0=true, ...)
                                   src/FlouTime/FlouTime.jl:34-54
  ORK256 / CarpenterKennedy2N54    OrdinaryDiffEq v6.49.1 `LowStorageRK2N` tableaus
                                   (third-party; used at test/tests.jl:19,56,93,139 and
                                   examples/src/Convergence.jl:27)
The 2N recurrence  tmp = A_s*tmp + dt*k ; u = u + B_s*tmp  runs fused with the RHS inside
the stage kernel; the host only enqueues stages.
"""
import ctypes as C
import time as _time

import numpy as np

from . import _lib as L
from .disc import _ptr, _state


class _LowStorageRK2N:
    A = B = c = ()

    def __init__(self, williamson_condition=False, stage_limiter=None, step_limiter=None):
        if williamson_condition:
            raise ValueError("williamson_condition=true (ArrayFuse path) is not supported; the "
                             "reference always passes williamson_condition=false")
        if step_limiter is not None:
            raise ValueError("step limiters are outside the B200 hot path (the reference uses stage_limiter!)")
        # ORK256(stage_limiter! = get_limiter_callback(dg, eq, :zhang_shu, minval)) as in
        # examples/src/3D_Euler.jl:76-80; applied on the device after every stage
        if stage_limiter is not None and not hasattr(stage_limiter, "minval"):
            raise ValueError("stage_limiter must come from get_limiter_callback(disc, eq, 'zhang_shu', minval)")
        self.stage_limiter = stage_limiter

    @property
    def nstages(self):
        return len(self.B)


class ORK256(_LowStorageRK2N):
    A = (0.0, -1.0, -1.55798, -1.0, -0.45031)
    B = (0.2, 0.83204, 0.6, 0.35394, 0.2)
    c = (0.0, 0.2, 0.2, 0.8, 0.8)


class CarpenterKennedy2N54(_LowStorageRK2N):
    A = (0.0, -567301805773 / 1357537059087, -2404267990393 / 2016746695238,
         -3550918686646 / 2091501179385, -1275806237668 / 842570457699)
    B = (1432997174477 / 9575080441755, 5161836677717 / 13612068292357,
         1720146321549 / 2090206949498, 3134564353537 / 4481467310338,
         2277821191437 / 14882151754819)
    c = (0.0, 1432997174477 / 9575080441755, 2526269341429 / 6820363962896,
         2006345519317 / 3224310063776, 2802321613138 / 2924317926251)


def _tab(solver):
    return (np.array(solver.A, dtype=np.float64), np.array(solver.B, dtype=np.float64),
            np.array(solver.c, dtype=np.float64))


def _set_limiter(disc, solver):
    lim = getattr(solver, "stage_limiter", None)
    L.check(L.lib().flou_b200_set_stage_limiter(disc.handle, 1 if lim is not None else 0,
                                                lim.minval if lim is not None else 0.0))


def _advance_stagewise(disc, solver, dt, nsteps, t0):
    """Source or boundary closures that depend on (Q, t): the host re-tabulates them before every
    stage at t + c_s dt (OrdinaryDiffEq evaluates f(u, p, t + c_s dt)) and launches one stage."""
    A, B, c = _tab(solver)
    lib = L.lib()
    for n in range(int(nsteps)):
        t = t0 + n * dt
        for s in range(solver.nstages):
            disc.refresh(t + c[s] * dt)
            L.check(lib.flou_b200_lsrk2n_stage(disc.handle, float(A[s]), float(B[s]), float(dt), 1 if s == 0 else 0))


def advance(disc, solver, dt, nsteps, t0=0.0):
    """Device-resident fast path: nsteps RK steps on the uploaded state (asynchronous)."""
    A, B, c = _tab(solver)
    _set_limiter(disc, solver)
    if disc.has_dynamic:
        return _advance_stagewise(disc, solver, dt, nsteps, t0)
    L.check(L.lib().flou_b200_lsrk2n_advance(disc.handle, solver.nstages, _ptr(A), _ptr(B), _ptr(c),
                                             float(dt), float(t0), int(nsteps)))


def get_max_dt(q, disc, equation, cfl):
    """FlouCommon.get_max_dt(q, disc, eq, cfl) (MultielementDiscontinuous.jl:162-178): what
    `get_cfl_callback` (FlouTime.jl:92-105) evaluates every step; `q=None` uses the state that
    lives on the device."""
    return disc.get_max_dt(cfl, q)


class Solution:
    """Minimal stand-in for the ODE solution: `u[0]` initial and `u[-1]` final state, `t`."""

    def __init__(self, u0, u1, t0, t1):
        self.u = [u0, u1]
        self.t = [t0, t1]


def timeintegrate(Q0, disc, equation, solver, tfinal, *, dt, adaptive=False, alias_u0=True,
                  saveat=None, save_everystep=False, callback=None, t0=0.0, nsteps=None,
                  save_start=True):
    """Returns (sol, exetime) like the reference; `Q0` is overwritten when alias_u0=True.

    Fixed-step integration (`adaptive=false`).  `callback`: a list from `make_callback_list` of
    monitor / CFL callbacks (flou_b200.monitors); they are evaluated on the device between steps
    (the save callback, FlouBiz output, is not on this path)."""
    if adaptive:
        raise ValueError("adaptive=true is not supported (the reference always uses adaptive=false)")
    if callback is not None:
        return _timeintegrate_callbacks(Q0, disc, equation, solver, tfinal, dt=dt, alias_u0=alias_u0,
                                        callback=callback, t0=t0, save_start=save_start)
    if equation is not disc.equation:
        raise ValueError("`equation` is not the one the discretisation was built with")
    Q = _state(Q0, disc.ndofs, disc.nv, writable=alias_u0)
    # sol.u[1] like `saveat=(0, tf)` in the reference's tests; skipped with save_start=False
    u0 = Q.copy(order="F") if (save_start or not alias_u0) else None
    if not alias_u0:
        Q = u0.copy(order="F")
    # OrdinaryDiffEq with adaptive=false takes full steps of dt and shortens the last one so that it
    # lands on tfinal (tstops); an explicit `nsteps` runs exactly that many full steps
    last = 0.0
    if nsteps is None:
        nsteps, last = _split_steps(t0, tfinal, dt)
    A, B, c = _tab(solver)
    _set_limiter(disc, solver)
    tic = _time.perf_counter()
    try:
        if last == 0.0 and not disc.has_dynamic:
            L.check(L.lib().flou_b200_timeintegrate(disc.handle, _ptr(Q), solver.nstages, _ptr(A),
                                                    _ptr(B), _ptr(c), float(dt), float(t0),
                                                    int(nsteps)))
        else:
            disc.upload(Q)
            for n, h, ts in ((nsteps, dt, t0), (1, last, t0 + nsteps * dt)):
                if h == 0.0:
                    continue
                if disc.has_dynamic:
                    _advance_stagewise(disc, solver, float(h), n, float(ts))
                else:
                    L.check(L.lib().flou_b200_lsrk2n_advance(disc.handle, solver.nstages, _ptr(A), _ptr(B),
                                                             _ptr(c), float(h), float(ts), int(n)))
            disc.download(Q)
            if disc.status() & 1:
                raise L.DomainError("non-positive density/pressure or NaN (Simulation crashed!)")
    except L.DomainError:
        # FlouTime.jl:39-51: log and return `nothing` for the solution
        print("ERROR: Simulation crashed!")
        return None, _time.perf_counter() - tic
    return Solution(u0, Q, t0, t0 + nsteps * dt + last), _time.perf_counter() - tic


def _split_steps(t0, tfinal, dt):
    """(number of full steps of dt, length of the shortened last step or 0.0) from t0 to tfinal."""
    r = (tfinal - t0) / dt
    n = int(round(r))
    if abs(r - n) <= 1e-9 * max(1.0, abs(r)):
        return max(n, 0), 0.0
    n = max(int(np.floor(r)), 0)
    return n, (tfinal - t0) - n * dt


class _Integrator:
    """The fields of OrdinaryDiffEq's integrator the reference's callbacks read."""

    def __init__(self, disc, equation, t, dt):
        self.disc, self.equation, self.t, self.dt, self.iter = disc, equation, t, dt, 0


def _timeintegrate_callbacks(Q0, disc, equation, solver, tfinal, *, dt, alias_u0, callback, t0,
                             save_start):
    """Step-by-step loop with device-side callbacks (FlouTime.jl:56-152): the state is uploaded
    once and downloaded once; each callback costs one small reduction kernel."""
    from .monitors import _Callback
    cbs = list(callback) if isinstance(callback, (list, tuple)) else [callback]
    for cb in cbs:
        if not isinstance(cb, _Callback):
            raise ValueError("only callbacks from flou_b200.monitors (monitor, CFL) run on the B200 path")
    if equation is not disc.equation:
        raise ValueError("`equation` is not the one the discretisation was built with")
    Q = _state(Q0, disc.ndofs, disc.nv, writable=alias_u0)
    u0 = Q.copy(order="F") if (save_start or not alias_u0) else None
    if not alias_u0:
        Q = u0.copy(order="F")
    A, B, c = _tab(solver)
    _set_limiter(disc, solver)
    tic = _time.perf_counter()
    disc.upload(Q)
    integ = _Integrator(disc, equation, float(t0), float(dt))
    for cb in cbs:
        if cb.initialize:
            cb.affect(integ)
    try:
        while integ.t < tfinal - 1e-14 * max(1.0, abs(tfinal)):
            h = min(integ.dt, tfinal - integ.t)       # the last step lands on tfinal (tstops)
            if disc.has_dynamic:
                _advance_stagewise(disc, solver, h, 1, integ.t)
            else:
                L.check(L.lib().flou_b200_lsrk2n_advance(disc.handle, solver.nstages, _ptr(A), _ptr(B),
                                                         _ptr(c), h, integ.t, 1))
            integ.t += h
            integ.iter += 1
            for cb in cbs:
                if cb.selected(integ.iter):
                    cb.affect(integ)
        disc.download(Q)
        if disc.status() & 1:
            raise L.DomainError("non-positive density/pressure or NaN (Simulation crashed!)")
    except L.DomainError:
        print("ERROR: Simulation crashed!")
        return None, _time.perf_counter() - tic
    sol = Solution(u0, Q, t0, integ.t)
    sol.iterations = integ.iter
    return sol, _time.perf_counter() - tic
