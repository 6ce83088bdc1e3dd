"""3-D Euler Taylor-Green vortex through the Python harness -- the harness's counterpart of the
reference's examples/src/3D_Euler.jl (same calls, same order): entropy-conservative split form with
matrix dissipation on GLL nodes, ORK256 with the Zhang-Shu stage limiter, entropy monitor and VTKHDF
snapshots written by the save callback.  Needs a B200 (there is no CPU fallback).

    python examples/tgv3d.py [elements per direction = 16] [p = 4] [steps = 20] [output basename]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "flou.jl_b200"))
import flou_b200 as F  # noqa: E402


def main(n=16, p=4, nsteps=20, basename=None):
    dt = 1e-3
    tf = nsteps * dt
    equation = F.EulerEquation(3, 1.4)

    basis = F.LagrangeBasis("GLL", p + 1)
    rec = F.DGSEMrec(basis)
    std = F.StdHex(basis, rec, F.nvariables(equation))
    mesh = F.CartesianMesh(3, (0.0, 0.0, 0.0), (2 * np.pi,) * 3, (n, n, n))
    F.apply_periodicBCs(mesh, ("1", "2"), ("3", "4"), ("5", "6"))

    div = F.SplitDivOperator(F.MatrixDissipation(F.ChandrasekharAverage(), 1.0))
    dg = F.MultielementDisc(mesh, std, equation, div, {})

    # Taylor-Green vortex at Mach 0.1
    x, y, z = dg.coords().T
    rho0, V0, p0 = 1.0, 1.0, 1.0 / (1.4 * 0.1 ** 2)
    u = V0 * np.sin(x) * np.cos(y) * np.cos(z)
    v = -V0 * np.cos(x) * np.sin(y) * np.cos(z)
    pr = p0 + rho0 * V0 ** 2 / 16 * (np.cos(2 * x) + np.cos(2 * y)) * (np.cos(2 * z) + 2)
    rho = np.full_like(x, rho0)
    # vars_prim2cons (src/FlouCommon/Euler.jl:255-271), node by node
    Q = np.asfortranarray(np.stack([rho, rho * u, rho * v, np.zeros_like(x),
                                    pr / (equation.gamma - 1) + rho * (u * u + v * v) / 2], axis=1))

    mb, mvals = F.get_monitor_callback(float, float, dg, equation, "entropy")
    callbacks = [mb]
    if basename:
        callbacks.append(F.get_save_callback(basename, iter=range(0, nsteps + 1, max(nsteps // 2, 1))))
    cb = F.make_callback_list(*callbacks)

    zslimiter = F.get_limiter_callback(dg, equation, "zhang_shu", 1e-10)
    solver = F.ORK256(stage_limiter=zslimiter, williamson_condition=False)

    print(f"Starting simulation: {mesh.nelements} elements, {dg.ndofs} DOF, {nsteps} steps of {dt}")
    sol, exetime = F.timeintegrate(Q, dg, equation, solver, tf, dt=dt, alias_u0=True, adaptive=False, callback=cb)
    if sol is None:
        return 1
    print(f"Elapsed time: {exetime:.3f} s")
    print(f"Time per iteration and DOF: {exetime / nsteps / dg.ndofs:.3e} s")
    print(f"Entropy monitor: first {mvals.value[0]:.12e}, last {mvals.value[-1]:.12e} ({len(mvals.value)} samples)")
    dg.close()
    return 0


if __name__ == "__main__":
    a = sys.argv[1:]
    sys.exit(main(int(a[0]) if len(a) > 0 else 16, int(a[1]) if len(a) > 1 else 4,
                  int(a[2]) if len(a) > 2 else 20, a[3] if len(a) > 3 else None))
