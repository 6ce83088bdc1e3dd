import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "flou.jl_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import flou_b200 as F
from common import Case, relerr, smooth_state, random_state
for case in (Case(1, (9,), 4, nodes="GL", op="split", nf="mat", avg="cha"),
             Case(1, (9,), 4, nodes="GL", op="split", nf="cha", avg="cha"),
             Case(1, (9,), 4, nodes="GL", op="split", tp="std", nf="std", avg="std"),
             Case(2, (4, 5), 4, nodes="GL", op="split", nf="cha", avg="cha"),
             Case(1, (9,), 4, nodes="CGL", op="split", nf="mat", avg="cha"),
             Case(1, (9,), 4, nodes="GL", op="strong", nf="mat", avg="cha")):
    for state in ("smooth", "random"):
        orc = case.oracle(); disc, eq = case.product()
        Q = random_state(orc.ndof, case.nd, "euler", amp=case.amp) if state == "random" else smooth_state(orc.coords, case.nd, "euler")
        dQ = disc.new_state(); F.rhs(dQ, Q, F.EquationConfig(disc, eq), 0.0)
        ref = orc.rhs(Q)
        print(f"{case!r} {state}: rhs err {relerr(dQ, ref):.3e}  max|ref| {np.abs(ref).max():.3e}", flush=True)
        if case.nd == 1 and relerr(dQ, ref) > 1e-10 and state == "smooth":
            e = np.abs(dQ - ref).reshape(-1, case.np, 3)
            print("  per-node err (elem 0):", e[0, :, 0], "\n  ref:", ref.reshape(-1, case.np, 3)[0, :, 0], "\n  got:", dQ.reshape(-1, case.np, 3)[0, :, 0])
        disc.close()
