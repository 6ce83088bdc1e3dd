"""Development check (GPU box): RHS + 5 RK steps of a 3-D p=4 case against the oracle with the
library selected by FLOU_B200_LIB."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "flou.jl_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import flou_b200 as F
import oracle as O
from common import Case, relerr, smooth_state, random_state
for case, state in ((Case(3, (2, 3, 2), 5), "random"), (Case(3, (3, 2, 3), 5), "smooth"),
                    (Case(3, (2, 2, 3), 5, perturb_amp=0.08), "random")):
    orc = case.oracle(); disc, eq = case.product()
    Q = random_state(orc.ndof, 3, "euler") if state == "random" else smooth_state(orc.coords, 3, "euler")
    dQ = disc.new_state(); F.rhs(dQ, Q, F.EquationConfig(disc, eq), 0.0)
    e1 = relerr(dQ, orc.rhs(Q))
    u = Q.copy(order="F")
    sol, _ = F.timeintegrate(u, disc, eq, F.ORK256(williamson_condition=False), 5e-4, dt=1e-4)
    e2 = relerr(sol.u[-1], orc.lsrk2n(Q, O.ORK256, 1e-4, 5))
    print(f"{case!r} {state}: rhs {e1:.2e} state {e2:.2e}", "OK" if e1 < 1e-12 and e2 < 1e-10 else "FAIL")
    disc.close()
