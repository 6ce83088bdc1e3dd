"""Development check (GPU box): a 3-D p=4 mesh large enough for several groups per persistent CTA
(12x12x13 elements by default: several groups per CTA and, for even element counts per group, a
tail group; usage: mid_parity.py [n [np]]), RHS + 6 RK steps against the oracle, with the library
selected by FLOU_B200_LIB."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "flou.jl_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import flou_b200 as F
import oracle as O
from common import Case, relerr, smooth_state, random_state
n = int(sys.argv[1]) if len(sys.argv) > 1 else 12
npn = int(sys.argv[2]) if len(sys.argv) > 2 else 5
case = Case(3, (n, n, n + 1), npn)
orc = case.oracle(); disc, eq = case.product()
Q = random_state(orc.ndof, 3, "euler", amp=0.2)
dQ = disc.new_state(); F.rhs(dQ, Q, F.EquationConfig(disc, eq), 0.0)
e1 = relerr(dQ, orc.rhs(Q))
u = Q.copy(order="F")
sol, _ = F.timeintegrate(u, disc, eq, F.ORK256(williamson_condition=False), 6e-5, dt=1e-5)
e2 = relerr(sol.u[-1], orc.lsrk2n(Q, O.ORK256, 1e-5, 6))
print(f"{case!r}: rhs {e1:.2e} state {e2:.2e}", "OK" if e1 < 1e-12 and e2 < 1e-10 else "FAIL")
disc.close()
