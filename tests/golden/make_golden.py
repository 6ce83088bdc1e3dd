"""Generates the committed golden fixtures (run in the build container, where the reference
is mounted at /root/reference):

  cylinder_p5.npz   config 5 on the reference's own test/meshes/2D_cylinder.msh (73 quads, 60
                    faces with orientation 1): mesh tables as parsed + the ORACLE's RHS for a
                    seeded state, p=5, EC split form + MatrixDissipation, Hole/Top/Bottom slip
                    walls, Left inflow, Right outflow.
  cart3d_p3.npz     config 3 family: 3-D Euler 3x3x3, p=3, periodic: seeded state + oracle RHS
                    + oracle state after 5 ORK256 steps.

The GPU tests compare the CUDA path against these files AND against the live oracle, so a
silent change of the oracle is caught as well.   python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "flou.jl_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import flou_b200 as F          # noqa: E402  (host-side reader only; no GPU needed)
import oracle as O             # noqa: E402
from common import Case, random_state, smooth_state   # noqa: E402
from unstructured import build_pair, euler_bcs        # noqa: E402


def main():
    msh = "/root/reference/test/meshes/2D_cylinder.msh"
    raw = F.read_msh(msh)
    names = [g[0] for g in raw.groups]
    bcs = euler_bcs(names)
    orc, _, _ = build_pair(msh, 6, bcs, create=False)
    Q = random_state(orc.ndof, 2, "euler")
    np.savez_compressed(
        os.path.join(HERE, "cylinder_p5.npz"), nodes=raw.nodes, quads=raw.quads, lines=raw.lines,
        line_tags=raw.line_tags, line_entity=raw.line_entity,
        group_names=np.array(names), group_entities=np.array([",".join(map(str, g[1])) for g in raw.groups]),
        Q=Q, dQ=orc.rhs(Q))
    case = Case(3, (3, 3, 3), 4)
    orc = case.oracle()
    Q = smooth_state(orc.coords, 3, "euler")
    np.savez_compressed(os.path.join(HERE, "cart3d_p3.npz"), Q=Q, dQ=orc.rhs(Q),
                        u5=orc.lsrk2n(Q, O.ORK256, 1e-3, 5))
    print("golden fixtures written")


if __name__ == "__main__":
    main()
