"""1-D operator tables: product host mirror vs oracle (both follow the reference's monomial
route) and the SBP identities the DGSEM relies on.  CPU only."""
import numpy as np
import pytest

import flou_b200 as F
from oracle import operators as oo


@pytest.mark.parametrize("nodes", ["GL", "GLL", "CGL"])
@pytest.mark.parametrize("n", [2, 3, 4, 5, 6, 8])
def test_product_operators_match_oracle(nodes, n):
    b = F.LagrangeBasis(nodes, n)
    std = F.StdSegment(b, F.DGSEMrec(b), 1)
    o = oo.operators_1d(nodes, n)
    assert np.max(np.abs(b.xi - o["xi"])) <= 2e-16
    for mine, ref in ((std.w1d, o["w"]), (std.D, o["D"]), (std.Ds, o["Ds"]), (std.Dsharp, o["Dsharp"]),
                      (std.l[0], o["lm"]), (std.l[1], o["lp"]), (std.dg[0], o["dgl"]), (std.dg[1], o["dgr"])):
        assert np.max(np.abs(mine - ref)) <= 1e-13 * max(1.0, np.max(np.abs(ref)))
    assert b.hasboundaries == o["hasboundaries"] == (nodes == "GLL")


@pytest.mark.parametrize("nodes", ["GL", "GLL"])
@pytest.mark.parametrize("n", [3, 4, 5, 6])
def test_sbp_identities(nodes, n):
    o = oo.operators_1d(nodes, n)
    D, w = o["D"], o["w"]
    assert abs(w.sum() - 2.0) < 1e-13
    assert np.max(np.abs(D @ np.ones(n))) < 1e-12                       # derivative of 1
    Qm = np.diag(w) @ D
    B = np.outer(o["lp"], o["lp"]) - np.outer(o["lm"], o["lm"])
    assert np.max(np.abs(Qm + Qm.T - B)) < 1e-12                        # summation by parts
    # Dsharp = 2D - B_lift and Ds = D - B_lift share the lifting matrix (StdSegment.jl:87-89)
    assert np.max(np.abs((o["Dsharp"] - o["Ds"]) - D)) < 1e-13
    if nodes == "GLL":
        e1, en = np.eye(n)[0], np.eye(n)[-1]
        assert np.max(np.abs(o["lm"] - e1)) < 1e-13 and np.max(np.abs(o["lp"] - en)) < 1e-13


def test_tensor_product_node_and_weight_order():
    b = F.LagrangeBasis("GLL", 3)
    rec = F.DGSEMrec(b)
    q, h = F.StdQuad(b, rec, 4), F.StdHex(b, rec, 5)
    # x fastest (StdQuad.jl:46-50, StdHex.jl:48-54)
    assert np.array_equal(q.xi[:3, 0], b.xi) and np.all(q.xi[:3, 1] == b.xi[0])
    assert np.array_equal(h.xi[:3, 0], b.xi) and np.all(h.xi[:9, 2] == b.xi[0])
    assert abs(q.w.sum() - 4) < 1e-13 and abs(h.w.sum() - 8) < 1e-13
    assert q.ndofs() == 9 and h.ndofs() == 27 and h.nfacedofs() == 9


def test_unknown_nodes_raise():
    with pytest.raises(ValueError):
        F.LagrangeBasis("XYZ", 4)
