"""BASELINE config 4 AT FULL SIZE against the oracle: 3-D Euler, 128^3 hexahedra, p=4, EC split
form + matrix dissipation -- 262 M DOF, 2 M elements, every persistent CTA of the element kernel
walks ~3500 groups.

The oracle cannot evaluate 262 M DOF in test time, but it does not have to: on a fully periodic
Cartesian mesh the discretisation commutes with periodic tiling.  The state of a 16^3 mesh on the
box B, repeated 8 x 8 x 8 times, is a state of the 128^3 mesh on the box 8 B with the same element
size, and its RHS is the same tiling of the small mesh's RHS, element by element -- every one of
the 2 097 152 elements, whatever CTA, iteration or ring slot computed it, has an oracle value.

One subtlety, and it is the reference's: with `ChandrasekharAverage` inside the 3-D numerical flux
the flux is not symmetric in (master, slave) (the `wl^2` quirk, Equations/Euler.jl:216), and the
merged periodic face keeps the roles of the low side (`_apply_periodicBCs!`, Mesh.jl:236-322), so a
face on the small mesh's periodic seam and the same face in the middle of the big mesh give
different energy fluxes.  The reference values are therefore taken from EIGHT oracle runs on the
small mesh with the state rolled by 0 or 8 elements per direction (the seam then sits at a tile
boundary or in the middle of a tile), and each big-mesh element takes the run in which its faces
have the roles they have in the big mesh: the big mesh's own seam (global index 0 / 127) matches
the unrolled run, tile boundaries inside the big mesh match the rolled run.  With `StdAverage`
inside the numerical flux the roles do not matter, the tiling is exact for the evolution too, and
the state after two RK steps is compared as well."""
import itertools

import numpy as np
import pytest

from common import Case, random_state, smooth_state

N0 = 16


def _tile(a_small, rep, n0, npts):
    """(n0^3 * npts,) in Flou's element order (x fastest) -> the rep^3 tiling, same order."""
    a = a_small.reshape(n0, n0, n0, npts)
    return np.tile(a, (rep, rep, rep, 1)).reshape(-1)


def _rolled_references(orc, Qs, n0, npts, run):
    """run(Q) on the small mesh for the state rolled by 0 / n0/2 elements per direction, rolled
    back: R[sz, sy, sx] of shape (n0, n0, n0, npts, nv) (axes k, j, i)."""
    nv = Qs.shape[1]
    Q5 = np.stack([Qs[:, v].reshape(n0, n0, n0, npts) for v in range(nv)], axis=-1)
    R = np.empty((2, 2, 2, n0, n0, n0, npts, nv))
    for sz, sy, sx in itertools.product((0, 1), repeat=3):
        shift = (sz * (n0 // 2), sy * (n0 // 2), sx * (n0 // 2))
        Qr = np.roll(Q5, shift, axis=(0, 1, 2))
        flat = np.asfortranarray(Qr.reshape(-1, nv))
        out = run(flat)
        out5 = np.stack([out[:, v].reshape(n0, n0, n0, npts) for v in range(nv)], axis=-1)
        R[sz, sy, sx] = np.roll(out5, tuple(-s for s in shift), axis=(0, 1, 2))
    return R


def _choice(nb, n0):
    """Per global index along one direction: 1 = take the rolled run (a tile boundary inside the
    big mesh: ordinary roles), 0 = the unrolled run (the big mesh's own seam, or the middle of a tile)."""
    g = np.arange(nb)
    t = g % n0
    sel = np.zeros(nb, dtype=np.int64)
    sel[((t == 0) | (t == n0 - 1)) & (g != 0) & (g != nb - 1)] = 1
    return sel, t


def _max_err(big, R, rep, n0, npts):
    nb = n0 * rep
    sel, t = _choice(nb, n0)
    err = 0.0
    for v in range(big.shape[1]):
        b = big[:, v].reshape(nb, nb, nb, npts)
        for kb in range(rep):
            ks = slice(kb * n0, (kb + 1) * n0)
            ref = R[sel[ks][:, None, None], sel[None, :, None], sel[None, None, :],
                    t[ks][:, None, None], t[None, :, None], t[None, None, :], :, v]
            err = max(err, float(np.max(np.abs(b[ks] - ref))))
    return err


@pytest.mark.gpu
@pytest.mark.parametrize("avg", ["cha", "std"])
def test_config4_at_full_size_against_the_oracle_by_periodic_tiling(gpu, avg, monkeypatch):
    monkeypatch.delenv("FLOU_B200_XTRACE", raising=False)
    import flou_b200 as F
    import oracle as O
    rep, npn = 8, 5
    npts = npn ** 3
    kw = dict(nodes="GLL", eq="euler", op="split", tp="cha", nf="mat", avg=avg)
    small = Case(3, (N0,) * 3, npn, **kw)
    big = Case(3, (N0 * rep,) * 3, npn, box_scale=float(rep), **kw)
    orc = small.oracle()
    Qs = np.asfortranarray(0.9 * smooth_state(orc.coords, 3, "euler", waves=2)
                           + 0.1 * random_state(orc.ndof, 3, "euler", amp=0.3))
    R = _rolled_references(orc, Qs, N0, npts, orc.rhs)
    scale = float(np.max(np.abs(R)))
    if avg == "cha":      # the quirk is real: the runs differ next to the seam (energy flux only)
        assert np.max(np.abs(R[1, 1, 1] - R[0, 0, 0])[..., 4]) > 1e-6 * scale
        assert np.max(np.abs(R[1, 1, 1] - R[0, 0, 0])[..., :4]) < 1e-12 * scale
    else:
        assert np.max(np.abs(R[1, 1, 1] - R[0, 0, 0])) < 1e-12 * scale

    disc, eq = big.product(kernel="auto")          # what a user gets: the two-kernel production path
    assert disc.ndofs == (N0 * rep) ** 3 * npts == 262144000
    info = disc.kernel_info()
    assert (N0 * rep) ** 3 // info["elems_per_cta_iter"] > 1000 * info["grid_ctas"]
    Q = np.empty((disc.ndofs, 5), order="F")
    for v in range(5):
        Q[:, v] = _tile(Qs[:, v], rep, N0, npts)
    dQ = disc.new_state()
    F.rhs(dQ, Q, F.EquationConfig(disc, eq), 0.0)
    assert _max_err(dQ, R, rep, N0, npts) <= 1e-12 * scale
    del dQ
    if avg == "std":
        dt, nsteps = 2e-5, 2
        ref_state = orc.lsrk2n(Qs, O.ORK256, dt, nsteps)
        Rs = np.broadcast_to(np.stack([ref_state[:, v].reshape(N0, N0, N0, npts) for v in range(5)], axis=-1),
                             (2, 2, 2, N0, N0, N0, npts, 5))
        sol, _ = F.timeintegrate(Q, disc, eq, F.ORK256(), nsteps * dt, dt=dt, save_start=False)
        assert sol is not None
        assert _max_err(sol.u[-1], Rs, rep, N0, npts) <= 1e-10 * float(np.max(np.abs(ref_state)))
    disc.close()


@pytest.mark.parametrize("avg", ["cha", "std"])
def test_tiling_argument_on_a_small_mesh(avg):
    """The argument itself, oracle against oracle (CPU): 8^3 tiled 2x2x2 = 16^3."""
    n0, rep, npn = 8, 2, 3
    npts = npn ** 3
    kw = dict(nodes="GLL", eq="euler", op="split", tp="cha", nf="mat", avg=avg)
    os_ = Case(3, (n0,) * 3, npn, **kw).oracle()
    ob = Case(3, (n0 * rep,) * 3, npn, box_scale=float(rep), **kw).oracle()
    Qs = random_state(os_.ndof, 3, "euler", amp=0.3)
    Qb = np.empty((ob.ndof, 5), order="F")
    for v in range(5):
        Qb[:, v] = _tile(Qs[:, v], rep, n0, npts)
    R = _rolled_references(os_, Qs, n0, npts, os_.rhs)
    rb = ob.rhs(Qb)
    scale = float(np.max(np.abs(rb)))
    assert _max_err(rb, R, rep, n0, npts) <= 1e-13 * scale
    if avg == "cha":      # ... and plain tiling is NOT enough with the reference's asymmetric flux
        plain = np.broadcast_to(R[0, 0, 0], R.shape)
        assert _max_err(rb, plain, rep, n0, npts) > 1e-6 * scale
