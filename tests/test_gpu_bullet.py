"""GPU parity for the device side of the bullet reduction (spg_bullet_*) against the reference's
own loop restated with the oracle's python ristretto255 (src/nizk/bullet.rs:72-119: L and R from
the FOLDED generators, then G_L[i] = u^-1 G_L[i] + u G_R[i]). The device never folds generators;
the group elements, hence the compressed bytes, must still be identical."""
import numpy as np
import pytest

from oracle import cbind as O
from oracle import ristretto as G
from tests.helpers import rand_scalars

pytestmark = pytest.mark.gpu

Q = (1 << 252) + 27742317777372353535851937790883648493


@pytest.fixture(scope="module")
def ctx():
    import spartan_parallel_b200 as sp

    return sp.Context(0)


@pytest.mark.parametrize("n", [2, 8, 16])
def test_bullet_rounds_match_folded_generators(ctx, n):
    import spartan_parallel_b200 as sp

    gens = G.MultiCommitGens(n, b"bullet-test")
    dg = sp.MultiCommitGens(ctx, gens.compressed())
    br = sp.BulletReduction(ctx, dg, n)
    a = [O.to_int(s) for s in rand_scalars(n, 40 + n)]
    if n >= 8:
        a[1] = 0  # zero scalars are skipped by the MSM
    Gf = list(gens.G)
    nk, rnd = n, 0
    while nk != 1:
        nh = nk // 2
        bl = [O.to_int(s) for s in rand_scalars(2, 900 + 10 * n + rnd)]
        # reference: L = sum a_L[i] G_R[i] + blind_L h, R = sum a_R[i] G_L[i] + blind_R h (without the c * Q terms)
        L = G.multiscalar_mul(a[:nh], Gf[nh:nk]) + gens.h.mul(bl[0])
        R = G.multiscalar_mul(a[nh:nk], Gf[:nh]) + gens.h.mul(bl[1])
        got = br.lr(np.stack([O.from_int(x) for x in a[:nk]]), O.from_int(bl[0]), O.from_int(bl[1]))
        assert got[0] == L.compress() and got[1] == R.compress(), (n, nk)
        u = O.to_int(rand_scalars(1, 7000 + 10 * n + rnd)[0]) or 1
        u_inv = pow(u, -1, Q)
        br.fold(nk, O.from_int(u), O.from_int(u_inv))
        a = [(a[i] * u + u_inv * a[nh + i]) % Q for i in range(nh)]
        Gf = [Gf[i].mul(u_inv) + Gf[nh + i].mul(u) for i in range(nh)]
        nk, rnd = nh, rnd + 1
    assert br.final() == Gf[0].compress()
    br.free()


def test_bullet_rejects_bad_sizes(ctx):
    import spartan_parallel_b200 as sp

    dg = sp.MultiCommitGens(ctx, G.MultiCommitGens(8, b"bullet-test").compressed())
    with pytest.raises(sp.SpgError):
        sp.BulletReduction(ctx, dg, 6)
    with pytest.raises(sp.SpgError):
        sp.BulletReduction(ctx, dg, 16)
