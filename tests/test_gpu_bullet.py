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


@pytest.mark.parametrize("n", [2, 16, 512])
def test_bullet_rounds_with_resident_vectors(ctx, n):
    """a and b on the device (spg_bullet_set_ab / _lr_resident / _final_ab): every round's L and R equal those
    of the host-fed rounds, c_L, c_R and the folded a, b equal python integers (src/nizk/bullet.rs:83-84,
    113-116); n = 512 goes through the few-row MSM kernels"""
    import spartan_parallel_b200 as sp

    gens = G.MultiCommitGens(n, b"bullet-test")
    dg = sp.MultiCommitGens(ctx, gens.compressed())
    ref, br = sp.BulletReduction(ctx, dg, n), sp.BulletReduction(ctx, dg, n)
    a = [O.to_int(s) for s in rand_scalars(n, 50 + n)]
    b = [O.to_int(s) for s in rand_scalars(n, 51 + n)]
    if n >= 16:
        a[1] = 0
        b[3] = 0
    br.set_ab(np.stack([O.from_int(x) for x in a]), np.stack([O.from_int(x) for x in b]))
    with pytest.raises(sp.SpgError):
        ref.lr_resident(n, O.from_int(1), O.from_int(2))  # no vectors uploaded
    nk, rnd = n, 0
    while nk != 1:
        nh = nk // 2
        bl = [O.to_int(s) for s in rand_scalars(2, 1900 + 10 * n + rnd)]
        want = ref.lr(np.stack([O.from_int(x) for x in a[:nk]]), O.from_int(bl[0]), O.from_int(bl[1]))
        L, R, cL, cR = br.lr_resident(nk, O.from_int(bl[0]), O.from_int(bl[1]))
        assert (L, R) == want, (n, nk)
        # the same two points as extended coordinates: canonical field elements on the curve's equivalence class
        Le, Re, cL2, cR2 = br.lr_resident(nk, O.from_int(bl[0]), O.from_int(bl[1]), ext=True)
        for enc, pt in ((L, Le), (R, Re)):
            co = [int.from_bytes(pt[32 * k: 32 * (k + 1)], "little") for k in range(4)]
            assert all(c < G.P for c in co)
            assert G.Point(*co).compress() == enc
        assert np.array_equal(cL, cL2) and np.array_equal(cR, cR2)
        assert O.to_int(cL) == sum(a[i] * b[nh + i] for i in range(nh)) % Q
        assert O.to_int(cR) == sum(a[nh + i] * b[i] for i in range(nh)) % Q
        u = O.to_int(rand_scalars(1, 8000 + 10 * n + rnd)[0]) or 1
        u_inv = pow(u, -1, Q)
        for x in (ref, br):
            x.fold(nk, O.from_int(u), O.from_int(u_inv))
        a = [(a[i] * u + u_inv * a[nh + i]) % Q for i in range(nh)]
        b = [(b[i] * u_inv + u * b[nh + i]) % Q for i in range(nh)]
        nk, rnd = nh, rnd + 1
    g_hat, a0, b0 = br.final_ab()
    assert g_hat == ref.final()
    assert O.to_int(a0) == a[0] and O.to_int(b0) == b[0]
    ref.free()
    br.free()


def test_bullet_rejects_bad_sizes(ctx):
    import spartan_parallel_b200 as sp

    dg = sp.MultiCommitGens(ctx, G.MultiCommitGens(8, b"bullet-test").compressed())
    with pytest.raises(sp.SpgError):
        sp.BulletReduction(ctx, dg, 6)
    with pytest.raises(sp.SpgError):
        sp.BulletReduction(ctx, dg, 16)
