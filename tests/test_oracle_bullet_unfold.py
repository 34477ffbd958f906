"""The algebra behind spg_bullet_* checked on the CPU with the oracle's python ristretto255:
folding the generators every round (src/nizk/bullet.rs:113-118) and keeping the ORIGINAL
generators with per-base scalars s[m] give the same L, R and G_hat."""
import random

from oracle import ristretto as G

Q = (1 << 252) + 27742317777372353535851937790883648493


def test_unfolded_generators_equal_folded():
    rnd = random.Random(5)
    n = 8
    gens = G.MultiCommitGens(n, b"unfold-test")
    a = [rnd.randrange(Q) for _ in range(n)]
    Gf = list(gens.G)          # folded every round, as the reference does
    s = [1] * n                # per-base scalars of the unfolded form
    nk = n
    while nk != 1:
        nh = nk // 2
        # reference: L = <a_L, G_R>, R = <a_R, G_L> over the folded generators
        L_ref = G.multiscalar_mul(a[:nh], Gf[nh:nk])
        R_ref = G.multiscalar_mul(a[nh:nk], Gf[:nh])
        # unfolded: scalars over all n original bases
        rows_L = [(a[(m % nk) - nh] * s[m]) % Q if (m % nk) >= nh else 0 for m in range(n)]
        rows_R = [(a[(m % nk) + nh] * s[m]) % Q if (m % nk) < nh else 0 for m in range(n)]
        assert G.multiscalar_mul(rows_L, gens.G) == L_ref
        assert G.multiscalar_mul(rows_R, gens.G) == R_ref
        u = rnd.randrange(1, Q)
        u_inv = pow(u, -1, Q)
        a = [(a[i] * u + u_inv * a[nh + i]) % Q for i in range(nh)]
        Gf = [Gf[i].mul(u_inv) + Gf[nh + i].mul(u) for i in range(nh)]
        s = [(s[m] * (u if (m % nk) >= nh else u_inv)) % Q for m in range(n)]
        nk = nh
    assert G.multiscalar_mul(s, gens.G) == Gf[0]
