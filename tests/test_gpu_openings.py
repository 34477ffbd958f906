"""GPU parity for the PolyEvalProof variants other than the one R1CSProof::prove uses
(prove_batched_points, prove_batched_instances, prove_uni_batched_instances;
src/dense_mlpoly.rs:531, 689, 1046; called from src/lib.rs:2587, 2657, 2673): the C++ host mirror
driving spg_dense_bound_L and the device MSMs must emit the oracle's bytes."""
import numpy as np
import pytest

from oracle import cbind as O
from oracle import protocol as P
from tests.helpers import rand_scalars

pytestmark = pytest.mark.gpu
QM = (1 << 252) + 27742317777372353535851937790883648493


@pytest.fixture(scope="module")
def ctx():
    import spartan_parallel_b200 as sp

    return sp.Context(0)


def _setup(label, seed):
    return P.Transcript(label), P.RandomTape(b"proof", seed)


@pytest.mark.parametrize("nv", [6, 10])
def test_prove_batched_points(ctx, nv):
    import spartan_parallel_b200 as sp
    from spartan_parallel_b200 import host

    gens_n = 1 << (nv - nv // 2)
    poly = rand_scalars(1 << nv, 600 + nv)
    left = nv // 2
    base = rand_scalars(nv, 601)
    p2 = base.copy(); p2[left:] = rand_scalars(nv - left, 602)
    p3 = rand_scalars(nv, 603)
    p4 = p3.copy(); p4[left:] = rand_scalars(nv - left, 604)
    pts = [list(base), list(p3), list(p2), list(p4)]  # groups interleaved: (0, 2) and (1, 3)
    Zr = [O.dense_evaluate(poly, np.stack(p)) for p in pts]
    seed = rand_scalars(1, 605)[0]
    t, tape = _setup(b"open-points", seed)
    want = P.serialize_polyeval_proofs(P.polyeval_prove_batched_points(poly, pts, Zr, P.DotProductProofGens(gens_n, b"gens-open"), t, tape))
    got = host.polyeval_prove(ctx, "points", [sp.DensePolynomial.new(ctx, poly)], pts, Zr, b"open-points", b"gens-open", seed, gens_n)
    assert got == want


def test_prove_batched_instances(ctx):
    """polynomials of different sizes: the point is padded with leading zeros for the larger ones and
    trimmed for the smaller ones; equal (size, R) pairs are combined"""
    import spartan_parallel_b200 as sp
    from spartan_parallel_b200 import host

    sizes = [10, 8, 10, 12, 8]
    polys = [rand_scalars(1 << nv, 610 + i) for i, nv in enumerate(sizes)]
    r = list(rand_scalars(10, 620))
    rs = [r] * len(polys)
    Zr = []
    for poly, nv in zip(polys, sizes):
        rr = [P.ZERO] * (nv - len(r)) + r if nv >= len(r) else r[len(r) - nv:]
        Zr.append(O.dense_evaluate(poly, np.stack(rr)))
    seed = rand_scalars(1, 621)[0]
    gens_n = 1 << 6
    t, tape = _setup(b"open-inst", seed)
    proofs = P.polyeval_prove_batched_instances(polys, rs, Zr, P.DotProductProofGens(gens_n, b"gens-open"), t, tape)
    assert len(proofs) == 3
    want = P.serialize_polyeval_proofs(proofs)
    got = host.polyeval_prove(ctx, "instances", [sp.DensePolynomial.new(ctx, p_) for p_ in polys], rs, Zr, b"open-inst", b"gens-open", seed, gens_n)
    assert got == want


def test_prove_uni_batched_instances(ctx):
    import spartan_parallel_b200 as sp
    from spartan_parallel_b200 import host

    sizes = [10, 6, 10, 2]
    polys = [rand_scalars(1 << nv, 630 + i) for i, nv in enumerate(sizes)]
    r = rand_scalars(1, 640)[0]
    ri = O.to_int(r)
    Zr = [O.from_int(sum(O.to_int(z) * pow(ri, i, QM) for i, z in enumerate(Z)) % QM) for Z in polys]
    seed = rand_scalars(1, 641)[0]
    gens_n = 1 << 5
    t, tape = _setup(b"open-uni", seed)
    pr, Cy = P.polyeval_prove_uni_batched_instances(polys, r, Zr, P.DotProductProofGens(gens_n, b"gens-open"), t, tape)
    w = P.Writer()
    P.w_dplog(w, pr)
    got, got_cy = host.polyeval_prove(ctx, "uni", [sp.DensePolynomial.new(ctx, p_) for p_ in polys], r, Zr, b"open-uni", b"gens-open", seed, gens_n)
    assert got == bytes(w.b) and got_cy == Cy
