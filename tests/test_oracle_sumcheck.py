"""Self-consistency of the oracle's sumcheck restatement (CPU): the round
polynomials it emits must satisfy the sumcheck relations, and the final claims must
equal independent MLE evaluations of the input tables. Also replays the reference's
UniPoly and DensePolynomial known-answer tests (unipoly.rs:122-182,
dense_mlpoly.rs:1233-1365)."""
import numpy as np
import pytest

from oracle import cbind as O
from tests.helpers import drive_sc1_oracle, log2, rand_scalars


def F(v):
    return O.from_u64(v)


def test_unipoly_kats():
    # unipoly.rs:126-152: 2x^2+3x+1 from evals (1, 6, 15)
    co = O.unipoly_from_evals(np.stack([F(1), F(6), F(15)]))
    assert [O.to_int(c) for c in co] == [1, 3, 2]
    assert O.to_int(O.unipoly_evaluate(co, F(3))) == 28
    # unipoly.rs:154-181: x^3+2x^2+3x+1 from evals (1, 7, 23, 55)
    co = O.unipoly_from_evals(np.stack([F(1), F(7), F(23), F(55)]))
    assert [O.to_int(c) for c in co] == [1, 3, 2, 1]
    assert O.to_int(O.unipoly_evaluate(co, F(4))) == 109


def test_dense_kats():
    # dense_mlpoly.rs:1233-1252
    Z = np.stack([F(1), F(2), F(1), F(4)])
    r = np.stack([F(4), F(3)])
    assert O.to_int(O.dense_evaluate(Z, r)) == 28
    L, R = O.eq_evals(r[:1]), O.eq_evals(r[1:])
    LZ = O.dense_bound_L(Z, L)
    assert O.to_int(O.dot(LZ, R)) == 28


def test_eq_evals_vs_naive():
    # dense_mlpoly.rs:1322-1365
    r = rand_scalars(6, 5)
    got = O.eq_evals(r)
    one = O.ONE
    for i in range(64):
        acc = one
        for j in range(6):
            bit = (i >> (5 - j)) & 1
            acc = O.mul(acc, r[j] if bit else O.sub(one, r[j]))
        assert np.array_equal(got[i], acc)
    L, R = O.eq_evals(r[:3]), O.eq_evals(r[3:])
    for i in range(64):
        assert np.array_equal(got[i], O.mul(L[i >> 3], R[i & 7]))


def _natural_weighted_sum(num_proofs, max_q, num_cons, max_x, Az, Bz, Cz, tau_p, tau_q, tau_x):
    """sum_{p,q,x} eq(tau,(p,q,x)) (Az*Bz - Cz) computed from natural-order tables with
    an independent eq formulation (bit k of x pairs with tau_x[k])."""
    P = len(num_proofs)
    Pp = 1 if P == 1 else 1 << (P - 1).bit_length()
    Ap = O.eq_evals(tau_p) if log2(Pp) else O.ONE.reshape(1, 4)
    total = O.ZERO
    off = 0
    one = O.ONE

    def eq_lsb(tau, idx):
        acc = one
        for k in range(len(tau)):
            acc = O.mul(acc, tau[k] if (idx >> k) & 1 else O.sub(one, tau[k]))
        return acc

    for p in range(P):
        for q in range(num_proofs[p]):
            wq = O.mul(Ap[p], eq_lsb(tau_q, q))
            for x in range(num_cons[p]):
                i = off + q * num_cons[p] + x
                f = O.sub(O.mul(Az[i], Bz[i]), Cz[i])
                total = O.add(total, O.mul(O.mul(wq, eq_lsb(tau_x, x)), f))
        off += num_proofs[p] * num_cons[p]
    return total


@pytest.mark.parametrize("num_proofs,num_cons", [([1], [8]), ([4], [16]), ([4, 2, 1], [8, 4, 8]), ([2, 2, 2, 1, 1], [4, 4, 2, 1, 4])])
def test_sc1_oracle_is_a_valid_sumcheck(num_proofs, num_cons):
    P = len(num_proofs)
    max_q, max_x = max(num_proofs), max(num_cons)
    Pp = 1 if P == 1 else 1 << (P - 1).bit_length()
    N = sum(q * x for q, x in zip(num_proofs, num_cons))
    Az, Bz, Cz = rand_scalars(N, 1), rand_scalars(N, 2), rand_scalars(N, 3)
    tau_p, tau_q, tau_x = rand_scalars(3, 4)[: log2(Pp)], rand_scalars(8, 5)[: log2(max_q)], rand_scalars(8, 6)[: log2(max_x)]
    rounds = log2(Pp) + log2(max_q) + log2(max_x)
    ch = rand_scalars(rounds, 7)
    evals, final = drive_sc1_oracle(num_proofs, max_q, num_cons, max_x, Az, Bz, Cz, tau_p, tau_q, tau_x, ch)
    claim = _natural_weighted_sum(num_proofs, max_q, num_cons, max_x, Az, Bz, Cz, tau_p, tau_q, tau_x)
    for j in range(rounds):
        e0, e2, e3 = evals[j]
        co = O.unipoly_from_evals(np.stack([e0, O.sub(claim, e0), e2, e3]))
        # degree-3 interpolation must reproduce the evaluations it was built from
        assert np.array_equal(O.unipoly_evaluate(co, F(2)), e2)
        assert np.array_equal(O.unipoly_evaluate(co, F(3)), e3)
        claim = O.unipoly_evaluate(co, ch[j])
    tau_claim, az, bz, cz = final
    assert np.array_equal(claim, O.mul(tau_claim, O.sub(O.mul(az, bz), cz)))
    # final claims are the MLE evaluations of the tables: check Az against Pqx.evaluate
    nx, nq = log2(max_x), log2(max_q)
    rx, rq, rp = ch[:nx], ch[nx:nx + nq], ch[nx + nq:]
    A = O.Pqx.new_rev(Az, 1, num_proofs, max_q, num_cons, max_x)
    assert np.array_equal(A.evaluate(rp, rq, np.zeros((0, 4), dtype=np.uint64), rx), az)
    # and tau_claim = eq(tau_x, rx) eq(tau_q, rq) eq(tau_p, rp)
    want = O.ONE
    for t, r in ((tau_x, rx), (tau_q, rq), (tau_p, rp)):
        if len(t):
            want = O.mul(want, O.eq_evaluate(t, r))
    assert np.array_equal(tau_claim, want)


@pytest.mark.parametrize("unit,P", [(True, 1), (False, 1), (False, 3)])
def test_first_round_of_a_satisfying_witness_vanishes_at_0_and_1(unit, P):
    """What spg_sc1_set_satisfied relies on (csrc/sc1.cu): for a witness that satisfies the instance
    Az * Bz - Cz is zero at EVERY row, so the first round polynomial of the phase-1 sumcheck
    (src/sumcheck.rs:1166-1245) is zero at 0 and at 1 -- the reference computes those zeros -- and, being
    l(t) * G(t) with G quadratic and G(0) = G(1) = 0, it is fixed by its value at 2: e(3) l(2) = 3 e(2) l(3)
    with l(t) = eq(tau, t) for the variable bound first. Unit and general coefficients, several instances."""
    from oracle import r1cs as R

    X, Q = 1 << 6, 4
    inst = R.synthetic_instance(X, num_instances=P, unit=unit, seed=3)
    secs = R.synthetic_witness(X, [Q] * P, seed=9)
    Pp = 1 if P == 1 else 1 << (P - 1).bit_length()
    np_, nq, nx = log2(Pp), log2(Q), log2(X)
    big = rand_scalars(64, 77)
    tau_p, tau_q, tau_x = big[:np_], big[8:8 + nq], big[16:16 + nx]
    ch1, ch2, r_abc = rand_scalars(np_ + nq + nx, 78), rand_scalars(np_ + 1 + nx, 79), rand_scalars(3, 80)
    t = R.prove_tables(inst, P, Q, [Q] * P, X, [X] * P, secs, tau_p, tau_q, tau_x, ch1, r_abc, ch2)
    # every row satisfied
    assert np.array_equal(O.vec_sub(O.vec_mul(t.Az.reshape(-1, 4), t.Bz.reshape(-1, 4)), t.Cz.reshape(-1, 4)),
                          np.zeros((t.Az.size // 4, 4), dtype=np.uint64))
    e0, e2, e3 = t.evals1[0]
    assert O.to_int(e0) == 0                      # e(1) = claim - e(0) = 0 as well: the claim of phase 1 is zero
    q = (1 << 252) + 27742317777372353535851937790883648493
    e2i, e3i = O.to_int(e2), O.to_int(e3)
    assert e2i != 0
    # the variable bound first pairs entries (2i, 2i + 1) of x: its eq factor is one of the tau_x -- the relation
    # holds for exactly that one
    def line(tau, tt):
        return ((1 - tau) * (1 - tt) + tau * tt) % q
    hits = [k for k in range(nx) if (e3i * line(O.to_int(tau_x[k]), 2) - 3 * e2i * line(O.to_int(tau_x[k]), 3)) % q == 0]
    assert len(hits) == 1
