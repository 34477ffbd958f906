"""The oracle's full R1CSProof restatement must be a sound proof system on its own terms:
prove -> serialize -> deserialize -> verify accepts, and tampering with any part rejects.
(The reference records no proof bytes; this is the strongest pin available without cargo.)"""
import numpy as np
import pytest

from oracle import cbind as O
from oracle import protocol as Pr
from oracle import r1cs as R
from oracle import ristretto as G
from tests.helpers import log2, rand_scalars


def sparse_evals(inst, rx, ry_full, rp, num_instances):
    """(A, B, C)(rp, rx, rw ++ ry): what SNARK::verify feeds R1CSProof::verify."""
    trx = O.eq_evals(np.stack(rx)) if len(rx) else O.ONE.reshape(1, 4)
    try_ = O.eq_evals(np.stack(ry_full))
    per = []
    for i in range(inst.num_instances):
        per.append([O.sparse_evaluate_with_tables(*inst.mats[3 * i + m], trx, try_) for m in range(3)])
    if inst.num_instances == 1:
        return per[0]
    Pp = 1 << (num_instances - 1).bit_length()
    eqp = O.eq_evals(np.stack(rp))
    out = []
    for m in range(3):
        acc = O.ZERO
        for i in range(inst.num_instances):
            acc = O.add(acc, O.mul(eqp[i], per[i][m]))
        out.append(acc)
    return out


def run_prove_verify(inst, P, num_proofs, num_inputs, max_y, secs, seed_scalar):
    max_q = max(num_proofs)
    # generators sized for the largest committed witness polynomial (SNARKGens::new does this
    # from the caller's num_vars, src/lib.rs:164-169)
    gens = Pr.R1CSGens(b"gens_r1cs_sat", max(q * y for q, y in zip(num_proofs, num_inputs)))
    t = Pr.Transcript(b"oracle-test")
    tape = Pr.RandomTape(b"proof", seed_scalar)
    proof, ch = Pr.r1cs_prove(inst, P, max_q, num_proofs, max_y, num_inputs, secs, gens, t, tape)
    blob = Pr.serialize_r1cs_proof(proof)
    back = Pr.deserialize_r1cs_proof(blob)
    assert Pr.serialize_r1cs_proof(back) == blob
    comms = [[Pr.poly_commit(ws.poly_w(p), gens.pc.gens_n) for p in range(len(ws.w_mat))] for ws in secs]
    wnp = [[len(ws.w_mat[p]) for p in range(len(ws.w_mat))] for ws in secs]
    wni = [list(ws.num_inputs[: len(ws.w_mat)]) for ws in secs]
    rp, rq_rev, rx, rwy = ch
    evals = sparse_evals(inst, rx, rwy, rp, P)

    def verify(pr, ev=evals):
        return Pr.r1cs_verify(pr, P, max_q, num_proofs, max_y, wnp, wni, comms, inst.max_num_cons, gens, ev, Pr.Transcript(b"oracle-test"))

    out = verify(back)
    assert out is not None, "honest proof rejected"
    for a, b in zip(out, ch):
        assert len(a) == len(b) and all(np.array_equal(x, y) for x, y in zip(a, b))
    return blob, back, verify, evals


def test_prove_verify_single_instance():
    X, Q = 1 << 4, 2
    inst = R.synthetic_instance(X, unit=False, seed=1)
    secs = R.synthetic_witness(X, [Q], seed=2)
    blob, proof, verify, evals = run_prove_verify(inst, 1, [Q], [X], X, secs, rand_scalars(1, 3)[0])
    # tampering: a response scalar, a commitment, the claimed matrix evaluation
    bad = Pr.deserialize_r1cs_proof(blob)
    bad["proof_eq_sc_phase2"]["z"] = O.add(bad["proof_eq_sc_phase2"]["z"], O.ONE)
    assert verify(bad) is None
    bad = Pr.deserialize_r1cs_proof(blob)
    bad["sc_proof_phase1"]["comm_polys"][1] = G.BASEPOINT_COMPRESSED
    assert verify(bad) is None
    assert verify(proof, [O.add(evals[0], O.ONE), evals[1], evals[2]]) is None


def test_unsatisfied_witness_is_rejected():
    X, Q = 1 << 3, 2
    inst = R.synthetic_instance(X)
    secs = R.synthetic_witness(X, [Q], seed=4)
    secs[1].w_mat[0][1][3] = O.add(secs[1].w_mat[0][1][3], O.ONE)  # break one constraint
    max_q = Q
    gens = Pr.R1CSGens(b"gens_r1cs_sat", inst.num_vars)
    proof, ch = Pr.r1cs_prove(inst, 1, max_q, [Q], X, [X], secs, gens, Pr.Transcript(b"oracle-test"), Pr.RandomTape(b"proof", O.ONE))
    comms = [[Pr.poly_commit(ws.poly_w(0), gens.pc.gens_n)] for ws in secs]
    evals = sparse_evals(inst, ch[2], ch[3], ch[0], 1)
    assert Pr.r1cs_verify(proof, 1, max_q, [Q], X, [[Q], [Q]], [[X], [X]], comms, X, gens, evals, Pr.Transcript(b"oracle-test")) is None


@pytest.mark.slow
def test_prove_verify_three_instances_ragged():
    P, X = 3, 1 << 3
    inst = R.synthetic_instance(X, num_instances=P, unit=False, seed=5)
    secs = R.synthetic_witness(X, [4, 2, 1], seed=6)
    run_prove_verify(inst, P, [4, 2, 1], [X] * P, X, secs, rand_scalars(1, 7)[0])


def test_other_opening_variants_verify():
    """PolyEvalProof::prove_batched_points / prove_batched_instances / prove_uni_batched_instances
    (src/dense_mlpoly.rs:531, 689, 1046) as restated in the oracle: every dot-product proof they
    emit is accepted by the oracle's DotProductProofLog verifier against the commitment the
    reference's verifiers derive homomorphically (sum_i L_i C_i) and the claimed evaluations."""
    from oracle import protocol as P
    from oracle import ristretto as G
    from tests.helpers import rand_scalars

    nv = 4
    gens = P.DotProductProofGens(1 << (nv - nv // 2), b"test-open")
    poly = rand_scalars(1 << nv, 500)
    comm = [G.decompress(c) for c in P.poly_commit(poly, gens.gens_n)]
    left = nv // 2
    # two points sharing the left half, one with another left half -> two proofs
    base = rand_scalars(nv, 501)
    p2 = base.copy(); p2[left:] = rand_scalars(nv - left, 502)
    p3 = rand_scalars(nv, 503)
    pts = [list(base), list(p2), list(p3)]
    Zr = [O.dense_evaluate(poly, np.stack(p)) for p in pts]
    t, tape = P.Transcript(b"open"), P.RandomTape(b"proof", rand_scalars(1, 504)[0])
    proofs = P.polyeval_prove_batched_points(poly, pts, Zr, gens, t, tape)
    assert len(proofs) == 2
    # verifier side of the first group (src/dense_mlpoly.rs:625-683): same transcript schedule
    tv = P.Transcript(b"open")
    tv.append_protocol_name(b"polynomial evaluation proof")
    c = tv.challenge_scalar(b"challenge_c")
    L1, R1 = P._factored(pts[0]); _, R2 = P._factored(pts[1])
    Rc = O.vec_add(R1, O.vec_mul(np.tile(c, (R2.shape[0], 1)), R2))
    Zc = P.add(Zr[0], P.mul(c, Zr[1]))
    C_LZ = G.multiscalar_mul([P.sint(x) for x in L1], comm).compress()
    C_Zc = P.commit1(Zc, P.ZERO, gens.gens_1).compress()
    assert P.dplog_verify(proofs[0], Rc.shape[0], gens, tv, list(Rc), C_LZ, C_Zc)
    L3, R3 = P._factored(pts[2])
    C_LZ3 = G.multiscalar_mul([P.sint(x) for x in L3], comm).compress()
    assert P.dplog_verify(proofs[1], R3.shape[0], gens, tv, list(R3), C_LZ3, P.commit1(Zr[2], P.ZERO, gens.gens_1).compress())

    # batched instances: two polynomials of the same size at the same point -> one combined proof
    poly_b = rand_scalars(1 << nv, 505)
    comm_b = [G.decompress(c) for c in P.poly_commit(poly_b, gens.gens_n)]
    r = list(rand_scalars(nv, 506))
    Zi = [O.dense_evaluate(poly, np.stack(r)), O.dense_evaluate(poly_b, np.stack(r))]
    t, tape = P.Transcript(b"open2"), P.RandomTape(b"proof", rand_scalars(1, 507)[0])
    proofs = P.polyeval_prove_batched_instances([poly, poly_b], [r, r], Zi, gens, t, tape)
    assert len(proofs) == 1
    tv = P.Transcript(b"open2")
    tv.append_protocol_name(b"polynomial evaluation proof")
    c = tv.challenge_scalar(b"challenge_c")
    L, R = P._factored(r)
    C_LZ = (G.multiscalar_mul([P.sint(x) for x in L], comm) + G.multiscalar_mul([P.sint(x) for x in L], comm_b).mul(P.sint(c))).compress()
    Zc = P.add(Zi[0], P.mul(c, Zi[1]))
    assert P.dplog_verify(proofs[0], R.shape[0], gens, tv, list(R), C_LZ, P.commit1(Zc, P.ZERO, gens.gens_1).compress())

    # univariate view: sizes 2^4 and 2^2 at one scalar point
    small = rand_scalars(4, 508)
    rr = rand_scalars(1, 509)[0]
    QM = (1 << 252) + 27742317777372353535851937790883648493
    uni = lambda Z: sum(P.sint(z) * pow(P.sint(rr), i, QM) for i, z in enumerate(Z)) % QM
    Zu = [O.from_int(uni(poly)), O.from_int(uni(small))]
    t, tape = P.Transcript(b"open3"), P.RandomTape(b"proof", rand_scalars(1, 510)[0])
    pr, Cy = P.polyeval_prove_uni_batched_instances([poly, small], rr, Zu, gens, t, tape)
    tv = P.Transcript(b"open3")
    tv.append_protocol_name(b"polynomial evaluation proof")
    c = tv.challenge_scalar(b"challenge_c")
    assert Cy == P.commit1(P.add(Zu[0], P.mul(c, Zu[1])), P.ZERO, gens.gens_1).compress()
    assert len(P.serialize_polyeval_proofs([pr])) > 8
