"""GPU parity for the R1CSProof::prove table pipeline through the C ABI:
z_mat assembly, multiply_vec_block (SpMV), phase-1 rounds, ABC table, Z bound to rq,
phase-2 rounds -- against oracle.r1cs.prove_tables with the same injected challenges.
Bit-exact at every round."""
import numpy as np
import pytest

from oracle import cbind as O
from oracle import r1cs as R
from tests.helpers import log2, rand_scalars, random_instance, random_witness_secs

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import spartan_parallel_b200 as sp

    return sp.Context(0)


def gpu_pipeline(ctx, inst, P, max_q, num_proofs, max_y, num_inputs, secs, tau_p, tau_q, tau_x, ch1, r_abc, ch2, claim=None,
                 satisfied=False):
    import spartan_parallel_b200 as sp

    A = [inst.mats[3 * i] for i in range(inst.num_instances)]
    B = [inst.mats[3 * i + 1] for i in range(inst.num_instances)]
    Cm = [inst.mats[3 * i + 2] for i in range(inst.num_instances)]
    dinst = sp.R1CSInstance(ctx, inst.num_instances, inst.max_num_cons, inst.num_cons, inst.num_vars, A, B, Cm)
    dsecs = []
    for ws in secs:
        nq = [len(ws.w_mat[p]) for p in range(len(ws.w_mat))]
        flat = np.concatenate([np.concatenate(ws.w_mat[p]) for p in range(len(ws.w_mat))])
        dsecs.append(sp.ProverWitnessSecInfo(ctx, nq, ws.num_inputs, flat))
    z = sp.ZMat(ctx, num_proofs, num_inputs, dsecs)
    block_cons = [inst.num_cons[0]] * P if inst.num_instances == 1 else inst.num_cons
    sc1 = sp.sumcheck_phase1(ctx, dinst, z, num_proofs, max_q, block_cons, inst.max_num_cons, max_y, tau_p, tau_q, tau_x)
    if satisfied:
        sc1.set_satisfied()
    elif claim is not None:
        sc1.set_claim(claim)
    e1 = []
    for j in range(sc1.num_rounds):
        e1.append(sc1.round_eval())
        sc1.round_bind(ch1[j])
    c1 = sc1.final()
    nx, nq_ = log2(inst.max_num_cons), log2(max_q)
    r = np.asarray(ch1).reshape(-1, 4)
    rx = r[:nx][::-1].copy()
    rq_rev, rp = r[nx:nx + nq_], r[nx + nq_:]
    sc2 = sp.SumcheckPhase2(ctx, dinst, z, num_proofs, max_q, num_inputs, max_y, len(secs), rx, rq_rev, rp, *r_abc)
    e2 = []
    for j in range(sc2.num_rounds):
        e2.append(sc2.round_eval())
        sc2.round_bind(ch2[j])
    c2 = sc2.final()
    return e1, c1, e2, c2, dinst, dsecs


def run_case(ctx, inst, P, num_proofs, num_inputs, max_y, secs, seed, claim=None, satisfied=False):
    max_q = max(num_proofs)
    Pp = 1 if P == 1 else 1 << (P - 1).bit_length()
    W = len(secs)
    Wp = 1 if W == 1 else 1 << (W - 1).bit_length()
    np_, nq, nx, ny, nw = log2(Pp), log2(max_q), log2(inst.max_num_cons), log2(max_y), log2(Wp)
    big = rand_scalars(64, seed)
    tau_p, tau_q, tau_x = big[:np_], big[8:8 + nq], big[16:16 + nx]
    ch1 = rand_scalars(max(np_ + nq + nx, 1), seed + 1)
    ch2 = rand_scalars(max(np_ + nw + ny, 1), seed + 2)
    r_abc = rand_scalars(3, seed + 3)
    want = R.prove_tables(inst, P, max_q, num_proofs, max_y, num_inputs, secs, tau_p, tau_q, tau_x, ch1, r_abc, ch2)
    e1, c1, e2, c2, dinst, _ = gpu_pipeline(ctx, inst, P, max_q, num_proofs, max_y, num_inputs, secs, tau_p, tau_q, tau_x, ch1, r_abc, ch2, claim,
                                            satisfied)
    assert len(e1) == len(want.evals1) and len(e2) == len(want.evals2)
    for j, (g, w) in enumerate(zip(e1, want.evals1)):
        assert np.array_equal(g, w), f"phase 1 round {j}"
    assert np.array_equal(c1, want.claims1)
    for j, (g, w) in enumerate(zip(e2, want.evals2)):
        assert np.array_equal(g, w), f"phase 2 round {j}"
    assert np.array_equal(c2, want.claims2)
    return dinst, want


def test_c1_synthetic(ctx):
    """BASELINE config C1: X = 2^10, Q = 4, one instance, sections (u, v)."""
    X, Q = 1 << 10, 4
    inst = R.synthetic_instance(X)
    secs = R.synthetic_witness(X, [Q], seed=11)
    dinst, want = run_case(ctx, inst, 1, [Q], [X], X, secs, seed=100)
    # a satisfying witness makes the phase-1 claim vanish at every round-0 point
    assert np.array_equal(want.evals1[0][0], O.ZERO)


@pytest.mark.parametrize("log_x,Q,P", [(10, 4, 1), (9, 8, 1), (8, 4, 3)])
def test_supplied_claim_two_point_first_round(ctx, log_x, Q, P):
    """spg_sc1_set_claim(0), what R1CSProof::prove passes (src/r1csproof.rs:330): the first round
    then evaluates two points per pair and takes e(1) from the claim like the reference
    (src/sumcheck.rs:1250-1256); bit-exact for a satisfying witness."""
    X = 1 << log_x
    inst = R.synthetic_instance(X, num_instances=P, unit=(P == 1), seed=7)
    secs = R.synthetic_witness(X, [Q] * P, seed=15)
    run_case(ctx, inst, P, [Q] * P, [X] * P, X, secs, seed=105, claim=O.ZERO)


@pytest.mark.parametrize("log_x,Q,P", [(10, 4, 1), (9, 8, 1), (8, 4, 3), (12, 16, 1)])
def test_satisfied_one_point_first_round(ctx, log_x, Q, P):
    """spg_sc1_set_satisfied: the witness satisfies the instance row by row, so the fused SpMV + first round
    evaluates t = 2 only (e(0) = e(1) = 0 are sums of zeros); every round and claim still equals the oracle's,
    which computes the zeros like the reference does. Unit rows (P = 1) and general rows (P = 3)."""
    X = 1 << log_x
    inst = R.synthetic_instance(X, num_instances=P, unit=(P == 1), seed=7)
    secs = R.synthetic_witness(X, [Q] * P, seed=15)
    run_case(ctx, inst, P, [Q] * P, [X] * P, X, secs, seed=106, satisfied=True)


def test_set_claim_after_first_round_is_refused(ctx):
    import spartan_parallel_b200 as sp

    X, Q = 1 << 9, 2
    tabs = [rand_scalars(X * Q, s) for s in (1, 2, 3)]
    sc1 = sp.SumcheckPhase1.from_tables(ctx, [Q], Q, [X], X, *tabs, rand_scalars(0, 1), rand_scalars(1, 4), rand_scalars(9, 5))
    sc1.round_eval()
    with pytest.raises(sp.SpgError, match="before the first round"):
        sc1.set_claim(O.ZERO)


def test_synthetic_non_unit_coefficients(ctx):
    X, Q = 1 << 6, 8
    inst = R.synthetic_instance(X, unit=False, seed=3)
    secs = R.synthetic_witness(X, [Q], seed=12)
    run_case(ctx, inst, 1, [Q], [X], X, secs, seed=101)


def test_three_distinct_instances(ctx):
    P, X = 3, 1 << 7
    inst = R.synthetic_instance(X, num_instances=P, unit=False, seed=5)
    secs = R.synthetic_witness(X, [4, 2, 2], seed=13)
    run_case(ctx, inst, P, [4, 2, 2], [X] * P, X, secs, seed=102)


def test_shared_instance_single_inst(ctx):
    """One R1CS instance shared by P proving instances (perm-root call site, src/lib.rs:2472)."""
    P, X = 3, 1 << 5
    inst = R.synthetic_instance(X)
    secs = R.synthetic_witness(X, [4, 4, 1], seed=14)
    run_case(ctx, inst, P, [4, 4, 1], [X] * P, X, secs, seed=103)


def test_c4_heterogeneous(ctx):
    """BASELINE config C4 shape: P = 5, Q_p = {64,16,16,4,1}, W = 5 with a single section,
    ragged num_inputs and num_cons."""
    P, W, Ymax = 5, 5, 1 << 6
    num_proofs = [16, 8, 8, 4, 1]
    num_cons = [128, 64, 128, 32, 16]
    Y = [64, 64, 32, 64, 16]
    inst = random_instance(P, num_cons, W, Ymax, Y, nnz=150, seed=21)
    kinds = ["full", "single", "full", "short", "full"]
    sec_inputs = [Y, [32] * P, [64, 128, 16, 64, 16], [8] * P, [8] * P]
    secs = random_witness_secs(P, num_proofs, W, sec_inputs, kinds, seed=22)
    run_case(ctx, inst, P, num_proofs, Y, Ymax, secs, seed=104)


def test_multi_evaluate(ctx):
    inst = random_instance(2, [32, 16], 3, 16, [16, 8], nnz=90, seed=31)
    import spartan_parallel_b200 as sp

    A = [inst.mats[3 * i] for i in range(2)]
    B = [inst.mats[3 * i + 1] for i in range(2)]
    Cm = [inst.mats[3 * i + 2] for i in range(2)]
    d = sp.R1CSInstance(ctx, 2, 32, [32, 16], inst.num_vars, A, B, Cm)
    rx, ry = rand_scalars(5, 1), rand_scalars(log2(inst.num_vars), 2)
    got = d.multi_evaluate(rx, ry)
    trx, try_ = O.eq_evals(rx), O.eq_evals(ry)
    for m in range(6):
        rows, cols, vals = inst.mats[m]
        assert np.array_equal(got[m], O.sparse_evaluate_with_tables(rows, cols, vals, trx, try_)), m


def test_witness_poly_evaluate(ctx):
    """polyeval step (src/r1csproof.rs:534-573): poly_w[p].evaluate(rq_short ++ ry_short)."""
    import spartan_parallel_b200 as sp

    Q, Y = 8, 32
    w = rand_scalars(Q * Y, 5)
    sec = sp.ProverWitnessSecInfo(ctx, [Q], [Y], w)
    r = rand_scalars(8, 6)
    assert np.array_equal(sec.poly_w(0).evaluate(r), O.dense_evaluate(w, r))


def test_bad_shapes_rejected(ctx):
    import spartan_parallel_b200 as sp

    inst = R.synthetic_instance(8)
    with pytest.raises(sp.SpgError):
        sp.R1CSInstance(ctx, 1, 12, [8], 16, [inst.mats[0]], [inst.mats[1]], [inst.mats[2]])  # not a power of two
    bad = (inst.mats[0][0], (inst.mats[0][1] + 100).astype(np.uint32), inst.mats[0][2])
    with pytest.raises(sp.SpgError):
        sp.R1CSInstance(ctx, 1, 8, [8], 16, [bad], [inst.mats[1]], [inst.mats[2]])  # column out of range


def test_golden_round_tables(ctx):
    """device rounds against the committed fixture tests/golden/tables_p3_ragged.npz"""
    import os

    P, X = 3, 1 << 4
    num_proofs = [4, 2, 1]
    inst = R.synthetic_instance(X, num_instances=P, unit=False, seed=5)
    secs = R.synthetic_witness(X, num_proofs, seed=6)
    big = rand_scalars(64, 40)
    tau_p, tau_q, tau_x = big[:2], big[8:10], big[16:20]
    ch1, ch2, r_abc = rand_scalars(8, 41), rand_scalars(2 + 1 + 4, 42), rand_scalars(3, 43)
    e1, c1, e2, c2, _, _ = gpu_pipeline(ctx, inst, P, 4, num_proofs, X, [X] * P, secs, tau_p, tau_q, tau_x, ch1, r_abc, ch2)
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tables_p3_ragged.npz"))
    assert np.array_equal(np.stack(e1), g["evals1"]) and np.array_equal(np.asarray(c1), g["claims1"])
    assert np.array_equal(np.stack(e2), g["evals2"]) and np.array_equal(np.asarray(c2), g["claims2"])


def test_checked_claim_detects_an_unsatisfied_witness(ctx):
    """spg_sc1_set_claim_checked: same rounds as the plain prover for a satisfying witness; a witness
    with one wrong entry makes the first round return an error instead of an unverifiable proof"""
    import spartan_parallel_b200 as sp

    X, Q = 1 << 9, 4
    inst = R.synthetic_instance(X)
    secs = R.synthetic_witness(X, [Q], seed=17)
    A, B, Cm = inst.mats
    dinst = sp.R1CSInstance(ctx, 1, X, [X], 2 * X, [A], [B], [Cm])
    big = rand_scalars(64, 170)
    tau_q, tau_x, ch = big[8:10], big[16:25], rand_scalars(11, 171)

    def prover(w_secs):
        dsecs = [sp.ProverWitnessSecInfo(ctx, [Q], [X], np.concatenate(ws.w_mat[0])) for ws in w_secs]
        z = sp.ZMat(ctx, [Q], [X], dsecs)
        sc = sp.sumcheck_phase1(ctx, dinst, z, [Q], Q, [X], X, X, big[:0], tau_q, tau_x)
        sc._hold = (dsecs, z)
        return sc

    plain, checked = prover(secs), prover(secs)
    plain.set_claim(O.ZERO)
    checked.set_claim_checked(O.ZERO)
    for j in range(plain.num_rounds):
        assert np.array_equal(plain.round_eval(), checked.round_eval()), j
        plain.round_bind(ch[j])
        checked.round_bind(ch[j])
    assert np.array_equal(plain.final(), checked.final())
    bad = R.synthetic_witness(X, [Q], seed=17)
    bad[1].w_mat[0][2][5] = O.add(bad[1].w_mat[0][2][5], O.ONE)  # v of proof 2, constraint 5
    sc = prover(bad)
    sc.set_claim_checked(O.ZERO)
    with pytest.raises(sp.SpgError, match="not the sum over the tables"):
        sc.round_eval()


def test_multi_evaluate_bound_rp(ctx):
    """R1CSInstance::multi_evaluate_bound_rp (src/r1csinstance.rs:597-629) for 3 instances (padded to 4)"""
    import spartan_parallel_b200 as sp

    inst = random_instance(3, [32, 16, 32], 3, 16, [16, 8, 16], nnz=70, seed=33)
    A = [inst.mats[3 * i] for i in range(3)]
    B = [inst.mats[3 * i + 1] for i in range(3)]
    Cm = [inst.mats[3 * i + 2] for i in range(3)]
    d = sp.R1CSInstance(ctx, 3, 32, [32, 16, 32], inst.num_vars, A, B, Cm)
    rx, ry, rp = rand_scalars(5, 1), rand_scalars(log2(inst.num_vars), 2), rand_scalars(2, 3)
    ev, bound = d.multi_evaluate_bound_rp(rp, rx, ry)
    trx, try_ = O.eq_evals(rx), O.eq_evals(ry)
    want = [O.sparse_evaluate_with_tables(*inst.mats[m], trx, try_) for m in range(9)]
    assert all(np.array_equal(ev[m], want[m]) for m in range(9))
    for k in range(3):
        col = np.stack([want[3 * i + k] for i in range(3)] + [O.ZERO])
        assert np.array_equal(bound[k], O.dense_evaluate(col, rp)), k
