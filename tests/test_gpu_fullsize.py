"""Full-size runs at the BASELINE.json config sizes on one GPU.

Two kinds of check, side by side:
  * bit-exact comparison of every round polynomial and every final claim with the oracle
    (OpenMP restatement of the reference loops) at C2 (2^20 x 1), C3 (2^16 x 256), full-size C4
    (P = 5, Q_p = {64,16,16,4,1}, W = 5 with a single section) and C5 (2^20 x 64, marked slow:
    ~30 s of oracle time on 16 cores);
  * size-independent properties of the same runs:
    - sumcheck verifier relation: with claim_0 = sum over the cube, every round polynomial
      interpolated from (e0, claim - e0, e2, e3) must hand the next round its claim, and the
      last claim must equal eq_claim * (Az*Bz - Cz) resp. eq * ABC * Z of the final claims;
    - a satisfying witness makes the phase-1 claim vanish (e0 = e1 = 0 in round 0);
    - linearity in the witness: phase-2 claims are checked against the phase-1 ones through
      r_A*Az + r_B*Bz + r_C*Cz = claim_phase2."""
import numpy as np
import pytest

from oracle import cbind as O

pytestmark = pytest.mark.gpu

ONE = O.ONE


@pytest.fixture(scope="module")
def ctx():
    import spartan_parallel_b200 as sp

    return sp.Context(0)


def canonical(rng, n):
    a = rng.integers(0, 1 << 64, size=(n, 4), dtype=np.uint64)
    a[:, 3] &= np.uint64((1 << 60) - 1)
    return a


def run_shape(ctx, log_x, Q, supply_claim=False):
    import spartan_parallel_b200 as sp

    X, N = 1 << log_x, (1 << log_x) * Q
    nq = Q.bit_length() - 1
    rng = np.random.default_rng(log_x * 1000 + Q)
    u = canonical(rng, N)
    du = sp.DensePolynomial.new(ctx, u)
    dun = sp.DensePolynomial.new(ctx, np.ascontiguousarray(np.roll(u.reshape(Q, X, 4), -1, axis=1).reshape(N, 4)))
    v = sp.vec_op(ctx, "mul", du, dun).to_host()
    del du, dun
    rows = np.arange(X, dtype=np.uint32)
    ones = np.tile(ONE, (X, 1))
    inst = sp.R1CSInstance(ctx, 1, X, [X], 2 * X, [(rows, rows, ones)], [(rows, ((rows + 1) % X).astype(np.uint32), ones)],
                           [(rows, (rows + X).astype(np.uint32), ones)])
    secs = [sp.ProverWitnessSecInfo(ctx, [Q], [X], u), sp.ProverWitnessSecInfo(ctx, [Q], [X], v)]
    z = sp.ZMat(ctx, [Q], [X], secs)
    tau_q, tau_x = canonical(rng, max(nq, 1))[:nq], canonical(rng, log_x)
    ch1, ch2 = canonical(rng, log_x + nq), canonical(rng, 1 + log_x)
    r_abc = canonical(rng, 3)
    e = np.zeros((0, 4), dtype=np.uint64)

    sc1 = sp.sumcheck_phase1(ctx, inst, z, [Q], Q, [X], X, X, e, tau_q, tau_x)
    claim = O.ZERO  # satisfied instance
    if supply_claim:
        sc1.set_claim(claim)
    for j in range(sc1.num_rounds):
        e0, e2, e3 = sc1.round_eval()
        if j == 0:
            assert np.array_equal(e0, O.ZERO), "round 0: e(0) must vanish for a satisfying witness"
        co = O.unipoly_from_evals(np.stack([e0, O.sub(claim, e0), e2, e3]))
        assert np.array_equal(O.unipoly_evaluate(co, O.from_u64(2)), e2)
        claim = O.unipoly_evaluate(co, ch1[j])
        sc1.round_bind(ch1[j])
    tau_claim, az, bz, cz = sc1.final()
    assert np.array_equal(claim, O.mul(tau_claim, O.sub(O.mul(az, bz), cz))), "phase-1 final check"
    want_tau = O.mul(O.eq_evaluate(tau_x, ch1[:log_x]), O.eq_evaluate(tau_q, ch1[log_x:]) if nq else ONE)
    assert np.array_equal(tau_claim, want_tau)

    rx = ch1[:log_x][::-1].copy()
    sc2 = sp.SumcheckPhase2(ctx, inst, z, [Q], Q, [X], X, 2, rx, ch1[log_x:], e, *r_abc)
    claim = O.add(O.add(O.mul(r_abc[0], az), O.mul(r_abc[1], bz)), O.mul(r_abc[2], cz))
    for j in range(sc2.num_rounds):
        e0, e2, e3 = sc2.round_eval()
        co = O.unipoly_from_evals(np.stack([e0, O.sub(claim, e0), e2, e3]))
        claim = O.unipoly_evaluate(co, ch2[j])
        sc2.round_bind(ch2[j])
    a, b, c = sc2.final()
    assert np.array_equal(claim, O.mul(O.mul(a, b), c)), "phase-2 final check"
    # the Z claim is the MLE of (u | v) at (rq, rw, ry): check against the witness evaluations
    ry = ch2[:log_x][::-1].copy()
    rw = ch2[log_x]
    rq = ch1[log_x:][::-1].copy()
    r = np.concatenate([rq, ry]) if nq else ry
    eu, ev = secs[0].poly_w(0).evaluate(r), secs[1].poly_w(0).evaluate(r)
    assert np.array_equal(c, O.add(O.mul(O.sub(ONE, rw), eu), O.mul(rw, ev))), "Z claim vs witness openings"
    sc1.free(); sc2.free(); z.free()
    for s in secs:
        s.free()


def test_c2_single_instance_2_20(ctx):
    run_shape(ctx, 20, 1)


def test_c3_2_16_x_256(ctx):
    run_shape(ctx, 16, 256)


def test_c5_2_20_x_64(ctx):
    run_shape(ctx, 20, 64)


def test_c5_2_20_x_64_supplied_claim(ctx):
    """same shape with spg_sc1_set_claim(0): two-point first round (what bench.py times)"""
    run_shape(ctx, 20, 64, supply_claim=True)


# ----------------------------------------------------------------------------------------------
# Bit-exact comparison with the oracle (OpenMP restatement of the reference loops) at the sizes
# BASELINE.json names: every round polynomial of both sumchecks and all final claims.
def test_c2_oracle_every_round(ctx):
    """C2: one instance, X = 2^20 constraints, one proof"""
    from oracle import r1cs as R
    from tests.test_gpu_r1cs import run_case

    X = 1 << 20
    run_case(ctx, R.synthetic_instance(X), 1, [1], [X], X, R.synthetic_witness(X, [1], seed=202), seed=2020, claim=O.ZERO)


def test_c3_oracle_every_round(ctx):
    """C3: X = 2^16 constraints x 256 proofs"""
    from oracle import r1cs as R
    from tests.test_gpu_r1cs import run_case

    X, Q = 1 << 16, 256
    run_case(ctx, R.synthetic_instance(X), 1, [Q], [X], X, R.synthetic_witness(X, [Q], seed=203), seed=2030)


def test_c4_full_size_oracle_every_round(ctx):
    """C4 as SURVEY 8(d) states it: P = 5 instances of 2^12 constraints, Q_p = {64,16,16,4,1},
    W = 5 sections with section 1 single (perm_w0, src/lib.rs:1703), Y_p in {2^12, 2^11},
    random sparse matrices with unit, zero and general coefficients"""
    from tests.helpers import random_instance, random_witness_secs
    from tests.test_gpu_r1cs import run_case

    P, W, Ymax = 5, 5, 1 << 12
    num_proofs = [64, 16, 16, 4, 1]
    num_cons = [1 << 12] * P
    Y = [1 << 12, 1 << 12, 1 << 11, 1 << 12, 1 << 11]
    inst = random_instance(P, num_cons, W, Ymax, Y, nnz=3 << 12, seed=204)
    kinds = ["full", "single", "full", "full", "full"]
    sec_inputs = [Y, [1 << 10] * P, [1 << 12, 1 << 13, 1 << 11, 1 << 12, 1 << 10], [8] * P, [8] * P]
    secs = random_witness_secs(P, num_proofs, W, sec_inputs, kinds, seed=205)
    run_case(ctx, inst, P, num_proofs, Y, Ymax, secs, seed=2040)


@pytest.mark.slow
def test_c5_oracle_every_round(ctx):
    """C5: X = 2^20 constraints x 64 proofs (2^26 constraints); about a minute of oracle time on
    16 host cores. SPG_SKIP_SLOW=1 skips it."""
    import os

    if os.environ.get("SPG_SKIP_SLOW"):
        pytest.skip("SPG_SKIP_SLOW set")
    from oracle import r1cs as R
    from tests.test_gpu_r1cs import run_case

    X, Q = 1 << 20, 64
    run_case(ctx, R.synthetic_instance(X), 1, [Q], [X], X, R.synthetic_witness(X, [Q], seed=205), seed=2050, claim=O.ZERO)
