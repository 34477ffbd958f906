"""End-to-end parity of the satisfiability proof: the C++ host mirror (libspghost.so)
driving the CUDA kernels through the C ABI must emit, byte for byte, the proof the oracle's
restatement of R1CSProof::prove emits for the same inputs, transcript label and tape seed
-- and the oracle's restatement of R1CSProof::verify must accept it."""
import os

import numpy as np
import pytest

from oracle import cbind as O
from oracle import protocol as Pr
from oracle import r1cs as R
from tests.helpers import rand_scalars
from tests.test_oracle_protocol import sparse_evals

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import spartan_parallel_b200 as sp

    return sp.Context(0)


def device_setup(ctx, inst, secs):
    import spartan_parallel_b200 as sp

    n = inst.num_instances
    dinst = sp.R1CSInstance(ctx, n, inst.max_num_cons, inst.num_cons, inst.num_vars, [inst.mats[3 * i] for i in range(n)],
                            [inst.mats[3 * i + 1] for i in range(n)], [inst.mats[3 * i + 2] for i in range(n)])
    dsecs = []
    for ws in secs:
        nq = [len(ws.w_mat[p]) for p in range(len(ws.w_mat))]
        flat = np.concatenate([np.concatenate(ws.w_mat[p]) for p in range(len(ws.w_mat))])
        dsecs.append(sp.ProverWitnessSecInfo(ctx, nq, ws.num_inputs[: len(nq)], flat))
    return dinst, dsecs


def run_case(ctx, inst, P, num_proofs, num_inputs, max_y, secs, seed, verify=True, device_gens=False):
    import spartan_parallel_b200 as sp
    from spartan_parallel_b200 import host

    max_q = max(num_proofs)
    gnv = max(q * y for q, y in zip(num_proofs, num_inputs))
    seed_scalar = rand_scalars(1, seed)[0]
    gens = Pr.R1CSGens(b"gens_r1cs_sat", gnv)
    trace = {}
    want, ch = Pr.r1cs_prove(inst, P, max_q, num_proofs, max_y, num_inputs, secs, gens, Pr.Transcript(b"spgpu-parity"),
                             Pr.RandomTape(b"proof", seed_scalar), trace=trace)
    want_bytes = Pr.serialize_r1cs_proof(want)
    dinst, dsecs = device_setup(ctx, inst, secs)
    dgens = host.R1CSGens(ctx, b"gens_r1cs_sat", gnv) if device_gens else None
    got_bytes, got_ch = host.r1cs_prove(ctx, dinst, dsecs, num_proofs, max_q, num_inputs, max_y, b"spgpu-parity", b"gens_r1cs_sat",
                                        seed_scalar, gnv, dgens)
    assert len(got_bytes) == len(want_bytes)
    if got_bytes != want_bytes:
        first = next(i for i in range(len(want_bytes)) if got_bytes[i] != want_bytes[i])
        raise AssertionError(f"proof bytes differ from offset {first} of {len(want_bytes)}")
    for a, b in zip(got_ch, ch):
        assert len(a) == len(b) and all(np.array_equal(x, y) for x, y in zip(a, b))
    if verify:
        # witness commitments from the device MSM (they are prover outputs in the protocol)
        dg = sp.MultiCommitGens(ctx, gens.pc.gens_n.compressed())
        comms = [[dg.commit_poly(w.poly_w(p)) for p in range(len(w.num_proofs))] for w in dsecs]
        wnp = [list(w.num_proofs) for w in dsecs]
        wni = [list(w.num_inputs) for w in dsecs]
        evals = sparse_evals(inst, ch[2], ch[3], ch[0], P)
        out = Pr.r1cs_verify(Pr.deserialize_r1cs_proof(got_bytes), P, max_q, num_proofs, max_y, wnp, wni, comms, inst.max_num_cons,
                             gens, evals, Pr.Transcript(b"spgpu-parity"))
        assert out is not None, "the oracle's verifier rejected the device proof"
        # and the device commitments equal the oracle's on one polynomial
        assert comms[0][0] == Pr.poly_commit(secs[0].poly_w(0), gens.pc.gens_n)
    return got_bytes


def test_small_single_instance(ctx):
    X, Q = 1 << 5, 2
    inst = R.synthetic_instance(X, unit=False, seed=1)
    secs = R.synthetic_witness(X, [Q], seed=2)
    blob = run_case(ctx, inst, 1, [Q], [X], X, secs, seed=3)
    golden = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "r1cs_proof_x32_q2.bin")
    assert blob == open(golden, "rb").read(), "device proof differs from the committed golden fixture"


def test_three_instances_ragged(ctx):
    P, X = 3, 1 << 3
    inst = R.synthetic_instance(X, num_instances=P, unit=False, seed=5)
    secs = R.synthetic_witness(X, [4, 2, 1], seed=6)
    run_case(ctx, inst, P, [4, 2, 1], [X] * P, X, secs, seed=7)


def test_c1_baseline_config(ctx):
    """BASELINE config C1: 2^10 constraints x 4 proofs: full proof parity."""
    X, Q = 1 << 10, 4
    inst = R.synthetic_instance(X)
    secs = R.synthetic_witness(X, [Q], seed=8)
    blob = run_case(ctx, inst, 1, [Q], [X], X, secs, seed=9, verify=False)
    assert len(blob) > 8000


def test_c1_device_openings(ctx):
    """same proof with the opening-proof MSMs (Cx, every bullet-reduction L / R, the folded
    generator) served by the device from unfolded fixed-base tables: identical bytes"""
    X, Q = 1 << 10, 4
    inst = R.synthetic_instance(X)
    secs = R.synthetic_witness(X, [Q], seed=8)
    run_case(ctx, inst, 1, [Q], [X], X, secs, seed=9, verify=False, device_gens=True)


def test_ragged_device_openings(ctx):
    P, X = 3, 1 << 6
    inst = R.synthetic_instance(X, num_instances=P, unit=False, seed=5)
    secs = R.synthetic_witness(X, [8, 4, 1], seed=6)
    run_case(ctx, inst, P, [8, 4, 1], [X] * P, X, secs, seed=7, device_gens=True)


@pytest.mark.slow
def test_wide_msm_device_openings(ctx):
    """2^16-entry witness polynomials: openings of size 256, the few-rows / many-bases MSM kernel"""
    X, Q = 1 << 13, 8
    inst = R.synthetic_instance(X)
    secs = R.synthetic_witness(X, [Q], seed=18)
    run_case(ctx, inst, 1, [Q], [X], X, secs, seed=19, verify=False, device_gens=True)
