"""GPU parity of the y-sharded phase 2 (spg_sc2_create_slice + spg_sc2_host_tail_*), on one GPU: G
slice provers stand in for G ranks, their partial round evaluations are added on the host, and the
cross-rank rounds run in parallel.HostTail2. Every round polynomial and the final claims must equal
the unsharded device prover's AND the oracle's (oracle.r1cs.prove_tables), bit for bit."""
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


def run_slices(ctx, sp, parallel, dinst, zrq, Y, W, G, rx, r_abc, ch2):
    n = W * Y // G
    eng = [sp.SumcheckPhase2.slice(ctx, dinst, zrq, Y, W, r * n, n, rx, *r_abc) for r in range(G)]
    evals = []
    n_local = log2(n)
    for j in range(n_local):
        evals.append(sp.host_sum(np.stack([e.round_eval() for e in eng])))
        for e in eng:
            e.round_bind(ch2[j])
    fin = np.stack([e.final() for e in eng])
    for e in eng:
        e.free()
    tail = parallel.HostTail2(fin[:, 1].copy(), fin[:, 2].copy(), log2(Y) - n_local)
    tail.set_scale(fin[0, 0])
    for j in range(n_local, log2(W * Y)):
        evals.append(tail.round_eval())
        tail.round_bind(ch2[j])
    return evals, tail.final()


@pytest.mark.parametrize("log_y,W,G,kinds", [(10, 2, 2, ("full", "full")), (10, 2, 8, ("full", "full")), (8, 4, 4, ("full", "short", "full", "full")),
                                             (6, 1, 64, ("full",)), (12, 2, 4, ("full", "full")), (5, 2, 64, ("full", "full"))])
def test_slices_match_unsharded_and_oracle(ctx, log_y, W, G, kinds):
    import spartan_parallel_b200 as sp
    from spartan_parallel_b200 import parallel

    Y, Q, X = 1 << log_y, 4, 1 << 7
    inst = random_instance(1, [X], W, Y, [Y], 6 * X, seed=40 + log_y)
    sec_inputs = [[Y if k == "full" else max(Y // 4, 1)] for k in kinds]
    secs = random_witness_secs(1, [Q], W, sec_inputs, kinds, seed=50 + log_y)
    nx, nq, ny, nw = log2(X), log2(Q), log_y, log2(W)
    tau_q, tau_x = rand_scalars(nq, 1), rand_scalars(nx, 2)
    ch1, ch2, r_abc = rand_scalars(nx + nq, 3), rand_scalars(ny + nw, 4), rand_scalars(3, 5)
    want = R.prove_tables(inst, 1, Q, [Q], Y, [Y], secs, tau_q[:0], tau_q, tau_x, ch1, r_abc, ch2)
    A, B, Cm = inst.mats
    dinst = sp.R1CSInstance(ctx, 1, X, [X], inst.num_vars, [A], [B], [Cm])
    dsecs = []
    for ws in secs:
        flat = np.concatenate([np.concatenate(ws.w_mat[p]) for p in range(len(ws.w_mat))])
        dsecs.append(sp.ProverWitnessSecInfo(ctx, [len(ws.w_mat[0])], ws.num_inputs, flat))
    z = sp.ZMat(ctx, [Q], [Y], dsecs)
    rx, rq_rev = ch1[:nx][::-1].copy(), ch1[nx:]
    zrq = sp.zmat_bind_rq(ctx, z, rq_rev)
    evals, claims = run_slices(ctx, sp, parallel, dinst, zrq, Y, W, G, rx, r_abc, ch2)
    assert len(evals) == len(want.evals2)
    for j, (g, w) in enumerate(zip(evals, want.evals2)):
        assert np.array_equal(g, w), f"phase 2 round {j} (G = {G})"
    assert np.array_equal(claims, want.claims2)
    # and the unsharded device prover on the same bound table
    sc2 = sp.SumcheckPhase2.from_zrq(ctx, dinst, zrq, [Y], Y, W, rx, ch1[:0], *r_abc)
    assert np.array_equal(sc2.run_rounds(ch2), np.stack(evals))
    assert np.array_equal(sc2.final(), claims)


def test_slice_rejects_bad_chunks(ctx):
    import spartan_parallel_b200 as sp

    X = Y = 1 << 6
    inst = R.synthetic_instance(X)
    A, B, Cm = inst.mats
    dinst = sp.R1CSInstance(ctx, 1, X, [X], inst.num_vars, [A], [B], [Cm])
    zrq = sp.DensePolynomial.new(ctx, rand_scalars(2 * Y, 9))
    rx, r_abc = rand_scalars(6, 1), rand_scalars(3, 2)
    for off, n in ((0, 2 * Y), (8, 16), (0, 24), (2 * Y, 16)):  # wider than a section, misaligned, not a power of two, outside
        with pytest.raises(sp.SpgError):
            sp.SumcheckPhase2.slice(ctx, dinst, zrq, Y, 2, off, n, rx, *r_abc)
    with pytest.raises(sp.SpgError):  # shorter table than W * Y
        sp.SumcheckPhase2.slice(ctx, dinst, sp.DensePolynomial.new(ctx, rand_scalars(Y, 9)), Y, 2, 0, 16, rx, *r_abc)
