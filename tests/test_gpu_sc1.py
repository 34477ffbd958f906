"""GPU parity for the phase-1 sumcheck loops (spg_sc1_*) against the oracle's
restatement of sumcheck.rs:1067-1380. Every round polynomial evaluation and every
final claim must be bit-identical."""
import numpy as np
import pytest

from oracle import cbind as O
from tests.helpers import drive_sc1_oracle, log2, rand_scalars

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import spartan_parallel_b200 as sp

    return sp.Context(0)


CASES = [
    # (num_proofs, num_cons)
    ([1], [1]),
    ([1], [2]),
    ([1], [16]),
    ([4], [1]),
    ([4], [1024]),          # BASELINE config C1 shape
    ([2, 1], [8, 8]),
    ([8, 4, 4, 1, 1], [64, 32, 64, 2, 1]),   # ragged in q and x, P = 5 -> P' = 8
    ([64, 16, 16, 4, 1], [256, 256, 128, 256, 64]),  # C4-like
    ([1, 1, 1], [4, 4, 4]),
]


@pytest.mark.parametrize("num_proofs,num_cons", CASES)
@pytest.mark.parametrize("fuse", [True])
def test_sc1_rounds_bit_exact(ctx, num_proofs, num_cons, fuse):
    import spartan_parallel_b200 as sp

    P = len(num_proofs)
    max_q, max_x = max(num_proofs), max(num_cons)
    Pp = 1 if P == 1 else 1 << (P - 1).bit_length()
    N = sum(q * x for q, x in zip(num_proofs, num_cons))
    seed = 1000 + 17 * N + P
    Az, Bz, Cz = rand_scalars(N, seed), rand_scalars(N, seed + 1), rand_scalars(N, seed + 2)
    tau_p, tau_q, tau_x = rand_scalars(max(log2(Pp), 1), seed + 3)[: log2(Pp)], rand_scalars(max(log2(max_q), 1), seed + 4)[: log2(max_q)], rand_scalars(max(log2(max_x), 1), seed + 5)[: log2(max_x)]
    rounds = log2(Pp) + log2(max_q) + log2(max_x)
    ch = rand_scalars(max(rounds, 1), seed + 6)
    want_evals, want_final = drive_sc1_oracle(num_proofs, max_q, num_cons, max_x, Az, Bz, Cz, tau_p, tau_q, tau_x, ch)

    sc = sp.SumcheckPhase1.from_tables(ctx, num_proofs, max_q, num_cons, max_x, Az, Bz, Cz, tau_p, tau_q, tau_x)
    assert sc.num_rounds == rounds
    for j in range(rounds):
        got = sc.round_eval()
        assert np.array_equal(got, want_evals[j]), f"round {j} evals differ"
        sc.round_bind(ch[j])
    assert np.array_equal(sc.final(), want_final)


def test_sc1_state_errors(ctx):
    import spartan_parallel_b200 as sp

    Az = rand_scalars(8, 1)
    tau = rand_scalars(3, 2)
    sc = sp.SumcheckPhase1.from_tables(ctx, [1], 1, [8], 8, Az, Az, Az, tau[:0], tau[:0], tau)
    with pytest.raises(sp.SpgError):
        sc.round_bind(tau[0])  # bind before eval
    sc.round_eval()
    with pytest.raises(sp.SpgError):
        sc.round_eval()  # eval twice
    with pytest.raises(sp.SpgError):
        sc.final()  # not all rounds bound
    with pytest.raises(sp.SpgError):
        sp.SumcheckPhase1.from_tables(ctx, [3], 4, [8], 8, Az, Az, Az, tau[:0], tau[:2], tau)  # non power of two


def test_run_rounds_equals_round_by_round(ctx):
    """spg_sc1_run_rounds (C loop with known challenges) == eval/bind driven round by round"""
    import spartan_parallel_b200 as sp

    X, Q = 1 << 10, 4
    tabs = [rand_scalars(X * Q, s) for s in (21, 22, 23)]
    tau_q, tau_x, ch = rand_scalars(2, 24), rand_scalars(10, 25), rand_scalars(12, 26)
    none = rand_scalars(0, 1)
    a = sp.SumcheckPhase1.from_tables(ctx, [Q], Q, [X], X, *tabs, none, tau_q, tau_x)
    want = []
    for j in range(a.num_rounds):
        want.append(a.round_eval())
        a.round_bind(ch[j])
    b = sp.SumcheckPhase1.from_tables(ctx, [Q], Q, [X], X, *tabs, none, tau_q, tau_x)
    got = b.run_rounds(ch)
    assert np.array_equal(got, np.stack(want))
    assert np.array_equal(a.final(), b.final())


def test_tables_read_between_fused_binds_are_canonical(ctx):
    """The fused bind leaves its tables in [0, 2q) while the next bind is the same kernel (csrc/sc1.cu,
    k_rows_rolled<true>). Anything else that reads them -- here spg_sc1_debug_tables in the middle of the
    run -- must see canonical scalars (k_canon3), and the rounds after such a read must not change."""
    import spartan_parallel_b200 as sp

    X, Q = 1 << 13, 2
    N = X * Q
    Az, Bz, Cz = rand_scalars(N, 71), rand_scalars(N, 72), rand_scalars(N, 73)
    tau_q, tau_x = rand_scalars(1, 74), rand_scalars(13, 75)
    ch = rand_scalars(14, 76)
    none = tau_q[:0]
    ref = sp.SumcheckPhase1.from_tables(ctx, [Q], Q, [X], X, Az, Bz, Cz, none, tau_q, tau_x)
    want = ref.run_rounds(ch)
    want_final = ref.final()
    sc = sp.SumcheckPhase1.from_tables(ctx, [Q], Q, [X], X, Az, Bz, Cz, none, tau_q, tau_x)
    tabs = [t.reshape(Q, X, 4) for t in (Az, Bz, Cz)]
    for j in range(14):
        assert np.array_equal(sc.round_eval(), want[j]), f"round {j}"
        sc.round_bind(ch[j])
        if j < 3:  # after binds 1 and 2 the tables are held unreduced (rows of 2^12, 2^11 >= 2^10)
            got = sc.debug_tables()
            for k in range(3):
                lo, hi = tabs[k][:, 0::2].reshape(-1, 4), tabs[k][:, 1::2].reshape(-1, 4)
                bound = O.vec_add(lo, O.vec_mul(np.tile(ch[j], (lo.shape[0], 1)), O.vec_sub(hi, lo)))
                tabs[k] = bound.reshape(Q, -1, 4)
                assert np.array_equal(got[k], bound), f"table {k} after bind {j}"
    assert np.array_equal(sc.final(), want_final)
