"""Oracle restatement of the (pi, D) recurrence of the w3 witness sections
(src/lib.rs:1378-1400, 862-880, 1533-1570) against an independent big-integer model.
The reference records no witness values: parity unpinned by the reference, pinned here by
the closed form (pi[q] = product of x over the valid suffix when every later v is 1)."""
import numpy as np

from oracle import cbind as O
from tests.helpers import rand_scalars

Q = (1 << 252) + 27742317777372353535851937790883648493


def model(v, x, seg_len):
    """plain integers mod q, the loop of lib.rs:1382-1399"""
    n = len(v)
    D, pi = [0] * n, [0] * n
    start = 0
    for L in seg_len:
        for k in reversed(range(L)):
            q = start + k
            D[q] = x[q] if k == L - 1 else x[q] * (pi[q + 1] + 1 - v[q + 1]) % Q
            pi[q] = v[q] * D[q] % Q
        start += L
    return pi, D


def table(v, x):
    n = len(v)
    w3 = np.zeros((n, 8, 4), dtype=np.uint64)
    w3[:, 0] = np.stack([O.from_int(a) for a in v])
    w3[:, 1] = np.stack([O.from_int(a) for a in x])
    return w3.reshape(n * 8, 4)


def test_perm_fill_matches_integer_model():
    rng = np.random.default_rng(11)
    for seg_len in ([1], [2], [7], [3, 1, 5], [64, 16, 16, 4, 1]):
        n = sum(seg_len)
        v = [int(b) for b in rng.integers(0, 2, size=n)]
        x = [O.to_int(s) for s in rand_scalars(n, 100 + n)]
        out = O.perm_fill(table(v, x), seg_len).reshape(n, 8, 4)
        pi, D = model(v, x, seg_len)
        assert [O.to_int(s) for s in out[:, 2]] == pi
        assert [O.to_int(s) for s in out[:, 3]] == D


def test_closed_form_all_valid():
    n = 9
    x = [O.to_int(s) for s in rand_scalars(n, 5)]
    out = O.perm_fill(table([1] * n, x), [n]).reshape(n, 8, 4)
    acc = 1
    for q in reversed(range(n)):
        acc = acc * x[q] % Q
        assert O.to_int(out[q, 2]) == acc == O.to_int(out[q, 3])


def test_general_field_valued_v():
    n = 6
    v = [O.to_int(s) for s in rand_scalars(n, 8)]
    x = [O.to_int(s) for s in rand_scalars(n, 9)]
    out = O.perm_fill(table(v, x), [4, 2]).reshape(n, 8, 4)
    pi, D = model(v, x, [4, 2])
    assert [O.to_int(s) for s in out[:, 2]] == pi and [O.to_int(s) for s in out[:, 3]] == D


def test_affine_composition_equals_the_recurrence():
    """the algebra spg_perm_scan relies on: D[q] = A_q D[q+1] + B_q with A_q = x[q] v[q+1],
    B_q = x[q] - A_q, composed right to left, reproduces the sequential loop"""
    n = 13
    v = [O.to_int(s) for s in rand_scalars(n, 21)]
    x = [O.to_int(s) for s in rand_scalars(n, 22)]
    pi, D = model(v, x, [n])
    A, B = 1, 0  # composition of the maps of rows q..n-1, applied to an arbitrary start value
    for q in reversed(range(n)):
        a_q = 0 if q == n - 1 else x[q] * v[q + 1] % Q
        b_q = (x[q] - a_q) % Q
        A, B = a_q * A % Q, (a_q * B + b_q) % Q
        assert A == 0 and B == D[q]  # the last row has A = 0, so every suffix map is constant


def test_derived_sections_satisfy_their_relations():
    """oracle restatement of block_w2 / block_w3 (src/lib.rs:1511-1613): the relations the block
    R1CS checks on those columns -- x = v (tau - sum w2[3..] - in[2]), PMC chain, pi = v D,
    D[q] = x[q] (pi[q+1] + 1 - v[q+1]) -- recomputed independently with python integers"""
    from tests.helpers import rand_scalars

    q_mod = (1 << 252) + 27742317777372353535851937790883648493
    rows, n, phy, vir = 6, 3, 2, 1
    tau, r = rand_scalars(2, 400)
    io_width, vars_width, w2_width = 2 * n, 16, 16
    w0 = O.wit_perm_w0(tau, r, 2 * n, 8)
    vars_ = rand_scalars(rows * vars_width, 401).reshape(rows, vars_width, 4)
    for q in range(rows):
        vars_[q, 0] = O.ONE if q < 4 else 0
    w2, w3 = O.wit_block(vars_, w0, tau, r, n, io_width, phy, vir, w2_width)
    I = lambda a: O.to_int(a)
    t, rr = I(tau), I(r)
    assert [I(x) for x in w0[:6]] == [t] + [pow(rr, k, q_mod) for k in range(1, 6)]
    for q in range(rows):
        v = I(vars_[q, 0])
        s = sum(I(w2[q, i]) for i in range(3, 2 * n)) % q_mod
        assert I(w3[q, 1]) == v * (t - s - I(vars_[q, 2])) % q_mod
        chain = v
        for i in range(phy):
            pmr = rr * I(vars_[q, io_width + 2 * i + 1]) % q_mod
            chain = chain * (t - I(vars_[q, io_width + 2 * i]) - pmr) % q_mod
            assert I(w2[q, 2 * n + 2 * i]) == pmr and I(w2[q, 2 * n + 2 * i + 1]) == chain
        for (pi_c, d_c, x) in ((2, 3, I(w3[q, 1])), (4, 5, chain)):
            nxt = (I(w3[q + 1, pi_c]) + 1 - I(w3[q + 1, 0])) % q_mod if q + 1 < rows else 1
            assert I(w3[q, d_c]) == x * nxt % q_mod and I(w3[q, pi_c]) == v * I(w3[q, d_c]) % q_mod
    sh = O.wit_shift(w3)
    assert np.array_equal(sh[:-1], w3[1:]) and not sh[-1].any()
