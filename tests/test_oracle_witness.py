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
