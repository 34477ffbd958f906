"""The algebra behind the many-row commitment path (csrc/msm.cu: k_hrecode, k_msm_hrows, k_hfinish)
checked on the CPU with python integers and the oracle's ristretto255:

* the carry-free signed recoding: with K = sum_{w < wins-1} 2^(c w + c - 1), the unsigned c-bit digits
  u_w of s + K give d_w = u_w - 2^(c-1) for w < wins - 1 and the top digit unchanged, and
  sum_w d_w 2^(c w) = s with |d_w| <= 2^(c-1) and 0 <= d_top <= 2^(c-1) -- for every table width the
  library picks and every scalar below q (wins = ceil(254 / c));
* one window table per base plus a Horner chain over the windows gives the same group element as the
  plain multiscalar multiplication (src/commitments.rs:87-92): sum_j s_j G_j = sum_w 2^(c w) S_w with
  S_w = sum_j d_{j,w} G_j;
* the width rule for a table that is built inside the call that uses it (pick_hwindow / pick_window)."""
import random

import pytest

from oracle import ristretto as G

Q = (1 << 252) + 27742317777372353535851937790883648493


def recode(s: int, c: int):
    """k_hrecode restated: digits of s (0 <= s < q) for window width c"""
    wins = -(-254 // c)
    K = sum(1 << (c * w + c - 1) for w in range(wins - 1))
    t = s + K
    assert t < 1 << 256                      # the device adds K in eight 32-bit limbs without a carry out
    half = 1 << (c - 1)
    digits = []
    for w in range(wins - 1):
        digits.append((t & ((1 << c) - 1)) - half)
        t >>= c
    digits.append(t)                         # what is left: the top digit, never negative
    return digits


def edge_scalars(c: int):
    wins = -(-254 // c)
    half = 1 << (c - 1)
    yield from (0, 1, 2, Q - 1, Q - 2, Q // 2, (Q + 1) // 2, 1 << 252, (1 << 252) - 1, (1 << 253) % Q, half, half - 1, half + 1,
                (1 << c) - 1, 1 << c)
    for digit in (half, half - 1, half + 1, (1 << c) - 1, 1):
        yield sum(digit << (c * w) for w in range(wins)) % Q


@pytest.mark.parametrize("c", list(range(5, 18)))
def test_recoding_is_exact_and_bounded(c):
    rnd = random.Random(100 + c)
    wins = -(-254 // c)
    half = 1 << (c - 1)
    for s in list(edge_scalars(c)) + [rnd.randrange(Q) for _ in range(300)]:
        d = recode(s, c)
        assert len(d) == wins
        assert sum(x << (c * w) for w, x in enumerate(d)) == s
        assert all(-half <= x <= half - 1 for x in d[:-1])      # table entries 1 .. 2^(c-1) cover every magnitude
        assert 0 <= d[-1] <= half                               # c * wins >= 254 leaves room for the carry of K
    assert recode(0, c) == [0] * wins                           # a zero scalar costs no addition


@pytest.mark.parametrize("c", [5, 13, 17])
def test_single_window_table_plus_horner_equals_msm(c):
    rnd = random.Random(7 + c)
    n = 3
    gens = G.MultiCommitGens(n, b"horner-test")
    scalars = [rnd.randrange(Q), Q - 1, 5]
    want = G.multiscalar_mul(scalars, gens.G)
    digits = [recode(s, c) for s in scalars]
    wins = len(digits[0])
    # S_w from the single table H[j][d - 1] = d G_j (negative digits: the negated entry)
    S = []
    for w in range(wins):
        acc = G.Point.identity()
        for j in range(n):
            d = digits[j][w]
            if d:
                e = gens.G[j].mul(abs(d))
                acc = acc + (-e if d < 0 else e)
        S.append(acc)
    # k_hfinish: top window first, c doublings per step
    acc = S[wins - 1]
    for w in range(wins - 2, -1, -1):
        for _ in range(c):
            acc = acc.double()
        acc = acc + S[w]
    assert acc.compress() == want.compress()


def test_lazily_built_tables_pay_for_themselves():
    """pick_hwindow / pick_window for a table built inside the call: the width that minimises build + use
    (43 products per entry built, 7 per addition); a table built ahead takes the widest window instead"""
    def lazy_single(L):
        return min(range(9, 18), key=lambda c: 43 * (1 << (c - 1)) + 7 * L * -(-254 // c))

    def lazy_per_window(L):
        return min(range(8, 14), key=lambda c: (43 * (1 << (c - 1)) + 7 * L) * -(-254 // c))

    assert lazy_single(8192) in (13, 14) and lazy_single(4096) in (12, 13) and lazy_single(128) == 9
    assert lazy_per_window(2) == 8 and lazy_per_window(1) == 8       # a bullet round's two rows: the narrowest table
    assert lazy_per_window(1 << 16) >= 12
    # the widest table only pays after several commitments of that size
    L, c_wide, c_lazy = 8192, 17, lazy_single(8192)
    build = lambda c: 43 * (1 << (c - 1))
    use = lambda c: 7 * L * -(-254 // c)
    n_break_even = (build(c_wide) - build(c_lazy)) / (use(c_lazy) - use(c_wide))
    assert 5 < n_break_even < 15
