"""GPU parity for the permutation-product suffix scan (spg_perm_scan) against the oracle's
sequential loop (src/lib.rs:1378-1400, 862-880, 1533-1570). Bit-exact, through the C ABI."""
import numpy as np
import pytest

from oracle import cbind as O
from tests.helpers import rand_scalars

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import spartan_parallel_b200 as sp

    return sp.Context(0)


def make_table(n, seed, bits=True, width=8):
    rng = np.random.default_rng(seed)
    w3 = rand_scalars(n * width, seed).reshape(n, width, 4)  # junk in the output columns: must be overwritten
    if bits:
        one, zero = O.ONE, np.zeros(4, dtype=np.uint64)
        w3[:, 0] = np.where(rng.integers(0, 2, size=(n, 1)).astype(bool), one, zero)
    return w3.reshape(n * width, 4)


@pytest.mark.parametrize("seg_len", [[1], [2], [5], [1024], [1025], [1023, 1], [3000, 7, 1, 1100], [64, 16, 16, 4, 1],
                                     [1 << 15, 1 << 15, 1 << 14, 4097]])
@pytest.mark.parametrize("bits", [True, False])
def test_perm_scan_matches_oracle(ctx, seg_len, bits):
    import spartan_parallel_b200 as sp
    from spartan_parallel_b200 import api

    n = sum(seg_len)
    w3 = make_table(n, 7 + n, bits)
    want = O.perm_fill(w3, seg_len)
    d = sp.DensePolynomial.new(ctx, w3)
    api.perm_scan(ctx, d, seg_len)
    got = d.to_host()[: n * 8]
    assert np.array_equal(got, want)


def test_second_pair_of_columns(ctx):
    """the PHY / VIR pairs of block_w3 use (pi, D) = columns (4, 5) and (6, 7) with their own x"""
    import spartan_parallel_b200 as sp
    from spartan_parallel_b200 import api

    seg_len = [300, 20, 1]
    n = sum(seg_len)
    w3 = make_table(n, 99)
    want = O.perm_fill(w3, seg_len, v_col=0, x_col=6, pi_col=4, d_col=5)
    d = sp.DensePolynomial.new(ctx, w3)
    api.perm_scan(ctx, d, seg_len, x_col=6, pi_col=4, d_col=5)
    assert np.array_equal(d.to_host()[: n * 8], want)


def test_full_size_closed_form(ctx):
    """2^20 rows (one segment), every v = 1: pi[q] is the suffix product of x; check it through
    size-independent properties: pi[q] = x[q] * pi[q+1] everywhere and pi == D."""
    import spartan_parallel_b200 as sp
    from spartan_parallel_b200 import api

    n = 1 << 20
    x = rand_scalars(n, 4242)
    w3 = np.zeros((n, 8, 4), dtype=np.uint64)
    w3[:, 0] = O.ONE
    w3[:, 1] = x
    d = sp.DensePolynomial.new(ctx, w3.reshape(n * 8, 4))
    api.perm_scan(ctx, d, [n])
    out = d.to_host().reshape(n, 8, 4)
    pi, D = np.ascontiguousarray(out[:, 2]), np.ascontiguousarray(out[:, 3])
    assert np.array_equal(pi, D)
    assert np.array_equal(pi[-1], x[-1])
    assert np.array_equal(pi[:-1], O.vec_mul(x[:-1], pi[1:]))


def test_bad_segments_rejected(ctx):
    import spartan_parallel_b200 as sp
    from spartan_parallel_b200 import api

    d = sp.DensePolynomial.new(ctx, make_table(8, 1))
    with pytest.raises(sp.SpgError):
        api.perm_scan(ctx, d, [4, 0, 4])
    with pytest.raises(sp.SpgError):  # D would overwrite the x column it is computed from
        api.perm_scan(ctx, d, [8], x_col=1, d_col=1)
