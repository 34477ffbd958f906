"""GPU parity for the Pedersen commitment kernels (spg_gens_upload, spg_poly_commit,
spg_commit_batch) against the oracle's python ristretto255: compressed bytes must match."""
import numpy as np
import pytest

from oracle import cbind as O
from oracle import ristretto as G
from tests.helpers import rand_scalars

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import spartan_parallel_b200 as sp

    return sp.Context(0)


@pytest.fixture(scope="module")
def gens16():
    return G.MultiCommitGens(16, b"test-gens")


def test_gens_upload_rejects_bad_point(ctx, gens16):
    import spartan_parallel_b200 as sp

    enc = bytearray(gens16.compressed())
    enc[32:64] = bytes.fromhex("0100000000000000000000000000000000000000000000000000000000000000")  # negative s
    with pytest.raises(sp.SpgError):
        sp.MultiCommitGens(ctx, bytes(enc))


def test_poly_commit_matches_oracle(ctx, gens16):
    import spartan_parallel_b200 as sp

    dg = sp.MultiCommitGens(ctx, gens16.compressed())
    Z = rand_scalars(64, 1)  # 6 variables -> 8 rows x 8 columns
    Z[5] = 0
    Z[8:16] = 0  # an all-zero row commits to the identity
    Z[17] = O.ONE
    got = dg.commit_poly(sp.DensePolynomial.new(ctx, Z))
    assert len(got) == 8
    for i in range(8):
        row = [O.to_int(Z[8 * i + j]) for j in range(8)]
        assert got[i] == G.commit_vec(row, 0, gens16).compress(), i
    assert got[1] == bytes(32)


def test_commit_batch_with_blinds(ctx, gens16):
    import spartan_parallel_b200 as sp

    g4 = G.MultiCommitGens(4, b"test-gens")
    dg = sp.MultiCommitGens(ctx, g4.compressed())
    s = rand_scalars(12, 2).reshape(3, 4, 4)
    blinds = rand_scalars(3, 3)
    blinds[1] = 0
    got = dg.commit_batch(s, blinds)
    for i in range(3):
        want = G.commit_vec([O.to_int(x) for x in s[i]], O.to_int(blinds[i]), g4).compress()
        assert got[i] == want, i
    # scalar commitment (gens_1): v*G0 + b*h  (src/commitments.rs:73-78)
    g1 = G.MultiCommitGens(1, b"test-gens")
    d1 = sp.MultiCommitGens(ctx, g1.compressed())
    v, b = rand_scalars(2, 4)
    assert d1.commit_batch(v.reshape(1, 1, 4), b.reshape(1, 4))[0] == G.commit_scalar(O.to_int(v), O.to_int(b), g1).compress()


def test_hyrax_row_identity_at_scale(ctx):
    """Size-independent property at a size the python oracle cannot reach directly:
    sum_i L_i * C_i == commit(L * Z) -- the relation PolyEvalProof::verify checks
    (src/dense_mlpoly.rs:505-513). 2^14 scalars, 128 rows x 128 columns, chunked path."""
    import spartan_parallel_b200 as sp

    gens = G.MultiCommitGens(128, b"scale-gens")
    dg = sp.MultiCommitGens(ctx, gens.compressed())
    Z = rand_scalars(1 << 14, 5)
    poly = sp.DensePolynomial.new(ctx, Z)
    rows = dg.commit_poly(poly)
    assert len(rows) == 128
    r = rand_scalars(14, 6)
    L = O.eq_evals(r[:7])
    LZ = poly.bound(L).to_host()
    lhs = G.multiscalar_mul([O.to_int(x) for x in L], [G.decompress(c) for c in rows]).compress()
    rhs = dg.commit_batch(LZ.reshape(1, 128, 4))[0]
    assert lhs == rhs


def test_few_rows_many_bases_kernel(ctx):
    """the wide MSM path (<= 16 rows, >= 256 bases): zeros, a sparse row, blinds; bytes vs the oracle"""
    import spartan_parallel_b200 as sp

    n = 300
    gens = G.MultiCommitGens(n, b"wide-gens")
    dg = sp.MultiCommitGens(ctx, gens.compressed())
    s = rand_scalars(3 * n, 7).reshape(3, n, 4)
    s[1, ::2] = 0
    s[2] = 0
    s[2, 299] = O.ONE
    blinds = rand_scalars(3, 8)
    blinds[2] = 0
    got = dg.commit_batch(s, blinds)
    for i in range(3):
        want = G.commit_vec([O.to_int(x) for x in s[i]], O.to_int(blinds[i]), gens).compress()
        assert got[i] == want, i
