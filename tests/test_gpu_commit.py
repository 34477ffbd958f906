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


def test_fe8_selftest(ctx):
    """the eight-limb GF(2^255-19) arithmetic of the MSM kernels (csrc/fe8.cuh) against the
    ten-limb code on random and edge operands: products, sums, differences, point additions"""
    import ctypes as C

    bad = C.c_uint32(0xFFFFFFFF)
    from spartan_parallel_b200._lib import check

    check(ctx.L.spg_debug_fe8_selftest(ctx.h, 1 << 16, 12345, C.byref(bad)), "spg_debug_fe8_selftest")
    assert bad.value == 0, f"failing checks: {bad.value:#x}"


@pytest.mark.parametrize("c", [5, 8, 11, 13, 16])
def test_window_widths(ctx, c, monkeypatch):
    """every table geometry (signed c-bit digits, ceil(254/c) windows) gives the oracle's bytes;
    scalars include q-1, 2^252, digit patterns that carry through every window, and blinds"""
    import spartan_parallel_b200 as sp

    monkeypatch.setenv("SPG_MSM_WINDOW", str(c))
    n = 6
    gens = G.MultiCommitGens(n, b"win-gens")
    dg = sp.MultiCommitGens(ctx, gens.compressed())
    q = O.Q if hasattr(O, "Q") else (1 << 252) + 27742317777372353535851937790883648493
    half = 1 << (c - 1)
    carry_all = sum((half + 1) << (c * w) for w in range(254 // c)) % q   # every digit > half: carries ripple to the top
    exact_half = sum(half << (c * w) for w in range(254 // c)) % q        # every digit == half: no carry
    ints = [[q - 1, 1 << 252, carry_all, exact_half, (1 << 253) % q, 0],
            [1, 2, half, half + 1, (1 << c) - 1, 1 << c]]
    s = np.stack([np.stack([O.from_int(v) for v in row]) for row in ints])
    rnd = rand_scalars(n, 40 + c).reshape(1, n, 4)
    s = np.concatenate([s, rnd])
    blinds = np.stack([O.from_int(q - 1), O.from_int(0), rand_scalars(1, 41)[0]])
    got = dg.commit_batch(s, blinds)
    assert dg.info()["window_bits"] == c
    for i in range(3):
        want = G.commit_vec([O.to_int(x) for x in s[i]], O.to_int(blinds[i]), gens).compress()
        assert got[i] == want, (c, i)
    dg.free()


@pytest.mark.parametrize("c,slab", [(9, None), (13, 50), (16, None), (17, 1)])
def test_many_rows_single_window_table(ctx, gens16, c, slab, monkeypatch):
    """the many-row path (one window table per base, Horner over the windows; csrc/msm.cu k_msm_hrows)
    forced at a size the python oracle reaches: 128 rows x 16 bases, every table width, one and several
    slabs, edge scalars (0, 1, q-1, 2^252, digits at +-half in every window), blinds; bytes against the
    oracle on chosen rows and against the per-window path on all rows"""
    import spartan_parallel_b200 as sp

    q = (1 << 252) + 27742317777372353535851937790883648493
    L, R = 128, 16
    half = 1 << (c - 1)
    nw = -(-254 // c)
    pats = [0, 1, q - 1, 1 << 252, (1 << 253) % q, half, half - 1, half + 1, (1 << c) - 1, 1 << c,
            sum(half << (c * w) for w in range(nw - 1)) % q,          # every digit becomes -half ... with carries
            sum((half - 1) << (c * w) for w in range(nw - 1)) % q,    # every digit just below the sign change
            sum(((1 << c) - 1) << (c * w) for w in range(nw - 1)) % q, q - 2, q // 2, (q + 1) // 2]
    s = rand_scalars(L * R, 70 + c).reshape(L, R, 4)
    s[0] = np.stack([O.from_int(v) for v in pats])
    s[1] = 0                                                          # identity
    s[2] = 0
    s[2, 15] = O.ONE
    s[3, ::2] = 0
    blinds = rand_scalars(L, 71 + c)
    blinds[0] = O.from_int(q - 1)
    blinds[5] = 0
    monkeypatch.setenv("SPG_MSM_HORNER_MIN", str(1 << 60))            # per-window path
    dg0 = sp.MultiCommitGens(ctx, gens16.compressed())
    want_all = dg0.commit_batch(s, blinds)
    assert dg0.info()["rows_table"]["table_bytes"] == 0
    dg0.free()
    monkeypatch.setenv("SPG_MSM_HORNER_MIN", "1")
    monkeypatch.setenv("SPG_MSM_HWINDOW", str(c))
    if slab is not None:
        monkeypatch.setenv("SPG_MSM_SLAB_BYTES", str(slab * nw * R * 4))
    dg = sp.MultiCommitGens(ctx, gens16.compressed())
    got = dg.commit_batch(s, blinds)
    info = dg.info()["rows_table"]
    assert info["window_bits"] == c and info["adds_per_scalar"] == nw and info["table_bases"] == R
    assert got == want_all
    for i in [0, 1, 2, 3, 5, 64, 127]:
        want = G.commit_vec([O.to_int(x) for x in s[i]], O.to_int(blinds[i]), gens16).compress()
        assert got[i] == want, (c, i)
    # the same rows without blinds through DensePolynomial::commit
    poly = sp.DensePolynomial.new(ctx, s.reshape(L * R, 4))
    rows = dg.commit_poly(poly, L)
    assert rows[1] == bytes(32)
    assert rows[64] == G.commit_vec([O.to_int(x) for x in s[64]], 0, gens16).compress()
    # a slice of rows (one rank's share of a sharded commitment) through the same path
    assert dg.commit_poly_rows(poly, L, 0, L) == b"".join(rows)
    dg.free()


def test_commit_rows_slices(ctx, gens16):
    """spg_poly_commit_rows: any slice of rows equals the same rows of the full commitment
    (the multi-GPU sharding of a commitment relies on it)"""
    import spartan_parallel_b200 as sp

    dg = sp.MultiCommitGens(ctx, gens16.compressed())
    poly = sp.DensePolynomial.new(ctx, rand_scalars(256, 9))  # 16 x 16
    full = b"".join(dg.commit_poly(poly))
    for row0, nrows in [(0, 16), (0, 5), (5, 11), (15, 1), (7, 0)]:
        assert dg.commit_poly_rows(poly, 16, row0, nrows) == full[32 * row0: 32 * (row0 + nrows)]
    with pytest.raises(sp.SpgError):
        dg.commit_poly_rows(poly, 16, 10, 7)


def test_witness_commit_at_config_size(ctx, monkeypatch):
    """BASELINE config C5's witness geometry: a 2^26-entry section = 8192 row commitments over
    8192 generators (src/dense_mlpoly.rs:214-239). Checked (a) against the python oracle on rows
    made sparse enough for it (full-width scalars in chosen columns, incl. the last generator),
    (b) by the relation PolyEvalProof::verify relies on, sum_i L_i C_i == commit(L * Z)
    (src/dense_mlpoly.rs:505-513), with the left side computed over the ROW COMMITMENTS as bases."""
    import hashlib

    import spartan_parallel_b200 as sp

    ell, Lr, R = 26, 8192, 8192
    base = bytes.fromhex("e2f2ae0a6abc4e71a884a961c500515f58e30b6aa582dd8db6a65945e08d2d76")
    uniform = hashlib.shake_256(b"gens_r1cs_sat" + base).digest(64 * (R + 1))
    dg = sp.MultiCommitGens.from_uniform(ctx, uniform)
    rng = np.random.default_rng(26)
    Z = rng.integers(0, 1 << 64, size=(1 << ell, 4), dtype=np.uint64)
    Z[:, 3] &= np.uint64((1 << 60) - 1)
    cols = [0, 1, 4095, 8190, 8191]
    sparse_rows = [0, 4097, 8191]
    for r in sparse_rows:
        keep = Z[r * R + np.array(cols)].copy()
        Z[r * R:(r + 1) * R] = 0
        Z[r * R + np.array(cols)] = keep
    poly = sp.DensePolynomial.new(ctx, Z)
    rows = dg.commit_poly(poly, Lr)
    info = dg.info()
    assert info["rows_table"]["table_bases"] >= R  # 8192 rows: the single-window table + Horner path
    assert info["table_bases"] == 0                # no blinds: the per-window table was not needed (yet)
    lazy_bits = info["rows_table"]["window_bits"]
    assert lazy_bits <= 14                          # built inside the call: sized to pay for itself in one commitment
    dg.prepare(R, Lr)                               # setup-time tables: the widest window the budget allows
    ahead = dg.info()                               # (17 bits on an otherwise empty B200; the allowance shrinks it when
    assert ahead["rows_table"]["window_bits"] > lazy_bits and ahead["table_bases"] >= R  # other tables are alive)
    assert dg.commit_poly(poly, Lr) == rows
    # (a) oracle on the sparse rows: the same generators derived by the oracle's own hash-to-group
    og = {c: G.from_uniform_bytes(uniform[64 * c: 64 * (c + 1)]) for c in cols + [R]}
    for r in sparse_rows:
        want = G.multiscalar_mul([O.to_int(Z[r * R + c]) for c in cols], [og[c] for c in cols]).compress()
        assert rows[r] == want, r
    # (b) linearity over all rows
    r = rand_scalars(ell, 27)
    Lv = O.eq_evals(r[:13])
    LZ = poly.bound(Lv).to_host()
    rhs = dg.commit_batch(LZ.reshape(1, R, 4))[0]
    monkeypatch.setenv("SPG_MSM_WINDOW", "8")  # a second, small table over the row commitments as bases
    row_gens = sp.MultiCommitGens(ctx, b"".join(rows) + og[R].compress())
    lhs = row_gens.commit_batch(Lv.reshape(1, Lr, 4))[0]
    assert lhs == rhs
