"""GPU parity for the device field library and dense-table kernels (through the C ABI)
against the oracle. Bit-exact."""
import numpy as np
import pytest

from oracle import cbind as O
from tests.helpers import rand_scalars

pytestmark = pytest.mark.gpu

Q = (1 << 252) + 27742317777372353535851937790883648493


@pytest.fixture(scope="module")
def ctx():
    import spartan_parallel_b200 as sp

    return sp.Context(0)


def _edge_values():
    vals = [0, 1, 2, Q - 1, Q - 2, (1 << 252), (1 << 252) - 1, (1 << 128), (1 << 128) - 1, Q // 2, Q // 2 + 1,
            0xFFFFFFFF, 0x100000000, (1 << 64) - 1, 1 << 64]
    return np.stack([O.from_int(v) for v in vals])


def test_field_ops_bit_exact(ctx):
    import spartan_parallel_b200 as sp

    a = np.concatenate([_edge_values(), rand_scalars(5000, 1)])
    b = np.concatenate([_edge_values()[::-1], rand_scalars(5000, 2)])
    # raw Montgomery limbs near the modulus too (canonical but extreme representations)
    extra = np.array([[0x5812631A5CF5D3EC, 0x14DEF9DEA2F79CD6, 0, 0x1000000000000000], [1, 0, 0, 0], [0, 0, 0, 0]], dtype=np.uint64)
    a = np.concatenate([a, extra])
    b = np.concatenate([b, extra[::-1]])
    da, db = sp.DensePolynomial.new(ctx, a), sp.DensePolynomial.new(ctx, b)
    n = a.shape[0]
    assert np.array_equal(sp.vec_op(ctx, "mul", da, db).to_host()[:n], O.vec_mul(a, b))
    assert np.array_equal(sp.vec_op(ctx, "add", da, db).to_host()[:n], O.vec_add(a, b))
    assert np.array_equal(sp.vec_op(ctx, "sub", da, db).to_host()[:n], O.vec_sub(a, b))
    assert np.array_equal(sp.vec_op(ctx, "square", da).to_host()[:n], O.vec_mul(a, a))
    neg = sp.vec_op(ctx, "neg", da).to_host()[:n]
    canon = sp.vec_op(ctx, "to_canonical", da).to_host()[:n]
    for i in range(0, n, 97):
        assert np.array_equal(neg[i], O.neg(a[i]))
        assert canon[i].tobytes() == O.to_bytes(a[i])


def test_invert_matches_reference_chain(ctx):
    """Scalar::invert (ristretto255.rs:541-595): the device's a^(q-2) equals the oracle's restatement of
    the reference addition chain on edge values and random scalars; 0 maps to 0; a * a^-1 = 1."""
    import spartan_parallel_b200 as sp

    a = np.concatenate([_edge_values(), rand_scalars(300, 9)])
    da = sp.DensePolynomial.new(ctx, a)
    inv = sp.vec_op(ctx, "invert", da)
    got = inv.to_host()[: a.shape[0]]
    for i in range(a.shape[0]):
        if O.to_int(a[i]) == 0:
            assert np.array_equal(got[i], O.ZERO)
        else:
            assert np.array_equal(got[i], O.invert(a[i])), i
    prod = sp.vec_op(ctx, "mul", da, inv).to_host()[: a.shape[0]]
    for i in range(a.shape[0]):
        if O.to_int(a[i]) != 0:
            assert np.array_equal(prod[i], O.ONE)


def test_from_u512(ctx):
    import spartan_parallel_b200 as sp

    rng = np.random.default_rng(3)
    wide = rng.integers(0, 1 << 64, size=(1000, 8), dtype=np.uint64)
    wide[0] = 0xFFFFFFFFFFFFFFFF
    wide[1] = 0
    got = sp.from_u512(ctx, wide).to_host()
    assert np.array_equal(got, O.vec_from_u512(wide))


@pytest.mark.parametrize("ell", [0, 1, 3, 9, 10, 13])
def test_eq_evals(ctx, ell):
    import spartan_parallel_b200 as sp

    r = rand_scalars(max(ell, 1), 10 + ell)[:ell]
    got = sp.EqPolynomial(ctx, r).evals().to_host()
    want = O.eq_evals(r) if ell else O.ONE.reshape(1, 4)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("ell", [1, 4, 11, 14])
def test_dense_bind_evaluate_boundL(ctx, ell):
    import spartan_parallel_b200 as sp

    Z = rand_scalars(1 << ell, 20 + ell)
    r = rand_scalars(ell, 30 + ell)
    p = sp.DensePolynomial.new(ctx, Z)
    assert np.array_equal(p.evaluate(r), O.dense_evaluate(Z, r))
    L = O.eq_evals(r[: ell // 2]) if ell // 2 else O.ONE.reshape(1, 4)
    assert np.array_equal(p.bound(L).to_host(), O.dense_bound_L(Z, L))
    p.bound_poly_var_top(r[0])
    assert np.array_equal(p.to_host(), O.dense_bound_top(Z, r[0]))
    p2 = sp.DensePolynomial.new(ctx, Z)
    p2.bound_poly_var_bot(r[0])
    assert np.array_equal(p2.to_host(), O.dense_bound_bot(Z, r[0]))
    a, b = sp.DensePolynomial.new(ctx, Z), sp.DensePolynomial.new(ctx, Z[::-1].copy())
    assert np.array_equal(sp.dot(ctx, a, b), O.dot(Z, Z[::-1].copy()))


def test_reference_dense_kat(ctx):
    """dense_mlpoly.rs:1233-1252: Z = [1,2,1,4], r = [4,3] -> 28."""
    import spartan_parallel_b200 as sp

    Z = np.stack([O.from_u64(v) for v in (1, 2, 1, 4)])
    r = np.stack([O.from_u64(4), O.from_u64(3)])
    assert np.array_equal(sp.DensePolynomial.new(ctx, Z).evaluate(r), O.from_u64(28))


def test_wide_range_forms_selftest(ctx):
    """csrc/fq.cuh's wide-range forms (a - b + 2q / 6q without a condition, one fold into [0, 2q), products of
    operands up to 6q) exactly as k_rows_rolled / k_rows / k_rows_spmv chain them, on every edge combination of
    {0, 1, q-1, q, q+1, 2q-1} plus random operands below 2q, against canonical arithmetic on the device"""
    import ctypes as C

    from spartan_parallel_b200._lib import check

    bad = C.c_uint32(0xFFFFFFFF)
    check(ctx.L.spg_debug_fq_wide_selftest(ctx.h, 1 << 18, 2024, C.byref(bad)), "spg_debug_fq_wide_selftest")
    assert bad.value == 0, f"failing checks: {bad.value:#x}"
