"""GPU parity for the derived witness sections SNARK::prove builds before committing
(src/lib.rs:1328-1400 perm_w0 / perm_exec, :1481-1613 block_w2 / block_w3,
:832-880 mem_gen, :1667-1676 shifted copies) against the oracle's sequential restatement:
every table bit for bit."""
import numpy as np
import pytest

from oracle import cbind as O
from tests.helpers import rand_scalars

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import spartan_parallel_b200 as sp

    return sp.Context(0)


def _rows(n_rows, width, seed, valid=None):
    a = rand_scalars(n_rows * width, seed).reshape(n_rows, width, 4)
    # column 0 is the validity flag v of the row: 0 / 1 like the reference's padded executions
    flags = np.ones(n_rows, dtype=bool) if valid is None else valid
    for q in range(n_rows):
        a[q, 0] = O.ONE if flags[q] else 0
    return a


@pytest.mark.parametrize("used,total", [(6, 8), (0, 4), (16, 16), (130, 256)])
def test_perm_w0(ctx, used, total):
    import spartan_parallel_b200 as sp

    tau, r = rand_scalars(2, 300 + used)
    got = sp.wit_perm_w0(ctx, tau, r, used, total).to_host()
    assert np.array_equal(got, O.wit_perm_w0(tau, r, used, total))


@pytest.mark.parametrize("rows,n,num_ios", [(1, 2, 4), (5, 3, 8), (64, 9, 32), (300, 33, 128)])
def test_perm_exec_sections(ctx, rows, n, num_ios):
    import spartan_parallel_b200 as sp

    tau, r = rand_scalars(2, 310 + rows)
    w0 = O.wit_perm_w0(tau, r, 2 * n, num_ios)
    valid = np.arange(rows) < max(1, rows - rows // 4)  # trailing invalid executions
    inputs = _rows(rows, num_ios, 311 + rows, valid)
    want2, want3 = O.wit_exec(inputs, w0, tau, n, num_ios)
    d_in = sp.DensePolynomial.new(ctx, inputs.reshape(-1, 4))
    d_w0 = sp.DensePolynomial.new(ctx, w0)
    w2, w3 = sp.wit_block(ctx, d_in, rows, num_ios, d_w0, tau, r, n, w2_width=num_ios, exec_mode=True)
    assert np.array_equal(w2.to_host()[: rows * num_ios].reshape(rows, num_ios, 4), want2)
    assert np.array_equal(w3.to_host()[: rows * 8].reshape(rows, 8, 4), want3)
    sh = sp.wit_shift(ctx, w3, rows).to_host()[: rows * 8].reshape(rows, 8, 4)
    assert np.array_equal(sh, O.wit_shift(want3))


@pytest.mark.parametrize("rows,n,phy,vir", [(8, 3, 0, 0), (16, 4, 3, 0), (16, 4, 0, 2), (128, 5, 4, 3), (33, 17, 1, 1)])
def test_block_sections(ctx, rows, n, phy, vir):
    """one block instance: INPUT part, PHY and VIR memory chains, the three (pi, D) pairs"""
    import spartan_parallel_b200 as sp

    tau, r = rand_scalars(2, 320 + rows + phy)
    io_width = 2 * n
    vars_width = 1 << (io_width + 2 * phy + 4 * vir - 1).bit_length()
    w2_width = 1 << (2 * n + 2 * phy + 4 * vir - 1).bit_length()
    w0 = O.wit_perm_w0(tau, r, 2 * n, max(2 * n, 4))
    valid = np.arange(rows) < max(1, rows - rows // 3)
    vars_ = _rows(rows, vars_width, 321 + rows, valid)
    want2, want3 = O.wit_block(vars_, w0, tau, r, n, io_width, phy, vir, w2_width)
    d_vars = sp.DensePolynomial.new(ctx, vars_.reshape(-1, 4))
    d_w0 = sp.DensePolynomial.new(ctx, w0)
    w2, w3 = sp.wit_block(ctx, d_vars, rows, vars_width, d_w0, tau, r, n, io_width, phy, vir, w2_width)
    assert np.array_equal(w2.to_host()[: rows * w2_width].reshape(rows, w2_width, 4), want2)
    assert np.array_equal(w3.to_host()[: rows * 8].reshape(rows, 8, 4), want3)


def test_block_sections_two_instances_in_one_call(ctx):
    """seg_len with several instances: the recurrences restart at each instance's last row,
    and the shifted copy puts a zero row at the end of EACH instance"""
    import spartan_parallel_b200 as sp

    n, phy, vir, seg = 3, 2, 1, [5, 11]
    rows = sum(seg)
    tau, r = rand_scalars(2, 330)
    io_width, vars_width, w2_width = 2 * n, 16, 16
    w0 = O.wit_perm_w0(tau, r, 2 * n, 8)
    vars_ = _rows(rows, vars_width, 331)
    parts = [O.wit_block(vars_[a:b], w0, tau, r, n, io_width, phy, vir, w2_width) for a, b in ((0, 5), (5, 16))]
    want2, want3 = np.concatenate([p[0] for p in parts]), np.concatenate([p[1] for p in parts])
    d_vars = sp.DensePolynomial.new(ctx, vars_.reshape(-1, 4))
    w2, w3 = sp.wit_block(ctx, d_vars, rows, vars_width, sp.DensePolynomial.new(ctx, w0), tau, r, n, io_width, phy, vir, w2_width, seg_len=seg)
    assert np.array_equal(w2.to_host()[: rows * w2_width].reshape(rows, w2_width, 4), want2)
    assert np.array_equal(w3.to_host()[: rows * 8].reshape(rows, 8, 4), want3)
    sh = sp.wit_shift(ctx, w3, rows, 8, seg).to_host()[: rows * 8].reshape(rows, 8, 4)
    assert np.array_equal(sh, np.concatenate([O.wit_shift(want3[:5]), O.wit_shift(want3[5:])]))


@pytest.mark.parametrize("rows,in_width,mem_width", [(1, 4, 4), (37, 4, 4), (500, 8, 8)])
def test_mem_sections(ctx, rows, in_width, mem_width):
    import spartan_parallel_b200 as sp

    tau, r = rand_scalars(2, 340 + rows)
    mems = _rows(rows, in_width, 341 + rows, np.arange(rows) < max(1, rows - 3))
    want2, want3 = O.wit_mem(mems, tau, r, mem_width)
    w2, w3 = sp.wit_mem(ctx, sp.DensePolynomial.new(ctx, mems.reshape(-1, 4)), rows, in_width, tau, r, mem_width)
    assert np.array_equal(w2.to_host()[: rows * mem_width].reshape(rows, mem_width, 4), want2)
    assert np.array_equal(w3.to_host()[: rows * 8].reshape(rows, 8, 4), want3)


def test_bad_shapes_rejected(ctx):
    import spartan_parallel_b200 as sp

    tau, r = rand_scalars(2, 350)
    v = sp.DensePolynomial.new(ctx, rand_scalars(64, 351))
    w0 = sp.DensePolynomial.new(ctx, rand_scalars(8, 352))
    with pytest.raises(sp.SpgError):
        sp.wit_block(ctx, v, 8, 8, w0, tau, r, 3, 6, 4, 0, 8)  # 6 + 8 memory scalars do not fit rows of 8
    with pytest.raises(sp.SpgError):
        sp.wit_block(ctx, v, 16, 8, w0, tau, r, 3, w2_width=8, exec_mode=True)  # 16 rows x 8 > 64
    with pytest.raises(sp.SpgError):
        sp.wit_shift(ctx, v, 8, 8, [3, 4])  # segments do not cover the rows
