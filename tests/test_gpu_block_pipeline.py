"""One block instance end to end, the way SNARK::prove handles it (src/lib.rs:1481-1676): the primary
section comes out of a RunTimeKnowledge file (formats.py), is uploaded once, the derived sections
block_w2 / block_w3 / block_w3_shifted are computed on the device and committed there without
returning to the host -- against the oracle's tables and the oracle's Pedersen commitments of them."""
import numpy as np
import pytest

from oracle import cbind as O
from oracle import protocol as P
from tests.helpers import rand_scalars

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import spartan_parallel_b200 as sp

    return sp.Context(0)


def test_block_sections_from_rtk_to_commitments(ctx, tmp_path):
    import spartan_parallel_b200 as sp
    from spartan_parallel_b200 import formats as F

    n, phy, vir, rows = 3, 2, 1, 16
    io_width, vars_width, w2_width = 2 * n, 16, 16
    vars_ = rand_scalars(rows * vars_width, 700).reshape(rows, vars_width, 4)
    for q in range(rows):
        vars_[q, 0] = O.ONE if q < 13 else 0  # three padded (invalid) executions at the end
    # the block's executions as the prover's front end hands them over (.rtk, examples/interface.rs:197-220)
    rtk = {"block_max_num_proofs": rows, "block_num_proofs": [rows], "consis_num_proofs": rows,
           "total_num_init_phy_mem_accesses": 0, "total_num_init_vir_mem_accesses": 0, "total_num_phy_mem_accesses": 0,
           "total_num_vir_mem_accesses": 0, "block_vars_matrix": [[{"assignment": list(vars_[q])} for q in range(rows)]],
           "exec_inputs": [], "init_phy_mems_list": [], "init_vir_mems_list": [], "addr_phy_mems_list": [], "addr_vir_mems_list": [],
           "addr_ts_bits_list": [], "input": [], "input_stack": [], "input_mem": [], "output": bytes(32), "output_exec_num": 0}
    path = tmp_path / "blk_bin.rtk"
    path.write_bytes(F.encode(F.RunTimeKnowledge, rtk))
    table = F.block_witness_tables(F.read_rtk(str(path)))[0]
    assert np.array_equal(table, vars_)

    tau, r = rand_scalars(2, 701)
    w0 = O.wit_perm_w0(tau, r, 2 * n, 8)
    want2, want3 = O.wit_block(vars_, w0, tau, r, n, io_width, phy, vir, w2_width)
    want3s = O.wit_shift(want3)

    d_vars = sp.DensePolynomial.new(ctx, table.reshape(-1, 4))
    d_w0 = sp.wit_perm_w0(ctx, tau, r, 2 * n, 8)
    w2, w3 = sp.wit_block(ctx, d_vars, rows, vars_width, d_w0, tau, r, n, io_width, phy, vir, w2_width)
    w3s = sp.wit_shift(ctx, w3, rows)
    # commitments of the four polynomials the prover appends to its transcript (src/lib.rs:1624-1676), computed
    # from the device-resident tables; generators as R1CSGens derives them
    gens = P.DotProductProofGens(16, b"gens_r1cs_sat")
    dg = sp.MultiCommitGens(ctx, gens.gens_n.compressed())
    for dev, want in ((d_vars, vars_), (w2, want2), (w3, want3), (w3s, want3s)):
        flat = want.reshape(-1, 4)
        assert np.array_equal(dev.to_host()[: flat.shape[0]], flat)
        got = dg.commit_poly(dev)
        assert got == P.poly_commit(flat, gens.gens_n)
