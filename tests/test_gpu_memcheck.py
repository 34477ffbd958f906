"""GPU parity for the sparse-polynomial memory-check kernels: hash layer, deref,
product circuits and the batched cubic sumcheck (product_tree.rs, sumcheck.rs:264-434)."""
import numpy as np
import pytest

from oracle import cbind as O
from tests.helpers import rand_scalars

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import spartan_parallel_b200 as sp

    return sp.Context(0)


def oracle_tree(leaves):
    n = leaves.shape[0]
    lefts, rights = [leaves[: n // 2]], [leaves[n // 2:]]
    while lefts[-1].shape[0] > 1:
        l, r = O.prod_layer(lefts[-1], rights[-1])
        lefts.append(l)
        rights.append(r)
    return lefts, rights


@pytest.mark.parametrize("logn", [1, 2, 5, 11, 13])
def test_product_circuit(ctx, logn):
    import spartan_parallel_b200 as sp

    leaves = rand_scalars(1 << logn, 40 + logn)
    pc = sp.ProductCircuit(ctx, sp.DensePolynomial.new(ctx, leaves))
    lefts, rights = oracle_tree(leaves)
    assert pc.num_layers == len(lefts) == logn
    for k in range(pc.num_layers):
        l, r = pc.layer(k)
        assert np.array_equal(l.to_host(), lefts[k]) and np.array_equal(r.to_host(), rights[k])
    assert np.array_equal(pc.evaluate(), O.mul(lefts[-1][0], rights[-1][0]))


def test_hash_layer_and_deref(ctx):
    import spartan_parallel_b200 as sp

    n = 1 << 10
    rng = np.random.default_rng(7)
    mem = rand_scalars(n, 50)
    addr = rng.integers(0, n, size=n, dtype=np.uint64)
    ts = rng.integers(0, 1 << 20, size=n, dtype=np.uint64)
    gamma, tau = rand_scalars(2, 51)
    dmem = sp.DensePolynomial.new(ctx, mem)
    dv = sp.deref(ctx, addr, dmem)
    vals = mem[addr.astype(np.int64)]
    assert np.array_equal(dv.to_host(), vals)
    g2 = O.mul(gamma, gamma)

    def want(a, v, t):
        return O.sub(O.add(O.add(O.mul(O.from_u64(int(t)), g2), O.mul(v, gamma)), O.from_u64(int(a))), tau)

    read = sp.hash_layer(ctx, addr, dv, ts, gamma, tau).to_host()
    write = sp.hash_layer(ctx, addr, dv, ts, gamma, tau, ts_plus_one=True).to_host()
    init = sp.hash_layer(ctx, None, dmem, None, gamma, tau).to_host()
    audit = sp.hash_layer(ctx, None, dmem, ts, gamma, tau).to_host()
    for i in range(0, n, 37):
        assert np.array_equal(read[i], want(addr[i], vals[i], ts[i]))
        assert np.array_equal(write[i], want(addr[i], vals[i], ts[i] + 1))
        assert np.array_equal(init[i], want(i, mem[i], 0))
        assert np.array_equal(audit[i], want(i, mem[i], ts[i]))
    with pytest.raises(sp.SpgError):
        sp.deref(ctx, np.array([n], dtype=np.uint64), dmem)


@pytest.mark.parametrize("logn,npar,nseq", [(1, 1, 0), (2, 2, 1), (6, 3, 0), (9, 4, 2), (12, 12, 6), (4, 0, 2), (15, 3, 1)])
def test_cubic_batched(ctx, logn, npar, nseq):
    import spartan_parallel_b200 as sp

    n = 1 << logn
    seed = 60 + logn
    A_par = [rand_scalars(n, seed + i) for i in range(npar)]
    B_par = [rand_scalars(n, seed + 20 + i) for i in range(npar)]
    C_par = rand_scalars(n, seed + 40)
    A_seq = [rand_scalars(n, seed + 50 + i) for i in range(nseq)]
    B_seq = [rand_scalars(n, seed + 60 + i) for i in range(nseq)]
    C_seq = [rand_scalars(n, seed + 70 + i) for i in range(nseq)]
    coeffs = rand_scalars(npar + nseq, seed + 80)
    ch = rand_scalars(logn, seed + 81)
    up = lambda xs: [sp.DensePolynomial.new(ctx, x) for x in xs]
    dA, dB, dAs, dBs, dCs = up(A_par), up(B_par), up(A_seq), up(B_seq), up(C_seq)
    dC = sp.DensePolynomial.new(ctx, C_par) if npar else None
    sc = sp.CubicBatched(ctx, dA, dB, dC, dAs, dBs, dCs, coeffs)
    oa, ob, oc = [x.copy() for x in A_par], [x.copy() for x in B_par], C_par.copy()
    oas, obs, ocs = [x.copy() for x in A_seq], [x.copy() for x in B_seq], [x.copy() for x in C_seq]
    for j in range(logn):
        want = O.cubic_batched_eval(oa, ob, oc if npar else None, oas, obs, ocs, coeffs)
        assert np.array_equal(sc.round_eval(), want), f"round {j}"
        sc.round_bind(ch[j])
        oa = [O.dense_bound_top(x, ch[j]) for x in oa]
        ob = [O.dense_bound_top(x, ch[j]) for x in ob]
        oc = O.dense_bound_top(oc, ch[j])
        oas = [O.dense_bound_top(x, ch[j]) for x in oas]
        obs = [O.dense_bound_top(x, ch[j]) for x in obs]
        ocs = [O.dense_bound_top(x, ch[j]) for x in ocs]
    got = sc.final()
    want = [x[0] for x in oa] + [x[0] for x in ob] + ([oc[0]] if npar else []) + [x[0] for x in oas] + [x[0] for x in obs] + [x[0] for x in ocs]
    assert np.array_equal(got, np.stack(want))
    # the caller's vectors were bound in place (the shared C table alternates between two buffers on the way):
    # each now holds its one bound scalar
    for vec, w in zip(dA + dB + ([dC] if npar else []) + dAs + dBs + dCs, want):
        assert np.array_equal(vec.to_host()[0], w)


def test_addr_timestamps_on_device(ctx):
    """AddrTimestamps::new (src/sparse_mlpoly.rs:222-253) computed on the device (stable sort by
    address + position in the run) against the reference's sequential counters: heavy repeats of a
    few addresses, address 0 shared with the zero padding of the shorter matrices, cells never
    touched, and the counters running ACROSS the matrices of the batch."""
    import spartan_parallel_b200 as sp

    rng = np.random.default_rng(91)
    nvx, nvy = 6, 9
    nnz = [300, 1, 512, 77]
    polys = []
    for n in nnz:
        rows = rng.integers(0, 1 << nvx, size=n).astype(np.uint32)
        cols = rng.integers(0, 1 << nvy, size=n).astype(np.uint32)
        rows[: n // 3] = 5          # one hot row
        cols[n // 2:] = 0           # address 0 is also what the padding reads
        polys.append((rows, cols, rand_scalars(n, 92 + n)))
    d = sp.MultiSparseMatPolynomialAsDense(ctx, polys, nvx, nvy)
    N, M = d.num_ops, d.num_mem_cells
    assert N == 512 and M == 1 << nvy
    for side, col in (("row", 0), ("col", 1)):
        audit = [0] * M
        for i, p in enumerate(polys):
            addr = np.zeros(N, dtype=np.int64)
            addr[: len(p[col])] = p[col]
            want_ts = []
            for a in addr:
                want_ts.append(audit[a])
                audit[a] += 1
            got_addr = d.view(f"{side}_addr", i).to_host()
            got_ts = d.view(f"{side}_read_ts", i).to_host()
            assert np.array_equal(got_addr, np.stack([O.from_u64(int(a)) for a in addr])), (side, i)
            assert np.array_equal(got_ts, np.stack([O.from_u64(t) for t in want_ts])), (side, i)
        got_audit = d.view(f"{side}_audit_ts").to_host()
        assert np.array_equal(got_audit, np.stack([O.from_u64(t) for t in audit])), side
    d.free()
