"""Parity of the sparse-polynomial evaluation proof (memory-check network): the device +
C++ host mirror must emit, byte for byte, the commitment and proof the oracle's restatement
of SparseMatPolynomial::multi_commit / SparseMatPolyEvalProof::prove emits, and the oracle's
restatement of ::verify must accept the device bytes."""
import os

import numpy as np
import pytest

from oracle import cbind as O
from oracle import sparse_proof as SP
from oracle.protocol import RandomTape, Transcript, Writer
from tests.helpers import rand_scalars
from tests.test_oracle_sparse_proof import evaluate, make_polys

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import spartan_parallel_b200 as sp

    return sp.Context(0)


def comm_bytes(c):
    w = Writer()
    w.u64(c["batch_size"])
    w.u64(c["num_ops"])
    w.u64(c["num_mem_cells"])
    w.points(c["comm_comb_ops"])
    w.points(c["comm_comb_mem"])
    return bytes(w.b)


@pytest.mark.parametrize("nvx,nvy,nnz_list", [(3, 3, [8, 8, 8]), (2, 4, [5, 7]), (4, 3, [16]), (6, 6, [100, 64, 90])])
def test_sparse_proof_bytes(ctx, nvx, nvy, nnz_list):
    from spartan_parallel_b200 import host

    rng = np.random.default_rng(31 + nvx + len(nnz_list))
    polys = make_polys(rng, nvx, nvy, nnz_list)
    rx, ry = list(rand_scalars(nvx, 11)), list(rand_scalars(nvy, 12))
    evals = [evaluate(p, rx, ry) for p in polys]
    seed = rand_scalars(1, 5)[0]

    dense = SP.MultiSparseDense(polys)
    gens = SP.SparseGens(b"gens_sparse_poly", nvx, nvy, max(nnz_list), len(polys))
    want_comm = SP.multi_commit(dense, gens)
    want = SP.serialize_sparse_proof(SP.sparse_prove(dense, rx, ry, evals, gens, Transcript(b"spark-parity"), RandomTape(b"proof", seed)))

    got_comm, got = host.sparse_prove(ctx, [(p[0], p[1], p[2]) for p in polys], nvx, nvy, np.stack(rx), np.stack(ry),
                                      np.stack(evals), b"spark-parity", b"gens_sparse_poly", seed)
    assert got_comm == comm_bytes(want_comm)
    assert len(got) == len(want)
    if got != want:
        first = next(i for i in range(len(want)) if got[i] != want[i])
        raise AssertionError(f"proof bytes differ from offset {first} of {len(want)}")
    assert SP.sparse_verify(SP.deserialize_sparse_proof(got), want_comm, rx, ry, evals, gens, Transcript(b"spark-parity"))
    if (nvx, nvy, nnz_list) == (3, 3, [8, 8, 8]):
        golden = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sparse_proof_3x8.bin")
        assert got == open(golden, "rb").read(), "device proof differs from the committed golden fixture"


def test_wrong_evaluation_is_refused(ctx):
    from spartan_parallel_b200 import SpgError, host

    rng = np.random.default_rng(3)
    polys = make_polys(rng, 3, 3, [8])
    rx, ry = list(rand_scalars(3, 11)), list(rand_scalars(3, 12))
    bad = [O.add(evaluate(polys[0], rx, ry), O.ONE)]
    with pytest.raises(SpgError, match="claimed evaluation is wrong"):
        host.sparse_prove(ctx, [(p[0], p[1], p[2]) for p in polys], 3, 3, np.stack(rx), np.stack(ry), np.stack(bad), b"x",
                          b"gens_sparse_poly", rand_scalars(1, 5)[0])
