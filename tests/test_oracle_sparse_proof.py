"""Oracle check for the sparse-polynomial evaluation proof: the reference's own live test for
this path is the prove -> verify round trip (src/sparse_mlpoly.rs:1605-1676); here with tamper checks."""
import copy

import numpy as np
import pytest

from oracle import cbind as O
from oracle import sparse_proof as SP
from oracle.protocol import RandomTape, Transcript, ONE, add

from .helpers import rand_scalars


def make_polys(rng, nvx, nvy, nnz_list):
    polys = []
    for nnz in nnz_list:
        rows = rng.integers(0, 1 << nvx, size=nnz)
        cols = rng.integers(0, 1 << nvy, size=nnz)
        vals = rand_scalars(nnz, int(rng.integers(1 << 30)))
        polys.append((rows, cols, vals, nvx, nvy))
    return polys


def evaluate(poly, rx, ry):
    rows, cols, vals, _, _ = poly
    return O.sparse_evaluate_with_tables(rows, cols, vals, SP.eq_evals(rx), SP.eq_evals(ry))


@pytest.mark.parametrize("nvx,nvy,nnz_list", [(3, 3, [8, 8, 8]), (2, 4, [5, 7]), (4, 3, [16])])
def test_sparse_proof_roundtrip(nvx, nvy, nnz_list):
    rng = np.random.default_rng(7 + nvx)
    polys = make_polys(rng, nvx, nvy, nnz_list)
    dense = SP.MultiSparseDense(polys)
    gens = SP.SparseGens(b"gens_sparse_poly", nvx, nvy, max(nnz_list), len(polys))
    comm = SP.multi_commit(dense, gens)
    rx, ry = list(rand_scalars(nvx, 11)), list(rand_scalars(nvy, 12))
    evals = [evaluate(p, rx, ry) for p in polys]

    tape = RandomTape(b"proof", O.from_u64(99))
    pt = Transcript(b"example")
    proof = SP.sparse_prove(dense, rx, ry, evals, gens, pt, tape)
    blob = SP.serialize_sparse_proof(proof)
    proof2 = SP.deserialize_sparse_proof(blob)
    assert SP.serialize_sparse_proof(proof2) == blob

    assert SP.sparse_verify(proof2, comm, rx, ry, evals, gens, Transcript(b"example"))
    # prover and verifier transcripts stay in lock step
    vt = Transcript(b"example")
    SP.sparse_verify(proof2, comm, rx, ry, evals, gens, vt)
    assert vt.challenge_bytes(b"x", 32) == pt.challenge_bytes(b"x", 32)

    bad_evals = [add(evals[0], ONE)] + evals[1:]
    assert not SP.sparse_verify(proof2, comm, rx, ry, bad_evals, gens, Transcript(b"example"))
    bad = copy.deepcopy(proof2)
    bad["hash_layer"]["eval_val"][0] = add(bad["hash_layer"]["eval_val"][0], ONE)
    assert not SP.sparse_verify(bad, comm, rx, ry, evals, gens, Transcript(b"example"))
    bad = copy.deepcopy(proof2)
    a, r_, au = bad["hash_layer"]["eval_row"]
    bad["hash_layer"]["eval_row"] = (a, r_, add(au, ONE))
    assert not SP.sparse_verify(bad, comm, rx, ry, evals, gens, Transcript(b"example"))
