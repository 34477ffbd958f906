"""Regenerates the golden fixtures in this directory from the oracle (run from the repo root):
    python tests/golden/make_golden.py
The reference is a Rust crate and cannot run in this image, so these vectors do not come from
it; they freeze the oracle's outputs (which are pinned to the reference's own KATs in
tests/test_oracle_field.py and validated by the oracle's verifier) so that neither the oracle
nor the device path can drift silently. Inputs are fully determined by the seeds below."""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import cbind as O  # noqa: E402
from oracle import protocol as Pr  # noqa: E402
from oracle import r1cs as R  # noqa: E402
from oracle import sparse_proof as SP  # noqa: E402
from tests.helpers import rand_scalars  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def r1cs_case():
    """X = 2^5 constraints x Q = 2 proofs, non-unit coefficients (tests/test_gpu_proof.py::test_small_single_instance)."""
    X, Q = 1 << 5, 2
    inst = R.synthetic_instance(X, unit=False, seed=1)
    secs = R.synthetic_witness(X, [Q], seed=2)
    seed = rand_scalars(1, 3)[0]
    gens = Pr.R1CSGens(b"gens_r1cs_sat", Q * X)
    proof, ch = Pr.r1cs_prove(inst, 1, Q, [Q], X, [X], secs, gens, Pr.Transcript(b"spgpu-parity"), Pr.RandomTape(b"proof", seed))
    return Pr.serialize_r1cs_proof(proof), ch


def tables_case():
    """phase-1 / phase-2 round evaluations for P = 3 ragged instances with injected challenges."""
    P, X = 3, 1 << 4
    num_proofs = [4, 2, 1]
    inst = R.synthetic_instance(X, num_instances=P, unit=False, seed=5)
    secs = R.synthetic_witness(X, num_proofs, seed=6)
    big = rand_scalars(64, 40)
    tau_p, tau_q, tau_x = big[:2], big[8:10], big[16:20]
    ch1, ch2, r_abc = rand_scalars(8, 41), rand_scalars(2 + 1 + 4, 42), rand_scalars(3, 43)
    out = R.prove_tables(inst, P, 4, num_proofs, X, [X] * P, secs, tau_p, tau_q, tau_x, ch1, r_abc, ch2)
    return out


def sparse_case():
    from tests.test_oracle_sparse_proof import evaluate, make_polys

    nvx, nvy, nnz = 3, 3, [8, 8, 8]
    polys = make_polys(np.random.default_rng(31 + nvx + len(nnz)), nvx, nvy, nnz)
    rx, ry = list(rand_scalars(nvx, 11)), list(rand_scalars(nvy, 12))
    evals = [evaluate(p, rx, ry) for p in polys]
    dense = SP.MultiSparseDense(polys)
    gens = SP.SparseGens(b"gens_sparse_poly", nvx, nvy, max(nnz), len(polys))
    pr = SP.sparse_prove(dense, rx, ry, evals, gens, Pr.Transcript(b"spark-parity"), Pr.RandomTape(b"proof", rand_scalars(1, 5)[0]))
    return SP.serialize_sparse_proof(pr)


def witness_case():
    """w3 table (rows of 8 scalars) for three proving instances of 37, 5 and 1 proofs: v in {0, 1}
    (one instance ends in invalid rows), x random; the oracle fills the (pi, D) columns
    (src/lib.rs:1378-1400)."""
    seg_len = [37, 5, 1]
    n = sum(seg_len)
    rng = np.random.default_rng(77)
    w3 = np.zeros((n, 8, 4), dtype=np.uint64)
    valid = rng.integers(0, 2, size=n).astype(bool)
    valid[30:37] = False  # trailing dummy executions of the first instance
    w3[:, 0] = np.where(valid[:, None], O.ONE, np.zeros(4, dtype=np.uint64))
    w3[:, 1] = rand_scalars(n, 78)
    w3[:, 4] = rand_scalars(n, 79)  # untouched columns must survive
    w3 = w3.reshape(n * 8, 4)
    return seg_len, w3, O.perm_fill(w3, seg_len)


def main():
    blob, ch = r1cs_case()
    open(os.path.join(HERE, "r1cs_proof_x32_q2.bin"), "wb").write(blob)
    t = tables_case()
    np.savez(os.path.join(HERE, "tables_p3_ragged.npz"), evals1=np.stack(t.evals1), claims1=np.stack(t.claims1),
             evals2=np.stack(t.evals2), claims2=np.stack(t.claims2))
    sblob = sparse_case()
    open(os.path.join(HERE, "sparse_proof_3x8.bin"), "wb").write(sblob)
    seg_len, w3_in, w3_out = witness_case()
    np.savez(os.path.join(HERE, "w3_perm_ragged.npz"), seg_len=np.array(seg_len), w3_in=w3_in, w3_out=w3_out)
    meta = {"r1cs_proof_x32_q2.bin": hashlib.sha256(blob).hexdigest(), "sparse_proof_3x8.bin": hashlib.sha256(sblob).hexdigest(),
            "challenge_counts": [len(c) for c in ch]}
    json.dump(meta, open(os.path.join(HERE, "MANIFEST.json"), "w"), indent=1)
    print(meta)


if __name__ == "__main__":
    main()
