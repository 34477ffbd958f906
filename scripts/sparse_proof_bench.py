"""Wall time of SparseMatPolynomial::multi_commit + SparseMatPolyEvalProof::prove on the device
(development aid). usage: sparse_proof_bench.py LOG_NNZ BATCH [NVX NVY]"""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import spartan_parallel_b200 as sp
from spartan_parallel_b200 import host

lg, batch = int(sys.argv[1]), int(sys.argv[2])
nvx = int(sys.argv[3]) if len(sys.argv) > 3 else lg
nvy = int(sys.argv[4]) if len(sys.argv) > 4 else lg + 1
ctx = sp.Context(0)
rng = np.random.default_rng(3)
nnz = 1 << lg
polys = []
for b in range(batch):
    rows = rng.integers(0, 1 << nvx, size=nnz).astype(np.uint32)
    cols = rng.integers(0, 1 << nvy, size=nnz).astype(np.uint32)
    vals = rng.integers(0, 1 << 64, size=(nnz, 4), dtype=np.uint64); vals[:, 3] &= np.uint64((1 << 60) - 1)
    polys.append((rows, cols, vals))
def canon(n):
    a = rng.integers(0, 1 << 64, size=(n, 4), dtype=np.uint64); a[:, 3] &= np.uint64((1 << 60) - 1); return a
rx, ry = canon(nvx), canon(nvy)
# evaluations M_i(rx, ry) on the device
mem_rx, mem_ry = sp.EqPolynomial(ctx, rx).evals(), sp.EqPolynomial(ctx, ry).evals()
hrx, hry = mem_rx.to_host(), mem_ry.to_host()
evals = []
for rows, cols, vals in polys:
    a = sp.DensePolynomial.new(ctx, hrx[rows]); b = sp.DensePolynomial.new(ctx, hry[cols]); v = sp.DensePolynomial.new(ctx, vals)
    evals.append(sp.dot(ctx, sp.vec_op(ctx, "mul", a, b), v))
seed = np.array([1, 2, 3, 4], dtype=np.uint64)
for it in range(3):
    t0 = time.perf_counter()
    comm, proof = host.sparse_prove(ctx, polys, nvx, nvy, rx, ry, np.stack(evals), b"bench", b"gens_sparse_poly", seed)
    print(f"sparse commit+prove nnz=2^{lg} x {batch}: {time.perf_counter() - t0:.3f} s, proof {len(proof)} B, commitment {len(comm)} B", flush=True)
