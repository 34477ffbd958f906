"""Wall time of the whole R1CSProof::prove (device kernels + C++ host mirror), development aid.
usage: full_proof_bench.py LOG_X Q [device_gens=1]"""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import spartan_parallel_b200 as sp
from spartan_parallel_b200 import host

log_x, Q = int(sys.argv[1]), int(sys.argv[2])
dev = int(sys.argv[3]) if len(sys.argv) > 3 else 1
ctx = sp.Context(0)
X, N = 1 << log_x, (1 << log_x) * Q
ONE = np.array([0xD6EC31748D98951D, 0xC6EF5BF4737DCF70, 0xFFFFFFFFFFFFFFFE, 0x0FFFFFFFFFFFFFFF], dtype=np.uint64)
rng = np.random.default_rng(1)
u = rng.integers(0, 1 << 64, size=(N, 4), dtype=np.uint64); u[:, 3] &= np.uint64((1 << 60) - 1)
du = sp.DensePolynomial.new(ctx, u)
dun = sp.DensePolynomial.new(ctx, np.ascontiguousarray(np.roll(u.reshape(Q, X, 4), -1, axis=1).reshape(N, 4)))
v = sp.vec_op(ctx, "mul", du, dun).to_host()
del du, dun
rows = np.arange(X, dtype=np.uint32); ones = np.tile(ONE, (X, 1))
inst = sp.R1CSInstance(ctx, 1, X, [X], 2 * X, [(rows, rows, ones)], [(rows, ((rows + 1) % X).astype(np.uint32), ones)], [(rows, (rows + X).astype(np.uint32), ones)])
secs = [sp.ProverWitnessSecInfo(ctx, [Q], [X], u), sp.ProverWitnessSecInfo(ctx, [Q], [X], v)]
t0 = time.perf_counter()
gens = host.R1CSGens(ctx, b"gens_r1cs_sat", N) if dev else None
print(f"gens: {time.perf_counter() - t0:.2f} s", flush=True)
seed = np.array([1, 2, 3, 4], dtype=np.uint64)
for it in range(2):
    t0 = time.perf_counter()
    proof, ch = host.r1cs_prove(ctx, inst, secs, [Q], Q, [X], X, b"bench", b"gens_r1cs_sat", seed, N, gens)
    print(f"R1CSProof::prove X=2^{log_x} Q={Q} device_gens={dev}: {time.perf_counter() - t0:.3f} s, {len(proof)} bytes", flush=True)
