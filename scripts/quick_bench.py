"""Quick device timing (development aid, not the contract bench)."""
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
import spartan_parallel_b200 as sp

ctx = sp.Context(0)
logx, Q = int(sys.argv[1]) if len(sys.argv) > 1 else 20, int(sys.argv[2]) if len(sys.argv) > 2 else 16
X = 1 << logx
N = X * Q
rng = np.random.default_rng(0)
def rnd(n):
    a = rng.integers(0, 1 << 64, size=(n, 4), dtype=np.uint64)
    a[:, 3] &= (1 << 59) - 1   # < 2^251 < q: canonical
    return a
Az, Bz, Cz = rnd(N), rnd(N), rnd(N)
nq = Q.bit_length() - 1
tau_q, tau_x = rnd(max(nq, 1))[:nq], rnd(logx)
ch = rnd(logx + nq)
# elementwise mul throughput
a = sp.DensePolynomial.new(ctx, Az); b = sp.DensePolynomial.new(ctx, Bz)
for _ in range(3):
    c = sp.vec_op(ctx, "mul", a, b)
ctx.sync(); t0 = time.perf_counter()
for _ in range(10):
    c = sp.vec_op(ctx, "mul", a, b)
ctx.sync(); dt = (time.perf_counter() - t0) / 10
print(f"vec mul: N={N} {dt*1e3:.3f} ms  {N/dt/1e9:.2f} Gmul/s  {96*N/dt/1e9:.1f} GB/s")
del a, b, c
for it in range(3):
    sc = sp.SumcheckPhase1.from_tables(ctx, [Q], Q, [X], X, Az, Bz, Cz, tau_q[:0], tau_q, tau_x)
    ctx.sync(); t0 = time.perf_counter()
    times = []
    for j in range(sc.num_rounds):
        t1 = time.perf_counter()
        e = sc.round_eval()
        sc.round_bind(ch[j])
        ctx.sync()
        times.append(time.perf_counter() - t1)
    dt = time.perf_counter() - t0
    print(f"sc1 X=2^{logx} Q={Q}: {dt*1e3:.2f} ms total, {N/dt/1e9:.3f} G cons/s; first rounds ms: {[round(t*1e3,3) for t in times[:6]]} tail: {[round(t*1e3,3) for t in times[-4:]]}")
    sc.free()
print("launches", ctx.launches)
