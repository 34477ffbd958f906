"""Per-round latency of the sumcheck loops on tiny tables (development aid)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import spartan_parallel_b200 as sp

ctx = sp.Context(0)
rng = np.random.default_rng(0)
def rnd(n):
    a = rng.integers(0, 1 << 64, size=(n, 4), dtype=np.uint64); a[:, 3] &= np.uint64((1 << 59) - 1); return a
for logx, Q in ((12, 1), (16, 1), (10, 16)):
    X = 1 << logx; N = X * Q; nq = Q.bit_length() - 1
    Az, Bz, Cz = rnd(N), rnd(N), rnd(N)
    tau_q, tau_x, ch = rnd(max(nq, 1))[:nq], rnd(logx), rnd(logx + nq)
    best = 1e9
    for it in range(5):
        sc = sp.SumcheckPhase1.from_tables(ctx, [Q], Q, [X], X, Az, Bz, Cz, tau_q[:0], tau_q, tau_x)
        ctx.sync(); t0 = time.perf_counter()
        sc.run_rounds(ch)
        dt = time.perf_counter() - t0
        best = min(best, dt)
        sc.free()
    print(f"X=2^{logx} Q={Q}: {len(ch)} rounds in {best*1e6:.0f} us = {best*1e6/len(ch):.1f} us/round")
