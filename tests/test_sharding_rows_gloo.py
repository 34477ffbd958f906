"""Multi-process (gloo, CPU) test of the row-sharded multi-instance driver
(parallel.ShardedRows, BASELINE config C4 shape: several instances with different numbers of
proofs and constraints): with a small field-arithmetic engine standing in for the device on each
rank, the sharded phase-1 sumcheck must reproduce the unsharded oracle bit for bit, and the
row-weighted Z binds must add up to the unsharded bound table."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import cbind as O
from tests.helpers import drive_sc1_oracle, log2, rand_scalars


class RowEngine:
    """x rounds of the phase-1 sumcheck over a set of rows with explicit row weights: what
    api.SumcheckPhase1 + spg_sc1_set_row_weights computes on a GPU (natural order, low bit first,
    exhausted rows scaled by 1 - r), written with the oracle's scalar ops."""

    def __init__(self, rows, weights, tau_x, max_x):
        # rows: list of (Az, Bz, Cz) arrays of the row's own length X_p
        self.rows = [[np.array(t, dtype=np.uint64) for t in r] for r in rows]
        self.w = weights
        self.tau, self.j, self.nx = tau_x, 0, log2(max_x)
        self.cx = O.ONE

    def _S(self, j):
        """eq table of ALL taus after round j (index bit k <-> tau[j+1+k]); a shorter row only reaches
        its low entries, which carry the (1 - tau) factors of the bits the row does not have"""
        rest = self.tau[j + 1: self.nx]
        return O.eq_evals(rest[::-1].copy()) if len(rest) else O.ONE.reshape(1, 4)

    def round_eval(self):
        j = self.j
        tau = self.tau[j]
        # line through (0, 1 - tau), (1, tau): l(t) = (1 - tau) + t (2 tau - 1)
        d = O.sub(tau, O.sub(O.ONE, tau))
        l = {0: O.sub(O.ONE, tau)}
        l[2] = O.add(O.add(l[0], d), d)
        l[3] = O.add(l[2], d)
        ev = {0: O.ZERO, 2: O.ZERO, 3: O.ZERO}
        S = self._S(j)
        for (A, B, C), w in zip(self.rows, self.w):
            n = A.shape[0]
            if n == 1:  # exhausted row: high half is zero
                pairs = [(A[0], O.ZERO, B[0], O.ZERO, C[0], O.ZERO)]
            else:
                pairs = [(A[2 * i], A[2 * i + 1], B[2 * i], B[2 * i + 1], C[2 * i], C[2 * i + 1]) for i in range(n // 2)]
            for i, (a0, a1, b0, b1, c0, c1) in enumerate(pairs):
                wi = O.mul(w, S[i])
                for t in (0, 2, 3):
                    tt = O.from_u64(t)
                    at = O.add(a0, O.mul(tt, O.sub(a1, a0)))
                    bt = O.add(b0, O.mul(tt, O.sub(b1, b0)))
                    ct = O.add(c0, O.mul(tt, O.sub(c1, c0)))
                    ev[t] = O.add(ev[t], O.mul(wi, O.sub(O.mul(at, bt), ct)))
        return np.stack([O.mul(O.mul(self.cx, l[t]), ev[t]) for t in (0, 2, 3)])

    def round_bind(self, r):
        tau = self.tau[self.j]
        for row in self.rows:
            for k in range(3):
                T = row[k]
                if T.shape[0] == 1:
                    row[k] = np.stack([O.mul(O.sub(O.ONE, r), T[0])])
                else:
                    row[k] = np.stack([O.add(T[2 * i], O.mul(r, O.sub(T[2 * i + 1], T[2 * i]))) for i in range(T.shape[0] // 2)])
        self.cx = O.mul(self.cx, O.add(O.mul(tau, r), O.mul(O.sub(O.ONE, tau), O.sub(O.ONE, r))))
        self.j += 1

    def debug_tables(self):
        return [np.stack([row[k][0] for row in self.rows]) for k in range(3)]


class OracleTail:
    def __init__(self, sc):
        self.sc, self.scale = sc, O.ONE

    def set_scale(self, c):
        self.scale = np.asarray(c, dtype=np.uint64)

    def round_eval(self):
        return np.stack([O.mul(x, self.scale) for x in self.sc.round_eval()])

    def round_bind(self, r):
        self.sc.round_bind(r)

    def final(self):
        f = self.sc.final()
        f[0] = O.mul(f[0], self.scale)
        return f


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, num_proofs, num_cons):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import spartan_parallel_b200 as sp
        from spartan_parallel_b200 import parallel

        P, Qmax, Xmax = len(num_proofs), max(num_proofs), max(num_cons)
        Pp = parallel.next_pow2(P)
        nx, nq, np_ = log2(Xmax), log2(Qmax), log2(Pp)
        N = sum(q * x for q, x in zip(num_proofs, num_cons))
        Az, Bz, Cz = rand_scalars(N, 1), rand_scalars(N, 2), rand_scalars(N, 3)
        big = rand_scalars(64, 4)
        tau_p, tau_q, tau_x = big[:np_], big[8:8 + nq], big[16:16 + nx]
        ch = rand_scalars(nx + nq + np_, 5)
        want, want_final = drive_sc1_oracle(num_proofs, Qmax, num_cons, Xmax, Az, Bz, Cz, tau_p, tau_q, tau_x, ch)
        off, t = [], 0
        for q, x in zip(num_proofs, num_cons):
            off.append(t)
            t += q * x
        comm = parallel.TorchComm()

        def make_engine(blocks, weights):
            rows = []
            for p, q0, cnt in blocks:
                X = num_cons[p]
                for q in range(q0, q0 + cnt):
                    lo = off[p] + q * X
                    rows.append((Az[lo:lo + X], Bz[lo:lo + X], Cz[lo:lo + X]))
            return RowEngine(rows, weights, tau_x, Xmax)

        def make_tail(a, b, c):
            one = O.ONE.reshape(1, 4)
            mk = lambda T: O.Pqx.new_rev(T, 1, num_proofs, Qmax, [1] * P, 1)
            Ap = O.eq_evals(tau_p) if np_ else one
            Aq = O.eq_evals(tau_q) if nq else one
            return OracleTail(O.Sc1(0, nq, np_, list(num_proofs), [1] * P, Ap, Aq, one, mk(a), mk(b), mk(c)))

        sh = parallel.ShardedRows(comm, num_proofs, Xmax, tau_p, tau_q, tau_x, make_engine, make_tail)
        assert sh.num_rounds == nx + nq + np_
        got = sh.run_rounds(ch)
        for j in range(sh.num_rounds):
            assert np.array_equal(got[j], want[j]), f"rank {rank} round {j}"
        assert np.array_equal(sh.final(), want_final), f"rank {rank} final claims"

        # row-weighted Z bind: the ranks' partial tables add up to the unsharded bind of every instance
        WY = 6
        rq_rev = rand_scalars(max(nq, 1), 6)[:nq]
        E = O.eq_evals(rq_rev[::-1].copy()) if nq else O.ONE.reshape(1, 4)
        part = np.zeros((P, WY, 4), dtype=np.uint64)
        wantZ = np.zeros((P, WY, 4), dtype=np.uint64)
        for p in range(P):
            Zp = rand_scalars(num_proofs[p] * WY, 70 + p).reshape(num_proofs[p], WY, 4)
            for q in range(num_proofs[p]):
                wantZ[p] = O.vec_add(wantZ[p], O.vec_mul(np.tile(E[q], (WY, 1)), Zp[q]))
            for (pp, q0, cnt) in sh.mine:
                if pp != p:
                    continue
                for q in range(q0, q0 + cnt):
                    part[p] = O.vec_add(part[p], O.vec_mul(np.tile(sp.host_eq_weight(rq_rev, q), (WY, 1)), Zp[q]))
        allp = comm.all_gather(part.reshape(P * WY, 4))
        assert np.array_equal(sp.host_sum(allp), wantZ.reshape(P * WY, 4))
    finally:
        dist.destroy_process_group()


def test_partition_covers_every_row_once():
    from spartan_parallel_b200 import parallel

    for world in (1, 2, 4, 8):
        for qs in ([64, 16, 16, 4, 1], [1], [2, 1, 1], [8, 8, 8]):
            seen = set()
            for r, blk in enumerate(parallel.partition_rows(qs, world)):
                assert len({p for p, _, _ in blk}) == len(blk)  # one block per instance and rank
                for p, q0, cnt in blk:
                    assert cnt & (cnt - 1) == 0
                    for q in range(q0, q0 + cnt):
                        assert (p, q) not in seen
                        seen.add((p, q))
            assert seen == {(p, q) for p, Q in enumerate(qs) for q in range(Q)}


@pytest.mark.parametrize("world,num_proofs,num_cons", [(2, [8, 4, 4, 2, 1], [8, 4, 8, 2, 4]), (4, [4, 2, 1], [4, 4, 2]), (2, [2], [8])])
def test_row_sharded_phase1_matches_unsharded(world, num_proofs, num_cons):
    mp.spawn(_worker, args=(world, _free_port(), num_proofs, num_cons), nprocs=world, join=True)
