"""Multi-process (gloo, CPU) tests of the proof-axis sharding logic in
spartan_parallel_b200/parallel.py: with the oracle standing in for the per-rank device
engine, the sharded phase-1 sumcheck must reproduce the unsharded one bit for bit, and
the sharded Z-bind must sum to the unsharded table."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import cbind as O
from tests.helpers import log2, rand_scalars

ONE1 = O.ONE.reshape(1, 4)


class OracleEngine:
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


def mk_sc1(nx, nq, Q, X, Az, Bz, Cz, tau_q, tau_x):
    mk = lambda T: O.Pqx.new_rev(T, 1, [Q], Q, [X], X)
    Aq = O.eq_evals(tau_q) if nq else ONE1
    Ax = O.eq_evals(tau_x) if nx else ONE1
    return O.Sc1(nx, nq, 0, [Q], [X], ONE1, Aq, Ax, mk(Az), mk(Bz), mk(Cz))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, Q_local, X, use_shm=False, host_tail=False):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from spartan_parallel_b200 import parallel

        Q = Q_local * world
        nx, nq = log2(X), log2(Q)
        N = Q * X
        Az, Bz, Cz = rand_scalars(N, 1), rand_scalars(N, 2), rand_scalars(N, 3)
        tau_q, tau_x = rand_scalars(max(nq, 1), 4)[:nq], rand_scalars(max(nx, 1), 5)[:nx]
        ch = rand_scalars(nx + nq, 6)
        # unsharded oracle
        full = mk_sc1(nx, nq, Q, X, Az, Bz, Cz, tau_q, tau_x)
        want = []
        for j in range(nx + nq):
            want.append(full.round_eval())
            full.round_bind(ch[j])
        want_final = full.final()
        # sharded: this rank owns proofs [rank * Q_local, (rank + 1) * Q_local)
        lo, hi = rank * Q_local * X, (rank + 1) * Q_local * X
        comm = parallel.ShmComm() if use_shm else parallel.TorchComm()

        def make_engine(tau_q_local):
            return OracleEngine(mk_sc1(nx, log2(Q_local), Q_local, X, Az[lo:hi], Bz[lo:hi], Cz[lo:hi], tau_q_local, tau_x))

        def make_tail(a, b, c, tau_high):
            G = a.shape[0]
            if host_tail:  # the library's host-side tail (spg_sc1_host_tail_*), what gpu_phase1 uses for G <= 64
                return parallel.HostTail(a, b, c, tau_high)
            return OracleEngine(mk_sc1(0, log2(G), G, 1, a, b, c, tau_high, tau_x[:0]))

        sh = parallel.ShardedPhase1(comm, Q_local, X, tau_q, tau_x, make_engine, make_tail)
        assert sh.num_rounds == nx + nq
        for j in range(sh.num_rounds):
            got = sh.round_eval()
            assert np.array_equal(got, want[j]), f"rank {rank} round {j}"
            sh.round_bind(ch[j])
        assert np.array_equal(sh.final(), want_final), f"rank {rank} final claims"
        # the all-rounds driver (the GPU engine runs its local rounds in one C loop; here the python path)
        sh2 = parallel.ShardedPhase1(comm, Q_local, X, tau_q, tau_x, make_engine, make_tail)
        assert np.array_equal(sh2.run_rounds(ch), np.stack(want)), f"rank {rank} run_rounds"
        assert np.array_equal(sh2.final(), want_final)

        # sharded Z-bind: partial tables scaled by the rank weight sum to the full bind
        import spartan_parallel_b200 as sp

        WY = 24
        Zfull = rand_scalars(Q * WY, 7).reshape(Q, WY, 4)
        rq_rev = rand_scalars(nq, 8)
        E = O.eq_evals(rq_rev[::-1].copy()) if nq else ONE1  # LSB-first table
        wantZ = np.stack([O.ZERO] * WY)
        for q in range(Q):
            wantZ = O.vec_add(wantZ, O.vec_mul(np.tile(E[q], (WY, 1)), Zfull[q]))
        nql = log2(Q_local)
        El = O.eq_evals(rq_rev[:nql][::-1].copy()) if nql else ONE1
        wgt = sp.host_eq_weight(rq_rev[nql:], rank)
        part = np.stack([O.ZERO] * WY)
        for ql in range(Q_local):
            part = O.vec_add(part, O.vec_mul(np.tile(O.mul(El[ql], wgt), (WY, 1)), Zfull[rank * Q_local + ql]))
        allp = comm.all_gather(part)
        assert np.array_equal(sp.host_sum(allp), wantZ)
        if use_shm:
            comm.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,Q_local,X", [(2, 4, 8), (4, 2, 4), (2, 1, 16)])
def test_sharded_phase1_matches_unsharded(world, Q_local, X):
    mp.spawn(_worker, args=(world, _free_port(), Q_local, X), nprocs=world, join=True)


def _live_worker(rank, world, port, Q_local, X):
    """a live transcript on rank 0 (the oracle's ZK sumcheck glue: merlin, tape, per-round dot-product
    proofs) over the sharded prover; the other ranks follow the challenges it publishes"""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import protocol as P
        from spartan_parallel_b200 import parallel

        Q = Q_local * world
        nx, nq = log2(X), log2(Q)
        N = Q * X
        Az, Bz, Cz = rand_scalars(N, 1), rand_scalars(N, 2), rand_scalars(N, 3)
        tau_q, tau_x = rand_scalars(max(nq, 1), 4)[:nq], rand_scalars(max(nx, 1), 5)[:nx]
        lo, hi = rank * Q_local * X, (rank + 1) * Q_local * X
        comm = parallel.TorchComm()
        make_engine = lambda tq: OracleEngine(mk_sc1(nx, log2(Q_local), Q_local, X, Az[lo:hi], Bz[lo:hi], Cz[lo:hi], tq, tau_x))
        make_tail = lambda a, b, c, th: parallel.HostTail(a, b, c, th)
        sh = parallel.ShardedPhase1(comm, Q_local, X, tau_q, tau_x, make_engine, make_tail)
        gens = P.R1CSGens(b"gens_live", 16)
        seed = rand_scalars(1, 9)[0]
        # the claim of these random tables (they satisfy nothing): sum of eq(tau, (q, x)) * (Az Bz - Cz), natural
        # bit k of q / x pairing with tau_q[k] / tau_x[k]
        import spartan_parallel_b200 as sp

        claim = O.ZERO
        for q in range(Q):
            wq = sp.host_eq_weight(tau_q, q)
            for x in range(X):
                i = q * X + x
                w = O.mul(wq, sp.host_eq_weight(tau_x, x))
                claim = O.add(claim, O.mul(w, O.sub(O.mul(Az[i], Bz[i]), Cz[i])))
        if rank == 0:
            t, tape = P.Transcript(b"live"), P.RandomTape(b"proof", seed)
            got, r_got, _ = P.zk_sumcheck_prove(claim, O.ZERO, nx + nq, parallel.LeaderRounds(sh, comm), gens.gens_1, gens.gens_4, t, tape)
            t2, tape2 = P.Transcript(b"live"), P.RandomTape(b"proof", seed)
            want, r_want, _ = P.zk_sumcheck_prove(claim, O.ZERO, nx + nq, OracleEngine(mk_sc1(nx, nq, Q, X, Az, Bz, Cz, tau_q, tau_x)),
                                                  gens.gens_1, gens.gens_4, t2, tape2)
            assert got["comm_polys"] == want["comm_polys"] and got["comm_evals"] == want["comm_evals"]
            assert all(np.array_equal(a, b) for a, b in zip(r_got, r_want))
            # and the proof is one the verifier accepts for that claim
            tv = P.Transcript(b"live")
            ok = P.zk_sumcheck_verify(got, P.commit1(claim, O.ZERO, gens.gens_1).compress(), nx + nq, gens.gens_1, gens.gens_4, tv)
            assert ok is not None and ok is not False
            rs = np.stack(r_got)
        else:
            rs = parallel.follow_rounds(sh, comm)
        allr = comm.all_gather(rs)
        assert all(np.array_equal(allr[0], allr[k]) for k in range(world)), "every rank ends with the leader's challenges"
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,Q_local,X", [(2, 2, 4), (4, 1, 4)])
def test_live_transcript_over_sharded_prover(world, Q_local, X):
    mp.spawn(_live_worker, args=(world, _free_port(), Q_local, X), nprocs=world, join=True)


@pytest.mark.parametrize("world,Q_local,X", [(2, 4, 8), (4, 2, 4), (8, 1, 2)])
def test_host_tail_matches_unsharded(world, Q_local, X):
    """the cross-rank rounds on the host (parallel.HostTail over spg_sc1_host_tail_eval / _bind)"""
    mp.spawn(_worker, args=(world, _free_port(), Q_local, X, False, True), nprocs=world, join=True)


def test_shared_memory_mailbox():
    mp.spawn(_worker, args=(2, _free_port(), 2, 8, True), nprocs=2, join=True)
