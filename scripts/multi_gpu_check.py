"""Multi-GPU parity check (run under torchrun, one rank per GPU): the sharded phase-1
sumcheck and the sharded Z-bind on real devices against the unsharded oracle.
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 scripts/multi_gpu_check.py"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spartan_parallel_b200 as sp
from oracle import cbind as O
from oracle import r1cs as R
from spartan_parallel_b200 import parallel
from tests.helpers import log2, rand_scalars


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = sp.Context(local)
    comm = parallel.TorchComm(device=torch.device("cuda", local))
    shm = parallel.ShmComm(device=torch.device("cuda", local))
    for log_x in (7, 9):
        check(ctx, comm, shm, rank, world, log_x)
    if "--full" in sys.argv:
        full_size(ctx, shm, rank, world)
    if "--c4" in sys.argv:
        c4_rows(ctx, shm, rank, world)
    if "--live" in sys.argv:
        live_transcript(ctx, shm, rank, world)
    dist.barrier()
    if rank == 0:
        print(f"multi-GPU parity ok: world={world}, every round bit-exact on every rank (per-round driver and C round loop)")
    shm.close()
    dist.destroy_process_group()


def full_size(ctx, shm, rank, world):
    """BASELINE config C3 (X = 2^16, Q = 256) sharded over all ranks against the SAME batch proven
    unsharded on every rank's own GPU: every round and both claim vectors must be bit-identical."""
    log_x, Q = 16, 256
    X, Ql = 1 << log_x, 256 // world
    nq = log2(Q)
    rng = np.random.default_rng(77)  # same stream on every rank

    def canon(n):
        a = rng.integers(0, 1 << 64, size=(n, 4), dtype=np.uint64)
        a[:, 3] &= np.uint64((1 << 60) - 1)
        return a

    u = canon(X * Q)
    du = sp.DensePolynomial.new(ctx, u)
    dun = sp.DensePolynomial.new(ctx, np.ascontiguousarray(np.roll(u.reshape(Q, X, 4), -1, axis=1).reshape(X * Q, 4)))
    v = sp.vec_op(ctx, "mul", du, dun).to_host()
    del du, dun
    rows = np.arange(X, dtype=np.uint32)
    ones = np.tile(O.ONE, (X, 1))
    inst = sp.R1CSInstance(ctx, 1, X, [X], 2 * X, [(rows, rows, ones)], [(rows, ((rows + 1) % X).astype(np.uint32), ones)],
                           [(rows, (rows + X).astype(np.uint32), ones)])
    tau_q, tau_x = canon(nq), canon(log_x)
    ch1, ch2, r_abc = canon(log_x + nq), canon(1 + log_x), canon(3)
    none = canon(1)[:0]
    rx = ch1[:log_x][::-1].copy()
    # unsharded
    secs = [sp.ProverWitnessSecInfo(ctx, [Q], [X], u), sp.ProverWitnessSecInfo(ctx, [Q], [X], v)]
    z = sp.ZMat(ctx, [Q], [X], secs)
    sc1 = sp.sumcheck_phase1(ctx, inst, z, [Q], Q, [X], X, X, none, tau_q, tau_x)
    want1 = sc1.run_rounds(ch1)
    wantc1 = sc1.final()
    sc1.free()
    sc2 = sp.SumcheckPhase2(ctx, inst, z, [Q], Q, [X], X, 2, rx, ch1[log_x:], none, *r_abc)
    want2 = sc2.run_rounds(ch2)
    wantc2 = sc2.final()
    sc2.free()
    z.free()
    for s_ in secs:
        s_.free()
    # sharded: this rank's slice of the proofs
    lo, hi = rank * Ql * X, (rank + 1) * Ql * X
    secs = [sp.ProverWitnessSecInfo(ctx, [Ql], [X], u[lo:hi]), sp.ProverWitnessSecInfo(ctx, [Ql], [X], v[lo:hi])]
    z = sp.ZMat(ctx, [Ql], [X], secs)
    peer = parallel.PeerTable(ctx, shm, 2 * X)
    for satisfied in (False, True):
        sh = parallel.gpu_phase1(ctx, shm, inst, z, Ql, X, X, tau_q, tau_x, satisfied=satisfied)
        got1 = sh.run_rounds(ch1)
        assert np.array_equal(got1, want1), f"rank {rank}: sharded C3 phase 1 differs (satisfied={satisfied})"
        assert np.array_equal(sh.final(), wantc1), f"rank {rank}: sharded C3 claims differ"
        zrq = parallel.gpu_bind_rq_sharded(ctx, shm, z, ch1[log_x:], Ql, peer)
        sc2 = sp.SumcheckPhase2.from_zrq(ctx, inst, zrq, [X], X, 2, rx, none, *r_abc)
        assert np.array_equal(sc2.run_rounds(ch2), want2), f"rank {rank}: sharded C3 phase 2 differs"
        assert np.array_equal(sc2.final(), wantc2)
        sc2.free()
    peer.close()
    if rank == 0:
        print(f"C3 (2^16 x 256) sharded over {world} GPUs == unsharded: every round and claim bit-identical")


def c4_rows(ctx, shm, rank, world):
    """BASELINE config C4 at full size (P = 5 instances of 2^12 constraints, Q_p = {64,16,16,4,1}, W = 5
    sections with a single one, ragged inputs) with its (instance, proof) rows spread over the ranks
    (parallel.ShardedRows) against the oracle's unsharded run: every round of both sumchecks and all
    final claims, on every rank."""
    from tests.helpers import random_instance, random_witness_secs

    P, W, Ymax = 5, 5, 1 << 12
    num_proofs = [64, 16, 16, 4, 1]
    num_cons = [1 << 12] * P
    Y = [1 << 12, 1 << 12, 1 << 11, 1 << 12, 1 << 11]
    inst = random_instance(P, num_cons, W, Ymax, Y, nnz=3 << 12, seed=204)
    kinds = ["full", "single", "full", "full", "full"]
    sec_inputs = [Y, [1 << 10] * P, [1 << 12, 1 << 13, 1 << 11, 1 << 12, 1 << 10], [8] * P, [8] * P]
    secs = random_witness_secs(P, num_proofs, W, sec_inputs, kinds, seed=205)
    np_, nq, nx, ny, nw = 3, 6, 12, 12, 3
    big = rand_scalars(64, 2040)
    tau_p, tau_q, tau_x = big[:np_], big[8:8 + nq], big[16:16 + nx]
    ch1, ch2, r_abc = rand_scalars(np_ + nq + nx, 2041), rand_scalars(np_ + nw + ny, 2042), rand_scalars(3, 2043)
    want = R.prove_tables(inst, P, 64, num_proofs, Ymax, Y, secs, tau_p, tau_q, tau_x, ch1, r_abc, ch2)
    mats = ([inst.mats[3 * i] for i in range(P)], [inst.mats[3 * i + 1] for i in range(P)], [inst.mats[3 * i + 2] for i in range(P)])
    sh, z_local, mine = parallel.gpu_phase1_rows(ctx, shm, mats, num_cons, 1 << 12, inst.num_vars, secs, num_proofs, Y, Ymax,
                                                 tau_p, tau_q, tau_x)
    got1 = sh.run_rounds(ch1)
    assert np.array_equal(got1, np.stack(want.evals1)), f"rank {rank}: row-sharded C4 phase 1 differs"
    assert np.array_equal(sh.final(), want.claims1), f"rank {rank}: row-sharded C4 phase-1 claims differ"
    peer = parallel.PeerTable(ctx, shm, W * sum(Y))
    zrq = parallel.gpu_bind_rq_rows(ctx, shm, z_local, mine, ch1[nx:nx + nq], Y, W, peer)
    dinst = sp.R1CSInstance(ctx, P, 1 << 12, num_cons, inst.num_vars, *mats)
    rx = ch1[:nx][::-1].copy()
    sc2 = sp.SumcheckPhase2.from_zrq(ctx, dinst, zrq, Y, Ymax, W, rx, ch1[nx + nq:], *r_abc)
    got2 = sc2.run_rounds(ch2)
    assert np.array_equal(got2, np.stack(want.evals2)), f"rank {rank}: row-sharded C4 phase 2 differs"
    assert np.array_equal(sc2.final(), want.claims2), f"rank {rank}: row-sharded C4 phase-2 claims differ"
    sc2.free()
    sh.free()
    peer.close()
    if rank == 0:
        print(f"C4 (P = 5, Q_p = {num_proofs}, 2^12 constraints, W = 5) row-sharded over {world} GPUs == oracle: "
              f"{len(got1)} + {len(got2)} rounds and all claims bit-identical; rows per rank "
              f"{[sum(c for _, _, c in b) for b in sh.blocks]}")


def live_transcript(ctx, shm, rank, world):
    """A live Fiat-Shamir transcript over the sharded device prover: rank 0 runs the ZK sumcheck glue
    (merlin transcript, RandomTape, per-round commitments and dot-product proofs; the oracle's python
    restatement stands in for the Rust host) and publishes each challenge through the mailbox
    (parallel.LeaderRounds); the other ranks follow (parallel.follow_rounds). The phase-1 proof must be
    the one the unsharded oracle prover emits."""
    from oracle import protocol as P

    X, Ql = 1 << 10, 4
    Q = Ql * world
    nx, nq = log2(X), log2(Q)
    inst = R.synthetic_instance(X)
    secs = R.synthetic_witness(X, [Q], seed=21)
    big = rand_scalars(64, 22)
    tau_q, tau_x = big[:nq], big[16:16 + nx]
    A, B, Cm = inst.mats
    dinst = sp.R1CSInstance(ctx, 1, X, [X], 2 * X, [A], [B], [Cm])
    dsecs = [sp.ProverWitnessSecInfo(ctx, [Ql], [X], np.concatenate(ws.w_mat[0][rank * Ql:(rank + 1) * Ql])) for ws in secs]
    z = sp.ZMat(ctx, [Ql], [X], dsecs)
    sh = parallel.gpu_phase1(ctx, shm, dinst, z, Ql, X, X, tau_q, tau_x, satisfied=True)
    gens = P.R1CSGens(b"gens_live", 16)
    seed = rand_scalars(1, 23)[0]
    if rank == 0:
        t, tape = P.Transcript(b"live"), P.RandomTape(b"proof", seed)
        got, r_got, _ = P.zk_sumcheck_prove(O.ZERO, O.ZERO, nx + nq, parallel.LeaderRounds(sh, shm), gens.gens_1, gens.gens_4, t, tape)
        z_mat = R.build_z_mat(1, [Q], [X], secs)
        Az, Bz, Cz = R.multiply_vec_block(inst, 1, [Q], X, [X], z_mat)
        mk = lambda T: O.Pqx.new_rev(T, 1, [Q], Q, [X], X)
        sc = O.Sc1(nx, nq, 0, [Q], [X], O.ONE.reshape(1, 4), O.eq_evals(tau_q), O.eq_evals(tau_x), mk(Az), mk(Bz), mk(Cz))
        t2, tape2 = P.Transcript(b"live"), P.RandomTape(b"proof", seed)
        want, r_want, _ = P.zk_sumcheck_prove(O.ZERO, O.ZERO, nx + nq, P._Sc1Engine(sc), gens.gens_1, gens.gens_4, t2, tape2)
        assert got["comm_polys"] == want["comm_polys"] and got["comm_evals"] == want["comm_evals"], "live sharded proof differs"
        assert all(np.array_equal(a, b) for a, b in zip(r_got, r_want))
        rs = np.stack(r_got)
        print(f"live transcript on rank 0 over {world} sharded device provers: phase-1 ZK sumcheck proof ({nx + nq} rounds) "
              "byte-identical to the unsharded oracle's")
    else:
        rs = parallel.follow_rounds(sh, shm)
    allr = parallel.TorchComm.all_gather(shm, rs)
    assert all(np.array_equal(allr[0], allr[k]) for k in range(world))
    sh.free()


def check(ctx, comm, shm, rank, world, log_x):
    X, Ql = 1 << log_x, 4
    Q = Ql * world
    nx, nq = log2(X), log2(Q)
    inst = R.synthetic_instance(X, unit=False, seed=9)
    secs = R.synthetic_witness(X, [Q], seed=10)  # the whole batch; this rank uploads its slice
    big = rand_scalars(64, 11)
    tau_q, tau_x = big[:nq], big[16:16 + nx]
    ch1, ch2, r_abc = rand_scalars(nx + nq, 12), rand_scalars(1 + nx, 13), rand_scalars(3, 14)
    want = R.prove_tables(inst, 1, Q, [Q], X, [X], secs, big[:0], tau_q, tau_x, ch1, r_abc, ch2)

    A, B, Cm = inst.mats
    dinst = sp.R1CSInstance(ctx, 1, X, [X], 2 * X, [A], [B], [Cm])
    dsecs = []
    for ws in secs:
        mine = np.concatenate(ws.w_mat[0][rank * Ql:(rank + 1) * Ql])
        dsecs.append(sp.ProverWitnessSecInfo(ctx, [Ql], [X], mine))
    z = sp.ZMat(ctx, [Ql], [X], dsecs)
    sc1 = parallel.gpu_phase1(ctx, comm, dinst, z, Ql, X, X, tau_q, tau_x)
    for j in range(sc1.num_rounds):
        got = sc1.round_eval()
        assert np.array_equal(got, want.evals1[j]), f"rank {rank}: phase-1 round {j} differs"
        sc1.round_bind(ch1[j])
    assert np.array_equal(sc1.final(), want.claims1), f"rank {rank}: phase-1 claims differ"
    rx = ch1[:nx][::-1].copy()
    zrq = parallel.gpu_bind_rq_sharded(ctx, comm, z, ch1[nx:nx + nq], Ql)
    sc2 = sp.SumcheckPhase2.from_zrq(ctx, dinst, zrq, [X], X, 2, rx, ch1[:0], *r_abc)
    for j in range(sc2.num_rounds):
        got = sc2.round_eval()
        assert np.array_equal(got, want.evals2[j]), f"rank {rank}: phase-2 round {j} differs"
        sc2.round_bind(ch2[j])
    assert np.array_equal(sc2.final(), want.claims2), f"rank {rank}: phase-2 claims differ"
    # the same phase 1 through the shared-memory mailbox and the C round loop
    sc1 = parallel.gpu_phase1(ctx, shm, dinst, z, Ql, X, X, tau_q, tau_x, satisfied=True)
    got = sc1.run_rounds(ch1)
    assert np.array_equal(got, np.stack(want.evals1)), f"rank {rank}: C round loop differs"
    assert np.array_equal(sc1.final(), want.claims1), f"rank {rank}: C round loop claims differ"
    # the Z table reduced over peer memory instead of NCCL + additions
    peer = parallel.PeerTable(ctx, shm, 2 * X)
    for _ in range(2):  # twice: the table is reused across proofs
        zrq2 = parallel.gpu_bind_rq_sharded(ctx, shm, z, ch1[nx:nx + nq], Ql, peer)
        assert np.array_equal(zrq2.to_host(), zrq.to_host()), f"rank {rank}: peer-memory reduction differs"
    sc2 = sp.SumcheckPhase2.from_zrq(ctx, dinst, zrq2, [X], X, 2, rx, ch1[:0], *r_abc)
    assert np.array_equal(sc2.run_rounds(ch2), np.stack(want.evals2)), f"rank {rank}: phase 2 on the peer table differs"
    sc2.free()
    peer.close()


if __name__ == "__main__":
    main()
