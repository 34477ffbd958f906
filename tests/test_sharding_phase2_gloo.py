"""Multi-process (gloo, CPU) test of the y-sharded phase-2 sumcheck (parallel.ShardedPhase2 + the
library's host-side tail spg_sc2_host_tail_*): with the oracle standing in for the per-rank device
engine on its chunk of the flat [w][y] tables, every round polynomial and the final claims must equal
the unsharded oracle's (src/sumcheck.rs:573-743 over one instance)."""
import os

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import cbind as O
from tests.helpers import log2, rand_scalars
from tests.test_sharding_gloo import ONE1, _free_port


def mk_sc2(W, Y, abc, z):
    mk = lambda T: O.Pqx.new_rev(T, W, [1], 1, [Y], Y)
    return O.Sc2(log2(Y), log2(W), 0, True, W, [Y], ONE1, mk(abc), mk(z))


def _worker(rank, world, port, W, Y, use_shm):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from spartan_parallel_b200 import parallel

        abc, z = rand_scalars(W * Y, 11), rand_scalars(W * Y, 12)
        ch = rand_scalars(log2(W * Y), 13)
        full = mk_sc2(W, Y, abc, z)
        want = []
        for j in range(full.num_rounds):
            want.append(full.round_eval())
            full.round_bind(ch[j])
        want_final = full.final()
        comm = parallel.ShmComm() if use_shm else parallel.TorchComm()
        make_engine = lambda off, n: mk_sc2(1, n, abc[off:off + n], z[off:off + n])
        sh = parallel.ShardedPhase2(comm, W, Y, make_engine)
        assert sh.num_rounds == full.num_rounds and sh.flat_len == W * Y // world
        for j in range(sh.num_rounds):
            assert np.array_equal(sh.round_eval(), want[j]), f"rank {rank} round {j}"
            sh.round_bind(ch[j])
        assert np.array_equal(sh.final(), want_final), f"rank {rank} final claims"
        sh2 = parallel.ShardedPhase2(comm, W, Y, make_engine)
        assert np.array_equal(sh2.run_rounds(ch), np.stack(want))
        assert np.array_equal(sh2.final(), want_final)
        if use_shm:
            comm.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,W,Y,use_shm", [(2, 2, 8, False), (4, 2, 8, False), (4, 4, 4, False), (4, 1, 16, True), (2, 2, 2, True)])
def test_sharded_phase2_matches_unsharded(world, W, Y, use_shm):
    mp.spawn(_worker, args=(world, _free_port(), W, Y, use_shm), nprocs=world, join=True)


def test_phase2_sharding_needs_aligned_chunks():
    from spartan_parallel_b200 import parallel

    class One:
        world, rank = 3, 0

    with pytest.raises(AssertionError):
        parallel.ShardedPhase2(One(), 2, 8, lambda off, n: None)
