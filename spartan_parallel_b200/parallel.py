"""Sharding of the proof axis across GPUs (one process per GPU, torch.distributed).

The protocol-level data parallelism of spartan-parallel is the proof index q: tables are
independent per proof up to the per-round sum. A rank owns the contiguous block of
proofs q = rank * Q_local + q_local, i.e. the HIGH bits of q are the rank. Because the
sumcheck binds q low bit first (the reference stores q bit-reversed and binds the top,
src/custom_dense_mlpoly.rs:83-99, 222-244), the x rounds and the first log2(Q_local) q
rounds never pair scalars of different ranks: each rank runs the unchanged single-GPU
kernels on its shard, scaled by the eq weight of its rank bits, and the only exchange
per round is 3 scalars per rank (all-gather, then Scalar::add on the host -- NCCL has
no modular reduction). After the local rounds every table is one scalar per rank; those
are gathered and the last log2(G) rounds run on a G-entry table.

Phase 2 needs Z bound to rq: sum over all proofs -> every rank binds its shard
(spg_zmat_bind_rq with its rank weight), the partial (W * Y)-scalar tables are
all-gathered once and added, and phase 2 (independent of Q) runs replicated.

The per-rank engine is injected so that the CPU tests can drive the same logic with the
oracle standing in for the device (tests/test_sharding_gloo.py).
"""
from __future__ import annotations

import os

import numpy as np

from . import api


def log2(n: int) -> int:
    return n.bit_length() - 1


def bind_host_to_gpu(device: int) -> dict:
    """Pins this process to the CPUs of the NUMA node the GPU hangs off, BEFORE any pinned host
    buffer is allocated: pinned pages are first-touch, so the staging buffers of the end-to-end
    path then sit in the memory local to the GPU's PCIe root instead of crossing the socket
    interconnect (8 ranks copying 4 GiB per batch each otherwise share one socket's links).
    Best effort: returns what was done, never raises."""
    info = {"device": device, "numa_node": None, "cpus": None}
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(device)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
        if len(bus.split(":")[0]) == 8:  # nvml prints an 8-digit PCI domain, sysfs a 4-digit one
            bus = bus[4:]
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        info["numa_node"] = node
        if node < 0:
            return info
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            info["cpus"] = len(cpus)
    except Exception as e:  # no NVML, no sysfs entry, container without the topology: leave the affinity alone
        info["error"] = str(e)[:120]
    return info


class TorchComm:
    """all_gather of small uint64 arrays over torch.distributed (nccl or gloo)."""

    def __init__(self, device=None):
        import torch.distributed as dist

        self.dist = dist
        self.rank = dist.get_rank()
        self.world = dist.get_world_size()
        self.device = device

    def all_gather(self, arr: np.ndarray) -> np.ndarray:
        import torch

        t = torch.from_numpy(np.ascontiguousarray(arr).view(np.int64).copy())
        if self.device is not None:
            t = t.to(self.device)
        outs = [torch.empty_like(t) for _ in range(self.world)]
        self.dist.all_gather(outs, t)
        return np.stack([o.cpu().numpy().view(np.uint64).reshape(arr.shape) for o in outs])

    def all_gather_device(self, tensor):
        import torch

        outs = [torch.empty_like(tensor) for _ in range(self.world)]
        self.dist.all_gather(outs, tensor)
        return outs

    def barrier(self):
        self.all_gather(np.zeros(1, dtype=np.uint64))


class ShmComm(TorchComm):
    """Per-round exchange through a POSIX shared-memory mailbox.

    The 3 round evaluations land in pinned HOST memory on every rank (the host needs them
    for the transcript), and all ranks of one box share that host: exchanging 96 bytes per
    rank through shared memory costs a few microseconds, against ~100 us for staging them
    back to the device for an NCCL all-gather. Bulk device tables still go over NCCL
    (all_gather_device). Slots are double-buffered by call parity: a rank can run at most
    one call ahead of the slowest reader."""

    SLOT = 4096

    def __init__(self, device=None):
        super().__init__(device)
        from multiprocessing import shared_memory

        self.calls = 0
        name = f"spg_mbox_{os.environ.get('MASTER_PORT', '0')}_{os.getppid()}"
        size = 2 * self.world * (self.SLOT + 64)
        if self.rank == 0:
            try:
                old = shared_memory.SharedMemory(name=name)
                old.close()
                old.unlink()
            except FileNotFoundError:
                pass
            self.shm = shared_memory.SharedMemory(name=name, create=True, size=size)
            self.shm.buf[:size] = bytes(size)
        self.dist.barrier()
        if self.rank != 0:
            self.shm = shared_memory.SharedMemory(name=name)
            try:  # rank 0 owns the segment; keep this process's resource tracker from unlinking it again
                from multiprocessing import resource_tracker

                resource_tracker.unregister(self.shm._name, "shared_memory")
            except Exception:
                pass
        self.dist.barrier()
        words = np.ndarray((2, self.world, (self.SLOT + 64) // 8), dtype=np.uint64, buffer=self.shm.buf)
        self.seq = words[:, :, 0]
        self.data = words[:, :, 8:]
        self.addr = words.ctypes.data  # for the C round loop (spg_sc1_run_rounds_sharded)
        self.slot_stride = self.SLOT + 64

    def all_gather(self, arr: np.ndarray) -> np.ndarray:
        """Through the C helper (spg_mailbox_all_gather): the same release / acquire ordering as the
        C round loop (plain numpy stores are ordered only on x86), a deadline on every wait, and a
        poisoned slot when a rank fails so that its peers raise instead of spinning forever."""
        import ctypes as C

        from ._lib import check, lib

        a = np.ascontiguousarray(arr, dtype=np.uint64).reshape(-1)
        assert a.size * 8 <= self.SLOT
        calls = C.c_uint64(self.calls)
        out = np.empty((self.world, a.size), dtype=np.uint64)
        rc = lib().spg_mailbox_all_gather(C.c_void_p(self.addr), self.slot_stride, self.rank, self.world, C.byref(calls),
                                          a.ctypes.data_as(C.c_void_p), a.size * 8, out.ctypes.data_as(C.c_void_p))
        self.calls = int(calls.value)
        check(rc, "spg_mailbox_all_gather")
        return out.reshape((self.world,) + tuple(np.shape(arr)))

    def poison(self):
        """Called by a rank that is about to fail: peers blocked in an exchange return an error."""
        from ._lib import lib

        lib().spg_mailbox_poison(self.addr, self.slot_stride, self.rank, self.world)

    def close(self):
        try:
            self.dist.barrier()
            self.shm.close()
            if self.rank == 0:
                self.shm.unlink()
        except Exception:
            pass


class LocalComm:
    """Single-process stand-in (world = 1)."""

    rank, world = 0, 1

    def all_gather(self, arr):
        return np.stack([arr])

    def barrier(self):
        pass


class PeerTable:
    """One table of n scalars per rank, all mapped into every rank's address space through
    CUDA IPC (spg_peer_alloc / spg_peer_open): the modular all-reduce between the two phases
    of a sharded proof runs as one kernel per rank over NVLink peer memory (spg_peer_sum)
    instead of an NCCL all-gather of whole tables followed by separate additions.
    Create once, reuse for every proof."""

    def __init__(self, ctx, comm, n: int):
        import ctypes as C

        from ._lib import check

        self.ctx, self.comm, self.n = ctx, comm, n
        handle = np.zeros(64, dtype=np.uint8)
        h = C.c_void_p()
        check(ctx.L.spg_peer_alloc(ctx.h, n, C.byref(h), handle.ctypes.data_as(C.c_void_p)), "spg_peer_alloc")
        self._vec = h
        self.device_ptr = int(ctx.L.spg_vec_device_ptr(h))
        self.poly = api.DensePolynomial.wrap(ctx, self.device_ptr, n, owner=self)
        handles = comm.all_gather(handle.view(np.uint64))  # (world, 8)
        self._opened = []
        ptrs = []
        for r in range(comm.world):
            if r == comm.rank:
                ptrs.append(self.device_ptr)
                continue
            hr = np.ascontiguousarray(handles[r]).view(np.uint8)
            pp = C.c_void_p()
            check(ctx.L.spg_peer_open(ctx.h, hr.ctypes.data_as(C.c_void_p), C.byref(pp)), "spg_peer_open")
            self._opened.append(pp)
            ptrs.append(pp.value)
        self._ptrs = (C.c_void_p * comm.world)(*ptrs)
        comm.barrier()

    def all_reduce(self):
        """Every rank has written its partial table into self.poly (on the context's stream)."""
        from ._lib import check

        self.ctx.sync()
        self.comm.barrier()  # all partial tables are complete and visible
        check(self.ctx.L.spg_peer_sum(self.ctx.h, self._ptrs, self.comm.world, self.comm.rank, self.n), "spg_peer_sum")
        self.ctx.sync()
        self.comm.barrier()  # every chunk has been written everywhere
        return self.poly

    def reduce_scatter(self):
        """like all_reduce, but rank r only receives the sums of chunk r (entries [r * n / G, (r + 1) * n / G)
        of its own table): half the NVLink traffic, for a consumer sharded the same way (ShardedPhase2)"""
        from ._lib import check

        self.ctx.sync()
        self.comm.barrier()
        check(self.ctx.L.spg_peer_reduce_scatter(self.ctx.h, self._ptrs, self.comm.world, self.comm.rank, self.n), "spg_peer_reduce_scatter")
        self.ctx.sync()
        self.comm.barrier()  # peers have finished reading this rank's table before it is overwritten again
        return self.poly

    def close(self):
        if getattr(self, "_vec", None) is None:
            return
        self.comm.barrier()  # nobody is still reading a peer's table
        for pp in self._opened:
            self.ctx.L.spg_peer_close(pp)
        self._opened = []
        self.comm.barrier()
        self.ctx.L.spg_peer_free(self._vec)
        self._vec = None


class ShardedPhase1:
    """Phase-1 sumcheck of one instance whose Q = Q_local * world proofs are sharded.

    ``make_engine(tau_q_local)`` returns the rank-local prover (an object with
    round_eval / round_bind / final / set_scale: ``api.SumcheckPhase1`` on a GPU);
    ``make_tail(Az, Bz, Cz, tau_q_high)`` builds the prover for the gathered G-entry
    tables."""

    def __init__(self, comm, Q_local: int, X: int, tau_q, tau_x, make_engine, make_tail):
        self.comm = comm
        G = comm.world
        assert G & (G - 1) == 0, "world size must be a power of two"
        self.nx, self.nql, self.ng = log2(X), log2(Q_local), log2(G)
        tau_q = np.asarray(tau_q, dtype=np.uint64).reshape(-1, 4)
        assert tau_q.shape[0] == self.nql + self.ng
        self.tau_q_high = tau_q[self.nql:]
        self.engine = make_engine(tau_q[: self.nql])
        self.engine.set_scale(api.host_eq_weight(self.tau_q_high, comm.rank))
        self.make_tail = make_tail
        self.tail = None
        self.round = 0
        self.num_rounds = self.nx + self.nql + self.ng

    def round_eval(self) -> np.ndarray:
        if self.round < self.nx + self.nql:
            part = self.engine.round_eval()
            return api.host_sum(self.comm.all_gather(part))
        return self._tail().round_eval()

    def round_bind(self, r):
        if self.round < self.nx + self.nql:
            self.engine.round_bind(r)
        else:
            self._tail().round_bind(r)
        self.round += 1

    def run_rounds(self, challenges) -> np.ndarray:
        """All rounds with challenges known in advance. On a GPU engine with the shared-memory
        mailbox the local rounds run in one C loop (no Python between rounds)."""
        ch = np.asarray(challenges, dtype=np.uint64).reshape(-1, 4)
        out = []
        n_local = self.nx + self.nql
        if self.round == 0 and hasattr(self.engine, "run_rounds_sharded") and isinstance(self.comm, ShmComm):
            calls = np.array([self.comm.calls], dtype=np.uint64)
            ev = self.engine.run_rounds_sharded(ch[:n_local], self.comm.addr, self.comm.slot_stride, self.comm.rank,
                                                self.comm.world, calls)
            self.comm.calls = int(calls[0])
            out.extend(ev)
            self.round = n_local
        while self.round < self.num_rounds:
            j = self.round
            out.append(self.round_eval())
            self.round_bind(ch[j])
        return np.stack(out)

    def _tail(self):
        if self.tail is None:
            # one scalar per table per rank; the bound eq products are identical on all ranks
            self.engine.set_scale(api.ONE)
            claims = self.engine.final()
            allc = self.comm.all_gather(claims)  # (G, 4, 4)
            self.tail = self.make_tail(allc[:, 1].copy(), allc[:, 2].copy(), allc[:, 3].copy(), self.tau_q_high)
            self.tail.set_scale(claims[0])
        return self.tail

    def final(self) -> np.ndarray:
        return self._tail().final()

    def free(self):
        """Releases the shard prover and the tail prover now (not when the object is collected):
        a lingering tail pins small pieces of the freed table memory until the NEXT proof has
        already allocated its tables."""
        for eng in (self.engine, self.tail):
            if eng is not None and hasattr(eng, "free"):
                eng.free()
        self.engine = self.tail = None


class HostTail:
    """The last log2(G) rounds of a sharded proof on G gathered scalars per table, on the host through
    spg_sc1_host_tail_eval / _bind (see include/spgpu.h): same interface as the device prover."""

    def __init__(self, Az, Bz, Cz, tau_high):
        import ctypes as C

        from ._lib import lib

        self.L, self.C = lib(), C
        G = Az.shape[0]
        self.G, self.len = G, G
        tau_high = np.asarray(tau_high, dtype=np.uint64).reshape(-1, 4)
        E = np.stack([api.host_eq_weight(tau_high, i) for i in range(G)])
        self.state = np.ascontiguousarray(np.concatenate([E, Az, Bz, Cz]).astype(np.uint64))
        self.scale = api.ONE.copy()

    def set_scale(self, c):
        self.scale = np.ascontiguousarray(np.asarray(c, dtype=np.uint64).reshape(4))

    def round_eval(self):
        from ._lib import check

        e = np.empty((3, 4), dtype=np.uint64)
        p = lambda a: a.ctypes.data_as(self.C.c_void_p)
        check(self.L.spg_sc1_host_tail_eval(p(self.state), self.G, self.len, p(self.scale), p(e)), "spg_sc1_host_tail_eval")
        return e

    def round_bind(self, r):
        from ._lib import check

        r = np.ascontiguousarray(np.asarray(r, dtype=np.uint64).reshape(4))
        p = lambda a: a.ctypes.data_as(self.C.c_void_p)
        check(self.L.spg_sc1_host_tail_bind(p(self.state), self.G, self.len, p(r)), "spg_sc1_host_tail_bind")
        self.len //= 2

    def final(self):
        s = self.state.reshape(4, self.G, 4)
        return np.stack([api.host_mul(self.scale, s[0, 0]), s[1, 0], s[2, 0], s[3, 0]])

    def free(self):
        pass


class HostTail2:
    """The last log2(G) rounds of a y-sharded phase 2 on the gathered scalars (spg_sc2_host_tail_*): first
    the remaining y rounds (adjacent pairs), then the w rounds (top bit first)."""

    def __init__(self, B, C_, n_y_rounds: int):
        import ctypes as C

        from ._lib import lib

        self.L, self.C = lib(), C
        self.G = self.len = B.shape[0]
        self.state = np.ascontiguousarray(np.concatenate([B, C_]).astype(np.uint64))
        self.scale = api.ONE.copy()
        self.n_y, self.done = n_y_rounds, 0

    def set_scale(self, c):
        self.scale = np.ascontiguousarray(np.asarray(c, dtype=np.uint64).reshape(4))

    def round_eval(self):
        from ._lib import check

        e = np.empty((3, 4), dtype=np.uint64)
        p = lambda a: a.ctypes.data_as(self.C.c_void_p)
        check(self.L.spg_sc2_host_tail_eval(p(self.state), self.G, self.len, int(self.done >= self.n_y), p(self.scale), p(e)),
              "spg_sc2_host_tail_eval")
        return e

    def round_bind(self, r):
        from ._lib import check

        r = np.ascontiguousarray(np.asarray(r, dtype=np.uint64).reshape(4))
        p = lambda a: a.ctypes.data_as(self.C.c_void_p)
        check(self.L.spg_sc2_host_tail_bind(p(self.state), self.G, self.len, int(self.done >= self.n_y), p(r)), "spg_sc2_host_tail_bind")
        self.len //= 2
        self.done += 1

    def final(self):
        s = self.state.reshape(2, self.G, 4)
        return np.stack([self.scale, s[0, 0], s[1, 0]])


class ShardedPhase2:
    """Phase-2 sumcheck of one instance sharded over y: rank r owns chunk r of the flat [w][y] tables
    (W sections of Y inputs, W a power of two dividing the world size). The summand eq_p * ABC * Z has no
    weight over (w, y), so the ranks' partial round evaluations add; after the log2(chunk) local rounds
    the per-rank scalars are gathered and the last log2(G) rounds run on the host (HostTail2).

    ``make_engine(flat_off, flat_len)`` returns the rank-local prover (round_eval / round_bind / final;
    ``api.SumcheckPhase2.slice`` on a GPU)."""

    def __init__(self, comm, W: int, Y: int, make_engine):
        self.comm = comm
        G = comm.world
        assert W & (W - 1) == 0 and G % W == 0 and (W * Y) % G == 0, "y-sharding needs W | G | W * Y"
        self.flat_len = W * Y // G
        self.n_local = log2(self.flat_len)
        self.n_y_tail = log2(Y) - self.n_local
        self.num_rounds = log2(Y) + log2(W)
        self.engine = make_engine(comm.rank * self.flat_len, self.flat_len)
        self.tail = None
        self.round = 0

    def round_eval(self) -> np.ndarray:
        if self.round < self.n_local:
            return api.host_sum(self.comm.all_gather(self.engine.round_eval()))
        return self._tail().round_eval()

    def round_bind(self, r):
        if self.round < self.n_local:
            self.engine.round_bind(r)
        else:
            self._tail().round_bind(r)
        self.round += 1

    def run_rounds(self, challenges) -> np.ndarray:
        ch = np.asarray(challenges, dtype=np.uint64).reshape(-1, 4)
        out = []
        if self.round == 0 and self.n_local and hasattr(self.engine, "run_rounds_sharded") and isinstance(self.comm, ShmComm):
            calls = np.array([self.comm.calls], dtype=np.uint64)
            out.extend(self.engine.run_rounds_sharded(ch[: self.n_local], self.comm.addr, self.comm.slot_stride, self.comm.rank,
                                                      self.comm.world, calls))
            self.comm.calls = int(calls[0])
            self.round = self.n_local
        while self.round < self.num_rounds:
            j = self.round
            out.append(self.round_eval())
            self.round_bind(ch[j])
        return np.stack(out)

    def _tail(self):
        if self.tail is None:
            claims = np.asarray(self.engine.final(), dtype=np.uint64).reshape(3, 4)
            allc = self.comm.all_gather(claims)  # (G, 3, 4), rank order = flat (w, y_high) order
            self.tail = HostTail2(allc[:, 1].copy(), allc[:, 2].copy(), self.n_y_tail)
            self.tail.set_scale(claims[0])  # eq_p bound so far (one instance: 1)
        return self.tail

    def final(self) -> np.ndarray:
        return self._tail().final()

    def free(self):
        if self.engine is not None and hasattr(self.engine, "free"):
            self.engine.free()
        self.engine = None


def gpu_phase2_sharded(ctx, comm, inst, zrq, max_num_inputs: int, W: int, rx, r_A, r_B, r_C) -> ShardedPhase2:
    """ShardedPhase2 on this rank's GPU; zrq holds this rank's reduce-scattered chunk in place
    (PeerTable.reduce_scatter)."""
    return ShardedPhase2(comm, W, max_num_inputs,
                         lambda off, n: api.SumcheckPhase2.slice(ctx, inst, zrq, max_num_inputs, W, off, n, rx, r_A, r_B, r_C))


def gpu_phase1(ctx, comm, inst, z, Q_local, X, max_num_inputs, tau_q, tau_x, satisfied: bool = False) -> ShardedPhase1:
    """ShardedPhase1 on this rank's GPU. satisfied=True asserts that the witness satisfies the
    instance: Az*Bz - Cz then vanishes entry by entry, so every shard's own sum is zero and the
    shard may take e(1) from that claim and skip e(0) in the first round (spg_sc1_set_satisfied) exactly like
    the unsharded prover."""
    empty = np.zeros((0, 4), dtype=np.uint64)

    def make_engine(tau_q_local):
        sc = api.sumcheck_phase1(ctx, inst, z, [Q_local], Q_local, [X], X, max_num_inputs, empty, tau_q_local, tau_x)
        if satisfied:
            sc.set_satisfied()
        return sc

    def make_tail(Az, Bz, Cz, tau_high):
        G = Az.shape[0]
        if G <= 64 and not os.environ.get("SPG_DEVICE_TAIL"):
            return HostTail(Az, Bz, Cz, tau_high)
        return api.SumcheckPhase1.from_tables(ctx, [G], G, [1], 1, Az, Bz, Cz, empty, tau_high, empty)

    return ShardedPhase1(comm, Q_local, X, tau_q, tau_x, make_engine, make_tail)


def gpu_bind_rq_sharded(ctx, comm, z, rq_rev, Q_local, peer: "PeerTable | None" = None, scatter: bool = False):
    """Z bound to rq over all shards: local bind scaled by the rank's eq weight, then the modular
    sum over ranks -- over peer memory when a PeerTable is given (one kernel per rank, 2 (G-1)/G
    of a table over NVLink), else one NCCL all-gather of the partial tables + device additions."""
    import torch

    rq_rev = np.asarray(rq_rev, dtype=np.uint64).reshape(-1, 4)
    nql = log2(Q_local)
    total = sum(len(z.witness_secs) * y for y in z.num_inputs)
    if peer is not None and comm.world > 1:
        assert peer.n == total
        if isinstance(comm, ShmComm):  # bind, barrier, peer sum, barrier in one C call (spg_zmat_bind_rq_sharded)
            import ctypes as C

            from ._lib import check

            rq = np.ascontiguousarray(rq_rev)
            calls = C.c_uint64(comm.calls)
            rc = ctx.L.spg_zmat_bind_rq_sharded(ctx.h, z.h, rq.ctypes.data_as(C.c_void_p), nql, rq.shape[0], peer._ptrs, comm.world,
                                                comm.rank, total, int(scatter), C.c_void_p(comm.addr), comm.slot_stride, C.byref(calls))
            comm.calls = int(calls.value)
            check(rc, "spg_zmat_bind_rq_sharded")
            return peer.poly
        api.zmat_bind_rq(ctx, z, rq_rev[:nql], api.host_eq_weight(rq_rev[nql:], comm.rank), peer.poly)
        return peer.reduce_scatter() if scatter else peer.all_reduce()
    dev = torch.device("cuda", ctx.device)
    mine = torch.empty((total, 4), dtype=torch.int64, device=dev)
    out = api.DensePolynomial.wrap(ctx, mine.data_ptr(), total, owner=mine)
    api.zmat_bind_rq(ctx, z, rq_rev[:nql], api.host_eq_weight(rq_rev[nql:], comm.rank), out)
    ctx.sync()
    if comm.world == 1:
        return out
    parts = comm.all_gather_device(mine)
    torch.cuda.synchronize()
    acc = api.DensePolynomial.wrap(ctx, parts[0].data_ptr(), total, owner=parts[0])
    for t in parts[1:]:
        nxt = api.DensePolynomial.wrap(ctx, t.data_ptr(), total, owner=t)
        acc2 = api.vec_op(ctx, "add", acc, nxt)
        acc = acc2
    ctx.sync()
    return acc


# ---------------------------------------------------------------------------------------------
# Several instances with different numbers of proofs (BASELINE config C4; src/lib.rs:1156-1270 sorts
# them by num_proofs): the (instance, proof) ROWS of the batch are spread over the ranks. The x rounds
# touch each row on its own, so a rank proves its rows with the global eq weights eq_p[p] * eq_q[q]
# handed to the engine (spg_sc1_set_row_weights); after the x rounds every row is one scalar per
# table, those are gathered, and the q and p rounds (sum_p Q_p entries: negligible) run on every
# rank alike. Unlike ShardedPhase1 nothing ties a rank to a bit pattern of q, so instances with fewer
# proofs than ranks are no special case.
def next_pow2(n: int) -> int:
    return 1 if n <= 1 else 1 << (n - 1).bit_length()


def partition_rows(num_proofs, world: int):
    """blocks[rank] = [(instance p, first proof, count)]: instance p's proofs are cut into
    min(world, Q_p) equal blocks, dealt to the ranks starting at rank p (so that instances with fewer
    proofs than ranks do not pile up on rank 0). Counts are powers of two like the Q_p."""
    blocks = [[] for _ in range(world)]
    for p, Q in enumerate(num_proofs):
        per = max(1, Q // world)
        nblk = Q // per
        stride = max(1, world // nblk)
        for b in range(nblk):
            blocks[(b * stride + p) % world].append((p, b * per, per))
    for r in range(world):
        blocks[r].sort()
    return blocks


def host_eq(a, b) -> np.ndarray:
    """eq(a, b) = a b + (1 - a)(1 - b) for two scalars, through the library's host helpers"""
    one_minus = lambda t: api.host_eq_weight(np.asarray(t, dtype=np.uint64).reshape(1, 4), 0)
    return api.host_sum(np.stack([api.host_mul(a, b), api.host_mul(one_minus(a), one_minus(b))]).reshape(2, 1, 4))[0]


def row_weights(blocks_of_rank, tau_p, tau_q):
    """eq_p[p] * eq_q[q] for the rows of one rank. eq_p is the reference's table order (index MSB <->
    tau_p[0], p is never bit-reversed); eq_q pairs bit k of the natural proof index with tau_q[k]."""
    tau_p = np.asarray(tau_p, dtype=np.uint64).reshape(-1, 4)
    tau_q = np.asarray(tau_q, dtype=np.uint64).reshape(-1, 4)
    tp = tau_p[::-1].copy()
    out = []
    for p, q0, cnt in blocks_of_rank:
        wp = api.host_eq_weight(tp, p)
        for q in range(q0, q0 + cnt):
            out.append(api.host_mul(wp, api.host_eq_weight(tau_q, q)))
    return np.stack(out) if out else np.zeros((0, 4), dtype=np.uint64)


class ShardedRows:
    """Phase-1 sumcheck of a multi-instance batch whose rows are spread over the ranks.

    ``make_engine(blocks, weights)`` returns this rank's prover over its rows (round_eval /
    round_bind / debug_tables; on a GPU ``api.SumcheckPhase1`` with set_row_weights) or None when
    the rank owns no row; ``make_tail(Az, Bz, Cz)`` builds the prover of the q and p rounds from the
    gathered per-row scalars (set_scale / round_eval / round_bind / final)."""

    def __init__(self, comm, num_proofs, max_num_cons: int, tau_p, tau_q, tau_x, make_engine, make_tail):
        self.comm = comm
        self.num_proofs = list(num_proofs)
        P = len(self.num_proofs)
        self.nx, self.nq, self.np_ = log2(max_num_cons), log2(max(self.num_proofs)), log2(next_pow2(P))
        self.tau_x = np.asarray(tau_x, dtype=np.uint64).reshape(-1, 4)
        self.blocks = partition_rows(self.num_proofs, comm.world)
        self.mine = self.blocks[comm.rank]
        self.engine = make_engine(self.mine, row_weights(self.mine, tau_p, tau_q)) if self.mine else None
        self.make_tail = make_tail
        self.tail = None
        self.round = 0
        self.num_rounds = self.nx + self.nq + self.np_
        self.rx = []

    def round_eval(self) -> np.ndarray:
        if self.round < self.nx:
            part = self.engine.round_eval() if self.engine is not None else np.zeros((3, 4), dtype=np.uint64)
            return api.host_sum(self.comm.all_gather(np.asarray(part, dtype=np.uint64).reshape(3, 4)))
        return self._tail().round_eval()

    def round_bind(self, r):
        r = np.asarray(r, dtype=np.uint64).reshape(4)
        if self.round < self.nx:
            if self.engine is not None:
                self.engine.round_bind(r)
            self.rx.append(r)
        else:
            self._tail().round_bind(r)
        self.round += 1

    def run_rounds(self, challenges) -> np.ndarray:
        ch = np.asarray(challenges, dtype=np.uint64).reshape(-1, 4)
        out = []
        if (self.round == 0 and self.nx and self.engine is not None and hasattr(self.engine, "run_rounds_sharded")
                and isinstance(self.comm, ShmComm) and all(self.blocks)):
            # every rank owns rows: the x rounds run in the C loop (eval, mailbox, bind; no Python in between)
            calls = np.array([self.comm.calls], dtype=np.uint64)
            out.extend(self.engine.run_rounds_sharded(ch[: self.nx], self.comm.addr, self.comm.slot_stride, self.comm.rank,
                                                      self.comm.world, calls))
            self.comm.calls = int(calls[0])
            self.rx = [ch[j] for j in range(self.nx)]
            self.round = self.nx
        while self.round < self.num_rounds:
            j = self.round
            out.append(self.round_eval())
            self.round_bind(ch[j])
        return np.stack(out)

    def _tail(self):
        if self.tail is None:
            n_rows = [sum(c for _, _, c in b) for b in self.blocks]
            cap = max(n_rows)
            mine = np.zeros((3, cap, 4), dtype=np.uint64)
            if self.engine is not None:
                tabs = self.engine.debug_tables()
                for k in range(3):
                    mine[k, : n_rows[self.comm.rank]] = np.asarray(tabs[k], dtype=np.uint64).reshape(-1, 4)[: n_rows[self.comm.rank]]
            # one exchange of 3 scalars per row; through torch.distributed: it may exceed a mailbox slot
            allr = TorchComm.all_gather(self.comm, mine) if self.comm.world > 1 else mine[None]
            off, tot = [], 0
            for Q in self.num_proofs:
                off.append(tot)
                tot += Q
            tabs = np.zeros((3, tot, 4), dtype=np.uint64)
            for r, blk in enumerate(self.blocks):
                pos = 0
                for p, q0, cnt in blk:
                    tabs[:, off[p] + q0: off[p] + q0 + cnt] = allr[r][:, pos: pos + cnt]
                    pos += cnt
            # the x rounds left the scalar prefix prod_j eq(tau_x[j], r_j) with the host
            cx = api.ONE.copy()
            for j in range(self.nx):
                cx = api.host_mul(cx, host_eq(self.tau_x[j], self.rx[j]))
            self.tail = self.make_tail(tabs[0], tabs[1], tabs[2])
            self.tail.set_scale(cx)
        return self.tail

    def final(self) -> np.ndarray:
        return self._tail().final()

    def free(self):
        for eng in (self.engine, self.tail):
            if eng is not None and hasattr(eng, "free"):
                eng.free()
        self.engine = self.tail = None


def gpu_phase1_rows(ctx, comm, mats, num_cons, max_num_cons, num_vars, secs, num_proofs, num_inputs, max_num_inputs,
                    tau_p, tau_q, tau_x, satisfied: bool = False):
    """ShardedRows on this rank's GPU. mats: (A_list, B_list, C_list) of host COO triples, one per
    instance (or one shared); secs: witness sections as objects with .num_inputs and .w_mat[p][q]
    (host arrays; a section with one instance / one proof is shared, as in the reference).
    Returns (prover, z_local, blocks of this rank): the caller keeps z_local for the Z bind."""
    empty = np.zeros((0, 4), dtype=np.uint64)
    P = len(num_proofs)
    state = {}

    def make_engine(blocks, weights):
        ps = sorted({p for p, _, _ in blocks})
        assert len(ps) == len(blocks), "one block per instance and rank"
        shared = len(mats[0]) == 1
        sel = [0] if shared else ps
        inst = api.R1CSInstance(ctx, len(sel), max_num_cons, [num_cons[0 if shared else p] for p in sel], num_vars,
                                [mats[0][i] for i in sel], [mats[1][i] for i in sel], [mats[2][i] for i in sel])
        Ql = [c for _, _, c in blocks]
        dsecs = []
        for ws in secs:
            single = len(ws.w_mat) == 1
            src_p = [0] if single else ps
            nq, rows = [], []
            for k, p in enumerate(src_p):
                short = len(ws.w_mat[p]) == 1
                if single or short:
                    q0, cnt = 0, 1
                else:
                    _, q0, cnt = blocks[k]
                nq.append(cnt)
                rows.extend(ws.w_mat[p][q0: q0 + cnt])
            dsecs.append(api.ProverWitnessSecInfo(ctx, nq, [ws.num_inputs[p] for p in src_p], np.concatenate(rows)))
        z = api.ZMat(ctx, Ql, [num_inputs[p] for p in ps], dsecs)
        sc = api.sumcheck_phase1(ctx, inst, z, Ql, max(Ql), [num_cons[0 if shared else p] for p in ps], max_num_cons,
                                 max_num_inputs, np.tile(api.ONE, (log2(next_pow2(len(ps))), 1)), np.tile(api.ONE, (log2(max(Ql)), 1)), tau_x)
        sc.set_row_weights(weights)
        if satisfied:
            sc.set_satisfied()
        state.update(z=z, inst=inst, secs=dsecs)
        return sc

    def make_tail(Az, Bz, Cz):
        return api.SumcheckPhase1.from_tables(ctx, list(num_proofs), max(num_proofs), [1] * P, 1, Az, Bz, Cz, tau_p, tau_q, empty)

    sh = ShardedRows(comm, num_proofs, max_num_cons, tau_p, tau_q, tau_x, make_engine, make_tail)
    sh._state = state  # keeps the rank-local instance, sections and z_mat alive
    return sh, state.get("z"), sh.mine


def gpu_bind_rq_rows(ctx, comm, z_local, blocks, rq_rev, num_inputs, W: int, peer: "PeerTable"):
    """Z bound to rq for a row-sharded batch: this rank's rows weighted by eq(rq, q) land at their
    instances' places in the batch-wide [p][w][y] table (zeros elsewhere), then the modular all-reduce
    over peer memory. peer.n = W * sum_p num_inputs[p]."""
    rq_rev = np.asarray(rq_rev, dtype=np.uint64).reshape(-1, 4)
    off, tot = [], 0
    for y in num_inputs:
        off.append(tot)
        tot += W * y
    assert peer.n == tot
    peer.poly.zero()
    if blocks:
        w = np.stack([api.host_eq_weight(rq_rev, q) for _, q0, cnt in blocks for q in range(q0, q0 + cnt)])
        api.zmat_bind_weights(ctx, z_local, w, peer.poly, [off[p] for p, _, _ in blocks])
    return peer.all_reduce() if comm.world > 1 else peer.poly


# ---------------------------------------------------------------------------------------------
# Live Fiat-Shamir transcript over a sharded proof. Only one rank owns the transcript (merlin state,
# RandomTape, the per-round sigma protocol); the others need nothing but the challenge of each round.
# The leader wraps the sharded prover so that its round_bind first publishes r_j through the same
# exchange the evaluations use; followers replay the rounds with the challenges they receive.
class LeaderRounds:
    """round_eval / round_bind / final of a sharded prover (ShardedPhase1 or ShardedRows) for the rank
    that runs the transcript: hand it to the code that drives an unsharded prover (the ZK sumcheck
    glue of the host)."""

    def __init__(self, sharded, comm):
        self.sh, self.comm = sharded, comm
        self.num_rounds = sharded.num_rounds

    def round_eval(self):
        return self.sh.round_eval()

    def round_bind(self, r):
        r = np.ascontiguousarray(np.asarray(r, dtype=np.uint64).reshape(4))
        self.comm.all_gather(r)  # broadcast: the followers take the leader's entry
        self.sh.round_bind(r)

    def final(self):
        return self.sh.final()


def follow_rounds(sharded, comm, leader: int = 0):
    """The other ranks' side of LeaderRounds: take part in every round's exchange, bind with the
    challenge the leader publishes. Returns the challenges (every rank ends with the same r)."""
    out = []
    for _ in range(sharded.num_rounds):
        sharded.round_eval()
        r = comm.all_gather(np.zeros(4, dtype=np.uint64))[leader]
        sharded.round_bind(r)
        out.append(r)
    return np.stack(out) if out else np.zeros((0, 4), dtype=np.uint64)
