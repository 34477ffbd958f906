"""CPU tests of the host mailbox a sharded proof exchanges its round evaluations through
(spg_mailbox_all_gather / spg_mailbox_poison in libspgpu.so; no device needed): ordering of the
gathered payloads, double buffering over many calls, and the two ways a wait ends without its
peers -- a poisoned slot and the deadline."""
import ctypes as C
import os
import subprocess
import sys
import threading

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SLOT, STRIDE = 4096, 4096 + 64


def _mailbox(world):
    return np.zeros(2 * world * STRIDE // 8, dtype=np.uint64)


def _gather(L, box, rank, world, calls, payload):
    out = np.empty((world, payload.size), dtype=np.uint64)
    rc = L.spg_mailbox_all_gather(C.c_void_p(box.ctypes.data), STRIDE, rank, world, C.byref(calls),
                                  payload.ctypes.data_as(C.c_void_p), payload.size * 8, out.ctypes.data_as(C.c_void_p))
    return rc, out


def test_all_gather_three_ranks_many_rounds():
    from spartan_parallel_b200 import _lib

    L = _lib.lib()
    world, rounds = 3, 200
    box = _mailbox(world)
    results = [None] * world

    def run(rank):
        calls = C.c_uint64(0)
        seen = []
        for j in range(rounds):
            payload = np.arange(12, dtype=np.uint64) + 1000 * rank + 7 * j
            rc, out = _gather(L, box, rank, world, calls, payload)
            assert rc == 0
            seen.append(out.copy())
        results[rank] = seen

    ts = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    [t.start() for t in ts]
    [t.join(60) for t in ts]
    for rank in range(world):
        for j in range(rounds):
            for r in range(world):
                assert np.array_equal(results[rank][j][r], np.arange(12, dtype=np.uint64) + 1000 * r + 7 * j)


def test_poisoned_peer_ends_the_wait():
    from spartan_parallel_b200 import _lib

    L = _lib.lib()
    world = 2
    box = _mailbox(world)
    L.spg_mailbox_poison(C.c_void_p(box.ctypes.data), STRIDE, 1, world)  # rank 1 failed before publishing
    rc, _ = _gather(L, box, 0, world, C.c_uint64(0), np.ones(12, dtype=np.uint64))
    assert rc != 0
    assert b"rank 1 reported a failure" in L.spg_last_error()
    # rank 0 poisoned its own slots on the way out, so a third party would see it too
    assert box[0] == np.uint64(0xFFFFFFFFFFFFFFFF)


def test_deadline_ends_the_wait():
    code = (
        "import ctypes as C, numpy as np, sys\n"
        f"sys.path.insert(0, {ROOT!r})\n"
        "from spartan_parallel_b200 import _lib\n"
        "L = _lib.lib()\n"
        f"box = np.zeros(2 * 2 * {STRIDE} // 8, dtype=np.uint64)\n"
        "calls = C.c_uint64(0)\n"
        "p = np.ones(12, dtype=np.uint64); out = np.empty((2, 12), dtype=np.uint64)\n"
        f"rc = L.spg_mailbox_all_gather(C.c_void_p(box.ctypes.data), {STRIDE}, 0, 2, C.byref(calls), p.ctypes.data_as(C.c_void_p), 96, out.ctypes.data_as(C.c_void_p))\n"
        "print(rc, L.spg_last_error().decode())\n"
    )
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=60,
                         env=dict(os.environ, SPG_MAILBOX_TIMEOUT_S="0.3"))
    assert out.returncode == 0, out.stderr[-1000:]
    assert "did not publish" in out.stdout and not out.stdout.startswith("0 ")
