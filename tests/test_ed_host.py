"""CPU check of the curve arithmetic the CUDA commitment kernels compile
(spartan_parallel_b200/csrc/ed25519.cuh is __host__ __device__): field ops, point
addition / doubling, ristretto encode / decode against the oracle."""
import os
import random
import subprocess

import pytest

from oracle import ristretto as G

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("ed") / "ed_host_check")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-Wno-unknown-pragmas", "-o", exe, os.path.join(ROOT, "tools", "ed_host_check.cpp")])
    return exe


def run(exe, cmds):
    return subprocess.run([exe], input="\n".join(cmds) + "\n", capture_output=True, text=True, check=True).stdout.split()


def h(x):
    return x.to_bytes(32, "little").hex()


def test_field_ops(harness):
    rng = random.Random(1)
    P = G.P
    cmds, want = [], []
    for _ in range(300):
        a = rng.choice([0, 1, P - 1, P - 2, 2**255 - 20, rng.randrange(P)])
        b = rng.choice([0, 1, P - 1, 19, rng.randrange(P)])
        cmds += [f"mul {h(a)} {h(b)}", f"sub {h(a)} {h(b)}"]
        want += [h(a * b % P), h((a - b) % P)]
    assert run(harness, cmds) == want


def test_group_ops_and_encoding(harness):
    rng = random.Random(2)
    pts = [G.BASEPOINT.mul(rng.randrange(1, G.L)) for _ in range(5)] + [G.Point.identity()]
    cmds, want = [], []
    for p in pts:
        for k in [0, 1, 2, G.L - 1, rng.randrange(G.L)]:
            cmds.append(f"smul {h(k)} {p.compress().hex()}")
            want.append(p.mul(k).compress().hex())
    for p, q in zip(pts, pts[1:] + pts[:1]):
        cmds += [f"add {p.compress().hex()} {q.compress().hex()}", f"add {p.compress().hex()} {p.compress().hex()}"]
        want += [(p + q).compress().hex(), (p + p).compress().hex()]
    assert run(harness, cmds) == want


def test_invalid_encodings_rejected(harness):
    from tests.test_oracle_group import RFC_BAD, RFC_MULTIPLES

    out = run(harness, [f"dec {x}" for x in RFC_BAD] + [f"dec {x}" for x in RFC_MULTIPLES])
    assert out == ["invalid"] * len(RFC_BAD) + ["ok"] * len(RFC_MULTIPLES)


def test_from_uniform_bytes(harness):
    import hashlib

    rng = random.Random(3)
    blobs = [hashlib.sha512(b"Ristretto is traditionally a short shot of espresso coffee").digest()]
    blobs += [rng.getrandbits(512).to_bytes(64, "little") for _ in range(20)]
    out = run(harness, [f"uni {b[:32].hex()} {b[32:].hex()}" for b in blobs])
    assert out == [G.from_uniform_bytes(b).compress().hex() for b in blobs]
    assert out[0] == "3066f82a1a747d45120d1740f14358531a8f04bbffe6a819f86dfe50f44a0a46"
