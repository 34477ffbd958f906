"""CPU checks of the C++ host mirror (libspghost.so) against the oracle: merlin transcript
and generator derivation. (The full prover needs a GPU: tests/test_gpu_proof.py.)"""
import os
import subprocess

import pytest

from oracle import merlin as M
from oracle import ristretto as G

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def host():
    from spartan_parallel_b200 import host as H

    if not os.path.exists(H.HOST_LIB_PATH):
        subprocess.check_call(["make", "-C", ROOT, "-j8"])
    return H


def test_merlin_vector(host):
    got = host.transcript_kat(b"test protocol", b"some label", b"some data", b"challenge", 32)
    assert got.hex() == "d5a21972d0d5fe320c0d263fac7fffb8145aa640af6e9bca177c03c7efcf0615"
    for n in (1, 64, 166, 167, 400):
        t = M.Transcript(b"lbl")
        t.append_message(b"a" * 7, b"x" * 300)
        assert host.transcript_kat(b"lbl", b"a" * 7, b"x" * 300, b"cc", n) == t.challenge_bytes(b"cc", n)


@pytest.mark.parametrize("n", [1, 4, 9])
def test_generators_match_oracle(host, n):
    assert host.gens_derive(b"gens_r1cs_sat", n) == G.MultiCommitGens(n, b"gens_r1cs_sat").compressed()
