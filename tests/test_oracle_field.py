"""Pins the oracle field against every known-answer test of the reference
(/root/reference/src/scalar/ristretto255.rs:776-1201), a pure-python big-int model
and libsodium (PyNaCl) as an independent implementation of the same field."""
import random

import numpy as np
import pytest

from oracle import cbind as O

Q = (1 << 252) + 27742317777372353535851937790883648493
RMONT = (1 << 256) % Q
M64 = (1 << 64) - 1

MODULUS = [0x5812631A5CF5D3ED, 0x14DEF9DEA2F79CD6, 0x0, 0x1000000000000000]
R = [0xD6EC31748D98951D, 0xC6EF5BF4737DCF70, 0xFFFFFFFFFFFFFFFE, 0x0FFFFFFFFFFFFFFF]
R2 = [0xA40611E3449C0F01, 0xD00E1BA768859347, 0xCEEC73D217F5BE65, 0x0399411B7C309A3D]
R3 = [0x2A9E49687B83A2DB, 0x278324E6AEF7F3EC, 0x8065DC6C04EC5B65, 0x0E530B773599CEC7]
LARGEST = [0x5812631A5CF5D3EC, 0x14DEF9DEA2F79CD6, 0x0, 0x1000000000000000]


def S(limbs):
    return np.array(limbs, dtype=np.uint64)


def limbs_int(a):
    return sum(int(a[i]) << (64 * i) for i in range(4))


def eq(a, b):
    return np.array_equal(np.asarray(a, dtype=np.uint64), np.asarray(b, dtype=np.uint64))


def test_inv_constant():
    # ristretto255.rs:776-789
    inv = 1
    for _ in range(63):
        inv = (inv * inv) & M64
        inv = (inv * MODULUS[0]) & M64
    inv = (-inv) & M64
    assert inv == 0xD2B51DA312547E1B


def test_constants_rederived():
    assert limbs_int(MODULUS) == Q
    assert limbs_int(R) == RMONT
    assert limbs_int(R2) == pow(2, 512, Q)
    assert limbs_int(R3) == pow(2, 768, Q)


def test_to_bytes_kats():
    # ristretto255.rs:818-851
    assert O.to_bytes(O.ZERO) == bytes(32)
    assert O.to_bytes(O.ONE) == bytes([1] + [0] * 31)
    assert O.to_bytes(S(R2)) == bytes([29, 149, 152, 141, 116, 49, 236, 214, 112, 207, 125, 115, 244, 91, 239, 198, 254] + [255] * 14 + [15])
    assert O.to_bytes(O.neg(O.ONE)) == bytes([236, 211, 245, 92, 26, 99, 18, 88, 214, 156, 247, 162, 222, 249, 222, 20] + [0] * 15 + [16])
    # test_debug: Debug prints to_bytes reversed
    assert O.to_bytes(S(R2))[::-1].hex() == "0ffffffffffffffffffffffffffffffec6ef5bf4737dcf70d6ec31748d98951d"


def test_from_bytes_kats():
    # ristretto255.rs:853-932
    v, ok = O.from_bytes(bytes(32))
    assert ok and eq(v, O.ZERO)
    v, ok = O.from_bytes(bytes([1] + [0] * 31))
    assert ok and eq(v, O.ONE)
    v, ok = O.from_bytes(bytes([29, 149, 152, 141, 116, 49, 236, 214, 112, 207, 125, 115, 244, 91, 239, 198, 254] + [255] * 14 + [15]))
    assert ok and eq(v, R2)
    _, ok = O.from_bytes(bytes([236, 211, 245, 92, 26, 99, 18, 88, 214, 156, 247, 162, 222, 249, 222, 20] + [0] * 15 + [16]))
    assert ok
    bad = [
        [1, 0, 0, 0, 255, 255, 255, 255, 254, 91, 254, 255, 2, 164, 189, 83, 5, 216, 161, 9, 8, 216, 57, 51, 72, 125, 157, 41, 83, 167, 237, 115],
        [2, 0, 0, 0, 255, 255, 255, 255, 254, 91, 254, 255, 2, 164, 189, 83, 5, 216, 161, 9, 8, 216, 57, 51, 72, 125, 157, 41, 83, 167, 237, 115],
        [1, 0, 0, 0, 255, 255, 255, 255, 254, 91, 254, 255, 2, 164, 189, 83, 5, 216, 161, 9, 8, 216, 58, 51, 72, 125, 157, 41, 83, 167, 237, 115],
        [1, 0, 0, 0, 255, 255, 255, 255, 254, 91, 254, 255, 2, 164, 189, 83, 5, 216, 161, 9, 8, 216, 57, 51, 72, 125, 157, 41, 83, 167, 237, 116],
    ]
    for b in bad:
        _, ok = O.from_bytes(bytes(b))
        assert not ok
    # the modulus itself and modulus+1 are non-canonical; modulus-1 is canonical
    assert not O.from_bytes(Q.to_bytes(32, "little"))[1]
    assert not O.from_bytes((Q + 1).to_bytes(32, "little"))[1]
    assert O.from_bytes((Q - 1).to_bytes(32, "little"))[1]


def test_from_u512_kats():
    # ristretto255.rs:934-968
    assert eq(O.from_u512(MODULUS + [0, 0, 0, 0]), O.ZERO)
    assert eq(O.from_u512([1, 0, 0, 0, 0, 0, 0, 0]), R)
    assert eq(O.from_u512([0, 0, 0, 0, 1, 0, 0, 0]), R2)
    assert eq(O.from_u512([M64] * 8), O.sub(S(R3), S(R)))


def test_from_bytes_wide_kats():
    # ristretto255.rs:970-1005
    b = bytes([29, 149, 152, 141, 116, 49, 236, 214, 112, 207, 125, 115, 244, 91, 239, 198, 254] + [255] * 14 + [15] + [0] * 32)
    assert eq(O.from_bytes_wide(b), R2)
    b = bytes([236, 211, 245, 92, 26, 99, 18, 88, 214, 156, 247, 162, 222, 249, 222, 20] + [0] * 15 + [16] + [0] * 32)
    assert eq(O.from_bytes_wide(b), O.neg(O.ONE))
    assert eq(O.from_bytes_wide(bytes([0xFF] * 64)),
              O.from_raw([0xA40611E3449C0F00, 0xD00E1BA768859347, 0xCEEC73D217F5BE65, 0x0399411B7C309A3D]))


def test_zero_add_neg_sub_kats():
    # ristretto255.rs:1007-1069
    Z = O.ZERO
    assert eq(O.neg(Z), Z) and eq(O.add(Z, Z), Z) and eq(O.sub(Z, Z), Z) and eq(O.mul(Z, Z), Z)
    assert eq(O.add(S(LARGEST), S(LARGEST)), [0x5812631A5CF5D3EB, 0x14DEF9DEA2F79CD6, 0, 0x1000000000000000])
    assert eq(O.add(S(LARGEST), S([1, 0, 0, 0])), Z)
    assert eq(O.neg(S(LARGEST)), [1, 0, 0, 0])
    assert eq(O.neg(S([1, 0, 0, 0])), LARGEST)
    assert eq(O.sub(S(LARGEST), S(LARGEST)), Z)
    assert eq(O.sub(Z, S(LARGEST)), O.sub(S(MODULUS), S(LARGEST)))


def _double_and_add(cur):
    acc = O.ZERO
    for byte in reversed(O.to_bytes(cur)):
        for i in reversed(range(8)):
            acc = O.add(acc, acc)
            if (byte >> i) & 1:
                acc = O.add(acc, cur)
    return acc


def test_mul_square_vs_double_and_add():
    # ristretto255.rs:1071-1127
    cur = S(LARGEST)
    for _ in range(100):
        want = _double_and_add(cur)
        assert eq(O.mul(cur, cur), want)
        assert eq(O.square(cur), want)
        cur = O.add(cur, S(LARGEST))


def test_inversion_kats():
    # ristretto255.rs:1129-1172 (invert of zero is "none": the chain returns 0)
    assert eq(O.invert(O.ZERO), O.ZERO)
    assert eq(O.invert(O.ONE), O.ONE)
    assert eq(O.invert(O.neg(O.ONE)), O.neg(O.ONE))
    tmp = S(R2)
    for _ in range(100):
        assert eq(O.mul(O.invert(tmp), tmp), O.ONE)
        tmp = O.add(tmp, S(R2))
    q_minus_2 = [0x5812631A5CF5D3EB, 0x14DEF9DEA2F79CD6, 0, 0x1000000000000000]
    r1 = S(R)
    for _ in range(100):
        a = O.invert(r1)
        b = O.pow_(r1, q_minus_2)
        assert eq(a, b)
        r1 = O.add(a, S(R))


def test_from_raw_and_double_kats():
    # ristretto255.rs:1174-1201
    assert eq(O.from_raw([0xD6EC31748D98951C, 0xC6EF5BF4737DCF70, 0xFFFFFFFFFFFFFFFE, 0x0FFFFFFFFFFFFFFF]), O.from_raw([M64] * 4))
    assert eq(O.from_raw(MODULUS), O.ZERO)
    assert eq(O.from_raw([1, 0, 0, 0]), R)
    a = O.from_raw([0x1FFF3231233FFFFD, 0x4884B7FA00034802, 0x998C4FEFECBC4FF3, 0x1824B159ACC50562])
    assert limbs_int(O.add(a, a)) == (2 * limbs_int(a)) % Q


def test_against_python_bigint_model():
    rng = random.Random(1234)
    rinv = pow(RMONT, -1, Q)
    for _ in range(300):
        x, y = rng.randrange(Q), rng.randrange(Q)
        a, b = O.from_int(x), O.from_int(y)
        assert limbs_int(a) == x * RMONT % Q
        assert O.to_int(O.mul(a, b)) == x * y % Q
        assert limbs_int(O.mul(a, b)) == limbs_int(a) * limbs_int(b) * rinv % Q
        assert O.to_int(O.add(a, b)) == (x + y) % Q
        assert O.to_int(O.sub(a, b)) == (x - y) % Q
        assert O.to_int(O.neg(a)) == (-x) % Q
        assert O.to_int(O.square(a)) == x * x % Q
    for x in (1, 2, Q - 1, 12345):
        assert O.to_int(O.invert(O.from_int(x))) == pow(x, -1, Q)


def test_batch_invert():
    rng = random.Random(7)
    xs = [rng.randrange(1, Q) for _ in range(17)]
    arr = np.stack([O.from_int(x) for x in xs])
    inv, ret = O.batch_invert(arr)
    prod = 1
    for i, x in enumerate(xs):
        assert O.to_int(inv[i]) == pow(x, -1, Q)
        prod = prod * x % Q
    assert O.to_int(ret) == pow(prod, -1, Q)


def test_against_libsodium():
    nacl = pytest.importorskip("nacl.bindings")
    rng = random.Random(99)
    for _ in range(100):
        x, y = rng.randrange(Q), rng.randrange(1, Q)
        xb, yb = x.to_bytes(32, "little"), y.to_bytes(32, "little")
        a, b = O.from_bytes(xb)[0], O.from_bytes(yb)[0]
        assert O.to_bytes(O.mul(a, b)) == nacl.crypto_core_ed25519_scalar_mul(xb, yb)
        assert O.to_bytes(O.add(a, b)) == nacl.crypto_core_ed25519_scalar_add(xb, yb)
        assert O.to_bytes(O.sub(a, b)) == nacl.crypto_core_ed25519_scalar_sub(xb, yb)
        assert O.to_bytes(O.invert(b)) == nacl.crypto_core_ed25519_scalar_invert(yb)
        wide = rng.getrandbits(512).to_bytes(64, "little")
        assert O.to_bytes(O.from_bytes_wide(wide)) == nacl.crypto_core_ed25519_scalar_reduce(wide)


def test_vector_helpers():
    rng = np.random.default_rng(5)
    wide = rng.integers(0, 2**64, size=(50, 8), dtype=np.uint64)
    a = O.vec_from_u512(wide)
    b = O.vec_from_u512(wide[::-1].copy())
    m, s, d = O.vec_mul(a, b), O.vec_add(a, b), O.vec_sub(a, b)
    for i in range(50):
        assert eq(a[i], O.from_u512(wide[i]))
        assert eq(m[i], O.mul(a[i], b[i])) and eq(s[i], O.add(a[i], b[i])) and eq(d[i], O.sub(a[i], b[i]))
