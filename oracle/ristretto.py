"""ORACLE (test infrastructure, NOT product code).

ristretto255 restated from RFC 9496 with python integers. The reference gets the group
from `curve25519-dalek ^4.1.1` (Cargo.toml:14-18; absent from /root/reference):
  src/group.rs:6-7 (RistrettoPoint / CompressedRistretto), :87-117 (multiscalar mul),
  src/commitments.rs:15-33 (generators via from_uniform_bytes), :69-92 (Pedersen commit).

Pinned by RFC 9496 appendix vectors (small multiples of the generator, the
element-derivation vectors) in tests/test_oracle_group.py, and cross-checked against
libsodium's crypto_core_ristretto255_* / crypto_scalarmult_ristretto255 (PyNaCl).
"""
from __future__ import annotations

import hashlib

P = 2**255 - 19
L = 2**252 + 27742317777372353535851937790883648493
D = (-121665 * pow(121666, -1, P)) % P
SQRT_M1 = pow(2, (P - 1) // 4, P)
ONE_MINUS_D_SQ = (1 - D * D) % P
D_MINUS_ONE_SQ = (D - 1) * (D - 1) % P


def _is_neg(x: int) -> bool:
    return (x % P) & 1 == 1


def _abs(x: int) -> int:
    x %= P
    return P - x if x & 1 else x


def sqrt_ratio_m1(u: int, v: int):
    """RFC 9496 4.2: (was_square, sqrt(u/v) or sqrt(i*u/v)), non-negative root."""
    u %= P
    v %= P
    v3 = v * v % P * v % P
    v7 = v3 * v3 % P * v % P
    r = u * v3 % P * pow(u * v7 % P, (P - 5) // 8, P) % P
    check = v * r % P * r % P
    correct = check == u
    flipped = check == (-u) % P
    flipped_i = check == (-u * SQRT_M1) % P
    if flipped or flipped_i:
        r = r * SQRT_M1 % P
    r = _abs(r)
    return (correct or flipped), r


INVSQRT_A_MINUS_D = sqrt_ratio_m1(1, (-1 - D) % P)[1]
SQRT_AD_MINUS_ONE = None  # fixed below (root choice pinned by the RFC's element-derivation vectors)


class Point:
    """Extended twisted Edwards coordinates (X:Y:Z:T), a = -1."""

    __slots__ = ("X", "Y", "Z", "T")

    def __init__(self, X, Y, Z, T):
        self.X, self.Y, self.Z, self.T = X % P, Y % P, Z % P, T % P

    @staticmethod
    def identity():
        return Point(0, 1, 1, 0)

    def __add__(self, o: "Point") -> "Point":
        A = (self.Y - self.X) * (o.Y - o.X) % P
        B = (self.Y + self.X) * (o.Y + o.X) % P
        C = self.T * 2 * D % P * o.T % P
        Dd = self.Z * 2 * o.Z % P
        E, F, G, H = B - A, Dd - C, Dd + C, B + A
        return Point(E * F, G * H, F * G, E * H)

    def double(self) -> "Point":
        A = self.X * self.X % P
        B = self.Y * self.Y % P
        C = 2 * self.Z * self.Z % P
        H = A + B
        E = H - (self.X + self.Y) ** 2 % P
        G = A - B
        F = C + G
        return Point(E * F, G * H, F * G, E * H)

    def __neg__(self):
        return Point(-self.X, self.Y, self.Z, -self.T)

    def __sub__(self, o):
        return self + (-o)

    def mul(self, k: int) -> "Point":
        k %= L
        acc = Point.identity()
        for bit in bin(k)[2:] if k else "":
            acc = acc.double()
            if bit == "1":
                acc = acc + self
        return acc

    def __eq__(self, o) -> bool:
        # ristretto equality (RFC 9496 4.3.3)
        return (self.X * o.Y - self.Y * o.X) % P == 0 or (self.Y * o.Y - self.X * o.X) % P == 0

    def compress(self) -> bytes:
        """RFC 9496 4.3.2 Encode."""
        X0, Y0, Z0, T0 = self.X, self.Y, self.Z, self.T
        u1 = (Z0 + Y0) * (Z0 - Y0) % P
        u2 = X0 * Y0 % P
        _, invsqrt = sqrt_ratio_m1(1, u1 * u2 % P * u2 % P)
        den1 = invsqrt * u1 % P
        den2 = invsqrt * u2 % P
        z_inv = den1 * den2 % P * T0 % P
        ix0 = X0 * SQRT_M1 % P
        iy0 = Y0 * SQRT_M1 % P
        enchanted = den1 * INVSQRT_A_MINUS_D % P
        rotate = _is_neg(T0 * z_inv)
        if rotate:
            x, y, den_inv = iy0, ix0, enchanted
        else:
            x, y, den_inv = X0, Y0, den2
        if _is_neg(x * z_inv):
            y = (-y) % P
        s = _abs(den_inv * ((Z0 - y) % P))
        return s.to_bytes(32, "little")


def decompress(b: bytes):
    """RFC 9496 4.3.1 Decode; returns None for an invalid encoding."""
    if len(b) != 32:
        return None
    s = int.from_bytes(b, "little")
    if s >= P or _is_neg(s):
        return None
    ss = s * s % P
    u1 = (1 - ss) % P
    u2 = (1 + ss) % P
    u2_sqr = u2 * u2 % P
    v = (-(D * u1 % P * u1) - u2_sqr) % P
    was_square, invsqrt = sqrt_ratio_m1(1, v * u2_sqr % P)
    den_x = invsqrt * u2 % P
    den_y = invsqrt * den_x % P * v % P
    x = _abs(2 * s * den_x)
    y = u1 * den_y % P
    t = x * y % P
    if not was_square or _is_neg(t) or y == 0:
        return None
    return Point(x, y, 1, t)


def _map(t: int) -> Point:
    """RFC 9496 4.3.4 MAP."""
    r = SQRT_M1 * t % P * t % P
    u = (r + 1) * ONE_MINUS_D_SQ % P
    v = (-1 - r * D) * (r + D) % P
    was_square, s = sqrt_ratio_m1(u, v)
    s_prime = (-_abs(s * t)) % P
    if not was_square:
        s = s_prime
    c = -1 if was_square else r
    N = (c * (r - 1) % P * D_MINUS_ONE_SQ - v) % P
    w0 = 2 * s * v % P
    w1 = N * SQRT_AD_MINUS_ONE % P
    w2 = (1 - s * s) % P
    w3 = (1 + s * s) % P
    return Point(w0 * w3, w2 * w1, w1 * w3, w0 * w2)


def from_uniform_bytes(b: bytes) -> Point:
    """RFC 9496 4.3.4 element derivation = dalek RistrettoPoint::from_uniform_bytes."""
    assert len(b) == 64
    t1 = int.from_bytes(b[:32], "little") & ((1 << 255) - 1)
    t2 = int.from_bytes(b[32:], "little") & ((1 << 255) - 1)
    return _map(t1 % P) + _map(t2 % P)


# SQRT_AD_MINUS_ONE = sqrt(a*d - 1) with a = -1; the RFC fixes the root
# 25063068953384623474111414158702152701244531502492656460079210482610430750235
SQRT_AD_MINUS_ONE = 25063068953384623474111414158702152701244531502492656460079210482610430750235
assert SQRT_AD_MINUS_ONE * SQRT_AD_MINUS_ONE % P == (-D - 1) % P

BASEPOINT_COMPRESSED = bytes.fromhex("e2f2ae0a6abc4e71a884a961c500515f58e30b6aa582dd8db6a65945e08d2d76")
BASEPOINT = decompress(BASEPOINT_COMPRESSED)


def multiscalar_mul(scalars, points) -> Point:
    acc = Point.identity()
    for s, p in zip(scalars, points):
        if s % L:
            acc = acc + p.mul(s)
    return acc


class MultiCommitGens:
    """MultiCommitGens::new (src/commitments.rs:15-33): SHAKE256(label || basepoint) read in
    64-byte blocks, n + 1 points; the last one is h."""

    def __init__(self, n: int, label: bytes, _G=None, _h=None):
        self.n = n
        if _G is not None:
            self.G, self.h = _G, _h
            return
        xof = hashlib.shake_256(label + BASEPOINT_COMPRESSED).digest(64 * (n + 1))
        pts = [from_uniform_bytes(xof[64 * i: 64 * (i + 1)]) for i in range(n + 1)]
        self.G, self.h = pts[:n], pts[n]

    def split_at(self, mid: int):
        return MultiCommitGens(mid, b"", self.G[:mid], self.h), MultiCommitGens(self.n - mid, b"", self.G[mid:], self.h)

    def scale(self, s: int):
        return MultiCommitGens(self.n, b"", [g.mul(s) for g in self.G], self.h)

    def compressed(self) -> bytes:
        return b"".join(g.compress() for g in self.G) + self.h.compress()


def commit_scalar(v: int, blind: int, gens: MultiCommitGens) -> Point:
    """Commitments for Scalar (src/commitments.rs:73-78)."""
    assert gens.n == 1
    return multiscalar_mul([v, blind], [gens.G[0], gens.h])


def commit_vec(vals, blind: int, gens: MultiCommitGens) -> Point:
    """Commitments for [Scalar] (src/commitments.rs:87-92)."""
    assert gens.n >= len(vals)
    return multiscalar_mul(vals, gens.G[: len(vals)]) + gens.h.mul(blind)
