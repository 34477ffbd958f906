"""ORACLE (test infrastructure, NOT product code).

Transcript-level restatement of the reference's satisfiability proof:
  R1CSProof::prove / verify          src/r1csproof.rs:210-954
  ZKSumcheckInstanceProof (ZK glue)  src/sumcheck.rs:94-190, 788-1380
  sigma protocols                    src/nizk/mod.rs:15-576
  BulletReductionProof               src/nizk/bullet.rs:32-243
  PolyEvalProof (batched, disjoint)  src/dense_mlpoly.rs:861-1043
  UniPoly                            src/unipoly.rs
  generators                         src/commitments.rs:15-67, src/r1csproof.rs:45-80
  bincode layout                     serde derives on the proof structs

Scalars are numpy uint64[4] Montgomery limbs (cbind); group elements are
oracle.ristretto.Point. RandomTape is seeded by the caller (the reference seeds from
OsRng, src/random.rs:11-20 -- the only deviation, needed for reproducibility).

parity: the reference contains no proof bytes or challenge values to pin this against
(its end-to-end tests are commented out); it is validated by its own
verifier restatement (prove -> verify accepts, tampering rejects) and by the pinned
pieces underneath (field KATs, merlin vector, RFC 9496 vectors).
"""
from __future__ import annotations

import struct

import numpy as np

from . import cbind as O
from . import merlin as M
from . import ristretto as G
from .r1cs import abc_table, build_z_mat, log2, multiply_vec_block, next_pow2

ZERO, ONE = O.ZERO, O.ONE


def add(a, b):
    return O.add(a, b)


def sub(a, b):
    return O.sub(a, b)


def mul(a, b):
    return O.mul(a, b)


def sint(a) -> int:
    """canonical integer of a Montgomery scalar (Scalar::decompress_scalar, src/scalar/mod.rs:32-36)"""
    return O.to_int(a)


def eq(a, b) -> bool:
    return np.array_equal(a, b)


# ------------------------------------------------------------------ transcript (src/transcript.rs)
class Transcript(M.Transcript):
    def append_protocol_name(self, name: bytes):
        self.append_message(b"protocol-name", name)

    def append_scalar(self, label: bytes, s):
        self.append_message(label, O.to_bytes(s))

    def append_point(self, label: bytes, c: bytes):
        self.append_message(label, c)

    def append_scalars(self, label: bytes, v):
        self.append_message(label, b"begin_append_vector")
        for s in v:
            self.append_scalar(label, s)
        self.append_message(label, b"end_append_vector")

    def challenge_scalar(self, label: bytes):
        return O.from_bytes_wide(self.challenge_bytes(label, 64))

    def challenge_vector(self, label: bytes, n: int):
        return [self.challenge_scalar(label) for _ in range(n)]


class RandomTape:
    def __init__(self, name: bytes, seed):
        self.tape = Transcript(name)
        self.tape.append_scalar(b"init_randomness", seed)

    def random_scalar(self, label):
        return self.tape.challenge_scalar(label)

    def random_vector(self, label, n):
        return self.tape.challenge_vector(label, n)


# ------------------------------------------------------------------ commitments
def commit1(v, blind, gens: G.MultiCommitGens) -> G.Point:
    return G.commit_scalar(sint(v), sint(blind), gens)


def commitn(vals, blind, gens: G.MultiCommitGens) -> G.Point:
    return G.commit_vec([sint(x) for x in vals], sint(blind), gens)


class DotProductProofGens:
    def __init__(self, n, label):
        self.n = n
        self.gens_n, self.gens_1 = G.MultiCommitGens(n + 1, label).split_at(n)


class R1CSGens:
    def __init__(self, label: bytes, num_vars: int):
        ell = log2(num_vars)
        right = ell - ell // 2
        self.pc = DotProductProofGens(1 << right, label)
        self.gens_1 = self.pc.gens_1
        self.gens_3 = G.MultiCommitGens(3, label)
        self.gens_4 = G.MultiCommitGens(4, label)


# ------------------------------------------------------------------ serialization (bincode 1.x)
class Writer:
    def __init__(self):
        self.b = bytearray()

    def u64(self, x):
        self.b += struct.pack("<Q", x)

    def scalar(self, s):
        self.b += np.asarray(s, dtype="<u8").tobytes()

    def point(self, c: bytes):
        assert len(c) == 32
        self.b += c

    def scalars(self, v):
        self.u64(len(v))
        for s in v:
            self.scalar(s)

    def points(self, v):
        self.u64(len(v))
        for p in v:
            self.point(p)


class Reader:
    def __init__(self, b: bytes):
        self.b, self.pos = b, 0

    def u64(self):
        v = struct.unpack_from("<Q", self.b, self.pos)[0]
        self.pos += 8
        return v

    def scalar(self):
        s = np.frombuffer(self.b, dtype="<u8", count=4, offset=self.pos).astype(np.uint64)
        self.pos += 32
        return s

    def point(self):
        c = bytes(self.b[self.pos:self.pos + 32])
        self.pos += 32
        return c

    def scalars(self):
        return [self.scalar() for _ in range(self.u64())]

    def points(self):
        return [self.point() for _ in range(self.u64())]


# ------------------------------------------------------------------ UniPoly
def unipoly_from_evals(evals):
    return list(O.unipoly_from_evals(np.stack(evals)))


def unipoly_eval(coeffs, r):
    return O.unipoly_evaluate(np.stack(coeffs), r)


# ------------------------------------------------------------------ sigma protocols
def knowledge_prove(gens, t, tape, x, r):
    t.append_protocol_name(b"knowledge proof")
    t1, t2 = tape.random_scalar(b"t1"), tape.random_scalar(b"t2")
    C = commit1(x, r, gens).compress()
    t.append_point(b"C", C)
    alpha = commit1(t1, t2, gens).compress()
    t.append_point(b"alpha", alpha)
    c = t.challenge_scalar(b"c")
    return {"alpha": alpha, "z1": add(mul(x, c), t1), "z2": add(mul(r, c), t2)}, C


def knowledge_verify(pr, gens, t, C):
    t.append_protocol_name(b"knowledge proof")
    t.append_point(b"C", C)
    t.append_point(b"alpha", pr["alpha"])
    c = t.challenge_scalar(b"c")
    lhs = commit1(pr["z1"], pr["z2"], gens).compress()
    rhs = (G.decompress(C).mul(sint(c)) + G.decompress(pr["alpha"])).compress()
    return lhs == rhs


def equality_prove(gens, t, tape, v1, s1, v2, s2):
    t.append_protocol_name(b"equality proof")
    r = tape.random_scalar(b"r")
    C1 = commit1(v1, s1, gens).compress()
    t.append_point(b"C1", C1)
    C2 = commit1(v2, s2, gens).compress()
    t.append_point(b"C2", C2)
    alpha = gens.h.mul(sint(r)).compress()
    t.append_point(b"alpha", alpha)
    c = t.challenge_scalar(b"c")
    return {"alpha": alpha, "z": add(mul(c, sub(s1, s2)), r)}, C1, C2


def equality_verify(pr, gens, t, C1, C2):
    t.append_protocol_name(b"equality proof")
    t.append_point(b"C1", C1)
    t.append_point(b"C2", C2)
    t.append_point(b"alpha", pr["alpha"])
    c = t.challenge_scalar(b"c")
    Cd = G.decompress(C1) - G.decompress(C2)
    rhs = (Cd.mul(sint(c)) + G.decompress(pr["alpha"])).compress()
    return gens.h.mul(sint(pr["z"])).compress() == rhs


def product_prove(gens, t, tape, x, rX, y, rY, z, rZ):
    t.append_protocol_name(b"product proof")
    b1, b2, b3, b4, b5 = (tape.random_scalar(l) for l in (b"b1", b"b2", b"b3", b"b4", b"b5"))
    X = commit1(x, rX, gens).compress()
    t.append_point(b"X", X)
    Y = commit1(y, rY, gens).compress()
    t.append_point(b"Y", Y)
    Z = commit1(z, rZ, gens).compress()
    t.append_point(b"Z", Z)
    alpha = commit1(b1, b2, gens).compress()
    t.append_point(b"alpha", alpha)
    beta = commit1(b3, b4, gens).compress()
    t.append_point(b"beta", beta)
    gX = G.MultiCommitGens(1, b"", [G.decompress(X)], gens.h)
    delta = commit1(b3, b5, gX).compress()
    t.append_point(b"delta", delta)
    c = t.challenge_scalar(b"c")
    zz = [add(b1, mul(c, x)), add(b2, mul(c, rX)), add(b3, mul(c, y)), add(b4, mul(c, rY)),
          add(b5, mul(c, sub(rZ, mul(rX, y))))]
    return {"alpha": alpha, "beta": beta, "delta": delta, "z": zz}, X, Y, Z


def product_verify(pr, gens, t, X, Y, Z):
    t.append_protocol_name(b"product proof")
    for l, p in ((b"X", X), (b"Y", Y), (b"Z", Z), (b"alpha", pr["alpha"]), (b"beta", pr["beta"]), (b"delta", pr["delta"])):
        t.append_point(l, p)
    z1, z2, z3, z4, z5 = pr["z"]
    c = t.challenge_scalar(b"c")

    def chk(Pc, Xc, g, a, b):
        return (G.decompress(Pc) + G.decompress(Xc).mul(sint(c))).compress() == commit1(a, b, g).compress()

    gX = G.MultiCommitGens(1, b"", [G.decompress(X)], gens.h)
    return chk(pr["alpha"], X, gens, z1, z2) and chk(pr["beta"], Y, gens, z3, z4) and chk(pr["delta"], Z, gX, z3, z5)


def dot(a, b):
    return O.dot(np.stack(a), np.stack(b))


def dotproduct_prove(g1, gn, t, tape, x, blind_x, a, y, blind_y):
    t.append_protocol_name(b"dot product proof")
    n = len(x)
    d = tape.random_vector(b"d_vec", n)
    r_delta, r_beta = tape.random_scalar(b"r_delta"), tape.random_scalar(b"r_beta")
    Cx = commitn(x, blind_x, gn).compress()
    t.append_point(b"Cx", Cx)
    Cy = commit1(y, blind_y, g1).compress()
    t.append_point(b"Cy", Cy)
    t.append_scalars(b"a", a)
    delta = commitn(d, r_delta, gn).compress()
    t.append_point(b"delta", delta)
    beta = commit1(dot(a, d), r_beta, g1).compress()
    t.append_point(b"beta", beta)
    c = t.challenge_scalar(b"c")
    z = [add(mul(c, x[i]), d[i]) for i in range(n)]
    return {"delta": delta, "beta": beta, "z": z, "z_delta": add(mul(c, blind_x), r_delta), "z_beta": add(mul(c, blind_y), r_beta)}, Cx, Cy


def dotproduct_verify(pr, g1, gn, t, a, Cx, Cy):
    t.append_protocol_name(b"dot product proof")
    t.append_point(b"Cx", Cx)
    t.append_point(b"Cy", Cy)
    t.append_scalars(b"a", a)
    t.append_point(b"delta", pr["delta"])
    t.append_point(b"beta", pr["beta"])
    c = t.challenge_scalar(b"c")
    ok = (G.decompress(Cx).mul(sint(c)) + G.decompress(pr["delta"])) == commitn(pr["z"], pr["z_delta"], gn)
    ok &= (G.decompress(Cy).mul(sint(c)) + G.decompress(pr["beta"])) == commit1(dot(pr["z"], a), pr["z_beta"], g1)
    return bool(ok)


def w_dotproduct(w: Writer, pr):
    w.point(pr["delta"])
    w.point(pr["beta"])
    w.scalars(pr["z"])
    w.scalar(pr["z_delta"])
    w.scalar(pr["z_beta"])


def r_dotproduct(r: Reader):
    return {"delta": r.point(), "beta": r.point(), "z": r.scalars(), "z_delta": r.scalar(), "z_beta": r.scalar()}


# ------------------------------------------------------------------ bullet reduction
def bullet_prove(t, Q, Gv, H, a, b, blind, blinds):
    Gv, a, b = list(Gv), list(a), list(b)
    n = len(Gv)
    L_vec, R_vec = [], []
    blind_fin = blind
    it = iter(blinds)
    while n != 1:
        n //= 2
        aL, aR, bL, bR, GL, GR = a[:n], a[n:2 * n], b[:n], b[n:2 * n], Gv[:n], Gv[n:2 * n]
        cL, cR = dot(aL, bR), dot(aR, bL)
        blind_L, blind_R = next(it)
        L = G.multiscalar_mul([sint(x) for x in aL] + [sint(cL), sint(blind_L)], GR + [Q, H])
        R = G.multiscalar_mul([sint(x) for x in aR] + [sint(cR), sint(blind_R)], GL + [Q, H])
        Lc, Rc = L.compress(), R.compress()
        t.append_point(b"L", Lc)
        t.append_point(b"R", Rc)
        u = t.challenge_scalar(b"u")
        u_inv = O.invert(u)
        ui, uinvi = sint(u), sint(u_inv)
        a = [add(mul(aL[i], u), mul(u_inv, aR[i])) for i in range(n)]
        b = [add(mul(bL[i], u_inv), mul(u, bR[i])) for i in range(n)]
        Gv = [GL[i].mul(uinvi) + GR[i].mul(ui) for i in range(n)]
        blind_fin = add(add(blind_fin, mul(mul(blind_L, u), u)), mul(mul(blind_R, u_inv), u_inv))
        L_vec.append(Lc)
        R_vec.append(Rc)
    return {"L": L_vec, "R": R_vec}, a[0], b[0], Gv[0], blind_fin


def bullet_verify(pr, n, a, t, Gamma, Gv):
    lg = len(pr["L"])
    assert n == 1 << lg
    ch = []
    for Lc, Rc in zip(pr["L"], pr["R"]):
        t.append_point(b"L", Lc)
        t.append_point(b"R", Rc)
        ch.append(t.challenge_scalar(b"u"))
    inv, allinv = O.batch_invert(np.stack(ch)) if ch else (np.zeros((0, 4), np.uint64), ONE)
    ch_sq = [O.square(c) for c in ch]
    inv_sq = [O.square(c) for c in inv]
    s = [allinv]
    for i in range(1, n):
        lg_i = i.bit_length() - 1
        k = 1 << lg_i
        s.append(mul(s[i - k], ch_sq[(lg - 1) - lg_i]))
    G_hat = G.multiscalar_mul([sint(x) for x in s], Gv)
    a_hat = dot(a, s)
    pts = [G.decompress(p) for p in pr["L"]] + [G.decompress(p) for p in pr["R"]] + [Gamma]
    Gamma_hat = G.multiscalar_mul([sint(x) for x in ch_sq] + [sint(x) for x in inv_sq] + [1], pts)
    return G_hat, Gamma_hat, a_hat


def dplog_prove(gens: DotProductProofGens, t, tape, x, blind_x, a, y, blind_y):
    t.append_protocol_name(b"dot product proof (log)")
    n = len(x)
    assert gens.n >= n
    d = tape.random_scalar(b"d")
    r_delta = tape.random_scalar(b"r_delta")
    r_beta = tape.random_scalar(b"r_delta")  # sic (nizk/mod.rs:454)
    lg = log2(n)
    v1, v2 = tape.random_vector(b"blinds_vec_1", 2 * lg), tape.random_vector(b"blinds_vec_2", 2 * lg)
    blinds = list(zip(v1, v2))
    Cx = commitn(x, blind_x, gens.gens_n).compress()
    t.append_point(b"Cx", Cx)
    Cy = commit1(y, blind_y, gens.gens_1).compress()
    t.append_point(b"Cy", Cy)
    t.append_scalars(b"a", a)
    r = t.challenge_scalar(b"r")
    g1s = gens.gens_1.scale(sint(r))
    blind_Gamma = add(blind_x, mul(r, blind_y))
    bp, x_hat, a_hat, g_hat, rhat = bullet_prove(t, g1s.G[0], gens.gens_n.G[:n], gens.gens_n.h, x, a, blind_Gamma, blinds)
    y_hat = mul(x_hat, a_hat)
    ghat = G.MultiCommitGens(1, b"", [g_hat], gens.gens_1.h)
    delta = commit1(d, r_delta, ghat).compress()
    t.append_point(b"delta", delta)
    beta = commit1(d, r_beta, g1s).compress()
    t.append_point(b"beta", beta)
    c = t.challenge_scalar(b"c")
    z1 = add(d, mul(c, y_hat))
    z2 = add(mul(a_hat, add(mul(c, rhat), r_beta)), r_delta)
    return {"bullet": bp, "delta": delta, "beta": beta, "z1": z1, "z2": z2}, Cx, Cy


def dplog_verify(pr, n, gens: DotProductProofGens, t, a, Cx, Cy):
    t.append_protocol_name(b"dot product proof (log)")
    t.append_point(b"Cx", Cx)
    t.append_point(b"Cy", Cy)
    t.append_scalars(b"a", a)
    r = t.challenge_scalar(b"r")
    g1s = gens.gens_1.scale(sint(r))
    Gamma = G.decompress(Cx) + G.decompress(Cy).mul(sint(r))
    g_hat, Gamma_hat, a_hat = bullet_verify(pr["bullet"], n, a, t, Gamma, gens.gens_n.G[:n])
    t.append_point(b"delta", pr["delta"])
    t.append_point(b"beta", pr["beta"])
    c = t.challenge_scalar(b"c")
    lhs = ((Gamma_hat.mul(sint(c)) + G.decompress(pr["beta"])).mul(sint(a_hat)) + G.decompress(pr["delta"])).compress()
    rhs = ((g_hat + g1s.G[0].mul(sint(a_hat))).mul(sint(pr["z1"])) + g1s.h.mul(sint(pr["z2"]))).compress()
    return lhs == rhs


def w_dplog(w: Writer, pr):
    w.points(pr["bullet"]["L"])
    w.points(pr["bullet"]["R"])
    w.point(pr["delta"])
    w.point(pr["beta"])
    w.scalar(pr["z1"])
    w.scalar(pr["z2"])


def r_dplog(r: Reader):
    return {"bullet": {"L": r.points(), "R": r.points()}, "delta": r.point(), "beta": r.point(), "z1": r.scalar(), "z2": r.scalar()}


# ------------------------------------------------------------------ ZK sumcheck
def zk_sumcheck_prove(claim, blind_claim, num_rounds, engine, g1, gn, t, tape, trace=None):
    """Shared body of the two disjoint-round provers (src/sumcheck.rs:1104-1367, 816-1054);
    `engine` supplies round_eval / round_bind (oracle C loops)."""
    blinds_poly = tape.random_vector(b"blinds_poly", num_rounds)
    blinds_evals = tape.random_vector(b"blinds_evals", num_rounds)
    claim_per_round = claim
    comm_claim = commit1(claim_per_round, blind_claim, g1).compress()
    r, comm_polys, comm_evals, proofs = [], [], [], []
    for j in range(num_rounds):
        e0, e2, e3 = engine.round_eval()
        if trace is not None:
            trace.append(np.stack([e0, e2, e3]))
        poly = unipoly_from_evals([e0, sub(claim_per_round, e0), e2, e3])
        comm_poly = commitn(poly, blinds_poly[j], gn).compress()
        t.append_point(b"comm_poly", comm_poly)
        comm_polys.append(comm_poly)
        r_j = t.challenge_scalar(b"challenge_nextround")
        engine.round_bind(r_j)
        ev = unipoly_eval(poly, r_j)
        comm_eval = commit1(ev, blinds_evals[j], g1).compress()
        t.append_point(b"comm_claim_per_round", comm_claim)
        t.append_point(b"comm_eval", comm_eval)
        w = t.challenge_vector(b"combine_two_claims_to_one", 2)
        target = add(mul(w[0], claim_per_round), mul(w[1], ev))
        blind_sc = blind_claim if j == 0 else blinds_evals[j - 1]
        blind = add(mul(w[0], blind_sc), mul(w[1], blinds_evals[j]))
        a_sc = [add(ONE, ONE), ONE, ONE, ONE]
        a_ev = [ONE]
        for _ in range(3):
            a_ev.append(mul(a_ev[-1], r_j))
        a = [add(mul(w[0], a_sc[k]), mul(w[1], a_ev[k])) for k in range(4)]
        pr, _, _ = dotproduct_prove(g1, gn, t, tape, poly, blinds_poly[j], a, target, blind)
        proofs.append(pr)
        claim_per_round, comm_claim = ev, comm_eval
        r.append(r_j)
        comm_evals.append(comm_eval)
    return {"comm_polys": comm_polys, "comm_evals": comm_evals, "proofs": proofs}, r, blinds_evals[num_rounds - 1]


def zk_sumcheck_verify(pr, comm_claim, num_rounds, g1, gn, t):
    assert len(pr["comm_polys"]) == num_rounds and len(pr["comm_evals"]) == num_rounds
    r = []
    for i in range(num_rounds):
        comm_poly = pr["comm_polys"][i]
        t.append_point(b"comm_poly", comm_poly)
        r_i = t.challenge_scalar(b"challenge_nextround")
        comm_claim_per_round = comm_claim if i == 0 else pr["comm_evals"][i - 1]
        comm_eval = pr["comm_evals"][i]
        t.append_point(b"comm_claim_per_round", comm_claim_per_round)
        t.append_point(b"comm_eval", comm_eval)
        w = t.challenge_vector(b"combine_two_claims_to_one", 2)
        comm_target = G.multiscalar_mul([sint(w[0]), sint(w[1])], [G.decompress(comm_claim_per_round), G.decompress(comm_eval)]).compress()
        a_sc = [add(ONE, ONE), ONE, ONE, ONE]
        a_ev = [ONE]
        for _ in range(3):
            a_ev.append(mul(a_ev[-1], r_i))
        a = [add(mul(w[0], a_sc[k]), mul(w[1], a_ev[k])) for k in range(4)]
        if not dotproduct_verify(pr["proofs"][i], g1, gn, t, a, comm_poly, comm_target):
            return None
        r.append(r_i)
    return pr["comm_evals"][-1], r


def w_zksc(w: Writer, pr):
    w.points(pr["comm_polys"])
    w.points(pr["comm_evals"])
    w.u64(len(pr["proofs"]))
    for p in pr["proofs"]:
        w_dotproduct(w, p)


def r_zksc(r: Reader):
    cp, ce = r.points(), r.points()
    return {"comm_polys": cp, "comm_evals": ce, "proofs": [r_dotproduct(r) for _ in range(r.u64())]}


# ------------------------------------------------------------------ PolyEvalProof (batched, disjoint rounds)
def _lr_for(num_proofs, num_inputs, rq, ry):
    nvq, nvy = log2(num_proofs), log2(num_inputs)
    if nvy >= len(ry):
        ry_short = [ZERO] * (nvy - len(ry)) + list(ry)
    else:
        ry_short = list(ry[len(ry) - nvy:])
    rq_short = list(rq[len(rq) - nvq:])
    r = rq_short + ry_short
    left = len(r) // 2
    one = ONE.reshape(1, 4)
    L = O.eq_evals(np.stack(r[:left])) if left else one
    R = O.eq_evals(np.stack(r[left:])) if len(r) - left else one
    return L, R


def polyeval_prove_batched(polys, num_proofs_list, num_inputs_list, rq, ry, Zr_list, gens: DotProductProofGens, t, tape):
    t.append_protocol_name(b"polynomial evaluation proof")
    index_map, LZ_list, Zc_list, L_list, R_list = {}, [], [], [], []
    c_base = t.challenge_scalar(b"challenge_c")
    c = ONE
    for i, poly in enumerate(polys):
        key = (num_proofs_list[i], num_inputs_list[i])
        if key in index_map:
            c = mul(c, c_base)
            idx = index_map[key]
            LZ = O.dense_bound_L(poly, L_list[idx])
            LZ_list[idx] = O.vec_add(LZ_list[idx], O.vec_mul(np.tile(c, (LZ.shape[0], 1)), LZ))
            Zc_list[idx] = add(Zc_list[idx], mul(c, Zr_list[i]))
        else:
            index_map[key] = len(LZ_list)
            Zc_list.append(Zr_list[i])
            L, R = _lr_for(key[0], key[1], rq, ry)
            LZ_list.append(O.dense_bound_L(poly, L))
            L_list.append(L)
            R_list.append(R)
    proofs = []
    for i in range(len(LZ_list)):
        pr, _, _ = dplog_prove(gens, t, tape, list(LZ_list[i]), ZERO, list(R_list[i]), Zc_list[i], ZERO)
        proofs.append(pr)
    return proofs


def polyeval_verify_batched(proofs, num_proofs_list, num_inputs_list, gens: DotProductProofGens, t, rq, ry, Zr_points, comm_list):
    t.append_protocol_name(b"polynomial evaluation proof")
    index_map, LZ_list, Zc_list, L_list, R_list = {}, [], [], [], []
    c_base = t.challenge_scalar(b"challenge_c")
    c = ONE
    for i, comm in enumerate(comm_list):
        Cd = [G.decompress(x) for x in comm]
        key = (num_proofs_list[i], num_inputs_list[i])
        if key in index_map:
            c = mul(c, c_base)
            idx = index_map[key]
            LZ = G.multiscalar_mul([sint(x) for x in L_list[idx]], Cd)
            LZ_list[idx] = LZ_list[idx] + LZ.mul(sint(c))
            Zc_list[idx] = Zc_list[idx] + Zr_points[i].mul(sint(c))
        else:
            index_map[key] = len(LZ_list)
            Zc_list.append(Zr_points[i])
            L, R = _lr_for(key[0], key[1], rq, ry)
            LZ_list.append(G.multiscalar_mul([sint(x) for x in L], Cd))
            L_list.append(L)
            R_list.append(R)
    if len(LZ_list) != len(proofs):
        return False
    for i in range(len(LZ_list)):
        R = R_list[i]
        if not dplog_verify(proofs[i], len(R), gens, t, list(R), LZ_list[i].compress(), Zc_list[i].compress()):
            return False
    return True


def _factored(r):
    """EqPolynomial::compute_factored_evals (src/dense_mlpoly.rs:122-130)."""
    left = len(r) // 2
    one = ONE.reshape(1, 4)
    L = O.eq_evals(np.stack(r[:left])) if left else one
    R = O.eq_evals(np.stack(r[left:])) if len(r) - left else one
    return L, R


def _key(v):
    return tuple(int(x) for x in np.asarray(v, dtype=np.uint64).reshape(-1))


def polyeval_prove_batched_points(poly, r_list, Zr_list, gens: DotProductProofGens, t, tape):
    """PolyEvalProof::prove_batched_points (src/dense_mlpoly.rs:531-622), zero blinds: several points
    on one polynomial; points sharing the left half of r share L and are combined by powers of c."""
    t.append_protocol_name(b"polynomial evaluation proof")
    left = len(r_list[0]) // 2
    index_map, L_list, R_list, Zc_list = {}, [], [], []
    c_base = t.challenge_scalar(b"challenge_c")
    c = ONE
    for i, r in enumerate(r_list):
        Li, Ri = _factored(list(r))
        k = _key(np.stack(r[:left])) if left else ()
        if k in index_map:
            c = mul(c, c_base)
            idx = index_map[k]
            R_list[idx] = O.vec_add(R_list[idx], O.vec_mul(np.tile(c, (Ri.shape[0], 1)), Ri))
            Zc_list[idx] = add(Zc_list[idx], mul(c, Zr_list[i]))
        else:
            index_map[k] = len(L_list)
            L_list.append(Li)
            R_list.append(Ri)
            Zc_list.append(Zr_list[i])
    proofs = []
    for i in range(len(L_list)):
        LZ = O.dense_bound_L(poly, L_list[i])
        pr, _, _ = dplog_prove(gens, t, tape, list(LZ), ZERO, list(R_list[i]), Zc_list[i], ZERO)
        proofs.append(pr)
    return proofs


def polyeval_prove_batched_instances(polys, r_list, Zr_list, gens: DotProductProofGens, t, tape):
    """PolyEvalProof::prove_batched_instances (src/dense_mlpoly.rs:689-780), zero blinds: one point per
    polynomial (padded with leading zeros or trimmed to the polynomial's size); instances with the same
    size and R are combined."""
    t.append_protocol_name(b"polynomial evaluation proof")
    index_map, LZ_list, Zc_list, R_list = {}, [], [], []
    c_base = t.challenge_scalar(b"challenge_c")
    c = ONE
    for i, poly in enumerate(polys):
        nv = log2(poly.shape[0])
        r = list(r_list[i])
        r = [ZERO] * (nv - len(r)) + r if nv >= len(r) else r[len(r) - nv:]
        L, R = _factored(r)
        k = (nv, _key(R))
        LZ = O.dense_bound_L(poly, L)
        if k in index_map:
            c = mul(c, c_base)
            idx = index_map[k]
            LZ_list[idx] = O.vec_add(LZ_list[idx], O.vec_mul(np.tile(c, (LZ.shape[0], 1)), LZ))
            Zc_list[idx] = add(Zc_list[idx], mul(c, Zr_list[i]))
        else:
            index_map[k] = len(LZ_list)
            Zc_list.append(Zr_list[i])
            R_list.append(R)
            LZ_list.append(LZ)
    proofs = []
    for i in range(len(LZ_list)):
        pr, _, _ = dplog_prove(gens, t, tape, list(LZ_list[i]), ZERO, list(R_list[i]), Zc_list[i], ZERO)
        proofs.append(pr)
    return proofs


def polyeval_prove_uni_batched_instances(polys, r, Zr, gens: DotProductProofGens, t, tape):
    """PolyEvalProof::prove_uni_batched_instances (src/dense_mlpoly.rs:1046-1130): the polynomials read as
    univariate ones, opened at the single point r; returns (proof, C_Zr_prime)."""
    t.append_protocol_name(b"polynomial evaluation proof")
    max_nv = max(log2(p.shape[0]) for p in polys)
    R_size = 1 << (max_nv - max_nv // 2)
    R, rb = [], ONE
    for _ in range(R_size):
        R.append(rb)
        rb = mul(rb, r)
    L_map = {}
    c_base = t.challenge_scalar(b"challenge_c")
    c = ONE
    LZ_comb = np.zeros((R_size, 4), dtype=np.uint64)
    Zr_comb = ZERO
    for i, poly in enumerate(polys):
        nv = log2(poly.shape[0])
        if nv not in L_map:
            Ls, Rs = 1 << (nv // 2), 1 << (nv - nv // 2)
            r_base = ONE
            for _ in range(Rs):
                r_base = mul(r_base, r)
            L, lb = [], ONE
            for _ in range(Ls):
                L.append(lb)
                lb = mul(lb, r_base)
            L_map[nv] = np.stack(L)
        LZ = O.dense_bound_L(poly, L_map[nv])
        pad = np.zeros((R_size, 4), dtype=np.uint64)
        pad[: LZ.shape[0]] = O.vec_mul(np.tile(c, (LZ.shape[0], 1)), LZ)
        LZ_comb = O.vec_add(LZ_comb, pad)
        Zr_comb = add(Zr_comb, mul(c, Zr[i]))
        c = mul(c, c_base)
    pr, _, Cy = dplog_prove(gens, t, tape, list(LZ_comb), ZERO, R, Zr_comb, ZERO)
    return pr, Cy


def serialize_polyeval_proofs(proofs) -> bytes:
    """bincode of Vec<PolyEvalProof>"""
    w = Writer()
    w.u64(len(proofs))
    for pr in proofs:
        w_dplog(w, pr)
    return bytes(w.b)


def poly_commit(Z, gens_n: G.MultiCommitGens):
    """DensePolynomial::commit with zero blinds (src/dense_mlpoly.rs:199-239)."""
    ell = log2(Z.shape[0])
    Ls = 1 << (ell // 2)
    Rs = Z.shape[0] // Ls
    return [commitn(list(Z[Rs * i:Rs * (i + 1)]), ZERO, gens_n).compress() for i in range(Ls)]


# ------------------------------------------------------------------ R1CSProof
class _Sc1Engine:
    def __init__(self, sc):
        self.sc = sc

    def round_eval(self):
        return self.sc.round_eval()

    def round_bind(self, r):
        self.sc.round_bind(r)


def _prefix_list(rw, W):
    Wp = next_pow2(W)
    nrw = log2(Wp)
    if Wp > 8:
        raise ValueError(f"Unsupported num_witness_secs: {W}")
    out = []
    for k in range(Wp):
        acc = ONE
        for b in range(nrw):
            acc = mul(acc, rw[b] if (k >> (nrw - 1 - b)) & 1 else sub(ONE, rw[b]))
        out.append(acc)
    return out


def r1cs_prove(inst, num_instances, max_num_proofs, num_proofs, max_num_inputs, num_inputs, witness_secs, gens: R1CSGens,
               t: Transcript, tape: RandomTape, trace=None):
    """R1CSProof::prove (src/r1csproof.rs:210-685). Returns (proof dict, [rp, rq_rev, rx, rw++ry])."""
    t.append_protocol_name(b"R1CS proof")
    W = len(witness_secs)
    num_cons = inst.max_num_cons
    block_num_cons = [inst.num_cons[0]] * num_instances if inst.num_instances == 1 else list(inst.num_cons)
    z_mat = build_z_mat(num_instances, num_proofs, num_inputs, witness_secs)
    nrp, nrq, nrx = log2(next_pow2(num_instances)), log2(max_num_proofs), log2(num_cons)
    nrw, nry = log2(next_pow2(W)), log2(max_num_inputs)
    tau_p = t.challenge_vector(b"challenge_tau_p", nrp)
    tau_q = t.challenge_vector(b"challenge_tau_q", nrq)
    tau_x = t.challenge_vector(b"challenge_tau_x", nrx)
    one = ONE.reshape(1, 4)
    ev = lambda taus: O.eq_evals(np.stack(taus)) if len(taus) else one
    Az, Bz, Cz = multiply_vec_block(inst, num_instances, num_proofs, max_num_inputs, block_num_cons, z_mat)
    mk = lambda T: O.Pqx.new_rev(T, 1, num_proofs, max_num_proofs, block_num_cons, num_cons)
    sc1 = O.Sc1(nrx, nrq, nrp, num_proofs, block_num_cons, ev(tau_p), ev(tau_q), ev(tau_x), mk(Az), mk(Bz), mk(Cz))
    tr1 = [] if trace is not None else None
    sc_proof_phase1, r1, blind_claim_postsc1 = zk_sumcheck_prove(ZERO, ZERO, nrx + nrq + nrp, _Sc1Engine(sc1), gens.gens_1,
                                                                 gens.gens_4, t, tape, tr1)
    tau_claim, Az_claim, Bz_claim, Cz_claim = sc1.final()
    Az_blind, Bz_blind, Cz_blind, prod_blind = (tape.random_scalar(l) for l in (b"Az_blind", b"Bz_blind", b"Cz_blind", b"prod_Az_Bz_blind"))
    pok_Cz, comm_Cz = knowledge_prove(gens.gens_1, t, tape, Cz_claim, Cz_blind)
    prod = mul(Az_claim, Bz_claim)
    proof_prod, comm_Az, comm_Bz, comm_prod = product_prove(gens.gens_1, t, tape, Az_claim, Az_blind, Bz_claim, Bz_blind, prod, prod_blind)
    t.append_point(b"comm_Az_claim", comm_Az)
    t.append_point(b"comm_Bz_claim", comm_Bz)
    t.append_point(b"comm_Cz_claim", comm_Cz)
    t.append_point(b"comm_prod_Az_Bz_claims", comm_prod)
    blind_expected1 = mul(tau_claim, sub(prod_blind, Cz_blind))
    claim_post1 = mul(sub(mul(Az_claim, Bz_claim), Cz_claim), tau_claim)
    proof_eq1, _, _ = equality_prove(gens.gens_1, t, tape, claim_post1, blind_expected1, claim_post1, blind_claim_postsc1)
    rx_rev, rq_rev, rp = r1[:nrx], r1[nrx:nrx + nrq], r1[nrx + nrq:]
    rx, rq = rx_rev[::-1], rq_rev[::-1]
    r_A, r_B, r_C = (t.challenge_scalar(l) for l in (b"challenge_Az", b"challenge_Bz", b"challenge_Cz"))
    claim_phase2 = add(add(mul(r_A, Az_claim), mul(r_B, Bz_claim)), mul(r_C, Cz_claim))
    blind_claim_phase2 = add(add(mul(r_A, Az_blind), mul(r_B, Bz_blind)), mul(r_C, Cz_blind))
    evals_rx = ev(rx)
    abc = abc_table(inst, W, max_num_inputs, num_inputs, evals_rx, r_A, r_B, r_C)
    abc_flat = np.concatenate([a.reshape(-1, 4) for a in abc])
    n_abc = inst.num_instances
    ABC = O.Pqx.new_rev(abc_flat, W, [1] * n_abc, 1, list(num_inputs[:n_abc]), max_num_inputs)
    z_flat = np.concatenate([z.reshape(-1, 4) for z in z_mat])
    Z = O.Pqx.new_rev(z_flat, W, num_proofs, max_num_proofs, num_inputs, max_num_inputs)
    for r in rq_rev:
        Z.bound_poly(r, O.MODE_Q)
    sc2 = O.Sc2(nry, nrw, nrp, inst.num_instances == 1, W, num_inputs, ev(rp), ABC, Z)
    tr2 = [] if trace is not None else None
    sc_proof_phase2, r2, blind_claim_postsc2 = zk_sumcheck_prove(claim_phase2, blind_claim_phase2, nry + nrw + nrp, _Sc1Engine(sc2),
                                                                 gens.gens_1, gens.gens_4, t, tape, tr2)
    claims2 = sc2.final()
    ry_rev, rw, rp2 = r2[:nry], r2[nry:nry + nrw], r2[nry + nrw:]
    ry = ry_rev[::-1]
    ry_factors = [ONE]
    for i in range(nry):
        ry_factors.append(mul(ry_factors[i], sub(ONE, ry[i])))
    poly_list, npl, nil, Zr_list = [], [], [], []
    eval_vars_at_ry_list = [[] for _ in range(W)]
    comm_vars_at_ry_list = [[] for _ in range(W)]
    for i, w in enumerate(witness_secs):
        for p in range(len(w.w_mat)):
            poly = w.poly_w(p)
            wq, wy = len(w.w_mat[p]), w.num_inputs[p]
            poly_list.append(poly)
            npl.append(wq)
            nil.append(wy)
            if wy >= max_num_inputs:
                ry_short = [ZERO] * (log2(wy) - log2(max_num_inputs)) + list(ry)
            else:
                ry_short = list(ry[nry - log2(wy):])
            rq_short = list(rq[nrq - log2(wq):])
            r = rq_short + ry_short
            e = O.dense_evaluate(poly, np.stack(r)) if r else poly[0]
            Zr_list.append(e)
            eval_vars_at_ry_list[i].append(e if wy >= max_num_inputs else mul(e, ry_factors[nry - log2(wy)]))
            comm_vars_at_ry_list[i].append(commit1(e, ZERO, gens.pc.gens_1).compress())
    proof_eval_vars = polyeval_prove_batched(poly_list, npl, nil, rq, ry, Zr_list, gens.pc, t, tape)
    prefix = _prefix_list(rw, W)
    comb_list = []
    for p in range(num_instances):
        comb = ZERO
        for i in range(W):
            wp = 0 if len(witness_secs[i].w_mat) == 1 else p
            comb = add(comb, mul(prefix[i], eval_vars_at_ry_list[i][wp]))
        for q in range(nrq - log2(num_proofs[p])):
            comb = mul(comb, sub(ONE, rq[q]))
        comb_list.append(comb)
    pv = np.stack(comb_list + [ZERO] * (next_pow2(num_instances) - num_instances))
    eval_vars_at_ry = O.dense_evaluate(pv, np.stack(rp2)) if nrp else pv[0]
    comm_vars_at_ry = commit1(eval_vars_at_ry, ZERO, gens.pc.gens_1).compress()
    claim_post2 = mul(mul(claims2[0], claims2[1]), claims2[2])
    proof_eq2, _, _ = equality_prove(gens.pc.gens_1, t, tape, claim_post2, ZERO, claim_post2, blind_claim_postsc2)
    proof = {
        "sc_proof_phase1": sc_proof_phase1,
        "claims_phase2": (comm_Az, comm_Bz, comm_Cz, comm_prod),
        "pok_claims_phase2": (pok_Cz, proof_prod),
        "proof_eq_sc_phase1": proof_eq1,
        "sc_proof_phase2": sc_proof_phase2,
        # the reference pre-sizes the list to W and then pushes W more empty Vecs (:532-538)
        "comm_vars_at_ry_list": comm_vars_at_ry_list + [[] for _ in range(W)],
        "comm_vars_at_ry": comm_vars_at_ry,
        "proof_eval_vars_at_ry_list": proof_eval_vars,
        "proof_eq_sc_phase2": proof_eq2,
    }
    if trace is not None:
        trace.update({"evals1": tr1, "evals2": tr2, "claims1": np.stack([tau_claim, Az_claim, Bz_claim, Cz_claim]), "claims2": claims2})
    return proof, [rp2, rq_rev, rx, list(rw) + list(ry)]


def serialize_r1cs_proof(pr) -> bytes:
    """bincode layout of R1CSProof (src/r1csproof.rs:25-43)."""
    w = Writer()
    w_zksc(w, pr["sc_proof_phase1"])
    for c in pr["claims_phase2"]:
        w.point(c)
    pok, pp = pr["pok_claims_phase2"]
    w.point(pok["alpha"])
    w.scalar(pok["z1"])
    w.scalar(pok["z2"])
    w.point(pp["alpha"])
    w.point(pp["beta"])
    w.point(pp["delta"])
    for z in pp["z"]:
        w.scalar(z)
    w.point(pr["proof_eq_sc_phase1"]["alpha"])
    w.scalar(pr["proof_eq_sc_phase1"]["z"])
    w_zksc(w, pr["sc_proof_phase2"])
    w.u64(len(pr["comm_vars_at_ry_list"]))
    for v in pr["comm_vars_at_ry_list"]:
        w.points(v)
    w.point(pr["comm_vars_at_ry"])
    w.u64(len(pr["proof_eval_vars_at_ry_list"]))
    for p in pr["proof_eval_vars_at_ry_list"]:
        w_dplog(w, p)
    w.point(pr["proof_eq_sc_phase2"]["alpha"])
    w.scalar(pr["proof_eq_sc_phase2"]["z"])
    return bytes(w.b)


def deserialize_r1cs_proof(b: bytes):
    r = Reader(b)
    pr = {"sc_proof_phase1": r_zksc(r)}
    pr["claims_phase2"] = tuple(r.point() for _ in range(4))
    pok = {"alpha": r.point(), "z1": r.scalar(), "z2": r.scalar()}
    pp = {"alpha": r.point(), "beta": r.point(), "delta": r.point(), "z": [r.scalar() for _ in range(5)]}
    pr["pok_claims_phase2"] = (pok, pp)
    pr["proof_eq_sc_phase1"] = {"alpha": r.point(), "z": r.scalar()}
    pr["sc_proof_phase2"] = r_zksc(r)
    pr["comm_vars_at_ry_list"] = [r.points() for _ in range(r.u64())]
    pr["comm_vars_at_ry"] = r.point()
    pr["proof_eval_vars_at_ry_list"] = [r_dplog(r) for _ in range(r.u64())]
    pr["proof_eq_sc_phase2"] = {"alpha": r.point(), "z": r.scalar()}
    assert r.pos == len(b), "trailing bytes"
    return pr


def r1cs_verify(pr, num_instances, max_num_proofs, num_proofs, max_num_inputs, wit_num_proofs, wit_num_inputs, wit_comms,
                num_cons, gens: R1CSGens, evals, t: Transcript):
    """R1CSProof::verify (src/r1csproof.rs:687-954). wit_*[i][p]: VerifierWitnessSecInfo of
    section i; evals = (A, B, C) evaluations supplied by the caller. Returns the challenge
    vectors on success, None on rejection."""
    t.append_protocol_name(b"R1CS proof")
    W = len(wit_comms)
    nrp, nrq, nrx = log2(next_pow2(num_instances)), log2(max_num_proofs), log2(num_cons)
    nrw, nry = log2(next_pow2(W)), log2(max_num_inputs)
    tau_p = t.challenge_vector(b"challenge_tau_p", nrp)
    tau_q = t.challenge_vector(b"challenge_tau_q", nrq)
    tau_x = t.challenge_vector(b"challenge_tau_x", nrx)
    claim_phase1 = commit1(ZERO, ZERO, gens.gens_1).compress()
    res = zk_sumcheck_verify(pr["sc_proof_phase1"], claim_phase1, nrx + nrq + nrp, gens.gens_1, gens.gens_4, t)
    if res is None:
        return None
    comm_claim_post1, r1 = res
    comm_Az, comm_Bz, comm_Cz, comm_prod = pr["claims_phase2"]
    pok, pp = pr["pok_claims_phase2"]
    if not knowledge_verify(pok, gens.gens_1, t, comm_Cz):
        return None
    if not product_verify(pp, gens.gens_1, t, comm_Az, comm_Bz, comm_prod):
        return None
    t.append_point(b"comm_Az_claim", comm_Az)
    t.append_point(b"comm_Bz_claim", comm_Bz)
    t.append_point(b"comm_Cz_claim", comm_Cz)
    t.append_point(b"comm_prod_Az_Bz_claims", comm_prod)
    rx_rev, rq_rev, rp_round1 = r1[:nrx], r1[nrx:nrx + nrq], r1[nrx + nrq:]
    rx, rq = rx_rev[::-1], rq_rev[::-1]

    def bound(rs, taus):
        acc = ONE
        for r, tau in zip(rs, taus):
            acc = mul(acc, add(mul(r, tau), mul(sub(ONE, r), sub(ONE, tau))))
        return acc

    taus_bound = mul(mul(bound(rp_round1, tau_p), bound(rq_rev, tau_q)), bound(rx_rev, tau_x))
    expected1 = (G.decompress(comm_prod) - G.decompress(comm_Cz)).mul(sint(taus_bound)).compress()
    if not equality_verify(pr["proof_eq_sc_phase1"], gens.gens_1, t, expected1, comm_claim_post1):
        return None
    r_A, r_B, r_C = (t.challenge_scalar(l) for l in (b"challenge_Az", b"challenge_Bz", b"challenge_Cz"))
    comm_claim_phase2 = G.multiscalar_mul([sint(r_A), sint(r_B), sint(r_C)], [G.decompress(c) for c in (comm_Az, comm_Bz, comm_Cz)]).compress()
    res = zk_sumcheck_verify(pr["sc_proof_phase2"], comm_claim_phase2, nry + nrw + nrp, gens.gens_1, gens.gens_4, t)
    if res is None:
        return None
    comm_claim_post2, r2 = res
    ry_rev, rw, rp = r2[:nry], r2[nry:nry + nrw], r2[nry + nrw:]
    ry = ry_rev[::-1]
    p_rp_bound = bound(rp, rp_round1)
    ry_factors = [ONE]
    for i in range(nry):
        ry_factors.append(mul(ry_factors[i], sub(ONE, ry[i])))
    comm_list, npl, nil, Zr_pts = [], [], [], []
    for i in range(W):
        for p in range(len(wit_num_proofs[i])):
            comm_list.append(wit_comms[i][p])
            npl.append(wit_num_proofs[i][p])
            nil.append(wit_num_inputs[i][p])
            Zr_pts.append(G.decompress(pr["comm_vars_at_ry_list"][i][p]))
    if not polyeval_verify_batched(pr["proof_eval_vars_at_ry_list"], npl, nil, gens.pc, t, rq, ry, Zr_pts, comm_list):
        return None
    prefix = _prefix_list(rw, W)
    expected = []
    for p in range(num_instances):
        comb = None
        for i in range(W):
            wp = 0 if len(wit_num_proofs[i]) == 1 else p
            cpt = G.decompress(pr["comm_vars_at_ry_list"][i][wp])
            if wit_num_inputs[i][wp] < max_num_inputs:
                cpt = cpt.mul(sint(ry_factors[nry - log2(wit_num_inputs[i][wp])]))
            term = cpt.mul(sint(prefix[i]))
            comb = term if comb is None else comb + term
        for q in range(nrq - log2(num_proofs[p])):
            comb = comb.mul(sint(sub(ONE, rq[q])))
        expected.append(comb)
    one = ONE.reshape(1, 4)
    EQ_p = (O.eq_evals(np.stack(rp)) if nrp else one)[:num_instances]
    if G.multiscalar_mul([sint(x) for x in EQ_p], expected).compress() != pr["comm_vars_at_ry"]:
        return None
    eA, eB, eC = evals
    s = mul(add(add(mul(r_A, eA), mul(r_B, eB)), mul(r_C, eC)), p_rp_bound)
    expected2 = G.decompress(pr["comm_vars_at_ry"]).mul(sint(s)).compress()
    if not equality_verify(pr["proof_eq_sc_phase2"], gens.gens_1, t, expected2, comm_claim_post2):
        return None
    return [rp, rq_rev, rx, list(rw) + list(ry)]
