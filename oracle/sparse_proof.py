"""ORACLE (test infrastructure, NOT product code).

Restatement of the sparse-polynomial evaluation proof ("Spark" memory check):
  SparseMatPolynomial::multi_commit / multi_sparse_to_dense_rep  src/sparse_mlpoly.rs:354-425, 566-586
  AddrTimestamps                                                 :212-271
  Layers::build_hash_layer / ProductCircuit                      :612-737, src/product_tree.rs:17-64
  ProductCircuitEvalProofBatched::prove / verify                 src/product_tree.rs:260-487
  SumcheckInstanceProof::prove_cubic_batched / verify            src/sumcheck.rs:37-71, 264-434
  ProductLayerProof, HashLayerProof, PolyEvalNetworkProof        :763-1480
  SparseMatPolyEvalProof::prove / verify                         :1482-1603
  PolyEvalProof::prove / verify_plain (single point)             src/dense_mlpoly.rs:437-528
Validated like protocol.py: its own verifier accepts honest proofs and rejects tampered ones
(the reference's live test for this path, sparse_mlpoly.rs:1605-1676, is the same round trip).
"""
from __future__ import annotations

import numpy as np

from . import cbind as O
from . import ristretto as G
from .protocol import (ONE, ZERO, DotProductProofGens, Reader, Transcript, Writer, add, commit1, commitn, dplog_prove,
                       dplog_verify, mul, poly_commit, r_dplog, sint, sub, w_dplog)
from .r1cs import log2, next_pow2

one1 = ONE.reshape(1, 4)


def from_usize(vals):
    return np.stack([O.from_u64(int(v)) for v in vals])


def eq_evals(r):
    return O.eq_evals(np.stack(r)) if len(r) else one1


def dense_eval(Z, r):
    return O.dense_evaluate(Z, np.stack(r)) if len(r) else Z[0]


# ------------------------------------------------------------------ dense representation
class AddrTimestamps:
    def __init__(self, num_cells, num_ops, ops_addr):
        audit = [0] * num_cells
        self.ops_addr_usize = ops_addr
        self.ops_addr, self.read_ts = [], []
        for addrs in ops_addr:
            assert len(addrs) == num_ops
            rts = [0] * num_ops
            for i, a in enumerate(addrs):
                assert a < num_cells
                rts[i] = audit[a]
                audit[a] += 1
            self.ops_addr.append(from_usize(addrs))
            self.read_ts.append(from_usize(rts))
            setattr(self, "_rts", getattr(self, "_rts", []) + [rts])
        self.audit_ts = from_usize(audit)
        self.audit_usize = audit

    def deref(self, mem_val):
        return [mem_val[np.asarray(a, dtype=np.int64)] for a in self.ops_addr_usize]


def merge(polys):
    Z = np.concatenate(polys)
    n = next_pow2(Z.shape[0])
    return np.concatenate([Z, np.zeros((n - Z.shape[0], 4), dtype=np.uint64)])


class MultiSparseDense:
    """multi_sparse_to_dense_rep (:368-425). polys: list of (rows, cols, vals, nvx, nvy)."""

    def __init__(self, polys):
        self.batch_size = len(polys)
        nvx, nvy = polys[0][3], polys[0][4]
        N = max(next_pow2(len(p[0])) for p in polys)
        rows_v, cols_v, self.val = [], [], []
        for rows, cols, vals, _, _ in polys:
            n = len(rows)
            rows_v.append(list(map(int, rows)) + [0] * (N - n))
            cols_v.append(list(map(int, cols)) + [0] * (N - n))
            self.val.append(np.concatenate([O.fq_array(vals).reshape(-1, 4), np.zeros((N - n, 4), dtype=np.uint64)]))
        self.N = N
        self.num_mem_cells = 1 << max(nvx, nvy)
        self.row = AddrTimestamps(self.num_mem_cells, N, rows_v)
        self.col = AddrTimestamps(self.num_mem_cells, N, cols_v)
        self.comb_ops = merge(self.row.ops_addr + self.row.read_ts + self.col.ops_addr + self.col.read_ts + self.val)
        self.comb_mem = np.concatenate([self.row.audit_ts, self.col.audit_ts])


class SparseGens:
    """SparseMatPolyCommitmentGens::new (:289-316)."""

    def __init__(self, label, nvx, nvy, num_nz, batch):
        pcg = lambda nv: DotProductProofGens(1 << (nv - nv // 2), label)
        lg = log2(next_pow2(num_nz))
        self.gens_ops = pcg(lg + log2(next_pow2(batch * 5)))
        self.gens_mem = pcg(max(nvx, nvy) + 1)
        self.gens_derefs = pcg(lg + log2(next_pow2(batch * 2)))


def multi_commit(dense: MultiSparseDense, gens: SparseGens):
    return {"batch_size": dense.batch_size, "num_ops": dense.N, "num_mem_cells": dense.num_mem_cells,
            "comm_comb_ops": poly_commit(dense.comb_ops, gens.gens_ops.gens_n),
            "comm_comb_mem": poly_commit(dense.comb_mem, gens.gens_mem.gens_n)}


def append_poly_commitment(t, label, C):
    t.append_message(label, b"poly_commitment_begin")
    for c in C:
        t.append_point(b"poly_commitment_share", c)
    t.append_message(label, b"poly_commitment_end")


# ------------------------------------------------------------------ PolyEvalProof (single point)
def polyeval_prove(Z, r, Zr, gens: DotProductProofGens, t, tape):
    t.append_protocol_name(b"polynomial evaluation proof")
    left = len(r) // 2
    L, R = eq_evals(r[:left]), eq_evals(r[left:])
    LZ = O.dense_bound_L(Z, L)
    pr, _, C_Zr_prime = dplog_prove(gens, t, tape, list(LZ), ZERO, list(R), Zr, ZERO)
    return pr


def polyeval_verify_plain(pr, gens: DotProductProofGens, t, r, Zr, comm):
    C_Zr = commit1(Zr, ZERO, gens.gens_1).compress()
    t.append_protocol_name(b"polynomial evaluation proof")
    left = len(r) // 2
    L, R = eq_evals(r[:left]), eq_evals(r[left:])
    C_LZ = G.multiscalar_mul([sint(x) for x in L], [G.decompress(c) for c in comm]).compress()
    return dplog_verify(pr, len(R), gens, t, list(R), C_LZ, C_Zr)


# ------------------------------------------------------------------ non-ZK cubic sumcheck
def append_unipoly(t, coeffs):
    t.append_message(b"poly", b"UniPoly_begin")
    for c in coeffs:
        t.append_scalar(b"coeff", c)
    t.append_message(b"poly", b"UniPoly_end")


def cubic_batched_prove(claim, num_rounds, A_par, B_par, C_par, A_seq, B_seq, C_seq, coeffs, t):
    """prove_cubic_batched (src/sumcheck.rs:264-434); tables are bound in place (lists of arrays)."""
    e = claim
    r, polys = [], []
    for _ in range(num_rounds):
        ev = O.cubic_batched_eval(A_par, B_par, C_par if A_par else None, A_seq, B_seq, C_seq, np.stack(coeffs))
        poly = list(O.unipoly_from_evals(np.stack([ev[0], sub(e, ev[0]), ev[1], ev[2]])))
        append_unipoly(t, poly)
        r_j = t.challenge_scalar(b"challenge_nextround")
        r.append(r_j)
        for lst in (A_par, B_par, A_seq, B_seq, C_seq):
            for i in range(len(lst)):
                lst[i] = O.dense_bound_top(lst[i], r_j)
        C_par = O.dense_bound_top(C_par, r_j)
        e = O.unipoly_evaluate(np.stack(poly), r_j)
        polys.append([poly[0], poly[2], poly[3]])  # compress(): drop the linear term
    claims_prod = ([a[0] for a in A_par], [b[0] for b in B_par], C_par[0])
    claims_dotp = ([a[0] for a in A_seq], [b[0] for b in B_seq], [c[0] for c in C_seq])
    return polys, r, claims_prod, claims_dotp


def sumcheck_verify(polys, claim, num_rounds, degree_bound, t):
    """SumcheckInstanceProof::verify (src/sumcheck.rs:37-71)."""
    e = claim
    r = []
    assert len(polys) == num_rounds
    for cp in polys:
        lin = sub(sub(e, cp[0]), cp[0])
        for c in cp[1:]:
            lin = sub(lin, c)
        poly = [cp[0], lin] + list(cp[1:])
        assert len(poly) - 1 == degree_bound
        s = ZERO
        for c in poly:
            s = add(s, c)
        assert np.array_equal(add(poly[0], s), e)  # eval_at_zero + eval_at_one == e
        append_unipoly(t, poly)
        r_i = t.challenge_scalar(b"challenge_nextround")
        r.append(r_i)
        e = O.unipoly_evaluate(np.stack(poly), r_i)
    return e, r


# ------------------------------------------------------------------ product circuits
class ProductCircuit:
    def __init__(self, poly):
        n = poly.shape[0]
        self.left, self.right = [poly[: n // 2].copy()], [poly[n // 2:].copy()]
        for _ in range(log2(n) - 1):
            l, r = O.prod_layer(self.left[-1], self.right[-1])
            self.left.append(l)
            self.right.append(r)

    def evaluate(self):
        return mul(self.left[-1][0], self.right[-1][0])


def pcepb_prove(prod_circuits, dotp_circuits, t):
    """ProductCircuitEvalProofBatched::prove (src/product_tree.rs:260-384).
    dotp_circuits: list of [left, right, weight] arrays."""
    claims_dotp_final = ([], [], [])
    layers = []
    num_layers = len(prod_circuits[0].left)
    claims_to_verify = [c.evaluate() for c in prod_circuits]
    rand = []
    for layer_id in reversed(range(num_layers)):
        length = prod_circuits[0].left[layer_id].shape[0] + prod_circuits[0].right[layer_id].shape[0]
        C_par = eq_evals(rand)
        assert C_par.shape[0] == length // 2
        num_rounds = log2(C_par.shape[0])
        A_par = [c.left[layer_id] for c in prod_circuits]
        B_par = [c.right[layer_id] for c in prod_circuits]
        A_seq, B_seq, C_seq = [], [], []
        if layer_id == 0 and dotp_circuits:
            for d in dotp_circuits:
                claims_to_verify.append(O.dot(O.vec_mul(d[0], d[1]), d[2]))
                assert d[0].shape[0] == length // 2
            A_seq, B_seq, C_seq = [d[0] for d in dotp_circuits], [d[1] for d in dotp_circuits], [d[2] for d in dotp_circuits]
        coeff = t.challenge_vector(b"rand_coeffs_next_layer", len(claims_to_verify))
        claim = ZERO
        for c, k in zip(claims_to_verify, coeff):
            claim = add(claim, mul(c, k))
        polys, rand_prod, claims_prod, claims_dotp = cubic_batched_prove(claim, num_rounds, A_par, B_par, C_par, A_seq, B_seq, C_seq, coeff, t)
        left, right, _ = claims_prod
        for i in range(len(prod_circuits)):
            t.append_scalar(b"claim_prod_left", left[i])
            t.append_scalar(b"claim_prod_right", right[i])
        if layer_id == 0 and dotp_circuits:
            dl, dr, dw = claims_dotp
            for i in range(len(dotp_circuits)):
                t.append_scalar(b"claim_dotp_left", dl[i])
                t.append_scalar(b"claim_dotp_right", dr[i])
                t.append_scalar(b"claim_dotp_weight", dw[i])
            claims_dotp_final = (dl, dr, dw)
        r_layer = t.challenge_scalar(b"challenge_r_layer")
        claims_to_verify = [add(left[i], mul(r_layer, sub(right[i], left[i]))) for i in range(len(prod_circuits))]
        rand = [r_layer] + rand_prod
        layers.append({"polys": polys, "left": left, "right": right})
    return {"proof": layers, "claims_dotp": claims_dotp_final}, rand


def pcepb_verify(pr, claims_prod_vec, claims_dotp_vec, length, t):
    num_layers = log2(length)
    rand = []
    assert len(pr["proof"]) == num_layers
    claims_to_verify = list(claims_prod_vec)
    claims_dotp_out = []
    for num_rounds, i in enumerate(range(num_layers)):
        if i == num_layers - 1:
            claims_to_verify = claims_to_verify + list(claims_dotp_vec)
        coeff = t.challenge_vector(b"rand_coeffs_next_layer", len(claims_to_verify))
        claim = ZERO
        for c, k in zip(claims_to_verify, coeff):
            claim = add(claim, mul(c, k))
        lay = pr["proof"][i]
        claim_last, rand_prod = sumcheck_verify(lay["polys"], claim, num_rounds, 3, t)
        left, right = lay["left"], lay["right"]
        assert len(left) == len(claims_prod_vec) and len(right) == len(claims_prod_vec)
        for a, b in zip(left, right):
            t.append_scalar(b"claim_prod_left", a)
            t.append_scalar(b"claim_prod_right", b)
        assert len(rand) == len(rand_prod)
        eqv = ONE
        for a, b in zip(rand, rand_prod):
            eqv = mul(eqv, add(mul(a, b), mul(sub(ONE, a), sub(ONE, b))))
        expected = ZERO
        for k in range(len(claims_prod_vec)):
            expected = add(expected, mul(coeff[k], mul(mul(left[k], right[k]), eqv)))
        if i == num_layers - 1:
            npi = len(claims_prod_vec)
            dl, dr, dw = pr["claims_dotp"]
            for k in range(len(dl)):
                t.append_scalar(b"claim_dotp_left", dl[k])
                t.append_scalar(b"claim_dotp_right", dr[k])
                t.append_scalar(b"claim_dotp_weight", dw[k])
                expected = add(expected, mul(mul(mul(coeff[k + npi], dl[k]), dr[k]), dw[k]))
        if not np.array_equal(expected, claim_last):
            return None
        r_layer = t.challenge_scalar(b"challenge_r_layer")
        claims_to_verify = [add(left[k], mul(r_layer, sub(right[k], left[k]))) for k in range(len(left))]
        if i == num_layers - 1:
            dl, dr, dw = pr["claims_dotp"]
            for k in range(len(claims_dotp_vec) // 2):
                for v in (dl, dr, dw):
                    claims_dotp_out.append(add(v[2 * k], mul(r_layer, sub(v[2 * k + 1], v[2 * k]))))
        rand = [r_layer] + rand_prod
    return claims_to_verify, claims_dotp_out, rand


# ------------------------------------------------------------------ hash layer
def build_hash_layer(eval_table, addrs_vec, derefs_vec, read_ts_vec, audit_ts, r_hash, r_multiset):
    g2 = mul(r_hash, r_hash)
    n = eval_table.shape[0]

    def hvec(addr, val, ts):
        m = addr.shape[0]
        h = O.vec_add(O.vec_add(O.vec_mul(ts, np.tile(g2, (m, 1))), O.vec_mul(val, np.tile(r_hash, (m, 1)))), addr)
        return O.vec_sub(h, np.tile(r_multiset, (m, 1)))

    idx = from_usize(range(n))
    zeros = np.zeros((n, 4), dtype=np.uint64)
    init = hvec(idx, eval_table, zeros)
    audit = hvec(idx, eval_table, audit_ts)
    reads, writes = [], []
    for addrs, derefs, rts in zip(addrs_vec, derefs_vec, read_ts_vec):
        reads.append(hvec(addrs, derefs, rts))
        wts = O.vec_add(rts, np.tile(ONE, (rts.shape[0], 1)))
        writes.append(hvec(addrs, derefs, wts))
    return init, reads, writes, audit


class Layers:
    def __init__(self, eval_table, at: AddrTimestamps, ops_val, r_mem_check):
        init, reads, writes, audit = build_hash_layer(eval_table, at.ops_addr, ops_val, at.read_ts, at.audit_ts, *r_mem_check)
        self.init, self.audit = ProductCircuit(init), ProductCircuit(audit)
        self.read = [ProductCircuit(x) for x in reads]
        self.write = [ProductCircuit(x) for x in writes]


# ------------------------------------------------------------------ the proof
def _n_to_one(t, label_chal, evals):
    ch = t.challenge_vector(label_chal, log2(len(evals)))
    Z = np.stack(evals)
    for c in reversed(ch):
        Z = O.dense_bound_bot(Z, c)
    assert Z.shape[0] == 1
    return ch, Z[0]


def sparse_prove(dense: MultiSparseDense, rx, ry, evals, gens: SparseGens, t: Transcript, tape):
    """SparseMatPolyEvalProof::prove (:1509-1564)."""
    t.append_protocol_name(b"Sparse polynomial evaluation proof")
    assert len(evals) == dense.batch_size
    rx_ext, ry_ext = equalize(rx, ry)
    mem_rx, mem_ry = eq_evals(rx_ext), eq_evals(ry_ext)
    row_ops_val, col_ops_val = dense.row.deref(mem_rx), dense.col.deref(mem_ry)
    comb = merge(row_ops_val + col_ops_val)
    comm_derefs = poly_commit(comb, gens.gens_derefs.gens_n)
    t.append_message(b"derefs_commitment", b"begin_derefs_commitment")
    append_poly_commitment(t, b"comm_poly_row_col_ops_val", comm_derefs)
    t.append_message(b"derefs_commitment", b"end_derefs_commitment")
    r_mem_check = t.challenge_vector(b"challenge_r_hash", 2)
    row_layers = Layers(mem_rx, dense.row, row_ops_val, r_mem_check)
    col_layers = Layers(mem_ry, dense.col, col_ops_val, r_mem_check)
    # PolyEvalNetworkProof::prove
    t.append_protocol_name(b"Sparse polynomial evaluation proof")
    # ProductLayerProof::prove (:1118-1263)
    t.append_protocol_name(b"Sparse polynomial product layer proof")
    ev = {}
    for name, lay in (("row", row_layers), ("col", col_layers)):
        init, audit = lay.init.evaluate(), lay.audit.evaluate()
        read, write = [c.evaluate() for c in lay.read], [c.evaluate() for c in lay.write]
        ws, rs = ONE, ONE
        for w_, r_ in zip(write, read):
            ws, rs = mul(ws, w_), mul(rs, r_)
        assert np.array_equal(mul(init, ws), mul(rs, audit)), "memory check does not balance"
        t.append_scalar(f"claim_{name}_eval_init".encode(), init)
        t.append_scalars(f"claim_{name}_eval_read".encode(), read)
        t.append_scalars(f"claim_{name}_eval_write".encode(), write)
        t.append_scalar(f"claim_{name}_eval_audit".encode(), audit)
        ev[name] = (init, read, write, audit)
    b = dense.batch_size
    dotp_list, dl_vec, dr_vec = [], [], []
    for i in range(b):
        left, right, weight = row_ops_val[i], col_ops_val[i], dense.val[i]
        h = left.shape[0] // 2
        halves = [[left[:h].copy(), right[:h].copy(), weight[:h].copy()], [left[h:].copy(), right[h:].copy(), weight[h:].copy()]]
        el, er = (O.dot(O.vec_mul(x[0], x[1]), x[2]) for x in halves)
        t.append_scalar(b"claim_eval_dotp_left", el)
        t.append_scalar(b"claim_eval_dotp_right", er)
        assert np.array_equal(add(el, er), evals[i]), "claimed evaluation is wrong"
        dl_vec.append(el)
        dr_vec.append(er)
        dotp_list += halves
    prod_list = row_layers.read + row_layers.write + col_layers.read + col_layers.write
    proof_ops, rand_ops = pcepb_prove(prod_list, dotp_list, t)
    proof_mem, rand_mem = pcepb_prove([row_layers.init, row_layers.audit, col_layers.init, col_layers.audit], [], t)
    prod_layer = {"eval_row": ev["row"], "eval_col": ev["col"], "eval_val": (dl_vec, dr_vec), "proof_mem": proof_mem, "proof_ops": proof_ops}
    # HashLayerProof::prove (:827-918)
    t.append_protocol_name(b"Sparse polynomial hash layer proof")
    e_row_val = [dense_eval(x, rand_ops) for x in row_ops_val]
    e_col_val = [dense_eval(x, rand_ops) for x in col_ops_val]
    t.append_protocol_name(b"Derefs evaluation proof")
    evs = e_row_val + e_col_val
    evs += [ZERO] * (next_pow2(len(evs)) - len(evs))
    t.append_scalars(b"evals_ops_val", evs)
    ch, joint = _n_to_one(t, b"challenge_combine_n_to_one", evs)
    r_joint = ch + list(rand_ops)
    t.append_scalar(b"joint_claim_eval", joint)
    proof_derefs = polyeval_prove(comb, r_joint, joint, gens.gens_derefs, t, tape)

    def helper(at):
        return ([dense_eval(x, rand_ops) for x in at.ops_addr], [dense_eval(x, rand_ops) for x in at.read_ts],
                dense_eval(at.audit_ts, rand_mem))

    e_row, e_col = helper(dense.row), helper(dense.col)
    e_val = [dense_eval(x, rand_ops) for x in dense.val]
    evals_ops = e_row[0] + e_row[1] + e_col[0] + e_col[1] + e_val
    evals_ops += [ZERO] * (next_pow2(len(evals_ops)) - len(evals_ops))
    t.append_scalars(b"claim_evals_ops", evals_ops)
    ch, joint_ops = _n_to_one(t, b"challenge_combine_n_to_one", evals_ops)
    r_joint_ops = ch + list(rand_ops)
    t.append_scalar(b"joint_claim_eval_ops", joint_ops)
    proof_ops_eval = polyeval_prove(dense.comb_ops, r_joint_ops, joint_ops, gens.gens_ops, t, tape)
    evals_mem = [e_row[2], e_col[2]]
    t.append_scalars(b"claim_evals_mem", evals_mem)
    ch, joint_mem = _n_to_one(t, b"challenge_combine_two_to_one", evals_mem)
    r_joint_mem = ch + list(rand_mem)
    t.append_scalar(b"joint_claim_eval_mem", joint_mem)
    proof_mem_eval = polyeval_prove(dense.comb_mem, r_joint_mem, joint_mem, gens.gens_mem, t, tape)
    hash_layer = {"eval_row": e_row, "eval_col": e_col, "eval_val": e_val, "eval_derefs": (e_row_val, e_col_val),
                  "proof_ops": proof_ops_eval, "proof_mem": proof_mem_eval, "proof_derefs": proof_derefs}
    return {"comm_derefs": comm_derefs, "prod_layer": prod_layer, "hash_layer": hash_layer}


def equalize(rx, ry):
    rx, ry = list(rx), list(ry)
    if len(rx) < len(ry):
        rx = [ZERO] * (len(ry) - len(rx)) + rx
    elif len(ry) < len(rx):
        ry = [ZERO] * (len(rx) - len(ry)) + ry
    return rx, ry


def _hash_verify_helper(rand_mem, claims, e_ops_val, e_addr, e_rts, e_audit, r, r_hash, r_multiset):
    g2 = mul(r_hash, r_hash)
    hf = lambda a, v, ts: add(add(mul(ts, g2), mul(v, r_hash)), a)
    claim_init, claim_read, claim_write, claim_audit = claims
    n = len(rand_mem)
    init_addr = ZERO
    for i in range(n):  # IdentityPolynomial::evaluate (src/dense_mlpoly.rs:142-148)
        init_addr = add(init_addr, mul(O.from_u64(1 << (n - i - 1)), rand_mem[i]))
    init_val = O.eq_evaluate(np.stack(r), np.stack(rand_mem)) if n else ONE
    if not np.array_equal(sub(hf(init_addr, init_val, ZERO), r_multiset), claim_init):
        return False
    for i in range(len(e_addr)):
        if not np.array_equal(sub(hf(e_addr[i], e_ops_val[i], e_rts[i]), r_multiset), claim_read[i]):
            return False
        if not np.array_equal(sub(hf(e_addr[i], e_ops_val[i], add(e_rts[i], ONE)), r_multiset), claim_write[i]):
            return False
    return np.array_equal(sub(hf(init_addr, init_val, e_audit), r_multiset), claim_audit)


def sparse_verify(pr, comm, rx, ry, evals, gens: SparseGens, t: Transcript):
    """SparseMatPolyEvalProof::verify (:1566-1602) and everything below it."""
    t.append_protocol_name(b"Sparse polynomial evaluation proof")
    rx_ext, ry_ext = equalize(rx, ry)
    nz, num_cells = comm["num_ops"], comm["num_mem_cells"]
    assert 1 << len(rx_ext) == num_cells
    t.append_message(b"derefs_commitment", b"begin_derefs_commitment")
    append_poly_commitment(t, b"comm_poly_row_col_ops_val", pr["comm_derefs"])
    t.append_message(b"derefs_commitment", b"end_derefs_commitment")
    r_hash, r_multiset = t.challenge_vector(b"challenge_r_hash", 2)
    t.append_protocol_name(b"Sparse polynomial evaluation proof")
    b = len(evals)
    num_ops = next_pow2(nz)
    # ProductLayerProof::verify
    pl = pr["prod_layer"]
    t.append_protocol_name(b"Sparse polynomial product layer proof")
    for name in ("row", "col"):
        init, read, write, audit = pl[f"eval_{name}"]
        if len(read) != b or len(write) != b:
            return False
        ws, rs = ONE, ONE
        for w_, r_ in zip(write, read):
            ws, rs = mul(ws, w_), mul(rs, r_)
        if not np.array_equal(mul(init, ws), mul(rs, audit)):
            return False
        t.append_scalar(f"claim_{name}_eval_init".encode(), init)
        t.append_scalars(f"claim_{name}_eval_read".encode(), read)
        t.append_scalars(f"claim_{name}_eval_write".encode(), write)
        t.append_scalar(f"claim_{name}_eval_audit".encode(), audit)
    dl, dr = pl["eval_val"]
    claims_dotp_circuit = []
    for i in range(b):
        if not np.array_equal(add(dl[i], dr[i]), evals[i]):
            return False
        t.append_scalar(b"claim_eval_dotp_left", dl[i])
        t.append_scalar(b"claim_eval_dotp_right", dr[i])
        claims_dotp_circuit += [dl[i], dr[i]]
    ri, rr, rw, ra = pl["eval_row"]
    ci, cr, cw, ca = pl["eval_col"]
    res = pcepb_verify(pl["proof_ops"], list(rr) + list(rw) + list(cr) + list(cw), claims_dotp_circuit, num_ops, t)
    if res is None:
        return False
    claims_ops, claims_dotp, rand_ops = res
    res = pcepb_verify(pl["proof_mem"], [ri, ra, ci, ca], [], num_cells, t)
    if res is None:
        return False
    claims_mem, _, rand_mem = res
    # HashLayerProof::verify
    hl = pr["hash_layer"]
    t.append_protocol_name(b"Sparse polynomial hash layer proof")
    e_row_val, e_col_val = hl["eval_derefs"]
    t.append_protocol_name(b"Derefs evaluation proof")
    evs = list(e_row_val) + list(e_col_val)
    evs += [ZERO] * (next_pow2(len(evs)) - len(evs))
    t.append_scalars(b"evals_ops_val", evs)
    ch, joint = _n_to_one(t, b"challenge_combine_n_to_one", evs)
    t.append_scalar(b"joint_claim_eval", joint)
    if not polyeval_verify_plain(hl["proof_derefs"], gens.gens_derefs, t, ch + list(rand_ops), joint, pr["comm_derefs"]):
        return False
    e_val = hl["eval_val"]
    if len(claims_dotp) != 3 * len(e_row_val):
        return False
    for i in range(len(e_row_val)):
        if not (np.array_equal(claims_dotp[3 * i], e_row_val[i]) and np.array_equal(claims_dotp[3 * i + 1], e_col_val[i])
                and np.array_equal(claims_dotp[3 * i + 2], e_val[i])):
            return False
    e_row, e_col = hl["eval_row"], hl["eval_col"]
    evals_ops = list(e_row[0]) + list(e_row[1]) + list(e_col[0]) + list(e_col[1]) + list(e_val)
    evals_ops += [ZERO] * (next_pow2(len(evals_ops)) - len(evals_ops))
    t.append_scalars(b"claim_evals_ops", evals_ops)
    ch, joint_ops = _n_to_one(t, b"challenge_combine_n_to_one", evals_ops)
    t.append_scalar(b"joint_claim_eval_ops", joint_ops)
    if not polyeval_verify_plain(hl["proof_ops"], gens.gens_ops, t, ch + list(rand_ops), joint_ops, comm["comm_comb_ops"]):
        return False
    evals_mem = [e_row[2], e_col[2]]
    t.append_scalars(b"claim_evals_mem", evals_mem)
    ch, joint_mem = _n_to_one(t, b"challenge_combine_two_to_one", evals_mem)
    t.append_scalar(b"joint_claim_eval_mem", joint_mem)
    if not polyeval_verify_plain(hl["proof_mem"], gens.gens_mem, t, ch + list(rand_mem), joint_mem, comm["comm_comb_mem"]):
        return False
    claims_row = (claims_mem[0], claims_ops[:b], claims_ops[b:2 * b], claims_mem[1])
    claims_col = (claims_mem[2], claims_ops[2 * b:3 * b], claims_ops[3 * b:4 * b], claims_mem[3])
    if not _hash_verify_helper(rand_mem, claims_row, e_row_val, e_row[0], e_row[1], e_row[2], rx_ext, r_hash, r_multiset):
        return False
    return bool(_hash_verify_helper(rand_mem, claims_col, e_col_val, e_col[0], e_col[1], e_col[2], ry_ext, r_hash, r_multiset))


# ------------------------------------------------------------------ bincode layout
def _w_pcepb(w: Writer, pr):
    w.u64(len(pr["proof"]))
    for lay in pr["proof"]:
        w.u64(len(lay["polys"]))
        for cp in lay["polys"]:
            w.scalars(cp)
        w.scalars(lay["left"])
        w.scalars(lay["right"])
    for v in pr["claims_dotp"]:
        w.scalars(v)


def serialize_sparse_proof(pr) -> bytes:
    w = Writer()
    w.points(pr["comm_derefs"])
    pl = pr["prod_layer"]
    for name in ("eval_row", "eval_col"):
        init, read, write, audit = pl[name]
        w.scalar(init)
        w.scalars(read)
        w.scalars(write)
        w.scalar(audit)
    w.scalars(pl["eval_val"][0])
    w.scalars(pl["eval_val"][1])
    _w_pcepb(w, pl["proof_mem"])
    _w_pcepb(w, pl["proof_ops"])
    hl = pr["hash_layer"]
    for name in ("eval_row", "eval_col"):
        a, r_, au = hl[name]
        w.scalars(a)
        w.scalars(r_)
        w.scalar(au)
    w.scalars(hl["eval_val"])
    w.scalars(hl["eval_derefs"][0])
    w.scalars(hl["eval_derefs"][1])
    w_dplog(w, hl["proof_ops"])
    w_dplog(w, hl["proof_mem"])
    w_dplog(w, hl["proof_derefs"])
    return bytes(w.b)


def _r_pcepb(r: Reader):
    layers = []
    for _ in range(r.u64()):
        polys = [r.scalars() for _ in range(r.u64())]
        layers.append({"polys": polys, "left": r.scalars(), "right": r.scalars()})
    return {"proof": layers, "claims_dotp": (r.scalars(), r.scalars(), r.scalars())}


def deserialize_sparse_proof(b: bytes):
    r = Reader(b)
    pr = {"comm_derefs": r.points()}
    pl = {}
    for name in ("eval_row", "eval_col"):
        pl[name] = (r.scalar(), r.scalars(), r.scalars(), r.scalar())
    pl["eval_val"] = (r.scalars(), r.scalars())
    pl["proof_mem"] = _r_pcepb(r)
    pl["proof_ops"] = _r_pcepb(r)
    hl = {}
    for name in ("eval_row", "eval_col"):
        hl[name] = (r.scalars(), r.scalars(), r.scalar())
    hl["eval_val"] = r.scalars()
    hl["eval_derefs"] = (r.scalars(), r.scalars())
    hl["proof_ops"] = r_dplog(r)
    hl["proof_mem"] = r_dplog(r)
    hl["proof_derefs"] = r_dplog(r)
    assert r.pos == len(b)
    pr["prod_layer"], pr["hash_layer"] = pl, hl
    return pr
