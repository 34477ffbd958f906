"""ORACLE (test infrastructure, NOT product code).

Restates the table pipeline of ``R1CSProof::prove`` (/root/reference/src/r1csproof.rs:
210-685) with the Fiat-Shamir challenges injected by the caller, so that every stage
can be compared with the CUDA path in isolation:

  z_mat assembly (:278-293) -> multiply_vec_block (src/r1csinstance.rs:363-436)
  -> phase-1 sumcheck loops (src/sumcheck.rs:1067-1380)
  -> ABC table (:431-465) -> Z_poly bound to rq (:469-479) -> phase-2 loops
  (src/sumcheck.rs:788-1065) -> witness evaluations (:534-573).

The heavy loops run in the C restatement (oracle/*.c) through ``cbind``.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import cbind as O


def log2(n: int) -> int:
    return n.bit_length() - 1


def next_pow2(n: int) -> int:
    return 1 if n <= 1 else 1 << (n - 1).bit_length()


@dataclass
class Instance:
    """R1CSInstance (src/r1csinstance.rs:19-31); matrices as COO (row, col, val) arrays."""

    num_instances: int
    max_num_cons: int
    num_cons: list
    num_vars: int
    mats: list  # 3 * num_instances entries of (rows u32, cols u32, vals (n,4) u64); A, B, C per instance


@dataclass
class WitnessSec:
    """ProverWitnessSecInfo: w_mat[p][q] is a (num_inputs[p], 4) array."""

    num_inputs: list
    w_mat: list  # list over p of list over q of arrays

    def poly_w(self, p):
        return np.concatenate(self.w_mat[p])


def build_z_mat(num_instances, num_proofs, num_inputs, witness_secs):
    """z_mat[p][q][w][i] (src/r1csproof.rs:278-293) as a list over p of (Q_p, W, Y_p, 4) arrays."""
    out = []
    W = len(witness_secs)
    for p in range(num_instances):
        z = np.zeros((num_proofs[p], W, num_inputs[p], 4), dtype=np.uint64)
        for q in range(num_proofs[p]):
            for w, ws in enumerate(witness_secs):
                p_w = 0 if len(ws.w_mat) == 1 else p
                q_w = 0 if len(ws.w_mat[p_w]) == 1 else q
                n = min(ws.num_inputs[p_w], num_inputs[p])
                z[q, w, :n] = ws.w_mat[p_w][q_w][:n]
        out.append(z)
    return out


def multiply_vec_block(inst: Instance, num_instances, num_proofs, max_num_inputs, num_cons, z_mat):
    """Az, Bz, Cz in NATURAL ragged [p][q][x] order, each a flat (sum Q_p X_p, 4) array
    (src/r1csinstance.rs:363-411 before new_rev)."""
    outs = [[], [], []]
    for p in range(num_instances):
        pi = 0 if inst.num_instances == 1 else p
        zp = np.ascontiguousarray(z_mat[p])  # (Q_p, W, Y_p, 4): the proofs are independent
        Qp, Wn, Yp = zp.shape[0], zp.shape[1], zp.shape[2]
        for m in range(3):
            rows, cols, vals = inst.mats[3 * pi + m]
            outs[m].append(O.spmv_batch(rows, cols, vals, num_cons[pi], max_num_inputs, zp.reshape(Qp, Wn * Yp, 4), Yp))
    return [np.concatenate(o) for o in outs]


def abc_table(inst: Instance, num_witness_secs, max_num_inputs, num_inputs, evals_rx, r_A, r_B, r_C):
    """evals_ABC[p_inst] as (W, Y_p, 4) arrays (src/r1csproof.rs:431-456)."""
    out = []
    for p in range(inst.num_instances):
        tabs = []
        for m in range(3):
            rows, cols, vals = inst.mats[3 * p + m]
            tabs.append(O.eval_table_sparse(rows, cols, vals, evals_rx, num_witness_secs, max_num_inputs, num_inputs[p]))
        shp = tabs[0].shape
        flat = [t.reshape(-1, 4) for t in tabs]
        n = flat[0].shape[0]
        comb = O.vec_add(O.vec_add(O.vec_mul(np.tile(r_A, (n, 1)), flat[0]), O.vec_mul(np.tile(r_B, (n, 1)), flat[1])),
                         O.vec_mul(np.tile(r_C, (n, 1)), flat[2]))
        out.append(comb.reshape(shp))
    return out


@dataclass
class Trace:
    Az: np.ndarray = None
    Bz: np.ndarray = None
    Cz: np.ndarray = None
    evals1: list = field(default_factory=list)
    claims1: np.ndarray = None
    evals2: list = field(default_factory=list)
    claims2: np.ndarray = None
    rx: np.ndarray = None
    rq_rev: np.ndarray = None
    rp: np.ndarray = None
    ry: np.ndarray = None
    rw: np.ndarray = None
    rp2: np.ndarray = None


def prove_tables(inst: Instance, num_instances, max_num_proofs, num_proofs, max_num_inputs, num_inputs, witness_secs,
                 tau_p, tau_q, tau_x, ch1, r_abc, ch2) -> Trace:
    """Runs both sumchecks with injected challenges. ch1 / ch2: per-round challenges."""
    t = Trace()
    W = len(witness_secs)
    num_cons = inst.max_num_cons
    block_num_cons = [inst.num_cons[0]] * num_instances if inst.num_instances == 1 else list(inst.num_cons)
    z_mat = build_z_mat(num_instances, num_proofs, num_inputs, witness_secs)
    np_, nq, nx = log2(next_pow2(num_instances)), log2(max_num_proofs), log2(num_cons)
    nw, ny = log2(next_pow2(W)), log2(max_num_inputs)
    one = O.ONE.reshape(1, 4)
    Ap = O.eq_evals(tau_p) if np_ else one
    Aq = O.eq_evals(tau_q) if nq else one
    Ax = O.eq_evals(tau_x) if nx else one
    t.Az, t.Bz, t.Cz = multiply_vec_block(inst, num_instances, num_proofs, max_num_inputs, block_num_cons, z_mat)
    mk = lambda T: O.Pqx.new_rev(T, 1, num_proofs, max_num_proofs, block_num_cons, num_cons)
    sc1 = O.Sc1(nx, nq, np_, num_proofs, block_num_cons, Ap, Aq, Ax, mk(t.Az), mk(t.Bz), mk(t.Cz))
    for j in range(sc1.num_rounds):
        t.evals1.append(sc1.round_eval())
        sc1.round_bind(ch1[j])
    t.claims1 = sc1.final()
    r = np.asarray(ch1, dtype=np.uint64).reshape(-1, 4)
    rx_rev, rq_rev, rp = r[:nx], r[nx:nx + nq], r[nx + nq:nx + nq + np_]
    t.rx = rx_rev[::-1].copy()
    t.rq_rev, t.rp = rq_rev.copy(), rp.copy()
    r_A, r_B, r_C = r_abc

    evals_rx = O.eq_evals(t.rx) if nx else one
    abc = abc_table(inst, W, max_num_inputs, num_inputs, evals_rx, r_A, r_B, r_C)
    abc_flat = np.concatenate([a.reshape(-1, 4) for a in abc])
    n_abc = inst.num_instances
    ABC = O.Pqx.new_rev(abc_flat, W, [1] * n_abc, 1, list(num_inputs[:n_abc]), max_num_inputs)
    z_flat = np.concatenate([z.reshape(-1, 4) for z in z_mat])
    Z = O.Pqx.new_rev(z_flat, W, num_proofs, max_num_proofs, num_inputs, max_num_inputs)
    for rq in rq_rev:
        Z.bound_poly(rq, O.MODE_Q)
    eq_p = O.eq_evals(rp) if np_ else one
    sc2 = O.Sc2(ny, nw, np_, inst.num_instances == 1, W, num_inputs, eq_p, ABC, Z)
    for j in range(sc2.num_rounds):
        t.evals2.append(sc2.round_eval())
        sc2.round_bind(ch2[j])
    t.claims2 = sc2.final()
    r2 = np.asarray(ch2, dtype=np.uint64).reshape(-1, 4)
    t.ry = r2[:ny][::-1].copy()
    t.rw, t.rp2 = r2[ny:ny + nw].copy(), r2[ny + nw:ny + nw + np_].copy()
    return t


def synthetic_instance(X: int, num_instances: int = 1, seed: int = 0, unit: bool = True) -> Instance:
    """variables u_0..u_{X-1} | v_0..v_{X-1}; constraint x: u_x * u_{(x+1)%X} = v_x.
    With unit=False the coefficients are random (a_x u_x)(b_x u_{x+1}) = (a_x b_x) v_x."""
    mats = []
    rows = np.arange(X, dtype=np.uint32)
    for i in range(num_instances):
        if unit:
            a = b = c = np.tile(O.ONE, (X, 1))
        else:
            rng = np.random.default_rng(seed * 1000 + i)
            a = O.vec_from_u512(rng.integers(0, 1 << 64, size=(X, 8), dtype=np.uint64))
            b = O.vec_from_u512(rng.integers(0, 1 << 64, size=(X, 8), dtype=np.uint64))
            c = O.vec_mul(a, b)
        mats.append((rows, rows.copy(), a))
        mats.append((rows, ((rows + 1) % X).astype(np.uint32), b))
        mats.append((rows, (rows + X).astype(np.uint32), c))
    return Instance(num_instances, X, [X] * num_instances, 2 * X, mats)


def synthetic_witness(X: int, num_proofs: list, seed: int):
    """Two witness sections (u, v) for synthetic_instance; u uniform, v_x = u_x u_{x+1}."""
    us, vs = [], []
    for p, Q in enumerate(num_proofs):
        rng = np.random.default_rng(seed * 7919 + p)
        up, vp = [], []
        for _ in range(Q):
            u = O.vec_from_u512(rng.integers(0, 1 << 64, size=(X, 8), dtype=np.uint64))
            v = O.vec_mul(u, np.roll(u, -1, axis=0))
            up.append(u)
            vp.append(v)
        us.append(up)
        vs.append(vp)
    P = len(num_proofs)
    return [WitnessSec([X] * P, us), WitnessSec([X] * P, vs)]
