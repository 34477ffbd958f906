"""Host-side mirror of the reference interface for the accelerated path.

Names follow the reference crate (scroll-tech/spartan-parallel): ``EqPolynomial``,
``DensePolynomial``, ``R1CSInstance``, the two disjoint-round sumcheck provers and
``prove_cubic_batched``. Scalars are numpy ``uint64[..., 4]`` arrays holding the
reference's Montgomery limbs (src/scalar/ristretto255.rs:193-199). Every call goes
through the C ABI in include/spgpu.h; nothing here computes field arithmetic on
the CPU.
"""
from __future__ import annotations

import ctypes as C
from typing import Sequence

import numpy as np

from . import _lib
from ._lib import SpgError, check

MODE_P, MODE_Q, MODE_W, MODE_X = 1, 2, 3, 4
# Scalar::one() = R mod q (src/scalar/ristretto255.rs:307-312)
ONE = np.array([0xD6EC31748D98951D, 0xC6EF5BF4737DCF70, 0xFFFFFFFFFFFFFFFE, 0x0FFFFFFFFFFFFFFF], dtype=np.uint64)


def _fq(a) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint64)
    if a.shape[-1] != 4:
        raise ValueError(f"scalar arrays must have a trailing axis of 4 limbs, got {a.shape}")
    return a


def _ptr(a: np.ndarray | None):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _sz(v) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(v, dtype=np.uint64).reshape(-1))


class Context:
    """One device + stream (spg_ctx). The reference prover is single-threaded; so is this."""

    def __init__(self, device: int = 0):
        self.L = _lib.lib()
        h = C.c_void_p()
        check(self.L.spg_ctx_create(device, C.byref(h)), "spg_ctx_create")
        self.h = h
        self.device = device

    def sync(self):
        check(self.L.spg_ctx_sync(self.h), "spg_ctx_sync")

    def profile_begin(self):
        check(self.L.spg_ctx_profile_begin(self.h), "spg_ctx_profile_begin")

    def profile_end(self) -> list:
        import json

        buf = C.create_string_buffer(1 << 16)
        check(self.L.spg_ctx_profile_end(self.h, buf, len(buf)), "spg_ctx_profile_end")
        return json.loads(buf.value.decode())

    @property
    def launches(self) -> int:
        return int(self.L.spg_ctx_launch_count(self.h))

    @property
    def stream(self) -> int:
        return int(self.L.spg_ctx_stream(self.h) or 0)

    def close(self):
        if getattr(self, "h", None):
            self.L.spg_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DensePolynomial:
    """Device-resident DensePolynomial (src/dense_mlpoly.rs:19-24)."""

    def __init__(self, ctx: Context, handle, owner=None):
        self.ctx, self.h, self._owner = ctx, handle, owner

    @classmethod
    def new(cls, ctx: Context, Z) -> "DensePolynomial":
        Z = _fq(Z).reshape(-1, 4)
        n = Z.shape[0]
        n2 = 1 if n == 0 else 1 << (n - 1).bit_length()
        if n2 != n:  # DensePolynomial::new zero-pads to a power of two (:152-161)
            Z = np.concatenate([Z, np.zeros((n2 - n, 4), dtype=np.uint64)])
        h = C.c_void_p()
        check(ctx.L.spg_vec_upload(ctx.h, _ptr(Z), Z.shape[0], C.byref(h)), "spg_vec_upload")
        return cls(ctx, h)

    @classmethod
    def wrap(cls, ctx: Context, device_ptr: int, n: int, owner=None) -> "DensePolynomial":
        h = C.c_void_p()
        check(ctx.L.spg_vec_wrap(ctx.h, C.c_void_p(device_ptr), n, C.byref(h)), "spg_vec_wrap")
        return cls(ctx, h, owner)

    @classmethod
    def empty(cls, ctx: Context, n: int) -> "DensePolynomial":
        h = C.c_void_p()
        check(ctx.L.spg_vec_alloc(ctx.h, n, C.byref(h)), "spg_vec_alloc")
        return cls(ctx, h)

    def __len__(self):
        return int(self.ctx.L.spg_vec_len(self.h))

    def len(self):
        return len(self)

    def get_num_vars(self):
        return len(self).bit_length() - 1

    @property
    def device_ptr(self) -> int:
        return int(self.ctx.L.spg_vec_device_ptr(self.h) or 0)

    def zero(self):
        check(self.ctx.L.spg_vec_zero(self.ctx.h, self.h), "spg_vec_zero")

    def to_host(self) -> np.ndarray:
        out = np.empty((len(self), 4), dtype=np.uint64)
        check(self.ctx.L.spg_vec_download(self.ctx.h, self.h, 0, len(self), _ptr(out)), "spg_vec_download")
        return out

    def bound_poly_var_top(self, r):
        check(self.ctx.L.spg_dense_bound_top(self.ctx.h, self.h, _ptr(_fq(r))), "spg_dense_bound_top")

    def bound_poly_var_bot(self, r):
        check(self.ctx.L.spg_dense_bound_bot(self.ctx.h, self.h, _ptr(_fq(r))), "spg_dense_bound_bot")

    def evaluate(self, r) -> np.ndarray:
        r = _fq(np.asarray(r, dtype=np.uint64).reshape(-1, 4))
        out = np.empty(4, dtype=np.uint64)
        check(self.ctx.L.spg_dense_evaluate(self.ctx.h, self.h, _ptr(r), r.shape[0], _ptr(out)), "spg_dense_evaluate")
        return out

    def bound(self, L) -> "DensePolynomial":
        L = _fq(L).reshape(-1, 4)
        h = C.c_void_p()
        check(self.ctx.L.spg_dense_bound_L(self.ctx.h, self.h, _ptr(L), L.shape[0], C.byref(h)), "spg_dense_bound_L")
        return DensePolynomial(self.ctx, h)

    def free(self):
        if getattr(self, "h", None):
            self.ctx.L.spg_vec_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def vec_op(ctx: Context, op: str, a: DensePolynomial, b: DensePolynomial | None = None) -> DensePolynomial:
    """Scalar::{mul,add,sub,neg,square,to_bytes} elementwise on the device."""
    code = {"mul": 0, "add": 1, "sub": 2, "neg": 3, "square": 4, "to_canonical": 5, "invert": 6}[op]
    out = DensePolynomial.empty(ctx, len(a))
    check(ctx.L.spg_fq_vec_op(ctx.h, code, a.h, b.h if b is not None else None, out.h), "spg_fq_vec_op")
    return out


def perm_scan(ctx: Context, w3: DensePolynomial, seg_len, width: int = 8, v_col: int = 0, x_col: int = 1,
              pi_col: int = 2, d_col: int = 3) -> DensePolynomial:
    """Fills the (pi, D) columns of a row-major w3 table in place (src/lib.rs:1378-1400, 862-880,
    1533-1570): D[q] = x[q] * (pi[q+1] + 1 - v[q+1]), pi[q] = v[q] * D[q], from the last row of
    every segment (proving instance) upwards. ``seg_len`` lists the rows per segment."""
    seg = np.ascontiguousarray(seg_len, dtype=np.uint64)
    n = int(seg.sum())
    assert len(w3) >= n * width
    check(ctx.L.spg_perm_scan(ctx.h, n, _ptr(seg), seg.size, w3.h, v_col, width, w3.h, x_col, width, w3.h, d_col, width,
                              w3.h, pi_col, width), "spg_perm_scan")
    return w3


def wit_perm_w0(ctx: Context, tau, r, used: int, total: int) -> DensePolynomial:
    """perm_w0 = (tau, r, r^2, ...) (src/lib.rs:1328-1338)."""
    h = C.c_void_p()
    check(ctx.L.spg_wit_perm_w0(ctx.h, _ptr(_fq(tau)), _ptr(_fq(r)), used, total, C.byref(h)), "spg_wit_perm_w0")
    return DensePolynomial(ctx, h)


def wit_block(ctx: Context, vars_: DensePolynomial, rows: int, vars_width: int, perm_w0: DensePolynomial, tau, r,
              num_inputs_unpadded: int, io_width: int = 0, phy_ops: int = 0, vir_ops: int = 0, w2_width: int | None = None,
              seg_len=None, exec_mode: bool = False):
    """block_w2 / block_w3 of one instance (src/lib.rs:1511-1613), or with exec_mode perm_exec_w2 /
    perm_exec_w3 (src/lib.rs:1346-1400). Returns (w2, w3) as device polynomials (row-major)."""
    seg = _sz([rows] if seg_len is None else seg_len)
    a, b = C.c_void_p(), C.c_void_p()
    check(ctx.L.spg_wit_block(ctx.h, int(exec_mode), vars_.h, rows, vars_width, perm_w0.h, _ptr(_fq(tau)), _ptr(_fq(r)),
                              num_inputs_unpadded, io_width, phy_ops, vir_ops, w2_width, _ptr(seg), seg.size, C.byref(a), C.byref(b)),
          "spg_wit_block")
    return DensePolynomial(ctx, a), DensePolynomial(ctx, b)


def wit_mem(ctx: Context, mems: DensePolynomial, rows: int, in_width: int, tau, r, mem_width: int):
    """mem_gen's w2 and w3 (src/lib.rs:832-880)."""
    a, b = C.c_void_p(), C.c_void_p()
    check(ctx.L.spg_wit_mem(ctx.h, mems.h, rows, in_width, _ptr(_fq(tau)), _ptr(_fq(r)), mem_width, C.byref(a), C.byref(b)), "spg_wit_mem")
    return DensePolynomial(ctx, a), DensePolynomial(ctx, b)


def wit_shift(ctx: Context, w3: DensePolynomial, rows: int, width: int = 8, seg_len=None) -> DensePolynomial:
    """w3_shifted (src/lib.rs:1667-1676): per instance, rows 1.. followed by a zero row."""
    seg = _sz([rows] if seg_len is None else seg_len)
    h = C.c_void_p()
    check(ctx.L.spg_wit_shift(ctx.h, w3.h, rows, width, _ptr(seg), seg.size, C.byref(h)), "spg_wit_shift")
    return DensePolynomial(ctx, h)


def from_u512(ctx: Context, wide) -> DensePolynomial:
    wide = np.ascontiguousarray(wide, dtype=np.uint64)
    assert wide.shape[-1] == 8
    h = C.c_void_p()
    check(ctx.L.spg_fq_from_u512(ctx.h, _ptr(wide), wide.size // 8, C.byref(h)), "spg_fq_from_u512")
    return DensePolynomial(ctx, h)


def dot(ctx: Context, a: DensePolynomial, b: DensePolynomial) -> np.ndarray:
    out = np.empty(4, dtype=np.uint64)
    check(ctx.L.spg_dot(ctx.h, a.h, b.h, _ptr(out)), "spg_dot")
    return out


class EqPolynomial:
    """EqPolynomial (src/dense_mlpoly.rs:60-131)."""

    def __init__(self, ctx: Context, r):
        self.ctx = ctx
        self.r = _fq(np.asarray(r, dtype=np.uint64).reshape(-1, 4))

    def evals(self) -> DensePolynomial:
        h = C.c_void_p()
        check(self.ctx.L.spg_eq_evals(self.ctx.h, _ptr(self.r), self.r.shape[0], C.byref(h)), "spg_eq_evals")
        return DensePolynomial(self.ctx, h)

    @staticmethod
    def compute_factored_lens(ell: int):
        return ell // 2, ell - ell // 2

    def compute_factored_evals(self):
        left, _ = self.compute_factored_lens(self.r.shape[0])
        return EqPolynomial(self.ctx, self.r[:left]).evals(), EqPolynomial(self.ctx, self.r[left:]).evals()


class SumcheckPhase1:
    """Device loops of prove_cubic_with_additive_term_disjoint_rounds (src/sumcheck.rs:1067-1380)."""

    def __init__(self, ctx: Context, handle):
        self.ctx, self.h = ctx, handle

    @classmethod
    def from_tables(cls, ctx: Context, num_proofs, max_num_proofs, num_cons, max_num_cons, Az, Bz, Cz, tau_p, tau_q, tau_x):
        npf, nc = _sz(num_proofs), _sz(num_cons)
        h = C.c_void_p()
        tp, tq, tx = (_fq(np.asarray(t, dtype=np.uint64).reshape(-1, 4)) for t in (tau_p, tau_q, tau_x))
        check(ctx.L.spg_sc1_create_from_tables(ctx.h, len(npf), _ptr(npf), max_num_proofs, _ptr(nc), max_num_cons,
                                               _ptr(_fq(Az)), _ptr(_fq(Bz)), _ptr(_fq(Cz)), _ptr(tp), _ptr(tq), _ptr(tx),
                                               C.byref(h)), "spg_sc1_create_from_tables")
        return cls(ctx, h)

    @property
    def num_rounds(self) -> int:
        return int(self.ctx.L.spg_sc1_num_rounds(self.h))

    def set_scale(self, c):
        """Multiply every evaluation and the eq claim by c (shards of the proof axis)."""
        check(self.ctx.L.spg_sc1_set_scale(self.h, _ptr(_fq(c))), "spg_sc1_set_scale")

    def run_rounds(self, challenges) -> np.ndarray:
        """eval + bind for len(challenges) rounds with challenges known in advance; returns (n, 3, 4)."""
        ch = _fq(challenges).reshape(-1, 4)
        out = np.empty((ch.shape[0], 3, 4), dtype=np.uint64)
        check(self.ctx.L.spg_sc1_run_rounds(self.h, ch.shape[0], _ptr(ch), _ptr(out)), "spg_sc1_run_rounds")
        return out

    def run_rounds_sharded(self, challenges, mailbox_addr: int, slot_stride: int, rank: int, world: int, calls: np.ndarray):
        """run_rounds for one shard: per-round exchange of the partial evaluations through the shared-memory
        mailbox at `mailbox_addr` (see parallel.ShmComm); `calls` is the shared call counter (1 uint64)."""
        ch = _fq(challenges).reshape(-1, 4)
        out = np.empty((ch.shape[0], 3, 4), dtype=np.uint64)
        check(self.ctx.L.spg_sc1_run_rounds_sharded(self.h, ch.shape[0], _ptr(ch), _ptr(out), C.c_void_p(mailbox_addr),
                                                    slot_stride, rank, world, _ptr(calls)), "spg_sc1_run_rounds_sharded")
        return out

    def set_row_weights(self, weights):
        """caller-supplied row weights for the x rounds (spg_sc1_set_row_weights): a rank of a sharded
        proof passes the global eq_p * eq_q weights of the rows it owns"""
        w = _fq(weights)
        check(self.ctx.L.spg_sc1_set_row_weights(self.h, _ptr(w), w.shape[0]), "spg_sc1_set_row_weights")

    def set_claim_checked(self, claim):
        """set_claim, and the first round checks the claim against the tables (SpgError if it is not the
        true sum, e.g. an unsatisfied witness)"""
        c = _fq(np.asarray(claim, dtype=np.uint64).reshape(4))
        check(self.ctx.L.spg_sc1_set_claim_checked(self.h, _ptr(c)), "spg_sc1_set_claim_checked")

    def set_satisfied(self):
        """set_claim(0) plus: the witness satisfies the instance row by row, so the first round (fused with
        the SpMV) evaluates one point per pair instead of two. Exact for a satisfying witness."""
        check(self.ctx.L.spg_sc1_set_satisfied(self.h), "spg_sc1_set_satisfied")

    def set_claim(self, claim):
        """The prover's `claim` argument (src/sumcheck.rs:1069; zero in R1CSProof::prove): lets round 0
        use e(1) = claim - e(0) like the reference does. Exact iff the claim is the true sum."""
        check(self.ctx.L.spg_sc1_set_claim(self.h, _ptr(_fq(claim))), "spg_sc1_set_claim")

    def round_eval(self) -> np.ndarray:
        out = np.empty((3, 4), dtype=np.uint64)
        check(self.ctx.L.spg_sc1_round_eval(self.h, _ptr(out)), "spg_sc1_round_eval")
        return out

    def round_bind(self, r):
        check(self.ctx.L.spg_sc1_round_bind(self.h, _ptr(_fq(r))), "spg_sc1_round_bind")

    def final(self) -> np.ndarray:
        out = np.empty((4, 4), dtype=np.uint64)
        check(self.ctx.L.spg_sc1_final(self.h, _ptr(out)), "spg_sc1_final")
        return out

    def debug_tables(self):
        n = C.c_size_t()
        check(self.ctx.L.spg_sc1_debug_tables(self.h, None, None, None, 0, C.byref(n)), "spg_sc1_debug_tables")
        out = [np.empty((n.value, 4), dtype=np.uint64) for _ in range(3)]
        check(self.ctx.L.spg_sc1_debug_tables(self.h, _ptr(out[0]), _ptr(out[1]), _ptr(out[2]), n.value, C.byref(n)),
              "spg_sc1_debug_tables")
        return out

    def free(self):
        if getattr(self, "h", None):
            self.ctx.L.spg_sc1_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class R1CSInstance:
    """Device copy of R1CSInstance (src/r1csinstance.rs:19-31, ::new :89-182).

    ``A_list``/``B_list``/``C_list``: per instance, (rows, cols, vals) COO arrays --
    the reference's ``Vec<(usize, usize, Scalar)>`` in struct-of-arrays form."""

    def __init__(self, ctx: Context, num_instances, max_num_cons, num_cons, num_vars, A_list, B_list, C_list):
        assert len(A_list) == len(B_list) == len(C_list) == num_instances
        self.ctx = ctx
        self.num_instances, self.max_num_cons, self.num_vars = num_instances, max_num_cons, num_vars
        self.num_cons = list(num_cons)
        rows, cols, vals, nnz = [], [], [], []
        for i in range(num_instances):
            for M in (A_list[i], B_list[i], C_list[i]):
                r, c, v = M
                rows.append(np.asarray(r, dtype=np.uint32))
                cols.append(np.asarray(c, dtype=np.uint32))
                vals.append(_fq(v).reshape(-1, 4))
                nnz.append(len(r))
        rows = np.ascontiguousarray(np.concatenate(rows)) if rows else np.zeros(0, np.uint32)
        cols = np.ascontiguousarray(np.concatenate(cols)) if cols else np.zeros(0, np.uint32)
        vals = np.ascontiguousarray(np.concatenate(vals)) if vals else np.zeros((0, 4), np.uint64)
        h = C.c_void_p()
        check(ctx.L.spg_r1cs_create(ctx.h, num_instances, max_num_cons, _ptr(_sz(num_cons)), num_vars, _ptr(_sz(nnz)),
                                    _ptr(rows), _ptr(cols), _ptr(vals), C.byref(h)), "spg_r1cs_create")
        self.h = h

    def get_num_instances(self):
        return self.num_instances

    def get_num_cons(self):
        return self.max_num_cons

    def get_inst_num_cons(self):
        return self.num_cons

    def multi_evaluate(self, rx, ry) -> np.ndarray:
        """R1CSInstance::multi_evaluate (src/r1csinstance.rs:583-595)."""
        rx = _fq(np.asarray(rx, dtype=np.uint64).reshape(-1, 4))
        ry = _fq(np.asarray(ry, dtype=np.uint64).reshape(-1, 4))
        out = np.empty((3 * self.num_instances, 4), dtype=np.uint64)
        check(self.ctx.L.spg_r1cs_multi_evaluate(self.ctx.h, self.h, _ptr(rx), rx.shape[0], _ptr(ry), ry.shape[0], _ptr(out)),
              "spg_r1cs_multi_evaluate")
        return out

    def multi_evaluate_bound_rp(self, rp, rx, ry):
        """R1CSInstance::multi_evaluate_bound_rp (src/r1csinstance.rs:597-629): the per-instance evaluations
        and the three MLEs over the instance index of the A, B and C evaluations at rp (zero padded to a
        power of two like DensePolynomial::new)."""
        ev = self.multi_evaluate(rx, ry)
        rp = _fq(np.asarray(rp, dtype=np.uint64).reshape(-1, 4))
        bound = []
        for m in range(3):
            poly = DensePolynomial.new(self.ctx, ev[m::3])
            bound.append(poly.evaluate(rp))
            poly.free()
        return ev, tuple(bound)

    def free(self):
        if getattr(self, "h", None):
            self.ctx.L.spg_r1cs_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class ProverWitnessSecInfo:
    """One witness section (src/lib.rs ProverWitnessSecInfo): ``w_mat`` flattened
    [p][q][i]; ``num_proofs[p]`` is 1 for a short section; one instance for a single one."""

    def __init__(self, ctx: Context, num_proofs, num_inputs, w_mat, asynchronous: bool = False):
        """asynchronous=True: the H2D copy runs on the context's copy stream (w_mat must be
        pinned host memory and is kept referenced until this object is freed)."""
        self.ctx = ctx
        self.num_proofs, self.num_inputs = list(num_proofs), list(num_inputs)
        w = _fq(w_mat).reshape(-1, 4)
        assert w.shape[0] == sum(a * b for a, b in zip(self.num_proofs, self.num_inputs))
        h = C.c_void_p()
        fn = ctx.L.spg_witness_upload_async if asynchronous else ctx.L.spg_witness_upload
        check(fn(ctx.h, len(self.num_proofs), _ptr(_sz(num_proofs)), _ptr(_sz(num_inputs)), _ptr(w), C.byref(h)),
              "spg_witness_upload")
        self._host = w if asynchronous else None
        self.h = h

    def poly_w(self, p: int) -> DensePolynomial:
        h = C.c_void_p()
        check(self.ctx.L.spg_witness_poly(self.h, p, C.byref(h)), "spg_witness_poly")
        d = DensePolynomial(self.ctx, h, owner=self)
        d.free = lambda: None  # view owned by the witness handle
        return d

    def free(self):
        if getattr(self, "h", None):
            self.ctx.L.spg_witness_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class ZMat:
    """z_mat of R1CSProof::prove (src/r1csproof.rs:278-293), device resident."""

    def __init__(self, ctx: Context, num_proofs, num_inputs, witness_secs: Sequence[ProverWitnessSecInfo]):
        self.ctx = ctx
        self.num_proofs, self.num_inputs = list(num_proofs), list(num_inputs)
        self.witness_secs = list(witness_secs)
        arr = (C.c_void_p * len(witness_secs))(*[w.h for w in witness_secs])
        h = C.c_void_p()
        check(ctx.L.spg_zmat_build(ctx.h, len(self.num_proofs), _ptr(_sz(num_proofs)), _ptr(_sz(num_inputs)),
                                   len(witness_secs), arr, C.byref(h)), "spg_zmat_build")
        self.h = h

    def free(self):
        if getattr(self, "h", None):
            self.ctx.L.spg_zmat_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def sumcheck_phase1(ctx: Context, inst: R1CSInstance, z: ZMat, num_proofs, max_num_proofs, num_cons, max_num_cons,
                    max_num_inputs, tau_p, tau_q, tau_x) -> SumcheckPhase1:
    """multiply_vec_block + eq tables, ready for the phase-1 rounds (src/r1csproof.rs:305-343)."""
    npf, nc = _sz(num_proofs), _sz(num_cons)
    tp, tq, tx = (_fq(np.asarray(t, dtype=np.uint64).reshape(-1, 4)) for t in (tau_p, tau_q, tau_x))
    h = C.c_void_p()
    check(ctx.L.spg_sc1_create(ctx.h, inst.h, z.h, len(npf), _ptr(npf), max_num_proofs, _ptr(nc), max_num_cons,
                               max_num_inputs, _ptr(tp), _ptr(tq), _ptr(tx), C.byref(h)), "spg_sc1_create")
    sc = SumcheckPhase1(ctx, h)
    sc._keep = (inst, z)  # multiply_vec_block is fused into the first round: both must outlive it
    return sc


def host_sum(vals) -> np.ndarray:
    """Scalar::add over axis 0 of a (count, width, 4) array, on the host side of the ABI."""
    v = _fq(vals)
    if v.ndim == 2:
        v = v.reshape(v.shape[0], 1, 4)
    out = np.empty((v.shape[1], 4), dtype=np.uint64)
    check(_lib.lib().spg_fq_host_sum(_ptr(v), v.shape[0], v.shape[1], _ptr(out)), "spg_fq_host_sum")
    return out


def host_mul(a, b) -> np.ndarray:
    out = np.empty(4, dtype=np.uint64)
    check(_lib.lib().spg_fq_host_mul(_ptr(_fq(a)), _ptr(_fq(b)), _ptr(out)), "spg_fq_host_mul")
    return out


def host_eq_weight(tau, index: int) -> np.ndarray:
    """prod_k eq(tau[k], bit_k(index)): eq weight of a shard index (low bit first)."""
    t = _fq(np.asarray(tau, dtype=np.uint64).reshape(-1, 4))
    out = np.empty(4, dtype=np.uint64)
    check(_lib.lib().spg_fq_host_eq_weight(_ptr(t), t.shape[0], C.c_uint64(index), _ptr(out)), "spg_fq_host_eq_weight")
    return out


def zmat_bind_rq(ctx: Context, z: "ZMat", rq_rev, scale=None, out: DensePolynomial | None = None) -> DensePolynomial:
    """Z_poly.bound_poly_vars_rq on its own (src/r1csproof.rs:478); natural [p][w][y] output."""
    rq = _fq(np.asarray(rq_rev, dtype=np.uint64).reshape(-1, 4))
    total = sum(len(z.witness_secs) * y for y in z.num_inputs)
    if out is None:
        out = DensePolynomial.empty(ctx, total)
    sc = None if scale is None else _fq(scale)
    check(ctx.L.spg_zmat_bind_rq(ctx.h, z.h, _ptr(rq), rq.shape[0], _ptr(sc), out.h), "spg_zmat_bind_rq")
    return out


def zmat_bind_weights(ctx: Context, z: "ZMat", weights, out: DensePolynomial | None = None, out_off=None) -> DensePolynomial:
    """sum_q weights[p][q] * Z[p][q][w][y] with explicit per-row weights (spg_zmat_bind_weights);
    out_off places instance p's table at out[out_off[p]:] (a batch-wide table)."""
    w = _fq(weights)
    total = sum(len(z.witness_secs) * y for y in z.num_inputs)
    if out is None:
        out = DensePolynomial.empty(ctx, total)
    off = None if out_off is None else _sz(out_off)
    check(ctx.L.spg_zmat_bind_weights(ctx.h, z.h, _ptr(w), w.shape[0], _ptr(off), out.h), "spg_zmat_bind_weights")
    return out


class SumcheckPhase2:
    """Device loops of prove_cubic_disjoint_rounds (src/sumcheck.rs:788-1065)."""

    @classmethod
    def from_zrq(cls, ctx: Context, inst: R1CSInstance, zrq: DensePolynomial, num_inputs, max_num_inputs,
                 num_witness_secs, rx, rp, r_A, r_B, r_C):
        self = cls.__new__(cls)
        self.ctx = ctx
        nin = _sz(num_inputs)
        a = [_fq(np.asarray(t, dtype=np.uint64).reshape(-1, 4)) for t in (rx, rp)]
        h = C.c_void_p()
        check(ctx.L.spg_sc2_create_from_zrq(ctx.h, inst.h, zrq.h, len(nin), _ptr(nin), max_num_inputs, num_witness_secs,
                                            _ptr(a[0]), _ptr(a[1]), _ptr(_fq(r_A)), _ptr(_fq(r_B)), _ptr(_fq(r_C)),
                                            C.byref(h)), "spg_sc2_create_from_zrq")
        self.h = h
        return self

    def __init__(self, ctx: Context, inst: R1CSInstance, z: ZMat, num_proofs, max_num_proofs, num_inputs, max_num_inputs,
                 num_witness_secs, rx, rq_rev, rp, r_A, r_B, r_C):
        self.ctx = ctx
        npf, nin = _sz(num_proofs), _sz(num_inputs)
        a = [_fq(np.asarray(t, dtype=np.uint64).reshape(-1, 4)) for t in (rx, rq_rev, rp)]
        h = C.c_void_p()
        check(ctx.L.spg_sc2_create(ctx.h, inst.h, z.h, len(npf), _ptr(npf), max_num_proofs, _ptr(nin), max_num_inputs,
                                   num_witness_secs, _ptr(a[0]), _ptr(a[1]), _ptr(a[2]), _ptr(_fq(r_A)), _ptr(_fq(r_B)),
                                   _ptr(_fq(r_C)), C.byref(h)), "spg_sc2_create")
        self.h = h

    @classmethod
    def slice(cls, ctx: Context, inst: R1CSInstance, zrq: DensePolynomial, max_num_inputs, num_witness_secs: int,
              flat_off: int, flat_len: int, rx, r_A, r_B, r_C) -> "SumcheckPhase2":
        """one rank's chunk [flat_off, flat_off + flat_len) of a y-sharded phase 2 (spg_sc2_create_slice)"""
        self = cls.__new__(cls)
        self.ctx = ctx
        h = C.c_void_p()
        f = lambda v: _fq(np.asarray(v, dtype=np.uint64).reshape(-1, 4))
        check(ctx.L.spg_sc2_create_slice(ctx.h, inst.h, zrq.h, max_num_inputs, num_witness_secs, flat_off, flat_len, _ptr(f(rx)),
                                         _ptr(f(r_A)), _ptr(f(r_B)), _ptr(f(r_C)), C.byref(h)), "spg_sc2_create_slice")
        self.h = h
        self._keep = (inst, zrq)
        return self

    def run_rounds_sharded(self, challenges, mailbox_addr: int, slot_stride: int, rank: int, world: int, calls: np.ndarray):
        """local rounds with the mailbox exchange in one C loop (spg_sc2_run_rounds_sharded)"""
        ch = _fq(np.asarray(challenges, dtype=np.uint64).reshape(-1, 4))
        out = np.empty((ch.shape[0], 3, 4), dtype=np.uint64)
        check(self.ctx.L.spg_sc2_run_rounds_sharded(self.h, ch.shape[0], _ptr(ch), _ptr(out), C.c_void_p(mailbox_addr), slot_stride,
                                                    rank, world, _ptr(calls)), "spg_sc2_run_rounds_sharded")
        return out

    @property
    def num_rounds(self) -> int:
        return int(self.ctx.L.spg_sc2_num_rounds(self.h))

    def run_rounds(self, challenges) -> np.ndarray:
        ch = _fq(challenges).reshape(-1, 4)
        out = np.empty((ch.shape[0], 3, 4), dtype=np.uint64)
        check(self.ctx.L.spg_sc2_run_rounds(self.h, ch.shape[0], _ptr(ch), _ptr(out)), "spg_sc2_run_rounds")
        return out

    def round_eval(self) -> np.ndarray:
        out = np.empty((3, 4), dtype=np.uint64)
        check(self.ctx.L.spg_sc2_round_eval(self.h, _ptr(out)), "spg_sc2_round_eval")
        return out

    def round_bind(self, r):
        check(self.ctx.L.spg_sc2_round_bind(self.h, _ptr(_fq(r))), "spg_sc2_round_bind")

    def final(self) -> np.ndarray:
        out = np.empty((3, 4), dtype=np.uint64)
        check(self.ctx.L.spg_sc2_final(self.h, _ptr(out)), "spg_sc2_final")
        return out

    def free(self):
        if getattr(self, "h", None):
            self.ctx.L.spg_sc2_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class ProductCircuit:
    """ProductCircuit (src/product_tree.rs:11-64), all layers device resident."""

    def __init__(self, ctx: Context, poly: DensePolynomial):
        self.ctx = ctx
        h = C.c_void_p()
        check(ctx.L.spg_prodtree_build(ctx.h, poly.h, C.byref(h)), "spg_prodtree_build")
        self.h = h

    @property
    def num_layers(self) -> int:
        return int(self.ctx.L.spg_prodtree_num_layers(self.h))

    def layer(self, k: int):
        """(left_vec[k], right_vec[k]) as views."""
        l, r = C.c_void_p(), C.c_void_p()
        check(self.ctx.L.spg_prodtree_layer(self.h, k, C.byref(l), C.byref(r)), "spg_prodtree_layer")
        out = []
        for hh in (l, r):
            d = DensePolynomial(self.ctx, hh, owner=self)
            d.free = lambda: None
            out.append(d)
        return out

    def evaluate(self) -> np.ndarray:
        out = np.empty(4, dtype=np.uint64)
        check(self.ctx.L.spg_prodtree_evaluate(self.ctx.h, self.h, _ptr(out)), "spg_prodtree_evaluate")
        return out

    def free(self):
        if getattr(self, "h", None):
            self.ctx.L.spg_prodtree_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class CubicBatched:
    """Device loops of SumcheckInstanceProof::prove_cubic_batched (src/sumcheck.rs:264-434)."""

    def __init__(self, ctx: Context, A_par, B_par, C_par, A_seq, B_seq, C_seq, coeffs):
        self.ctx = ctx
        self._keep = (A_par, B_par, C_par, A_seq, B_seq, C_seq)
        arr = lambda vs: (C.c_void_p * max(len(vs), 1))(*[v.h for v in vs])
        h = C.c_void_p()
        co = _fq(coeffs).reshape(-1, 4)
        check(ctx.L.spg_cubic_create(ctx.h, len(A_par), arr(A_par), arr(B_par), C_par.h if C_par is not None else None,
                                     len(A_seq), arr(A_seq), arr(B_seq), arr(C_seq), _ptr(co), C.byref(h)), "spg_cubic_create")
        self.h = h
        self.n_claims = 2 * len(A_par) + (1 if A_par else 0) + 3 * len(A_seq)

    def round_eval(self) -> np.ndarray:
        out = np.empty((3, 4), dtype=np.uint64)
        check(self.ctx.L.spg_cubic_round_eval(self.h, _ptr(out)), "spg_cubic_round_eval")
        return out

    def round_bind(self, r):
        check(self.ctx.L.spg_cubic_round_bind(self.h, _ptr(_fq(r))), "spg_cubic_round_bind")

    def final(self) -> np.ndarray:
        out = np.empty((self.n_claims, 4), dtype=np.uint64)
        check(self.ctx.L.spg_cubic_final(self.h, _ptr(out)), "spg_cubic_final")
        return out

    def free(self):
        if getattr(self, "h", None):
            self.ctx.L.spg_cubic_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def hash_layer(ctx: Context, addr, val: DensePolynomial, ts, gamma, tau, ts_plus_one=False) -> DensePolynomial:
    """One vector of Layers::build_hash_layer (src/sparse_mlpoly.rs:612-687):
    ts*gamma^2 + val*gamma + addr - tau; addr=None means the cell index, ts=None means 0."""
    n = len(val)
    a = None if addr is None else np.ascontiguousarray(addr, dtype=np.uint64)
    t = None if ts is None else np.ascontiguousarray(ts, dtype=np.uint64)
    h = C.c_void_p()
    check(ctx.L.spg_hash_layer(ctx.h, _ptr(a), val.h, _ptr(t), n, _ptr(_fq(gamma)), _ptr(_fq(tau)), int(ts_plus_one),
                               C.byref(h)), "spg_hash_layer")
    return DensePolynomial(ctx, h)


def deref(ctx: Context, addr, mem: DensePolynomial) -> DensePolynomial:
    """AddrTimestamps::deref_mem (src/sparse_mlpoly.rs:255-264)."""
    a = np.ascontiguousarray(addr, dtype=np.uint64)
    h = C.c_void_p()
    check(ctx.L.spg_deref(ctx.h, _ptr(a), len(a), mem.h, C.byref(h)), "spg_deref")
    return DensePolynomial(ctx, h)


class MultiSparseMatPolynomialAsDense:
    """SparseMatPolynomial::multi_sparse_to_dense_rep (src/sparse_mlpoly.rs:368-425) on the device:
    padded address vectors, read / audit timestamps (AddrTimestamps::new, :222-253) and values of a
    batch of sparse matrices `polys` = [(rows, cols, vals), ...]."""

    KINDS = {"row_addr": 0, "row_read_ts": 1, "col_addr": 2, "col_read_ts": 3, "val": 4, "row_audit_ts": 5,
             "col_audit_ts": 6, "comb_ops": 7, "comb_mem": 8}

    def __init__(self, ctx: Context, polys, num_vars_x: int, num_vars_y: int):
        self.ctx = ctx
        nnz = _sz([len(p[0]) for p in polys])
        rows = np.ascontiguousarray(np.concatenate([np.asarray(p[0], dtype=np.uint32) for p in polys]))
        cols = np.ascontiguousarray(np.concatenate([np.asarray(p[1], dtype=np.uint32) for p in polys]))
        vals = _fq(np.concatenate([np.asarray(p[2], dtype=np.uint64).reshape(-1, 4) for p in polys]))
        h = C.c_void_p()
        check(ctx.L.spg_sparse_create(ctx.h, len(polys), num_vars_x, num_vars_y, _ptr(nnz), _ptr(rows), _ptr(cols), _ptr(vals),
                                      C.byref(h)), "spg_sparse_create")
        self.h = h
        self.batch = len(polys)
        self.num_ops = int(ctx.L.spg_sparse_num_ops(h))
        self.num_mem_cells = int(ctx.L.spg_sparse_num_mem_cells(h))

    def view(self, kind: str, i: int = 0) -> DensePolynomial:
        v = C.c_void_p()
        check(self.ctx.L.spg_sparse_view(self.h, self.KINDS[kind], i, C.byref(v)), "spg_sparse_view")
        return DensePolynomial(self.ctx, v, owner=self)

    def free(self):
        if getattr(self, "h", None):
            self.ctx.L.spg_sparse_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class BulletReduction:
    """Device side of BulletReductionProof::prove (src/nizk/bullet.rs:72-119) with the generator
    fold unrolled onto the scalars (spg_bullet_*): ``lr`` returns the group parts of a round's L
    and R over the original generators, ``fold`` applies the round's challenge, ``final`` returns
    the folded generator G_hat. The caller keeps a, b, c_L * Q / c_R * Q and the transcript."""

    def __init__(self, ctx: Context, gens: "MultiCommitGens", n: int):
        self.ctx, self.n, self.gens = ctx, n, gens
        h = C.c_void_p()
        check(ctx.L.spg_bullet_create(ctx.h, gens.h, n, C.byref(h)), "spg_bullet_create")
        self.h = h

    def lr(self, a, blind_L, blind_R):
        a = _fq(a)
        bl = np.stack([_fq(blind_L).reshape(4), _fq(blind_R).reshape(4)])
        out = np.empty(64, dtype=np.uint8)
        check(self.ctx.L.spg_bullet_lr(self.h, a.shape[0], _ptr(a), _ptr(bl), _ptr(out)), "spg_bullet_lr")
        return out[:32].tobytes(), out[32:].tobytes()

    def fold(self, nk: int, u, u_inv):
        u, ui = _fq(u).reshape(4), _fq(u_inv).reshape(4)
        check(self.ctx.L.spg_bullet_fold(self.h, nk, _ptr(u), _ptr(ui)), "spg_bullet_fold")

    def final(self) -> bytes:
        out = np.empty(32, dtype=np.uint8)
        check(self.ctx.L.spg_bullet_final(self.h, _ptr(out)), "spg_bullet_final")
        return out.tobytes()

    # a and b resident on the device: the host loops over them (bullet.rs:83-84, 113-116) move along
    def set_ab(self, a, b):
        a, b = _fq(a), _fq(b)
        assert a.shape[0] == self.n and b.shape[0] == self.n
        check(self.ctx.L.spg_bullet_set_ab(self.h, _ptr(a), _ptr(b)), "spg_bullet_set_ab")

    def lr_resident(self, nk: int, blind_L, blind_R, ext: bool = False):
        """(L, R, c_L, c_R) of the round over the device's a and b; L and R as ristretto encodings, or with
        ``ext`` as 128 bytes of extended coordinates X, Y, Z, T each"""
        bl = np.stack([_fq(blind_L).reshape(4), _fq(blind_R).reshape(4)])
        per = 128 if ext else 32
        out = np.empty(2 * per, dtype=np.uint8)
        c = np.empty((2, 4), dtype=np.uint64)
        check(self.ctx.L.spg_bullet_lr_resident(self.h, nk, _ptr(bl), int(ext), _ptr(out), _ptr(c)), "spg_bullet_lr_resident")
        return out[:per].tobytes(), out[per:].tobytes(), c[0], c[1]

    def final_ab(self):
        """(G_hat, a[0], b[0]) after the last fold"""
        out = np.empty(32, dtype=np.uint8)
        ab = np.empty((2, 4), dtype=np.uint64)
        check(self.ctx.L.spg_bullet_final_ab(self.h, _ptr(out), _ptr(ab)), "spg_bullet_final_ab")
        return out.tobytes(), ab[0], ab[1]

    def free(self):
        if getattr(self, "h", None) is not None and self.h.value:
            self.ctx.L.spg_bullet_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class MultiCommitGens:
    """Device copy of MultiCommitGens (src/commitments.rs:8-67). Generator derivation
    (SHAKE256 -> from_uniform_bytes) is one-off host setup; the caller passes the n + 1
    compressed ristretto points G[0..n], h."""

    def __init__(self, ctx: Context, compressed: bytes):
        assert len(compressed) % 32 == 0 and len(compressed) >= 64
        self.ctx = ctx
        self.n = len(compressed) // 32 - 1
        buf = np.frombuffer(bytes(compressed), dtype=np.uint8).copy()
        h = C.c_void_p()
        check(ctx.L.spg_gens_upload(ctx.h, _ptr(buf), self.n + 1, C.byref(h)), "spg_gens_upload")
        self.h = h

    @classmethod
    def from_uniform(cls, ctx: Context, uniform: bytes):
        """MultiCommitGens::new on the device from its SHAKE256 output (64 bytes per point, h last)."""
        assert len(uniform) % 64 == 0 and len(uniform) >= 128
        self = cls.__new__(cls)
        self.ctx = ctx
        self.n = len(uniform) // 64 - 1
        buf = np.frombuffer(bytes(uniform), dtype=np.uint8).copy()
        h = C.c_void_p()
        check(ctx.L.spg_gens_from_uniform(ctx.h, _ptr(buf), self.n + 1, C.byref(h)), "spg_gens_from_uniform")
        self.h = h
        return self

    def commit_poly(self, poly: DensePolynomial, L_size: int | None = None) -> list:
        """DensePolynomial::commit with zero blinds (src/dense_mlpoly.rs:214-239): L_size
        compressed row commitments."""
        if L_size is None:
            L_size = 1 << (poly.get_num_vars() // 2)
        out = np.empty(32 * L_size, dtype=np.uint8)
        check(self.ctx.L.spg_poly_commit(self.ctx.h, self.h, poly.h, L_size, _ptr(out)), "spg_poly_commit")
        return [out[32 * i: 32 * (i + 1)].tobytes() for i in range(L_size)]

    def commit_poly_rows(self, poly: DensePolynomial, L_size: int, row0: int, nrows: int) -> bytes:
        """Rows [row0, row0 + nrows) of DensePolynomial::commit (the rows are independent:
        src/dense_mlpoly.rs:199-212); nrows * 32 bytes."""
        out = np.empty(32 * max(nrows, 1), dtype=np.uint8)
        check(self.ctx.L.spg_poly_commit_rows(self.ctx.h, self.h, poly.h, L_size, row0, nrows, _ptr(out)), "spg_poly_commit_rows")
        return out[:32 * nrows].tobytes()

    def prepare(self, R: int, rows: int | None = None):
        """Build the fixed-base tables for the first R generators now (setup, not prove time);
        with `rows`, also the single-window table a commitment of that many rows would use."""
        if rows is None:
            check(self.ctx.L.spg_gens_prepare(self.ctx.h, self.h, R), "spg_gens_prepare")
        else:
            check(self.ctx.L.spg_gens_prepare_rows(self.ctx.h, self.h, rows, R), "spg_gens_prepare_rows")

    def info(self) -> dict:
        o = np.zeros(4, dtype=np.uint64)
        check(self.ctx.L.spg_gens_info(self.h, _ptr(o)), "spg_gens_info")
        d = {"window_bits": int(o[0]), "adds_per_scalar": int(o[1]), "table_bytes": int(o[2]), "table_bases": int(o[3])}
        check(self.ctx.L.spg_gens_info_rows(self.h, _ptr(o)), "spg_gens_info_rows")
        d["rows_table"] = {"window_bits": int(o[0]), "adds_per_scalar": int(o[1]), "table_bytes": int(o[2]), "table_bases": int(o[3])}
        return d

    def commit_batch(self, scalars, blinds=None) -> list:
        """Commitments::commit for `count` vectors of equal length sharing these generators."""
        s = _fq(scalars)
        assert s.ndim == 3
        count, length = s.shape[0], s.shape[1]
        b = None if blinds is None else _fq(blinds).reshape(count, 4)
        out = np.empty(32 * count, dtype=np.uint8)
        check(self.ctx.L.spg_commit_batch(self.ctx.h, self.h, _ptr(s), length, _ptr(b), count, _ptr(out)), "spg_commit_batch")
        return [out[32 * i: 32 * (i + 1)].tobytes() for i in range(count)]

    def free(self):
        if getattr(self, "h", None):
            self.ctx.L.spg_gens_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass
