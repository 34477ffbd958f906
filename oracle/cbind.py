"""ORACLE (test infrastructure, NOT product code): ctypes binding over liboracle.so.

Scalars are numpy ``uint64`` arrays whose last axis has length 4: the reference's
Montgomery limbs (``/root/reference/src/scalar/ristretto255.rs:193-199``).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")

MODE_P, MODE_Q, MODE_W, MODE_X = 1, 2, 3, 4


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in ("fq.c", "polys.c", "sumcheck.c", "witness.c", "fq.h", "polys.h", "sumcheck.h", "witness.h")]
    stale = not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs)
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-s"], env={**os.environ, "CC": "gcc"})
    return _SO


class Fq(C.Structure):
    _fields_ = [("l", C.c_uint64 * 4)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _proto(_lib)
    return _lib


_P = C.c_void_p
_SZ = C.c_size_t


def _proto(L):
    for name in ("ofq_add", "ofq_sub", "ofq_mul"):
        f = getattr(L, name)
        f.restype = Fq
        f.argtypes = [_P, _P]
    for name in ("ofq_neg", "ofq_square", "ofq_invert"):
        f = getattr(L, name)
        f.restype = Fq
        f.argtypes = [_P]
    L.ofq_from_u64.restype = Fq
    L.ofq_from_u64.argtypes = [C.c_uint64]
    L.ofq_from_raw.restype = Fq
    L.ofq_from_raw.argtypes = [_P]
    L.ofq_from_u512.restype = Fq
    L.ofq_from_u512.argtypes = [_P]
    L.ofq_from_bytes.restype = C.c_int
    L.ofq_from_bytes.argtypes = [_P, _P]
    L.ofq_to_bytes.restype = None
    L.ofq_to_bytes.argtypes = [_P, _P]
    L.ofq_from_bytes_wide.restype = Fq
    L.ofq_from_bytes_wide.argtypes = [_P]
    L.ofq_pow.restype = Fq
    L.ofq_pow.argtypes = [_P, _P]
    L.ofq_batch_invert.restype = Fq
    L.ofq_batch_invert.argtypes = [_P, _SZ]
    L.ofq_montgomery_reduce.restype = Fq
    L.ofq_montgomery_reduce.argtypes = [_P]
    for name in ("ofq_vec_mul", "ofq_vec_add", "ofq_vec_sub"):
        f = getattr(L, name)
        f.restype = None
        f.argtypes = [_P, _P, _P, _SZ]
    L.ofq_vec_from_u512.restype = None
    L.ofq_vec_from_u512.argtypes = [_P, _P, _SZ]

    L.oeq_evals.restype = None
    L.oeq_evals.argtypes = [_P, _SZ, _P]
    L.oeq_evaluate.restype = Fq
    L.oeq_evaluate.argtypes = [_P, _P, _SZ]
    L.odense_bound_top.restype = _SZ
    L.odense_bound_top.argtypes = [_P, _SZ, _P]
    L.odense_bound_bot.restype = _SZ
    L.odense_bound_bot.argtypes = [_P, _SZ, _P]
    L.odense_evaluate.restype = Fq
    L.odense_evaluate.argtypes = [_P, _SZ, _P, _SZ]
    L.odense_bound_L.restype = None
    L.odense_bound_L.argtypes = [_P, _SZ, _P, _P]
    for name in ("owit_perm_w0", "owit_exec", "owit_block", "owit_mem", "owit_shift"):
        getattr(L, name).restype = None
    L.owit_perm_w0.argtypes = [_P, _P, _SZ, _SZ, _P]
    L.owit_exec.argtypes = [_P, _SZ, _SZ, _P, _P, _SZ, _SZ, _P, _P]
    L.owit_block.argtypes = [_P, _SZ, _SZ, _P, _P, _P, _SZ, _SZ, _SZ, _SZ, _SZ, _P, _P]
    L.owit_mem.argtypes = [_P, _SZ, _SZ, _P, _P, _SZ, _P, _P]
    L.owit_shift.argtypes = [_P, _SZ, _SZ, _P]
    L.owit_perm_fill.restype = None
    L.owit_perm_fill.argtypes = [_P, _SZ, _P, _SZ, _SZ, _SZ, _SZ, _SZ]
    L.odot.restype = Fq
    L.odot.argtypes = [_P, _P, _SZ]
    L.ounipoly_from_evals.restype = None
    L.ounipoly_from_evals.argtypes = [_P, _SZ, _P]
    L.ounipoly_evaluate.restype = Fq
    L.ounipoly_evaluate.argtypes = [_P, _SZ, _P]
    L.orev_bits.restype = _SZ
    L.orev_bits.argtypes = [_SZ, _SZ]

    for name in ("opqx_new_rev", "opqx_new"):
        f = getattr(L, name)
        f.restype = _P
        f.argtypes = [_P, _SZ, _SZ, _P, _SZ, _P, _SZ]
    L.opqx_clone.restype = _P
    L.opqx_clone.argtypes = [_P]
    L.opqx_free.restype = None
    L.opqx_free.argtypes = [_P]
    L.opqx_total.restype = _SZ
    L.opqx_total.argtypes = [_P]
    L.opqx_copy_out.restype = None
    L.opqx_copy_out.argtypes = [_P, _P]
    L.opqx_index.restype = Fq
    L.opqx_index.argtypes = [_P, _SZ, _SZ, _SZ, _SZ]
    L.opqx_index_high.restype = Fq
    L.opqx_index_high.argtypes = [_P, _SZ, _SZ, _SZ, _SZ, C.c_int]
    L.opqx_bound_poly.restype = None
    L.opqx_bound_poly.argtypes = [_P, _P, C.c_int]
    L.opqx_len.restype = _SZ
    L.opqx_len.argtypes = [_P]
    L.opqx_evaluate.restype = Fq
    L.opqx_evaluate.argtypes = [_P, _P, _SZ, _P, _SZ, _P, _SZ, _P, _SZ]

    L.osc1_new.restype = _P
    L.osc1_new.argtypes = [_SZ, _SZ, _SZ, _SZ, _P, _P, _P, _P, _P, _P, _P, _P]
    L.osc1_free.restype = None
    L.osc1_free.argtypes = [_P]
    L.osc1_round_eval.restype = None
    L.osc1_round_eval.argtypes = [_P, _P]
    L.osc1_round_bind.restype = None
    L.osc1_round_bind.argtypes = [_P, _P]
    L.osc1_final.restype = None
    L.osc1_final.argtypes = [_P, _P]

    L.osc2_new.restype = _P
    L.osc2_new.argtypes = [_SZ, _SZ, _SZ, C.c_int, _SZ, _SZ, _P, _P, _P, _P]
    L.osc2_free.restype = None
    L.osc2_free.argtypes = [_P]
    L.osc2_round_eval.restype = None
    L.osc2_round_eval.argtypes = [_P, _P]
    L.osc2_round_bind.restype = None
    L.osc2_round_bind.argtypes = [_P, _P]
    L.osc2_final.restype = None
    L.osc2_final.argtypes = [_P, _P]

    L.ocubic_batched_eval.restype = None
    L.ocubic_batched_eval.argtypes = [_SZ, _SZ, _P, _P, _P, _SZ, _P, _P, _P, _P, _P]
    L.ospmv.restype = None
    L.ospmv.argtypes = [_SZ, _P, _P, _P, _SZ, _SZ, _P, _SZ, _P]
    L.ospmv_batch.restype = None
    L.ospmv_batch.argtypes = [_SZ, _P, _P, _P, _SZ, _SZ, _P, _SZ, _SZ, _SZ, _P]
    L.oeval_table_sparse.restype = None
    L.oeval_table_sparse.argtypes = [_SZ, _P, _P, _P, _P, _SZ, _SZ, _SZ, _P]
    L.osparse_evaluate_with_tables.restype = Fq
    L.osparse_evaluate_with_tables.argtypes = [_SZ, _P, _P, _P, _P, _P]
    L.oprod_layer.restype = None
    L.oprod_layer.argtypes = [_P, _P, _SZ, _P, _P]


# ---------------------------------------------------------------------------- numpy helpers

def fq_array(a) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint64)
    assert a.shape[-1] == 4, a.shape
    return a


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _ret(f: Fq) -> np.ndarray:
    return np.array(list(f.l), dtype=np.uint64)


def _sz_arr(v) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(v, dtype=np.uint64))


ZERO = np.zeros(4, dtype=np.uint64)
ONE = np.array([0xd6ec31748d98951d, 0xc6ef5bf4737dcf70, 0xfffffffffffffffe, 0x0fffffffffffffff], dtype=np.uint64)


def add(a, b):
    return _ret(lib().ofq_add(_ptr(fq_array(a)), _ptr(fq_array(b))))


def sub(a, b):
    return _ret(lib().ofq_sub(_ptr(fq_array(a)), _ptr(fq_array(b))))


def mul(a, b):
    return _ret(lib().ofq_mul(_ptr(fq_array(a)), _ptr(fq_array(b))))


def neg(a):
    return _ret(lib().ofq_neg(_ptr(fq_array(a))))


def square(a):
    return _ret(lib().ofq_square(_ptr(fq_array(a))))


def invert(a):
    return _ret(lib().ofq_invert(_ptr(fq_array(a))))


def from_u64(v: int):
    return _ret(lib().ofq_from_u64(C.c_uint64(v)))


def from_raw(limbs):
    return _ret(lib().ofq_from_raw(_ptr(_sz_arr(limbs))))


def from_u512(limbs):
    return _ret(lib().ofq_from_u512(_ptr(_sz_arr(limbs))))


def from_bytes(b: bytes):
    """Returns (scalar, is_canonical) like CtOption (ristretto255.rs:391-415)."""
    out = np.zeros(4, dtype=np.uint64)
    buf = (C.c_uint8 * 32).from_buffer_copy(bytes(b))
    ok = lib().ofq_from_bytes(buf, _ptr(out))
    return out, bool(ok)


def to_bytes(a) -> bytes:
    buf = (C.c_uint8 * 32)()
    lib().ofq_to_bytes(_ptr(fq_array(a)), buf)
    return bytes(buf)


def from_bytes_wide(b: bytes):
    buf = (C.c_uint8 * 64).from_buffer_copy(bytes(b))
    return _ret(lib().ofq_from_bytes_wide(buf))


def pow_(a, by_limbs):
    return _ret(lib().ofq_pow(_ptr(fq_array(a)), _ptr(_sz_arr(by_limbs))))


def batch_invert(arr):
    arr = fq_array(arr).copy()
    ret = _ret(lib().ofq_batch_invert(_ptr(arr), arr.shape[0]))
    return arr, ret


def montgomery_reduce(r8):
    return _ret(lib().ofq_montgomery_reduce(_ptr(_sz_arr(r8))))


def to_int(a) -> int:
    """Canonical integer value (leaves Montgomery form)."""
    return int.from_bytes(to_bytes(a), "little")


def from_int(v: int):
    q = (1 << 252) + 27742317777372353535851937790883648493
    v %= q
    return from_raw([(v >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)])


def vec_mul(a, b):
    a, b = fq_array(a), fq_array(b)
    out = np.empty_like(a)
    lib().ofq_vec_mul(_ptr(a), _ptr(b), _ptr(out), a.size // 4)
    return out


def vec_add(a, b):
    a, b = fq_array(a), fq_array(b)
    out = np.empty_like(a)
    lib().ofq_vec_add(_ptr(a), _ptr(b), _ptr(out), a.size // 4)
    return out


def vec_sub(a, b):
    a, b = fq_array(a), fq_array(b)
    out = np.empty_like(a)
    lib().ofq_vec_sub(_ptr(a), _ptr(b), _ptr(out), a.size // 4)
    return out


def vec_from_u512(wide):
    wide = np.ascontiguousarray(wide, dtype=np.uint64)
    assert wide.shape[-1] == 8
    out = np.empty(wide.shape[:-1] + (4,), dtype=np.uint64)
    lib().ofq_vec_from_u512(_ptr(wide), _ptr(out), wide.size // 8)
    return out


# ---------------------------------------------------------------------------- tables

def eq_evals(r):
    r = fq_array(r).reshape(-1, 4)
    out = np.empty((1 << r.shape[0], 4), dtype=np.uint64)
    lib().oeq_evals(_ptr(r), r.shape[0], _ptr(out))
    return out


def eq_evaluate(r, rx):
    r, rx = fq_array(r).reshape(-1, 4), fq_array(rx).reshape(-1, 4)
    return _ret(lib().oeq_evaluate(_ptr(r), _ptr(rx), r.shape[0]))


def dense_bound_top(Z, r):
    Z = fq_array(Z).copy()
    n = lib().odense_bound_top(_ptr(Z), Z.shape[0], _ptr(fq_array(r)))
    return Z[:n].copy()


def dense_bound_bot(Z, r):
    Z = fq_array(Z).copy()
    n = lib().odense_bound_bot(_ptr(Z), Z.shape[0], _ptr(fq_array(r)))
    return Z[:n].copy()


def dense_evaluate(Z, r):
    Z, r = fq_array(Z), fq_array(r).reshape(-1, 4)
    return _ret(lib().odense_evaluate(_ptr(Z), Z.shape[0], _ptr(r), r.shape[0]))


def dense_bound_L(Z, L):
    Z, L = fq_array(Z), fq_array(L)
    ell = int(Z.shape[0]).bit_length() - 1
    out = np.empty((1 << (ell - ell // 2), 4), dtype=np.uint64)
    lib().odense_bound_L(_ptr(Z), ell, _ptr(L), _ptr(out))
    return out


def perm_fill(w3, seg_len, width=8, v_col=0, x_col=1, pi_col=2, d_col=3):
    """(pi, D) columns of a row-major w3 table, src/lib.rs:1378-1400 (sequential, last proof first)."""
    w3 = np.ascontiguousarray(fq_array(w3).copy())
    seg = np.ascontiguousarray(seg_len, dtype=np.uint64)
    assert w3.shape[0] == int(seg.sum()) * width
    lib().owit_perm_fill(_ptr(w3), width, _ptr(seg), seg.size, v_col, x_col, pi_col, d_col)
    return w3


def wit_perm_w0(tau, r, used: int, total: int):
    """perm_w0 (src/lib.rs:1328-1338)"""
    out = np.zeros((total, 4), dtype=np.uint64)
    lib().owit_perm_w0(_ptr(fq_array(tau)), _ptr(fq_array(r)), used, total, _ptr(out))
    return out


def wit_exec(inputs, w0, tau, n: int, num_ios: int):
    """perm_exec_w2, perm_exec_w3 (src/lib.rs:1346-1400); inputs: (rows, in_width, 4)"""
    a = fq_array(inputs)
    rows, in_width = a.shape[0], a.shape[1]
    w2, w3 = np.zeros((rows, num_ios, 4), dtype=np.uint64), np.zeros((rows, 8, 4), dtype=np.uint64)
    lib().owit_exec(_ptr(a), rows, in_width, _ptr(fq_array(w0)), _ptr(fq_array(tau)), n, num_ios, _ptr(w2), _ptr(w3))
    return w2, w3


def wit_block(vars_, w0, tau, r, n: int, io_width: int, phy_ops: int, vir_ops: int, w2_width: int):
    """block_w2, block_w3 of one instance (src/lib.rs:1511-1613); vars_: (rows, vars_width, 4)"""
    a = fq_array(vars_)
    rows, width = a.shape[0], a.shape[1]
    w2, w3 = np.zeros((rows, w2_width, 4), dtype=np.uint64), np.zeros((rows, 8, 4), dtype=np.uint64)
    lib().owit_block(_ptr(a), rows, width, _ptr(fq_array(w0)), _ptr(fq_array(tau)), _ptr(fq_array(r)), n, io_width, phy_ops,
                     vir_ops, w2_width, _ptr(w2), _ptr(w3))
    return w2, w3


def wit_mem(mems, tau, r, mem_width: int):
    """mem_gen (src/lib.rs:832-880); mems: (rows, in_width, 4)"""
    a = fq_array(mems)
    rows, width = a.shape[0], a.shape[1]
    w2, w3 = np.zeros((rows, mem_width, 4), dtype=np.uint64), np.zeros((rows, 8, 4), dtype=np.uint64)
    lib().owit_mem(_ptr(a), rows, width, _ptr(fq_array(tau)), _ptr(fq_array(r)), mem_width, _ptr(w2), _ptr(w3))
    return w2, w3


def wit_shift(w3):
    a = fq_array(w3)
    out = np.zeros_like(a)
    lib().owit_shift(_ptr(a), a.shape[0], a.shape[1], _ptr(out))
    return out


def dot(a, b):
    a, b = fq_array(a), fq_array(b)
    return _ret(lib().odot(_ptr(a), _ptr(b), a.shape[0]))


def unipoly_from_evals(evals):
    e = fq_array(evals)
    out = np.empty_like(e)
    lib().ounipoly_from_evals(_ptr(e), e.shape[0], _ptr(out))
    return out


def unipoly_evaluate(coeffs, r):
    c = fq_array(coeffs)
    return _ret(lib().ounipoly_evaluate(_ptr(c), c.shape[0], _ptr(fq_array(r))))


def rev_bits(q: int, n: int) -> int:
    return lib().orev_bits(q, n)


class Pqx:
    """DensePolynomialPqx (custom_dense_mlpoly.rs). ``z`` is the ragged natural
    [p][q][w][x] table flattened (instance-major), shape (sum_p Q_p*W*X_p, 4)."""

    def __init__(self, handle):
        self.h = handle
        self.owned = True

    @classmethod
    def new_rev(cls, z, W, num_proofs, max_num_proofs, num_inputs, max_num_inputs):
        z = fq_array(z)
        npf, nin = _sz_arr(num_proofs), _sz_arr(num_inputs)
        assert z.shape[0] == int(sum(int(a) * W * int(b) for a, b in zip(npf, nin)))
        return cls(lib().opqx_new_rev(_ptr(z), len(npf), W, _ptr(npf), max_num_proofs, _ptr(nin), max_num_inputs))

    @classmethod
    def new(cls, z, W, num_proofs, max_num_proofs, num_inputs, max_num_inputs):
        z = fq_array(z)
        npf, nin = _sz_arr(num_proofs), _sz_arr(num_inputs)
        return cls(lib().opqx_new(_ptr(z), len(npf), W, _ptr(npf), max_num_proofs, _ptr(nin), max_num_inputs))

    def clone(self):
        return Pqx(lib().opqx_clone(self.h))

    def release(self):
        """Hand ownership to a C prover object."""
        self.owned = False
        return self.h

    def __del__(self):
        if getattr(self, "owned", False) and self.h:
            lib().opqx_free(self.h)
            self.h = None

    def index(self, p, q, w, x):
        return _ret(lib().opqx_index(self.h, p, q, w, x))

    def index_high(self, p, q, w, x, mode):
        return _ret(lib().opqx_index_high(self.h, p, q, w, x, mode))

    def bound_poly(self, r, mode):
        lib().opqx_bound_poly(self.h, _ptr(fq_array(r)), mode)

    def len(self):
        return lib().opqx_len(self.h)

    def raw(self):
        out = np.empty((lib().opqx_total(self.h), 4), dtype=np.uint64)
        lib().opqx_copy_out(self.h, _ptr(out))
        return out

    def evaluate(self, rp, rq, rw, rx):
        a = [fq_array(np.asarray(v, dtype=np.uint64).reshape(-1, 4)) for v in (rp, rq, rw, rx)]
        return _ret(lib().opqx_evaluate(self.h, _ptr(a[0]), a[0].shape[0], _ptr(a[1]), a[1].shape[0],
                                        _ptr(a[2]), a[2].shape[0], _ptr(a[3]), a[3].shape[0]))


class Sc1:
    """State machine over prove_cubic_with_additive_term_disjoint_rounds' loops."""

    def __init__(self, nx, nq, np_, num_proofs, num_cons, Ap, Aq, Ax, B: Pqx, C_: Pqx, D: Pqx):
        npf, nc = _sz_arr(num_proofs), _sz_arr(num_cons)
        self.h = lib().osc1_new(nx, nq, np_, len(npf), _ptr(npf), _ptr(nc), _ptr(fq_array(Ap)), _ptr(fq_array(Aq)),
                                _ptr(fq_array(Ax)), B.release(), C_.release(), D.release())
        self.num_rounds = nx + nq + np_

    def round_eval(self):
        out = np.empty((3, 4), dtype=np.uint64)
        lib().osc1_round_eval(self.h, _ptr(out))
        return out

    def round_bind(self, r):
        lib().osc1_round_bind(self.h, _ptr(fq_array(r)))

    def final(self):
        out = np.empty((4, 4), dtype=np.uint64)
        lib().osc1_final(self.h, _ptr(out))
        return out

    def __del__(self):
        if self.h:
            lib().osc1_free(self.h)
            self.h = None


class Sc2:
    """State machine over prove_cubic_disjoint_rounds' loops."""

    def __init__(self, ny, nw, np_, single_inst, num_witness_secs, num_inputs, A, B: Pqx, C_: Pqx):
        nin = _sz_arr(num_inputs)
        self.h = lib().osc2_new(ny, nw, np_, int(single_inst), num_witness_secs, len(nin), _ptr(nin),
                                _ptr(fq_array(A)), B.release(), C_.release())
        self.num_rounds = ny + nw + np_

    def round_eval(self):
        out = np.empty((3, 4), dtype=np.uint64)
        lib().osc2_round_eval(self.h, _ptr(out))
        return out

    def round_bind(self, r):
        lib().osc2_round_bind(self.h, _ptr(fq_array(r)))

    def final(self):
        out = np.empty((3, 4), dtype=np.uint64)
        lib().osc2_final(self.h, _ptr(out))
        return out

    def __del__(self):
        if self.h:
            lib().osc2_free(self.h)
            self.h = None


def _ptr_array(arrs):
    arr_t = C.c_void_p * max(len(arrs), 1)
    return arr_t(*[a.ctypes.data for a in arrs])


def cubic_batched_eval(A_par, B_par, C_par, A_seq, B_seq, C_seq, coeffs):
    """One round of prove_cubic_batched (sumcheck.rs:297-371); tables are lists of (len,4) arrays."""
    A_par = [fq_array(a) for a in A_par]
    B_par = [fq_array(a) for a in B_par]
    A_seq = [fq_array(a) for a in A_seq]
    B_seq = [fq_array(a) for a in B_seq]
    C_seq = [fq_array(a) for a in C_seq]
    length = (A_par[0] if A_par else A_seq[0]).shape[0]
    Cp = fq_array(C_par) if C_par is not None else np.zeros((length, 4), dtype=np.uint64)
    out = np.empty((3, 4), dtype=np.uint64)
    lib().ocubic_batched_eval(length, len(A_par), _ptr_array(A_par), _ptr_array(B_par), _ptr(Cp), len(A_seq),
                              _ptr_array(A_seq), _ptr_array(B_seq), _ptr_array(C_seq), _ptr(fq_array(coeffs)), _ptr(out))
    return out


def spmv(row, col, val, num_rows, max_num_cols, z, seg_stride):
    row = np.ascontiguousarray(row, dtype=np.uint32)
    col = np.ascontiguousarray(col, dtype=np.uint32)
    val, z = fq_array(val), fq_array(z)
    out = np.empty((num_rows, 4), dtype=np.uint64)
    lib().ospmv(len(row), _ptr(row), _ptr(col), _ptr(val), num_rows, max_num_cols, _ptr(z), seg_stride, _ptr(out))
    return out


def spmv_batch(row, col, val, num_rows, max_num_cols, z_batch, seg_stride):
    """spmv for every proof of one instance: z_batch is (Q, W * Y, 4); returns (Q * num_rows, 4)."""
    row = np.ascontiguousarray(row, dtype=np.uint32)
    col = np.ascontiguousarray(col, dtype=np.uint32)
    val, zb = fq_array(val), fq_array(z_batch)
    nq = zb.shape[0]
    out = np.empty((nq * num_rows, 4), dtype=np.uint64)
    lib().ospmv_batch(len(row), _ptr(row), _ptr(col), _ptr(val), num_rows, max_num_cols, _ptr(zb), seg_stride, nq,
                      zb.shape[1], _ptr(out))
    return out


def eval_table_sparse(row, col, val, rx, num_segs, max_num_cols, num_cols):
    row = np.ascontiguousarray(row, dtype=np.uint32)
    col = np.ascontiguousarray(col, dtype=np.uint32)
    val, rx = fq_array(val), fq_array(rx)
    out = np.empty((num_segs, num_cols, 4), dtype=np.uint64)
    lib().oeval_table_sparse(len(row), _ptr(row), _ptr(col), _ptr(val), _ptr(rx), num_segs, max_num_cols, num_cols, _ptr(out))
    return out


def sparse_evaluate_with_tables(row, col, val, trx, try_):
    row = np.ascontiguousarray(row, dtype=np.uint32)
    col = np.ascontiguousarray(col, dtype=np.uint32)
    return _ret(lib().osparse_evaluate_with_tables(len(row), _ptr(row), _ptr(col), _ptr(fq_array(val)),
                                                   _ptr(fq_array(trx)), _ptr(fq_array(try_))))


def prod_layer(left, right):
    left, right = fq_array(left), fq_array(right)
    n = left.shape[0]
    ol = np.empty((n // 2, 4), dtype=np.uint64)
    orr = np.empty((n // 2, 4), dtype=np.uint64)
    lib().oprod_layer(_ptr(left), _ptr(right), n, _ptr(ol), _ptr(orr))
    return ol, orr
