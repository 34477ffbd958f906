"""ctypes binding of libspghost.so: the C++ mirror of the reference's transcript-side host
code (merlin transcript, RandomTape, sigma protocols, UniPoly, ZK sumcheck glue, opening
proofs, bincode layout). In production this layer is the unmodified Rust crate; see
INTEGRATION.md. It drives the device exclusively through the C ABI of libspgpu.so."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from ._lib import SpgError

_HERE = os.path.dirname(os.path.abspath(__file__))
HOST_LIB_PATH = os.path.join(_HERE, "libspghost.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(HOST_LIB_PATH):
            raise SpgError(f"{HOST_LIB_PATH} is missing: build it with `make`")
        from . import _lib as dev

        dev.lib()  # libspgpu.so first (libspghost links against it)
        L = C.CDLL(HOST_LIB_PATH)
        L.sph_last_error.restype = C.c_char_p
        L.sph_free.argtypes = [C.c_void_p]
        L.sph_free.restype = None
        L.sph_gens_derive.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p]
        L.sph_transcript_kat.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_size_t, C.c_void_p]
        L.sph_r1cs_prove.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p,
                                     C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.POINTER(C.c_void_p),
                                     C.POINTER(C.c_size_t), C.c_void_p, C.c_void_p]
        L.sph_r1cs_gens_new.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
        L.sph_r1cs_gens_new.restype = C.c_void_p
        L.sph_r1cs_gens_free.argtypes = [C.c_void_p]
        L.sph_r1cs_gens_free.restype = None
        L.sph_r1cs_gens_device_pc.argtypes = [C.c_void_p]
        L.sph_r1cs_gens_device_pc.restype = C.c_void_p
        L.sph_sparse_prove.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.POINTER(C.c_void_p),
                                       C.POINTER(C.c_size_t)]
        L.sph_sparse_gens_new.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t]
        L.sph_sparse_gens_new.restype = C.c_void_p
        L.sph_sparse_gens_free.argtypes = [C.c_void_p]
        L.sph_sparse_gens_free.restype = None
        L.sph_polyeval_prove.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_char_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p,
                                         C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t),
                                         C.c_void_p]
        L.sph_timings.argtypes = [C.c_char_p, C.c_size_t]
        L.sph_timings.restype = C.c_size_t
        L.sph_timings_reset.restype = None
        _lib = L
    return _lib


def timings_reset():
    lib().sph_timings_reset()


def timings() -> list:
    """[(stage label, milliseconds)] recorded by the host mirror since timings_reset()."""
    n = lib().sph_timings(None, 0)
    buf = C.create_string_buffer(n)
    lib().sph_timings(buf, n)
    out = []
    for ln in buf.value.decode().splitlines():
        k, _, v = ln.rpartition("\t")
        out.append((k.strip(), float(v)))
    return out


def _check(rc, what):
    if rc != 0:
        raise SpgError(f"{what}: {lib().sph_last_error().decode('utf-8', 'replace')}")


def gens_derive(label: bytes, n: int) -> bytes:
    """MultiCommitGens::new(n, label) (src/commitments.rs:15-33) as n + 1 compressed points."""
    out = np.empty(32 * (n + 1), dtype=np.uint8)
    _check(lib().sph_gens_derive(label, n, out.ctypes.data_as(C.c_void_p)), "sph_gens_derive")
    return out.tobytes()


def transcript_kat(label: bytes, l: bytes, m: bytes, c: bytes, n: int) -> bytes:
    out = np.empty(n, dtype=np.uint8)
    lib().sph_transcript_kat(label, l, m, c, n, out.ctypes.data_as(C.c_void_p))
    return out.tobytes()


def _sz(v):
    return np.ascontiguousarray(np.asarray(v, dtype=np.uint64).reshape(-1))


class R1CSGens:
    """R1CSGens::new (src/r1csproof.rs:45-80) with the opening-proof generators also resident on the
    device: the n-sized multiscalar multiplications of DotProductProofLog / BulletReductionProof then
    run there. Create once, reuse across proofs (like the reference's SNARKGens)."""

    def __init__(self, ctx, label: bytes, num_vars: int):
        self.ctx = ctx
        self.h = lib().sph_r1cs_gens_new(ctx.h, label, num_vars)
        if not self.h:
            raise SpgError(f"sph_r1cs_gens_new: {lib().sph_last_error().decode('utf-8', 'replace')}")

    def gens_pc(self):
        """gens_pc.gens.gens_n on the device as a MultiCommitGens view (not owned): the bases of the
        witness commitments (src/dense_mlpoly.rs:214-239) and of the openings' MSMs."""
        from . import api

        g = api.MultiCommitGens.__new__(api.MultiCommitGens)
        g.ctx, g.h, g.n, g._borrowed = self.ctx, C.c_void_p(lib().sph_r1cs_gens_device_pc(self.h)), None, self
        g.free = lambda: None
        return g

    def free(self):
        if getattr(self, "h", None):
            lib().sph_r1cs_gens_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def r1cs_prove(ctx, inst, witness_secs, num_proofs, max_num_proofs, num_inputs, max_num_inputs, transcript_label: bytes,
               gens_label: bytes, tape_seed, gens_num_vars: int, gens: "R1CSGens | None" = None):
    """R1CSProof::prove (src/r1csproof.rs:210-685) with a caller-seeded RandomTape.
    Returns (proof bytes in bincode layout, [rp, rq_rev, rx, rw ++ ry])."""
    P = len(num_proofs)
    secs = (C.c_void_p * len(witness_secs))(*[w.h for w in witness_secs])
    sec_ni = _sz([len(w.num_proofs) for w in witness_secs])
    sec_np = _sz([q for w in witness_secs for q in w.num_proofs])
    sec_nin = _sz([y for w in witness_secs for y in w.num_inputs])
    seed = np.ascontiguousarray(tape_seed, dtype=np.uint64)
    out_bytes, out_len = C.c_void_p(), C.c_size_t()
    ch = np.zeros((256, 4), dtype=np.uint64)
    counts = np.zeros(4, dtype=np.uint64)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    npf, nin, inc = _sz(num_proofs), _sz(num_inputs), _sz(inst.num_cons)
    _check(lib().sph_r1cs_prove(ctx.h, transcript_label, gens_label, p(seed), P, max_num_proofs, p(npf), max_num_inputs, p(nin),
                                len(witness_secs), secs, p(sec_ni), p(sec_np), p(sec_nin), inst.h, inst.num_instances,
                                inst.max_num_cons, p(inc), gens_num_vars, gens.h if gens else None, C.byref(out_bytes), C.byref(out_len), p(ch), p(counts)),
           "sph_r1cs_prove")
    proof = C.string_at(out_bytes, out_len.value)
    lib().sph_free(out_bytes)
    outs, pos = [], 0
    for c in counts:
        outs.append(ch[pos:pos + int(c)].copy())
        pos += int(c)
    return proof, outs


class SparseGens:
    """SparseMatPolyCommitmentGens::new (src/sparse_mlpoly.rs:289-316) with bases and window tables on the
    device. Create once, reuse across proofs (like the reference's SNARKGens)."""

    def __init__(self, ctx, label: bytes, num_vars_x: int, num_vars_y: int, max_nz: int, batch: int):
        self.ctx = ctx
        self.h = lib().sph_sparse_gens_new(ctx.h, label, num_vars_x, num_vars_y, max_nz, batch)
        if not self.h:
            raise SpgError(f"sph_sparse_gens_new: {lib().sph_last_error().decode('utf-8', 'replace')}")

    def free(self):
        if getattr(self, "h", None):
            lib().sph_sparse_gens_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def sparse_prove(ctx, polys, num_vars_x: int, num_vars_y: int, rx, ry, evals, transcript_label: bytes, gens_label: bytes,
                 tape_seed, gens: "SparseGens | None" = None):
    """SparseMatPolynomial::multi_commit + SparseMatPolyEvalProof::prove (src/sparse_mlpoly.rs:566-586,
    1509-1564) for a batch of matrices `polys` = [(rows, cols, vals), ...].
    Returns (SparseMatPolyCommitment bytes, SparseMatPolyEvalProof bytes), both in bincode layout."""
    nnz = _sz([len(p_[0]) for p_ in polys])
    rows = np.ascontiguousarray(np.concatenate([np.asarray(p_[0], dtype=np.uint32) for p_ in polys]))
    cols = np.ascontiguousarray(np.concatenate([np.asarray(p_[1], dtype=np.uint32) for p_ in polys]))
    vals = np.ascontiguousarray(np.concatenate([np.asarray(p_[2], dtype=np.uint64).reshape(-1, 4) for p_ in polys]))
    fq = lambda a, n: np.ascontiguousarray(np.asarray(a, dtype=np.uint64).reshape(n, 4)) if n else np.zeros((1, 4), dtype=np.uint64)
    rx_, ry_, ev = fq(rx, num_vars_x), fq(ry, num_vars_y), fq(evals, len(polys))
    seed = np.ascontiguousarray(tape_seed, dtype=np.uint64)
    oc, ocl, op, opl = C.c_void_p(), C.c_size_t(), C.c_void_p(), C.c_size_t()
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    _check(lib().sph_sparse_prove(ctx.h, transcript_label, gens_label, p(seed), len(polys), num_vars_x, num_vars_y, p(nnz),
                                  p(rows), p(cols), p(vals), p(rx_), p(ry_), p(ev), gens.h if gens else None, C.byref(oc),
                                  C.byref(ocl), C.byref(op), C.byref(opl)), "sph_sparse_prove")
    comm, proof = C.string_at(oc, ocl.value), C.string_at(op, opl.value)
    lib().sph_free(oc)
    lib().sph_free(op)
    return comm, proof


def polyeval_prove(ctx, variant: str, polys, r, Zr, transcript_label: bytes, gens_label: bytes, tape_seed, gens_n: int):
    """PolyEvalProof::prove_batched_points ("points": one polynomial, r = list of points),
    ::prove_batched_instances ("instances": one point per polynomial) or ::prove_uni_batched_instances
    ("uni": r = one scalar) over device polynomials (src/dense_mlpoly.rs:531, 689, 1046).
    Returns the bincode bytes (and C_Zr_prime for "uni")."""
    v = {"points": 0, "instances": 1, "uni": 2}[variant]
    hs = (C.c_void_p * len(polys))(*[p.h for p in polys])
    Zr_ = np.ascontiguousarray(np.asarray(Zr, dtype=np.uint64).reshape(-1, 4))
    if v == 2:
        num_points, r_len = len(polys), 1
        r_ = np.ascontiguousarray(np.tile(np.asarray(r, dtype=np.uint64).reshape(1, 4), (num_points, 1)))
    else:
        num_points = len(r)
        r_len = len(r[0])
        assert all(len(x) == r_len for x in r), "points must have equal length (pad on the caller's side)"
        r_ = np.ascontiguousarray(np.asarray([np.asarray(x, dtype=np.uint64).reshape(r_len, 4) for x in r], dtype=np.uint64).reshape(-1, 4))
    seed = np.ascontiguousarray(tape_seed, dtype=np.uint64)
    out, out_len = C.c_void_p(), C.c_size_t()
    extra = np.zeros(32, dtype=np.uint8)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    _check(lib().sph_polyeval_prove(ctx.h, v, transcript_label, gens_label, p(seed), gens_n, len(polys), hs, num_points, r_len, p(r_), p(Zr_),
                                    C.byref(out), C.byref(out_len), p(extra)), "sph_polyeval_prove")
    blob = C.string_at(out, out_len.value)
    lib().sph_free(out)
    return (blob, extra.tobytes()) if v == 2 else blob
