"""ctypes loader for libspgpu.so (the C-ABI boundary declared in include/spgpu.h).

There is no CPU fallback: if the shared library is missing, or no CUDA device is
visible when a context is created, the error is raised to the caller.
"""
from __future__ import annotations

import ctypes as C
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SPG_LIB") or os.path.join(_HERE, "libspgpu.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "spgpu.h")


class SpgError(RuntimeError):
    pass


class SpgFq(C.Structure):
    _fields_ = [("l", C.c_uint64 * 4)]


_lib = None


def declared_symbols() -> list[str]:
    """Every function name declared in include/spgpu.h."""
    text = open(HEADER_PATH).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(spg_[A-Za-z0-9_]+)\s*\(", text)))


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SpgError(
                f"{LIB_PATH} is missing: build it with `make` (or __graft_entry__.build()); "
                "this backend has no CPU fallback"
            )
        _lib = C.CDLL(LIB_PATH)
        _proto(_lib)
    return _lib


P = C.c_void_p
PP = C.POINTER(C.c_void_p)
SZ = C.c_size_t
INT = C.c_int


def _proto(L):
    L.spg_last_error.restype = C.c_char_p
    L.spg_last_error.argtypes = []
    L.spg_version.restype = INT
    sigs = {
        "spg_ctx_create": [INT, PP],
        "spg_ctx_sync": [P],
        "spg_ctx_profile_begin": [P],
        "spg_ctx_profile_end": [P, P, SZ],
        "spg_host_alloc": [SZ, PP],
        "spg_vec_alloc": [P, SZ, PP],
        "spg_vec_upload": [P, P, SZ, PP],
        "spg_vec_wrap": [P, P, SZ, PP],
        "spg_vec_download": [P, P, SZ, SZ, P],
        "spg_fq_vec_op": [P, INT, P, P, P],
        "spg_fq_from_u512": [P, P, SZ, PP],
        "spg_eq_evals": [P, P, SZ, PP],
        "spg_dense_bound_top": [P, P, P],
        "spg_dense_bound_bot": [P, P, P],
        "spg_dense_evaluate": [P, P, P, SZ, P],
        "spg_dense_bound_L": [P, P, P, SZ, PP],
        "spg_dot": [P, P, P, P],
        "spg_r1cs_create": [P, SZ, SZ, P, SZ, P, P, P, P, PP],
        "spg_r1cs_multi_evaluate": [P, P, P, SZ, P, SZ, P],
        "spg_witness_upload": [P, SZ, P, P, P, PP],
        "spg_witness_upload_async": [P, SZ, P, P, P, PP],
        "spg_sparse_create": [P, SZ, SZ, SZ, P, P, P, P, PP],
        "spg_sparse_view": [P, INT, SZ, PP],
        "spg_sparse_deref": [P, P, P, P, PP],
        "spg_hash_layer_fq": [P, P, P, P, INT, P, P, PP],
        "spg_vec_clone": [P, P, SZ, SZ, PP],
        "spg_witness_poly": [P, SZ, PP],
        "spg_zmat_build": [P, SZ, P, P, SZ, P, PP],
        "spg_sc1_create": [P, P, P, SZ, P, SZ, P, SZ, SZ, P, P, P, PP],
        "spg_sc1_create_from_tables": [P, SZ, P, SZ, P, SZ, P, P, P, P, P, P, PP],
        "spg_sc1_set_scale": [P, P],
        "spg_sc1_set_claim": [P, P],
        "spg_fq_host_sum": [P, SZ, SZ, P],
        "spg_fq_host_mul": [P, P, P],
        "spg_fq_host_eq_weight": [P, SZ, C.c_uint64, P],
        "spg_sc1_round_eval": [P, P],
        "spg_sc1_round_bind": [P, P],
        "spg_sc1_run_rounds": [P, SZ, P, P],
        "spg_sc1_run_rounds_sharded": [P, SZ, P, P, P, SZ, INT, INT, P],
        "spg_sc2_run_rounds": [P, SZ, P, P],
        "spg_sc1_final": [P, P],
        "spg_sc1_debug_tables": [P, P, P, P, SZ, P],
        "spg_sc2_create": [P, P, P, SZ, P, SZ, P, SZ, SZ, P, P, P, P, P, P, PP],
        "spg_sc2_create_from_zrq": [P, P, P, SZ, P, SZ, SZ, P, P, P, P, P, PP],
        "spg_zmat_bind_rq": [P, P, P, SZ, P, P],
        "spg_sc2_round_eval": [P, P],
        "spg_sc2_round_bind": [P, P],
        "spg_sc2_final": [P, P],
        "spg_prodtree_build": [P, P, PP],
        "spg_prodtree_layer": [P, SZ, PP, PP],
        "spg_prodtree_evaluate": [P, P, P],
        "spg_cubic_create": [P, SZ, P, P, P, SZ, P, P, P, P, PP],
        "spg_cubic_round_eval": [P, P],
        "spg_cubic_round_bind": [P, P],
        "spg_cubic_final": [P, P],
        "spg_hash_layer": [P, P, P, P, SZ, P, P, INT, PP],
        "spg_deref": [P, P, SZ, P, PP],
        "spg_perm_scan": [P, SZ, P, SZ, P, SZ, SZ, P, SZ, SZ, P, SZ, SZ, P, SZ, SZ],
        "spg_peer_alloc": [P, SZ, PP, P],
        "spg_peer_free": [P],
        "spg_peer_open": [P, P, PP],
        "spg_peer_close": [P],
        "spg_peer_sum": [P, P, INT, INT, SZ],
        "spg_bullet_create": [P, P, SZ, PP],
        "spg_bullet_lr": [P, SZ, P, P, P],
        "spg_bullet_fold": [P, SZ, P, P],
        "spg_bullet_final": [P, P],
        "spg_bullet_set_ab": [P, P, P],
        "spg_bullet_lr_resident": [P, SZ, P, INT, P, P],
        "spg_bullet_final_ab": [P, P, P],
        "spg_gens_upload": [P, P, SZ, PP],
        "spg_gens_from_uniform": [P, P, SZ, PP],
        "spg_poly_commit": [P, P, P, SZ, P],
        "spg_commit_batch": [P, P, P, SZ, P, SZ, P],
        "spg_poly_commit_rows": [P, P, P, SZ, SZ, SZ, P],
        "spg_mailbox_all_gather": [P, SZ, INT, INT, P, P, SZ, P],
        "spg_sc1_set_row_weights": [P, P, SZ],
        "spg_sc1_host_tail_eval": [P, SZ, SZ, P, P],
        "spg_sc1_host_tail_bind": [P, SZ, SZ, P],
        "spg_sc2_create_slice": [P, P, P, SZ, SZ, SZ, SZ, P, P, P, P, PP],
        "spg_sc2_run_rounds_sharded": [P, SZ, P, P, P, SZ, INT, INT, P],
        "spg_sc2_host_tail_eval": [P, SZ, SZ, INT, P, P],
        "spg_sc2_host_tail_bind": [P, SZ, SZ, INT, P],
        "spg_peer_reduce_scatter": [P, P, INT, INT, SZ],
        "spg_zmat_bind_rq_sharded": [P, P, P, SZ, SZ, P, INT, INT, SZ, INT, P, SZ, P],
        "spg_sc1_set_claim_checked": [P, P],
        "spg_sc1_set_satisfied": [P],
        "spg_zmat_bind_weights": [P, P, P, SZ, P, P],
        "spg_vec_zero": [P, P],
        "spg_wit_perm_w0": [P, P, P, SZ, SZ, PP],
        "spg_wit_block": [P, INT, P, SZ, SZ, P, P, P, SZ, SZ, SZ, SZ, SZ, P, SZ, PP, PP],
        "spg_wit_mem": [P, P, SZ, SZ, P, P, SZ, PP, PP],
        "spg_wit_shift": [P, P, SZ, SZ, P, SZ, PP],
        "spg_gens_prepare": [P, P, SZ],
        "spg_gens_info": [P, P],
        "spg_gens_prepare_rows": [P, P, SZ, SZ],
        "spg_gens_info_rows": [P, P],
        "spg_debug_fe8_selftest": [P, SZ, C.c_uint64, P],
        "spg_debug_fq_wide_selftest": [P, SZ, C.c_uint64, P],
    }
    for name, args in sigs.items():
        f = getattr(L, name, None)
        if f is None:
            continue
        f.restype = INT
        f.argtypes = args
    for name in ("spg_ctx_destroy", "spg_host_free", "spg_vec_free", "spg_r1cs_destroy", "spg_witness_destroy",
                 "spg_zmat_destroy", "spg_sc1_destroy", "spg_sc2_destroy", "spg_prodtree_destroy",
                 "spg_cubic_destroy", "spg_gens_destroy", "spg_sparse_destroy", "spg_bullet_destroy"):
        f = getattr(L, name, None)
        if f is not None:
            f.restype = None
            f.argtypes = [P]
    for name in ("spg_vec_len", "spg_sc1_num_rounds", "spg_sc2_num_rounds", "spg_prodtree_num_layers", "spg_sparse_num_ops",
                 "spg_sparse_num_mem_cells"):
        f = getattr(L, name, None)
        if f is not None:
            f.restype = SZ
            f.argtypes = [P]
    L.spg_mailbox_poison.restype = None
    L.spg_mailbox_poison.argtypes = [P, SZ, INT, INT]
    L.spg_ctx_launch_count.restype = C.c_uint64
    L.spg_ctx_launch_count.argtypes = [P]
    L.spg_ctx_stream.restype = P
    L.spg_ctx_stream.argtypes = [P]
    L.spg_vec_device_ptr.restype = P
    L.spg_vec_device_ptr.argtypes = [P]


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = lib().spg_last_error().decode("utf-8", "replace")
        raise SpgError(f"{what or 'libspgpu'} failed with status {rc}: {msg}")
