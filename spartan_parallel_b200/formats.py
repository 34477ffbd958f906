"""Wire and on-disk formats around the accelerated path, as schemas for a small
bincode 1.x codec (default options: little-endian fixed-width integers, u64 length prefix for
Vec, nothing for tuples / arrays / structs, bool = one byte):

  * the proof structs the prover emits -- `R1CSProof` (src/r1csproof.rs:25-43),
    `SparseMatPolyEvalProof` / `R1CSEvalProof` (src/sparse_mlpoly.rs:1466-1476,
    src/r1csinstance.rs:737-740), `PolyEvalProof`, `ShiftProofs`, `IOProofs` and the whole `SNARK`
    (src/lib.rs:700-756);
  * the inputs `examples/interface.rs` reads: `CompileTimeKnowledge` (.ctk, :46-71) and
    `RunTimeKnowledge` (.rtk, :197-220).

A `Scalar` serialises as its four Montgomery limbs (serde derive on `Scalar([u64; 4])`,
src/scalar/ristretto255.rs:198), a `CompressedRistretto` as 32 bytes, `usize` as u64.

decode(schema, bytes) -> python value (dict / list / int / bytes / numpy uint64[4]);
encode(schema, value) -> bytes. Both are exact inverses on well-formed input, which is what
tests/test_formats.py checks against the committed proof fixtures.
"""
from __future__ import annotations

import struct

import numpy as np

# ---------------------------------------------------------------- schema constructors
U64, BOOL, SCALAR, POINT, BYTES32 = "u64", "bool", "scalar", "point", "bytes32"


def Vec(t):
    return ("vec", t)


def Tuple(*ts):
    return ("tuple", list(ts))


def Array(t, n):
    return ("array", t, n)


def Struct(*fields):
    return ("struct", list(fields))


# ---------------------------------------------------------------- codec
class FormatError(ValueError):
    pass


def _decode(schema, b: memoryview, pos: int):
    if schema == U64:
        if pos + 8 > len(b):
            raise FormatError("truncated u64")
        return struct.unpack_from("<Q", b, pos)[0], pos + 8
    if schema == BOOL:
        if pos + 1 > len(b) or b[pos] > 1:
            raise FormatError("bad bool")
        return bool(b[pos]), pos + 1
    if schema == SCALAR:
        if pos + 32 > len(b):
            raise FormatError("truncated scalar")
        return np.frombuffer(b, dtype="<u8", count=4, offset=pos).astype(np.uint64), pos + 32
    if schema in (POINT, BYTES32):
        if pos + 32 > len(b):
            raise FormatError("truncated 32-byte field")
        return bytes(b[pos: pos + 32]), pos + 32
    kind = schema[0]
    if kind == "vec":
        n, pos = _decode(U64, b, pos)
        if n > len(b):  # every element takes at least one byte
            raise FormatError(f"vector length {n} exceeds the input")
        out = []
        for _ in range(n):
            v, pos = _decode(schema[1], b, pos)
            out.append(v)
        return out, pos
    if kind == "tuple":
        out = []
        for t in schema[1]:
            v, pos = _decode(t, b, pos)
            out.append(v)
        return tuple(out), pos
    if kind == "array":
        out = []
        for _ in range(schema[2]):
            v, pos = _decode(schema[1], b, pos)
            out.append(v)
        return out, pos
    if kind == "struct":
        out = {}
        for name, t in schema[1]:
            out[name], pos = _decode(t, b, pos)
        return out, pos
    raise FormatError(f"unknown schema {schema!r}")


def decode(schema, data: bytes, exact: bool = True):
    v, pos = _decode(schema, memoryview(data), 0)
    if exact and pos != len(data):
        raise FormatError(f"{len(data) - pos} trailing bytes")
    return v


def _encode(schema, v, out: bytearray):
    if schema == U64:
        out += struct.pack("<Q", int(v))
    elif schema == BOOL:
        out.append(1 if v else 0)
    elif schema == SCALAR:
        a = np.asarray(v, dtype=np.uint64).reshape(4)
        out += a.astype("<u8").tobytes()
    elif schema in (POINT, BYTES32):
        if len(v) != 32:
            raise FormatError("32-byte field expected")
        out += bytes(v)
    else:
        kind = schema[0]
        if kind == "vec":
            out += struct.pack("<Q", len(v))
            for x in v:
                _encode(schema[1], x, out)
        elif kind == "tuple":
            if len(v) != len(schema[1]):
                raise FormatError("tuple arity")
            for t, x in zip(schema[1], v):
                _encode(t, x, out)
        elif kind == "array":
            if len(v) != schema[2]:
                raise FormatError("array length")
            for x in v:
                _encode(schema[1], x, out)
        elif kind == "struct":
            for name, t in schema[1]:
                _encode(t, v[name], out)
        else:
            raise FormatError(f"unknown schema {schema!r}")


def encode(schema, value) -> bytes:
    out = bytearray()
    _encode(schema, value, out)
    return bytes(out)


# ---------------------------------------------------------------- proof structs
PolyCommitment = Struct(("C", Vec(POINT)))                                   # src/dense_mlpoly.rs:44-47
BulletReductionProof = Struct(("L_vec", Vec(POINT)), ("R_vec", Vec(POINT)))  # src/nizk/bullet.rs:18-22
DotProductProofLog = Struct(("bullet_reduction_proof", BulletReductionProof), ("delta", POINT), ("beta", POINT),
                            ("z1", SCALAR), ("z2", SCALAR))                  # src/nizk/mod.rs:421-428
PolyEvalProof = Struct(("proof", DotProductProofLog))                        # src/dense_mlpoly.rs:425-428
DotProductProof = Struct(("delta", POINT), ("beta", POINT), ("z", Vec(SCALAR)), ("z_delta", SCALAR), ("z_beta", SCALAR))
KnowledgeProof = Struct(("alpha", POINT), ("z1", SCALAR), ("z2", SCALAR))    # src/nizk/mod.rs:15-20
ProductProof = Struct(("alpha", POINT), ("beta", POINT), ("delta", POINT), ("z", Array(SCALAR, 5)))
EqualityProof = Struct(("alpha", POINT), ("z", SCALAR))
ZKSumcheckInstanceProof = Struct(("comm_polys", Vec(POINT)), ("comm_evals", Vec(POINT)), ("proofs", Vec(DotProductProof)))
R1CSProof = Struct(                                                          # src/r1csproof.rs:25-43
    ("sc_proof_phase1", ZKSumcheckInstanceProof),
    ("claims_phase2", Tuple(POINT, POINT, POINT, POINT)),
    ("pok_claims_phase2", Tuple(KnowledgeProof, ProductProof)),
    ("proof_eq_sc_phase1", EqualityProof),
    ("sc_proof_phase2", ZKSumcheckInstanceProof),
    ("comm_vars_at_ry_list", Vec(Vec(POINT))),
    ("comm_vars_at_ry", POINT),
    ("proof_eval_vars_at_ry_list", Vec(PolyEvalProof)),
    ("proof_eq_sc_phase2", EqualityProof),
)
CompressedUniPoly = Struct(("coeffs_except_linear_term", Vec(SCALAR)))       # src/unipoly.rs:14-18
SumcheckInstanceProof = Struct(("compressed_polys", Vec(CompressedUniPoly)))
LayerProofBatched = Struct(("proof", SumcheckInstanceProof), ("claims_prod_left", Vec(SCALAR)), ("claims_prod_right", Vec(SCALAR)))
ProductCircuitEvalProofBatched = Struct(("proof", Vec(LayerProofBatched)),
                                        ("claims_dotp", Tuple(Vec(SCALAR), Vec(SCALAR), Vec(SCALAR))))  # src/product_tree.rs:128-131
_side = Tuple(SCALAR, Vec(SCALAR), Vec(SCALAR), SCALAR)
ProductLayerProof = Struct(("eval_row", _side), ("eval_col", _side), ("eval_val", Tuple(Vec(SCALAR), Vec(SCALAR))),
                           ("proof_mem", ProductCircuitEvalProofBatched), ("proof_ops", ProductCircuitEvalProofBatched))
_hside = Tuple(Vec(SCALAR), Vec(SCALAR), SCALAR)
HashLayerProof = Struct(("eval_row", _hside), ("eval_col", _hside), ("eval_val", Vec(SCALAR)),
                        ("eval_derefs", Tuple(Vec(SCALAR), Vec(SCALAR))), ("proof_ops", PolyEvalProof), ("proof_mem", PolyEvalProof),
                        ("proof_derefs", PolyEvalProof))                     # src/sparse_mlpoly.rs:739-748
PolyEvalNetworkProof = Struct(("proof_prod_layer", ProductLayerProof), ("proof_hash_layer", HashLayerProof))
DerefsCommitment = Struct(("comm_ops_val", PolyCommitment))
SparseMatPolyEvalProof = Struct(("comm_derefs", DerefsCommitment), ("poly_eval_network_proof", PolyEvalNetworkProof))
R1CSEvalProof = Struct(("proof", SparseMatPolyEvalProof))                    # src/r1csinstance.rs:737-740
SparseMatPolyCommitment = Struct(("batch_size", U64), ("num_ops", U64), ("num_mem_cells", U64),
                                 ("comm_comb_ops", PolyCommitment), ("comm_comb_mem", PolyCommitment))
ShiftProofs = Struct(("proof", PolyEvalProof), ("C_orig_evals", Vec(POINT)), ("C_shifted_evals", Vec(POINT)),
                     ("openings", Vec(Vec(POINT))))                         # src/lib.rs:365-370
IOProofs = Struct(("proofs", Vec(PolyEvalProof)))                            # src/lib.rs:189-196

_three = lambda stem: [(f"{stem}_comm_w2", PolyCommitment), (f"{stem}_comm_w3", PolyCommitment), (f"{stem}_comm_w3_shifted", PolyCommitment)]
SNARK = Struct(                                                              # src/lib.rs:700-756
    ("block_comm_vars_list", Vec(PolyCommitment)),
    ("exec_comm_inputs", Vec(PolyCommitment)),
    ("addr_comm_phy_mems", PolyCommitment),
    ("addr_comm_phy_mems_shifted", PolyCommitment),
    ("addr_comm_vir_mems", PolyCommitment),
    ("addr_comm_vir_mems_shifted", PolyCommitment),
    ("addr_comm_ts_bits", PolyCommitment),
    ("perm_exec_comm_w2_list", PolyCommitment),
    ("perm_exec_comm_w3_list", PolyCommitment),
    ("perm_exec_comm_w3_shifted", PolyCommitment),
    ("block_comm_w2_list", Vec(PolyCommitment)),
    ("block_comm_w3_list", Vec(PolyCommitment)),
    ("block_comm_w3_list_shifted", Vec(PolyCommitment)),
    *_three("init_phy_mem"), *_three("init_vir_mem"), *_three("phy_mem_addr"), *_three("vir_mem_addr"),
    ("block_r1cs_sat_proof", R1CSProof),
    ("block_inst_evals_bound_rp", Array(SCALAR, 3)),
    ("block_inst_evals_list", Vec(SCALAR)),
    ("block_r1cs_eval_proof_list", Vec(R1CSEvalProof)),
    ("pairwise_check_r1cs_sat_proof", R1CSProof),
    ("pairwise_check_inst_evals_bound_rp", Array(SCALAR, 3)),
    ("pairwise_check_inst_evals_list", Vec(SCALAR)),
    ("pairwise_check_r1cs_eval_proof", R1CSEvalProof),
    ("perm_root_r1cs_sat_proof", R1CSProof),
    ("perm_root_inst_evals", Array(SCALAR, 3)),
    ("perm_root_r1cs_eval_proof", R1CSEvalProof),
    ("perm_poly_poly_list", Vec(SCALAR)),
    ("proof_eval_perm_poly_prod_list", Vec(PolyEvalProof)),
    ("shift_proof", ShiftProofs),
    ("io_proof", IOProofs),
)

# ---------------------------------------------------------------- .ctk / .rtk (examples/interface.rs)
_entry = Tuple(U64, BYTES32)  # (variable index, coefficient as 32 little-endian bytes)
CompileTimeKnowledge = Struct(                                               # examples/interface.rs:46-71
    ("block_num_instances", U64), ("num_vars", U64), ("num_inputs_unpadded", U64), ("num_vars_per_block", Vec(U64)),
    ("block_num_phy_ops", Vec(U64)), ("block_num_vir_ops", Vec(U64)), ("max_ts_width", U64),
    ("args", Vec(Vec(Tuple(Vec(_entry), Vec(_entry), Vec(_entry))))),        # per block, per constraint: (A, B, C) rows
    ("input_liveness", Vec(BOOL)), ("func_input_width", U64), ("input_offset", U64), ("input_block_num", U64),
    ("output_offset", U64), ("output_block_num", U64),
)
Assignment = Struct(("assignment", Vec(SCALAR)))                             # src/lib.rs:88-92
RunTimeKnowledge = Struct(                                                   # examples/interface.rs:197-220
    ("block_max_num_proofs", U64), ("block_num_proofs", Vec(U64)), ("consis_num_proofs", U64),
    ("total_num_init_phy_mem_accesses", U64), ("total_num_init_vir_mem_accesses", U64),
    ("total_num_phy_mem_accesses", U64), ("total_num_vir_mem_accesses", U64),
    ("block_vars_matrix", Vec(Vec(Assignment))), ("exec_inputs", Vec(Assignment)),
    ("init_phy_mems_list", Vec(Assignment)), ("init_vir_mems_list", Vec(Assignment)),
    ("addr_phy_mems_list", Vec(Assignment)), ("addr_vir_mems_list", Vec(Assignment)), ("addr_ts_bits_list", Vec(Assignment)),
    ("input", Vec(BYTES32)), ("input_stack", Vec(BYTES32)), ("input_mem", Vec(BYTES32)), ("output", BYTES32),
    ("output_exec_num", U64),
)


def read_ctk(path: str) -> dict:
    return decode(CompileTimeKnowledge, open(path, "rb").read())


def read_rtk(path: str) -> dict:
    return decode(RunTimeKnowledge, open(path, "rb").read())


def block_witness_tables(rtk: dict):
    """block_vars_matrix of a RunTimeKnowledge as what the device sections take: per block instance a
    (num_proofs, width, 4) uint64 array of Montgomery scalars (rows = executions of the block)."""
    out = []
    for inst in rtk["block_vars_matrix"]:
        out.append(np.stack([np.stack(a["assignment"]) for a in inst]) if inst else np.zeros((0, 0, 4), dtype=np.uint64))
    return out


def ctk_matrices(ctk: dict, block: int):
    """COO triples (rows, cols, 32-byte little-endian coefficients) of block `block`'s A, B, C as the
    front end lists them (one (A, B, C) row per constraint)."""
    mats = [([], [], []) for _ in range(3)]
    for row, cons in enumerate(ctk["args"][block]):
        for m in range(3):
            for col, coeff in cons[m]:
                mats[m][0].append(row)
                mats[m][1].append(col)
                mats[m][2].append(coeff)
    return [(np.asarray(r, dtype=np.uint32), np.asarray(c, dtype=np.uint32), list(v)) for r, c, v in mats]
