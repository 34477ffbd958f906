"""CPU tests of the wire / on-disk formats (spartan_parallel_b200/formats.py):
the proof schemas must consume the committed proof fixtures byte for byte and reproduce them,
a whole `SNARK` (src/lib.rs:700-756) assembled from those proofs must round-trip, and the
`.ctk` / `.rtk` layouts of examples/interface.rs must round-trip and feed the device tables."""
import os

import numpy as np
import pytest

from spartan_parallel_b200 import formats as F

HERE = os.path.dirname(os.path.abspath(__file__))
ONE = np.array([0xD6EC31748D98951D, 0xC6EF5BF4737DCF70, 0xFFFFFFFFFFFFFFFE, 0x0FFFFFFFFFFFFFFF], dtype=np.uint64)


def _golden(name):
    return open(os.path.join(HERE, "golden", name), "rb").read()


def test_r1cs_proof_fixture_round_trips():
    blob = _golden("r1cs_proof_x32_q2.bin")
    pr = F.decode(F.R1CSProof, blob)
    assert len(pr["sc_proof_phase1"]["comm_polys"]) == 6 and len(pr["sc_proof_phase2"]["comm_polys"]) == 6  # log 32 + log 2, 1 + log 32
    assert all(len(p["z"]) == 4 for p in pr["sc_proof_phase1"]["proofs"])  # cubic round polynomials: 4 coefficients
    assert F.encode(F.R1CSProof, pr) == blob
    with pytest.raises(F.FormatError):
        F.decode(F.R1CSProof, blob[:-1])
    with pytest.raises(F.FormatError):
        F.decode(F.R1CSProof, blob + b"\0")


def test_sparse_proof_fixture_round_trips():
    blob = _golden("sparse_proof_3x8.bin")
    pr = F.decode(F.SparseMatPolyEvalProof, blob)
    net = pr["poly_eval_network_proof"]
    assert len(net["proof_prod_layer"]["eval_row"][1]) == 3  # one read claim per matrix of the batch
    assert F.encode(F.SparseMatPolyEvalProof, pr) == blob
    assert F.encode(F.R1CSEvalProof, {"proof": pr}) == blob  # the wrapper adds no bytes


def test_snark_layout_round_trips():
    r1 = F.decode(F.R1CSProof, _golden("r1cs_proof_x32_q2.bin"))
    ev = {"proof": F.decode(F.SparseMatPolyEvalProof, _golden("sparse_proof_3x8.bin"))}
    pe = r1["proof_eval_vars_at_ry_list"][0]
    pc = lambda n, tag: {"C": [bytes([tag, i] + [0] * 30) for i in range(n)]}
    snark = {name: None for name, _ in F.SNARK[1]}
    for name, schema in F.SNARK[1]:
        if schema == F.PolyCommitment:
            snark[name] = pc(2, len(name))
        elif schema == F.Vec(F.PolyCommitment):
            snark[name] = [pc(1, 1), pc(3, 2)]
        elif schema == F.R1CSProof:
            snark[name] = r1
        elif schema == F.R1CSEvalProof:
            snark[name] = ev
        elif schema == F.Vec(F.R1CSEvalProof):
            snark[name] = [ev, ev]
        elif schema == F.Array(F.SCALAR, 3):
            snark[name] = [ONE, ONE * 0, ONE]
        elif schema == F.Vec(F.SCALAR):
            snark[name] = [ONE] * 5
        elif schema == F.Vec(F.PolyEvalProof):
            snark[name] = [pe, pe]
    snark["shift_proof"] = {"proof": pe, "C_orig_evals": [bytes(32)] * 2, "C_shifted_evals": [bytes(32)] * 2,
                            "openings": [[bytes(32)], [bytes(32)] * 3]}
    snark["io_proof"] = {"proofs": [pe]}
    assert all(v is not None for v in snark.values())
    blob = F.encode(F.SNARK, snark)
    back = F.decode(F.SNARK, blob)
    assert F.encode(F.SNARK, back) == blob
    assert back["block_comm_w3_list_shifted"][1]["C"][2] == bytes([2, 2] + [0] * 30)
    assert len(blob) > 3 * len(_golden("r1cs_proof_x32_q2.bin")) + 4 * len(_golden("sparse_proof_3x8.bin"))


def test_ctk_rtk_round_trip(tmp_path):
    le = lambda v: int(v).to_bytes(32, "little")
    ctk = {"block_num_instances": 2, "num_vars": 8, "num_inputs_unpadded": 3, "num_vars_per_block": [8, 4],
           "block_num_phy_ops": [1, 0], "block_num_vir_ops": [0, 2], "max_ts_width": 5,
           "args": [[([(0, le(1))], [(1, le(1)), (2, le(7))], [(3, le(1))]), ([], [(0, le(2))], [])], [([(1, le(1))], [(1, le(1))], [(2, le(1))])]],
           "input_liveness": [True, False, True], "func_input_width": 2, "input_offset": 1, "input_block_num": 0,
           "output_offset": 2, "output_block_num": 1}
    path = tmp_path / "demo_bin.ctk"
    path.write_bytes(F.encode(F.CompileTimeKnowledge, ctk))
    back = F.read_ctk(str(path))
    assert back["args"][0][0][1][1] == (2, le(7)) and back["input_liveness"] == [True, False, True]
    A, B, C = F.ctk_matrices(back, 0)
    assert list(B[0]) == [0, 0, 1] and list(B[1]) == [1, 2, 0] and B[2][1] == le(7)
    assg = lambda n, k: {"assignment": [ONE * 0 + np.uint64(k + i) for i in range(n)]}
    rtk = {"block_max_num_proofs": 2, "block_num_proofs": [2, 1], "consis_num_proofs": 3,
           "total_num_init_phy_mem_accesses": 0, "total_num_init_vir_mem_accesses": 0, "total_num_phy_mem_accesses": 1,
           "total_num_vir_mem_accesses": 0, "block_vars_matrix": [[assg(8, 1), assg(8, 20)], [assg(4, 40)]],
           "exec_inputs": [assg(4, 60), assg(4, 70), assg(4, 80)], "init_phy_mems_list": [], "init_vir_mems_list": [],
           "addr_phy_mems_list": [assg(4, 90)], "addr_vir_mems_list": [], "addr_ts_bits_list": [],
           "input": [le(5)], "input_stack": [], "input_mem": [le(6), le(7)], "output": le(9), "output_exec_num": 2}
    p2 = tmp_path / "demo_bin.rtk"
    p2.write_bytes(F.encode(F.RunTimeKnowledge, rtk))
    back = F.read_rtk(str(p2))
    assert F.encode(F.RunTimeKnowledge, back) == p2.read_bytes()
    tabs = F.block_witness_tables(back)
    assert tabs[0].shape == (2, 8, 4) and tabs[1].shape == (1, 4, 4) and int(tabs[0][1, 3, 0]) == 23
    with pytest.raises(F.FormatError):
        F.decode(F.RunTimeKnowledge, p2.read_bytes()[:40])


def test_hostile_lengths_are_refused():
    with pytest.raises(F.FormatError):
        F.decode(F.Vec(F.SCALAR), (2 ** 40).to_bytes(8, "little"))
    with pytest.raises(F.FormatError):
        F.decode(F.BOOL, b"\x02")
