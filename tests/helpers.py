"""Shared helpers for the tests: seeded inputs, conversions between the device's
natural-order tables and the oracle's (p, q_rev, w, x_rev) storage."""
import numpy as np

from oracle import cbind as O

M64 = (1 << 64) - 1


def splitmix64(state: int):
    state = (state + 0x9E3779B97F4A7C15) & M64
    z = state
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
    return state, z ^ (z >> 31)


def rand_scalars(n: int, seed: int) -> np.ndarray:
    """n uniform field elements: numpy PCG -> 8 u64 -> from_u512 (oracle)."""
    rng = np.random.default_rng(seed)
    wide = rng.integers(0, 1 << 64, size=(n, 8), dtype=np.uint64)
    return O.vec_from_u512(wide)


def rand_scalar(seed: int) -> np.ndarray:
    return rand_scalars(1, seed)[0]


def log2(n: int) -> int:
    return n.bit_length() - 1


def drive_sc1_oracle(num_proofs, max_q, num_cons, max_x, Az, Bz, Cz, tau_p, tau_q, tau_x, challenges):
    """Runs the oracle's phase-1 loops on natural-order ragged tables; returns
    (list of per-round (e0,e2,e3), final claims)."""
    P = len(num_proofs)
    nx, nq, np_ = log2(max_x), log2(max_q), log2(1 << (P - 1).bit_length() if P > 1 else 1)
    mk = lambda T: O.Pqx.new_rev(T, 1, num_proofs, max_q, num_cons, max_x)
    Ap = O.eq_evals(tau_p) if np_ else O.ONE.reshape(1, 4)
    Aq = O.eq_evals(tau_q) if nq else O.ONE.reshape(1, 4)
    Ax = O.eq_evals(tau_x) if nx else O.ONE.reshape(1, 4)
    sc = O.Sc1(nx, nq, np_, num_proofs, num_cons, Ap, Aq, Ax, mk(Az), mk(Bz), mk(Cz))
    evals = []
    for j in range(sc.num_rounds):
        evals.append(sc.round_eval())
        sc.round_bind(challenges[j])
    return evals, sc.final()


def random_instance(num_inst, num_cons, W, Ymax, Y, nnz, seed):
    """Random sparse A, B, C per instance with columns confined to the first Y[i] inputs of
    the first W sections (what R1CSProof::prove's z indexing requires). Not satisfiable in
    general -- the prover loops do not care, and parity must hold for any tables."""
    from oracle.r1cs import Instance

    rng = np.random.default_rng(seed)
    Wp = 1 << (W - 1).bit_length() if W > 1 else 1
    mats = []
    for i in range(num_inst):
        for m in range(3):
            n = nnz
            rows = rng.integers(0, num_cons[i], size=n, dtype=np.uint32)
            cols = (rng.integers(0, W, size=n, dtype=np.uint32) * Ymax + rng.integers(0, Y[i], size=n, dtype=np.uint32)).astype(np.uint32)
            vals = O.vec_from_u512(rng.integers(0, 1 << 64, size=(n, 8), dtype=np.uint64))
            # sprinkle unit and zero coefficients
            vals[::3] = O.ONE
            if n > 4:
                vals[4] = 0
            mats.append((rows, cols, vals))
    return Instance(num_inst, max(num_cons), list(num_cons), Wp * Ymax, mats)


def random_witness_secs(P, num_proofs, W, sec_inputs, kinds, seed):
    """kinds[w] in {'full', 'short', 'single'}; sec_inputs[w][p] = section width."""
    from oracle.r1cs import WitnessSec

    secs = []
    for w in range(W):
        n_inst = 1 if kinds[w] == "single" else P
        w_mat = []
        for p in range(n_inst):
            nq = 1 if kinds[w] in ("single", "short") else num_proofs[p]
            w_mat.append([rand_scalars(sec_inputs[w][p], seed + 1000 * w + 10 * p + q) for q in range(nq)])
        secs.append(WitnessSec(list(sec_inputs[w][:n_inst]), w_mat))
    return secs
