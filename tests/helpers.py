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
