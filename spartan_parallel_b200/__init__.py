"""B200-native prover backend for spartan-parallel's data-parallel R1CS proving path.

The product is ``libspgpu.so`` (hand-written sm_100a CUDA kernels behind the C ABI
of ``include/spgpu.h``); this package is the thin host-side mirror of the reference
interface used by the tests and the benchmark. It never imports ``oracle``.
"""
from ._lib import LIB_PATH, SpgError, declared_symbols  # noqa: F401
from .api import (  # noqa: F401
    MODE_P,
    MODE_Q,
    MODE_W,
    MODE_X,
    BulletReduction,
    Context,
    CubicBatched,
    DensePolynomial,
    EqPolynomial,
    MultiCommitGens,
    MultiSparseMatPolynomialAsDense,
    ProductCircuit,
    ProverWitnessSecInfo,
    R1CSInstance,
    SumcheckPhase1,
    SumcheckPhase2,
    ZMat,
    deref,
    dot,
    from_u512,
    hash_layer,
    host_eq_weight,
    host_mul,
    host_sum,
    sumcheck_phase1,
    vec_op,
    wit_block,
    wit_mem,
    wit_perm_w0,
    wit_shift,
    zmat_bind_rq,
    zmat_bind_weights,
)
