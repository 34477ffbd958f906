/*
 * spgpu.h -- C ABI of the B200-native prover backend for spartan-parallel's
 * data-parallel R1CS proving path.
 *
 * The reference crate (scroll-tech/spartan-parallel, pure Rust) has no FFI seam;
 * the seam is the set of method bodies listed below. A Rust `-sys` crate binds
 * exactly these symbols; the python/ctypes mirror in
 * spartan_parallel_b200/ binds the same ones.
 *
 * Conventions
 *  - spg_fq is the reference's `Scalar`: four little-endian u64 limbs holding
 *    a*2^256 mod q, fully reduced (src/scalar/ristretto255.rs:193-199). Every
 *    scalar crossing this boundary, in either direction, is in that form.
 *  - Every function returns 0 on success, a negative SPG_E* code otherwise;
 *    spg_last_error() gives the message. The reference's prover asserts/panics
 *    on bad shapes (e.g. src/r1csproof.rs:240-263); the Rust wrapper turns a
 *    non-zero status into panic!(). There is no CPU fallback anywhere.
 *  - Handles are opaque, own device memory, and are not thread-safe: one host
 *    thread drives one context (the reference prover is single-threaded).
 *  - Tables live on the device in NATURAL (p, q, w, x) order; the reference's
 *    bit-reversed (p, q_rev, w, x_rev) storage (src/custom_dense_mlpoly.rs:67-111)
 *    is an implementation detail of its top-binding loops and never crosses the ABI.
 */
#ifndef SPGPU_H
#define SPGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { uint64_t l[4]; } spg_fq;

typedef struct spg_ctx spg_ctx;       /* one device + stream + scratch */
typedef struct spg_vec spg_vec;       /* device vector of spg_fq (DensePolynomial.Z) */
typedef struct spg_r1cs spg_r1cs;     /* device copy of an R1CSInstance */
typedef struct spg_witness spg_witness; /* one witness section (ProverWitnessSecInfo) */
typedef struct spg_zmat spg_zmat;     /* z_mat[p][q][w][x] of R1CSProof::prove */
typedef struct spg_sc1 spg_sc1;       /* phase-1 sumcheck state */
typedef struct spg_sc2 spg_sc2;       /* phase-2 sumcheck state */
typedef struct spg_cubic spg_cubic;   /* prove_cubic_batched state */
typedef struct spg_gens spg_gens;     /* MultiCommitGens on the device */
typedef struct spg_bullet spg_bullet; /* BulletReductionProof::prove state (unfolded generators) */

enum {
  SPG_OK = 0,
  SPG_EINVAL = -1,  /* bad shape / argument (reference: assert!) */
  SPG_ECUDA = -2,   /* CUDA runtime error */
  SPG_ENOMEM = -3,
  SPG_ESTATE = -4   /* call out of order (e.g. round_bind before round_eval) */
};

const char *spg_last_error(void);
int spg_version(void);

/* ---------------------------------------------------------------- context */
int spg_ctx_create(int device, spg_ctx **out);
void spg_ctx_destroy(spg_ctx *ctx);
int spg_ctx_sync(spg_ctx *ctx);
/* number of kernel launches issued through this context so far */
uint64_t spg_ctx_launch_count(const spg_ctx *ctx);
/* Per-kernel timing with CUDA events recorded on the launching stream. Between begin
 * and end every launch is bracketed by an event pair; end writes a JSON array
 * [{"kernel", "launches", "total_ms", "units", "max_ms", "max_units"}] where "units" is
 * the algorithmic byte count the launcher declared for the kernel (0 if none). */
int spg_ctx_profile_begin(spg_ctx *ctx);
int spg_ctx_profile_end(spg_ctx *ctx, char *out_json, size_t cap);
/* cudaStream_t the context launches on (for CUDA-event timing by the caller) */
void *spg_ctx_stream(const spg_ctx *ctx);
/* pinned host allocation for fast uploads (optional convenience) */
int spg_host_alloc(size_t bytes, void **out);
void spg_host_free(void *p);

/* ---------------------------------------------------------------- vectors
 * DensePolynomial (src/dense_mlpoly.rs:19-24). */
int spg_vec_alloc(spg_ctx *ctx, size_t n, spg_vec **out);
/* v[i] = 0 */
int spg_vec_zero(spg_ctx *ctx, spg_vec *v);
int spg_vec_upload(spg_ctx *ctx, const spg_fq *host, size_t n, spg_vec **out);
/* wrap caller-owned device memory (e.g. a torch tensor); not freed by spg_vec_free */
int spg_vec_wrap(spg_ctx *ctx, void *device_ptr, size_t n, spg_vec **out);
int spg_vec_download(spg_ctx *ctx, const spg_vec *v, size_t offset, size_t n, spg_fq *host);
size_t spg_vec_len(const spg_vec *v);
void *spg_vec_device_ptr(const spg_vec *v);
void spg_vec_free(spg_vec *v);

/* ---------------------------------------------------------------- field (a1)
 * Scalar::{mul,add,sub,neg} elementwise, src/scalar/ristretto255.rs:690-763.
 * op: 0 mul, 1 add, 2 sub, 3 neg(a), 4 square(a), 5 to canonical integer (to_bytes),
 *     6 invert(a) (Scalar::invert, :541-595; 0 -> 0) */
int spg_fq_vec_op(spg_ctx *ctx, int op, const spg_vec *a, const spg_vec *b, spg_vec *out);
/* Scalar::from_u512 / from_bytes_wide on n wide values (8 u64 each), :435-466 */
int spg_fq_from_u512(spg_ctx *ctx, const uint64_t *host_wide, size_t n, spg_vec **out);

/* Host-side scalar helpers (no device involved) for the glue around sharded proving:
 * out[k] = sum_i in[i*width + k]  (Scalar::add) */
int spg_fq_host_sum(const spg_fq *in, size_t count, size_t width, spg_fq *out);
int spg_fq_host_mul(const spg_fq *a, const spg_fq *b, spg_fq *out);
/* prod_k eq(tau[k], bit_k(index)), k < nbits: the eq weight of a shard index when the
 * variable is bound low bit first */
int spg_fq_host_eq_weight(const spg_fq *tau, size_t nbits, uint64_t index, spg_fq *out);

/* ---------------------------------------------------------------- eq / dense MLE (a2, a3) */
/* EqPolynomial::evals, src/dense_mlpoly.rs:76-92: out has 2^ell entries, MSB <-> r[0] */
int spg_eq_evals(spg_ctx *ctx, const spg_fq *r, size_t ell, spg_vec **out);
/* DensePolynomial::bound_poly_var_top / _bot, :267-275 / :350-358 (in place; len halves) */
int spg_dense_bound_top(spg_ctx *ctx, spg_vec *v, const spg_fq *r);
int spg_dense_bound_bot(spg_ctx *ctx, spg_vec *v, const spg_fq *r);
/* DensePolynomial::evaluate, :361-367; r has log2(len) entries */
int spg_dense_evaluate(spg_ctx *ctx, const spg_vec *v, const spg_fq *r, size_t ell, spg_fq *out);
/* DensePolynomial::bound(L), :258-265: out[i] = sum_j L[j] * Z[j*R_size + i] */
int spg_dense_bound_L(spg_ctx *ctx, const spg_vec *v, const spg_fq *L, size_t L_size,
                      spg_vec **out);
/* DotProductProofLog::compute_dotproduct */
int spg_dot(spg_ctx *ctx, const spg_vec *a, const spg_vec *b, spg_fq *out);

/* ---------------------------------------------------------------- R1CS instance (a9, a10, a12)
 * R1CSInstance::new, src/r1csinstance.rs:89-182. Matrices are COO triples
 * (row, col, val) exactly as the reference stores them (src/sparse_mlpoly.rs:19-24);
 * mat index m = 3*inst + {0:A, 1:B, 2:C}; nnz[m] entries each, concatenated. */
int spg_r1cs_create(spg_ctx *ctx, size_t num_instances, size_t max_num_cons,
                    const size_t *num_cons, size_t num_vars, const size_t *nnz,
                    const uint32_t *rows, const uint32_t *cols, const spg_fq *vals,
                    spg_r1cs **out);
void spg_r1cs_destroy(spg_r1cs *inst);
/* SparseMatPolynomial::multi_evaluate via R1CSInstance::multi_evaluate,
 * src/r1csinstance.rs:583-595: out[3*num_instances] */
int spg_r1cs_multi_evaluate(spg_ctx *ctx, const spg_r1cs *inst, const spg_fq *rx, size_t nrx,
                            const spg_fq *ry, size_t nry, spg_fq *out);

/* ---------------------------------------------------------------- witnesses and z_mat (a11)
 * One ProverWitnessSecInfo (src/lib.rs): w_mat[p][q][i] flattened, poly_w[p] = w_mat[p].
 * num_instances is 1 (single) or P; num_proofs[p] is 1 (short) or Q_p. */
int spg_witness_upload(spg_ctx *ctx, size_t num_instances, const size_t *num_proofs,
                       const size_t *num_inputs, const spg_fq *host_w_mat, spg_witness **out);
/* same, but the copy runs on the context's copy stream and returns immediately: host_w_mat
 * must be pinned and stay valid until a consumer (spg_zmat_build, spg_witness_poly) has
 * been enqueued AND the context is next synchronised. Lets the upload of batch i+1 overlap
 * the proving of batch i. */
int spg_witness_upload_async(spg_ctx *ctx, size_t num_instances, const size_t *num_proofs,
                       const size_t *num_inputs, const spg_fq *host_w_mat, spg_witness **out);
void spg_witness_destroy(spg_witness *w);
/* poly_w[p] as a dense vector view (not owned by the caller) */
int spg_witness_poly(spg_witness *w, size_t p, spg_vec **out);
/* z_mat assembly, src/r1csproof.rs:278-293. The z_mat is a VIEW over the witness sections
 * (nothing is copied): the witness handles must outlive it. */
int spg_zmat_build(spg_ctx *ctx, size_t num_instances, const size_t *num_proofs,
                   const size_t *num_inputs, size_t num_witness_secs,
                   spg_witness *const *witness_secs, spg_zmat **out);
void spg_zmat_destroy(spg_zmat *z);

/* ---------------------------------------------------------------- phase-1 sumcheck (a4, a5, a9)
 * ZKSumcheckInstanceProof::prove_cubic_with_additive_term_disjoint_rounds,
 * src/sumcheck.rs:1067-1380, with comb = A*(B*C - D) (src/r1csproof.rs:100-104).
 * create = multiply_vec_block (src/r1csinstance.rs:363-436) + the three eq tables
 * (src/r1csproof.rs:305-322). num_cons is per proving instance (block_num_cons).
 * The matrix-vector products are validated here but computed inside the first round kernel
 * (one pass: read z, write Az/Bz/Cz, evaluate round 0), so `inst` and `z` must stay alive
 * until the first spg_sc1_round_eval (or spg_sc1_final / _debug_tables) has returned. */
int spg_sc1_create(spg_ctx *ctx, const spg_r1cs *inst, const spg_zmat *z, size_t num_instances,
                   const size_t *num_proofs, size_t max_num_proofs, const size_t *num_cons,
                   size_t max_num_cons, size_t max_num_inputs, const spg_fq *tau_p,
                   const spg_fq *tau_q, const spg_fq *tau_x, spg_sc1 **out);
/* same, from caller-provided Az/Bz/Cz in natural ragged [p][q][x] order (tests, partial offload) */
int spg_sc1_create_from_tables(spg_ctx *ctx, size_t num_instances, const size_t *num_proofs,
                               size_t max_num_proofs, const size_t *num_cons, size_t max_num_cons,
                               const spg_fq *Az, const spg_fq *Bz, const spg_fq *Cz,
                               const spg_fq *tau_p, const spg_fq *tau_q, const spg_fq *tau_x,
                               spg_sc1 **out);
/* Multiply every round evaluation and the eq claim by a constant. A shard of the proof
 * axis (multi-GPU) carries the eq factor of the proof bits it does not hold; the tail
 * rounds after a gather carry the eq products already bound on the shards. */
int spg_sc1_set_scale(spg_sc1 *s, const spg_fq *c);
/* The claim of the sumcheck, as the reference's prover receives it (`claim`,
 * src/sumcheck.rs:1069; R1CSProof::prove passes Scalar::zero(), src/r1csproof.rs:330). The
 * reference never evaluates the round polynomial at 1: it uses e(1) = claim - e(0)
 * (:1250-1256). With the claim supplied the device does the same from the first round on
 * (two evaluation points per pair instead of three). Results are identical to the
 * reference's whenever the claim really is the sum over the tables, i.e. for a satisfying
 * witness; without this call every round is exact for arbitrary tables. Call before round 0. */
int spg_sc1_set_claim(spg_sc1 *s, const spg_fq *claim);
/* the same, and the first round verifies the claim against the tables (three evaluation points
 * instead of two): SPG_EINVAL from spg_sc1_round_eval if the claim is not the true sum, e.g. for a
 * witness that does not satisfy the instance. Applies to tables the row-tiled kernels handle
 * (every row of at least 2^8 constraints); smaller tables evaluate three points anyway and ignore
 * the supplied claim's value. */
int spg_sc1_set_claim_checked(spg_sc1 *s, const spg_fq *claim);
/* spg_sc1_set_claim(0) plus the statement that the witness satisfies the instance row by row:
 * Az Bz - Cz is then zero at every (p, q, x), e(0) = e(1) = 0 in the first round are sums of zeros, and the
 * kernel that fuses the SpMV with that round evaluates the single point t = 2 (half its products). This is
 * what R1CSProof::prove knows about its inputs (src/r1csproof.rs:330 passes the zero claim for that reason).
 * Bit-identical to the reference for a satisfying witness; for any other input use _set_claim_checked. */
int spg_sc1_set_satisfied(spg_sc1 *s);
size_t spg_sc1_num_rounds(const spg_sc1 *s);
/* e = (eval_point_0, eval_point_2, eval_point_3) of the current round, :1166-1245 */
int spg_sc1_round_eval(spg_sc1 *s, spg_fq e[3]);
/* bind the round's variable to r_j, :1265-1275 */
int spg_sc1_round_bind(spg_sc1 *s, const spg_fq *r);
/* num_rounds x (round_eval, round_bind) with challenges known in advance (replaying a
 * transcript, benchmarking): evals_out[3 * j ..] = (e0, e2, e3) of round j. Every round still
 * returns its evaluations to the host before the bind is issued. */
int spg_sc1_run_rounds(spg_sc1 *s, size_t num_rounds, const spg_fq *challenges, spg_fq *evals_out);
/* the same loop for one shard of a proof whose proofs are spread over `world` processes of
 * one host: after each round_eval the 3 partial evaluations are exchanged through a POSIX
 * shared-memory mailbox (slot (b, r) at mailbox + (b * world + r) * slot_stride; word 0 =
 * sequence number, payload at +64; double-buffered by the parity of *calls) and summed with
 * Scalar::add; evals_out receives the sums. See spartan_parallel_b200/parallel.py. */
int spg_sc1_run_rounds_sharded(spg_sc1 *s, size_t num_rounds, const spg_fq *challenges,
                               spg_fq *evals_out, void *mailbox, size_t slot_stride, int rank,
                               int world, uint64_t *calls);
/* One exchange through the same mailbox: every rank contributes nbytes (<= slot_stride - 64) and
 * receives all ranks' contributions in rank order (out: world * nbytes). Release / acquire
 * ordering on the sequence words; waits are bounded (SPG_MAILBOX_TIMEOUT_S, default 120 s) and
 * a rank that fails poisons its slots (spg_mailbox_poison, done automatically on error) so
 * that its peers return SPG_ESTATE instead of spinning forever. */
int spg_mailbox_all_gather(void *mailbox, size_t slot_stride, int rank, int world, uint64_t *calls,
                           const void *data, size_t nbytes, void *out);
void spg_mailbox_poison(void *mailbox, size_t slot_stride, int rank, int world);
/* Sharded proofs whose rows (instance, proof) are spread over ranks in any way: a rank creates the
 * prover over ITS rows and replaces the row weights eq_p[p] * eq_q[q] of the x rounds by the global
 * eq weights of those rows (n_rows = sum_p num_proofs[p], table order); call before the first round.
 * The q and p rounds then run on the gathered per-row scalars (spg_sc1_debug_tables after the x
 * rounds -> spg_sc1_create_from_tables with num_cons = 1). */
int spg_sc1_set_row_weights(spg_sc1 *s, const spg_fq *weights, size_t n_rows);
/* The last log2(G) rounds of a proof sharded over G <= 64 ranks, on the host: after the rounds each
 * rank runs alone every table is one scalar per rank, and combining G scalars is mailbox-side work
 * like the per-round sum of partial evaluations (no device prover is stood up for them).
 * state = [E | A | B | C], G scalars each: E = eq table of the rank bits (index bit k <-> tau_q[nq_local
 * + k]), A, B, C = the gathered Az, Bz, Cz scalars in rank order. A round evaluates on the first
 * `len` entries of each (len = G >> round) and binds them in place to len / 2.
 * e = scale * sum E(t) (A(t) B(t) - C(t)) at t = 0, 2, 3. */
int spg_sc1_host_tail_eval(const spg_fq *state, size_t G, size_t len, const spg_fq *scale, spg_fq e[3]);
int spg_sc1_host_tail_bind(spg_fq *state, size_t G, size_t len, const spg_fq *r);
/* (tau_claim, Az, Bz, Cz) after the last bind, :1372-1377 */
int spg_sc1_final(spg_sc1 *s, spg_fq claims[4]);
/* copy the current Az/Bz/Cz tables back in natural ragged order (tests) */
int spg_sc1_debug_tables(spg_sc1 *s, spg_fq *Az, spg_fq *Bz, spg_fq *Cz, size_t cap, size_t *n);
void spg_sc1_destroy(spg_sc1 *s);

/* ---------------------------------------------------------------- phase-2 sumcheck (a6, a10)
 * ZKSumcheckInstanceProof::prove_cubic_disjoint_rounds, src/sumcheck.rs:788-1065,
 * comb = A*B*C. create = ABC table (src/r1csproof.rs:431-465), Z_poly bound to rq
 * (:469-479) and eq(rp) (:482). rx has log2(max_num_cons) entries (natural order,
 * i.e. already reversed as at :413), rq_rev log2(max_num_proofs), rp log2(P'). */
int spg_sc2_create(spg_ctx *ctx, const spg_r1cs *inst, const spg_zmat *z, size_t num_instances,
                   const size_t *num_proofs, size_t max_num_proofs, const size_t *num_inputs,
                   size_t max_num_inputs, size_t num_witness_secs, const spg_fq *rx,
                   const spg_fq *rq_rev, const spg_fq *rp, const spg_fq *r_A, const spg_fq *r_B,
                   const spg_fq *r_C, spg_sc2 **out);
/* The same, from a Z table already bound to rq (natural [p][w][y], sum_p W*num_inputs[p]
 * scalars). Used when the proof axis is sharded: every rank binds its shard with
 * spg_zmat_bind_rq, the partial tables are summed (Scalar::add) and phase 2 runs once. */
int spg_sc2_create_from_zrq(spg_ctx *ctx, const spg_r1cs *inst, const spg_vec *zrq, size_t num_instances,
                            const size_t *num_inputs, size_t max_num_inputs, size_t num_witness_secs,
                            const spg_fq *rx, const spg_fq *rp, const spg_fq *r_A, const spg_fq *r_B,
                            const spg_fq *r_C, spg_sc2 **out);
/* A y-sharded phase 2 (one instance, G ranks, G a multiple of the power-of-two number of witness
 * sections): rank r owns entries [flat_off, flat_off + flat_len) of the flat [w][y] tables (flat_len =
 * W * Y / G) and proves them as a (P = 1, W = 1, Y = flat_len) prover -- the summand has no weight over
 * (w, y), so the partial round evaluations of the ranks simply add. The ABC slice is built from the
 * instance, the Z slice is read from zrq at flat_off (spg_peer_reduce_scatter leaves it there).
 * spg_sc2_run_rounds_sharded runs the log2(flat_len) local rounds with the mailbox exchange; the remaining
 * log2(G) rounds run on the gathered scalars on the host: state = [B | C] (G each, flat (w, y_high) order),
 * mode 0 = a y round (adjacent pairs), mode 1 = a w round (top bit first). */
int spg_sc2_create_slice(spg_ctx *ctx, const spg_r1cs *inst, const spg_vec *zrq, size_t max_num_inputs,
                         size_t num_witness_secs, size_t flat_off, size_t flat_len, const spg_fq *rx,
                         const spg_fq *r_A, const spg_fq *r_B, const spg_fq *r_C, spg_sc2 **out);
int spg_sc2_run_rounds_sharded(spg_sc2 *s, size_t num_rounds, const spg_fq *challenges, spg_fq *evals_out,
                               void *mailbox, size_t slot_stride, int rank, int world, uint64_t *calls);
int spg_sc2_host_tail_eval(const spg_fq *state, size_t G, size_t len, int mode, const spg_fq *scale, spg_fq e[3]);
int spg_sc2_host_tail_bind(spg_fq *state, size_t G, size_t len, int mode, const spg_fq *r);
/* Z_poly.bound_poly_vars_rq (src/r1csproof.rs:478, src/custom_dense_mlpoly.rs:222-244, 300-304)
 * on its own: out[p][w][y] = scale * sum_q eq_lsb(rq_rev, q) z[p][q][w][y]; scale may be NULL (= 1). */
int spg_zmat_bind_rq(spg_ctx *ctx, const spg_zmat *z, const spg_fq *rq_rev, size_t nq, const spg_fq *scale,
                     spg_vec *out);
/* the same sum with explicit per-row weights (sum_p num_proofs[p], instance major) instead of the eq
 * table of rq: the sharded counterpart of spg_sc1_set_row_weights. out_off (may be NULL): where
 * instance p's W * num_inputs[p] scalars go inside `out` (a rank that owns some of the batch's
 * instances writes them at their place in the batch-wide table). */
int spg_zmat_bind_weights(spg_ctx *ctx, const spg_zmat *z, const spg_fq *weights, size_t n_weights,
                          const size_t *out_off, spg_vec *out);
size_t spg_sc2_num_rounds(const spg_sc2 *s);
int spg_sc2_round_eval(spg_sc2 *s, spg_fq e[3]);
int spg_sc2_round_bind(spg_sc2 *s, const spg_fq *r);
int spg_sc2_run_rounds(spg_sc2 *s, size_t num_rounds, const spg_fq *challenges, spg_fq *evals_out);
/* (eq claim, ABC claim, Z claim), :1058-1062 */
int spg_sc2_final(spg_sc2 *s, spg_fq claims[3]);
void spg_sc2_destroy(spg_sc2 *s);

/* ---------------------------------------------------------------- product trees (a13, a14, a7)
 * ProductCircuit::new, src/product_tree.rs:36-56: builds all layers of one circuit on
 * the device. layer k (0 = leaves) has left/right halves of length len/2^(k+1). */
typedef struct spg_prodtree spg_prodtree;
int spg_prodtree_build(spg_ctx *ctx, const spg_vec *leaves, spg_prodtree **out);
size_t spg_prodtree_num_layers(const spg_prodtree *t);
int spg_prodtree_layer(spg_prodtree *t, size_t layer, spg_vec **left, spg_vec **right);
/* ProductCircuit::evaluate, :58-63 */
int spg_prodtree_evaluate(spg_ctx *ctx, spg_prodtree *t, spg_fq *out);
void spg_prodtree_destroy(spg_prodtree *t);

/* SumcheckInstanceProof::prove_cubic_batched, src/sumcheck.rs:264-434, comb = A*B*C.
 * npar (A_i, B_i) pairs share C_par; nseq (A, B, C) triples. Tables are consumed
 * (bound in place). */
int spg_cubic_create(spg_ctx *ctx, size_t npar, spg_vec *const *A_par, spg_vec *const *B_par,
                     spg_vec *C_par, size_t nseq, spg_vec *const *A_seq, spg_vec *const *B_seq,
                     spg_vec *const *C_seq, const spg_fq *coeffs, spg_cubic **out);
int spg_cubic_round_eval(spg_cubic *s, spg_fq e[3]);
int spg_cubic_round_bind(spg_cubic *s, const spg_fq *r);
/* claims: A_par[npar], B_par[npar], C_par, A_seq[nseq], B_seq[nseq], C_seq[nseq] */
int spg_cubic_final(spg_cubic *s, spg_fq *claims);
void spg_cubic_destroy(spg_cubic *s);

/* ---------------------------------------------------------------- sparse-poly memory check (a15)
 * Layers::build_hash_layer, src/sparse_mlpoly.rs:612-687:
 *   out[i] = ts[i]*gamma^2 + val[i]*gamma + addr[i] - tau   (hash_func at :623-626)
 * addr/ts given as u64 integers (converted like DensePolynomial::from_usize), val as scalars. */
int spg_hash_layer(spg_ctx *ctx, const uint64_t *addr, const spg_vec *val, const uint64_t *ts,
                   size_t n, const spg_fq *gamma, const spg_fq *tau, int ts_plus_one,
                   spg_vec **out);
/* AddrTimestamps::deref_mem, :255-264: out[i] = mem[addr[i]] */
int spg_deref(spg_ctx *ctx, const uint64_t *addr, size_t n, const spg_vec *mem, spg_vec **out);

/* hash_func over scalar tables already on the device (the ops/timestamp polynomials of a
 * spg_sparse): addr == NULL means addr[i] = i, ts == NULL means ts = 0. */
int spg_hash_layer_fq(spg_ctx *ctx, const spg_vec *addr, const spg_vec *val, const spg_vec *ts,
                      int ts_plus_one, const spg_fq *gamma, const spg_fq *tau, spg_vec **out);

/* MultiSparseMatPolynomialAsDense, src/sparse_mlpoly.rs:273-280, 368-425: `batch` sparse
 * matrices (entries concatenated, nnz[i] each) padded to N = max next_pow2(nnz), their
 * address / read-timestamp / audit-timestamp polynomials and the merged comb_ops / comb_mem. */
typedef struct spg_sparse spg_sparse;
int spg_sparse_create(spg_ctx *ctx, size_t batch, size_t num_vars_x, size_t num_vars_y,
                      const size_t *nnz, const uint32_t *rows, const uint32_t *cols,
                      const spg_fq *vals, spg_sparse **out);
void spg_sparse_destroy(spg_sparse *s);
size_t spg_sparse_num_ops(const spg_sparse *s);       /* N */
size_t spg_sparse_num_mem_cells(const spg_sparse *s); /* 2^max(num_vars_x, num_vars_y) */
enum {
  SPG_SPARSE_ROW_ADDR = 0,
  SPG_SPARSE_ROW_READ_TS = 1,
  SPG_SPARSE_COL_ADDR = 2,
  SPG_SPARSE_COL_READ_TS = 3,
  SPG_SPARSE_VAL = 4,
  SPG_SPARSE_ROW_AUDIT_TS = 5,
  SPG_SPARSE_COL_AUDIT_TS = 6,
  SPG_SPARSE_COMB_OPS = 7,
  SPG_SPARSE_COMB_MEM = 8
};
/* non-owning view of one polynomial (i = matrix index for kinds 0-4); free with spg_vec_free,
 * the spg_sparse must outlive it */
int spg_sparse_view(spg_sparse *s, int kind, size_t i, spg_vec **out);
/* MultiSparseMatPolynomialAsDense::deref + Derefs::new, :34-62, 588-598: the merged lookup
 * table [row_ops_val[batch] | col_ops_val[batch] | zero pad], blocks of N */
int spg_sparse_deref(spg_ctx *ctx, const spg_sparse *s, const spg_vec *mem_rx,
                     const spg_vec *mem_ry, spg_vec **out);
/* device copy of v[offset, offset + n) */
int spg_vec_clone(spg_ctx *ctx, const spg_vec *v, size_t offset, size_t n, spg_vec **out);

/* ---------------------------------------------------------------- witness sections (f2)
 * The permutation-product recurrence behind every (pi, D) pair of the w3 witness sections
 * SNARK::prove builds before committing them (src/lib.rs:1378-1400 perm_exec_w3, :862-880
 * mem_gen, :1533-1570 block_w3 and its PHY / VIR pairs). Per segment (one proving
 * instance), from its last proof q down to its first:
 *     D[q]  = x[q] * (pi[q+1] + 1 - v[q+1])     (D[last] = x[last])
 *     pi[q] = v[q] * D[q]
 * The reference loops sequentially; here it is a parallel suffix scan of affine maps (exact
 * field arithmetic, so bit-identical). v, x, D, pi are strided views: entry q of a view is
 * vec[off + q * stride], which lets the caller point them at the columns of a row-major
 * w3 table (stride 8: v = col 0, x = col 1, pi = col 2, D = col 3, ...). seg_len[n_seg] sums
 * to n. */
int spg_perm_scan(spg_ctx *ctx, size_t n, const size_t *seg_len, size_t n_seg,
                  const spg_vec *v, size_t v_off, size_t v_stride,
                  const spg_vec *x, size_t x_off, size_t x_stride,
                  spg_vec *D, size_t D_off, size_t D_stride,
                  spg_vec *pi, size_t pi_off, size_t pi_stride);

/* ---------------------------------------------------------------- peer-memory all-reduce (e)
 * Modular sum of one table held at the same offset on `world` GPUs of a node (the rq-bound Z
 * table of a sharded proof, src/r1csproof.rs:478). Each process allocates the table with
 * spg_peer_alloc, passes the 64-byte IPC handle to its peers (any host channel), opens theirs
 * with spg_peer_open, and after every rank has finished writing its partial table (stream
 * sync + host barrier) calls spg_peer_sum: rank r sums chunk r over all peers with P2P loads
 * and writes the result into every peer's table with P2P stores. After a second stream sync +
 * host barrier every table holds the full sum. peer_ptrs[rank] is the rank's own table. */
int spg_peer_alloc(spg_ctx *ctx, size_t n, spg_vec **out, uint8_t handle[64]);
int spg_peer_free(spg_vec *v);
int spg_peer_open(spg_ctx *ctx, const uint8_t handle[64], void **ptr);
int spg_peer_close(void *ptr);
int spg_peer_sum(spg_ctx *ctx, void *const *peer_ptrs, int world, int rank, size_t n);
/* the same reduction, but the sum of chunk r stays with rank r only (half the NVLink traffic): for a
 * consumer that is sharded the same way (spg_sc2_create_slice) */
int spg_peer_reduce_scatter(spg_ctx *ctx, void *const *peer_ptrs, int world, int rank, size_t n);
/* the whole step between the phases of a proof sharded over the proof axis in one call: bind this rank's
 * Z rows to rq_rev[0 .. nq_local) scaled by the eq weight of its shard index under rq_rev[nq_local .. nq_total)
 * into peer_ptrs[rank], barrier (mailbox), spg_peer_sum or (scatter_only) spg_peer_reduce_scatter, barrier.
 * Replaces Z_poly.bound_poly_vars_rq (src/r1csproof.rs:478) of the unsharded prover. */
int spg_zmat_bind_rq_sharded(spg_ctx *ctx, const spg_zmat *z, const spg_fq *rq_rev, size_t nq_local, size_t nq_total,
                             void *const *peer_ptrs, int world, int rank, size_t n, int scatter_only, void *mailbox,
                             size_t slot_stride, uint64_t *calls);

/* ---------------------------------------------------------------- derived witness sections (f2)
 * What SNARK::prove computes from the primary sections (block_vars, exec_inputs, the memory
 * lists) and the two challenges (comb_tau, comb_r) right before committing: with these on the
 * device only the primary sections cross the host boundary. Tables are row-major, exactly the
 * flattened lists the reference commits (src/lib.rs:1424, 1443, 1630, 1655).
 *  spg_wit_perm_w0: perm_w0 = (tau, r, r^2, ..., r^(used-1), 0 ...), src/lib.rs:1328-1338
 *    (used = 2 * num_inputs_unpadded, total = num_ios).
 *  spg_wit_block: exec_mode = 0: block_w2 and block_w3 of ONE instance (src/lib.rs:1511-1613):
 *    vars = block_vars_mat[p] (rows x vars_width; inputs first, memory operations from io_width
 *    on), w2 rows of w2_width scalars, w3 rows of 8 = (v, x, pi, D, pi_phy, D_phy, pi_vir, D_vir).
 *    exec_mode = 1: perm_exec_w2 / perm_exec_w3 (src/lib.rs:1346-1400): vars = exec_inputs,
 *    w2_width = num_ios, w3 columns 4, 5 = w2[0], w2[1], no memory operations.
 *    seg_len: rows per proving instance (the (pi, D) recurrences restart at each).
 *  spg_wit_mem: mem_gen, src/lib.rs:832-880: mems rows (v, _, addr, data, ...).
 *  spg_wit_shift: w3_shifted, src/lib.rs:1667-1676: each instance's rows 1.. and a zero row. */
int spg_wit_perm_w0(spg_ctx *ctx, const spg_fq *tau, const spg_fq *r, size_t used, size_t total, spg_vec **out);
int spg_wit_block(spg_ctx *ctx, int exec_mode, const spg_vec *vars, size_t rows, size_t vars_width,
                  const spg_vec *perm_w0, const spg_fq *tau, const spg_fq *r, size_t num_inputs_unpadded,
                  size_t io_width, size_t phy_ops, size_t vir_ops, size_t w2_width, const size_t *seg_len,
                  size_t n_seg, spg_vec **w2_out, spg_vec **w3_out);
int spg_wit_mem(spg_ctx *ctx, const spg_vec *mems, size_t rows, size_t in_width, const spg_fq *tau,
                const spg_fq *r, size_t mem_width, spg_vec **w2_out, spg_vec **w3_out);
int spg_wit_shift(spg_ctx *ctx, const spg_vec *w3, size_t rows, size_t width, const size_t *seg_len,
                  size_t n_seg, spg_vec **out);

/* ---------------------------------------------------------------- commitments (a16)
 * MultiCommitGens::new is host-side setup (src/commitments.rs:15-33); the caller
 * passes the n+1 generators as compressed ristretto points (G[0..n], h). */
int spg_gens_upload(spg_ctx *ctx, const uint8_t *compressed, size_t n_plus_1, spg_gens **out);
/* MultiCommitGens::new on the device (src/commitments.rs:15-33): the caller supplies the
 * SHAKE256 output, 64 bytes per point (n generators, then h); each point is
 * RistrettoPoint::from_uniform_bytes of its 64 bytes. */
int spg_gens_from_uniform(spg_ctx *ctx, const uint8_t *uniform, size_t n_plus_1, spg_gens **out);
void spg_gens_destroy(spg_gens *g);
/* DensePolynomial::commit with zero blinds, src/dense_mlpoly.rs:199-239:
 * out = L_size compressed row commitments (32 bytes each). */
int spg_poly_commit(spg_ctx *ctx, const spg_gens *gens, const spg_vec *poly, size_t L_size,
                    uint8_t *out_compressed);
/* The same for rows [row0, row0 + nrows) only: the rows of a commitment are independent
 * (src/dense_mlpoly.rs:199-212 maps over them), which is how a commitment is sharded over
 * GPUs (each rank commits a slice, 32 bytes per row are gathered). out = nrows * 32 bytes. */
int spg_poly_commit_rows(spg_ctx *ctx, const spg_gens *gens, const spg_vec *poly, size_t L_size, size_t row0,
                         size_t nrows, uint8_t *out_compressed);
/* Builds the fixed-base window tables for the first R generators (and h) now instead of inside
 * the first commitment that needs them -- setup cost, like MultiCommitGens::new itself. */
int spg_gens_prepare(spg_ctx *ctx, spg_gens *gens, size_t R);
/* out = {window bits c, windows per scalar (= point additions per non-zero scalar), table
 * bytes, bases covered}; zeros before the first table is built. */
int spg_gens_info(const spg_gens *gens, size_t out[4]);
/* The same as spg_gens_prepare for a commitment of L rows over R bases (DensePolynomial::commit,
 * src/dense_mlpoly.rs:214-239): commitments with many rows use a second, single-window table per
 * base (the window factor 2^(c w) is applied once per row by a Horner chain instead of being folded
 * into the table: 15 additions per scalar at c = 17), which this call builds ahead as well. */
int spg_gens_prepare_rows(spg_ctx *ctx, spg_gens *gens, size_t L, size_t R);
/* spg_gens_info for that single-window table: {c, additions per non-zero scalar, bytes, bases};
 * zeros while commitments with many rows go through the per-window table. */
int spg_gens_info_rows(const spg_gens *gens, size_t out[4]);
/* Development aid: n pseudo-random and edge-case operand pairs through the eight-limb
 * GF(2^255-19) arithmetic of the commitment kernels, compared on the device with the ten-limb
 * code; *out_bad = OR of the failing checks' bits (0 = all agree). */
int spg_debug_fe8_selftest(spg_ctx *ctx, size_t n, uint64_t seed, uint32_t *out_bad);
/* F_q wide-range forms (csrc/fq.cuh: differences made non-negative by adding 2q / 6q, one fold into
 * [0, 2q)) as the row kernels use them, on n operand tuples that include every edge combination of
 * {0, 1, q-1, q, q+1, 2q-1}, against canonical arithmetic; *out_bad = OR of the failing checks (0 = pass) */
int spg_debug_fq_wide_selftest(spg_ctx *ctx, size_t n, uint64_t seed, uint32_t *out_bad);
/* Commitments::commit for a batch of short vectors sharing the bases (sumcheck
 * round polynomials etc.): out[i] = sum_j s[i*len+j] G[j] + blind[i] h */
int spg_commit_batch(spg_ctx *ctx, const spg_gens *gens, const spg_fq *scalars, size_t len,
                     const spg_fq *blinds, size_t count, uint8_t *out_compressed);

/* ---------------------------------------------------------------- bullet reduction (f1)
 * BulletReductionProof::prove, src/nizk/bullet.rs:72-119, without folding the generators: the
 * library keeps s[m] = prod_j (u_j or u_j^-1 by the top bits of m) on the device, and a round's
 * L and R are multiscalar multiplications over the ORIGINAL n bases (fixed-base tables, no
 * doublings). The host keeps the transcript and the O(nk) folds of a and b. Per round:
 *   spg_bullet_lr(nk, a[0..nk), {blind_L, blind_R}) -> sum_m scalar(m) G[m] + blind h for L and R
 *     (the caller adds c_L Q / c_R Q and compresses: src/nizk/bullet.rs:86-110);
 *   spg_bullet_fold(nk, u, u^-1) after the challenge (bullet.rs:113-118).
 * spg_bullet_final returns the fully folded generator G_hat (compressed). n: power of two. */
int spg_bullet_create(spg_ctx *ctx, const spg_gens *gens, size_t n, spg_bullet **out);
int spg_bullet_lr(spg_bullet *b, size_t nk, const spg_fq *a, const spg_fq blinds[2], uint8_t out_LR[64]);
int spg_bullet_fold(spg_bullet *b, size_t nk, const spg_fq *u, const spg_fq *u_inv);
int spg_bullet_final(spg_bullet *b, uint8_t out_G[32]);
/* The same rounds with a and b resident on the device (the O(nk) host loops of bullet.rs:83-84 and
 * :113-116 move with them): spg_bullet_set_ab uploads both vectors (n scalars each) once;
 * spg_bullet_lr_resident returns the round's L and R as above plus c_L = <a_L, b_R>, c_R = <a_R, b_L>;
 * spg_bullet_fold then folds s, a and b; spg_bullet_final_ab returns G_hat and the folded a[0], b[0].
 * ext = 0: out_LR = two ristretto encodings (64 bytes). ext = 1: the two points in extended coordinates,
 * X, Y, Z, T as canonical little-endian field elements (2 x 128 bytes): the caller adds c Q and encodes
 * the sum anyway, so the encoding here (an inversion and a square root on one thread) and the decoding
 * there are skipped. */
int spg_bullet_set_ab(spg_bullet *b, const spg_fq *a, const spg_fq *bvec);
int spg_bullet_lr_resident(spg_bullet *b, size_t nk, const spg_fq blinds[2], int ext, uint8_t *out_LR, spg_fq out_c[2]);
int spg_bullet_final_ab(spg_bullet *b, uint8_t out_G[32], spg_fq out_ab[2]);
void spg_bullet_destroy(spg_bullet *b);

#ifdef __cplusplus
}
#endif
#endif /* SPGPU_H */
