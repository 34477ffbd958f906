/*
 * ORACLE (test infrastructure, NOT product code).
 * CPU restatement of the reference's multilinear-polynomial tables:
 *   EqPolynomial / DensePolynomial   -> /root/reference/src/dense_mlpoly.rs
 *   DensePolynomialPqx               -> /root/reference/src/custom_dense_mlpoly.rs
 *   UniPoly                          -> /root/reference/src/unipoly.rs
 * Keeps the reference's storage order (p, q_rev, w, x_rev) and its loops, so
 * the CUDA path (which uses a different, natural-order layout) is checked
 * against an independent formulation.
 */
#ifndef SPG_ORACLE_POLYS_H
#define SPG_ORACLE_POLYS_H
#include "fq.h"

#ifdef __cplusplus
extern "C" {
#endif

#define OMODE_P 1
#define OMODE_Q 2
#define OMODE_W 3
#define OMODE_X 4

/* dense_mlpoly.rs:76-92; out has 2^ell entries */
void oeq_evals(const ofq *r, size_t ell, ofq *out);
/* dense_mlpoly.rs:69-74 */
ofq oeq_evaluate(const ofq *r, const ofq *rx, size_t ell);
/* dense_mlpoly.rs:267-275 (in place, new length n/2 is returned) */
size_t odense_bound_top(ofq *Z, size_t len, const ofq *r);
/* dense_mlpoly.rs:350-358 */
size_t odense_bound_bot(ofq *Z, size_t len, const ofq *r);
/* dense_mlpoly.rs:361-367 (chis built by oeq_evals, then dot product) */
ofq odense_evaluate(const ofq *Z, size_t len, const ofq *r, size_t ell);
/* dense_mlpoly.rs:258-265; out has R_size entries */
void odense_bound_L(const ofq *Z, size_t ell, const ofq *L, ofq *out);
/* nizk/mod.rs compute_dotproduct */
ofq odot(const ofq *a, const ofq *b, size_t n);

/* unipoly.rs:23-54: evals (3 or 4) -> coeffs, lowest degree first */
void ounipoly_from_evals(const ofq *evals, size_t n, ofq *coeffs);
/* unipoly.rs:72-80 */
ofq ounipoly_evaluate(const ofq *coeffs, size_t n, const ofq *r);

/* custom_dense_mlpoly.rs:36-41 */
size_t orev_bits(size_t q, size_t max_num_proofs);

typedef struct opqx {
  size_t P;                /* Z.len() */
  size_t W;                /* Z[p][q].len() */
  size_t *alloc_q;         /* Z[p].len() (never shrinks) */
  size_t *alloc_x;         /* Z[p][q][w].len() (never shrinks) */
  size_t *off;             /* offset of instance p in data */
  ofq *data;               /* [p][q][w][x] */
  /* the struct fields of DensePolynomialPqx (custom_dense_mlpoly.rs:22-33) */
  size_t num_instances;
  size_t *num_proofs;
  size_t max_num_proofs;
  size_t num_witness_secs;
  size_t *num_inputs;
  size_t max_num_inputs;
} opqx;

/* custom_dense_mlpoly.rs:67-111: z_nat is [p][q][w][x] flattened with the
 * ragged shape (num_proofs[p], W, num_inputs[p]) in natural q / x order. */
opqx *opqx_new_rev(const ofq *z_nat, size_t P, size_t W, const size_t *num_proofs,
                   size_t max_num_proofs, const size_t *num_inputs, size_t max_num_inputs);
/* custom_dense_mlpoly.rs:45-63: same shape, stored as given (no reversal) */
opqx *opqx_new(const ofq *z, size_t P, size_t W, const size_t *num_proofs,
               size_t max_num_proofs, const size_t *num_inputs, size_t max_num_inputs);
opqx *opqx_clone(const opqx *s);
void opqx_free(opqx *s);
size_t opqx_total(const opqx *s);            /* number of stored scalars */
void opqx_copy_out(const opqx *s, ofq *out); /* raw storage, stored order */
ofq opqx_index(const opqx *s, size_t p, size_t q, size_t w, size_t x);                /* :118-128 */
ofq opqx_index_high(const opqx *s, size_t p, size_t q, size_t w, size_t x, int mode); /* :136-173 */
void opqx_bound_poly(opqx *s, const ofq *r, int mode);                                /* :180-289 */
size_t opqx_len(const opqx *s);                                                       /* :113-115 */
/* :320-333 */
ofq opqx_evaluate(const opqx *s, const ofq *rp, size_t np, const ofq *rq, size_t nq,
                  const ofq *rw, size_t nw, const ofq *rx, size_t nx);

#ifdef __cplusplus
}
#endif
#endif
