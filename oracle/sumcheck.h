/*
 * ORACLE (test infrastructure, NOT product code).
 * CPU restatement of the round loops of /root/reference/src/sumcheck.rs, split
 * at the host<->device seam the product uses: round_eval (the triple loop that
 * yields e(0), e(2), e(3)) and round_bind (the bound_poly calls). The
 * Fiat-Shamir / ZK glue around them is restated separately in python
 * (oracle/protocol.py).
 */
#ifndef SPG_ORACLE_SUMCHECK_H
#define SPG_ORACLE_SUMCHECK_H
#include "polys.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- phase 1: prove_cubic_with_additive_term_disjoint_rounds, sumcheck.rs:1067-1380 */
typedef struct osc1 {
  size_t num_rounds_x, num_rounds_q, num_rounds_p, round;
  size_t cons_len, proof_len, instance_len;
  size_t P;
  size_t *num_proofs, *num_cons;
  ofq *Ap, *Aq, *Ax;
  size_t lenAp, lenAq, lenAx;
  opqx *B, *C, *D; /* owned */
  int mode;        /* mode of the round last evaluated */
} osc1;

/* Ap/Aq/Ax: eq tables (copied). B,C,D: ownership is taken. */
osc1 *osc1_new(size_t num_rounds_x, size_t num_rounds_q, size_t num_rounds_p, size_t P,
               const size_t *num_proofs, const size_t *num_cons, const ofq *Ap, const ofq *Aq,
               const ofq *Ax, opqx *B, opqx *C, opqx *D);
void osc1_free(osc1 *s);
void osc1_round_eval(osc1 *s, ofq out[3]);   /* sumcheck.rs:1150-1245 */
void osc1_round_bind(osc1 *s, const ofq *r); /* sumcheck.rs:1265-1275 */
void osc1_final(const osc1 *s, ofq out[4]);  /* sumcheck.rs:1372-1377 */

/* ---- phase 2: prove_cubic_disjoint_rounds, sumcheck.rs:788-1065 */
typedef struct osc2 {
  size_t num_rounds_y, num_rounds_w, num_rounds_p, round;
  size_t inputs_len, witness_secs_len, instance_len;
  int single_inst;
  size_t num_witness_secs;
  size_t P;
  size_t *num_inputs;
  ofq *A;
  size_t lenA;
  opqx *B, *C; /* owned */
  int mode;
} osc2;

osc2 *osc2_new(size_t num_rounds_y, size_t num_rounds_w, size_t num_rounds_p, int single_inst,
               size_t num_witness_secs, size_t P, const size_t *num_inputs, const ofq *A,
               opqx *B, opqx *C);
void osc2_free(osc2 *s);
void osc2_round_eval(osc2 *s, ofq out[3]);   /* sumcheck.rs:858-941 */
void osc2_round_bind(osc2 *s, const ofq *r); /* sumcheck.rs:961-968 */
void osc2_final(const osc2 *s, ofq out[3]);  /* sumcheck.rs:1058-1062 */

/* ---- prove_cubic_batched round, sumcheck.rs:297-371. Tables are arrays of
 * pointers to caller-owned dense vectors, all of length len (current).
 * comb = A*B*C. out = (e0,e2,e3) already combined with coeffs. */
void ocubic_batched_eval(size_t len, size_t npar, ofq *const *A_par, ofq *const *B_par,
                         const ofq *C_par, size_t nseq, ofq *const *A_seq, ofq *const *B_seq,
                         ofq *const *C_seq, const ofq *coeffs, ofq out[3]);

/* ---- sparse / R1CS glue */
/* sparse_mlpoly.rs:454-472 : out[num_rows] = M * z, z given as segments
 * z[col / max_num_cols][col % max_num_cols] with segment stride seg_stride */
void ospmv(size_t nnz, const uint32_t *row, const uint32_t *col, const ofq *val, size_t num_rows,
           size_t max_num_cols, const ofq *z, size_t seg_stride, ofq *out);
void ospmv_batch(size_t nnz, const uint32_t *row, const uint32_t *col, const ofq *val, size_t num_rows,
                 size_t max_num_cols, const ofq *z, size_t seg_stride, size_t nq, size_t z_stride, ofq *out);
/* sparse_mlpoly.rs:524-541 : out[num_segs][num_cols] (zeroed here) */
void oeval_table_sparse(size_t nnz, const uint32_t *row, const uint32_t *col, const ofq *val,
                        const ofq *rx, size_t num_segs, size_t max_num_cols, size_t num_cols,
                        ofq *out);
/* sparse_mlpoly.rs:427-436 */
ofq osparse_evaluate_with_tables(size_t nnz, const uint32_t *row, const uint32_t *col,
                                 const ofq *val, const ofq *trx, const ofq *try_);

/* ---- product_tree.rs:18-34: one layer; in: left,right of length n; out: left',right' of n/2 */
void oprod_layer(const ofq *left, const ofq *right, size_t n, ofq *out_left, ofq *out_right);

#ifdef __cplusplus
}
#endif
#endif
