/* ORACLE (test infrastructure, NOT product code) -- see sumcheck.h. */
#include "sumcheck.h"
#include <assert.h>
#include <stdlib.h>
#include <string.h>

static size_t min_sz(size_t a, size_t b) { return a < b ? a : b; }
static size_t *dup_sz(const size_t *a, size_t n) {
  size_t *r = (size_t *)malloc(sizeof(size_t) * (n ? n : 1));
  memcpy(r, a, sizeof(size_t) * n);
  return r;
}
static ofq *dup_fq(const ofq *a, size_t n) {
  ofq *r = (ofq *)malloc(sizeof(ofq) * (n ? n : 1));
  memcpy(r, a, sizeof(ofq) * n);
  return r;
}

/* value at 2 and at 3 of the line through (0, lo), (1, hi):
 * sumcheck.rs:1210-1236 computes hi+hi-lo, then +hi-lo again. */
static void extrap(const ofq *lo, const ofq *hi, ofq *at2, ofq *at3) {
  ofq t = ofq_add(hi, hi);
  *at2 = ofq_sub(&t, lo);
  t = ofq_add(at2, hi);
  *at3 = ofq_sub(&t, lo);
}

/* ------------------------------------------------------------------ phase 1 */

osc1 *osc1_new(size_t nx, size_t nq, size_t np, size_t P, const size_t *num_proofs,
               const size_t *num_cons, const ofq *Ap, const ofq *Aq, const ofq *Ax, opqx *B,
               opqx *C, opqx *D) {
  osc1 *s = (osc1 *)calloc(1, sizeof(osc1));
  s->num_rounds_x = nx; s->num_rounds_q = nq; s->num_rounds_p = np;
  s->cons_len = (size_t)1 << nx; s->proof_len = (size_t)1 << nq; s->instance_len = (size_t)1 << np;
  s->P = P;
  s->num_proofs = dup_sz(num_proofs, P);
  s->num_cons = dup_sz(num_cons, P);
  s->lenAp = s->instance_len; s->lenAq = s->proof_len; s->lenAx = s->cons_len;
  s->Ap = dup_fq(Ap, s->lenAp); s->Aq = dup_fq(Aq, s->lenAq); s->Ax = dup_fq(Ax, s->lenAx);
  s->B = B; s->C = C; s->D = D;
  return s;
}

void osc1_free(osc1 *s) {
  if (!s) return;
  free(s->num_proofs); free(s->num_cons); free(s->Ap); free(s->Aq); free(s->Ax);
  opqx_free(s->B); opqx_free(s->C); opqx_free(s->D);
  free(s);
}

/* comb_func of r1csproof.rs:100-104: A * (B*C - D) */
static ofq comb1(const ofq *a, const ofq *b, const ofq *c, const ofq *d) {
  ofq bc = ofq_mul(b, c);
  ofq t = ofq_sub(&bc, d);
  return ofq_mul(a, &t);
}

void osc1_round_eval(osc1 *s, ofq out[3]) {
  size_t j = s->round;
  int mode = j < s->num_rounds_x ? OMODE_X : (j < s->num_rounds_x + s->num_rounds_q ? OMODE_Q : OMODE_P);
  s->mode = mode;
  if (s->cons_len > 1) s->cons_len /= 2;
  else if (s->proof_len > 1) s->proof_len /= 2;
  else s->instance_len /= 2;

  ofq e0 = ofq_zero(), e2 = ofq_zero(), e3 = ofq_zero();
  for (size_t p = 0; p < min_sz(s->instance_len, s->P); p++) {
    if (mode == OMODE_X && s->num_cons[p] > 1) s->num_cons[p] /= 2;
    if (mode == OMODE_Q && s->num_proofs[p] > 1) s->num_proofs[p] /= 2;
    /* the (q, x) iterations are independent; field sums are exact, so per-thread partial
     * sums combined at the end equal the reference's sequential accumulation bit for bit */
    size_t nq = s->num_proofs[p], nxp = s->num_cons[p];
#pragma omp parallel if (nq * nxp >= 8192)
    {
      ofq t0 = ofq_zero(), t2 = ofq_zero(), t3 = ofq_zero();
#pragma omp for collapse(2) schedule(static)
      for (size_t q = 0; q < nq; q++) {
        for (size_t x = 0; x < nxp; x++) {
          size_t step_q = s->proof_len / nq;
          size_t step_x = s->cons_len / nxp;
          ofq pq = ofq_mul(&s->Ap[p], &s->Aq[q * step_q]);
          ofq A_lo = ofq_mul(&pq, &s->Ax[x * step_x]);
          ofq A_hi;
          if (mode == OMODE_P) {
            ofq t = ofq_mul(&s->Ap[p + s->instance_len], &s->Aq[q * step_q]);
            A_hi = ofq_mul(&t, &s->Ax[x * step_x]);
          } else if (mode == OMODE_Q) {
            ofq t = ofq_mul(&s->Ap[p], &s->Aq[q * step_q + s->proof_len]);
            A_hi = ofq_mul(&t, &s->Ax[x * step_x]);
          } else {
            A_hi = ofq_mul(&pq, &s->Ax[x * step_x + s->cons_len]);
          }
          ofq B_lo = opqx_index(s->B, p, q, 0, x), B_hi = opqx_index_high(s->B, p, q, 0, x, mode);
          ofq C_lo = opqx_index(s->C, p, q, 0, x), C_hi = opqx_index_high(s->C, p, q, 0, x, mode);
          ofq D_lo = opqx_index(s->D, p, q, 0, x), D_hi = opqx_index_high(s->D, p, q, 0, x, mode);
          ofq t = comb1(&A_lo, &B_lo, &C_lo, &D_lo);
          t0 = ofq_add(&t0, &t);
          ofq A2, A3, B2, B3, C2, C3, D2, D3;
          extrap(&A_lo, &A_hi, &A2, &A3);
          extrap(&B_lo, &B_hi, &B2, &B3);
          extrap(&C_lo, &C_hi, &C2, &C3);
          extrap(&D_lo, &D_hi, &D2, &D3);
          t = comb1(&A2, &B2, &C2, &D2);
          t2 = ofq_add(&t2, &t);
          t = comb1(&A3, &B3, &C3, &D3);
          t3 = ofq_add(&t3, &t);
        }
      }
#pragma omp critical
      {
        e0 = ofq_add(&e0, &t0);
        e2 = ofq_add(&e2, &t2);
        e3 = ofq_add(&e3, &t3);
      }
    }
  }
  out[0] = e0; out[1] = e2; out[2] = e3;
}

void osc1_round_bind(osc1 *s, const ofq *r) {
  int mode = s->mode;
  if (mode == OMODE_P) s->lenAp = odense_bound_top(s->Ap, s->lenAp, r);
  else if (mode == OMODE_Q) s->lenAq = odense_bound_top(s->Aq, s->lenAq, r);
  else s->lenAx = odense_bound_top(s->Ax, s->lenAx, r);
  opqx_bound_poly(s->B, r, mode);
  opqx_bound_poly(s->C, r, mode);
  opqx_bound_poly(s->D, r, mode);
  s->round++;
}

void osc1_final(const osc1 *s, ofq out[4]) {
  ofq t = ofq_mul(&s->Ap[0], &s->Aq[0]);
  out[0] = ofq_mul(&t, &s->Ax[0]);
  out[1] = opqx_index(s->B, 0, 0, 0, 0);
  out[2] = opqx_index(s->C, 0, 0, 0, 0);
  out[3] = opqx_index(s->D, 0, 0, 0, 0);
}

/* ------------------------------------------------------------------ phase 2 */

osc2 *osc2_new(size_t ny, size_t nw, size_t np, int single_inst, size_t num_witness_secs, size_t P,
               const size_t *num_inputs, const ofq *A, opqx *B, opqx *C) {
  osc2 *s = (osc2 *)calloc(1, sizeof(osc2));
  s->num_rounds_y = ny; s->num_rounds_w = nw; s->num_rounds_p = np;
  s->inputs_len = (size_t)1 << ny; s->witness_secs_len = (size_t)1 << nw; s->instance_len = (size_t)1 << np;
  s->single_inst = single_inst;
  s->num_witness_secs = num_witness_secs;
  s->P = P;
  s->num_inputs = dup_sz(num_inputs, P);
  s->lenA = s->instance_len;
  s->A = dup_fq(A, s->lenA);
  s->B = B; s->C = C;
  return s;
}

void osc2_free(osc2 *s) {
  if (!s) return;
  free(s->num_inputs); free(s->A);
  opqx_free(s->B); opqx_free(s->C);
  free(s);
}

static ofq comb2(const ofq *a, const ofq *b, const ofq *c) {
  ofq ab = ofq_mul(a, b);
  return ofq_mul(&ab, c);
}

void osc2_round_eval(osc2 *s, ofq out[3]) {
  size_t j = s->round;
  int mode = j < s->num_rounds_y ? OMODE_X : (j < s->num_rounds_y + s->num_rounds_w ? OMODE_W : OMODE_P);
  s->mode = mode;
  if (s->inputs_len > 1) s->inputs_len /= 2;
  else if (s->witness_secs_len > 1) s->witness_secs_len /= 2;
  else s->instance_len /= 2;

  ofq e0 = ofq_zero(), e2 = ofq_zero(), e3 = ofq_zero();
  for (size_t p = 0; p < min_sz(s->instance_len, s->P); p++) {
    size_t p_inst = s->single_inst ? 0 : p;
    if (mode == OMODE_X && s->num_inputs[p] > 1) s->num_inputs[p] /= 2;
    for (size_t w = 0; w < min_sz(s->witness_secs_len, s->num_witness_secs); w++) {
      for (size_t y = 0; y < s->num_inputs[p]; y++) {
        ofq A_lo = s->A[p];
        ofq A_hi = (mode == OMODE_P) ? s->A[p + s->instance_len] : s->A[p];
        ofq B_lo = opqx_index(s->B, p_inst, 0, w, y), B_hi = opqx_index_high(s->B, p_inst, 0, w, y, mode);
        ofq C_lo = opqx_index(s->C, p, 0, w, y), C_hi = opqx_index_high(s->C, p, 0, w, y, mode);
        ofq t = comb2(&A_lo, &B_lo, &C_lo);
        e0 = ofq_add(&e0, &t);
        ofq A2, A3, B2, B3, C2, C3;
        extrap(&A_lo, &A_hi, &A2, &A3);
        extrap(&B_lo, &B_hi, &B2, &B3);
        extrap(&C_lo, &C_hi, &C2, &C3);
        t = comb2(&A2, &B2, &C2);
        e2 = ofq_add(&e2, &t);
        t = comb2(&A3, &B3, &C3);
        e3 = ofq_add(&e3, &t);
      }
    }
  }
  out[0] = e0; out[1] = e2; out[2] = e3;
}

void osc2_round_bind(osc2 *s, const ofq *r) {
  int mode = s->mode;
  if (mode == OMODE_P) s->lenA = odense_bound_top(s->A, s->lenA, r);
  if (mode != OMODE_P || !s->single_inst) opqx_bound_poly(s->B, r, mode);
  opqx_bound_poly(s->C, r, mode);
  s->round++;
}

void osc2_final(const osc2 *s, ofq out[3]) {
  out[0] = s->A[0];
  out[1] = opqx_index(s->B, 0, 0, 0, 0);
  out[2] = opqx_index(s->C, 0, 0, 0, 0);
}

/* ------------------------------------------------------------------ batched cubic */

static void cubic_triple(const ofq *A, const ofq *B, const ofq *C, size_t len, ofq e[3]) {
  size_t h = len / 2;
  e[0] = e[1] = e[2] = ofq_zero();
  for (size_t i = 0; i < h; i++) {
    ofq t = comb2(&A[i], &B[i], &C[i]);
    e[0] = ofq_add(&e[0], &t);
    ofq A2, A3, B2, B3, C2, C3;
    extrap(&A[i], &A[h + i], &A2, &A3);
    extrap(&B[i], &B[h + i], &B2, &B3);
    extrap(&C[i], &C[h + i], &C2, &C3);
    t = comb2(&A2, &B2, &C2);
    e[1] = ofq_add(&e[1], &t);
    t = comb2(&A3, &B3, &C3);
    e[2] = ofq_add(&e[2], &t);
  }
}

void ocubic_batched_eval(size_t len, size_t npar, ofq *const *A_par, ofq *const *B_par,
                         const ofq *C_par, size_t nseq, ofq *const *A_seq, ofq *const *B_seq,
                         ofq *const *C_seq, const ofq *coeffs, ofq out[3]) {
  out[0] = out[1] = out[2] = ofq_zero();
  for (size_t k = 0; k < npar + nseq; k++) {
    ofq e[3];
    if (k < npar) cubic_triple(A_par[k], B_par[k], C_par, len, e);
    else cubic_triple(A_seq[k - npar], B_seq[k - npar], C_seq[k - npar], len, e);
    for (int t = 0; t < 3; t++) {
      ofq m = ofq_mul(&e[t], &coeffs[k]);
      out[t] = ofq_add(&out[t], &m);
    }
  }
}

/* ------------------------------------------------------------------ sparse */

void ospmv(size_t nnz, const uint32_t *row, const uint32_t *col, const ofq *val, size_t num_rows,
           size_t max_num_cols, const ofq *z, size_t seg_stride, ofq *out) {
  for (size_t i = 0; i < num_rows; i++) out[i] = ofq_zero();
  for (size_t i = 0; i < nnz; i++) {
    const ofq *zz = &z[(col[i] / max_num_cols) * seg_stride + (col[i] % max_num_cols)];
    ofq t = ofq_mul(&val[i], zz);
    out[row[i]] = ofq_add(&out[row[i]], &t);
  }
}

/* the same product for `nq` proofs of one instance (z and out advance by z_stride / num_rows
 * scalars per proof): the proofs are independent, which is the reference's data-parallel axis */
void ospmv_batch(size_t nnz, const uint32_t *row, const uint32_t *col, const ofq *val, size_t num_rows,
                 size_t max_num_cols, const ofq *z, size_t seg_stride, size_t nq, size_t z_stride, ofq *out) {
#pragma omp parallel for schedule(static) if (nq > 1 && nq * nnz >= 8192)
  for (size_t q = 0; q < nq; q++)
    ospmv(nnz, row, col, val, num_rows, max_num_cols, z + q * z_stride, seg_stride, out + q * num_rows);
}

void oeval_table_sparse(size_t nnz, const uint32_t *row, const uint32_t *col, const ofq *val,
                        const ofq *rx, size_t num_segs, size_t max_num_cols, size_t num_cols,
                        ofq *out) {
  for (size_t i = 0; i < num_segs * num_cols; i++) out[i] = ofq_zero();
  for (size_t i = 0; i < nnz; i++) {
    size_t seg = col[i] / max_num_cols, c = col[i] % max_num_cols;
    assert(seg < num_segs && c < num_cols);
    ofq t = ofq_mul(&rx[row[i]], &val[i]);
    out[seg * num_cols + c] = ofq_add(&out[seg * num_cols + c], &t);
  }
}

ofq osparse_evaluate_with_tables(size_t nnz, const uint32_t *row, const uint32_t *col,
                                 const ofq *val, const ofq *trx, const ofq *try_) {
  ofq acc = ofq_zero();
  for (size_t i = 0; i < nnz; i++) {
    ofq t = ofq_mul(&trx[row[i]], &try_[col[i]]);
    t = ofq_mul(&t, &val[i]);
    acc = ofq_add(&acc, &t);
  }
  return acc;
}

/* ------------------------------------------------------------------ product tree */

void oprod_layer(const ofq *left, const ofq *right, size_t n, ofq *out_left, ofq *out_right) {
  size_t h = n / 2;
  for (size_t i = 0; i < h; i++) out_left[i] = ofq_mul(&left[i], &right[i]);
  for (size_t i = h; i < n; i++) out_right[i - h] = ofq_mul(&left[i], &right[i]);
}
