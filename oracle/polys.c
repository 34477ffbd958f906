/* ORACLE (test infrastructure, NOT product code) -- see polys.h. */
#include "polys.h"
#include <assert.h>
#include <stdlib.h>
#include <string.h>

static size_t min_sz(size_t a, size_t b) { return a < b ? a : b; }
static size_t next_pow2(size_t n) { size_t p = 1; while (p < n) p <<= 1; return p; }
static size_t log2_sz(size_t n) { size_t l = 0; while (((size_t)1 << l) < n) l++; return l; }

/* ------------------------------------------------------------------ Eq / dense */

void oeq_evals(const ofq *r, size_t ell, ofq *out) {
  size_t n = (size_t)1 << ell;
  for (size_t i = 0; i < n; i++) out[i] = ofq_one();
  size_t size = 1;
  for (size_t j = 0; j < ell; j++) {
    size *= 2;
    /* walk the odd indices from the top: each parent is copied to two children */
    for (size_t i = size - 1;; i -= 2) {
      ofq s = out[i / 2];
      out[i] = ofq_mul(&s, &r[j]);
      out[i - 1] = ofq_sub(&s, &out[i]);
      if (i == 1) break;
    }
  }
}

ofq oeq_evaluate(const ofq *r, const ofq *rx, size_t ell) {
  ofq acc = ofq_one();
  ofq one = ofq_one();
  for (size_t i = 0; i < ell; i++) {
    ofq a = ofq_mul(&r[i], &rx[i]);
    ofq b = ofq_sub(&one, &r[i]);
    ofq c = ofq_sub(&one, &rx[i]);
    ofq d = ofq_mul(&b, &c);
    ofq t = ofq_add(&a, &d);
    acc = ofq_mul(&acc, &t);
  }
  return acc;
}

size_t odense_bound_top(ofq *Z, size_t len, const ofq *r) {
  size_t n = len / 2;
  for (size_t i = 0; i < n; i++) {
    ofq d = ofq_sub(&Z[i + n], &Z[i]);
    ofq t = ofq_mul(r, &d);
    Z[i] = ofq_add(&Z[i], &t);
  }
  return n;
}

size_t odense_bound_bot(ofq *Z, size_t len, const ofq *r) {
  size_t n = len / 2;
  for (size_t i = 0; i < n; i++) {
    ofq d = ofq_sub(&Z[2 * i + 1], &Z[2 * i]);
    ofq t = ofq_mul(r, &d);
    Z[i] = ofq_add(&Z[2 * i], &t);
  }
  return n;
}

ofq odot(const ofq *a, const ofq *b, size_t n) {
  ofq acc = ofq_zero();
  for (size_t i = 0; i < n; i++) {
    ofq t = ofq_mul(&a[i], &b[i]);
    acc = ofq_add(&acc, &t);
  }
  return acc;
}

ofq odense_evaluate(const ofq *Z, size_t len, const ofq *r, size_t ell) {
  assert(len == ((size_t)1 << ell));
  ofq *chis = (ofq *)malloc(sizeof(ofq) * len);
  oeq_evals(r, ell, chis);
  ofq res = odot(Z, chis, len);
  free(chis);
  return res;
}

void odense_bound_L(const ofq *Z, size_t ell, const ofq *L, ofq *out) {
  size_t left = ell / 2, right = ell - left;
  size_t Ls = (size_t)1 << left, Rs = (size_t)1 << right;
  for (size_t i = 0; i < Rs; i++) {
    ofq acc = ofq_zero();
    for (size_t j = 0; j < Ls; j++) {
      ofq t = ofq_mul(&L[j], &Z[j * Rs + i]);
      acc = ofq_add(&acc, &t);
    }
    out[i] = acc;
  }
}

/* ------------------------------------------------------------------ UniPoly */

static ofq small_inv(unsigned v) {
  /* (v_usize).to_scalar().invert(): scalar/mod.rs:10-15 sums v ones */
  ofq acc = ofq_zero(), one = ofq_one();
  for (unsigned i = 0; i < v; i++) acc = ofq_add(&acc, &one);
  return ofq_invert(&acc);
}

void ounipoly_from_evals(const ofq *e, size_t n, ofq *co) {
  assert(n == 3 || n == 4);
  ofq two_inv = small_inv(2);
  if (n == 3) {
    ofq c = e[0];
    ofq t = ofq_sub(&e[2], &e[1]);
    t = ofq_sub(&t, &e[1]);
    t = ofq_add(&t, &c);
    ofq a = ofq_mul(&two_inv, &t);
    ofq b = ofq_sub(&e[1], &c);
    b = ofq_sub(&b, &a);
    co[0] = c; co[1] = b; co[2] = a;
  } else {
    ofq six_inv = small_inv(6);
    ofq d = e[0];
    ofq t = ofq_sub(&e[3], &e[2]);
    t = ofq_sub(&t, &e[2]);
    t = ofq_sub(&t, &e[2]);
    t = ofq_add(&t, &e[1]);
    t = ofq_add(&t, &e[1]);
    t = ofq_add(&t, &e[1]);
    t = ofq_sub(&t, &e[0]);
    ofq a = ofq_mul(&six_inv, &t);
    ofq u = ofq_add(&e[0], &e[0]);
    for (int k = 0; k < 5; k++) u = ofq_sub(&u, &e[1]);
    for (int k = 0; k < 4; k++) u = ofq_add(&u, &e[2]);
    u = ofq_sub(&u, &e[3]);
    ofq b = ofq_mul(&two_inv, &u);
    ofq c = ofq_sub(&e[1], &d);
    c = ofq_sub(&c, &a);
    c = ofq_sub(&c, &b);
    co[0] = d; co[1] = c; co[2] = b; co[3] = a;
  }
}

ofq ounipoly_evaluate(const ofq *co, size_t n, const ofq *r) {
  ofq eval = co[0];
  ofq power = *r;
  for (size_t i = 1; i < n; i++) {
    ofq t = ofq_mul(&power, &co[i]);
    eval = ofq_add(&eval, &t);
    power = ofq_mul(&power, r);
  }
  return eval;
}

/* ------------------------------------------------------------------ Pqx */

size_t orev_bits(size_t q, size_t maxn) {
  size_t lg = log2_sz(maxn), out = 0;
  for (size_t i = 0; i < lg; i++)
    if ((q >> i) & 1) out += maxn >> (i + 1);
  return out;
}

static opqx *opqx_alloc(size_t P, size_t W, const size_t *num_proofs, size_t max_num_proofs,
                        const size_t *num_inputs, size_t max_num_inputs) {
  opqx *s = (opqx *)calloc(1, sizeof(opqx));
  s->P = P;
  s->W = W;
  s->alloc_q = (size_t *)malloc(sizeof(size_t) * P);
  s->alloc_x = (size_t *)malloc(sizeof(size_t) * P);
  s->num_proofs = (size_t *)malloc(sizeof(size_t) * P);
  s->num_inputs = (size_t *)malloc(sizeof(size_t) * P);
  s->off = (size_t *)malloc(sizeof(size_t) * (P + 1));
  size_t tot = 0;
  for (size_t p = 0; p < P; p++) {
    s->alloc_q[p] = s->num_proofs[p] = num_proofs[p];
    s->alloc_x[p] = s->num_inputs[p] = num_inputs[p];
    s->off[p] = tot;
    tot += num_proofs[p] * W * num_inputs[p];
  }
  s->off[P] = tot;
  s->data = (ofq *)calloc(tot ? tot : 1, sizeof(ofq));
  s->num_instances = next_pow2(P);
  s->max_num_proofs = max_num_proofs;
  s->num_witness_secs = next_pow2(W);
  s->max_num_inputs = max_num_inputs;
  return s;
}

static inline ofq *at(const opqx *s, size_t p, size_t q, size_t w, size_t x) {
  assert(p < s->P && q < s->alloc_q[p] && w < s->W && x < s->alloc_x[p]);
  return &s->data[s->off[p] + (q * s->W + w) * s->alloc_x[p] + x];
}

opqx *opqx_new(const ofq *z, size_t P, size_t W, const size_t *num_proofs, size_t max_num_proofs,
               const size_t *num_inputs, size_t max_num_inputs) {
  opqx *s = opqx_alloc(P, W, num_proofs, max_num_proofs, num_inputs, max_num_inputs);
  memcpy(s->data, z, sizeof(ofq) * s->off[P]);
  return s;
}

opqx *opqx_new_rev(const ofq *z, size_t P, size_t W, const size_t *num_proofs,
                   size_t max_num_proofs, const size_t *num_inputs, size_t max_num_inputs) {
  opqx *s = opqx_alloc(P, W, num_proofs, max_num_proofs, num_inputs, max_num_inputs);
  for (size_t p = 0; p < P; p++) {
    size_t step_q = max_num_proofs / num_proofs[p];
    size_t step_x = max_num_inputs / num_inputs[p];
#pragma omp parallel for collapse(2) schedule(static) if (num_proofs[p] * num_inputs[p] >= 8192)
    for (size_t q = 0; q < num_proofs[p]; q++) {
      for (size_t x = 0; x < num_inputs[p]; x++) {
        size_t q_rev = orev_bits(q, max_num_proofs) / step_q;
        size_t x_rev = orev_bits(x, max_num_inputs) / step_x;
        for (size_t w = 0; w < W; w++)
          *at(s, p, q_rev, w, x_rev) = z[s->off[p] + (q * W + w) * num_inputs[p] + x];
      }
    }
  }
  return s;
}

opqx *opqx_clone(const opqx *o) {
  opqx *s = opqx_alloc(o->P, o->W, o->alloc_q, o->max_num_proofs, o->alloc_x, o->max_num_inputs);
  memcpy(s->data, o->data, sizeof(ofq) * o->off[o->P]);
  memcpy(s->num_proofs, o->num_proofs, sizeof(size_t) * o->P);
  memcpy(s->num_inputs, o->num_inputs, sizeof(size_t) * o->P);
  s->num_instances = o->num_instances;
  s->num_witness_secs = o->num_witness_secs;
  return s;
}

void opqx_free(opqx *s) {
  if (!s) return;
  free(s->alloc_q); free(s->alloc_x); free(s->num_proofs); free(s->num_inputs);
  free(s->off); free(s->data); free(s);
}

size_t opqx_total(const opqx *s) { return s->off[s->P]; }
void opqx_copy_out(const opqx *s, ofq *out) { memcpy(out, s->data, sizeof(ofq) * s->off[s->P]); }
size_t opqx_len(const opqx *s) { return s->num_instances * s->max_num_proofs * s->max_num_inputs; }

ofq opqx_index(const opqx *s, size_t p, size_t q, size_t w, size_t x) {
  if (p < s->P && q < s->alloc_q[p] && w < s->W && x < s->alloc_x[p]) return *at(s, p, q, w, x);
  return ofq_zero();
}

ofq opqx_index_high(const opqx *s, size_t p, size_t q, size_t w, size_t x, int mode) {
  switch (mode) {
    case OMODE_P:
      if (p + s->num_instances / 2 < s->P) return *at(s, p + s->num_instances / 2, q, w, x);
      return ofq_zero();
    case OMODE_Q:
      if (s->num_proofs[p] == 1) return ofq_zero();
      return *at(s, p, q + s->num_proofs[p] / 2, w, x);
    case OMODE_W:
      if (w + s->num_witness_secs / 2 < s->W) return *at(s, p, q, w + s->num_witness_secs / 2, x);
      return ofq_zero();
    case OMODE_X:
      if (s->num_inputs[p] == 1) return ofq_zero();
      return *at(s, p, q, w, x + s->num_inputs[p] / 2);
    default:
      assert(!"unrecognized mode");
      return ofq_zero();
  }
}

static void fold(ofq *lo, const ofq *hi, const ofq *r) {
  ofq d = ofq_sub(hi, lo);
  ofq t = ofq_mul(r, &d);
  *lo = ofq_add(lo, &t);
}

void opqx_bound_poly(opqx *s, const ofq *r, int mode) {
  ofq one = ofq_one(), zero = ofq_zero();
  ofq one_minus_r = ofq_sub(&one, r);
  switch (mode) {
    case OMODE_P:
      assert(s->max_num_proofs == 1 && s->max_num_inputs == 1);
      s->num_instances /= 2;
      for (size_t p = 0; p < s->num_instances; p++)
        for (size_t w = 0; w < min_sz(s->num_witness_secs, s->W); w++) {
          ofq hi = (p + s->num_instances < s->P) ? *at(s, p + s->num_instances, 0, w, 0) : zero;
          fold(at(s, p, 0, w, 0), &hi, r);
        }
      break;
    case OMODE_Q:
      s->max_num_proofs /= 2;
      for (size_t p = 0; p < min_sz(s->num_instances, s->P); p++) {
        if (s->num_proofs[p] == 1) {
          for (size_t w = 0; w < min_sz(s->num_witness_secs, s->W); w++)
            for (size_t x = 0; x < s->num_inputs[p]; x++) {
              ofq *z = at(s, p, 0, w, x);
              *z = ofq_mul(&one_minus_r, z);
            }
        } else {
          s->num_proofs[p] /= 2;
#pragma omp parallel for schedule(static) if (s->num_proofs[p] * s->num_inputs[p] >= 8192)
          for (size_t q = 0; q < s->num_proofs[p]; q++)
            for (size_t w = 0; w < min_sz(s->num_witness_secs, s->W); w++)
              for (size_t x = 0; x < s->num_inputs[p]; x++)
                fold(at(s, p, q, w, x), at(s, p, q + s->num_proofs[p], w, x), r);
        }
      }
      break;
    case OMODE_W:
      s->num_witness_secs /= 2;
      for (size_t p = 0; p < min_sz(s->num_instances, s->P); p++)
        for (size_t q = 0; q < s->num_proofs[p]; q++)
          for (size_t w = 0; w < s->num_witness_secs; w++)
            for (size_t x = 0; x < s->num_inputs[p]; x++) {
              ofq hi = (w + s->num_witness_secs < s->W) ? *at(s, p, q, w + s->num_witness_secs, x) : zero;
              fold(at(s, p, q, w, x), &hi, r);
            }
      break;
    case OMODE_X:
      s->max_num_inputs /= 2;
      for (size_t p = 0; p < min_sz(s->num_instances, s->P); p++) {
        if (s->num_inputs[p] == 1) {
          for (size_t q = 0; q < s->num_proofs[p]; q++)
            for (size_t w = 0; w < min_sz(s->num_witness_secs, s->W); w++) {
              ofq *z = at(s, p, q, w, 0);
              *z = ofq_mul(&one_minus_r, z);
            }
        } else {
          s->num_inputs[p] /= 2;
#pragma omp parallel for collapse(2) schedule(static) if (s->num_proofs[p] * s->num_inputs[p] >= 8192)
          for (size_t q = 0; q < s->num_proofs[p]; q++)
            for (size_t w = 0; w < min_sz(s->num_witness_secs, s->W); w++)
              for (size_t x = 0; x < s->num_inputs[p]; x++)
                fold(at(s, p, q, w, x), at(s, p, q, w, x + s->num_inputs[p]), r);
        }
      }
      break;
    default:
      assert(!"unrecognized mode");
  }
}

ofq opqx_evaluate(const opqx *s, const ofq *rp, size_t np, const ofq *rq, size_t nq,
                  const ofq *rw, size_t nw, const ofq *rx, size_t nx) {
  opqx *c = opqx_clone(s);
  for (size_t i = 0; i < nx; i++) opqx_bound_poly(c, &rx[i], OMODE_X);
  for (size_t i = 0; i < nw; i++) opqx_bound_poly(c, &rw[i], OMODE_W);
  for (size_t i = 0; i < nq; i++) opqx_bound_poly(c, &rq[i], OMODE_Q);
  for (size_t i = 0; i < np; i++) opqx_bound_poly(c, &rp[i], OMODE_P);
  ofq res = opqx_index(c, 0, 0, 0, 0);
  opqx_free(c);
  return res;
}
