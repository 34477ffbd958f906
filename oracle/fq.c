/*
 * ORACLE (test infrastructure, NOT product code) -- see fq.h.
 * Restates /root/reference/src/scalar/ristretto255.rs with the same limb
 * schedule (4x64-bit schoolbook + HAC 14.32 Montgomery reduction) so that any
 * deviation in the CUDA field library shows up as a bit difference.
 */
#include "fq.h"
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;

const ofq OFQ_MODULUS = {{0x5812631a5cf5d3edULL, 0x14def9dea2f79cd6ULL, 0x0ULL, 0x1000000000000000ULL}};
const ofq OFQ_R = {{0xd6ec31748d98951dULL, 0xc6ef5bf4737dcf70ULL, 0xfffffffffffffffeULL, 0x0fffffffffffffffULL}};
const ofq OFQ_R2 = {{0xa40611e3449c0f01ULL, 0xd00e1ba768859347ULL, 0xceec73d217f5be65ULL, 0x0399411b7c309a3dULL}};
const ofq OFQ_R3 = {{0x2a9e49687b83a2dbULL, 0x278324e6aef7f3ecULL, 0x8065dc6c04ec5b65ULL, 0x0e530b773599cec7ULL}};

/* ristretto255.rs:20-37 */
static inline uint64_t adc(uint64_t a, uint64_t b, uint64_t *carry) {
  u128 t = (u128)a + b + *carry;
  *carry = (uint64_t)(t >> 64);
  return (uint64_t)t;
}
static inline uint64_t sbb(uint64_t a, uint64_t b, uint64_t *borrow) {
  u128 t = (u128)a - ((u128)b + (*borrow >> 63));
  *borrow = (uint64_t)(t >> 64);
  return (uint64_t)t;
}
static inline uint64_t mac(uint64_t a, uint64_t b, uint64_t c, uint64_t *carry) {
  u128 t = (u128)a + (u128)b * c + *carry;
  *carry = (uint64_t)(t >> 64);
  return (uint64_t)t;
}

ofq ofq_zero(void) { ofq z = {{0, 0, 0, 0}}; return z; }
ofq ofq_one(void) { return OFQ_R; }

int ofq_eq(const ofq *a, const ofq *b) {
  return a->l[0] == b->l[0] && a->l[1] == b->l[1] && a->l[2] == b->l[2] && a->l[3] == b->l[3];
}

ofq ofq_sub(const ofq *a, const ofq *b) {
  ofq d;
  uint64_t bw = 0, c = 0;
  for (int i = 0; i < 4; i++) d.l[i] = sbb(a->l[i], b->l[i], &bw);
  /* borrow is all-ones on underflow: add the modulus back under that mask */
  for (int i = 0; i < 4; i++) d.l[i] = adc(d.l[i], OFQ_MODULUS.l[i] & bw, &c);
  return d;
}

ofq ofq_add(const ofq *a, const ofq *b) {
  ofq s;
  uint64_t c = 0;
  for (int i = 0; i < 4; i++) s.l[i] = adc(a->l[i], b->l[i], &c);
  return ofq_sub(&s, &OFQ_MODULUS);
}

ofq ofq_neg(const ofq *a) {
  ofq d;
  uint64_t bw = 0;
  for (int i = 0; i < 4; i++) d.l[i] = sbb(OFQ_MODULUS.l[i], a->l[i], &bw);
  uint64_t nz = (a->l[0] | a->l[1] | a->l[2] | a->l[3]) != 0;
  uint64_t mask = (uint64_t)0 - nz;
  for (int i = 0; i < 4; i++) d.l[i] &= mask;
  return d;
}

ofq ofq_montgomery_reduce(const uint64_t rin[8]) {
  uint64_t r[8];
  memcpy(r, rin, sizeof r);
  uint64_t carry2 = 0;
  for (int i = 0; i < 4; i++) {
    uint64_t k = r[i] * OFQ_INV;
    uint64_t carry = 0;
    (void)mac(r[i], k, OFQ_MODULUS.l[0], &carry);
    for (int j = 1; j < 4; j++) r[i + j] = mac(r[i + j], k, OFQ_MODULUS.l[j], &carry);
    /* r[i+4] += carry2 + carry, new carry2 */
    u128 t = (u128)r[i + 4] + carry2 + carry;
    r[i + 4] = (uint64_t)t;
    carry2 = (uint64_t)(t >> 64);
  }
  ofq hi = {{r[4], r[5], r[6], r[7]}};
  return ofq_sub(&hi, &OFQ_MODULUS);
}

ofq ofq_mul(const ofq *a, const ofq *b) {
  uint64_t r[8] = {0};
  for (int i = 0; i < 4; i++) {
    uint64_t carry = 0;
    for (int j = 0; j < 4; j++) r[i + j] = mac(r[i + j], a->l[i], b->l[j], &carry);
    r[i + 4] = carry;
  }
  return ofq_montgomery_reduce(r);
}

ofq ofq_square(const ofq *a) {
  /* ristretto255.rs:476-504: off-diagonal terms doubled by a shift, then the
   * diagonal added. */
  uint64_t r[8] = {0};
  uint64_t carry;
  carry = 0;
  r[1] = mac(0, a->l[0], a->l[1], &carry);
  r[2] = mac(0, a->l[0], a->l[2], &carry);
  r[3] = mac(0, a->l[0], a->l[3], &carry);
  r[4] = carry;
  carry = 0;
  r[3] = mac(r[3], a->l[1], a->l[2], &carry);
  r[4] = mac(r[4], a->l[1], a->l[3], &carry);
  r[5] = carry;
  carry = 0;
  r[5] = mac(r[5], a->l[2], a->l[3], &carry);
  r[6] = carry;

  r[7] = r[6] >> 63;
  r[6] = (r[6] << 1) | (r[5] >> 63);
  r[5] = (r[5] << 1) | (r[4] >> 63);
  r[4] = (r[4] << 1) | (r[3] >> 63);
  r[3] = (r[3] << 1) | (r[2] >> 63);
  r[2] = (r[2] << 1) | (r[1] >> 63);
  r[1] = r[1] << 1;

  carry = 0;
  r[0] = mac(0, a->l[0], a->l[0], &carry);
  r[1] = adc(0, r[1], &carry);
  r[2] = mac(r[2], a->l[1], a->l[1], &carry);
  r[3] = adc(0, r[3], &carry);
  r[4] = mac(r[4], a->l[2], a->l[2], &carry);
  r[5] = adc(0, r[5], &carry);
  r[6] = mac(r[6], a->l[3], a->l[3], &carry);
  r[7] = adc(0, r[7], &carry);
  return ofq_montgomery_reduce(r);
}

ofq ofq_from_raw(const uint64_t v[4]) {
  ofq t = {{v[0], v[1], v[2], v[3]}};
  return ofq_mul(&t, &OFQ_R2);
}

ofq ofq_from_u64(uint64_t v) {
  uint64_t t[4] = {v, 0, 0, 0};
  return ofq_from_raw(t);
}

ofq ofq_from_u512(const uint64_t v[8]) {
  ofq d0 = {{v[0], v[1], v[2], v[3]}};
  ofq d1 = {{v[4], v[5], v[6], v[7]}};
  ofq x = ofq_mul(&d0, &OFQ_R2);
  ofq y = ofq_mul(&d1, &OFQ_R3);
  return ofq_add(&x, &y);
}

static uint64_t load64(const uint8_t *b) {
  uint64_t v = 0;
  for (int i = 7; i >= 0; i--) v = (v << 8) | b[i];
  return v;
}

int ofq_from_bytes(const uint8_t b[32], ofq *out) {
  ofq t;
  for (int i = 0; i < 4; i++) t.l[i] = load64(b + 8 * i);
  uint64_t bw = 0;
  for (int i = 0; i < 4; i++) (void)sbb(t.l[i], OFQ_MODULUS.l[i], &bw);
  int is_some = (int)(bw & 1);
  *out = ofq_mul(&t, &OFQ_R2);
  return is_some;
}

void ofq_to_bytes(const ofq *a, uint8_t out[32]) {
  uint64_t r[8] = {a->l[0], a->l[1], a->l[2], a->l[3], 0, 0, 0, 0};
  ofq t = ofq_montgomery_reduce(r);
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 8; j++) out[8 * i + j] = (uint8_t)(t.l[i] >> (8 * j));
}

ofq ofq_from_bytes_wide(const uint8_t b[64]) {
  uint64_t v[8];
  for (int i = 0; i < 8; i++) v[i] = load64(b + 8 * i);
  return ofq_from_u512(v);
}

ofq ofq_pow(const ofq *a, const uint64_t by[4]) {
  ofq res = ofq_one();
  for (int e = 3; e >= 0; e--)
    for (int i = 63; i >= 0; i--) {
      res = ofq_square(&res);
      if ((by[e] >> i) & 1) res = ofq_mul(&res, a);
    }
  return res;
}

static void sqmul(ofq *y, int squarings, const ofq *x) {
  for (int i = 0; i < squarings; i++) *y = ofq_square(y);
  *y = ofq_mul(y, x);
}

ofq ofq_invert(const ofq *a) {
  /* same addition chain as ristretto255.rs:541-595 */
  ofq _1 = *a;
  ofq _10 = ofq_square(&_1);
  ofq _100 = ofq_square(&_10);
  ofq _11 = ofq_mul(&_10, &_1);
  ofq _101 = ofq_mul(&_10, &_11);
  ofq _111 = ofq_mul(&_10, &_101);
  ofq _1001 = ofq_mul(&_10, &_111);
  ofq _1011 = ofq_mul(&_10, &_1001);
  ofq _1111 = ofq_mul(&_100, &_1011);
  ofq y = ofq_mul(&_1111, &_1);
  sqmul(&y, 123 + 3, &_101);
  sqmul(&y, 2 + 2, &_11);
  sqmul(&y, 1 + 4, &_1111);
  sqmul(&y, 1 + 4, &_1111);
  sqmul(&y, 4, &_1001);
  sqmul(&y, 2, &_11);
  sqmul(&y, 1 + 4, &_1111);
  sqmul(&y, 1 + 3, &_101);
  sqmul(&y, 3 + 3, &_101);
  sqmul(&y, 3, &_111);
  sqmul(&y, 1 + 4, &_1111);
  sqmul(&y, 2 + 3, &_111);
  sqmul(&y, 2 + 2, &_11);
  sqmul(&y, 1 + 4, &_1011);
  sqmul(&y, 2 + 4, &_1011);
  sqmul(&y, 6 + 4, &_1001);
  sqmul(&y, 2 + 2, &_11);
  sqmul(&y, 3 + 2, &_11);
  sqmul(&y, 3 + 2, &_11);
  sqmul(&y, 1 + 4, &_1001);
  sqmul(&y, 1 + 3, &_111);
  sqmul(&y, 2 + 4, &_1111);
  sqmul(&y, 1 + 4, &_1011);
  sqmul(&y, 3, &_101);
  sqmul(&y, 2 + 4, &_1111);
  sqmul(&y, 3, &_101);
  sqmul(&y, 1 + 2, &_11);
  return y;
}

ofq ofq_batch_invert(ofq *inputs, size_t n) {
  ofq *scratch = (ofq *)malloc(sizeof(ofq) * (n ? n : 1));
  ofq acc = ofq_one();
  for (size_t i = 0; i < n; i++) {
    scratch[i] = acc;
    acc = ofq_mul(&acc, &inputs[i]);
  }
  acc = ofq_invert(&acc);
  ofq ret = acc;
  for (size_t i = n; i-- > 0;) {
    ofq tmp = ofq_mul(&acc, &inputs[i]);
    inputs[i] = ofq_mul(&acc, &scratch[i]);
    acc = tmp;
  }
  free(scratch);
  return ret;
}

void ofq_vec_mul(const ofq *a, const ofq *b, ofq *out, size_t n) {
  for (size_t i = 0; i < n; i++) out[i] = ofq_mul(&a[i], &b[i]);
}
void ofq_vec_add(const ofq *a, const ofq *b, ofq *out, size_t n) {
  for (size_t i = 0; i < n; i++) out[i] = ofq_add(&a[i], &b[i]);
}
void ofq_vec_sub(const ofq *a, const ofq *b, ofq *out, size_t n) {
  for (size_t i = 0; i < n; i++) out[i] = ofq_sub(&a[i], &b[i]);
}
void ofq_vec_from_u512(const uint64_t *wide, ofq *out, size_t n) {
  for (size_t i = 0; i < n; i++) out[i] = ofq_from_u512(wide + 8 * i);
}
