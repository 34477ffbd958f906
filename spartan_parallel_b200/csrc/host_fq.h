// Host-side F_q arithmetic for the scalar glue around the kernels (round-polynomial
// scale factors, eq(tau, r) products, final claims) and for the C++ host mirror of
// the reference's transcript-side code. Same representation as the reference's
// `Scalar` (src/scalar/ristretto255.rs:193-199): 4 x u64, a*2^256 mod q, canonical.
// This is product code: it does not share anything with oracle/.
#pragma once
#include <cstdint>
#include <cstring>

#include "../../include/spgpu.h"

namespace spg {

struct hfq {
  uint64_t l[4];
};

namespace hfqc {
static const uint64_t Q[4] = {0x5812631a5cf5d3edULL, 0x14def9dea2f79cd6ULL, 0x0ULL,
                              0x1000000000000000ULL};
static const uint64_t ONE[4] = {0xd6ec31748d98951dULL, 0xc6ef5bf4737dcf70ULL,
                                0xfffffffffffffffeULL, 0x0fffffffffffffffULL};
static const uint64_t R2[4] = {0xa40611e3449c0f01ULL, 0xd00e1ba768859347ULL,
                               0xceec73d217f5be65ULL, 0x0399411b7c309a3dULL};
static const uint64_t R3[4] = {0x2a9e49687b83a2dbULL, 0x278324e6aef7f3ecULL,
                               0x8065dc6c04ec5b65ULL, 0x0e530b773599cec7ULL};
static const uint64_t NINV = 0xd2b51da312547e1bULL;  // -q^-1 mod 2^64
}  // namespace hfqc

static inline hfq hfq_zero() { return hfq{{0, 0, 0, 0}}; }
static inline hfq hfq_one() { return hfq{{hfqc::ONE[0], hfqc::ONE[1], hfqc::ONE[2], hfqc::ONE[3]}}; }
static inline hfq hfq_from(const spg_fq &a) { return hfq{{a.l[0], a.l[1], a.l[2], a.l[3]}}; }
static inline spg_fq hfq_to(const hfq &a) { return spg_fq{{a.l[0], a.l[1], a.l[2], a.l[3]}}; }
static inline bool hfq_eq(const hfq &a, const hfq &b) { return memcmp(a.l, b.l, 32) == 0; }
static inline bool hfq_is_zero(const hfq &a) { return (a.l[0] | a.l[1] | a.l[2] | a.l[3]) == 0; }

// returns a - q if a >= q else a (a < 2q)
static inline hfq hfq_reduce_once(const uint64_t a[4], uint64_t top) {
  uint64_t d[4];
  unsigned __int128 bw = 0;
  for (int i = 0; i < 4; i++) {
    unsigned __int128 t = (unsigned __int128)a[i] - hfqc::Q[i] - (uint64_t)bw;
    d[i] = (uint64_t)t;
    bw = (t >> 64) & 1;
  }
  bool ge = top || !bw;
  hfq r;
  for (int i = 0; i < 4; i++) r.l[i] = ge ? d[i] : a[i];
  return r;
}

static inline hfq hfq_add(const hfq &a, const hfq &b) {
  uint64_t s[4];
  unsigned __int128 c = 0;
  for (int i = 0; i < 4; i++) {
    c += (unsigned __int128)a.l[i] + b.l[i];
    s[i] = (uint64_t)c;
    c >>= 64;
  }
  return hfq_reduce_once(s, (uint64_t)c);
}

static inline hfq hfq_sub(const hfq &a, const hfq &b) {
  uint64_t d[4];
  unsigned __int128 bw = 0;
  for (int i = 0; i < 4; i++) {
    unsigned __int128 t = (unsigned __int128)a.l[i] - b.l[i] - (uint64_t)bw;
    d[i] = (uint64_t)t;
    bw = (t >> 64) & 1;
  }
  if (bw) {
    unsigned __int128 c = 0;
    for (int i = 0; i < 4; i++) {
      c += (unsigned __int128)d[i] + hfqc::Q[i];
      d[i] = (uint64_t)c;
      c >>= 64;
    }
  }
  return hfq{{d[0], d[1], d[2], d[3]}};
}

static inline hfq hfq_neg(const hfq &a) { return hfq_sub(hfq_zero(), a); }

// coarsely-integrated Montgomery product, 64-bit limbs
static inline hfq hfq_mul(const hfq &a, const hfq &b) {
  uint64_t t[6] = {0, 0, 0, 0, 0, 0};
  for (int i = 0; i < 4; i++) {
    unsigned __int128 c = 0;
    for (int j = 0; j < 4; j++) {
      c += (unsigned __int128)a.l[j] * b.l[i] + t[j];
      t[j] = (uint64_t)c;
      c >>= 64;
    }
    c += t[4];
    t[4] = (uint64_t)c;
    t[5] = (uint64_t)(c >> 64);
    uint64_t m = t[0] * hfqc::NINV;
    c = ((unsigned __int128)m * hfqc::Q[0] + t[0]) >> 64;
    for (int j = 1; j < 4; j++) {
      c += (unsigned __int128)m * hfqc::Q[j] + t[j];
      t[j - 1] = (uint64_t)c;
      c >>= 64;
    }
    c += t[4];
    t[3] = (uint64_t)c;
    t[4] = t[5] + (uint64_t)(c >> 64);
  }
  return hfq_reduce_once(t, t[4]);
}

static inline hfq hfq_from_u64(uint64_t v) {
  hfq t{{v, 0, 0, 0}};
  hfq r2{{hfqc::R2[0], hfqc::R2[1], hfqc::R2[2], hfqc::R2[3]}};
  return hfq_mul(t, r2);
}

// Scalar::from_u512 (src/scalar/ristretto255.rs:448-466)
static inline hfq hfq_from_u512(const uint64_t w[8]) {
  hfq d0{{w[0], w[1], w[2], w[3]}}, d1{{w[4], w[5], w[6], w[7]}};
  hfq r2{{hfqc::R2[0], hfqc::R2[1], hfqc::R2[2], hfqc::R2[3]}};
  hfq r3{{hfqc::R3[0], hfqc::R3[1], hfqc::R3[2], hfqc::R3[3]}};
  return hfq_add(hfq_mul(d0, r2), hfq_mul(d1, r3));
}

// Scalar::to_bytes (src/scalar/ristretto255.rs:419-431): canonical little-endian integer
static inline void hfq_to_bytes(const hfq &a, uint8_t out[32]) {
  hfq one_raw{{1, 0, 0, 0}};
  hfq t = hfq_mul(a, one_raw);
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 8; j++) out[8 * i + j] = (uint8_t)(t.l[i] >> (8 * j));
}

static inline hfq hfq_pow(const hfq &a, const uint64_t e[4]) {
  hfq r = hfq_one();
  for (int w = 3; w >= 0; w--)
    for (int i = 63; i >= 0; i--) {
      r = hfq_mul(r, r);
      if ((e[w] >> i) & 1) r = hfq_mul(r, a);
    }
  return r;
}

static inline hfq hfq_invert(const hfq &a) {
  static const uint64_t QM2[4] = {0x5812631a5cf5d3ebULL, 0x14def9dea2f79cd6ULL, 0x0ULL,
                                  0x1000000000000000ULL};
  return hfq_pow(a, QM2);
}

// eq(a, b) for one variable: a*b + (1-a)*(1-b)  (EqPolynomial::evaluate, dense_mlpoly.rs:69-74)
static inline hfq hfq_eq1(const hfq &a, const hfq &b) {
  hfq one = hfq_one();
  return hfq_add(hfq_mul(a, b), hfq_mul(hfq_sub(one, a), hfq_sub(one, b)));
}

// value at t in {0, 2, 3} of the line through (0, 1-a), (1, a): the per-round
// eq factor l_j(t) that the device leaves to the host
static inline void hfq_eq_line_023(const hfq &a, hfq out[3]) {
  hfq one = hfq_one();
  hfq lo = hfq_sub(one, a);          // t = 0
  hfq d = hfq_sub(a, lo);            // slope
  hfq at2 = hfq_add(a, d);           // t = 2
  hfq at3 = hfq_add(at2, d);         // t = 3
  out[0] = lo;
  out[1] = at2;
  out[2] = at3;
}

}  // namespace spg
