// GF(2^255 - 19) and ristretto255 group arithmetic for the commitment kernels.
// Replaces the curve25519-dalek calls behind src/group.rs:87-117 and
// src/commitments.rs:69-92 (Pedersen vector commitments = multiscalar multiplication),
// and CompressedRistretto encode/decode (RFC 9496 4.3.1 / 4.3.2).
//
// Field elements are ten unsigned limbs in radix 2^25.5 (26/25/26/... bits): every
// limb product fits a 64-bit accumulator without carries (ptxas emits IMAD.WIDE.U32 products
// summed by three-input IADD3 pairs), and the reduction by 19 is folded into the
// operands. All functions take and return "reduced" elements (limbs <= 2^26 / 2^25 plus
// a few units) so no bound bookkeeping leaks to the callers.
//
// The code is __host__ __device__ on purpose: tools/ed_host_check.cu runs it on the CPU
// against vectors produced by the oracle before any GPU time is spent.
#pragma once
#include <cstdint>

#ifndef __CUDACC__
#define __host__
#define __device__
#define __forceinline__ inline
#endif

namespace spg {

struct fe {
  uint32_t v[10];
};

}  // namespace spg
#include "ed_consts.cuh"
namespace spg {

#define SPG_FE_M26 0x3ffffffu
#define SPG_FE_M25 0x1ffffffu

__host__ __device__ __forceinline__ fe fe_zero() {
  fe r;
  for (int i = 0; i < 10; i++) r.v[i] = 0;
  return r;
}
__host__ __device__ __forceinline__ fe fe_one() {
  fe r = fe_zero();
  r.v[0] = 1;
  return r;
}

// carry a vector of 64-bit column sums into reduced limbs
__host__ __device__ __forceinline__ fe fe_carry64(uint64_t h[10]) {
  uint64_t c;
#pragma unroll
  for (int i = 0; i < 9; i++) {
    int bits = (i & 1) ? 25 : 26;
    c = h[i] >> bits;
    h[i] &= ((uint64_t)1 << bits) - 1;
    h[i + 1] += c;
  }
  c = h[9] >> 25;
  h[9] &= SPG_FE_M25;
  h[0] += 19 * c;
  c = h[0] >> 26;
  h[0] &= SPG_FE_M26;
  h[1] += c;
  fe r;
#pragma unroll
  for (int i = 0; i < 10; i++) r.v[i] = (uint32_t)h[i];
  return r;
}

__host__ __device__ __forceinline__ fe fe_add(const fe &a, const fe &b) {
  uint64_t h[10];
#pragma unroll
  for (int i = 0; i < 10; i++) h[i] = (uint64_t)a.v[i] + b.v[i];
  return fe_carry64(h);
}

// a - b computed as a + 4p - b (limbs of 4p: 4*(2^26-19), 4*(2^25-1), 4*(2^26-1), ...)
__host__ __device__ __forceinline__ fe fe_sub(const fe &a, const fe &b) {
  uint64_t h[10];
  h[0] = (uint64_t)a.v[0] + 0xfffffb4u - b.v[0];
#pragma unroll
  for (int i = 1; i < 10; i++) h[i] = (uint64_t)a.v[i] + ((i & 1) ? 0x7fffffcu : 0xffffffcu) - b.v[i];
  return fe_carry64(h);
}

__host__ __device__ __forceinline__ fe fe_neg(const fe &a) { return fe_sub(fe_zero(), a); }

__host__ __device__ __forceinline__ fe fe_mul(const fe &f, const fe &g) {
#if !defined(__CUDA_ARCH__) && defined(__SIZEOF_INT128__)
  // Host build (the transcript-side mirror, host/group.hpp): pair the limbs into five of
  // radix 2^51 and use 64 x 64 -> 128 products: 25 multiplications instead of 100. The result
  // is unpacked into the same reduced ten-limb form, so callers cannot tell the difference
  // (equal modulo p; fe_tobytes canonicalises).
  typedef unsigned __int128 u128;
  uint64_t a[5], b[5], b19[5];
  for (int i = 0; i < 5; i++) {
    a[i] = (uint64_t)f.v[2 * i] + ((uint64_t)f.v[2 * i + 1] << 26);
    b[i] = (uint64_t)g.v[2 * i] + ((uint64_t)g.v[2 * i + 1] << 26);
    b19[i] = 19 * b[i];
  }
  u128 r0 = (u128)a[0] * b[0] + (u128)a[1] * b19[4] + (u128)a[2] * b19[3] + (u128)a[3] * b19[2] + (u128)a[4] * b19[1];
  u128 r1 = (u128)a[0] * b[1] + (u128)a[1] * b[0] + (u128)a[2] * b19[4] + (u128)a[3] * b19[3] + (u128)a[4] * b19[2];
  u128 r2 = (u128)a[0] * b[2] + (u128)a[1] * b[1] + (u128)a[2] * b[0] + (u128)a[3] * b19[4] + (u128)a[4] * b19[3];
  u128 r3 = (u128)a[0] * b[3] + (u128)a[1] * b[2] + (u128)a[2] * b[1] + (u128)a[3] * b[0] + (u128)a[4] * b19[4];
  u128 r4 = (u128)a[0] * b[4] + (u128)a[1] * b[3] + (u128)a[2] * b[2] + (u128)a[3] * b[1] + (u128)a[4] * b[0];
  const uint64_t M51 = ((uint64_t)1 << 51) - 1;
  r1 += (uint64_t)(r0 >> 51);
  r2 += (uint64_t)(r1 >> 51);
  r3 += (uint64_t)(r2 >> 51);
  r4 += (uint64_t)(r3 >> 51);
  uint64_t t0 = (uint64_t)r0 & M51, t1 = (uint64_t)r1 & M51, t2 = (uint64_t)r2 & M51, t3 = (uint64_t)r3 & M51,
           t4 = (uint64_t)r4 & M51;
  t0 += 19 * (uint64_t)(r4 >> 51);
  t1 += t0 >> 51;
  t0 &= M51;
  const uint64_t t[5] = {t0, t1, t2, t3, t4};
  fe r;
  for (int i = 0; i < 5; i++) {
    r.v[2 * i] = (uint32_t)(t[i] & SPG_FE_M26);
    r.v[2 * i + 1] = (uint32_t)(t[i] >> 26);
  }
  return r;
#else
  uint32_t g19[10], f2[10];
#pragma unroll
  for (int i = 0; i < 10; i++) {
    g19[i] = 19u * g.v[i];
    f2[i] = (i & 1) ? 2u * f.v[i] : f.v[i];
  }
  uint64_t h[10];
#pragma unroll
  for (int k = 0; k < 10; k++) h[k] = 0;
#pragma unroll
  for (int i = 0; i < 10; i++) {
#pragma unroll
    for (int j = 0; j < 10; j++) {
      int k = i + j;
      // both odd -> the product sits one bit above limb k's position
      uint32_t fi = ((i & 1) && (j & 1)) ? f2[i] : f.v[i];
      if (k >= 10) h[k - 10] += (uint64_t)fi * g19[j];
      else h[k] += (uint64_t)fi * g.v[j];
    }
  }
  return fe_carry64(h);
#endif
}

__host__ __device__ __forceinline__ fe fe_sq(const fe &f) { return fe_mul(f, f); }

__host__ __device__ __forceinline__ fe fe_sqn(fe f, int n) {
  for (int i = 0; i < n; i++) f = fe_sq(f);
  return f;
}

// f^(2^252 - 3) = f^((p-5)/8)
__host__ __device__ inline fe fe_pow22523(const fe &z) {
  fe t0 = fe_sq(z);
  fe t1 = fe_sqn(t0, 2);
  t1 = fe_mul(z, t1);
  t0 = fe_mul(t0, t1);
  t0 = fe_sq(t0);
  t0 = fe_mul(t1, t0);
  t1 = fe_sqn(t0, 5);
  t0 = fe_mul(t1, t0);
  t1 = fe_sqn(t0, 10);
  t1 = fe_mul(t1, t0);
  fe t2 = fe_sqn(t1, 20);
  t1 = fe_mul(t2, t1);
  t1 = fe_sqn(t1, 10);
  t0 = fe_mul(t1, t0);
  t1 = fe_sqn(t0, 50);
  t1 = fe_mul(t1, t0);
  t2 = fe_sqn(t1, 100);
  t1 = fe_mul(t2, t1);
  t1 = fe_sqn(t1, 50);
  t0 = fe_mul(t1, t0);
  t0 = fe_sqn(t0, 2);
  return fe_mul(t0, z);
}

// canonical little-endian bytes (value fully reduced mod p)
__host__ __device__ inline void fe_tobytes(const fe &f, uint8_t out[32]) {
  uint64_t h[10];
  for (int i = 0; i < 10; i++) h[i] = f.v[i];
  fe t = fe_carry64(h);
  // q = floor((t + 19) / 2^255): 1 iff t >= p
  uint64_t q = (19ull * t.v[9] + ((uint64_t)1 << 24)) >> 25;
  for (int i = 0; i < 10; i++) {
    int bits = (i & 1) ? 25 : 26;
    q = (t.v[i] + q) >> bits;
  }
  uint64_t c = 19 * q;
  uint32_t l[10];
  for (int i = 0; i < 10; i++) {
    int bits = (i & 1) ? 25 : 26;
    uint64_t s = t.v[i] + c;
    l[i] = (uint32_t)(s & (((uint64_t)1 << bits) - 1));
    c = s >> bits;
  }
  // pack 255 bits
  const int shifts[10] = {0, 26, 51, 77, 102, 128, 153, 179, 204, 230};
  for (int i = 0; i < 32; i++) out[i] = 0;
  for (int i = 0; i < 10; i++) {
    uint64_t v = l[i];
    int bit = shifts[i];
    int byte = bit >> 3, sh = bit & 7;
    uint64_t w = v << sh;
    for (int k = 0; k < 5 && byte + k < 32; k++) out[byte + k] |= (uint8_t)(w >> (8 * k));
  }
}

// little-endian bytes -> limbs; the top bit (bit 255) is ignored
__host__ __device__ inline fe fe_frombytes(const uint8_t in[32]) {
  uint64_t w[4];
  for (int i = 0; i < 4; i++) {
    w[i] = 0;
    for (int k = 7; k >= 0; k--) w[i] = (w[i] << 8) | in[8 * i + k];
  }
  w[3] &= 0x7fffffffffffffffull;
  const int shifts[10] = {0, 26, 51, 77, 102, 128, 153, 179, 204, 230};
  fe r;
  for (int i = 0; i < 10; i++) {
    int bit = shifts[i], bits = (i & 1) ? 25 : 26;
    int word = bit >> 6, sh = bit & 63;
    uint64_t v = w[word] >> sh;
    if (sh + bits > 64 && word < 3) v |= w[word + 1] << (64 - sh);
    r.v[i] = (uint32_t)(v & (((uint64_t)1 << bits) - 1));
  }
  return r;
}

__host__ __device__ inline bool fe_is_negative(const fe &f) {
  uint8_t b[32];
  fe_tobytes(f, b);
  return b[0] & 1;
}

__host__ __device__ inline bool fe_equal(const fe &a, const fe &b) {
  uint8_t x[32], y[32];
  fe_tobytes(a, x);
  fe_tobytes(b, y);
  uint8_t d = 0;
  for (int i = 0; i < 32; i++) d |= x[i] ^ y[i];
  return d == 0;
}

__host__ __device__ inline bool fe_is_zero(const fe &a) { return fe_equal(a, fe_zero()); }

__host__ __device__ inline fe fe_abs(const fe &a) { return fe_is_negative(a) ? fe_neg(a) : a; }

// RFC 9496 4.2 SQRT_RATIO_M1
__host__ __device__ inline bool fe_sqrt_ratio_m1(const fe &u, const fe &v, fe *out) {
  fe v3 = fe_mul(fe_sq(v), v);
  fe v7 = fe_mul(fe_sq(v3), v);
  fe r = fe_mul(fe_mul(u, v3), fe_pow22523(fe_mul(u, v7)));
  fe check = fe_mul(v, fe_sq(r));
  fe neg_u = fe_neg(u);
  bool correct = fe_equal(check, u);
  bool flipped = fe_equal(check, neg_u);
  bool flipped_i = fe_equal(check, fe_mul(neg_u, fe_sqrt_m1()));
  if (flipped || flipped_i) r = fe_mul(r, fe_sqrt_m1());
  *out = fe_abs(r);
  return correct || flipped;
}

// ---------------------------------------------------------------- points
struct ge {       // extended coordinates (X : Y : Z : T), x = X/Z, y = Y/Z, xy = T/Z
  fe X, Y, Z, T;
};
struct ge_cached {  // (Y+X, Y-X, Z, 2dT): the second operand of an addition
  fe YpX, YmX, Z, T2d;
};

__host__ __device__ __forceinline__ ge ge_identity() {
  ge r;
  r.X = fe_zero();
  r.Y = fe_one();
  r.Z = fe_one();
  r.T = fe_zero();
  return r;
}

__host__ __device__ __forceinline__ ge_cached ge_to_cached(const ge &p) {
  ge_cached c;
  c.YpX = fe_add(p.Y, p.X);
  c.YmX = fe_sub(p.Y, p.X);
  c.Z = p.Z;
  c.T2d = fe_mul(p.T, fe_2d());
  return c;
}

// unified addition (add-2008-hwcd-3), 9 multiplications
__host__ __device__ __forceinline__ ge ge_add(const ge &p, const ge_cached &q) {
  fe A = fe_mul(fe_sub(p.Y, p.X), q.YmX);
  fe B = fe_mul(fe_add(p.Y, p.X), q.YpX);
  fe C = fe_mul(p.T, q.T2d);
  fe ZZ = fe_mul(p.Z, q.Z);
  fe D = fe_add(ZZ, ZZ);
  fe E = fe_sub(B, A), F = fe_sub(D, C), G = fe_add(D, C), H = fe_add(B, A);
  ge r;
  r.X = fe_mul(E, F);
  r.Y = fe_mul(G, H);
  r.Z = fe_mul(F, G);
  r.T = fe_mul(E, H);
  return r;
}

// doubling (dbl-2008-hwcd)
__host__ __device__ __forceinline__ ge ge_double(const ge &p) {
  fe A = fe_sq(p.X), B = fe_sq(p.Y);
  fe ZZ = fe_sq(p.Z);
  fe C = fe_add(ZZ, ZZ);
  fe H = fe_add(A, B);
  fe XY = fe_add(p.X, p.Y);
  fe E = fe_sub(H, fe_sq(XY));
  fe G = fe_sub(A, B);
  fe F = fe_add(C, G);
  ge r;
  r.X = fe_mul(E, F);
  r.Y = fe_mul(G, H);
  r.Z = fe_mul(F, G);
  r.T = fe_mul(E, H);
  return r;
}

// RFC 9496 4.3.2 Encode
__host__ __device__ inline void ristretto_compress(const ge &p, uint8_t out[32]) {
  fe u1 = fe_mul(fe_add(p.Z, p.Y), fe_sub(p.Z, p.Y));
  fe u2 = fe_mul(p.X, p.Y);
  fe invsqrt;
  fe_sqrt_ratio_m1(fe_one(), fe_mul(u1, fe_sq(u2)), &invsqrt);
  fe den1 = fe_mul(invsqrt, u1);
  fe den2 = fe_mul(invsqrt, u2);
  fe z_inv = fe_mul(fe_mul(den1, den2), p.T);
  fe ix0 = fe_mul(p.X, fe_sqrt_m1());
  fe iy0 = fe_mul(p.Y, fe_sqrt_m1());
  fe enchanted = fe_mul(den1, fe_invsqrt_a_minus_d());
  bool rotate = fe_is_negative(fe_mul(p.T, z_inv));
  fe x = rotate ? iy0 : p.X;
  fe y = rotate ? ix0 : p.Y;
  fe den_inv = rotate ? enchanted : den2;
  if (fe_is_negative(fe_mul(x, z_inv))) y = fe_neg(y);
  fe s = fe_abs(fe_mul(den_inv, fe_sub(p.Z, y)));
  fe_tobytes(s, out);
}

// RFC 9496 4.3.1 Decode; false for an invalid encoding
__host__ __device__ inline bool ristretto_decompress(const uint8_t in[32], ge *out) {
  fe s = fe_frombytes(in);
  uint8_t chk[32];
  fe_tobytes(s, chk);
  uint8_t d = 0;
  for (int i = 0; i < 32; i++) d |= chk[i] ^ in[i];
  if (d != 0 || (in[0] & 1)) return false;  // non-canonical or negative
  fe ss = fe_sq(s);
  fe u1 = fe_sub(fe_one(), ss);
  fe u2 = fe_add(fe_one(), ss);
  fe u2_sqr = fe_sq(u2);
  fe v = fe_sub(fe_neg(fe_mul(fe_d(), fe_sq(u1))), u2_sqr);
  fe invsqrt;
  bool was_square = fe_sqrt_ratio_m1(fe_one(), fe_mul(v, u2_sqr), &invsqrt);
  fe den_x = fe_mul(invsqrt, u2);
  fe den_y = fe_mul(fe_mul(invsqrt, den_x), v);
  fe x = fe_abs(fe_mul(fe_add(s, s), den_x));
  fe y = fe_mul(u1, den_y);
  fe t = fe_mul(x, y);
  if (!was_square || fe_is_negative(t) || fe_is_zero(y)) return false;
  out->X = x;
  out->Y = y;
  out->Z = fe_one();
  out->T = t;
  return true;
}

}  // namespace spg

namespace spg {

// RFC 9496 4.3.4 MAP (one half of the element derivation / dalek's elligator_ristretto_flavor)
__host__ __device__ inline ge ristretto_map(const fe &t) {
  fe one = fe_one();
  fe r = fe_mul(fe_sqrt_m1(), fe_sq(t));
  fe u = fe_mul(fe_add(r, one), fe_one_minus_d_sq());
  fe v = fe_mul(fe_sub(fe_neg(one), fe_mul(r, fe_d())), fe_add(r, fe_d()));
  fe s;
  bool was_square = fe_sqrt_ratio_m1(u, v, &s);
  fe s_prime = fe_neg(fe_abs(fe_mul(s, t)));
  if (!was_square) s = s_prime;
  fe c = was_square ? fe_neg(one) : r;
  fe N = fe_sub(fe_mul(fe_mul(c, fe_sub(r, one)), fe_d_minus_one_sq()), v);
  fe ss = fe_sq(s);
  fe w0 = fe_mul(fe_add(s, s), v);
  fe w1 = fe_mul(N, fe_sqrt_ad_minus_one());
  fe w2 = fe_sub(one, ss);
  fe w3 = fe_add(one, ss);
  ge p;
  p.X = fe_mul(w0, w3);
  p.Y = fe_mul(w2, w1);
  p.Z = fe_mul(w1, w3);
  p.T = fe_mul(w0, w2);
  return p;
}

// RistrettoPoint::from_uniform_bytes (used by MultiCommitGens::new, src/commitments.rs:23-27)
__host__ __device__ inline ge ristretto_from_uniform_bytes(const uint8_t b[64]) {
  ge p1 = ristretto_map(fe_frombytes(b));       // fe_frombytes masks bit 255
  ge p2 = ristretto_map(fe_frombytes(b + 32));
  return ge_add(p1, ge_to_cached(p2));
}

}  // namespace spg
