// GF(2^255 - 19) on eight saturated 32-bit limbs, for the inner loop of the commitment
// kernels (the point additions behind src/commitments.rs:69-92 / src/dense_mlpoly.rs:199-239).
//
// Why a second representation next to ed25519.cuh's ten limbs of radix 2^25.5: a
// multiplication there is 100 IMAD.WIDE.U32 plus ~200 IADD3 (ptxas does not fuse the 64-bit
// addend on sm_100a) plus the carry pass. Here a product is 64 + 8 carry-chained
// IMAD.WIDE.U32.X (mad.lo.cc / madc.hi.cc pairs on aligned register pairs, the same split
// even/odd accumulator trick as fq.cuh) and ~45 add-with-carry: the integer-multiplier pipe is
// the bound, and this form issues 28 % fewer multiplier instructions and a quarter of the adds.
//
// A value is any integer in [0, 2^256) congruent to the field element (2^256 = 38 mod p), so
// additions and subtractions only fold their carry / borrow back in (times 38); nothing is
// fully reduced until fe8_to_fe.
//
// Device only (inline PTX). Checked against the ten-limb code, which the CPU tests pin to the
// oracle, by spg_debug_fe8_selftest (csrc/msm.cu, tests/test_gpu_commit.py).
#pragma once
#include <cstdint>

#include "ed25519.cuh"

namespace spg {

struct fe8 {
  uint32_t v[8];
};

__device__ __forceinline__ fe8 fe8_zero() {
  fe8 r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = 0;
  return r;
}
__device__ __forceinline__ fe8 fe8_one() {
  fe8 r = fe8_zero();
  r.v[0] = 1;
  return r;
}

// r[0..7] += (a0, a1, a2, a3) * b as four aligned 64-bit products; the carry goes into `top`
// (the next limb, which by the bound on the partial sums never carries out itself)
__device__ __forceinline__ void fe8_chain(uint32_t *r, uint32_t &top, uint32_t a0, uint32_t a1, uint32_t a2,
                                          uint32_t a3, uint32_t b) {
  asm("{\n\t"
      "mad.lo.cc.u32   %0, %9,  %13, %0;\n\t"
      "madc.hi.cc.u32  %1, %9,  %13, %1;\n\t"
      "madc.lo.cc.u32  %2, %10, %13, %2;\n\t"
      "madc.hi.cc.u32  %3, %10, %13, %3;\n\t"
      "madc.lo.cc.u32  %4, %11, %13, %4;\n\t"
      "madc.hi.cc.u32  %5, %11, %13, %5;\n\t"
      "madc.lo.cc.u32  %6, %12, %13, %6;\n\t"
      "madc.hi.cc.u32  %7, %12, %13, %7;\n\t"
      "addc.u32        %8, %8, 0;\n\t"
      "}"
      : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(top)
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b));
}
// the same without a limb above (the last rows of the product: the carry is provably zero)
__device__ __forceinline__ void fe8_chain_last(uint32_t *r, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                               uint32_t b) {
  asm("{\n\t"
      "mad.lo.cc.u32   %0, %8,  %12, %0;\n\t"
      "madc.hi.cc.u32  %1, %8,  %12, %1;\n\t"
      "madc.lo.cc.u32  %2, %9,  %12, %2;\n\t"
      "madc.hi.cc.u32  %3, %9,  %12, %3;\n\t"
      "madc.lo.cc.u32  %4, %10, %12, %4;\n\t"
      "madc.hi.cc.u32  %5, %10, %12, %5;\n\t"
      "madc.lo.cc.u32  %6, %11, %12, %6;\n\t"
      "madc.hi.u32     %7, %11, %12, %7;\n\t"
      "}"
      : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b));
}

// fold the 512-bit value T (sixteen limbs) to eight: T_lo + 38 T_hi, then the small overflow again
__device__ __forceinline__ fe8 fe8_fold(const uint32_t *T) {
  uint32_t R[9];
#pragma unroll
  for (int i = 0; i < 8; i++) R[i] = T[i];
  R[8] = 0;
  const uint32_t c38 = 38u;
  // even limbs of T_hi land on aligned pairs (0,1) (2,3) (4,5) (6,7), odd ones on (1,2) ... (7,8)
  fe8_chain(R, R[8], T[8], T[10], T[12], T[14], c38);
  fe8_chain_last(R + 1, T[9], T[11], T[13], T[15], c38);
  // R[8] <= 38: fold it, and the (rare) carry of that fold once more
  uint32_t k = R[8] * 38u, c2;
  fe8 r;
  asm("{\n\t"
      "add.cc.u32  %0, %9,  %17;\n\t"
      "addc.cc.u32 %1, %10, 0;\n\t"
      "addc.cc.u32 %2, %11, 0;\n\t"
      "addc.cc.u32 %3, %12, 0;\n\t"
      "addc.cc.u32 %4, %13, 0;\n\t"
      "addc.cc.u32 %5, %14, 0;\n\t"
      "addc.cc.u32 %6, %15, 0;\n\t"
      "addc.cc.u32 %7, %16, 0;\n\t"
      "addc.u32    %8, 0, 0;\n\t"
      "}"
      : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]),
        "=r"(r.v[7]), "=r"(c2)
      : "r"(R[0]), "r"(R[1]), "r"(R[2]), "r"(R[3]), "r"(R[4]), "r"(R[5]), "r"(R[6]), "r"(R[7]), "r"(k));
  r.v[0] += 38u * c2;  // after a wrap the value is below 1444: no further carry
  return r;
}

__device__ __forceinline__ fe8 fe8_mul(const fe8 &a, const fe8 &b) {
  // E[k] sits at limb position k, O[k] at position k + 1: products whose low limb position is
  // even accumulate in E, the others in O
  uint32_t E[16], O[16];
#pragma unroll
  for (int i = 0; i < 16; i++) E[i] = O[i] = 0;
#pragma unroll
  for (int i = 0; i < 8; i += 2) {
    // b_i, i even: a_even * b_i at even positions i + 2j, a_odd * b_i at odd positions i + 2j + 1
    fe8_chain(E + i, E[i + 8], a.v[0], a.v[2], a.v[4], a.v[6], b.v[i]);
    fe8_chain(O + i, O[i + 8], a.v[1], a.v[3], a.v[5], a.v[7], b.v[i]);
    // b_{i+1}: a_even * b at odd positions i + 1 + 2j, a_odd * b at even positions i + 2 + 2j
    fe8_chain(O + i, O[i + 8], a.v[0], a.v[2], a.v[4], a.v[6], b.v[i + 1]);
    if (i + 2 < 8) fe8_chain(E + i + 2, E[i + 10], a.v[1], a.v[3], a.v[5], a.v[7], b.v[i + 1]);
    else fe8_chain_last(E + i + 2, a.v[1], a.v[3], a.v[5], a.v[7], b.v[i + 1]);
  }
  // T = E + (O << 32)
  uint32_t T[16];
  T[0] = E[0];
  asm("{\n\t"
      "add.cc.u32  %0,  %15, %30;\n\t"
      "addc.cc.u32 %1,  %16, %31;\n\t"
      "addc.cc.u32 %2,  %17, %32;\n\t"
      "addc.cc.u32 %3,  %18, %33;\n\t"
      "addc.cc.u32 %4,  %19, %34;\n\t"
      "addc.cc.u32 %5,  %20, %35;\n\t"
      "addc.cc.u32 %6,  %21, %36;\n\t"
      "addc.cc.u32 %7,  %22, %37;\n\t"
      "addc.cc.u32 %8,  %23, %38;\n\t"
      "addc.cc.u32 %9,  %24, %39;\n\t"
      "addc.cc.u32 %10, %25, %40;\n\t"
      "addc.cc.u32 %11, %26, %41;\n\t"
      "addc.cc.u32 %12, %27, %42;\n\t"
      "addc.cc.u32 %13, %28, %43;\n\t"
      "addc.u32    %14, %29, %44;\n\t"
      "}"
      : "=r"(T[1]), "=r"(T[2]), "=r"(T[3]), "=r"(T[4]), "=r"(T[5]), "=r"(T[6]), "=r"(T[7]), "=r"(T[8]), "=r"(T[9]),
        "=r"(T[10]), "=r"(T[11]), "=r"(T[12]), "=r"(T[13]), "=r"(T[14]), "=r"(T[15])
      : "r"(E[1]), "r"(E[2]), "r"(E[3]), "r"(E[4]), "r"(E[5]), "r"(E[6]), "r"(E[7]), "r"(E[8]), "r"(E[9]), "r"(E[10]),
        "r"(E[11]), "r"(E[12]), "r"(E[13]), "r"(E[14]), "r"(E[15]), "r"(O[0]), "r"(O[1]), "r"(O[2]), "r"(O[3]),
        "r"(O[4]), "r"(O[5]), "r"(O[6]), "r"(O[7]), "r"(O[8]), "r"(O[9]), "r"(O[10]), "r"(O[11]), "r"(O[12]),
        "r"(O[13]), "r"(O[14]));
  return fe8_fold(T);
}

__device__ __forceinline__ fe8 fe8_add(const fe8 &a, const fe8 &b) {
  fe8 r;
  uint32_t c;
  asm("{\n\t"
      "add.cc.u32  %0, %9,  %17;\n\t"
      "addc.cc.u32 %1, %10, %18;\n\t"
      "addc.cc.u32 %2, %11, %19;\n\t"
      "addc.cc.u32 %3, %12, %20;\n\t"
      "addc.cc.u32 %4, %13, %21;\n\t"
      "addc.cc.u32 %5, %14, %22;\n\t"
      "addc.cc.u32 %6, %15, %23;\n\t"
      "addc.cc.u32 %7, %16, %24;\n\t"
      "addc.u32    %8, 0, 0;\n\t"
      "}"
      : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]),
        "=r"(r.v[7]), "=r"(c)
      : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]),
        "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]), "r"(b.v[6]), "r"(b.v[7]));
  uint32_t k = 38u * c, c2;
  asm("{\n\t"
      "add.cc.u32  %0, %0, %9;\n\t"
      "addc.cc.u32 %1, %1, 0;\n\t"
      "addc.cc.u32 %2, %2, 0;\n\t"
      "addc.cc.u32 %3, %3, 0;\n\t"
      "addc.cc.u32 %4, %4, 0;\n\t"
      "addc.cc.u32 %5, %5, 0;\n\t"
      "addc.cc.u32 %6, %6, 0;\n\t"
      "addc.cc.u32 %7, %7, 0;\n\t"
      "addc.u32    %8, 0, 0;\n\t"
      "}"
      : "+r"(r.v[0]), "+r"(r.v[1]), "+r"(r.v[2]), "+r"(r.v[3]), "+r"(r.v[4]), "+r"(r.v[5]), "+r"(r.v[6]),
        "+r"(r.v[7]), "=r"(c2)
      : "r"(k));
  r.v[0] += 38u * c2;
  return r;
}

__device__ __forceinline__ fe8 fe8_sub(const fe8 &a, const fe8 &b) {
  fe8 r;
  uint32_t bw;
  asm("{\n\t"
      "sub.cc.u32  %0, %9,  %17;\n\t"
      "subc.cc.u32 %1, %10, %18;\n\t"
      "subc.cc.u32 %2, %11, %19;\n\t"
      "subc.cc.u32 %3, %12, %20;\n\t"
      "subc.cc.u32 %4, %13, %21;\n\t"
      "subc.cc.u32 %5, %14, %22;\n\t"
      "subc.cc.u32 %6, %15, %23;\n\t"
      "subc.cc.u32 %7, %16, %24;\n\t"
      "subc.u32    %8, 0, 0;\n\t"
      "}"
      : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]),
        "=r"(r.v[7]), "=r"(bw)
      : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]),
        "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]), "r"(b.v[6]), "r"(b.v[7]));
  // bw is 0 or 0xffffffff: a borrow means the stored value is 2^256 too large, i.e. 38 too large
  uint32_t k = 38u & bw, bw2;
  asm("{\n\t"
      "sub.cc.u32  %0, %0, %9;\n\t"
      "subc.cc.u32 %1, %1, 0;\n\t"
      "subc.cc.u32 %2, %2, 0;\n\t"
      "subc.cc.u32 %3, %3, 0;\n\t"
      "subc.cc.u32 %4, %4, 0;\n\t"
      "subc.cc.u32 %5, %5, 0;\n\t"
      "subc.cc.u32 %6, %6, 0;\n\t"
      "subc.cc.u32 %7, %7, 0;\n\t"
      "subc.u32    %8, 0, 0;\n\t"
      "}"
      : "+r"(r.v[0]), "+r"(r.v[1]), "+r"(r.v[2]), "+r"(r.v[3]), "+r"(r.v[4]), "+r"(r.v[5]), "+r"(r.v[6]),
        "+r"(r.v[7]), "=r"(bw2)
      : "r"(k));
  r.v[0] -= 38u & bw2;  // after the second wrap the value is >= 2^256 - 38: no further borrow
  return r;
}

__device__ __forceinline__ fe8 fe8_select(bool c, const fe8 &a, const fe8 &b) {
  fe8 r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = c ? a.v[i] : b.v[i];
  return r;
}

// ---------------------------------------------------------------- conversions (not hot)
__device__ inline fe8 fe8_from_fe(const fe &f) {
  uint8_t b[32];
  fe_tobytes(f, b);
  fe8 r;
  for (int i = 0; i < 8; i++)
    r.v[i] = (uint32_t)b[4 * i] | ((uint32_t)b[4 * i + 1] << 8) | ((uint32_t)b[4 * i + 2] << 16) | ((uint32_t)b[4 * i + 3] << 24);
  return r;
}
__device__ inline fe fe8_to_fe(const fe8 &a) {
  // v = lo255 + 2^255 * top  ->  lo255 + 19 * top (< 2^255 + 19); fe_frombytes takes 255 bits,
  // so fold once more if that sum reaches 2^255 (it then is below 19 + 19)
  uint32_t t[8];
  for (int i = 0; i < 8; i++) t[i] = a.v[i];
  for (int pass = 0; pass < 2; pass++) {
    uint64_t c = 19ull * (t[7] >> 31);
    t[7] &= 0x7fffffffu;
    for (int i = 0; i < 8; i++) {
      c += t[i];
      t[i] = (uint32_t)c;
      c >>= 32;
    }
  }
  uint8_t b[32];
  for (int i = 0; i < 8; i++)
    for (int k = 0; k < 4; k++) b[4 * i + k] = (uint8_t)(t[i] >> (8 * k));
  return fe_frombytes(b);  // fe_tobytes canonicalises the [p, 2^255) corner later
}

// ---------------------------------------------------------------- points
struct ge8 {  // extended coordinates
  fe8 X, Y, Z, T;
};
struct niels8 {  // affine precomputed point (y + x, y - x, 2 d x y): 96 bytes
  fe8 ypx, ymx, t2d;
};

__device__ __forceinline__ ge8 ge8_identity() {
  ge8 r;
  r.X = fe8_zero();
  r.Y = fe8_one();
  r.Z = fe8_one();
  r.T = fe8_zero();
  return r;
}

// mixed addition p + (neg ? -q : q) (madd-2008-hwcd-3 with Z2 = 1): 7 multiplications.
// Negating a precomputed point swaps y + x with y - x and negates 2dxy; the latter is done by
// exchanging the roles of D + C and D - C instead of a field negation.
__device__ __forceinline__ ge8 ge8_madd(const ge8 &p, const niels8 &q, bool neg) {
  fe8 ymx = fe8_select(neg, q.ypx, q.ymx), ypx = fe8_select(neg, q.ymx, q.ypx);
  fe8 A = fe8_mul(fe8_sub(p.Y, p.X), ymx);
  fe8 B = fe8_mul(fe8_add(p.Y, p.X), ypx);
  fe8 C = fe8_mul(p.T, q.t2d);
  fe8 D = fe8_add(p.Z, p.Z);
  fe8 E = fe8_sub(B, A), H = fe8_add(B, A);
  fe8 S = fe8_add(D, C), M = fe8_sub(D, C);
  fe8 F = fe8_select(neg, S, M), G = fe8_select(neg, M, S);
  ge8 r;
  r.X = fe8_mul(E, F);
  r.Y = fe8_mul(G, H);
  r.Z = fe8_mul(F, G);
  r.T = fe8_mul(E, H);
  return r;
}

// 2d mod p on eight limbs (d = -121665/121666)
__device__ __forceinline__ fe8 fe8_2d() {
  fe8 r;
  r.v[0] = 0x26b2f159u; r.v[1] = 0xebd69b94u; r.v[2] = 0x8283b156u; r.v[3] = 0x00e0149au;
  r.v[4] = 0xeef3d130u; r.v[5] = 0x198e80f2u; r.v[6] = 0x56dffce7u; r.v[7] = 0x2406d9dcu;
  return r;
}

// full addition of two extended points (add-2008-hwcd-3): 9 multiplications; only in the
// reductions of per-thread partial sums
__device__ __forceinline__ ge8 ge8_add(const ge8 &p, const ge8 &q) {
  fe8 A = fe8_mul(fe8_sub(p.Y, p.X), fe8_sub(q.Y, q.X));
  fe8 B = fe8_mul(fe8_add(p.Y, p.X), fe8_add(q.Y, q.X));
  fe8 C = fe8_mul(fe8_mul(p.T, q.T), fe8_2d());
  fe8 ZZ = fe8_mul(p.Z, q.Z);
  fe8 D = fe8_add(ZZ, ZZ);
  fe8 E = fe8_sub(B, A), F = fe8_sub(D, C), G = fe8_add(D, C), H = fe8_add(B, A);
  ge8 r;
  r.X = fe8_mul(E, F);
  r.Y = fe8_mul(G, H);
  r.Z = fe8_mul(F, G);
  r.T = fe8_mul(E, H);
  return r;
}

__device__ inline ge ge8_to_ge(const ge8 &p) {
  ge r;
  r.X = fe8_to_fe(p.X);
  r.Y = fe8_to_fe(p.Y);
  r.Z = fe8_to_fe(p.Z);
  r.T = fe8_to_fe(p.T);
  return r;
}

// 96-byte table entries: three 256-bit loads
__device__ __forceinline__ fe8 fe8_load(const fe8 *p) {
  fe8 r;
  asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]),
                 "=r"(r.v[7])
               : "l"(p));
  return r;
}
__device__ __forceinline__ niels8 niels8_load(const niels8 *p) {
  niels8 r;
  r.ypx = fe8_load(&p->ypx);
  r.ymx = fe8_load(&p->ymx);
  r.t2d = fe8_load(&p->t2d);
  return r;
}

}  // namespace spg
