// Device arithmetic in the Curve25519 scalar field F_q,
//   q = 2^252 + 27742317777372353535851937790883648493,
// in exactly the representation of the reference's `Scalar`
// (/root/reference/src/scalar/ristretto255.rs:193-199): a*R mod q with R = 2^256,
// four little-endian u64 limbs, always fully reduced at kernel boundaries.
// On the device the same 32 bytes are viewed as eight u32 limbs so that every
// product is one IMAD.WIDE.U32 (mad.lo.cc / madc.hi.cc pairs fused by ptxas).
//
// Replaces: Scalar::{add,sub,neg,mul,square,montgomery_reduce}
//           (ristretto255.rs:641-763). Any algorithm that returns the canonical
// representative of a*b*R^-1 mod q is bit-identical to the reference, so the
// reduction below is free to exploit q = 2^252 + c (limbs 4..6 of q are zero,
// limb 7 is 2^28): a CIOS Montgomery row costs 8 wide multiplies for a*b_i and
// only 5 for k*q.
//
// Two value ranges are used:
//   canonical  [0, q)   -- what is stored in HBM and crosses the C ABI
//   lazy       [0, 2q)  -- inside a kernel; fq_mul_lazy accepts and returns it
//                          (inputs < 2q give outputs < 1.26 q), which removes the
//                          conditional subtraction from the inner loops.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace spg {

struct __align__(32) fq {
  uint32_t v[8];
};

// q, 2q and the Montgomery constants as 32-bit limbs.
#define SPG_Q0 0x5cf5d3edu
#define SPG_Q1 0x5812631au
#define SPG_Q2 0xa2f79cd6u
#define SPG_Q3 0x14def9deu
#define SPG_Q7 0x10000000u
#define SPG_2Q0 0xb9eba7dau
#define SPG_2Q1 0xb024c634u
#define SPG_2Q2 0x45ef39acu
#define SPG_2Q3 0x29bdf3bdu
#define SPG_2Q7 0x20000000u
#define SPG_INV32 0x12547e1bu  // -q^-1 mod 2^32 (low half of INV, ristretto255.rs:304)

__device__ __forceinline__ fq fq_zero() {
  fq r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = 0;
  return r;
}

// R mod q = Scalar::one() (ristretto255.rs:307-312)
__device__ __forceinline__ fq fq_one() {
  fq r;
  r.v[0] = 0x8d98951du; r.v[1] = 0xd6ec3174u; r.v[2] = 0x737dcf70u; r.v[3] = 0xc6ef5bf4u;
  r.v[4] = 0xfffffffeu; r.v[5] = 0xffffffffu; r.v[6] = 0xffffffffu; r.v[7] = 0x0fffffffu;
  return r;
}

__device__ __forceinline__ bool fq_is_zero(const fq &a) {
  return (a.v[0] | a.v[1] | a.v[2] | a.v[3] | a.v[4] | a.v[5] | a.v[6] | a.v[7]) == 0;
}

__device__ __forceinline__ bool fq_equal(const fq &a, const fq &b) {
  uint32_t d = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) d |= a.v[i] ^ b.v[i];
  return d == 0;
}

// r = a - m (m given as the sparse limbs m0..m3, m7) if that does not borrow, else a.
// Brings [0, 2m) into [0, m).
__device__ __forceinline__ fq fq_cond_sub(const fq &a, uint32_t m0, uint32_t m1, uint32_t m2,
                                          uint32_t m3, uint32_t m7) {
  fq t;
  uint32_t borrow;
  asm("{\n\t"
      "sub.cc.u32  %0, %9,  %17;\n\t"
      "subc.cc.u32 %1, %10, %18;\n\t"
      "subc.cc.u32 %2, %11, %19;\n\t"
      "subc.cc.u32 %3, %12, %20;\n\t"
      "subc.cc.u32 %4, %13, 0;\n\t"
      "subc.cc.u32 %5, %14, 0;\n\t"
      "subc.cc.u32 %6, %15, 0;\n\t"
      "subc.cc.u32 %7, %16, %21;\n\t"
      "subc.u32    %8, 0, 0;\n\t"
      "}"
      : "=r"(t.v[0]), "=r"(t.v[1]), "=r"(t.v[2]), "=r"(t.v[3]), "=r"(t.v[4]), "=r"(t.v[5]),
        "=r"(t.v[6]), "=r"(t.v[7]), "=r"(borrow)
      : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]),
        "r"(a.v[7]), "r"(m0), "r"(m1), "r"(m2), "r"(m3), "r"(m7));
  fq r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = borrow ? a.v[i] : t.v[i];
  return r;
}

// [0,2q) -> [0,q)
__device__ __forceinline__ fq fq_canon(const fq &a) {
  return fq_cond_sub(a, SPG_Q0, SPG_Q1, SPG_Q2, SPG_Q3, SPG_Q7);
}

__device__ __forceinline__ fq fq_raw_add(const fq &a, const fq &b) {
  fq s;
  asm("{\n\t"
      "add.cc.u32  %0, %8,  %16;\n\t"
      "addc.cc.u32 %1, %9,  %17;\n\t"
      "addc.cc.u32 %2, %10, %18;\n\t"
      "addc.cc.u32 %3, %11, %19;\n\t"
      "addc.cc.u32 %4, %12, %20;\n\t"
      "addc.cc.u32 %5, %13, %21;\n\t"
      "addc.cc.u32 %6, %14, %22;\n\t"
      "addc.u32    %7, %15, %23;\n\t"
      "}"
      : "=r"(s.v[0]), "=r"(s.v[1]), "=r"(s.v[2]), "=r"(s.v[3]), "=r"(s.v[4]), "=r"(s.v[5]),
        "=r"(s.v[6]), "=r"(s.v[7])
      : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]),
        "r"(a.v[7]), "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]),
        "r"(b.v[6]), "r"(b.v[7]));
  return s;
}

// a - b, adding back m (sparse limbs) when the subtraction borrows.
__device__ __forceinline__ fq fq_sub_mod(const fq &a, const fq &b, uint32_t m0, uint32_t m1,
                                         uint32_t m2, uint32_t m3, uint32_t m7) {
  fq d;
  uint32_t mask;
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
      : "=r"(d.v[0]), "=r"(d.v[1]), "=r"(d.v[2]), "=r"(d.v[3]), "=r"(d.v[4]), "=r"(d.v[5]),
        "=r"(d.v[6]), "=r"(d.v[7]), "=r"(mask)
      : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]),
        "r"(a.v[7]), "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]),
        "r"(b.v[6]), "r"(b.v[7]));
  asm("{\n\t"
      "add.cc.u32  %0, %0, %8;\n\t"
      "addc.cc.u32 %1, %1, %9;\n\t"
      "addc.cc.u32 %2, %2, %10;\n\t"
      "addc.cc.u32 %3, %3, %11;\n\t"
      "addc.cc.u32 %4, %4, 0;\n\t"
      "addc.cc.u32 %5, %5, 0;\n\t"
      "addc.cc.u32 %6, %6, 0;\n\t"
      "addc.u32    %7, %7, %12;\n\t"
      "}"
      : "+r"(d.v[0]), "+r"(d.v[1]), "+r"(d.v[2]), "+r"(d.v[3]), "+r"(d.v[4]), "+r"(d.v[5]),
        "+r"(d.v[6]), "+r"(d.v[7])
      : "r"(m0 & mask), "r"(m1 & mask), "r"(m2 & mask), "r"(m3 & mask), "r"(m7 & mask));
  return d;
}

// canonical in, canonical out (Scalar::add / Scalar::sub / Scalar::neg)
__device__ __forceinline__ fq fq_add(const fq &a, const fq &b) { return fq_canon(fq_raw_add(a, b)); }
__device__ __forceinline__ fq fq_sub(const fq &a, const fq &b) {
  return fq_sub_mod(a, b, SPG_Q0, SPG_Q1, SPG_Q2, SPG_Q3, SPG_Q7);
}
__device__ __forceinline__ fq fq_neg(const fq &a) { return fq_sub(fq_zero(), a); }

// lazy [0,2q) in, lazy [0,2q) out
__device__ __forceinline__ fq fq_add_lazy(const fq &a, const fq &b) {
  return fq_cond_sub(fq_raw_add(a, b), SPG_2Q0, SPG_2Q1, SPG_2Q2, SPG_2Q3, SPG_2Q7);
}
__device__ __forceinline__ fq fq_sub_lazy(const fq &a, const fq &b) {
  return fq_sub_mod(a, b, SPG_2Q0, SPG_2Q1, SPG_2Q2, SPG_2Q3, SPG_2Q7);
}

// Wide-range forms for the inner loops of the row kernels. A 256-bit word holds values up to
// 15.99 q, and fq_mul_lazy(a, b) only needs a + q < 2^256 (its running value stays below
// (a + q) 2^32 + 2^256 for ANY 256-bit b) and returns less than a*b/R + q, so differences can be made
// non-negative by adding a constant multiple of q instead of a conditional correction (a borrow
// mask, five LOP3 and a second carry chain per subtraction in fq_sub_lazy). Each use states its
// range. None of this changes a result: everything is exact arithmetic mod q, canonicalised before it
// leaves the kernel (or, for tables kept in [0, 2q) between two launches, by the last of them).
#define SPG_6Q0 0x2dc2f78eu
#define SPG_6Q1 0x106e529eu
#define SPG_6Q2 0xd1cdad06u
#define SPG_6Q3 0x7d39db37u
#define SPG_6Q7 0x60000000u

// a - b + k q for k q given by its sparse limbs (limbs 4..6 of k q are zero for k <= 15); needs b <= k q
__device__ __forceinline__ fq fq_sub_plus(const fq &a, const fq &b, uint32_t m0, uint32_t m1, uint32_t m2,
                                          uint32_t m3, uint32_t m7) {
  fq d;
  asm("{\n\t"
      "sub.cc.u32  %0, %8,  %16;\n\t"
      "subc.cc.u32 %1, %9,  %17;\n\t"
      "subc.cc.u32 %2, %10, %18;\n\t"
      "subc.cc.u32 %3, %11, %19;\n\t"
      "subc.cc.u32 %4, %12, %20;\n\t"
      "subc.cc.u32 %5, %13, %21;\n\t"
      "subc.cc.u32 %6, %14, %22;\n\t"
      "subc.u32    %7, %15, %23;\n\t"
      "}"
      : "=r"(d.v[0]), "=r"(d.v[1]), "=r"(d.v[2]), "=r"(d.v[3]), "=r"(d.v[4]), "=r"(d.v[5]), "=r"(d.v[6]), "=r"(d.v[7])
      : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]),
        "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]), "r"(b.v[6]), "r"(b.v[7]));
  asm("{\n\t"
      "add.cc.u32  %0, %0, %8;\n\t"
      "addc.cc.u32 %1, %1, %9;\n\t"
      "addc.cc.u32 %2, %2, %10;\n\t"
      "addc.cc.u32 %3, %3, %11;\n\t"
      "addc.cc.u32 %4, %4, 0;\n\t"
      "addc.cc.u32 %5, %5, 0;\n\t"
      "addc.cc.u32 %6, %6, 0;\n\t"
      "addc.u32    %7, %7, %12;\n\t"
      "}"
      : "+r"(d.v[0]), "+r"(d.v[1]), "+r"(d.v[2]), "+r"(d.v[3]), "+r"(d.v[4]), "+r"(d.v[5]), "+r"(d.v[6]), "+r"(d.v[7])
      : "r"(m0), "r"(m1), "r"(m2), "r"(m3), "r"(m7));
  return d;
}
// a - b + 2q: a in [0, 14q), b in [0, 2q]
__device__ __forceinline__ fq fq_sub_plus2q(const fq &a, const fq &b) {
  return fq_sub_plus(a, b, SPG_2Q0, SPG_2Q1, SPG_2Q2, SPG_2Q3, SPG_2Q7);
}
// a - b + 6q: a in [0, 9.9q), b in [0, 6q]
__device__ __forceinline__ fq fq_sub_plus6q(const fq &a, const fq &b) {
  return fq_sub_plus(a, b, SPG_6Q0, SPG_6Q1, SPG_6Q2, SPG_6Q3, SPG_6Q7);
}
// [0, 4q) -> [0, 2q)
__device__ __forceinline__ fq fq_fold2q(const fq &a) {
  return fq_cond_sub(a, SPG_2Q0, SPG_2Q1, SPG_2Q2, SPG_2Q3, SPG_2Q7);
}

// ---------------------------------------------------------------------------
// Montgomery multiplication, CIOS over 32-bit limbs with split accumulators.
//
// The running value is T = sum_k E[k] 2^(32k) + sum_k O[k] 2^(32(k+1)): products of
// the even limbs of `a` accumulate in E, products of the odd limbs in O, so every
// lo/hi pair lands on an aligned register pair and ptxas emits one
// IMAD.WIDE.U32(.X) per 32x32 product with no register shuffling. After a row is
// reduced (T divisible by 2^32) the two arrays swap roles instead of shifting.
// T stays below 2^288 (a, b < 2q), so the top word never carries out.

// first row: E = a_even * b0, O = a_odd * b0
__device__ __forceinline__ void fq_row0(uint32_t E[8], uint32_t O[8], const fq &a, uint32_t b0) {
#pragma unroll
  for (int j = 0; j < 4; j++) {
    asm("mul.lo.u32 %0, %2, %3;\n\tmul.hi.u32 %1, %2, %3;"
        : "=r"(E[2 * j]), "=r"(E[2 * j + 1])
        : "r"(a.v[2 * j]), "r"(b0));
    asm("mul.lo.u32 %0, %2, %3;\n\tmul.hi.u32 %1, %2, %3;"
        : "=r"(O[2 * j]), "=r"(O[2 * j + 1])
        : "r"(a.v[2 * j + 1]), "r"(b0));
  }
}

// later rows. On entry X (aligned at position 0, X[0] == 0 after the previous
// reduction) and Y (position 1) hold T; on exit Y is the position-0 array and X
// the position-1 array of T/2^32 + a*bi.
__device__ __forceinline__ void fq_row_mul(uint32_t X[8], uint32_t Y[8], const fq &a, uint32_t bi) {
  // Y[0] += X[1]; the carry rides into the odd chain, which also shifts X down by 64 bits
  asm("{\n\t"
      "add.cc.u32      %8,  %8,  %1;\n\t"
      "madc.lo.cc.u32  %0,  %9,  %13, %2;\n\t"
      "madc.hi.cc.u32  %1,  %9,  %13, %3;\n\t"
      "madc.lo.cc.u32  %2,  %10, %13, %4;\n\t"
      "madc.hi.cc.u32  %3,  %10, %13, %5;\n\t"
      "madc.lo.cc.u32  %4,  %11, %13, %6;\n\t"
      "madc.hi.cc.u32  %5,  %11, %13, %7;\n\t"
      "madc.lo.cc.u32  %6,  %12, %13, 0;\n\t"
      "madc.hi.u32     %7,  %12, %13, 0;\n\t"
      "}"
      : "+r"(X[0]), "+r"(X[1]), "+r"(X[2]), "+r"(X[3]), "+r"(X[4]), "+r"(X[5]), "+r"(X[6]),
        "+r"(X[7]), "+r"(Y[0])
      : "r"(a.v[1]), "r"(a.v[3]), "r"(a.v[5]), "r"(a.v[7]), "r"(bi));
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
      : "+r"(Y[0]), "+r"(Y[1]), "+r"(Y[2]), "+r"(Y[3]), "+r"(Y[4]), "+r"(Y[5]), "+r"(Y[6]),
        "+r"(Y[7]), "+r"(X[7])
      : "r"(a.v[0]), "r"(a.v[2]), "r"(a.v[4]), "r"(a.v[6]), "r"(bi));
}

// T += k*q with k = E[0] * (-q^-1) mod 2^32, E at position 0, O at position 1.
// q's limbs 4..6 are zero: those columns only propagate carries (ALU pipe).
__device__ __forceinline__ void fq_row_red(uint32_t E[8], uint32_t O[8]) {
  uint32_t k = E[0] * SPG_INV32;
  // SPG_Q7_SHIFT: k * q7 = k * 2^28 as two funnel shifts and two carry adds on the ALU pipe instead of a
  // wide multiply on the FMA pipe. Measured (tools/imad_peak, bench): 8 % fewer FMA cycles per product
  // but no faster -- the ALU pipe is the co-limiter of these kernels -- so it stays off.
#ifdef SPG_Q7_SHIFT
  uint32_t klo, khi;
  asm("shf.l.wrap.b32 %0, 0, %2, 28;\n\tshf.r.wrap.b32 %1, %2, 0, 4;" : "=r"(klo), "=r"(khi) : "r"(k));
  asm("{\n\t"
      "mad.lo.cc.u32   %0, %8,  %9,  %0;\n\t"
      "madc.hi.cc.u32  %1, %8,  %9,  %1;\n\t"
      "madc.lo.cc.u32  %2, %8,  %10, %2;\n\t"
      "madc.hi.cc.u32  %3, %8,  %10, %3;\n\t"
      "addc.cc.u32     %4, %4, 0;\n\t"
      "addc.cc.u32     %5, %5, 0;\n\t"
      "addc.cc.u32     %6, %6, %11;\n\t"
      "addc.u32        %7, %7, %12;\n\t"
      "}"
      : "+r"(O[0]), "+r"(O[1]), "+r"(O[2]), "+r"(O[3]), "+r"(O[4]), "+r"(O[5]), "+r"(O[6]),
        "+r"(O[7])
      : "r"(k), "r"(SPG_Q1), "r"(SPG_Q3), "r"(klo), "r"(khi));
#else
  asm("{\n\t"
      "mad.lo.cc.u32   %0, %8,  %9,  %0;\n\t"
      "madc.hi.cc.u32  %1, %8,  %9,  %1;\n\t"
      "madc.lo.cc.u32  %2, %8,  %10, %2;\n\t"
      "madc.hi.cc.u32  %3, %8,  %10, %3;\n\t"
      "addc.cc.u32     %4, %4, 0;\n\t"
      "addc.cc.u32     %5, %5, 0;\n\t"
      "madc.lo.cc.u32  %6, %8,  %11, %6;\n\t"
      "madc.hi.u32     %7, %8,  %11, %7;\n\t"
      "}"
      : "+r"(O[0]), "+r"(O[1]), "+r"(O[2]), "+r"(O[3]), "+r"(O[4]), "+r"(O[5]), "+r"(O[6]),
        "+r"(O[7])
      : "r"(k), "r"(SPG_Q1), "r"(SPG_Q3), "r"(SPG_Q7));
#endif
  asm("{\n\t"
      "mad.lo.cc.u32   %0, %9,  %10, %0;\n\t"
      "madc.hi.cc.u32  %1, %9,  %10, %1;\n\t"
      "madc.lo.cc.u32  %2, %9,  %11, %2;\n\t"
      "madc.hi.cc.u32  %3, %9,  %11, %3;\n\t"
      "addc.cc.u32     %4, %4, 0;\n\t"
      "addc.cc.u32     %5, %5, 0;\n\t"
      "addc.cc.u32     %6, %6, 0;\n\t"
      "addc.cc.u32     %7, %7, 0;\n\t"
      "addc.u32        %8, %8, 0;\n\t"
      "}"
      : "+r"(E[0]), "+r"(E[1]), "+r"(E[2]), "+r"(E[3]), "+r"(E[4]), "+r"(E[5]), "+r"(E[6]),
        "+r"(E[7]), "+r"(O[7])
      : "r"(k), "r"(SPG_Q0), "r"(SPG_Q2));
}

// a, b in [0, 2q)  ->  a*b*R^-1 mod q as a value in [0, 1.26 q), a subset of [0, 2q)
__device__ __forceinline__ fq fq_mul_lazy(const fq &a, const fq &b) {
  uint32_t E[8], O[8];
  fq_row0(E, O, a, b.v[0]);
  fq_row_red(E, O);
#pragma unroll
  for (int i = 1; i < 8; i += 2) {
    fq_row_mul(E, O, a, b.v[i]);  // O is now the position-0 array
    fq_row_red(O, E);
    if (i + 1 < 8) {
      fq_row_mul(O, E, a, b.v[i + 1]);  // and E again
      fq_row_red(E, O);
    }
  }
  // after 8 rows O is at position 0 with O[0] == 0 and E at position 1:
  // result = T / 2^32 = E + (O >> 32)
  fq r;
  asm("{\n\t"
      "add.cc.u32  %0, %8,  %16;\n\t"
      "addc.cc.u32 %1, %9,  %17;\n\t"
      "addc.cc.u32 %2, %10, %18;\n\t"
      "addc.cc.u32 %3, %11, %19;\n\t"
      "addc.cc.u32 %4, %12, %20;\n\t"
      "addc.cc.u32 %5, %13, %21;\n\t"
      "addc.cc.u32 %6, %14, %22;\n\t"
      "addc.u32    %7, %15, 0;\n\t"
      "}"
      : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]),
        "=r"(r.v[6]), "=r"(r.v[7])
      : "r"(E[0]), "r"(E[1]), "r"(E[2]), "r"(E[3]), "r"(E[4]), "r"(E[5]), "r"(E[6]), "r"(E[7]),
        "r"(O[1]), "r"(O[2]), "r"(O[3]), "r"(O[4]), "r"(O[5]), "r"(O[6]), "r"(O[7]));
  return r;
}

// further products of the SAME row: T += a * bi with P0 the position-0 array and P1 the
// position-1 array (no shift). Used by the dot product below.
__device__ __forceinline__ void fq_row_acc(uint32_t P0[8], uint32_t P1[8], const fq &a, uint32_t bi) {
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
      : "+r"(P1[0]), "+r"(P1[1]), "+r"(P1[2]), "+r"(P1[3]), "+r"(P1[4]), "+r"(P1[5]), "+r"(P1[6]),
        "+r"(P1[7])
      : "r"(a.v[1]), "r"(a.v[3]), "r"(a.v[5]), "r"(a.v[7]), "r"(bi));
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
      : "+r"(P0[0]), "+r"(P0[1]), "+r"(P0[2]), "+r"(P0[3]), "+r"(P0[4]), "+r"(P0[5]), "+r"(P0[6]),
        "+r"(P0[7]), "+r"(P1[7])
      : "r"(a.v[0]), "r"(a.v[2]), "r"(a.v[4]), "r"(a.v[6]), "r"(bi));
}

// sum_{k<4} a_k * b_k * R^-1 mod q with ONE reduction per row instead of four: 8*(4*8+5) = 296
// wide multiplies instead of 4*104. Inputs must be CANONICAL (< q < 2^253): then the running
// value stays below 4*2^253*2^32 + 2^257 + 2^285 < 2^288, and the result is
// < 4 q^2 / R + q < 1.26 q, i.e. in the lazy range.
__device__ __forceinline__ fq fq_dot4_lazy(const fq (&a)[4], const fq (&b)[4]) {
  uint32_t E[8], O[8];
  fq_row0(E, O, a[0], b[0].v[0]);
#pragma unroll
  for (int k = 1; k < 4; k++) fq_row_acc(E, O, a[k], b[k].v[0]);
  fq_row_red(E, O);
#pragma unroll
  for (int i = 1; i < 8; i += 2) {
    fq_row_mul(E, O, a[0], b[0].v[i]);  // O is now the position-0 array
#pragma unroll
    for (int k = 1; k < 4; k++) fq_row_acc(O, E, a[k], b[k].v[i]);
    fq_row_red(O, E);
    if (i + 1 < 8) {
      fq_row_mul(O, E, a[0], b[0].v[i + 1]);  // and E again
#pragma unroll
      for (int k = 1; k < 4; k++) fq_row_acc(E, O, a[k], b[k].v[i + 1]);
      fq_row_red(E, O);
    }
  }
  fq r;
  asm("{\n\t"
      "add.cc.u32  %0, %8,  %16;\n\t"
      "addc.cc.u32 %1, %9,  %17;\n\t"
      "addc.cc.u32 %2, %10, %18;\n\t"
      "addc.cc.u32 %3, %11, %19;\n\t"
      "addc.cc.u32 %4, %12, %20;\n\t"
      "addc.cc.u32 %5, %13, %21;\n\t"
      "addc.cc.u32 %6, %14, %22;\n\t"
      "addc.u32    %7, %15, 0;\n\t"
      "}"
      : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]),
        "=r"(r.v[6]), "=r"(r.v[7])
      : "r"(E[0]), "r"(E[1]), "r"(E[2]), "r"(E[3]), "r"(E[4]), "r"(E[5]), "r"(E[6]), "r"(E[7]),
        "r"(O[1]), "r"(O[2]), "r"(O[3]), "r"(O[4]), "r"(O[5]), "r"(O[6]), "r"(O[7]));
  return r;
}

// canonical in, canonical out: Scalar::mul (ristretto255.rs:690-714)
__device__ __forceinline__ fq fq_mul(const fq &a, const fq &b) { return fq_canon(fq_mul_lazy(a, b)); }
__device__ __forceinline__ fq fq_sqr(const fq &a) { return fq_mul(a, a); }

// Scalar::invert (ristretto255.rs:541-595) as a^(q-2) by square-and-multiply over the bits of
// q - 2 (any exponentiation schedule gives the same canonical value); 0 maps to 0
// (the reference returns CtOption::none there).
__device__ __forceinline__ fq fq_invert(const fq &a) {
  // q - 2, little-endian 32-bit words
  const uint32_t e[8] = {SPG_Q0 - 2u, SPG_Q1, SPG_Q2, SPG_Q3, 0u, 0u, 0u, SPG_Q7};
  fq acc = fq_one();
#pragma unroll 1
  for (int i = 252; i >= 0; i--) {
    acc = fq_mul_lazy(acc, acc);
    if ((e[i >> 5] >> (i & 31)) & 1u) acc = fq_mul_lazy(acc, a);
  }
  return fq_canon(acc);
}

// to_bytes(): leave Montgomery form (ristretto255.rs:419-431)
__device__ __forceinline__ fq fq_from_mont(const fq &a) {
  fq one = fq_zero();
  one.v[0] = 1;
  return fq_mul(a, one);
}

// ---------------------------------------------------------------------------
// vectorised global memory access: one LDG.256 / STG.256 per scalar
__device__ __forceinline__ fq fq_load(const fq *p) {
  fq r;
  asm volatile("ld.global.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]),
                 "=r"(r.v[6]), "=r"(r.v[7])
               : "l"(p));
  return r;
}
// read-only, streaming (evict-first) variant for tables touched once per round
__device__ __forceinline__ fq fq_load_stream(const fq *p) {
  fq r;
  asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]),
                 "=r"(r.v[6]), "=r"(r.v[7])
               : "l"(p));
  return r;
}
// L1-bypassing load for data written by other blocks of the same kernel
__device__ __forceinline__ fq fq_load_cg(const fq *p) {
  fq r;
  asm volatile("ld.global.cg.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]),
                 "=r"(r.v[6]), "=r"(r.v[7])
               : "l"(p)
               : "memory");
  return r;
}
__device__ __forceinline__ void fq_store(fq *p, const fq &a) {
  asm volatile("st.global.v8.u32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(a.v[0]),
               "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]),
               "r"(a.v[7])
               : "memory");
}

// streaming store (evict-first): tables written once and read back only after far more than
// an L2's worth of other traffic, so they should not displace data that is re-read (CSR arrays)
__device__ __forceinline__ void fq_store_stream(fq *p, const fq &a) {
  asm volatile("st.global.cs.v8.u32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(a.v[0]),
               "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]),
               "r"(a.v[7])
               : "memory");
}

// warp-wide modular sum of canonical values; result valid in lane 0
__device__ __forceinline__ fq fq_warp_sum(fq x) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    fq y;
#pragma unroll
    for (int i = 0; i < 8; i++) y.v[i] = __shfl_down_sync(0xffffffffu, x.v[i], off);
    x = fq_add(x, y);
  }
  return x;
}

}  // namespace spg
