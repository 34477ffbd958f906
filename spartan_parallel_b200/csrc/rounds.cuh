// Round kernels shared by the sumcheck provers.
//
// Device layout (differs from the reference on purpose). The reference stores
// DensePolynomialPqx as nested Vecs in (p, q_rev, w, x_rev) order so that binding
// the LOW bit of x / q becomes a "top" bind (src/custom_dense_mlpoly.rs:67-111,
// 267-289). Here a table is one flat HBM buffer in NATURAL order; binding the low
// bit pairs ADJACENT scalars (2i, 2i+1), so a thread reads 64 contiguous bytes per
// table and the bound table is written densely, halving every round.
//
// A table is a list of segments (one per instance p). A segment has n_rows rows of
// row_len = 2^log_len scalars; the round binds the low bit inside each row. For
// the x rounds a row is one (p, q) pair; for the q rounds (x fully bound) a row is
// one instance. Rows of length 1 are "exhausted" (the instance has fewer
// constraints / proofs than the maximum): their high half is zero
// (index_high, custom_dense_mlpoly.rs:145-165) and binding scales by (1 - r).
//
// The eq polynomial is never bound: for the j-th round of a variable group with
// challenges r_0..r_{j-1} already fixed,
//   eq_bound[i_lo | i_hi] = (prod_{k<j} eq(tau_k, r_k)) * eq(tau_j, bit) * S_{j+1}[i]
// with S_{j+1} the eq table of the remaining taus. The kernels accumulate
//   sum_i RW[row] * S_{j+1}[i] * F_i(t),  t in {0, 2, 3}
// and the host multiplies by the scalar prefix and by l_j(t) = eq(tau_j, t).
#pragma once
#include "common.cuh"

namespace spg {

struct Seg {
  unsigned long long in_off;      // first scalar of the segment in the input tables
  unsigned long long out_off;     // first scalar of the segment in the bound tables
  unsigned long long item_start;  // prefix sum of work items before this segment
  unsigned int log_len;           // log2(row_len)
  unsigned int n_rows;
  unsigned int rw_off;            // first row weight of the segment
  unsigned int log_tiles;         // tile kernels: log2(tiles per row); item_start counts tiles
};

__device__ __forceinline__ int find_seg(const Seg *__restrict__ segs, int nseg, unsigned long long item) {
  int lo = 0, hi = nseg - 1;
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    if (segs[mid].item_start <= item) lo = mid;
    else hi = mid - 1;
  }
  return lo;
}

// Up to SEG_INLINE segments travel as a kernel argument (constant bank), so the per-round
// segment table costs no host-to-device copy; larger instance counts use the device array.
constexpr int SEG_INLINE = 16;
struct SegPack {
  Seg s[SEG_INLINE];
};

__device__ __forceinline__ Seg pick_seg(const SegPack &pk, const Seg *__restrict__ segs, int nseg,
                                        unsigned long long item) {
  if (nseg == 1) return pk.s[0];
  if (nseg > SEG_INLINE) return segs[find_seg(segs, nseg, item)];
  int lo = 0, hi = nseg - 1;
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    if (pk.s[mid].item_start <= item) lo = mid;
    else hi = mid - 1;
  }
  return pk.s[lo];
}

static inline SegPack make_pack(const std::vector<Seg> &v) {
  SegPack pk;
  memset(&pk, 0, sizeof pk);
  for (size_t i = 0; i < v.size() && i < (size_t)SEG_INLINE; i++) pk.s[i] = v[i];
  return pk;
}

// value of the line through (0, lo), (1, hi) at 2 and 3
__device__ __forceinline__ void line23(const fq &lo, const fq &hi, fq &at2, fq &at3) {
  fq d = fq_sub(hi, lo);
  at2 = fq_add(hi, d);
  at3 = fq_add(at2, d);
}

// COMB 1: T0*T1 - T2   (phase 1, src/r1csproof.rs:100-104 without the eq factor)
// COMB 2: T0*T1        (phase 2 / product circuits: B*C, the A factor is the weight)
template <int COMB>
__device__ __forceinline__ void comb_accumulate(fq (&acc)[3], const fq &w, const fq &a0, const fq &a1,
                                                const fq &b0, const fq &b1, const fq &c0, const fq &c1) {
  fq a2, a3, b2, b3;
  line23(a0, a1, a2, a3);
  line23(b0, b1, b2, b3);
  fq f0 = fq_mul(a0, b0), f2 = fq_mul(a2, b2), f3 = fq_mul(a3, b3);
  if (COMB == 1) {
    fq c2, c3;
    line23(c0, c1, c2, c3);
    f0 = fq_sub(f0, c0);
    f2 = fq_sub(f2, c2);
    f3 = fq_sub(f3, c3);
  }
  acc[0] = fq_add(acc[0], fq_mul(w, f0));
  acc[1] = fq_add(acc[1], fq_mul(w, f2));
  acc[2] = fq_add(acc[2], fq_mul(w, f3));
}

}  // namespace spg
