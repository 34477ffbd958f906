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

// ---------------------------------------------------------------- late rounds: one item over eight lanes
// A late round (tables of a few thousand scalars or less) is not throughput but a dependent chain: with
// one item per thread, k_quad_bind_eval / k2_quad_bind_eval execute ~6700 dependent instructions (the
// binds one after the other, three evaluation points, three accumulators reduced one after the other,
// the in-kernel final reduction again) on one warp per scheduler, ~8 cycles each: 20 - 27 us at ONE block
// (profiles/r2_ncu_launches.csv). Nothing in an item is sequential except bind -> evaluate, so here an
// item is spread over eight lanes and every stage runs once:
//   lane 0 .. 2 NT - 1   bind of (table l / 2, pair l % 2): z + x * y with (x, y, z) = (r, hi - lo, lo)
//   lane 6               the item's weight, the same z + x * y with (RW[row], S[i], 0) or (A[p], 1, 0)
//   lanes 0, 1, 2        after an exchange by shuffles: evaluation at t = 0, 2, 3 (the point is a
//                        select of the addend -d, d, 2d on top of x1 -- no divergent branch)
// and the three sums live in DIFFERENT lanes, so one butterfly reduces all three (the final reduction over
// the blocks' partial sums is laid out the same way). ~1500 instructions end to end. Same values as the
// per-thread kernels: exact arithmetic, canonical results.
struct SplitTabs {
  const fq *in[3];
  fq *out[3];
};
constexpr int SPLIT_ITEMS_PER_BLOCK = 16;        // 128 threads / 8 lanes
constexpr unsigned long long SPLIT_MAX_ITEMS = 4096;  // up to 256 blocks; beyond, one item per thread wins

__device__ __forceinline__ fq fq_shfl(const fq &x, int src) {
  fq y;
#pragma unroll
  for (int i = 0; i < 8; i++) y.v[i] = __shfl_sync(0xffffffffu, x.v[i], src);
  return y;
}
__device__ __forceinline__ fq fq_shfl_down(const fq &x, int off) {
  fq y;
#pragma unroll
  for (int i = 0; i < 8; i++) y.v[i] = __shfl_down_sync(0xffffffffu, x.v[i], off);
  return y;
}

// x0 + t (x1 - x0) at t = 0, 2, 3 for lane role l = 0, 1, 2: x1 + {x0 - x1, d, 2d}
__device__ __forceinline__ fq split_point(const fq &x0, const fq &x1, unsigned int l) {
  fq d = fq_sub(x1, x0), nd = fq_sub(x0, x1), d2 = fq_add(d, d), s;
#pragma unroll
  for (int i = 0; i < 8; i++) s.v[i] = l == 0 ? nd.v[i] : (l == 1 ? d.v[i] : d2.v[i]);
  return fq_add(x1, s);
}

// Final reduction over the blocks for kernels whose three sums are reduced side by side: `mine` (valid in
// thread 0) = this block's sums; single block: published directly; else the block that draws the last
// ticket adds the partial rows with thread t on sum t % 4 and one butterfly for all three (finish_block,
// common.cuh, does the same with three block_sums one after the other). 128 threads, sm = fq[16].
__device__ __forceinline__ void finish_block_lanes3(const FinishArgs &fa, const fq (&mine)[3], fq *sm) {
  __shared__ int is_last;
  const unsigned int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (gridDim.x == 1 && fa.seq) {
    if (threadIdx.x == 0) {
#pragma unroll
      for (int k = 0; k < 3; k++) fq_store(fa.result + k, mine[k]);
      __threadfence_system();
      *(volatile unsigned long long *)fa.flag = fa.seq;
    }
    return;
  }
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < 3; k++) fq_store(fa.partials + (size_t)blockIdx.x * 3 + k, mine[k]);
    int last = 0;
    if (fa.seq) {
      __threadfence();
      last = atomicAdd(fa.counter, 1u) == gridDim.x - 1;
    }
    is_last = last;
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  const unsigned int k4 = threadIdx.x & 3;
  fq a = fq_zero();
  if (k4 < 3)
    for (unsigned int b = threadIdx.x >> 2; b < gridDim.x; b += 32) a = fq_add(a, fq_load_cg(fa.partials + (size_t)b * 3 + k4));
  a = fq_add(a, fq_shfl_down(a, 4));
  a = fq_add(a, fq_shfl_down(a, 8));
  a = fq_add(a, fq_shfl_down(a, 16));
  if (lane < 3) sm[warp * 4 + lane] = a;  // (the caller's use of sm was consumed before the barrier above)
  __syncthreads();
  if (warp == 0) {
    fq t = (lane < 16 && (lane & 3) < 3) ? sm[(lane >> 2) * 4 + (lane & 3)] : fq_zero();
    t = fq_add(t, fq_shfl_down(t, 4));
    t = fq_add(t, fq_shfl_down(t, 8));
    fq s1 = fq_shfl(t, 1), s2 = fq_shfl(t, 2);
    if (lane == 0) {
      fq_store(fa.result + 0, t);
      fq_store(fa.result + 1, s1);
      fq_store(fa.result + 2, s2);
      *fa.counter = 0;
      __threadfence_system();
      *(volatile unsigned long long *)fa.flag = fa.seq;
    }
  }
}

template <int NT, int COMB>
__global__ void __launch_bounds__(128)
k_quad_split(const __grid_constant__ SplitTabs T, const Seg *__restrict__ segs, int nseg, const __grid_constant__ SegPack pk,
             unsigned long long total_items, fq r, const fq *__restrict__ RW, const fq *__restrict__ Snext, FinishArgs fa) {
  static_assert((NT == 3 && COMB == 1) || (NT == 2 && COMB == 2), "phase 1: three tables, A B - C; phase 2: two tables, B C");
  __shared__ fq sm[4 * 4];  // [warp][point]
  const unsigned int lane = threadIdx.x & 31, l = lane & 7, warp = threadIdx.x >> 5;
  const unsigned long long item = (unsigned long long)blockIdx.x * SPLIT_ITEMS_PER_BLOCK + (threadIdx.x >> 3);
  const bool valid = item < total_items;
  const Seg sg = pick_seg(pk, segs, nseg, valid ? item : 0);
  const unsigned long long local = valid ? item - sg.item_start : 0;
  const unsigned long long idx = sg.in_off + 4 * local, o = sg.out_off + 2 * local;
  const bool is_bind = l < 2 * NT, is_weight = l == 6;
  const unsigned int tab = is_bind ? l >> 1 : 0, pair = l & 1;
  // operands of z + x * y
  fq x = fq_zero(), y = fq_zero(), z = fq_zero();
  if (valid && is_bind) {
    const fq *p = T.in[tab] + idx + 2 * pair;
    z = fq_load_stream(p);
    y = fq_sub(fq_load_stream(p + 1), z);
    x = r;
  } else if (valid && is_weight) {
    if (Snext) {  // phase 1: row weight times the suffix eq table of the remaining variables
      unsigned int ql = sg.log_len - 2;
      x = fq_load(RW + sg.rw_off + (local >> ql));
      y = fq_load(Snext + (local & ((1ull << ql) - 1)));
    } else {      // phase 2: the instance's eq_p weight
      x = fq_load(RW + sg.rw_off);
      y = fq_one();
    }
  }
  const fq v = fq_add(z, fq_mul(x, y));
  if (valid && is_bind) fq_store(T.out[tab] + o + pair, v);
  // every lane of the group receives the bound values and the weight
  const int base = lane & ~7u;
  fq g[2 * NT];
#pragma unroll
  for (int k = 0; k < 2 * NT; k++) g[k] = fq_shfl(v, base + k);
  const fq w = fq_shfl(v, base + 6);
  fq f;
  if (COMB == 1) {
    fq a = split_point(g[0], g[1], l), b = split_point(g[2], g[3], l), c = split_point(g[4], g[5], l);
    f = fq_sub(fq_mul(a, b), c);
  } else {
    fq b = split_point(g[0], g[1], l), c = split_point(g[2], g[3], l);
    f = fq_mul(b, c);
  }
  fq acc = fq_mul(w, f);
  if (!(valid && l < 3)) acc = fq_zero();
  // lanes 8 j + k hold point k of item j: two butterfly steps leave the warp's sums in lanes 0 .. 2
  acc = fq_add(acc, fq_shfl_down(acc, 8));
  acc = fq_add(acc, fq_shfl_down(acc, 16));
  if (lane < 3) sm[warp * 4 + lane] = acc;
  __syncthreads();
  fq mine[3] = {fq_zero(), fq_zero(), fq_zero()};
  if (warp == 0) {
    fq t = l < 3 ? sm[(lane >> 3) * 4 + l] : fq_zero();
    t = fq_add(t, fq_shfl_down(t, 8));
    t = fq_add(t, fq_shfl_down(t, 16));
    mine[0] = t;
    mine[1] = fq_shfl(t, 1);
    mine[2] = fq_shfl(t, 2);
  }
  finish_block_lanes3(fa, mine, sm);
}

}  // namespace spg
