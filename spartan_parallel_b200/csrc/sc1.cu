// Phase-1 sumcheck of R1CSProof::prove:
//   ZKSumcheckInstanceProof::prove_cubic_with_additive_term_disjoint_rounds
//   (/root/reference/src/sumcheck.rs:1067-1380), comb = A*(B*C - D)
//   over DensePolynomialPqx tables (/root/reference/src/custom_dense_mlpoly.rs).
// The transcript / ZK glue of that function stays on the host; this file owns the
// two hot loops: round evaluation (:1166-1245) and binding (:1265-1275).
// See rounds.cuh for the device layout and the eq factorisation.
#include <chrono>
#include <cstdlib>

#include "rounds.cuh"
#include "r1cs.cuh"

namespace spg {

int eq_evals_device(spg_ctx *ctx, const fq *d_r, const spg_fq *h_r, size_t ell, fq *out, fq *scratch);
__global__ void k_eq_expand(const fq *__restrict__ prev, fq *__restrict__ out, size_t n, fq r);
int eq_expand_steps(spg_ctx *ctx, const fq *prev, size_t n, fq *out, unsigned int m0, const spg_fq *r, int steps, bool all);

// 128-thread blocks capped at 128 registers: four blocks (16 warps) per SM. Measured on
// B200 for the 2^20 x 64 batch: 256 threads x 1 block 11.5 ms, 128 x 3 (160 regs) 10.16 ms,
// 128 x 4 (128 regs, no spills) 9.88 ms, 64 x 6 10.25 ms.
#ifndef SPG_RB
#define SPG_RB 128
#endif
#ifndef SPG_MINB
#define SPG_MINB 4
#endif
constexpr int RB = SPG_RB;  // threads per block of the round kernels

// ---------------------------------------------------------------- round evaluation
// One work item = one (lo, hi) pair of adjacent scalars (or a single scalar of an
// exhausted row). 192 B read per item, 7 modmul.
template <int COMB>
__global__ void __launch_bounds__(RB, SPG_MINB)
k_pair_eval(const fq *__restrict__ T0, const fq *__restrict__ T1, const fq *__restrict__ T2,
            const Seg *__restrict__ segs, int nseg, const __grid_constant__ SegPack pk, unsigned long long total_items,
            const fq *__restrict__ RW, const fq *__restrict__ S, fq *__restrict__ partials) {
  __shared__ fq sm[3 * 32];
  fq acc[3] = {fq_zero(), fq_zero(), fq_zero()};
  for (unsigned long long item = (unsigned long long)blockIdx.x * RB + threadIdx.x; item < total_items;
       item += (unsigned long long)gridDim.x * RB) {
    Seg sg = pick_seg(pk, segs, nseg, item);
    unsigned long long local = item - sg.item_start;
    fq a0, a1, b0, b1, c0, c1, w;
    if (sg.log_len >= 1) {
      unsigned int hl = sg.log_len - 1;
      unsigned long long row = local >> hl, i = local & ((1ull << hl) - 1);
      unsigned long long idx = sg.in_off + 2 * local;
      a0 = fq_load_stream(T0 + idx); a1 = fq_load_stream(T0 + idx + 1);
      b0 = fq_load_stream(T1 + idx); b1 = fq_load_stream(T1 + idx + 1);
      c0 = fq_load_stream(T2 + idx); c1 = fq_load_stream(T2 + idx + 1);
      w = fq_mul(fq_load(RW + sg.rw_off + row), fq_load(S + i));
    } else {
      unsigned long long idx = sg.in_off + local;
      a0 = fq_load_stream(T0 + idx); b0 = fq_load_stream(T1 + idx); c0 = fq_load_stream(T2 + idx);
      a1 = b1 = c1 = fq_zero();
      w = fq_mul(fq_load(RW + sg.rw_off + local), fq_load(S));
    }
    comb_accumulate<COMB>(acc, w, a0, a1, b0, b1, c0, c1);
  }
  block_sum<3>(acc, sm);
  if (threadIdx.x == 0) {
    partials[blockIdx.x * 3 + 0] = acc[0];
    partials[blockIdx.x * 3 + 1] = acc[1];
    partials[blockIdx.x * 3 + 2] = acc[2];
  }
}

// bind: out = lo + r*(hi - lo); 192 B read + 96 B written per item, 3 modmul
__global__ void __launch_bounds__(RB, SPG_MINB)
k_pair_bind(const fq *__restrict__ T0, const fq *__restrict__ T1, const fq *__restrict__ T2,
            fq *__restrict__ O0, fq *__restrict__ O1, fq *__restrict__ O2,
            const Seg *__restrict__ segs, int nseg, const __grid_constant__ SegPack pk, unsigned long long total_items, fq r) {
  for (unsigned long long item = (unsigned long long)blockIdx.x * RB + threadIdx.x; item < total_items;
       item += (unsigned long long)gridDim.x * RB) {
    Seg sg = pick_seg(pk, segs, nseg, item);
    unsigned long long local = item - sg.item_start;
    unsigned long long o = sg.out_off + local;
    if (sg.log_len >= 1) {
      unsigned long long idx = sg.in_off + 2 * local;
      fq lo = fq_load_stream(T0 + idx), hi = fq_load_stream(T0 + idx + 1);
      fq_store(O0 + o, fq_add(lo, fq_mul(r, fq_sub(hi, lo))));
      lo = fq_load_stream(T1 + idx); hi = fq_load_stream(T1 + idx + 1);
      fq_store(O1 + o, fq_add(lo, fq_mul(r, fq_sub(hi, lo))));
      lo = fq_load_stream(T2 + idx); hi = fq_load_stream(T2 + idx + 1);
      fq_store(O2 + o, fq_add(lo, fq_mul(r, fq_sub(hi, lo))));
    } else {
      unsigned long long idx = sg.in_off + local;
      fq lo = fq_load_stream(T0 + idx);
      fq_store(O0 + o, fq_sub(lo, fq_mul(r, lo)));
      lo = fq_load_stream(T1 + idx);
      fq_store(O1 + o, fq_sub(lo, fq_mul(r, lo)));
      lo = fq_load_stream(T2 + idx);
      fq_store(O2 + o, fq_sub(lo, fq_mul(r, lo)));
    }
  }
}

// fused bind_j + eval_{j+1}: one item = four adjacent input scalars per table ->
// two bound scalars (written) -> one (lo, hi) pair of the next round.
// 384 B read + 192 B written per item, 6 + 7 modmul. Requires log_len >= 2.
template <int COMB>
__global__ void __launch_bounds__(RB, SPG_MINB)
k_quad_bind_eval(const fq *__restrict__ T0, const fq *__restrict__ T1, const fq *__restrict__ T2,
                 fq *__restrict__ O0, fq *__restrict__ O1, fq *__restrict__ O2,
                 const Seg *__restrict__ segs, int nseg, const __grid_constant__ SegPack pk, unsigned long long total_items, fq r,
                 const fq *__restrict__ RW, const fq *__restrict__ Snext, FinishArgs fa) {
  __shared__ fq sm[3 * 32];
  fq acc[3] = {fq_zero(), fq_zero(), fq_zero()};
  for (unsigned long long item = (unsigned long long)blockIdx.x * RB + threadIdx.x; item < total_items;
       item += (unsigned long long)gridDim.x * RB) {
    Seg sg = pick_seg(pk, segs, nseg, item);
    unsigned long long local = item - sg.item_start;
    unsigned int ql = sg.log_len - 2;
    unsigned long long row = local >> ql, i = local & ((1ull << ql) - 1);
    unsigned long long idx = sg.in_off + 4 * local, o = sg.out_off + 2 * local;
    fq lo, hi, a0, a1, b0, b1, c0, c1;
    lo = fq_load_stream(T0 + idx); hi = fq_load_stream(T0 + idx + 1);
    a0 = fq_add(lo, fq_mul(r, fq_sub(hi, lo)));
    lo = fq_load_stream(T0 + idx + 2); hi = fq_load_stream(T0 + idx + 3);
    a1 = fq_add(lo, fq_mul(r, fq_sub(hi, lo)));
    fq_store(O0 + o, a0); fq_store(O0 + o + 1, a1);
    lo = fq_load_stream(T1 + idx); hi = fq_load_stream(T1 + idx + 1);
    b0 = fq_add(lo, fq_mul(r, fq_sub(hi, lo)));
    lo = fq_load_stream(T1 + idx + 2); hi = fq_load_stream(T1 + idx + 3);
    b1 = fq_add(lo, fq_mul(r, fq_sub(hi, lo)));
    fq_store(O1 + o, b0); fq_store(O1 + o + 1, b1);
    lo = fq_load_stream(T2 + idx); hi = fq_load_stream(T2 + idx + 1);
    c0 = fq_add(lo, fq_mul(r, fq_sub(hi, lo)));
    lo = fq_load_stream(T2 + idx + 2); hi = fq_load_stream(T2 + idx + 3);
    c1 = fq_add(lo, fq_mul(r, fq_sub(hi, lo)));
    fq_store(O2 + o, c0); fq_store(O2 + o + 1, c1);
    fq w = fq_mul(fq_load(RW + sg.rw_off + row), fq_load(Snext + i));
    comb_accumulate<COMB>(acc, w, a0, a1, b0, b1, c0, c1);
  }
  block_sum<3>(acc, sm);
  finish_block<3>(fa, acc, sm);
}

// ---------------------------------------------------------------- row-tiled fast path
// One block = one tile of one row, so the row weight RW[row] is applied once per block
// (after the block reduction) instead of once per item, and all arithmetic between the
// loads and the stores runs in the lazy range [0, 2q).
//   k_rows<NE>:    evaluation only; NE = 3 points t = 0, 1, 2 (first round: also yields the true
//                  claim) or NE = 2 points t = 0, 2 when the caller supplied the claim
//                  (spg_sc1_set_claim)
//   k_rows_rolled: bind with r, then evaluate the bound pair at t = 0, 2. The round polynomial is
//                  l_j(t) * G(t) with G quadratic, so G(0), G(2) and the running claim determine
//                  it (the host side solves for G(1), G(3); exact field arithmetic, see
//                  spg_sc1_round_eval).
constexpr int ROWS_LOG_TILE = 10;  // 1024 items per tile = 8 per thread at 128 threads

// Tile -> (row, tile within the row), ROW FASTEST: consecutive blocks work on the same item
// range of different rows (proofs), so what every row re-reads for that range -- the suffix eq
// table S and, in the fused first round, the CSR arrays of the matrices -- is fetched from DRAM
// once and then hits in L2. With the tile index running fastest (the first form) each row
// streamed those arrays again after > L2's worth of table traffic: 2.6 GB of extra DRAM reads
// per first round at 2^20 x 64 (ncu: 13.4 GB against 10.7 GB algorithmic).
// n_rows is a power of two (num_proofs[p] is).
__device__ __forceinline__ void tile_to_row(const Seg &sg, unsigned long long tl, unsigned long long &row,
                                            unsigned long long &tr) {
  unsigned int lr = 31 - __clz(sg.n_rows);
  row = tl & (((unsigned long long)1 << lr) - 1);
  tr = tl >> lr;
}

template <int NE>
__global__ void __launch_bounds__(RB, SPG_MINB)
k_rows(const fq *__restrict__ T0, const fq *__restrict__ T1, const fq *__restrict__ T2, const Seg *__restrict__ segs,
       int nseg, const __grid_constant__ SegPack pk, const fq *__restrict__ RW, const fq *__restrict__ S, FinishArgs fa) {
  static_assert(NE == 2 || NE == 3, "k_rows: 2 or 3 evaluation points");
  __shared__ fq sm[NE * 32];
  unsigned long long tile = blockIdx.x;
  Seg sg = pick_seg(pk, segs, nseg, tile);
  unsigned long long tl = tile - sg.item_start;
  unsigned long long row, tr;
  tile_to_row(sg, tl, row, tr);
  unsigned int li = sg.log_len - 1;
  unsigned long long items_row = 1ull << li, tile_items = items_row >> sg.log_tiles;
  unsigned long long base = tr * tile_items;
  fq acc[NE];
#pragma unroll
  for (int k = 0; k < NE; k++) acc[k] = fq_zero();
  for (unsigned long long it = base + threadIdx.x; it < base + tile_items; it += RB) {
    unsigned long long idx = sg.in_off + 2 * (row * items_row + it);
    fq a0 = fq_load_stream(T0 + idx), a1 = fq_load_stream(T0 + idx + 1);
    fq b0 = fq_load_stream(T1 + idx), b1 = fq_load_stream(T1 + idx + 1);
    fq c0 = fq_load_stream(T2 + idx), c1 = fq_load_stream(T2 + idx + 1);
    fq w = fq_load(S + it);
    // wide-range forms (fq.cuh), bounds as in k_rows_rolled: table entries are canonical here, w is
    // canonical, acc stays in [0, 2q)
    // t = 0
    acc[0] = fq_fold2q(fq_raw_add(acc[0], fq_mul_lazy(w, fq_sub_plus2q(fq_mul_lazy(a0, b0), c0))));
    if (NE == 3)  // t = 1
      acc[1] = fq_fold2q(fq_raw_add(acc[1], fq_mul_lazy(w, fq_sub_plus2q(fq_mul_lazy(a1, b1), c1))));
    // t = 2: 2*hi - lo + 2q in (0, 4q)
    fq a2 = fq_raw_add(a1, fq_sub_plus2q(a1, a0));
    fq b2 = fq_raw_add(b1, fq_sub_plus2q(b1, b0));
    fq c2 = fq_raw_add(c1, fq_sub_plus2q(c1, c0));
    acc[NE - 1] = fq_fold2q(fq_raw_add(acc[NE - 1], fq_mul_lazy(w, fq_sub_plus6q(fq_mul_lazy(a2, b2), c2))));
  }
#pragma unroll
  for (int k = 0; k < NE; k++) acc[k] = fq_canon(acc[k]);
  block_sum<NE>(acc, sm);
  if (threadIdx.x == 0) {
    fq rw = RW[sg.rw_off + row];
#pragma unroll
    for (int k = 0; k < NE; k++) acc[k] = fq_mul(rw, acc[k]);
  }
  finish_block<NE>(fa, acc, sm);
}

// Fused bind_j + eval_{j+1} with ONE copy of the bind code: the six binds of an item (three
// tables x two pairs) run as a rolled loop whose body -- one Montgomery product, the lazy add /
// sub, the canonicalisation and the store -- is about 5 KB of SASS, and the bound scalars are
// handed to the evaluation part through thread-private shared-memory slots. Fully unrolled (the
// first form of this kernel) the item loop is ~45 KB, more than the 32 KB instruction cache (ncu:
// 22 % of its stall samples were no_instructions) and measured 5.99 ms per pass at 2^20 x 64; this
// loop is 21 KB and takes 5.69 ms. The loads of the next (table, pair) are issued
// before the current product, across items too, so the rolled loop does not expose one DRAM
// latency per bind. (Two binds per trip with the operand registers ping-ponging removes the 22
// register copies at the end of a trip, 7 % of the bind's instructions, and measured SLOWER: 5.53 ms
// against 5.41 ms.)
//
// Ranges (fq.cuh, wide-range forms): the tables may come in unreduced, in [0, 2q) -- that is how this
// kernel leaves them when LAZY_OUT, i.e. when the host knows the next bind is this kernel again -- so
// the bind is d = hi - lo + 2q in (0, 4q), r d / R + q < 1.25 q, lo + that < 3.25 q, folded once into
// [0, 2q); the last launch of a run (LAZY_OUT = false) also canonicalises. The evaluation adds 2q or 6q
// instead of correcting conditionally; bounds are stated line by line. Against the form with a
// conditional correction after every add and sub this is ~220 fewer ALU instructions per item.
template <bool LAZY_OUT>
__global__ void __launch_bounds__(RB, SPG_MINB)
k_rows_rolled(const fq *__restrict__ T0, const fq *__restrict__ T1, const fq *__restrict__ T2,
              fq *__restrict__ O0, fq *__restrict__ O1, fq *__restrict__ O2, const Seg *__restrict__ segs,
              int nseg, const __grid_constant__ SegPack pk, fq r, const fq *__restrict__ RW, const fq *__restrict__ S,
              FinishArgs fa) {
  extern __shared__ __align__(32) unsigned char stash_raw[];
  fq *stash = reinterpret_cast<fq *>(stash_raw) + threadIdx.x;  // slot k at stash[k * RB]
  __shared__ fq sm[2 * 32];
  unsigned long long tile = blockIdx.x;
  Seg sg = pick_seg(pk, segs, nseg, tile);
  unsigned long long tl = tile - sg.item_start;
  unsigned long long row, tr;
  tile_to_row(sg, tl, row, tr);
  unsigned int li = sg.log_len - 2;
  unsigned long long items_row = 1ull << li, tile_items = items_row >> sg.log_tiles;
  unsigned long long base = tr * tile_items, end = base + tile_items;
  fq acc[2] = {fq_zero(), fq_zero()};
  unsigned long long it = base + threadIdx.x;
  fq lo, hi;
  if (it < end) {
    unsigned long long idx = sg.in_off + 4 * (row * items_row + it);
    lo = fq_load_stream(T0 + idx);
    hi = fq_load_stream(T0 + idx + 1);
  }
  for (; it < end; it += RB) {
    unsigned long long local = row * items_row + it;
    unsigned long long idx = sg.in_off + 4 * local, o = sg.out_off + 2 * local;
    const bool more = it + RB < end;
#pragma unroll 1
    for (int k = 0; k < 6; k++) {
      // prefetch the operands of bind k + 1 (or of the next item's first bind)
      fq nlo, nhi;
      {
        int kn = k == 5 ? 0 : k + 1;
        unsigned long long nidx = (k == 5 ? idx + 4 * (unsigned long long)RB : idx) + 2 * (kn & 1);
        const fq *Tn = (kn >> 1) == 0 ? T0 : ((kn >> 1) == 1 ? T1 : T2);
        if (k < 5 || more) {
          nlo = fq_load_stream(Tn + nidx);
          nhi = fq_load_stream(Tn + nidx + 1);
        }
      }
      fq v = fq_fold2q(fq_raw_add(lo, fq_mul_lazy(r, fq_sub_plus2q(hi, lo))));
      if (!LAZY_OUT) v = fq_canon(v);
      fq *To = (k >> 1) == 0 ? O0 : ((k >> 1) == 1 ? O1 : O2);
      fq_store(To + o + (k & 1), v);
      stash[k * RB] = v;
      lo = nlo;
      hi = nhi;
    }
    fq w = fq_load(S + it);
    {
      // all six in [0, 2q); w (suffix eq table) is canonical; acc stays in [0, 2q)
      fq a0 = stash[0], b0 = stash[2 * RB], c0 = stash[4 * RB];
      // a0 b0 / R + q < 1.25 q; - c0 + 2q: (0, 3.25 q); w * that / R + q < 1.21 q; acc + that < 3.21 q
      acc[0] = fq_fold2q(fq_raw_add(acc[0], fq_mul_lazy(w, fq_sub_plus2q(fq_mul_lazy(a0, b0), c0))));
      fq a1 = stash[RB], b1 = stash[3 * RB], c1 = stash[5 * RB];
      // 2 x1 - x0 + 2q in (0, 6q)
      fq a2 = fq_raw_add(a1, fq_sub_plus2q(a1, a0));
      fq b2 = fq_raw_add(b1, fq_sub_plus2q(b1, b0));
      fq c2 = fq_raw_add(c1, fq_sub_plus2q(c1, c0));
      // a2 b2 / R + q < 3.26 q; - c2 + 6q: (0, 9.26 q); w * that / R + q < 1.58 q; acc + that < 3.58 q
      acc[1] = fq_fold2q(fq_raw_add(acc[1], fq_mul_lazy(w, fq_sub_plus6q(fq_mul_lazy(a2, b2), c2))));
    }
  }
  acc[0] = fq_canon(acc[0]);
  acc[1] = fq_canon(acc[1]);
  block_sum<2>(acc, sm);
  if (threadIdx.x == 0) {
    fq rw = RW[sg.rw_off + row];
    acc[0] = fq_mul(rw, acc[0]);
    acc[1] = fq_mul(rw, acc[1]);
  }
  finish_block<2>(fa, acc, sm);
}

// tables left in [0, 2q) by k_rows_rolled<true> -> canonical (only when the bind that follows is not the
// one the host predicted, e.g. fusing was switched off in between)
__global__ void k_canon3(fq *__restrict__ T0, fq *__restrict__ T1, fq *__restrict__ T2, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    fq_store(T0 + i, fq_canon(fq_load(T0 + i)));
    fq_store(T1 + i, fq_canon(fq_load(T1 + i)));
    fq_store(T2 + i, fq_canon(fq_load(T2 + i)));
  }
}

// First round with the SpMV fused in (multiply_vec_block + round 0 in one pass): the block
// computes Az, Bz, Cz of its tile straight from the witness sections, writes them for the
// next round and evaluates the round polynomial on the fly, so the three tables are written
// once and not read back (k_spmv3 + k_rows<2> wrote 96 N bytes and read them again).
struct SpmvSegs {
  CsxView3 mats[SEG_INLINE];
  const SecView *secs[SEG_INLINE];
};

template <int NE, bool COLD = false>
__global__ void __launch_bounds__(RB, SPG_MINB)
k_rows_spmv(const __grid_constant__ SpmvSegs SP, unsigned int log_ymax, fq *__restrict__ O0, fq *__restrict__ O1,
            fq *__restrict__ O2, int nseg, const __grid_constant__ SegPack pk, const fq *__restrict__ RW,
            const fq *__restrict__ S, fq *__restrict__ partials, int skip_t0) {
  __shared__ fq sm[NE * 32];
  unsigned long long tile = blockIdx.x;
  int si = 0;
  if (nseg > 1) {
    int lo = 0, hi = nseg - 1;
    while (lo < hi) {
      int mid = (lo + hi + 1) >> 1;
      if (pk.s[mid].item_start <= tile) lo = mid;
      else hi = mid - 1;
    }
    si = lo;
  }
  const Seg sg = pk.s[si];
  const SecView *__restrict__ secs = SP.secs[si];
  unsigned long long tl = tile - sg.item_start;
  unsigned long long row, tr;
  tile_to_row(sg, tl, row, tr);
  unsigned int li = sg.log_len - 1;
  unsigned long long items_row = 1ull << li, tile_items = items_row >> sg.log_tiles;
  unsigned long long base = tr * tile_items;
  fq acc[NE];
#pragma unroll
  for (int k = 0; k < NE; k++) acc[k] = fq_zero();
  // software pipeline: the row heads (pointer + first column index, two dependent L2 hits) of
  // the next item are in flight while the current item's products run, so an iteration waits
  // for one DRAM latency (z) instead of three dependent ones. (Measured: 4.05 ms unpipelined,
  // 3.41 ms with this; a second stage that also L2-prefetches the next z operands spills at
  // 128 registers and is slower, 3.96 ms.)
  unsigned long long it = base + threadIdx.x;
  RowHead3 h_lo, h_hi;
  if (it < base + tile_items) {
    h_lo = spmv_head3(SP.mats[si], (unsigned int)(2 * it));
    h_hi = spmv_head3(SP.mats[si], (unsigned int)(2 * it + 1));
  }
  for (; it < base + tile_items; it += RB) {
    unsigned long long idx = sg.in_off + 2 * (row * items_row + it);
    fq lo3[3], hi3[3];
    spmv_finish3<COLD>(SP.mats[si], h_lo, secs, row, log_ymax, lo3);
    spmv_finish3<COLD>(SP.mats[si], h_hi, secs, row, log_ymax, hi3);
    if (it + RB < base + tile_items) {
      h_lo = spmv_head3(SP.mats[si], (unsigned int)(2 * (it + RB)));
      h_hi = spmv_head3(SP.mats[si], (unsigned int)(2 * (it + RB) + 1));
    }
    fq_store_stream(O0 + idx, lo3[0]); fq_store_stream(O0 + idx + 1, hi3[0]);
    fq_store_stream(O1 + idx, lo3[1]); fq_store_stream(O1 + idx + 1, hi3[1]);
    fq_store_stream(O2 + idx, lo3[2]); fq_store_stream(O2 + idx + 1, hi3[2]);
    fq w = fq_load(S + it);
    // wide-range forms (fq.cuh), bounds as in k_rows_rolled; lo3 / hi3 are canonical
    // skip_t0: the caller vouches that the witness satisfies the instance (spg_sc1_set_satisfied), so
    // Az Bz - Cz vanishes at every row and only the point t = 2 carries information (a run-time flag, uniform
    // over the grid: a separate instantiation without the t = 0 sum crashes ptxas 12.9)
    if (!skip_t0) acc[0] = fq_fold2q(fq_raw_add(acc[0], fq_mul_lazy(w, fq_sub_plus2q(fq_mul_lazy(lo3[0], lo3[1]), lo3[2]))));
    if (NE == 3) acc[1] = fq_fold2q(fq_raw_add(acc[1], fq_mul_lazy(w, fq_sub_plus2q(fq_mul_lazy(hi3[0], hi3[1]), hi3[2]))));
    fq a2 = fq_raw_add(hi3[0], fq_sub_plus2q(hi3[0], lo3[0]));
    fq b2 = fq_raw_add(hi3[1], fq_sub_plus2q(hi3[1], lo3[1]));
    fq c2 = fq_raw_add(hi3[2], fq_sub_plus2q(hi3[2], lo3[2]));
    acc[NE - 1] = fq_fold2q(fq_raw_add(acc[NE - 1], fq_mul_lazy(w, fq_sub_plus6q(fq_mul_lazy(a2, b2), c2))));
  }
#pragma unroll
  for (int k = 0; k < NE; k++) acc[k] = fq_canon(acc[k]);
  block_sum<NE>(acc, sm);
  if (threadIdx.x == 0) {
    fq rw = RW[sg.rw_off + row];
#pragma unroll
    for (int k = 0; k < NE; k++) partials[(unsigned long long)blockIdx.x * NE + k] = fq_mul(rw, acc[k]);
  }
}

// ---------------------------------------------------------------- p rounds (tiny)
// tables hold one scalar per instance, zero padded to P'. MODE_P binds the TOP bit
// of p (no reversal): pairs (p, p + half). sumcheck.rs:1186-1245 with mode P.
__global__ void k_p_eval(const fq *__restrict__ Ap, const fq *__restrict__ T0,
                         const fq *__restrict__ T1, const fq *__restrict__ T2, size_t half,
                         size_t limit, fq *__restrict__ out) {
  __shared__ fq sm[3 * 32];
  fq acc[3] = {fq_zero(), fq_zero(), fq_zero()};
  for (size_t p = threadIdx.x; p < limit; p += blockDim.x) {
    fq A0 = Ap[p], A1 = Ap[p + half], A2, A3;
    line23(A0, A1, A2, A3);
    fq a0 = T0[p], a1 = T0[p + half], b0 = T1[p], b1 = T1[p + half], c0 = T2[p], c1 = T2[p + half];
    fq a2, a3, b2, b3, c2, c3;
    line23(a0, a1, a2, a3);
    line23(b0, b1, b2, b3);
    line23(c0, c1, c2, c3);
    acc[0] = fq_add(acc[0], fq_mul(A0, fq_sub(fq_mul(a0, b0), c0)));
    acc[1] = fq_add(acc[1], fq_mul(A2, fq_sub(fq_mul(a2, b2), c2)));
    acc[2] = fq_add(acc[2], fq_mul(A3, fq_sub(fq_mul(a3, b3), c3)));
  }
  block_sum<3>(acc, sm);
  if (threadIdx.x == 0) {
    out[0] = acc[0];
    out[1] = acc[1];
    out[2] = acc[2];
  }
}

__global__ void k_p_bind(fq *__restrict__ Ap, fq *__restrict__ T0, fq *__restrict__ T1,
                         fq *__restrict__ T2, size_t half, fq r) {
  for (size_t p = threadIdx.x; p < half; p += blockDim.x) {
    fq lo = Ap[p], hi = Ap[p + half];
    Ap[p] = fq_add(lo, fq_mul(r, fq_sub(hi, lo)));
    lo = T0[p]; hi = T0[p + half];
    T0[p] = fq_add(lo, fq_mul(r, fq_sub(hi, lo)));
    lo = T1[p]; hi = T1[p + half];
    T1[p] = fq_add(lo, fq_mul(r, fq_sub(hi, lo)));
    lo = T2[p]; hi = T2[p + half];
    T2[p] = fq_add(lo, fq_mul(r, fq_sub(hi, lo)));
  }
}

// RW[off_p + q] = Ap[p] * Eq_nat[q]
__global__ void k_row_weights(const fq *__restrict__ Ap, const fq *__restrict__ Eq,
                              const unsigned long long *__restrict__ rw_off,
                              const unsigned int *__restrict__ Qp, int P, fq *__restrict__ RW) {
  int p = blockIdx.y;
  if (p >= P) return;
  fq ap = Ap[p];
  for (unsigned int q = blockIdx.x * blockDim.x + threadIdx.x; q < Qp[p]; q += gridDim.x * blockDim.x)
    RW[rw_off[p] + q] = fq_mul(ap, Eq[q]);
}

// suffix eq tables for LSB-first binding: level m (2^m entries) lives at buf + 2^m,
//   level_m[2i + b] = level_{m-1}[i] * eq(tau[n - m], b),  level_0 = [1]
// so level m is the eq table of tau[n-m .. n-1] with index bit k <-> tau[n-m+k].
// The first SUFFIX_SMALL levels are built by one block in one launch (no host round trip for
// the leading one either); the larger levels take one streaming launch each.
constexpr int SUFFIX_SMALL = 9;
struct SmallTaus {
  fq t[SUFFIX_SMALL];
};
__global__ void k_suffix_small(fq *__restrict__ buf, const __grid_constant__ SmallTaus taus, int levels) {
  if (threadIdx.x == 0) buf[1] = fq_one();
  __syncthreads();
  for (int m = 1; m <= levels; m++) {
    size_t cnt = (size_t)1 << (m - 1);
    fq r = taus.t[m - 1];
    for (size_t i = threadIdx.x; i < cnt; i += blockDim.x) {
      fq s = buf[cnt + i];
      fq hi = fq_mul(s, r);
      buf[2 * cnt + 2 * i + 1] = hi;
      buf[2 * cnt + 2 * i] = fq_sub(s, hi);
    }
    __syncthreads();
  }
}

int build_suffix_tables(spg_ctx *ctx, const std::vector<hfq> &tau, size_t max_level, fq *buf) {
  size_t n = tau.size();
  SmallTaus st;
  memset(&st, 0, sizeof st);
  int small = (int)(max_level < (size_t)SUFFIX_SMALL ? max_level : SUFFIX_SMALL);
  for (int m = 1; m <= small; m++) memcpy(&st.t[m - 1], &tau[n - m], sizeof(fq));
  SPG_LAUNCH(ctx, k_suffix_small, 1, 256, 0, buf, st, small);
  // the larger levels three per launch (k_eq_expand_multi, field_ops.cu): level m0 -> m0 + 1 .. m0 + steps
  for (size_t m0 = small; m0 < max_level;) {
    int steps = (int)(max_level - m0 < 3 ? max_level - m0 : 3);
    spg_fq r[3];
    for (int k = 0; k < steps; k++) memcpy(&r[k], &tau[n - (m0 + 1 + k)], sizeof(spg_fq));
    SPG_TRY(eq_expand_steps(ctx, buf + ((size_t)1 << m0), (size_t)1 << m0, buf, (unsigned int)m0, r, steps, true));
    m0 += steps;
  }
  return SPG_OK;
}

}  // namespace spg

using namespace spg;

struct spg_sc1 {
  spg_ctx *ctx = nullptr;
  size_t P = 0, Pp = 1;  // instances, padded to a power of two
  size_t nx = 0, nq = 0, np = 0;
  std::vector<size_t> Q, X;
  std::vector<hfq> tau_p, tau_q, tau_x;
  fq *tab[2][3] = {{nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr}};
  size_t cap[2] = {0, 0};
  int cur = 0;
  fq *Sx = nullptr, *Sq = nullptr, *Ap = nullptr, *RWx = nullptr;
  Seg *d_segs = nullptr;
  unsigned long long *d_rw_off = nullptr;
  unsigned int *d_Qp = nullptr;
  std::vector<Seg> segs;
  std::vector<unsigned> loglen;  // current log row length per instance
  size_t round = 0;
  bool evaluated = false;
  // raw device sums for the current round, produced by the fused bind+eval of the previous
  // bind: 0 = none, 3 = sums at t = 0, 2, 3, 2 = sums at t = 0, 2 (needs the running claim)
  int cached_kind = 0;
  hfq cached[3];
  // the true running claim s_{j-1}(r_{j-1}) (= e(0) + e(1) of the current round), maintained
  // from the evaluations this object itself produced: exact for any input tables
  bool tab_lazy = false;   // tab[cur] holds values in [0, 2q) (k_rows_rolled<true>); only k_rows_rolled reads those
  size_t tab_lazy_n = 0;
  bool claim_known = false;
  bool check_claim = false;  // verify a supplied claim against the tables in the first round (spg_sc1_set_claim_checked)
  bool satisfied = false;    // spg_sc1_set_satisfied: every row of Az Bz - Cz is zero, the first round needs only t = 2
  hfq supplied_claim;
  hfq claim;
  hfq last_e[3];
  // the same claim divided by the scalar prefix c of the current round: with
  // s_j(t) = c_j l_j(t) G_j(t) and c_{j+1} = c_j l_j(r_j) it is simply G_j(r_j), so the
  // two-point rounds need no field inversion (tau^-1 is precomputed)
  bool g_known = false, lastG_valid = false;
  hfq gclaim;
  hfq lastG[3];  // G(0), G(1), G(2) of the round just evaluated
  std::vector<hfq> tau_x_inv, tau_q_inv;  // zero where tau is zero
  hfq cx, cq;    // prod eq(tau_k, r_k) over the bound x / q variables
  hfq scale;     // external factor on every evaluation (spg_sc1_set_scale); one by default
  // multiply_vec_block deferred into the first round (see k_rows_spmv); the instance and
  // the z_mat must stay alive until the first spg_sc1_round_eval has returned
  const spg_r1cs *pend_inst = nullptr;
  const spg_zmat *pend_z = nullptr;
  size_t pend_max_num_inputs = 0;
  size_t p_len = 1;  // current instance_len during the p rounds
  bool fuse = true;
  bool p_ready = false;
};

namespace {

int sc1_materialize(spg_sc1 *s);
bool sc1_rows_eligible0(const spg_sc1 *s);

int phase_of(const spg_sc1 *s, size_t round) {
  if (round < s->nx) return 0;
  if (round < s->nx + s->nq) return 1;
  return 2;
}

// segments for the current phase; item counts for pair kernels (div = 1) or quad kernels (div = 2)
void build_segs(spg_sc1 *s, int phase, int quad, unsigned long long *total_items,
                unsigned long long *out_total) {
  unsigned long long in_off = 0, out_off = 0, items = 0, rw = 0;
  s->segs.resize(s->P);
  for (size_t p = 0; p < s->P; p++) {
    Seg &g = s->segs[p];
    unsigned ll = s->loglen[p];
    unsigned long long rows = phase == 0 ? s->Q[p] : 1;
    g.in_off = in_off;
    g.out_off = out_off;
    g.item_start = items;
    g.log_len = ll;
    g.n_rows = (unsigned)rows;
    g.rw_off = (unsigned)rw;
    g.log_tiles = 0;
    unsigned long long in_sz = rows << ll;
    unsigned long long out_sz = ll >= 1 ? in_sz >> 1 : in_sz;
    unsigned long long it = quad ? (in_sz >> 2) : out_sz;
    in_off += in_sz;
    out_off += out_sz;
    items += it;
    rw += rows;
  }
  *total_items = items;
  *out_total = out_off;
}

int upload_segs(spg_sc1 *s) {
  if (s->P <= (size_t)SEG_INLINE) return SPG_OK;  // the segments ride in the kernel arguments
  SPG_CUDA(cudaMemcpyAsync(s->d_segs, s->segs.data(), s->P * sizeof(Seg), cudaMemcpyHostToDevice,
                           s->ctx->stream));
  return SPG_OK;
}

const fq *s_table(const spg_sc1 *s, int phase, size_t level) {
  const fq *base = phase == 0 ? s->Sx : s->Sq;
  return base + ((size_t)1 << level);
}

// rounds already bound inside the current phase
size_t phase_round(const spg_sc1 *s, size_t round) {
  int ph = phase_of(s, round);
  return ph == 0 ? round : (ph == 1 ? round - s->nx : round - s->nx - s->nq);
}

constexpr size_t ROWS_STASH_BYTES = 6 * RB * sizeof(fq);  // k_rows_rolled: six bound scalars per thread
// The fused first round keeps the multi-entry / non-unit row handling out of line
// (spmv_row_rest, r1cs.cuh): 3.42 -> 3.14 ms per pass at 2^20 x 64. SPG_SPMV_COLD=0 selects the
// fully inlined form again (A/B switch).
bool spmv_cold_enabled() {
  const char *e = getenv("SPG_SPMV_COLD");
  return !(e && *e == '0');
}

int sc1_alloc_common(spg_ctx *ctx, size_t P, const size_t *num_proofs, size_t max_num_proofs,
                     const size_t *num_cons, size_t max_num_cons, const spg_fq *tau_p,
                     const spg_fq *tau_q, const spg_fq *tau_x, spg_sc1 **out) {
  SPG_CHECK(ctx && out && num_proofs && num_cons, "spg_sc1_create: null argument");
  SPG_CHECK(P >= 1, "spg_sc1_create: num_instances must be >= 1");
  SPG_CHECK(is_pow2(max_num_proofs) && is_pow2(max_num_cons),
            "spg_sc1_create: max_num_proofs / max_num_cons must be powers of two");
  for (size_t p = 0; p < P; p++) {
    SPG_CHECK(is_pow2(num_proofs[p]) && num_proofs[p] <= max_num_proofs,
              "spg_sc1_create: num_proofs[%zu] = %zu is not a power of two <= %zu", p, num_proofs[p],
              max_num_proofs);
    SPG_CHECK(is_pow2(num_cons[p]) && num_cons[p] <= max_num_cons,
              "spg_sc1_create: num_cons[%zu] = %zu is not a power of two <= %zu", p, num_cons[p],
              max_num_cons);
  }
  spg_sc1 *s = new (std::nothrow) spg_sc1();
  if (!s) return SPG_ENOMEM;
  s->ctx = ctx;
  s->P = P;
  s->Pp = next_pow2(P);
  s->nx = log2u(max_num_cons);
  s->nq = log2u(max_num_proofs);
  s->np = log2u(s->Pp);
  SPG_CHECK((s->np == 0 || tau_p) && (s->nq == 0 || tau_q) && (s->nx == 0 || tau_x),
            "spg_sc1_create: null tau");
  s->Q.assign(num_proofs, num_proofs + P);
  s->X.assign(num_cons, num_cons + P);
  for (size_t i = 0; i < s->np; i++) s->tau_p.push_back(hfq_from(tau_p[i]));
  for (size_t i = 0; i < s->nq; i++) s->tau_q.push_back(hfq_from(tau_q[i]));
  for (size_t i = 0; i < s->nx; i++) s->tau_x.push_back(hfq_from(tau_x[i]));
  {
    // Montgomery batch inversion of all taus (one field inversion)
    std::vector<hfq> all(s->tau_x);
    all.insert(all.end(), s->tau_q.begin(), s->tau_q.end());
    std::vector<hfq> pre(all.size() + 1, hfq_one());
    for (size_t i = 0; i < all.size(); i++) pre[i + 1] = hfq_is_zero(all[i]) ? pre[i] : hfq_mul(pre[i], all[i]);
    hfq inv = hfq_invert(pre[all.size()]);
    std::vector<hfq> out(all.size(), hfq_zero());
    for (size_t i = all.size(); i-- > 0;) {
      if (hfq_is_zero(all[i])) continue;
      out[i] = hfq_mul(inv, pre[i]);
      inv = hfq_mul(inv, all[i]);
    }
    s->tau_x_inv.assign(out.begin(), out.begin() + s->nx);
    s->tau_q_inv.assign(out.begin() + s->nx, out.end());
  }
  size_t N = 0, rows = 0;
  for (size_t p = 0; p < P; p++) {
    N += s->Q[p] * s->X[p];
    rows += s->Q[p];
  }
  size_t N1 = 0;  // size after the first bind: exhausted rows (length 1) do not shrink
  for (size_t p = 0; p < P; p++) {
    size_t len = s->nx ? s->X[p] : s->Q[p], rws = s->nx ? s->Q[p] : 1;
    N1 += rws * (len > 1 ? len / 2 : 1);
  }
  s->cap[0] = N > s->Pp ? N : s->Pp;
  s->cap[1] = N1 > s->Pp ? N1 : s->Pp;
  for (int b = 0; b < 2; b++)
    for (int k = 0; k < 3; k++) SPG_CUDA(dev_alloc(ctx, &s->tab[b][k], s->cap[b] * sizeof(fq)));
  SPG_CUDA(dev_alloc(ctx, &s->Sx, ((size_t)2 << s->nx) * sizeof(fq)));
  SPG_CUDA(dev_alloc(ctx, &s->Sq, ((size_t)2 << s->nq) * sizeof(fq)));
  SPG_CUDA(dev_alloc(ctx, &s->Ap, s->Pp * sizeof(fq)));
  SPG_CUDA(dev_alloc(ctx, &s->RWx, rows * sizeof(fq)));
  SPG_CUDA(dev_alloc(ctx, &s->d_segs, P * sizeof(Seg)));
  SPG_CUDA(dev_alloc(ctx, &s->d_rw_off, P * sizeof(unsigned long long)));
  SPG_CUDA(dev_alloc(ctx, &s->d_Qp, P * sizeof(unsigned int)));
  s->cx = hfq_one();
  s->cq = hfq_one();
  s->scale = hfq_one();
  s->p_len = s->Pp;
  s->loglen.resize(P);
  for (size_t p = 0; p < P; p++) s->loglen[p] = s->nx ? log2u(s->X[p]) : log2u(s->Q[p]);
  *out = s;
  return SPG_OK;
}

// eq tables + row weights (src/r1csproof.rs:305-312 and the Ap*Aq products of sumcheck.rs:1186)
int sc1_build_weights(spg_sc1 *s) {
  spg_ctx *ctx = s->ctx;
  // Ap = eq(tau_p).evals() in the reference's order (p is never bit-reversed)
  {
    fq *d_r = nullptr, *scratch = nullptr;
    std::vector<spg_fq> hr(s->np ? s->np : 1);
    for (size_t i = 0; i < s->np; i++) hr[i] = hfq_to(s->tau_p[i]);
    SPG_CUDA(dev_alloc(ctx, &d_r, hr.size() * sizeof(fq)));
    SPG_CUDA(dev_alloc(ctx, &scratch, s->Pp * sizeof(fq)));
    SPG_CUDA(cudaMemcpyAsync(d_r, hr.data(), s->np * sizeof(fq), cudaMemcpyHostToDevice, ctx->stream));
    // (the pageable upload above is staged before cudaMemcpyAsync returns, so hr may go out of scope)
    int rc = eq_evals_device(ctx, d_r, hr.data(), s->np, s->Ap, scratch);
    dev_free(ctx, d_r);
    dev_free(ctx, scratch);
    SPG_TRY(rc);
  }
  SPG_TRY(build_suffix_tables(ctx, s->tau_x, s->nx, s->Sx));
  SPG_TRY(build_suffix_tables(ctx, s->tau_q, s->nq, s->Sq));
  std::vector<unsigned long long> rw_off(s->P);
  std::vector<unsigned int> Qp(s->P);
  unsigned long long rw = 0;
  unsigned int maxQ = 1;
  for (size_t p = 0; p < s->P; p++) {
    rw_off[p] = rw;
    Qp[p] = (unsigned)s->Q[p];
    rw += s->Q[p];
    if (Qp[p] > maxQ) maxQ = Qp[p];
  }
  SPG_CUDA(cudaMemcpyAsync(s->d_rw_off, rw_off.data(), s->P * sizeof(unsigned long long),
                           cudaMemcpyHostToDevice, ctx->stream));
  SPG_CUDA(cudaMemcpyAsync(s->d_Qp, Qp.data(), s->P * sizeof(unsigned int), cudaMemcpyHostToDevice,
                           ctx->stream));
  dim3 grid((maxQ + 127) / 128, (unsigned)s->P);
  SPG_LAUNCH(ctx, k_row_weights, grid, 128, 0, s->Ap, s->Sq + ((size_t)1 << s->nq), s->d_rw_off,
             s->d_Qp, (int)s->P, s->RWx);
  return SPG_OK;
}

// tables left in [0, 2q) by the fused bind: canonicalise in place (see k_canon3)
int sc1_canon_tables(spg_sc1 *s) {
  if (!s->tab_lazy) return SPG_OK;
  spg_ctx *ctx = s->ctx;
  SPG_LAUNCH(ctx, k_canon3, grid_for(ctx, s->tab_lazy_n, 256, 4), 256, 0, s->tab[s->cur][0], s->tab[s->cur][1], s->tab[s->cur][2],
             s->tab_lazy_n);
  s->tab_lazy = false;
  return SPG_OK;
}

// runs the deferred multiply_vec_block (anything but the fused first round needs the tables)
int sc1_materialize(spg_sc1 *s) {
  if (s->tab_lazy) return sc1_canon_tables(s);  // (never together with a pending SpMV: that is round 0)
  if (!s->pend_inst) return SPG_OK;
  const spg_r1cs *inst = s->pend_inst;
  const spg_zmat *z = s->pend_z;
  s->pend_inst = nullptr;
  s->pend_z = nullptr;
  return r1cs_multiply_vec_block(s->ctx, inst, z, s->P, s->Q.data(), s->X.data(), s->pend_max_num_inputs, s->tab[0][0],
                                 s->tab[0][1], s->tab[0][2]);
}

bool sc1_rows_eligible0(const spg_sc1 *s) {
  for (size_t p = 0; p < s->P; p++)
    if (s->loglen[p] < 8) return false;
  return true;
}

// after the last q round every table holds one scalar per instance: pad to P' with zeros
int sc1_enter_p_phase(spg_sc1 *s) {
  if (s->p_ready) return SPG_OK;
  if (s->Pp > s->P)
    for (int k = 0; k < 3; k++)
      SPG_CUDA(cudaMemsetAsync(s->tab[s->cur][k] + s->P, 0, (s->Pp - s->P) * sizeof(fq), s->ctx->stream));
  s->p_ready = true;
  return SPG_OK;
}

}  // namespace

extern "C" {

int spg_sc1_create_from_tables(spg_ctx *ctx, size_t num_instances, const size_t *num_proofs,
                               size_t max_num_proofs, const size_t *num_cons, size_t max_num_cons,
                               const spg_fq *Az, const spg_fq *Bz, const spg_fq *Cz,
                               const spg_fq *tau_p, const spg_fq *tau_q, const spg_fq *tau_x,
                               spg_sc1 **out) {
  spg::DeviceGuard _dev(spg::ctx_of(ctx));
  SPG_CHECK(Az && Bz && Cz, "spg_sc1_create_from_tables: null table");
  spg_sc1 *s = nullptr;
  SPG_TRY(sc1_alloc_common(ctx, num_instances, num_proofs, max_num_proofs, num_cons, max_num_cons,
                           tau_p, tau_q, tau_x, &s));
  size_t N = 0;
  for (size_t p = 0; p < s->P; p++) N += s->Q[p] * s->X[p];
  const spg_fq *src[3] = {Az, Bz, Cz};
  for (int k = 0; k < 3; k++) {
    cudaError_t e = cudaMemcpyAsync(s->tab[0][k], src[k], N * sizeof(fq), cudaMemcpyHostToDevice, ctx->stream);
    if (e != cudaSuccess) {
      spg_sc1_destroy(s);
      return cuda_fail(e, "table upload", __FILE__, __LINE__);
    }
  }
  int rc = sc1_build_weights(s);
  if (rc != SPG_OK) {
    spg_sc1_destroy(s);
    return rc;
  }
  *out = s;
  return SPG_OK;
}

int spg_sc1_create(spg_ctx *ctx, const spg_r1cs *inst, const spg_zmat *z, size_t num_instances,
                   const size_t *num_proofs, size_t max_num_proofs, const size_t *num_cons,
                   size_t max_num_cons, size_t max_num_inputs, const spg_fq *tau_p,
                   const spg_fq *tau_q, const spg_fq *tau_x, spg_sc1 **out) {
  spg::DeviceGuard _dev(spg::ctx_of(ctx));
  SPG_CHECK(inst && z, "spg_sc1_create: null instance / z_mat");
  SPG_CHECK(max_num_cons == inst->max_num_cons, "spg_sc1_create: max_num_cons %zu != instance's %zu",
            max_num_cons, inst->max_num_cons);
  spg_sc1 *s = nullptr;
  SPG_TRY(sc1_alloc_common(ctx, num_instances, num_proofs, max_num_proofs, num_cons, max_num_cons,
                           tau_p, tau_q, tau_x, &s));
  int rc = r1cs_spmv_validate(inst, z, num_instances, num_proofs, num_cons, max_num_inputs);
  if (rc == SPG_OK) {
    s->pend_inst = inst;
    s->pend_z = z;
    s->pend_max_num_inputs = max_num_inputs;
    // the fused first round needs tile-able rows and an inline segment table; otherwise multiply now
    if (!(s->nx >= 1 && s->P <= (size_t)SEG_INLINE && sc1_rows_eligible0(s))) rc = sc1_materialize(s);
  }
  if (rc == SPG_OK) rc = sc1_build_weights(s);
  if (rc != SPG_OK) {
    spg_sc1_destroy(s);
    return rc;
  }
  *out = s;
  return SPG_OK;
}

int spg_sc1_set_scale(spg_sc1 *s, const spg_fq *c) {
  SPG_CHECK(s && c, "spg_sc1_set_scale: null argument");
  s->scale = hfq_from(*c);
  return SPG_OK;
}

int spg_sc1_set_claim(spg_sc1 *s, const spg_fq *claim) {
  SPG_CHECK(s && claim, "spg_sc1_set_claim: null argument");
  if (s->round != 0 || s->evaluated) {
    set_error("spg_sc1_set_claim: must be called before the first round");
    return SPG_ESTATE;
  }
  s->claim = hfq_from(*claim);
  s->claim_known = true;
  return SPG_OK;
}

// The caller vouches that the witness satisfies the instance: (Az Bz - Cz)[p][q][x] = 0 for EVERY row, not only
// in the weighted sum (for random tau the two are the same statement up to negligible probability, and the
// prover of a correct execution knows it outright). Implies spg_sc1_set_claim(0); in addition the first
// round -- the one fused with the SpMV -- then evaluates ONE point per pair, t = 2: e(0) and e(1) are sums of
// zeros. Half the products of that kernel. Exact (bit-identical to the reference, which computes the zeros)
// iff the statement holds; spg_sc1_set_claim_checked is the variant that verifies instead of trusting.
int spg_sc1_set_satisfied(spg_sc1 *s) {
  SPG_CHECK(s, "spg_sc1_set_satisfied: null argument");
  spg_fq zero;
  memset(&zero, 0, sizeof zero);
  SPG_TRY(spg_sc1_set_claim(s, &zero));
  s->satisfied = true;
  return SPG_OK;
}

// Replaces the row weights RW[p][q] = eq_p[p] * eq_q[q] of the x rounds by caller-supplied ones (one
// per (instance, proof) row, in table order). A rank of a sharded proof that owns an arbitrary subset of
// the batch's rows passes the GLOBAL eq weights of its rows here; the q and p rounds of such a
// proof run on the gathered per-row scalars (parallel.ShardedRows), not on this handle.
int spg_sc1_set_row_weights(spg_sc1 *s, const spg_fq *weights, size_t n_rows) {
  spg::DeviceGuard _dev(spg::ctx_of(s));
  SPG_CHECK(s && weights, "spg_sc1_set_row_weights: null argument");
  if (s->round != 0 || s->evaluated) {
    set_error("spg_sc1_set_row_weights: must be called before the first round");
    return SPG_ESTATE;
  }
  size_t rows = 0;
  for (size_t p = 0; p < s->P; p++) rows += s->Q[p];
  SPG_CHECK(n_rows == rows, "spg_sc1_set_row_weights: %zu weights for %zu rows", n_rows, rows);
  SPG_CUDA(cudaMemcpyAsync(s->RWx, weights, rows * sizeof(fq), cudaMemcpyHostToDevice, s->ctx->stream));
  SPG_CUDA(cudaStreamSynchronize(s->ctx->stream));  // the caller's buffer may be pageable and short-lived
  return SPG_OK;
}

// spg_sc1_set_claim plus a check: the first round evaluates three points (like a prover without a
// claim), which yields the true sum over the tables; if it differs from `claim` the round returns
// SPG_EINVAL instead of emitting a proof that cannot verify. Costs one evaluation point of the first
// round. With a plain spg_sc1_set_claim a wrong claim (an unsatisfied witness) goes unnoticed here --
// as it does in the reference, whose verifier rejects the proof later.
int spg_sc1_set_claim_checked(spg_sc1 *s, const spg_fq *claim) {
  SPG_TRY(spg_sc1_set_claim(s, claim));
  s->check_claim = true;
  s->supplied_claim = hfq_from(*claim);
  return SPG_OK;
}

size_t spg_sc1_num_rounds(const spg_sc1 *s) { return s ? s->nx + s->nq + s->np : 0; }

namespace {

// segments for the row-tiled kernels: item_start counts tiles
void build_tile_segs(spg_sc1 *s, int phase, int fused, unsigned long long *tiles_out, unsigned long long *out_total) {
  unsigned long long in_off = 0, out_off = 0, tiles = 0, rw = 0;
  s->segs.resize(s->P);
  // tile size: 1024 items while that still gives every SM several blocks; smaller tiles (down to
  // one item per thread) for the later rounds and for shards with few rows, which otherwise run
  // on a handful of blocks (8 rows of 2^12 items: 32 blocks of 1024 instead of 256 of 128)
  unsigned long long total_items = 0;
  for (size_t p = 0; p < s->P; p++) total_items += (unsigned long long)(phase == 0 ? s->Q[p] : 1) << (s->loglen[p] - (fused ? 2 : 1));
  unsigned log_tile = ROWS_LOG_TILE;
  const unsigned long long want_blocks = (unsigned long long)s->ctx->sm_count * 8;
  while (log_tile > 7 && (total_items >> log_tile) < want_blocks) log_tile--;
  for (size_t p = 0; p < s->P; p++) {
    Seg &g = s->segs[p];
    unsigned ll = s->loglen[p];
    unsigned long long rows = phase == 0 ? s->Q[p] : 1;
    unsigned li = ll - (fused ? 2 : 1);
    unsigned lt = li > log_tile ? li - log_tile : 0;
    g.in_off = in_off;
    g.out_off = out_off;
    g.item_start = tiles;
    g.log_len = ll;
    g.n_rows = (unsigned)rows;
    g.rw_off = (unsigned)rw;
    g.log_tiles = lt;
    in_off += rows << ll;
    out_off += (rows << ll) >> 1;
    tiles += rows << lt;
    rw += rows;
  }
  *tiles_out = tiles;
  *out_total = out_off;
}

// the tiled kernels need every row to hold at least one item per thread of a block
bool rows_eligible(const spg_sc1 *s, int fused, unsigned rounds_ahead = 0) {
  unsigned need = (fused ? 2 : 1) + 7 + rounds_ahead;
  for (size_t p = 0; p < s->P; p++)
    if (s->loglen[p] < need) return false;
  return true;
}


// per-segment matrices and witness views for the fused first round
SpmvSegs make_spmv_segs(const spg_sc1 *s) {
  SpmvSegs sp;
  memset(&sp, 0, sizeof sp);
  const spg_r1cs *inst = s->pend_inst;
  for (size_t p = 0; p < s->P && p < (size_t)SEG_INLINE; p++) {
    size_t pi = inst->num_instances == 1 ? 0 : p;
    for (int m = 0; m < 3; m++) sp.mats[p].M[m] = csx_view(inst->by_row[3 * pi + m]);
    sp.secs[p] = s->pend_z->views + p * s->pend_z->W;
  }
  return sp;
}

// value at r of the cubic through (0, e0), (1, e1), (2, e2), (3, e3)  (UniPoly::from_evals + evaluate)
hfq cubic_at(const hfq &e0, const hfq &e1, const hfq &e2, const hfq &e3, const hfq &r) {
  static const hfq two_inv = hfq_invert(hfq_from_u64(2)), six_inv = hfq_invert(hfq_from_u64(6));
  hfq three_e1 = hfq_add(hfq_add(e1, e1), e1), three_e2 = hfq_add(hfq_add(e2, e2), e2);
  hfq a = hfq_mul(six_inv, hfq_sub(hfq_add(hfq_sub(e3, three_e2), three_e1), e0));
  hfq five_e1 = hfq_add(hfq_add(three_e1, e1), e1), four_e2 = hfq_add(three_e2, e2);
  hfq b = hfq_mul(two_inv, hfq_sub(hfq_add(hfq_sub(hfq_add(e0, e0), five_e1), four_e2), e3));
  hfq c = hfq_sub(hfq_sub(hfq_sub(e1, e0), a), b);
  return hfq_add(e0, hfq_mul(r, hfq_add(c, hfq_mul(r, hfq_add(b, hfq_mul(r, a))))));
}

}  // namespace

int spg_sc1_round_eval(spg_sc1 *s, spg_fq e[3]) {
  spg::DeviceGuard _dev(spg::ctx_of(s));
  SPG_CHECK(s && e, "spg_sc1_round_eval: null argument");
  if (s->round >= spg_sc1_num_rounds(s)) {
    set_error("spg_sc1_round_eval: all %zu rounds are done", spg_sc1_num_rounds(s));
    return SPG_ESTATE;
  }
  if (s->evaluated) {
    set_error("spg_sc1_round_eval: round %zu already evaluated; call round_bind", s->round);
    return SPG_ESTATE;
  }
  spg_ctx *ctx = s->ctx;
  int phase = phase_of(s, s->round);
  size_t j = phase_round(s, s->round);
  hfq ev[3];  // e(0), e(2), e(3)
  if (phase == 2) {
    SPG_TRY(sc1_materialize(s));
    SPG_TRY(sc1_enter_p_phase(s));
    size_t half = s->p_len / 2;
    size_t limit = half < s->P ? half : s->P;
    SPG_LAUNCH(ctx, k_p_eval, 1, 128, 0, s->Ap, s->tab[s->cur][0], s->tab[s->cur][1],
               s->tab[s->cur][2], half, limit, ctx->d_result);
    spg_fq tmp[3];
    SPG_TRY(fetch_result(ctx, 3, tmp));
    hfq c = hfq_mul(hfq_mul(s->cx, s->cq), s->scale);
    for (int t = 0; t < 3; t++) ev[t] = hfq_mul(c, hfq_from(tmp[t]));
  } else {
    // scalar prefix and the active variable's eq line l(t), left to the host by the kernels
    const hfq &tau = phase == 0 ? s->tau_x[j] : s->tau_q[j];
    hfq line[3];
    hfq_eq_line_023(tau, line);  // l(0), l(2), l(3); l(1) = tau
    hfq c = hfq_mul(phase == 0 ? s->cx : hfq_mul(s->cx, s->cq), s->scale);
    size_t n_phase = phase == 0 ? s->nx : s->nq;
    const fq *S = s_table(s, phase, n_phase - j - 1);
    const fq *RW = phase == 0 ? s->RWx : s->Ap;
    bool done = false;
    const bool verify_now = s->check_claim && s->round == 0;  // then the three-point path below yields the true sum
    if (s->cached_kind == 0 && s->claim_known && !verify_now && rows_eligible(s, 0) && !hfq_is_zero(hfq_mul(c, tau))) {
      // the claim is known (spg_sc1_set_claim, or tracked from earlier rounds): two points suffice
      unsigned long long tiles = 0, out_total = 0;
      build_tile_segs(s, phase, 0, &tiles, &out_total);
      SPG_TRY(upload_segs(s));
      SPG_TRY(ensure_partials(ctx, (size_t)tiles * 2));
      double pairs = 0;
      for (size_t p = 0; p < s->P; p++) pairs += (double)((phase == 0 ? s->Q[p] : 1) << (s->loglen[p] - 1));
      ctx->next_units = 192.0 * pairs;
      FinishArgs fa = {};  // seq == 0: separate reduce
      if (s->pend_inst) {
        // tables do not exist yet: compute them in the same pass (read z once, write 96 N bytes once)
        ctx->next_units = 288.0 * pairs;
        static const bool cold = spmv_cold_enabled();
        const int skip_t0 = s->satisfied && s->round == 0;
        if (cold)
          SPG_LAUNCH(ctx, (k_rows_spmv<2, true>), (unsigned)tiles, RB, 0, make_spmv_segs(s), log2u(s->pend_max_num_inputs),
                     s->tab[0][0], s->tab[0][1], s->tab[0][2], (int)s->P, make_pack(s->segs), RW, S, ctx->d_partials, skip_t0);
        else
          SPG_LAUNCH(ctx, (k_rows_spmv<2>), (unsigned)tiles, RB, 0, make_spmv_segs(s), log2u(s->pend_max_num_inputs),
                     s->tab[0][0], s->tab[0][1], s->tab[0][2], (int)s->P, make_pack(s->segs), RW, S, ctx->d_partials, skip_t0);
        s->pend_inst = nullptr;
        s->pend_z = nullptr;
      } else {
        fa = finish_args(ctx, tiles);
        SPG_LAUNCH(ctx, k_rows<2>, (unsigned)tiles, RB, 0, s->tab[s->cur][0], s->tab[s->cur][1], s->tab[s->cur][2],
                   s->d_segs, (int)s->P, make_pack(s->segs), RW, S, fa);
      }
      spg_fq tmp[2];
      SPG_TRY(finish_result(ctx, fa, tiles, 2, tmp));  // (with skip_t0 the first sum is a sum of zeros)
      s->cached[0] = hfq_from(tmp[0]);
      s->cached[1] = hfq_from(tmp[1]);
      s->cached_kind = 2;
    }
    if (s->cached_kind == 2) {
      // G(0), G(2) from the kernel; G(1) from the running claim, G(3) by extrapolation
      const hfq &tau_inv = phase == 0 ? s->tau_x_inv[j] : s->tau_q_inv[j];
      if (!hfq_is_zero(tau_inv) && !hfq_is_zero(c)) {
        if (!s->g_known) {
          s->gclaim = hfq_mul(s->claim, hfq_invert(c));
          s->g_known = true;
        }
        const hfq &G0 = s->cached[0], &G2 = s->cached[1];
        hfq G1 = hfq_mul(hfq_sub(s->gclaim, hfq_mul(line[0], G0)), tau_inv);  // l(1) = tau
        hfq d = hfq_sub(G2, G1);
        hfq G3 = hfq_add(G0, hfq_add(hfq_add(d, d), d));
        ev[0] = hfq_mul(hfq_mul(c, line[0]), G0);
        ev[1] = hfq_mul(hfq_mul(c, line[1]), G2);
        ev[2] = hfq_mul(hfq_mul(c, line[2]), G3);
        s->lastG[0] = G0;
        s->lastG[1] = G1;
        s->lastG[2] = G2;
        s->lastG_valid = true;
        done = true;
      }
      s->cached_kind = 0;  // tau == 0 or c == 0 (probability ~2^-252): recompute from the bound tables below
    } else if (s->cached_kind == 3) {
      for (int t = 0; t < 3; t++) ev[t] = hfq_mul(hfq_mul(c, line[t]), s->cached[t]);
      s->cached_kind = 0;
      done = true;
    }
    if (!done && rows_eligible(s, 0)) {
      unsigned long long tiles = 0, out_total = 0;
      build_tile_segs(s, phase, 0, &tiles, &out_total);
      SPG_TRY(upload_segs(s));
      SPG_TRY(ensure_partials(ctx, (size_t)tiles * 3));
      double pairs = 0;
      for (size_t p = 0; p < s->P; p++) pairs += (double)((phase == 0 ? s->Q[p] : 1) << (s->loglen[p] - 1));
      ctx->next_units = 192.0 * pairs;  // 2 scalars x 3 tables read per pair
      FinishArgs fa = {};
      if (s->pend_inst) {
        ctx->next_units = 288.0 * pairs;
        SPG_LAUNCH(ctx, (k_rows_spmv<3>), (unsigned)tiles, RB, 0, make_spmv_segs(s), log2u(s->pend_max_num_inputs),
                   s->tab[0][0], s->tab[0][1], s->tab[0][2], (int)s->P, make_pack(s->segs), RW, S, ctx->d_partials, 0);
        s->pend_inst = nullptr;
        s->pend_z = nullptr;
      } else {
        fa = finish_args(ctx, tiles);
        SPG_LAUNCH(ctx, k_rows<3>, (unsigned)tiles, RB, 0, s->tab[s->cur][0], s->tab[s->cur][1], s->tab[s->cur][2],
                   s->d_segs, (int)s->P, make_pack(s->segs), RW, S, fa);
      }
      spg_fq tmp[3];
      SPG_TRY(finish_result(ctx, fa, tiles, 3, tmp));
      hfq G0 = hfq_from(tmp[0]), G1 = hfq_from(tmp[1]), G2 = hfq_from(tmp[2]);
      hfq d = hfq_sub(G2, G1);
      hfq G3 = hfq_add(G0, hfq_add(hfq_add(d, d), d));
      ev[0] = hfq_mul(hfq_mul(c, line[0]), G0);
      ev[1] = hfq_mul(hfq_mul(c, line[1]), G2);
      ev[2] = hfq_mul(hfq_mul(c, line[2]), G3);
      s->claim = hfq_add(ev[0], hfq_mul(hfq_mul(c, tau), G1));  // e(0) + e(1): the true claim
      if (verify_now && !hfq_eq(s->claim, s->supplied_claim)) {
        set_error("spg_sc1_round_eval: the supplied claim is not the sum over the tables (the witness does not satisfy the instance?)");
        return SPG_EINVAL;
      }
      s->claim_known = true;
      s->lastG[0] = G0;
      s->lastG[1] = G1;
      s->lastG[2] = G2;
      s->lastG_valid = true;
      done = true;
    }
    if (!done) {
      SPG_TRY(sc1_materialize(s));
      unsigned long long items = 0, out_total = 0;
      build_segs(s, phase, 0, &items, &out_total);
      SPG_TRY(upload_segs(s));
      int grid = grid_for(ctx, items, RB, 4);
      SPG_TRY(ensure_partials(ctx, (size_t)grid * 3));
      ctx->next_units = 192.0 * (double)items;
      SPG_LAUNCH(ctx, k_pair_eval<1>, grid, RB, 0, s->tab[s->cur][0], s->tab[s->cur][1],
                 s->tab[s->cur][2], s->d_segs, (int)s->P, make_pack(s->segs), items, RW, S, ctx->d_partials);
      SPG_TRY(reduce_partials(ctx, ctx->d_partials, grid, 3, ctx->d_result));
      spg_fq tmp[3];
      SPG_TRY(fetch_result(ctx, 3, tmp));
      for (int t = 0; t < 3; t++) ev[t] = hfq_mul(hfq_mul(c, line[t]), hfq_from(tmp[t]));
    }
  }
  for (int t = 0; t < 3; t++) {
    s->last_e[t] = ev[t];
    e[t] = hfq_to(ev[t]);
  }
  s->evaluated = true;
  return SPG_OK;
}

int spg_sc1_round_bind(spg_sc1 *s, const spg_fq *r) {
  spg::DeviceGuard _dev(spg::ctx_of(s));
  SPG_CHECK(s && r, "spg_sc1_round_bind: null argument");
  if (!s->evaluated) {
    set_error("spg_sc1_round_bind: round %zu has not been evaluated", s->round);
    return SPG_ESTATE;
  }
  spg_ctx *ctx = s->ctx;
  int phase = phase_of(s, s->round);
  size_t j = phase_round(s, s->round);
  fq rr;
  memcpy(&rr, r, sizeof rr);
  hfq rh = hfq_from(*r);
  // claim_{j+1} = s_j(r_j)
  if (s->claim_known)
    s->claim = cubic_at(s->last_e[0], hfq_sub(s->claim, s->last_e[0]), s->last_e[1], s->last_e[2], rh);
  if (s->lastG_valid) {
    // G(r) for the quadratic through (0, G0), (1, G1), (2, G2)
    static const hfq two_inv = hfq_invert(hfq_from_u64(2));
    const hfq &G0 = s->lastG[0], &G1 = s->lastG[1], &G2 = s->lastG[2];
    hfq d1 = hfq_sub(G1, G0);
    hfq d2 = hfq_mul(two_inv, hfq_add(hfq_sub(hfq_sub(G2, G1), G1), G0));
    hfq rr1 = hfq_mul(rh, hfq_sub(rh, hfq_one()));
    s->gclaim = hfq_add(G0, hfq_add(hfq_mul(rh, d1), hfq_mul(rr1, d2)));
    s->g_known = true;
  } else {
    s->g_known = false;
  }
  s->lastG_valid = false;
  if (phase == 2) {
    size_t half = s->p_len / 2;
    SPG_LAUNCH(ctx, k_p_bind, 1, 128, 0, s->Ap, s->tab[s->cur][0], s->tab[s->cur][1],
               s->tab[s->cur][2], half, rr);
    s->p_len = half;
  } else {
    size_t n_phase = phase == 0 ? s->nx : s->nq;
    bool next_same = j + 1 < n_phase;
    unsigned minlen = 64;
    for (size_t p = 0; p < s->P; p++) minlen = s->loglen[p] < minlen ? s->loglen[p] : minlen;
    int nxt = s->cur ^ 1;
    unsigned long long items = 0, out_total = 0;
    const fq *RW = phase == 0 ? s->RWx : s->Ap;
    bool rolled = s->fuse && next_same && s->claim_known && rows_eligible(s, 1);
    if (s->tab_lazy && !rolled) SPG_TRY(sc1_canon_tables(s));
    if (rolled) {
      // the bind after this one is this kernel again iff the same conditions hold one round later
      bool lazy_out = j + 2 < n_phase && rows_eligible(s, 1, 1);
      unsigned long long tiles = 0;
      build_tile_segs(s, phase, 1, &tiles, &out_total);
      SPG_CHECK(out_total <= s->cap[nxt], "internal: bound table exceeds buffer");
      SPG_TRY(upload_segs(s));
      const fq *Snext = s_table(s, phase, n_phase - j - 2);
      SPG_TRY(ensure_partials(ctx, (size_t)tiles * 2));
      ctx->next_units = 288.0 * (double)out_total;  // per bound pair: 4 read + 2 written scalars x 3 tables
      FinishArgs fa = finish_args(ctx, tiles);
      if (lazy_out)
        SPG_LAUNCH(ctx, k_rows_rolled<true>, (unsigned)tiles, RB, ROWS_STASH_BYTES, s->tab[s->cur][0], s->tab[s->cur][1],
                   s->tab[s->cur][2], s->tab[nxt][0], s->tab[nxt][1], s->tab[nxt][2], s->d_segs, (int)s->P,
                   make_pack(s->segs), rr, RW, Snext, fa);
      else
        SPG_LAUNCH(ctx, k_rows_rolled<false>, (unsigned)tiles, RB, ROWS_STASH_BYTES, s->tab[s->cur][0], s->tab[s->cur][1],
                   s->tab[s->cur][2], s->tab[nxt][0], s->tab[nxt][1], s->tab[nxt][2], s->d_segs, (int)s->P,
                   make_pack(s->segs), rr, RW, Snext, fa);
      s->tab_lazy = lazy_out;
      s->tab_lazy_n = (size_t)out_total;
      spg_fq tmp[2];
      SPG_TRY(finish_result(ctx, fa, tiles, 2, tmp));
      s->cached[0] = hfq_from(tmp[0]);
      s->cached[1] = hfq_from(tmp[1]);
      s->cached_kind = 2;
    } else if (s->fuse && next_same && minlen >= 2) {
      build_segs(s, phase, 1, &items, &out_total);
      SPG_CHECK(out_total <= s->cap[nxt], "internal: bound table exceeds buffer");
      SPG_TRY(upload_segs(s));
      const fq *Snext = s_table(s, phase, n_phase - j - 2);
      const bool split = items <= SPLIT_MAX_ITEMS;  // a late round: latency, not throughput (rounds.cuh)
      int grid = split ? (int)((items + SPLIT_ITEMS_PER_BLOCK - 1) / SPLIT_ITEMS_PER_BLOCK) : grid_for(ctx, items, RB, 4);
      SPG_TRY(ensure_partials(ctx, (size_t)grid * 3));
      FinishArgs fa = finish_args(ctx, grid);
      ctx->next_units = 576.0 * (double)items;  // 4 read + 2 written scalars x 3 tables per item
      if (split) {
        SplitTabs T = {{s->tab[s->cur][0], s->tab[s->cur][1], s->tab[s->cur][2]}, {s->tab[nxt][0], s->tab[nxt][1], s->tab[nxt][2]}};
        SPG_LAUNCH(ctx, (k_quad_split<3, 1>), grid, 128, 0, T, s->d_segs, (int)s->P, make_pack(s->segs), items, rr, RW, Snext, fa);
      } else {
        SPG_LAUNCH(ctx, k_quad_bind_eval<1>, grid, RB, 0, s->tab[s->cur][0], s->tab[s->cur][1],
                   s->tab[s->cur][2], s->tab[nxt][0], s->tab[nxt][1], s->tab[nxt][2], s->d_segs,
                   (int)s->P, make_pack(s->segs), items, rr, RW, Snext, fa);
      }
      spg_fq tmp[3];
      SPG_TRY(finish_result(ctx, fa, grid, 3, tmp));
      for (int t = 0; t < 3; t++) s->cached[t] = hfq_from(tmp[t]);
      s->cached_kind = 3;
    } else {
      build_segs(s, phase, 0, &items, &out_total);
      SPG_CHECK(out_total <= s->cap[nxt], "internal: bound table exceeds buffer");
      SPG_TRY(upload_segs(s));
      ctx->next_units = 288.0 * (double)items;
      SPG_LAUNCH(ctx, k_pair_bind, grid_for(ctx, items, RB, 8), RB, 0, s->tab[s->cur][0],
                 s->tab[s->cur][1], s->tab[s->cur][2], s->tab[nxt][0], s->tab[nxt][1],
                 s->tab[nxt][2], s->d_segs, (int)s->P, make_pack(s->segs), items, rr);
    }
    s->cur = nxt;
    for (size_t p = 0; p < s->P; p++)
      if (s->loglen[p] > 0) s->loglen[p]--;
    const hfq &tau = phase == 0 ? s->tau_x[j] : s->tau_q[j];
    if (phase == 0) s->cx = hfq_mul(s->cx, hfq_eq1(tau, rh));
    else s->cq = hfq_mul(s->cq, hfq_eq1(tau, rh));
    // x rounds done: rows of the q phase are whole instances
    if (phase == 0 && j + 1 == s->nx)
      for (size_t p = 0; p < s->P; p++) s->loglen[p] = log2u(s->Q[p]);
  }
  s->round++;
  s->evaluated = false;
  return SPG_OK;
}

int spg_sc1_run_rounds(spg_sc1 *s, size_t num_rounds, const spg_fq *challenges, spg_fq *evals_out) {
  spg::DeviceGuard _dev(spg::ctx_of(s));
  SPG_CHECK(s && challenges && evals_out, "spg_sc1_run_rounds: null argument");
  static const bool trace = getenv("SPG_TRACE_ROUNDS") != nullptr;
  for (size_t j = 0; j < num_rounds; j++) {
    auto t0 = std::chrono::steady_clock::now();
    SPG_TRY(spg_sc1_round_eval(s, evals_out + 3 * j));
    auto t1 = std::chrono::steady_clock::now();
    SPG_TRY(spg_sc1_round_bind(s, challenges + j));
    auto t2 = std::chrono::steady_clock::now();
    if (trace)
      fprintf(stderr, "[spg] round %zu: eval %.1f us, bind %.1f us\n", j,
              std::chrono::duration<double, std::micro>(t1 - t0).count(),
              std::chrono::duration<double, std::micro>(t2 - t1).count());
  }
  return SPG_OK;
}

// ---------------------------------------------------------------- host mailbox of a sharded proof
// The mailbox is the shared-memory segment spartan_parallel_b200.parallel.ShmComm maps:
//   slot(b, r) = mailbox + (b * world + r) * slot_stride; word 0 = sequence number, data at +64
// double-buffered by the parity of the call counter *calls (shared with the Python side).
// A rank that fails publishes the poison sequence in both of its slots, so that its peers
// return an error instead of spinning forever; every wait also has a deadline
// (SPG_MAILBOX_TIMEOUT_S, default 120 s).
extern "C++" {
namespace spg {
constexpr uint64_t MAILBOX_POISON = UINT64_MAX;

double mailbox_timeout_s() {
  static const double t = [] {
    const char *e = getenv("SPG_MAILBOX_TIMEOUT_S");
    double v = e ? atof(e) : 120.0;
    return v > 0 ? v : 120.0;
  }();
  return t;
}

int mailbox_exchange(char *base, size_t slot_stride, int rank, int world, uint64_t c, const void *mine, size_t nbytes,
                     void *out) {
  size_t b = c & 1;
  char *me = base + (b * world + rank) * slot_stride;
  memcpy(me + 64, mine, nbytes);
  __atomic_store_n((uint64_t *)me, c, __ATOMIC_RELEASE);
  auto t0 = std::chrono::steady_clock::now();
  for (int r = 0; r < world; r++) {
    char *slot = base + (b * world + r) * slot_stride;
    unsigned spins = 0;
    for (;;) {
      uint64_t seq = __atomic_load_n((uint64_t *)slot, __ATOMIC_ACQUIRE);
      if (seq == c) break;
      if (seq == MAILBOX_POISON) {
        set_error("sharded proof: rank %d reported a failure (see its own error message)", r);
        return SPG_ESTATE;
      }
      spin_pause();
      if ((++spins & 0xfffu) == 0 &&
          std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > mailbox_timeout_s()) {
        set_error("sharded proof: rank %d did not publish exchange %llu within %.0f s", r, (unsigned long long)c,
                  mailbox_timeout_s());
        return SPG_ESTATE;
      }
    }
    memcpy((char *)out + (size_t)r * nbytes, slot + 64, nbytes);
  }
  return SPG_OK;
}
}  // namespace spg
}  // extern "C++"

void spg_mailbox_poison(void *mailbox, size_t slot_stride, int rank, int world) {
  if (!mailbox || rank < 0 || rank >= world) return;
  for (size_t b = 0; b < 2; b++)
    __atomic_store_n((uint64_t *)((char *)mailbox + (b * world + rank) * slot_stride), MAILBOX_POISON, __ATOMIC_RELEASE);
}

int spg_mailbox_all_gather(void *mailbox, size_t slot_stride, int rank, int world, uint64_t *calls, const void *data,
                           size_t nbytes, void *out) {
  SPG_CHECK(mailbox && calls && data && out, "spg_mailbox_all_gather: null argument");
  SPG_CHECK(world >= 1 && rank >= 0 && rank < world && slot_stride >= 64 + nbytes, "spg_mailbox_all_gather: bad mailbox geometry");
  uint64_t c = ++*calls;
  int rc = mailbox_exchange((char *)mailbox, slot_stride, rank, world, c, data, nbytes, out);
  if (rc != SPG_OK) spg_mailbox_poison(mailbox, slot_stride, rank, world);
  return rc;
}

// Rounds of a proof whose proofs are sharded over `world` processes of one host (one GPU
// each): every rank evaluates its shard, the 3 partial evaluations are exchanged through the
// mailbox and summed with Scalar::add, and every rank binds with the same challenge.
int spg_sc1_run_rounds_sharded(spg_sc1 *s, size_t num_rounds, const spg_fq *challenges, spg_fq *evals_out,
                               void *mailbox, size_t slot_stride, int rank, int world, uint64_t *calls) {
  spg::DeviceGuard _dev(spg::ctx_of(s));
  SPG_CHECK(s && challenges && evals_out && mailbox && calls, "spg_sc1_run_rounds_sharded: null argument");
  SPG_CHECK(world >= 1 && world <= 64 && rank >= 0 && rank < world && slot_stride >= 64 + 3 * sizeof(spg_fq),
            "spg_sc1_run_rounds_sharded: bad mailbox geometry");
  char *base = (char *)mailbox;
  static const bool trace = getenv("SPG_TRACE_ROUNDS") != nullptr;
  int rc = SPG_OK;
  for (size_t j = 0; j < num_rounds && rc == SPG_OK; j++) {
    spg_fq part[3], all[64 * 3];
    auto t0 = std::chrono::steady_clock::now();
    if ((rc = spg_sc1_round_eval(s, part)) != SPG_OK) break;
    auto t1 = std::chrono::steady_clock::now();
    if ((rc = mailbox_exchange(base, slot_stride, rank, world, ++*calls, part, sizeof part, all)) != SPG_OK) break;
    hfq acc[3] = {hfq_zero(), hfq_zero(), hfq_zero()};
    for (int r = 0; r < world; r++)
      for (int t = 0; t < 3; t++) acc[t] = hfq_add(acc[t], hfq_from(all[3 * r + t]));
    for (int t = 0; t < 3; t++) evals_out[3 * j + t] = hfq_to(acc[t]);
    auto t2 = std::chrono::steady_clock::now();
    rc = spg_sc1_round_bind(s, challenges + j);
    if (trace && rank == 0) {
      auto t3 = std::chrono::steady_clock::now();
      fprintf(stderr, "[spg] sharded round %zu: eval %.1f us, mailbox %.1f us, bind %.1f us\n", j,
              std::chrono::duration<double, std::micro>(t1 - t0).count(),
              std::chrono::duration<double, std::micro>(t2 - t1).count(),
              std::chrono::duration<double, std::micro>(t3 - t2).count());
    }
  }
  if (rc != SPG_OK) spg_mailbox_poison(mailbox, slot_stride, rank, world);  // peers abort instead of spinning
  return rc;
}

// The cross-rank end of a sharded proof. After the rounds a rank can run alone, every table of a
// proof sharded over G ranks is one scalar per rank; the last log2(G) rounds combine those G
// scalars. Like the per-round sum of the ranks' partial evaluations they run on the host, next to the
// mailbox the scalars arrived through: G <= 64 entries, microseconds of exact field arithmetic,
// against ~0.2 ms for standing up a device prover for them. Natural order, low bit first, the eq table
// of the rank bits bound along with the tables (the reference binds its eq tables the same way,
// src/sumcheck.rs:1265-1275):
//   e_j(t) = scale * sum_i E_j(t)[i] * (A_j(t)[i] * B_j(t)[i] - C_j(t)[i]),  t = 0, 2, 3.
// state: [E | A | B | C], G scalars each, overwritten in place; round j works on the first G >> j of each.
int spg_sc1_host_tail_eval(const spg_fq *state, size_t G, size_t len, const spg_fq *scale, spg_fq e[3]) {
  SPG_CHECK(state && scale && e, "spg_sc1_host_tail_eval: null argument");
  SPG_CHECK(is_pow2(G) && G <= 64 && is_pow2(len) && len >= 2 && len <= G, "spg_sc1_host_tail_eval: bad sizes %zu / %zu", len, G);
  hfq acc[3] = {hfq_zero(), hfq_zero(), hfq_zero()};
  for (size_t i = 0; i < len / 2; i++) {
    hfq lo[4], d[4];
    for (int k = 0; k < 4; k++) {
      lo[k] = hfq_from(state[k * G + 2 * i]);
      d[k] = hfq_sub(hfq_from(state[k * G + 2 * i + 1]), lo[k]);
    }
    for (int pt = 0; pt < 3; pt++) {
      if (pt >= 1)
        for (int k = 0; k < 4; k++) {
          lo[k] = hfq_add(lo[k], d[k]);
          if (pt == 1) lo[k] = hfq_add(lo[k], d[k]);  // 0 -> 2, then 2 -> 3
        }
      acc[pt] = hfq_add(acc[pt], hfq_mul(lo[0], hfq_sub(hfq_mul(lo[1], lo[2]), lo[3])));
    }
  }
  hfq sc = hfq_from(*scale);
  for (int pt = 0; pt < 3; pt++) e[pt] = hfq_to(hfq_mul(sc, acc[pt]));
  return SPG_OK;
}

int spg_sc1_host_tail_bind(spg_fq *state, size_t G, size_t len, const spg_fq *r) {
  SPG_CHECK(state && r, "spg_sc1_host_tail_bind: null argument");
  SPG_CHECK(is_pow2(G) && G <= 64 && is_pow2(len) && len >= 2 && len <= G, "spg_sc1_host_tail_bind: bad sizes %zu / %zu", len, G);
  hfq rr = hfq_from(*r);
  for (int k = 0; k < 4; k++)
    for (size_t i = 0; i < len / 2; i++) {
      hfq lo = hfq_from(state[k * G + 2 * i]), hi = hfq_from(state[k * G + 2 * i + 1]);
      state[k * G + i] = hfq_to(hfq_add(lo, hfq_mul(rr, hfq_sub(hi, lo))));
    }
  return SPG_OK;
}

int spg_sc1_final(spg_sc1 *s, spg_fq claims[4]) {
  spg::DeviceGuard _dev(spg::ctx_of(s));
  SPG_CHECK(s && claims, "spg_sc1_final: null argument");
  if (s->round != spg_sc1_num_rounds(s)) {
    set_error("spg_sc1_final: %zu of %zu rounds bound", s->round, spg_sc1_num_rounds(s));
    return SPG_ESTATE;
  }
  spg_ctx *ctx = s->ctx;
  SPG_TRY(sc1_materialize(s));
  // with zero q rounds the x phase never re-keys the segments; everything is length one anyway
  spg_fq h[4];
  const fq *heads[4] = {s->Ap, s->tab[s->cur][0], s->tab[s->cur][1], s->tab[s->cur][2]};
  SPG_TRY(gather_heads(ctx, heads, 4, h));
  hfq ap;
  memcpy(&ap, &h[0], sizeof ap);
  claims[0] = hfq_to(hfq_mul(hfq_mul(hfq_mul(ap, s->cq), s->cx), s->scale));
  for (int k = 0; k < 3; k++) memcpy(&claims[1 + k], &h[1 + k], sizeof(spg_fq));
  return SPG_OK;
}

int spg_sc1_debug_tables(spg_sc1 *s, spg_fq *Az, spg_fq *Bz, spg_fq *Cz, size_t cap, size_t *n) {
  spg::DeviceGuard _dev(spg::ctx_of(s));
  SPG_CHECK(s && n, "spg_sc1_debug_tables: null argument");
  SPG_TRY(sc1_materialize(s));
  size_t total = 0;
  int phase = s->round < spg_sc1_num_rounds(s) ? phase_of(s, s->round) : 2;
  if (phase == 2) total = s->p_len < s->P ? s->p_len : s->P;
  else
    for (size_t p = 0; p < s->P; p++) total += ((phase == 0 ? s->Q[p] : 1) << s->loglen[p]);
  *n = total;
  if (!Az) return SPG_OK;
  SPG_CHECK(cap >= total, "spg_sc1_debug_tables: capacity %zu < %zu", cap, total);
  spg_fq *dst[3] = {Az, Bz, Cz};
  for (int k = 0; k < 3; k++)
    SPG_CUDA(cudaMemcpyAsync(dst[k], s->tab[s->cur][k], total * sizeof(fq), cudaMemcpyDeviceToHost,
                             s->ctx->stream));
  SPG_CUDA(cudaStreamSynchronize(s->ctx->stream));
  return SPG_OK;
}

void spg_sc1_destroy(spg_sc1 *s) {
  spg::DeviceGuard _dev(spg::ctx_of(s));
  if (!s) return;
  for (int b = 0; b < 2; b++)
    for (int k = 0; k < 3; k++) dev_free(s->ctx, s->tab[b][k]);
  dev_free(s->ctx, s->Sx);
  dev_free(s->ctx, s->Sq);
  dev_free(s->ctx, s->Ap);
  dev_free(s->ctx, s->RWx);
  dev_free(s->ctx, s->d_segs);
  dev_free(s->ctx, s->d_rw_off);
  dev_free(s->ctx, s->d_Qp);
  delete s;
}

}  // extern "C"
