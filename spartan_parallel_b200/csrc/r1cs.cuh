// Device-side R1CS instance, witness sections and z_mat (shared declarations).
#pragma once
#include "common.cuh"

namespace spg {

// compressed sparse rows (or columns): entries of major index i are ptr[i]..ptr[i+1].
// idx holds the minor index with bit 31 set when val == 1 (Montgomery R): a unit
// coefficient needs no multiplication, and most R1CS coefficients are units.
struct Csx {
  uint32_t *ptr = nullptr;
  uint32_t *idx = nullptr;
  uint32_t *major = nullptr;  // expanded major index per entry (COO view)
  fq *val = nullptr;
  size_t n_major = 0;
  size_t nnz = 0;
};

constexpr uint32_t UNIT_FLAG = 0x80000000u;

// z_mat[p][q][w][y] is never materialised: it is a view over the witness sections,
//   z[p][q][w][y] = y < copy ? ptr[q * q_stride + y] : 0      (src/r1csproof.rs:282-290)
// with q_stride = 0 for a short section (one row shared by every proof).
struct SecView {
  const fq *ptr;
  unsigned long long q_stride;
  unsigned long long copy;
};

__device__ __forceinline__ fq z_load(const SecView &v, size_t q, size_t y) {
  return y < v.copy ? fq_load(v.ptr + q * v.q_stride + y) : fq_zero();
}

struct CsxView {
  const uint32_t *ptr, *idx;
  const fq *val;
};

__device__ __forceinline__ fq spmv_row(const CsxView &M, unsigned int x, const SecView *__restrict__ secs,
                                       size_t q, unsigned int log_ymax) {
  fq acc = fq_zero();
  for (uint32_t e = M.ptr[x]; e < M.ptr[x + 1]; e++) {
    uint32_t c = M.idx[e];
    bool unit = c & UNIT_FLAG;
    c &= ~UNIT_FLAG;
    size_t w = c >> log_ymax, y = c & ((1u << log_ymax) - 1);
    fq zz = z_load(secs[w], q, y);
    acc = fq_add(acc, unit ? zz : fq_mul(fq_load(M.val + e), zz));
  }
  return acc;
}

// rest of a row after its first entry (which the caller prefetched)
__device__ __forceinline__ fq spmv_row_tail(const CsxView &M, uint32_t e0, uint32_t e1, fq acc,
                                            const SecView *__restrict__ secs, size_t q, unsigned int log_ymax) {
  for (uint32_t e = e0; e < e1; e++) {
    uint32_t c = M.idx[e];
    bool unit = c & UNIT_FLAG;
    c &= ~UNIT_FLAG;
    size_t w = c >> log_ymax, y = c & ((1u << log_ymax) - 1);
    fq zz = z_load(secs[w], q, y);
    acc = fq_add(acc, unit ? zz : fq_mul(fq_load(M.val + e), zz));
  }
  return acc;
}

// the first entries of the three matrices' rows x are fetched together (three independent
// ptr -> idx -> z chains in flight), then each row is finished: Az[x], Bz[x], Cz[x] for proof q
struct CsxView3 {
  CsxView M[3];
};
// row heads: entry range and first column index of row x in each of the three matrices
struct RowHead3 {
  uint32_t e0[3], e1[3], c[3];
};
__device__ __forceinline__ RowHead3 spmv_head3(const CsxView3 &V, unsigned int x) {
  RowHead3 h;
#pragma unroll
  for (int m = 0; m < 3; m++) {
    h.e0[m] = V.M[m].ptr[x];
    h.e1[m] = V.M[m].ptr[x + 1];
  }
#pragma unroll
  for (int m = 0; m < 3; m++) h.c[m] = h.e0[m] < h.e1[m] ? V.M[m].idx[h.e0[m]] : 0u;
  return h;
}
// Everything a row needs beyond "one entry with a unit coefficient": the product with the first
// coefficient and the remaining entries. Out of line on purpose (COLD = true): inlined six times
// into the fused first-round kernel these loops, each with its own Montgomery product, are
// ~40 KB of SASS threaded through the hot path of a kernel whose common case (R1CS rows are
// mostly single unit entries per matrix) never executes them.
static __device__ __noinline__ fq spmv_row_rest(const CsxView M, uint32_t e0, uint32_t e1, uint32_t c0, fq first,
                                         const SecView *__restrict__ secs, size_t q, unsigned int log_ymax) {
  if (!(c0 & UNIT_FLAG)) first = fq_mul(fq_load(M.val + e0), first);
  return spmv_row_tail(M, e0 + 1, e1, first, secs, q, log_ymax);
}

template <bool COLD = false>
__device__ __forceinline__ void spmv_finish3(const CsxView3 &V, const RowHead3 &h, const SecView *__restrict__ secs,
                                             size_t q, unsigned int log_ymax, fq (&out)[3]) {
  const uint32_t ymask = (1u << log_ymax) - 1;
#pragma unroll
  for (int m = 0; m < 3; m++) {
    uint32_t cc = h.c[m] & ~UNIT_FLAG;
    out[m] = h.e0[m] < h.e1[m] ? z_load(secs[cc >> log_ymax], q, cc & ymask) : fq_zero();
  }
#pragma unroll
  for (int m = 0; m < 3; m++) {
    if (COLD) {
      if (h.e0[m] < h.e1[m] && (!(h.c[m] & UNIT_FLAG) || h.e0[m] + 1 < h.e1[m]))
        out[m] = spmv_row_rest(V.M[m], h.e0[m], h.e1[m], h.c[m], out[m], secs, q, log_ymax);
    } else {
      if (h.e0[m] < h.e1[m] && !(h.c[m] & UNIT_FLAG)) out[m] = fq_mul(fq_load(V.M[m].val + h.e0[m]), out[m]);
      if (h.e0[m] + 1 < h.e1[m]) out[m] = spmv_row_tail(V.M[m], h.e0[m] + 1, h.e1[m], out[m], secs, q, log_ymax);
    }
  }
}
__device__ __forceinline__ void spmv_rows3(const CsxView3 &V, unsigned int x, const SecView *__restrict__ secs,
                                           size_t q, unsigned int log_ymax, fq (&out)[3]) {
  RowHead3 h = spmv_head3(V, x);
  spmv_finish3(V, h, secs, q, log_ymax, out);
}

}  // namespace spg

struct spg_r1cs {
  spg_ctx *ctx = nullptr;
  size_t num_instances = 0, max_num_cons = 0, num_vars = 0;
  std::vector<size_t> num_cons;
  std::vector<spg::Csx> by_row, by_col;      // 3 per instance: A, B, C
  std::vector<uint32_t> max_col;             // per matrix: largest column index used
  std::vector<std::vector<uint32_t>> h_cols;  // host copy of the column indices (validation)
};

struct spg_witness {
  spg_ctx *ctx = nullptr;
  size_t num_instances = 0;
  std::vector<size_t> num_proofs, num_inputs, off;  // off[p]: first scalar of w_mat[p]
  spg::fq *d = nullptr;
  size_t total = 0;
  std::vector<spg_vec *> views;
  cudaEvent_t ready = nullptr;  // async upload: recorded on the copy stream, awaited by the first consumer
};

struct spg_zmat {
  spg_ctx *ctx = nullptr;
  size_t P = 0, W = 0;
  std::vector<size_t> num_proofs, num_inputs;
  spg::SecView *views = nullptr;  // device array [P][W]; the witness handles must outlive the z_mat
};

namespace spg {

// orders the compute stream after a pending asynchronous upload of the section
int witness_wait(spg_witness *w);

CsxView csx_view(const Csx &c);
// the shape / range checks of multiply_vec_block without the launch
int r1cs_spmv_validate(const spg_r1cs *inst, const spg_zmat *z, size_t P, const size_t *num_proofs,
                       const size_t *num_cons, size_t max_num_inputs);
// multiply_vec_block (src/r1csinstance.rs:363-436): Az/Bz/Cz in natural ragged [p][q][x] order
int r1cs_multiply_vec_block(spg_ctx *ctx, const spg_r1cs *inst, const spg_zmat *z, size_t P,
                            const size_t *num_proofs, const size_t *num_cons, size_t max_num_inputs,
                            fq *Az, fq *Bz, fq *Cz);

// compute_eval_table_sparse_disjoint_rounds + the r_A/r_B/r_C combination
// (src/r1csinstance.rs:484-534, src/r1csproof.rs:431-456): out[p_inst][w][y], natural y,
// with per-instance offsets out_off[p_inst] and row length num_inputs[p_inst]
int r1cs_abc_table(spg_ctx *ctx, const spg_r1cs *inst, const fq *evals_rx, size_t num_segs,
                   size_t max_num_cols, const size_t *num_cols, const size_t *out_off,
                   const spg_fq *r_A, const spg_fq *r_B, const spg_fq *r_C, fq *out);

// entries [t0, t0 + count) of a shared / single instance's ABC table (the y-sharded phase 2)
int r1cs_abc_slice(spg_ctx *ctx, const spg_r1cs *inst, const fq *evals_rx, size_t num_segs, size_t max_num_cols,
                   size_t t0, size_t count, const spg_fq *r_A, const spg_fq *r_B, const spg_fq *r_C, fq *out);

}  // namespace spg
