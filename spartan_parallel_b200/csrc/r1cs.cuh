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

}  // namespace spg
