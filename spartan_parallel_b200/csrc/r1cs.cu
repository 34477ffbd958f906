// R1CS instance on the device: SpMV (Az, Bz, Cz), the transposed SpMV that builds the
// phase-2 ABC table, sparse evaluation, witness sections and z_mat assembly.
//   reference: src/r1csinstance.rs:89-182, 363-436, 484-534, 583-595
//              src/sparse_mlpoly.rs:427-472, 524-541
//              src/r1csproof.rs:278-293 (z_mat), 431-456 (ABC)
// The reference scatters `+=` over an unsorted COO list; field addition is exact, so
// regrouping by row (CSR) or by column (CSC) gives bit-identical sums with no atomics.
#include "r1cs.cuh"

#include <algorithm>

namespace spg {

// ---------------------------------------------------------------- kernels
// thread t = q * X + x computes row x of A, B, C against z[p][q]. The kernel is bound by
// the latency of its dependent loads (row pointer -> column index -> z), so the three
// matrices' first entries are fetched together (spmv_rows3): three independent chains in
// flight per thread instead of one after the other (most R1CS rows have one or two entries).
__global__ void k_spmv3(CsxView A, CsxView B, CsxView C, const SecView *__restrict__ secs, size_t Q,
                        unsigned int log_x, unsigned int log_ymax,
                        fq *__restrict__ outA, fq *__restrict__ outB, fq *__restrict__ outC) {
  size_t total = Q << log_x;
  CsxView3 V;
  V.M[0] = A;
  V.M[1] = B;
  V.M[2] = C;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (size_t)gridDim.x * blockDim.x) {
    size_t q = t >> log_x;
    unsigned int x = (unsigned int)(t & (((size_t)1 << log_x) - 1));
    fq r[3];
    spmv_rows3(V, x, secs, q, log_ymax, r);
    fq_store(outA + t, r[0]);
    fq_store(outB + t, r[1]);
    fq_store(outC + t, r[2]);
  }
}

__device__ __forceinline__ fq spmv_col(const CsxView &M, uint32_t c, const fq *__restrict__ rx) {
  fq acc = fq_zero();
  for (uint32_t e = M.ptr[c]; e < M.ptr[c + 1]; e++) {
    uint32_t r = M.idx[e];
    bool unit = r & UNIT_FLAG;
    r &= ~UNIT_FLAG;
    fq t = fq_load(rx + r);
    acc = fq_add(acc, unit ? t : fq_mul(t, fq_load(M.val + e)));
  }
  return acc;
}

// out[w * num_cols + y] = r_A * sum_A + r_B * sum_B + r_C * sum_C over column w*max_cols + y
// (entries [t0, t0 + count) of the table only: a rank of a y-sharded phase 2 builds its own slice)
__global__ void k_abc_table(CsxView A, CsxView B, CsxView C, const fq *__restrict__ rx,
                            size_t t0, size_t count, unsigned int log_max_cols, unsigned int log_cols,
                            fq rA, fq rB, fq rC, fq *__restrict__ out) {
  for (size_t l = (size_t)blockIdx.x * blockDim.x + threadIdx.x; l < count;
       l += (size_t)gridDim.x * blockDim.x) {
    size_t t = t0 + l;
    size_t w = t >> log_cols, y = t & (((size_t)1 << log_cols) - 1);
    uint32_t c = (uint32_t)((w << log_max_cols) + y);
    fq a = fq_mul(rA, spmv_col(A, c, rx));
    fq b = fq_mul(rB, spmv_col(B, c, rx));
    fq cc = fq_mul(rC, spmv_col(C, c, rx));
    fq_store(out + l, fq_add(fq_add(a, b), cc));
  }
}

// sum over entries of trx[row] * try[col] * val (evaluate_with_tables, sparse_mlpoly.rs:427-436)
__global__ void k_sparse_eval(const uint32_t *__restrict__ row, const uint32_t *__restrict__ col,
                              const fq *__restrict__ val, size_t nnz, const fq *__restrict__ trx,
                              const fq *__restrict__ try_, fq *__restrict__ partials) {
  __shared__ fq sm[32];
  fq acc[1] = {fq_zero()};
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < nnz;
       e += (size_t)gridDim.x * blockDim.x) {
    uint32_t c = col[e] & ~UNIT_FLAG;
    fq t = fq_mul(fq_load(trx + row[e]), fq_load(try_ + c));
    acc[0] = fq_add(acc[0], (col[e] & UNIT_FLAG) ? t : fq_mul(t, fq_load(val + e)));
  }
  block_sum<1>(acc, sm);
  if (threadIdx.x == 0) partials[blockIdx.x] = acc[0];
}

// ---------------------------------------------------------------- host helpers
static int build_csx(spg_ctx *ctx, size_t n_major, size_t nnz, const uint32_t *major,
                     const uint32_t *minor, const spg_fq *vals, Csx *out) {
  std::vector<uint32_t> ptr(n_major + 1, 0), idx(nnz), maj(nnz);
  std::vector<spg_fq> val(nnz ? nnz : 1);
  for (size_t e = 0; e < nnz; e++) ptr[major[e] + 1]++;
  for (size_t i = 0; i < n_major; i++) ptr[i + 1] += ptr[i];
  std::vector<uint32_t> fill(ptr.begin(), ptr.end() - 1);
  hfq one = hfq_one();
  for (size_t e = 0; e < nnz; e++) {
    uint32_t pos = fill[major[e]]++;
    bool unit = memcmp(&vals[e], &one, 32) == 0;
    idx[pos] = minor[e] | (unit ? UNIT_FLAG : 0u);
    maj[pos] = major[e];
    val[pos] = vals[e];
  }
  out->n_major = n_major;
  out->nnz = nnz;
  SPG_CUDA(cudaMalloc(&out->ptr, (n_major + 1) * sizeof(uint32_t)));
  SPG_CUDA(cudaMalloc(&out->idx, (nnz ? nnz : 1) * sizeof(uint32_t)));
  SPG_CUDA(cudaMalloc(&out->major, (nnz ? nnz : 1) * sizeof(uint32_t)));
  SPG_CUDA(cudaMalloc(&out->val, (nnz ? nnz : 1) * sizeof(fq)));
  SPG_CUDA(cudaMemcpyAsync(out->ptr, ptr.data(), (n_major + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
  if (nnz) {
    SPG_CUDA(cudaMemcpyAsync(out->idx, idx.data(), nnz * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
    SPG_CUDA(cudaMemcpyAsync(out->major, maj.data(), nnz * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
    SPG_CUDA(cudaMemcpyAsync(out->val, val.data(), nnz * sizeof(fq), cudaMemcpyHostToDevice, ctx->stream));
  }
  SPG_CUDA(cudaStreamSynchronize(ctx->stream));
  return SPG_OK;
}

static void free_csx(Csx &c) {
  if (c.ptr) cudaFree(c.ptr);
  if (c.idx) cudaFree(c.idx);
  if (c.major) cudaFree(c.major);
  if (c.val) cudaFree(c.val);
  c = Csx();
}

CsxView csx_view(const Csx &c) { return CsxView{c.ptr, c.idx, c.val}; }
static CsxView view(const Csx &c) { return csx_view(c); }

int r1cs_spmv_validate(const spg_r1cs *inst, const spg_zmat *z, size_t P, const size_t *num_proofs,
                       const size_t *num_cons, size_t max_num_inputs) {
  SPG_CHECK(inst->num_instances == 1 || inst->num_instances == P,
            "multiply_vec_block: instance has %zu blocks, proving %zu", inst->num_instances, P);
  SPG_CHECK(z->P == P, "multiply_vec_block: z_mat has %zu instances, expected %zu", z->P, P);
  SPG_CHECK(is_pow2(max_num_inputs), "multiply_vec_block: max_num_inputs must be a power of two");
  unsigned log_ymax = log2u(max_num_inputs);
  for (size_t p = 0; p < P; p++) {
    size_t pi = inst->num_instances == 1 ? 0 : p;
    SPG_CHECK(num_cons[p] == inst->num_cons[pi], "multiply_vec_block: num_cons[%zu] = %zu, instance has %zu",
              p, num_cons[p], inst->num_cons[pi]);
    SPG_CHECK(z->num_proofs[p] == num_proofs[p], "multiply_vec_block: z_mat num_proofs mismatch at %zu", p);
    size_t Yp = z->num_inputs[p];
    // the reference indexes z[col / max][col % max] and would panic out of range
    for (int m = 0; m < 3; m++) {
      uint32_t mc = inst->max_col[3 * pi + m];
      if (inst->by_row[3 * pi + m].nnz == 0) continue;
      SPG_CHECK((mc >> log_ymax) < z->W, "multiply_vec_block: column %u addresses witness section %u >= %zu",
                mc, mc >> log_ymax, z->W);
      if (Yp < max_num_inputs)
        for (uint32_t c : inst->h_cols[3 * pi + m])
          SPG_CHECK((c & (max_num_inputs - 1)) < Yp,
                    "multiply_vec_block: column %u exceeds num_inputs[%zu] = %zu", c, p, Yp);
    }
  }
  return SPG_OK;
}

int r1cs_multiply_vec_block(spg_ctx *ctx, const spg_r1cs *inst, const spg_zmat *z, size_t P,
                            const size_t *num_proofs, const size_t *num_cons, size_t max_num_inputs,
                            fq *Az, fq *Bz, fq *Cz) {
  SPG_TRY(r1cs_spmv_validate(inst, z, P, num_proofs, num_cons, max_num_inputs));
  unsigned log_ymax = log2u(max_num_inputs);
  size_t off = 0;
  for (size_t p = 0; p < P; p++) {
    size_t pi = inst->num_instances == 1 ? 0 : p;
    size_t items = num_proofs[p] * num_cons[p];
    ctx->next_units = 32.0 * (double)items * 3.0 * 2.0;  // >= one z read + one write per (row, q, matrix)
    SPG_LAUNCH(ctx, k_spmv3, grid_for(ctx, items, 256), 256, 0, view(inst->by_row[3 * pi]),
               view(inst->by_row[3 * pi + 1]), view(inst->by_row[3 * pi + 2]), z->views + p * z->W,
               num_proofs[p], log2u(num_cons[p]), log_ymax, Az + off, Bz + off, Cz + off);
    off += items;
  }
  return SPG_OK;
}

int r1cs_abc_table(spg_ctx *ctx, const spg_r1cs *inst, const fq *evals_rx, size_t num_segs,
                   size_t max_num_cols, const size_t *num_cols, const size_t *out_off,
                   const spg_fq *r_A, const spg_fq *r_B, const spg_fq *r_C, fq *out) {
  fq rA, rB, rC;
  memcpy(&rA, r_A, 32);
  memcpy(&rB, r_B, 32);
  memcpy(&rC, r_C, 32);
  unsigned log_max = log2u(max_num_cols);
  for (size_t p = 0; p < inst->num_instances; p++) {
    SPG_CHECK(is_pow2(num_cols[p]) && num_cols[p] <= max_num_cols, "abc_table: bad num_cols[%zu]", p);
    for (int m = 0; m < 3; m++) {
      if (inst->by_col[3 * p + m].nnz == 0) continue;
      uint32_t mc = inst->max_col[3 * p + m];
      SPG_CHECK((mc >> log_max) < num_segs, "abc_table: column %u addresses segment >= %zu", mc, num_segs);
      if (num_cols[p] < max_num_cols)
        for (uint32_t c : inst->h_cols[3 * p + m])
          SPG_CHECK((c & (max_num_cols - 1)) < num_cols[p], "abc_table: column %u exceeds num_cols[%zu]", c, p);
    }
    size_t items = num_segs * num_cols[p];
    SPG_LAUNCH(ctx, k_abc_table, grid_for(ctx, items, 128), 128, 0, view(inst->by_col[3 * p]),
               view(inst->by_col[3 * p + 1]), view(inst->by_col[3 * p + 2]), evals_rx, (size_t)0, items, log_max,
               log2u(num_cols[p]), rA, rB, rC, out + out_off[p]);
  }
  return SPG_OK;
}

// entries [t0, t0 + count) of instance 0's [w][y] table (num_cols = max_num_cols columns per segment)
int r1cs_abc_slice(spg_ctx *ctx, const spg_r1cs *inst, const fq *evals_rx, size_t num_segs, size_t max_num_cols,
                   size_t t0, size_t count, const spg_fq *r_A, const spg_fq *r_B, const spg_fq *r_C, fq *out) {
  fq rA, rB, rC;
  memcpy(&rA, r_A, 32);
  memcpy(&rB, r_B, 32);
  memcpy(&rC, r_C, 32);
  unsigned log_max = log2u(max_num_cols);
  SPG_CHECK(t0 + count <= num_segs * max_num_cols, "abc_slice: [%zu, %zu) exceeds the table", t0, t0 + count);
  for (int m = 0; m < 3; m++)
    if (inst->by_col[m].nnz) SPG_CHECK((inst->max_col[m] >> log_max) < num_segs, "abc_slice: column %u addresses segment >= %zu", inst->max_col[m], num_segs);
  SPG_LAUNCH(ctx, k_abc_table, grid_for(ctx, count, 128), 128, 0, view(inst->by_col[0]), view(inst->by_col[1]),
             view(inst->by_col[2]), evals_rx, t0, count, log_max, log_max, rA, rB, rC, out);
  return SPG_OK;
}

int eq_evals_device(spg_ctx *ctx, const fq *d_r, const spg_fq *h_r, size_t ell, fq *out, fq *scratch);

int witness_wait(spg_witness *w) {
  if (w->ready) {
    SPG_CUDA(cudaStreamWaitEvent(w->ctx->stream, w->ready, 0));
    SPG_CUDA(cudaEventDestroy(w->ready));  // released once the recorded work completes
    w->ready = nullptr;
  }
  return SPG_OK;
}

}  // namespace spg

using namespace spg;

extern "C" {

int spg_r1cs_create(spg_ctx *ctx, size_t num_instances, size_t max_num_cons, const size_t *num_cons,
                    size_t num_vars, const size_t *nnz, const uint32_t *rows, const uint32_t *cols,
                    const spg_fq *vals, spg_r1cs **out) {
  spg::DeviceGuard _dev(spg::ctx_of(ctx));
  SPG_CHECK(ctx && out && num_cons && nnz, "spg_r1cs_create: null argument");
  SPG_CHECK(num_instances >= 1, "spg_r1cs_create: need at least one instance");
  // R1CSInstance::new asserts (src/r1csinstance.rs:105-117)
  SPG_CHECK(is_pow2(max_num_cons), "spg_r1cs_create: max_num_cons %zu is not a power of two", max_num_cons);
  SPG_CHECK(is_pow2(num_vars), "spg_r1cs_create: num_vars %zu is not a power of two", num_vars);
  SPG_CHECK(num_vars < (1ull << 31) && max_num_cons < (1ull << 31), "spg_r1cs_create: dimension too large");
  for (size_t i = 0; i < num_instances; i++)
    SPG_CHECK(is_pow2(num_cons[i]) && num_cons[i] <= max_num_cons,
              "spg_r1cs_create: num_cons[%zu] = %zu is not a power of two <= %zu", i, num_cons[i], max_num_cons);
  spg_r1cs *r = new (std::nothrow) spg_r1cs();
  if (!r) return SPG_ENOMEM;
  r->ctx = ctx;
  r->num_instances = num_instances;
  r->max_num_cons = max_num_cons;
  r->num_vars = num_vars;
  r->num_cons.assign(num_cons, num_cons + num_instances);
  r->by_row.resize(3 * num_instances);
  r->by_col.resize(3 * num_instances);
  r->max_col.assign(3 * num_instances, 0);
  r->h_cols.resize(3 * num_instances);
  size_t base = 0;
  for (size_t m = 0; m < 3 * num_instances; m++) {
    size_t n = nnz[m];
    SPG_CHECK(n == 0 || (rows && cols && vals), "spg_r1cs_create: null matrix data");
    for (size_t e = 0; e < n; e++) {
      if (rows[base + e] >= r->num_cons[m / 3] || cols[base + e] >= num_vars) {
        set_error("spg_r1cs_create: entry %zu of matrix %zu (%u, %u) is out of range", e, m, rows[base + e], cols[base + e]);
        spg_r1cs_destroy(r);
        return SPG_EINVAL;
      }
      if (cols[base + e] > r->max_col[m]) r->max_col[m] = cols[base + e];
    }
    r->h_cols[m].assign(cols + base, cols + base + n);
    int rc = build_csx(ctx, r->num_cons[m / 3], n, rows + base, cols + base, vals + base, &r->by_row[m]);
    if (rc == SPG_OK) rc = build_csx(ctx, num_vars, n, cols + base, rows + base, vals + base, &r->by_col[m]);
    if (rc != SPG_OK) {
      spg_r1cs_destroy(r);
      return rc;
    }
    base += n;
  }
  *out = r;
  return SPG_OK;
}

void spg_r1cs_destroy(spg_r1cs *r) {
  spg::DeviceGuard _dev(spg::ctx_of(r));
  if (!r) return;
  for (auto &c : r->by_row) free_csx(c);
  for (auto &c : r->by_col) free_csx(c);
  delete r;
}

int spg_r1cs_multi_evaluate(spg_ctx *ctx, const spg_r1cs *inst, const spg_fq *rx, size_t nrx,
                            const spg_fq *ry, size_t nry, spg_fq *out) {
  spg::DeviceGuard _dev(spg::ctx_of(ctx));
  SPG_CHECK(ctx && inst && out, "spg_r1cs_multi_evaluate: null argument");
  SPG_CHECK(((size_t)1 << nrx) == inst->max_num_cons && ((size_t)1 << nry) == inst->num_vars,
            "spg_r1cs_multi_evaluate: |rx| = %zu, |ry| = %zu do not match %zu x %zu", nrx, nry,
            inst->max_num_cons, inst->num_vars);
  size_t nx = (size_t)1 << nrx, ny = (size_t)1 << nry;
  CudaTmp t_tabs, t_r;
  SPG_CUDA(t_tabs.alloc((nx + ny + (nx > ny ? nx : ny)) * sizeof(fq)));
  SPG_CUDA(t_r.alloc((nrx + nry + 1) * sizeof(fq)));
  fq *tabs = t_tabs.as<fq>(), *d_r = t_r.as<fq>();
  SPG_CUDA(cudaMemcpyAsync(d_r, rx, nrx * sizeof(fq), cudaMemcpyHostToDevice, ctx->stream));
  SPG_CUDA(cudaMemcpyAsync(d_r + nrx, ry, nry * sizeof(fq), cudaMemcpyHostToDevice, ctx->stream));
  fq *trx = tabs, *try_ = tabs + nx, *scratch = tabs + nx + ny;
  int rc = eq_evals_device(ctx, d_r, rx, nrx, trx, scratch);
  if (rc == SPG_OK) rc = eq_evals_device(ctx, d_r + nrx, ry, nry, try_, scratch);
  for (size_t m = 0; rc == SPG_OK && m < 3 * inst->num_instances; m++) {
    const Csx &c = inst->by_row[m];
    int grid = grid_for(ctx, c.nnz, 256, 2);
    rc = ensure_partials(ctx, grid);
    if (rc != SPG_OK) break;
    k_sparse_eval<<<grid, 256, 0, ctx->stream>>>(c.major, c.idx, c.val, c.nnz, trx, try_, ctx->d_partials);
    ctx->launches++;
    rc = reduce_partials(ctx, ctx->d_partials, grid, 1, ctx->d_result + (m % 48));
    if (rc == SPG_OK && (m % 48 == 47 || m + 1 == 3 * inst->num_instances)) {
      size_t cnt = m % 48 + 1;
      rc = fetch_result(ctx, (int)cnt, out + (m - (cnt - 1)));
    }
  }
  cudaStreamSynchronize(ctx->stream);  // before the guards free the tables
  return rc;
}

static int witness_upload_impl(spg_ctx *ctx, size_t num_instances, const size_t *num_proofs,
                               const size_t *num_inputs, const spg_fq *host_w_mat, bool async, spg_witness **out) {
  SPG_CHECK(ctx && out && num_proofs && num_inputs && host_w_mat, "spg_witness_upload: null argument");
  SPG_CHECK(num_instances >= 1, "spg_witness_upload: need at least one instance");
  spg_witness *w = new (std::nothrow) spg_witness();
  if (!w) return SPG_ENOMEM;
  w->ctx = ctx;
  w->num_instances = num_instances;
  w->num_proofs.assign(num_proofs, num_proofs + num_instances);
  w->num_inputs.assign(num_inputs, num_inputs + num_instances);
  size_t tot = 0;
  for (size_t p = 0; p < num_instances; p++) {
    if (!is_pow2(num_proofs[p]) || !is_pow2(num_inputs[p])) {
      set_error("spg_witness_upload: num_proofs[%zu] = %zu / num_inputs = %zu must be powers of two", p,
                num_proofs[p], num_inputs[p]);
      delete w;
      return SPG_EINVAL;
    }
    w->off.push_back(tot);
    tot += num_proofs[p] * num_inputs[p];
  }
  w->total = tot;
  cudaError_t e = dev_alloc(ctx, &w->d, tot * sizeof(fq));
  if (async) {
    // the buffer comes from the compute stream's pool: order the copy stream after the
    // allocation, copy there, and leave an event for the first consumer
    cudaEvent_t alloc_done = nullptr;
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&alloc_done, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventRecord(alloc_done, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx->copy_stream, alloc_done, 0);
    if (e == cudaSuccess) e = cudaMemcpyAsync(w->d, host_w_mat, tot * sizeof(fq), cudaMemcpyHostToDevice, ctx->copy_stream);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&w->ready, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventRecord(w->ready, ctx->copy_stream);
    if (alloc_done) cudaEventDestroy(alloc_done);
  } else {
    if (e == cudaSuccess) e = cudaMemcpyAsync(w->d, host_w_mat, tot * sizeof(fq), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  }
  if (e != cudaSuccess) {
    spg_witness_destroy(w);
    return cuda_fail(e, "witness upload", __FILE__, __LINE__);
  }
  w->views.assign(num_instances, nullptr);
  *out = w;
  return SPG_OK;
}

int spg_witness_upload(spg_ctx *ctx, size_t num_instances, const size_t *num_proofs,
                       const size_t *num_inputs, const spg_fq *host_w_mat, spg_witness **out) {
  spg::DeviceGuard _dev(spg::ctx_of(ctx));
  return witness_upload_impl(ctx, num_instances, num_proofs, num_inputs, host_w_mat, false, out);
}

int spg_witness_upload_async(spg_ctx *ctx, size_t num_instances, const size_t *num_proofs,
                             const size_t *num_inputs, const spg_fq *host_w_mat, spg_witness **out) {
  spg::DeviceGuard _dev(spg::ctx_of(ctx));
  return witness_upload_impl(ctx, num_instances, num_proofs, num_inputs, host_w_mat, true, out);
}

void spg_witness_destroy(spg_witness *w) {
  spg::DeviceGuard _dev(spg::ctx_of(w));
  if (!w) return;
  if (w->ready) {
    // a section freed before anyone consumed it: the free below is stream-ordered on the
    // compute stream, so the copy must be ordered before it
    if (spg::ctx_alive(w->ctx)) cudaStreamWaitEvent(w->ctx->stream, w->ready, 0);
    cudaEventDestroy(w->ready);
  }
  for (spg_vec *v : w->views)
    if (v) spg_vec_free(v);
  if (w->d) dev_free(w->ctx, w->d);
  delete w;
}

int spg_witness_poly(spg_witness *w, size_t p, spg_vec **out) {
  spg::DeviceGuard _dev(spg::ctx_of(w));
  SPG_CHECK(w && out, "spg_witness_poly: null argument");
  SPG_CHECK(p < w->num_instances, "spg_witness_poly: instance %zu out of range", p);
  SPG_TRY(witness_wait(w));
  if (!w->views[p]) {
    spg_vec *v = nullptr;
    SPG_TRY(spg_vec_wrap(w->ctx, w->d + w->off[p], w->num_proofs[p] * w->num_inputs[p], &v));
    w->views[p] = v;
  }
  *out = w->views[p];
  return SPG_OK;
}

int spg_zmat_build(spg_ctx *ctx, size_t num_instances, const size_t *num_proofs, const size_t *num_inputs,
                   size_t num_witness_secs, spg_witness *const *witness_secs, spg_zmat **out) {
  spg::DeviceGuard _dev(spg::ctx_of(ctx));
  SPG_CHECK(ctx && out && num_proofs && num_inputs && witness_secs, "spg_zmat_build: null argument");
  // asserts of R1CSProof::prove (src/r1csproof.rs:240-263)
  SPG_CHECK(num_witness_secs >= 1 && num_witness_secs <= 16, "spg_zmat_build: num_witness_secs must be in 1..=16");
  for (size_t p = 0; p < num_instances; p++)
    SPG_CHECK(is_pow2(num_proofs[p]) && is_pow2(num_inputs[p]), "spg_zmat_build: sizes must be powers of two");
  for (size_t w = 0; w < num_witness_secs; w++) {
    const spg_witness *ws = witness_secs[w];
    SPG_CHECK(ws, "spg_zmat_build: witness section %zu is null", w);
    SPG_CHECK(ws->num_instances == 1 || ws->num_instances == num_instances,
              "spg_zmat_build: section %zu has %zu instances, expected 1 or %zu", w, ws->num_instances, num_instances);
    for (size_t p = 0; p < ws->num_instances; p++)
      SPG_CHECK(ws->num_proofs[p] == 1 || ws->num_proofs[p] == num_proofs[p],
                "spg_zmat_build: section %zu instance %zu has %zu proofs, expected 1 or %zu", w, p,
                ws->num_proofs[p], num_proofs[p]);
  }
  for (size_t w = 0; w < num_witness_secs; w++) SPG_TRY(witness_wait(witness_secs[w]));
  spg_zmat *z = new (std::nothrow) spg_zmat();
  if (!z) return SPG_ENOMEM;
  z->ctx = ctx;
  z->P = num_instances;
  z->W = num_witness_secs;
  z->num_proofs.assign(num_proofs, num_proofs + num_instances);
  z->num_inputs.assign(num_inputs, num_inputs + num_instances);
  std::vector<SecView> hv(num_instances * num_witness_secs);
  for (size_t p = 0; p < num_instances; p++)
    for (size_t w = 0; w < num_witness_secs; w++) {
      const spg_witness *ws = witness_secs[w];
      size_t pw = ws->num_instances == 1 ? 0 : p;
      size_t ni = ws->num_inputs[pw];
      SecView &v = hv[p * num_witness_secs + w];
      v.ptr = ws->d + ws->off[pw];
      v.q_stride = ws->num_proofs[pw] == 1 ? 0 : ni;
      v.copy = ni < num_inputs[p] ? ni : num_inputs[p];
    }
  cudaError_t e = dev_alloc(ctx, &z->views, hv.size() * sizeof(SecView));
  // (pageable source: staged before the call returns, so hv may go out of scope)
  if (e == cudaSuccess) e = cudaMemcpyAsync(z->views, hv.data(), hv.size() * sizeof(SecView), cudaMemcpyHostToDevice, ctx->stream);
  if (e != cudaSuccess) {
    spg_zmat_destroy(z);
    return cuda_fail(e, "z_mat views", __FILE__, __LINE__);
  }
  *out = z;
  return SPG_OK;
}

void spg_zmat_destroy(spg_zmat *z) {
  spg::DeviceGuard _dev(spg::ctx_of(z));
  if (!z) return;
  if (z->views) dev_free(z->ctx, z->views);
  delete z;
}

}  // extern "C"
