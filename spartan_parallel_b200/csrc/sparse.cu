// Dense representation of a batch of sparse matrix polynomials, resident on the device,
// and the table builders of the memory-check network around it.
//   SparseMatPolynomial::multi_sparse_to_dense_rep   src/sparse_mlpoly.rs:368-425
//   AddrTimestamps::new / deref                       src/sparse_mlpoly.rs:212-271
//   Derefs::new (merge of row/col lookups)            src/sparse_mlpoly.rs:34-62
//   Layers::build_hash_layer                          src/sparse_mlpoly.rs:612-687
// Layout: comb_ops is ONE buffer [row_addr[b] | row_read_ts[b] | col_addr[b] | col_read_ts[b] |
// val[b] | zero pad], each block N scalars, which is exactly DensePolynomial::merge's order
// (:409-416), so the polynomial that is committed and the per-poly tables the sumchecks
// read are the same memory. comb_mem = [row_audit_ts | col_audit_ts].
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"

namespace spg {

// ---------------------------------------------------------------- AddrTimestamps::new on the device
// The reference walks the operations in order with one counter per memory cell
// (src/sparse_mlpoly.rs:222-253): read_ts[k] = number of EARLIER operations on the same address,
// audit_ts[a] = number of operations on a. Order-dependent but not sequential: a STABLE sort of the
// operation indices by address (CUB radix sort over the address bits; the one library primitive in
// this backend, preprocessing only) puts each address's operations in a run, in index order, and
// the timestamp is the position inside the run.
__global__ void k_iota_u32(uint32_t *__restrict__ v, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) v[i] = (uint32_t)i;
}
// keys: sorted addresses, idx: the operations in that order. run start by binary search (the keys are
// L2 resident); the last element of a run writes the cell's final counter.
__global__ void k_run_ranks(const uint32_t *__restrict__ keys, const uint32_t *__restrict__ idx, size_t n,
                            uint32_t *__restrict__ read_ts, uint32_t *__restrict__ audit) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    uint32_t key = keys[i];
    size_t lo = 0, hi = i;  // first position holding `key`
    while (lo < hi) {
      size_t mid = (lo + hi) >> 1;
      if (keys[mid] < key) lo = mid + 1;
      else hi = mid;
    }
    uint32_t rank = (uint32_t)(i - lo);
    read_ts[idx[i]] = rank;
    if (i + 1 == n || keys[i + 1] != key) audit[key] = rank + 1;
  }
}

// read_ts[n] and audit[cells] (zeroed here) from addr[n] < cells; all device pointers
int addr_timestamps(spg_ctx *ctx, const uint32_t *d_addr, size_t n, size_t cells, uint32_t *d_read_ts, uint32_t *d_audit) {
  uint32_t *buf = nullptr;  // keys_out | idx_in | idx_out
  void *tmp = nullptr;
  size_t tmp_bytes = 0;
  int end_bit = (int)log2u(cells);
  if (end_bit < 1) end_bit = 1;
  SPG_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_addr, (uint32_t *)nullptr, (const uint32_t *)nullptr,
                                           (uint32_t *)nullptr, (int)n, 0, end_bit, ctx->stream));
  SPG_CUDA(dev_alloc(ctx, &buf, 3 * n * sizeof(uint32_t)));
  cudaError_t e = dev_alloc_bytes(ctx, &tmp, tmp_bytes ? tmp_bytes : 16);
  if (e != cudaSuccess) {
    dev_free(ctx, buf);
    return cuda_fail(e, "sort scratch", __FILE__, __LINE__);
  }
  uint32_t *keys_out = buf, *idx_in = buf + n, *idx_out = buf + 2 * n;
  int rc = [&]() -> int {
    SPG_LAUNCH(ctx, k_iota_u32, grid_for(ctx, n, 256), 256, 0, idx_in, n);
    SPG_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, d_addr, keys_out, idx_in, idx_out, (int)n, 0, end_bit, ctx->stream));
    ctx->launches += 4;  // CUB's histogram + onesweep passes (a lower bound; not launched through SPG_LAUNCH)
    SPG_CUDA(cudaMemsetAsync(d_audit, 0, cells * sizeof(uint32_t), ctx->stream));
    SPG_LAUNCH(ctx, k_run_ranks, grid_for(ctx, n, 256), 256, 0, keys_out, idx_out, n, d_read_ts, d_audit);
    return SPG_OK;
  }();
  dev_free(ctx, tmp);
  dev_free(ctx, buf);
  return rc;
}

__device__ __forceinline__ fq fq_from_u32_dev(unsigned int v) {
  // Scalar::from(u64) = [v,0,0,0] * R^2 (ristretto255.rs:212-216)
  fq R2, raw = fq_zero();
  R2.v[0] = 0x449c0f01u; R2.v[1] = 0xa40611e3u; R2.v[2] = 0x68859347u; R2.v[3] = 0xd00e1ba7u;
  R2.v[4] = 0x17f5be65u; R2.v[5] = 0xceec73d2u; R2.v[6] = 0x7c309a3du; R2.v[7] = 0x0399411bu;
  raw.v[0] = v;
  return fq_mul(R2, raw);
}

__global__ void k_from_u32(const uint32_t *__restrict__ in, size_t n, fq *__restrict__ out) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    fq_store(out + i, fq_from_u32_dev(in[i]));
}

// out[s*bN + k] = mem_s[addr_s[k]] for the two sides s (row: mem_rx, col: mem_ry); zero pad after 2bN
__global__ void k_sparse_deref(const uint32_t *__restrict__ row, const uint32_t *__restrict__ col, size_t bN,
                               const fq *__restrict__ mem_rx, const fq *__restrict__ mem_ry, size_t total,
                               fq *__restrict__ out) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    fq v = fq_zero();
    if (i < bN) v = fq_load(mem_rx + row[i]);
    else if (i < 2 * bN) v = fq_load(mem_ry + col[i - bN]);
    fq_store(out + i, v);
  }
}

// hash_func(addr, val, ts) - tau = ts*gamma^2 + val*gamma + addr - tau  (:623-626) over scalar tables;
// addr == nullptr means addr[i] = i (init / audit), ts == nullptr means ts = 0
__global__ void k_hash_layer_fq(const fq *__restrict__ addr, const fq *__restrict__ val, const fq *__restrict__ ts,
                                size_t n, fq gamma, fq gamma2, fq tau, int ts_plus_one, fq *__restrict__ out) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    fq a = addr ? fq_load(addr + i) : fq_from_u32_dev((unsigned int)i);
    fq h = fq_add(fq_mul(fq_load(val + i), gamma), a);
    if (ts) {
      fq t = fq_load(ts + i);
      if (ts_plus_one) t = fq_add(t, fq_one());
      h = fq_add(h, fq_mul(t, gamma2));
    } else if (ts_plus_one) {
      h = fq_add(h, gamma2);
    }
    fq_store(out + i, fq_sub(h, tau));
  }
}

}  // namespace spg

using namespace spg;

struct spg_sparse {
  spg_ctx *ctx = nullptr;
  size_t batch = 0, N = 0, M = 0;
  uint32_t *d_row = nullptr, *d_col = nullptr;  // [batch][N]
  fq *comb_ops = nullptr;
  size_t comb_ops_len = 0;
  fq *comb_mem = nullptr;  // [2M]
};

extern "C" {

int spg_sparse_create(spg_ctx *ctx, size_t batch, size_t num_vars_x, size_t num_vars_y, const size_t *nnz,
                      const uint32_t *rows, const uint32_t *cols, const spg_fq *vals, spg_sparse **out) {
  spg::DeviceGuard _dev(spg::ctx_of(ctx));
  SPG_CHECK(ctx && nnz && rows && cols && vals && out, "spg_sparse_create: null argument");
  SPG_CHECK(batch >= 1, "spg_sparse_create: empty batch");
  SPG_CHECK(num_vars_x < 32 && num_vars_y < 32, "spg_sparse_create: at most 2^31 rows/columns");
  size_t N = 1, total = 0;
  for (size_t i = 0; i < batch; i++) {
    if (next_pow2(nnz[i]) > N) N = next_pow2(nnz[i]);
    total += nnz[i];
  }
  size_t M = (size_t)1 << (num_vars_x > num_vars_y ? num_vars_x : num_vars_y);
  SPG_CHECK(batch * N < ((size_t)1 << 32), "spg_sparse_create: batch * N = %zu exceeds 2^32 timestamps", batch * N);
  // integer side: the host only pads the address vectors; the read / audit timestamps
  // (AddrTimestamps::new, :222-253: one counter per memory cell running across the batch) are
  // computed on the device from them (addr_timestamps above)
  size_t bN = batch * N;
  SPG_CHECK(bN < ((size_t)1 << 31), "spg_sparse_create: batch * N = %zu operations exceed 2^31", bN);
  std::vector<uint32_t> h(2 * bN, 0);
  uint32_t *h_row = h.data(), *h_col = h_row + bN;
  size_t pos = 0;
  for (size_t i = 0; i < batch; i++) {
    for (size_t k = 0; k < nnz[i]; k++, pos++) {
      SPG_CHECK(rows[pos] < ((size_t)1 << num_vars_x) && cols[pos] < ((size_t)1 << num_vars_y),
                "spg_sparse_create: entry %zu of matrix %zu is out of range", k, i);
      h_row[i * N + k] = rows[pos];
      h_col[i * N + k] = cols[pos];
    }
  }
  spg_sparse *s = new (std::nothrow) spg_sparse();
  if (!s) return SPG_ENOMEM;
  s->ctx = ctx;
  s->batch = batch;
  s->N = N;
  s->M = M;
  s->comb_ops_len = next_pow2(5 * bN);
  uint32_t *d_int = nullptr, *d_audit = nullptr;
  cudaError_t e = cudaMalloc(&d_int, 4 * bN * 4);
  if (e == cudaSuccess) e = cudaMalloc(&d_audit, 2 * M * 4);
  if (e == cudaSuccess) e = cudaMalloc(&s->d_row, bN * 4);
  if (e == cudaSuccess) e = cudaMalloc(&s->d_col, bN * 4);
  if (e == cudaSuccess) e = cudaMalloc(&s->comb_ops, s->comb_ops_len * sizeof(fq));
  if (e == cudaSuccess) e = cudaMalloc(&s->comb_mem, 2 * M * sizeof(fq));
  if (e != cudaSuccess) {
    cudaFree(d_int);
    cudaFree(d_audit);
    spg_sparse_destroy(s);
    return cuda_fail(e, "cudaMalloc(sparse)", __FILE__, __LINE__);
  }
  int rc = [&]() -> int {
    // d_int = [row_addr | row_read_ts | col_addr | col_read_ts], d_audit = [row cells | col cells]
    SPG_CUDA(cudaMemcpyAsync(d_int, h_row, bN * 4, cudaMemcpyHostToDevice, ctx->stream));
    SPG_CUDA(cudaMemcpyAsync(d_int + 2 * bN, h_col, bN * 4, cudaMemcpyHostToDevice, ctx->stream));
    SPG_TRY(addr_timestamps(ctx, d_int, bN, M, d_int + bN, d_audit));
    SPG_TRY(addr_timestamps(ctx, d_int + 2 * bN, bN, M, d_int + 3 * bN, d_audit + M));
    SPG_CUDA(cudaMemcpyAsync(s->d_row, d_int, bN * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    SPG_CUDA(cudaMemcpyAsync(s->d_col, d_int + 2 * bN, bN * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    SPG_LAUNCH(ctx, k_from_u32, grid_for(ctx, 4 * bN, 256), 256, 0, d_int, 4 * bN, s->comb_ops);
    SPG_LAUNCH(ctx, k_from_u32, grid_for(ctx, 2 * M, 256), 256, 0, d_audit, 2 * M, s->comb_mem);
    // val[b][N] (zero padded), then the zero tail of the merged polynomial
    SPG_CUDA(cudaMemsetAsync(s->comb_ops + 4 * bN, 0, (s->comb_ops_len - 4 * bN) * sizeof(fq), ctx->stream));
    pos = 0;
    for (size_t i = 0; i < batch; i++) {
      if (nnz[i])
        SPG_CUDA(cudaMemcpyAsync(s->comb_ops + 4 * bN + i * N, vals + pos, nnz[i] * sizeof(fq), cudaMemcpyHostToDevice,
                                 ctx->stream));
      pos += nnz[i];
    }
    SPG_CUDA(cudaStreamSynchronize(ctx->stream));
    return SPG_OK;
  }();
  cudaFree(d_int);
  cudaFree(d_audit);
  if (rc != SPG_OK) {
    spg_sparse_destroy(s);
    return rc;
  }
  *out = s;
  return SPG_OK;
}

void spg_sparse_destroy(spg_sparse *s) {
  spg::DeviceGuard _dev(spg::ctx_of(s));
  if (!s) return;
  cudaFree(s->d_row);
  cudaFree(s->d_col);
  cudaFree(s->comb_ops);
  cudaFree(s->comb_mem);
  delete s;
}

size_t spg_sparse_num_ops(const spg_sparse *s) { return s ? s->N : 0; }
size_t spg_sparse_num_mem_cells(const spg_sparse *s) { return s ? s->M : 0; }

int spg_sparse_view(spg_sparse *s, int kind, size_t i, spg_vec **out) {
  spg::DeviceGuard _dev(spg::ctx_of(s));
  SPG_CHECK(s && out, "spg_sparse_view: null argument");
  size_t bN = s->batch * s->N;
  fq *p = nullptr;
  size_t n = s->N;
  switch (kind) {
    case SPG_SPARSE_ROW_ADDR: p = s->comb_ops + i * s->N; break;
    case SPG_SPARSE_ROW_READ_TS: p = s->comb_ops + bN + i * s->N; break;
    case SPG_SPARSE_COL_ADDR: p = s->comb_ops + 2 * bN + i * s->N; break;
    case SPG_SPARSE_COL_READ_TS: p = s->comb_ops + 3 * bN + i * s->N; break;
    case SPG_SPARSE_VAL: p = s->comb_ops + 4 * bN + i * s->N; break;
    case SPG_SPARSE_ROW_AUDIT_TS: p = s->comb_mem; n = s->M; i = 0; break;
    case SPG_SPARSE_COL_AUDIT_TS: p = s->comb_mem + s->M; n = s->M; i = 0; break;
    case SPG_SPARSE_COMB_OPS: p = s->comb_ops; n = s->comb_ops_len; i = 0; break;
    case SPG_SPARSE_COMB_MEM: p = s->comb_mem; n = 2 * s->M; i = 0; break;
    default: SPG_CHECK(false, "spg_sparse_view: unknown kind %d", kind);
  }
  SPG_CHECK(i < s->batch, "spg_sparse_view: index %zu out of range (batch %zu)", i, s->batch);
  return spg_vec_wrap(s->ctx, p, n, out);
}

int spg_sparse_deref(spg_ctx *ctx, const spg_sparse *s, const spg_vec *mem_rx, const spg_vec *mem_ry, spg_vec **out) {
  spg::DeviceGuard _dev(spg::ctx_of(ctx));
  SPG_CHECK(ctx && s && mem_rx && mem_ry && out, "spg_sparse_deref: null argument");
  SPG_CHECK(mem_rx->n == s->M && mem_ry->n == s->M, "spg_sparse_deref: memories must have %zu cells (got %zu, %zu)", s->M,
            mem_rx->n, mem_ry->n);
  size_t bN = s->batch * s->N, total = next_pow2(2 * bN);
  spg_vec *o = nullptr;
  SPG_TRY(vec_new(ctx, total, &o));
  ctx->next_units = 72.0 * (double)(2 * bN);
  SPG_LAUNCH(ctx, k_sparse_deref, grid_for(ctx, total, 256), 256, 0, s->d_row, s->d_col, bN, mem_rx->d, mem_ry->d, total,
             o->d);
  *out = o;
  return SPG_OK;
}

int spg_hash_layer_fq(spg_ctx *ctx, const spg_vec *addr, const spg_vec *val, const spg_vec *ts, int ts_plus_one,
                      const spg_fq *gamma, const spg_fq *tau, spg_vec **out) {
  spg::DeviceGuard _dev(spg::ctx_of(ctx));
  SPG_CHECK(ctx && val && gamma && tau && out, "spg_hash_layer_fq: null argument");
  size_t n = val->n;
  SPG_CHECK((!addr || addr->n == n) && (!ts || ts->n == n), "spg_hash_layer_fq: tables must share the length %zu", n);
  SPG_CHECK(n < ((size_t)1 << 32), "spg_hash_layer_fq: at most 2^32 entries");
  spg_vec *o = nullptr;
  SPG_TRY(vec_new(ctx, n, &o));
  hfq g = hfq_from(*gamma);
  hfq g2 = hfq_mul(g, g);
  fq fg, fg2, ft;
  memcpy(&fg, &g, 32);
  memcpy(&fg2, &g2, 32);
  memcpy(&ft, tau, 32);
  ctx->next_units = (double)n * 32.0 * (2 + (addr ? 1 : 0) + (ts ? 1 : 0));
  SPG_LAUNCH(ctx, k_hash_layer_fq, grid_for(ctx, n, 256), 256, 0, addr ? addr->d : nullptr, val->d, ts ? ts->d : nullptr, n,
             fg, fg2, ft, ts_plus_one, o->d);
  *out = o;
  return SPG_OK;
}

int spg_vec_clone(spg_ctx *ctx, const spg_vec *v, size_t offset, size_t n, spg_vec **out) {
  spg::DeviceGuard _dev(spg::ctx_of(ctx));
  SPG_CHECK(ctx && v && out, "spg_vec_clone: null argument");
  SPG_CHECK(offset + n <= v->n, "spg_vec_clone: range [%zu, %zu) exceeds length %zu", offset, offset + n, v->n);
  spg_vec *o = nullptr;
  SPG_TRY(vec_new(ctx, n, &o));
  if (n) SPG_CUDA(cudaMemcpyAsync(o->d, v->d + offset, n * sizeof(fq), cudaMemcpyDeviceToDevice, ctx->stream));
  *out = o;
  return SPG_OK;
}

}  // extern "C"
