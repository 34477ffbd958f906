// Permutation-product recurrences of the w3 witness sections built inside SNARK::prove.
//   reference: src/lib.rs:1378-1400 (perm_exec_w3), 862-880 (mem_gen), 1533-1570 and the
//   PHY / VIR pairs after it (block_w3). All of them are, per proving instance, from the last
//   proof q down to the first:
//       D[q]  = x[q] * (pi[q+1] + 1 - v[q+1])      (D[last] = x[last])
//       pi[q] = v[q] * D[q]
//   The reference walks q sequentially (`for q in (0..n).rev()`), one core, right before the
//   sections are committed.
//
// On the device the recurrence is a suffix scan of affine maps: with pi[q+1] = v[q+1] D[q+1],
//   D[q] = A_q * D[q+1] + B_q,   A_q = x[q] * v[q+1],   B_q = x[q] - A_q,
// and composition (A, B) o (A', B') = (A A', A B' + B) is associative. The last entry of a
// segment has A = 0, which also cuts the scan between instances. Field arithmetic is exact, so
// the scanned values are bit-identical to the sequential ones.
//
// Three launches: per-block aggregates, one block that scans the aggregates, apply.
#include "common.cuh"

namespace spg {

constexpr int PS_THREADS = 256;
constexpr int PS_ITEMS = 4;
constexpr int PS_CHUNK = PS_THREADS * PS_ITEMS;

struct Strided {
  const fq *base;
  unsigned long long off, stride;
  __device__ __forceinline__ const fq *at(unsigned long long q) const { return base + off + q * stride; }
};
struct StridedOut {
  fq *base;
  unsigned long long off, stride;
  __device__ __forceinline__ fq *at(unsigned long long q) const { return base + off + q * stride; }
};

struct Affine {
  fq A, B;
};
// (hi o lo)(D) = hi.A * (lo.A * D + lo.B) + hi.B : lo is applied first
__device__ __forceinline__ Affine affine_after(const Affine &hi, const Affine &lo) {
  Affine r;
  r.A = fq_mul(hi.A, lo.A);
  r.B = fq_add(fq_mul(hi.A, lo.B), hi.B);
  return r;
}

// scan position j <-> proof q = n - 1 - j (the scan runs from the last proof to the first)
__device__ __forceinline__ Affine perm_map(const Strided &v, const Strided &x, const unsigned char *__restrict__ seg_last,
                                           unsigned long long q, fq &xq) {
  xq = fq_load(x.at(q));
  Affine m;
  if (seg_last[q]) {
    m.A = fq_zero();
    m.B = xq;
  } else {
    m.A = fq_mul(xq, fq_load(v.at(q + 1)));
    m.B = fq_sub(xq, m.A);
  }
  return m;
}

// inclusive scan of one Affine per thread across the block (Hillis-Steele over shared memory);
// returns the composition of the maps of threads 0..threadIdx.x
__device__ __forceinline__ Affine block_scan_affine(Affine mine, Affine *sh /* [2][PS_THREADS] */) {
  int t = threadIdx.x, cur = 0;
  sh[t] = mine;
  __syncthreads();
  for (int d = 1; d < PS_THREADS; d <<= 1) {
    Affine r = sh[cur * PS_THREADS + t];
    if (t >= d) r = affine_after(r, sh[cur * PS_THREADS + t - d]);
    sh[(cur ^ 1) * PS_THREADS + t] = r;
    __syncthreads();
    cur ^= 1;
  }
  return sh[cur * PS_THREADS + t];
}

// pass 1: aggregate map of every chunk of PS_CHUNK scan positions
__global__ void __launch_bounds__(PS_THREADS)
k_perm_aggregate(Strided v, Strided x, const unsigned char *__restrict__ seg_last, unsigned long long n,
                 Affine *__restrict__ agg) {
  extern __shared__ __align__(32) unsigned char smem_raw[];
  Affine *sh = reinterpret_cast<Affine *>(smem_raw);
  unsigned long long j0 = (unsigned long long)blockIdx.x * PS_CHUNK + (unsigned long long)threadIdx.x * PS_ITEMS;
  Affine run;
  run.A = fq_one();
  run.B = fq_zero();
#pragma unroll 1
  for (int k = 0; k < PS_ITEMS; k++) {
    unsigned long long j = j0 + k;
    if (j >= n) break;
    fq xq;
    run = affine_after(perm_map(v, x, seg_last, n - 1 - j, xq), run);
  }
  Affine tot = block_scan_affine(run, sh);
  if (threadIdx.x == PS_THREADS - 1) agg[blockIdx.x] = tot;
}

// pass 2: one block turns the chunk aggregates into the value entering every chunk:
// enter[b] = D at the scan position just before chunk b (0 for chunk 0, where it is unused:
// the first scan position is the end of a segment, so its A is zero)
__global__ void __launch_bounds__(PS_THREADS)
k_perm_chunks(const Affine *__restrict__ agg, unsigned long long nchunks, fq *__restrict__ enter) {
  extern __shared__ __align__(32) unsigned char smem_raw[];
  Affine *sh = reinterpret_cast<Affine *>(smem_raw);
  __shared__ fq carry;
  if (threadIdx.x == 0) carry = fq_zero();
  __syncthreads();
  for (unsigned long long base = 0; base < nchunks; base += PS_THREADS) {
    unsigned long long b = base + threadIdx.x;
    Affine mine;
    if (b < nchunks) mine = agg[b];
    else {
      mine.A = fq_one();
      mine.B = fq_zero();
    }
    Affine inc = block_scan_affine(mine, sh);
    fq c = carry;
    fq out_after = fq_add(fq_mul(inc.A, c), inc.B);  // D after chunk b
    __syncthreads();
    // exclusive: the value entering chunk b is the value after chunk b-1
    sh[threadIdx.x].B = out_after;
    __syncthreads();
    if (b < nchunks) enter[b] = threadIdx.x == 0 ? c : sh[threadIdx.x - 1].B;
    if (threadIdx.x == PS_THREADS - 1) carry = out_after;
    __syncthreads();
  }
}

// pass 3: every chunk replays its maps from the entering value and writes D and pi
__global__ void __launch_bounds__(PS_THREADS)
k_perm_apply(Strided v, Strided x, const unsigned char *__restrict__ seg_last, unsigned long long n,
             const fq *__restrict__ enter, StridedOut D, StridedOut pi) {
  extern __shared__ __align__(32) unsigned char smem_raw[];
  Affine *sh = reinterpret_cast<Affine *>(smem_raw);
  unsigned long long j0 = (unsigned long long)blockIdx.x * PS_CHUNK + (unsigned long long)threadIdx.x * PS_ITEMS;
  Affine run;
  run.A = fq_one();
  run.B = fq_zero();
#pragma unroll 1
  for (int k = 0; k < PS_ITEMS; k++) {
    unsigned long long j = j0 + k;
    if (j >= n) break;
    fq xq;
    run = affine_after(perm_map(v, x, seg_last, n - 1 - j, xq), run);
  }
  Affine inc = block_scan_affine(run, sh);
  __syncthreads();
  sh[threadIdx.x] = inc;
  __syncthreads();
  fq d = enter[blockIdx.x];
  if (threadIdx.x > 0) {
    Affine prev = sh[threadIdx.x - 1];
    d = fq_add(fq_mul(prev.A, d), prev.B);
  }
#pragma unroll 1
  for (int k = 0; k < PS_ITEMS; k++) {
    unsigned long long j = j0 + k;
    if (j >= n) break;
    unsigned long long q = n - 1 - j;
    fq xq;
    Affine m = perm_map(v, x, seg_last, q, xq);
    d = fq_add(fq_mul(m.A, d), m.B);
    fq_store(D.at(q), d);
    fq_store(pi.at(q), fq_mul(fq_load(v.at(q)), d));
  }
}


// ---------------------------------------------------------------- derived witness sections
// Everything else SNARK::prove computes from the primary sections before committing
// (src/lib.rs:1328-1400, 1481-1613, mem_gen :832-880) is per row of the batch; only the (pi, D)
// columns couple rows, and those are the scan above. One warp per row; the sums over the row's
// entries are warp reductions (field addition is exact, so the order is free).

// out = (tau, r, r^2, ..., r^(used-1), 0, ...): thread i raises r to the i-th power
__global__ void k_perm_w0(fq tau, fq r, size_t used, size_t total, fq *__restrict__ out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  fq v = fq_zero();
  if (i == 0 && used > 0) v = tau;
  else if (i < used) {
    fq acc = fq_one(), base = r;
    for (size_t e = i; e; e >>= 1) {
      if (e & 1) acc = fq_mul(acc, base);
      base = fq_mul(base, base);
    }
    v = acc;
  }
  fq_store(out + i, v);
}

struct WitArgs {
  const fq *vars;  // rows x vars_width
  const fq *w0;    // perm_w0
  fq *w2, *w3;     // rows x w2_width, rows x 8
  unsigned long long rows, vars_width, w2_width;
  unsigned int n, io_width, phy_ops, vir_ops;
  int exec_mode;
  fq tau, r, r2, r3;
};

__global__ void __launch_bounds__(128)
k_wit_rows(const __grid_constant__ WitArgs a) {
  const unsigned long long q = (unsigned long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (q >= a.rows) return;
  const fq *__restrict__ in = a.vars + q * a.vars_width;
  fq *__restrict__ w2 = a.w2 + q * a.w2_width;
  fq *__restrict__ w3 = a.w3 + q * 8;
  const unsigned int n = a.n;
  // zero what the loops below do not write: entries 0..2 (lane 0 later), [2n, w2_width)
  for (unsigned long long i = 2ull * n + lane; i < a.w2_width; i += 32) fq_store(w2 + i, fq_zero());
  // w2[2 + i] = perm_w0[i] * in[i + 2], i in 1 .. 2(n-1)-1, and their sum (= sum of w2[3..], :1385 / :1535)
  fq s3 = fq_zero();
  for (unsigned int i = 1 + lane; i < 2 * (n - 1); i += 32) {
    fq t = fq_mul(fq_load(a.w0 + i), fq_load(in + i + 2));
    fq_store(w2 + 2 + i, t);
    s3 = fq_add(s3, t);
  }
  // the two weighted sums over the inputs and the outputs (:1362-1366 / :1522-1526)
  fq acc0 = fq_zero(), acc2 = fq_zero();
  for (unsigned int i = lane; i + 1 < n; i += 32) {
    fq perm = i == 0 ? fq_one() : fq_load(a.w0 + i);
    acc0 = fq_add(acc0, fq_mul(perm, fq_load(in + 2 + i)));
    acc2 = fq_add(acc2, fq_mul(perm, fq_load(in + 2 + (n - 1) + i)));
  }
  s3 = fq_warp_sum(s3);
  acc0 = fq_warp_sum(acc0);
  acc2 = fq_warp_sum(acc2);
  if (lane != 0) return;
  const fq v = fq_load(in);
  fq w20 = fq_mul(fq_add(v, acc0), v);
  fq w21 = fq_mul(fq_add(v, acc2), v);
  fq_store(w2, w20);
  fq_store(w2 + 1, w21);
  fq_store(w2 + 2, acc2);
  // (entries 3 .. 2n-1 were written by the loop above, [2n, w2_width) zeroed or filled below)
  fq_store(w3, v);
#pragma unroll 1
  for (int k = 2; k < 8; k++) fq_store(w3 + k, fq_zero());  // (pi, D) columns: filled by the scans
  fq_store(w3 + 1, fq_mul(v, fq_sub(fq_sub(a.tau, s3), fq_load(in + 2))));
  if (a.exec_mode) {  // consistency-check columns of perm_exec_w3 (:1389-1390)
    fq_store(w3 + 4, w20);
    fq_store(w3 + 5, w21);
    return;
  }
  // PHY: PMR = r * PD, PMC = (cnst or PMC[i-1]) * (tau - PA - PMR)  (:1541-1553)
  fq chain = v;
  unsigned long long src = a.io_width, dst = 2ull * n;
  for (unsigned int i = 0; i < a.phy_ops; i++, src += 2, dst += 2) {
    fq pmr = fq_mul(a.r, fq_load(in + src + 1));
    chain = fq_mul(chain, fq_sub(fq_sub(a.tau, fq_load(in + src)), pmr));
    fq_store(w2 + dst, pmr);
    fq_store(w2 + dst + 1, chain);
  }
  // VIR: VMR1..3 = r, r^2, r^3 times (VD, VL, VT), VMC chain  (:1570-1597)
  chain = v;
  for (unsigned int i = 0; i < a.vir_ops; i++, src += 4, dst += 4) {
    fq m1 = fq_mul(a.r, fq_load(in + src + 1)), m2 = fq_mul(a.r2, fq_load(in + src + 2)), m3 = fq_mul(a.r3, fq_load(in + src + 3));
    chain = fq_mul(chain, fq_sub(fq_sub(fq_sub(fq_sub(a.tau, fq_load(in + src)), m1), m2), m3));
    fq_store(w2 + dst, m1);
    fq_store(w2 + dst + 1, m2);
    fq_store(w2 + dst + 2, m3);
    fq_store(w2 + dst + 3, chain);
  }
}

// mem_gen rows (src/lib.rs:832-880): w2 = (0, 0, 0, r * data, 0 ...), w3 = (v, x, _, _, I, O, 0, 0)
__global__ void k_wit_mem(const fq *__restrict__ mems, size_t rows, size_t in_width, fq tau, fq r, size_t mem_width,
                          fq *__restrict__ w2, fq *__restrict__ w3) {
  size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= rows) return;
  const fq *m = mems + q * in_width;
  fq v = fq_load(m), addr = fq_load(m + 2), rd = fq_mul(r, fq_load(m + 3));
  for (size_t i = 0; i < mem_width; i++) fq_store(w2 + q * mem_width + i, i == 3 ? rd : fq_zero());
  fq *b = w3 + q * 8;
  fq_store(b, v);
  fq_store(b + 1, fq_mul(v, fq_sub(fq_sub(tau, addr), rd)));
  fq_store(b + 2, fq_zero());
  fq_store(b + 3, fq_zero());
  fq_store(b + 4, fq_mul(v, fq_add(fq_add(v, addr), rd)));
  fq_store(b + 5, v);
  fq_store(b + 6, fq_zero());
  fq_store(b + 7, fq_zero());
}

// w3_shifted: the instance's rows 1.. followed by a zero row (src/lib.rs:1667-1676, :925-929)
__global__ void k_wit_shift(const fq *__restrict__ w3, size_t rows, size_t width, const unsigned char *__restrict__ seg_last,
                            fq *__restrict__ out) {
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < rows * width; t += (size_t)gridDim.x * blockDim.x) {
    size_t q = t / width;
    fq_store(out + t, seg_last[q] ? fq_zero() : fq_load(w3 + t + width));
  }
}

__global__ void k_seg_flags(const unsigned long long *__restrict__ seg_end, int nseg, unsigned char *__restrict__ flags) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < nseg) flags[seg_end[s] - 1] = 1;
}

}  // namespace spg

using namespace spg;

extern "C" {

int spg_perm_scan(spg_ctx *ctx, size_t n, const size_t *seg_len, size_t n_seg, const spg_vec *v, size_t v_off,
                  size_t v_stride, const spg_vec *x, size_t x_off, size_t x_stride, spg_vec *D, size_t D_off,
                  size_t D_stride, spg_vec *pi, size_t pi_off, size_t pi_stride) {
  spg::DeviceGuard _dev(spg::ctx_of(ctx));
  SPG_CHECK(ctx && v && x && D && pi && seg_len, "spg_perm_scan: null argument");
  SPG_CHECK(n_seg >= 1 && n_seg <= (1u << 20), "spg_perm_scan: %zu segments", n_seg);
  SPG_CHECK(v_stride && x_stride && D_stride && pi_stride, "spg_perm_scan: zero stride");
  std::vector<unsigned long long> ends(n_seg);
  size_t tot = 0;
  for (size_t s = 0; s < n_seg; s++) {
    SPG_CHECK(seg_len[s] >= 1, "spg_perm_scan: empty segment %zu", s);
    tot += seg_len[s];
    ends[s] = tot;
  }
  SPG_CHECK(tot == n && n >= 1, "spg_perm_scan: segments cover %zu of %zu entries", tot, n);
  auto fits = [&](const spg_vec *w, size_t off, size_t stride) { return off + (n - 1) * stride < w->n; };
  SPG_CHECK(fits(v, v_off, v_stride) && fits(x, x_off, x_stride) && fits(D, D_off, D_stride) && fits(pi, pi_off, pi_stride),
            "spg_perm_scan: a strided view runs past its vector");
  // an output column must not coincide with an input column (the apply pass reads v[q+1] and
  // x[q] of rows it has not written yet, but only from columns it never writes)
  auto same = [](const spg_vec *a, size_t ao, size_t as, const spg_vec *b, size_t bo, size_t bs) {
    return a->d == b->d && as == bs && ao % as == bo % bs;
  };
  SPG_CHECK(!same(D, D_off, D_stride, v, v_off, v_stride) && !same(D, D_off, D_stride, x, x_off, x_stride) &&
                !same(pi, pi_off, pi_stride, v, v_off, v_stride) && !same(pi, pi_off, pi_stride, x, x_off, x_stride) &&
                !same(D, D_off, D_stride, pi, pi_off, pi_stride),
            "spg_perm_scan: output columns must differ from the input columns and from each other");
  size_t nchunks = (n + PS_CHUNK - 1) / PS_CHUNK;
  unsigned char *flags = nullptr;
  unsigned long long *d_ends = nullptr;
  Affine *agg = nullptr;
  fq *enter = nullptr;
  SPG_CUDA(dev_alloc(ctx, &flags, n));
  SPG_CUDA(dev_alloc(ctx, &d_ends, n_seg * sizeof(unsigned long long)));
  SPG_CUDA(dev_alloc(ctx, &agg, nchunks * sizeof(Affine)));
  SPG_CUDA(dev_alloc(ctx, &enter, nchunks * sizeof(fq)));
  SPG_CUDA(cudaMemsetAsync(flags, 0, n, ctx->stream));
  SPG_CUDA(cudaMemcpyAsync(d_ends, ends.data(), n_seg * sizeof(unsigned long long), cudaMemcpyHostToDevice, ctx->stream));
  SPG_LAUNCH(ctx, k_seg_flags, (unsigned)((n_seg + 255) / 256), 256, 0, d_ends, (int)n_seg, flags);
  Strided sv{v->d, v_off, v_stride}, sx{x->d, x_off, x_stride};
  StridedOut sD{D->d, D_off, D_stride}, sp{pi->d, pi_off, pi_stride};
  const size_t smem = 2 * PS_THREADS * sizeof(Affine);
  SPG_LAUNCH(ctx, k_perm_aggregate, (unsigned)nchunks, PS_THREADS, smem, sv, sx, flags, (unsigned long long)n, agg);
  SPG_LAUNCH(ctx, k_perm_chunks, 1, PS_THREADS, smem, agg, (unsigned long long)nchunks, enter);
  SPG_LAUNCH(ctx, k_perm_apply, (unsigned)nchunks, PS_THREADS, smem, sv, sx, flags, (unsigned long long)n, enter, sD, sp);
  // the host array `ends` was copied asynchronously from pageable memory: staged by the runtime
  // before the call returns, so it may go out of scope here
  dev_free(ctx, flags);
  dev_free(ctx, d_ends);
  dev_free(ctx, agg);
  dev_free(ctx, enter);
  return SPG_OK;
}


namespace {
// device flags marking the last row of every segment; caller frees
int seg_last_flags(spg_ctx *ctx, size_t n, const size_t *seg_len, size_t n_seg, const char *who, unsigned char **out) {
  SPG_CHECK(seg_len && n_seg >= 1 && n_seg <= (1u << 20), "%s: bad segment list", who);
  std::vector<unsigned long long> ends(n_seg);
  size_t tot = 0;
  for (size_t s = 0; s < n_seg; s++) {
    SPG_CHECK(seg_len[s] >= 1, "%s: empty segment %zu", who, s);
    tot += seg_len[s];
    ends[s] = tot;
  }
  SPG_CHECK(tot == n, "%s: segments cover %zu of %zu rows", who, tot, n);
  unsigned char *flags = nullptr;
  unsigned long long *d_ends = nullptr;
  SPG_CUDA(dev_alloc(ctx, &flags, n));
  cudaError_t e = dev_alloc(ctx, &d_ends, n_seg * sizeof(unsigned long long));
  if (e != cudaSuccess) {
    dev_free(ctx, flags);
    return cuda_fail(e, who, __FILE__, __LINE__);
  }
  int rc = [&]() -> int {
    SPG_CUDA(cudaMemsetAsync(flags, 0, n, ctx->stream));
    SPG_CUDA(cudaMemcpyAsync(d_ends, ends.data(), n_seg * sizeof(unsigned long long), cudaMemcpyHostToDevice, ctx->stream));
    SPG_LAUNCH(ctx, k_seg_flags, (unsigned)((n_seg + 255) / 256), 256, 0, d_ends, (int)n_seg, flags);
    return SPG_OK;
  }();
  dev_free(ctx, d_ends);
  if (rc != SPG_OK) {
    dev_free(ctx, flags);
    return rc;
  }
  *out = flags;
  return SPG_OK;
}
}  // namespace

int spg_wit_perm_w0(spg_ctx *ctx, const spg_fq *tau, const spg_fq *r, size_t used, size_t total, spg_vec **out) {
  spg::DeviceGuard _dev(spg::ctx_of(ctx));
  SPG_CHECK(ctx && tau && r && out, "spg_wit_perm_w0: null argument");
  SPG_CHECK(total >= 1 && used <= total, "spg_wit_perm_w0: %zu used entries of %zu", used, total);
  spg_vec *o = nullptr;
  SPG_TRY(vec_new(ctx, total, &o));
  fq ft, fr;
  memcpy(&ft, tau, 32);
  memcpy(&fr, r, 32);
  SPG_LAUNCH(ctx, k_perm_w0, (unsigned)((total + 127) / 128), 128, 0, ft, fr, used, total, o->d);
  *out = o;
  return SPG_OK;
}

int spg_wit_block(spg_ctx *ctx, int exec_mode, const spg_vec *vars, size_t rows, size_t vars_width, const spg_vec *perm_w0,
                  const spg_fq *tau, const spg_fq *r, size_t num_inputs_unpadded, size_t io_width, size_t phy_ops,
                  size_t vir_ops, size_t w2_width, const size_t *seg_len, size_t n_seg, spg_vec **w2_out, spg_vec **w3_out) {
  spg::DeviceGuard _dev(spg::ctx_of(ctx));
  SPG_CHECK(ctx && vars && perm_w0 && tau && r && seg_len && w2_out && w3_out, "spg_wit_block: null argument");
  const size_t n = num_inputs_unpadded;
  SPG_CHECK(n >= 2 && rows >= 1, "spg_wit_block: num_inputs_unpadded %zu, rows %zu", n, rows);
  if (exec_mode) phy_ops = vir_ops = 0;
  SPG_CHECK(vars->n >= rows * vars_width, "spg_wit_block: %zu rows of %zu scalars exceed the vector (%zu)", rows, vars_width, vars->n);
  SPG_CHECK(vars_width >= 2 * n && perm_w0->n >= 2 * (n - 1), "spg_wit_block: rows / perm_w0 shorter than the %zu inputs", 2 * n);
  SPG_CHECK(phy_ops + vir_ops == 0 || vars_width >= io_width + 2 * phy_ops + 4 * vir_ops,
            "spg_wit_block: rows of %zu scalars cannot hold %zu + %zu memory operations after %zu", vars_width, phy_ops, vir_ops, io_width);
  SPG_CHECK(w2_width >= 2 * n + 2 * phy_ops + 4 * vir_ops, "spg_wit_block: w2 rows of %zu scalars are too short", w2_width);
  spg_vec *w2 = nullptr, *w3 = nullptr;
  SPG_TRY(vec_new(ctx, rows * w2_width, &w2));
  int rc = vec_new(ctx, rows * 8, &w3);
  if (rc == SPG_OK) rc = [&]() -> int {
    WitArgs a;
    a.vars = vars->d;
    a.w0 = perm_w0->d;
    a.w2 = w2->d;
    a.w3 = w3->d;
    a.rows = rows;
    a.vars_width = vars_width;
    a.w2_width = w2_width;
    a.n = (unsigned)n;
    a.io_width = (unsigned)io_width;
    a.phy_ops = (unsigned)phy_ops;
    a.vir_ops = (unsigned)vir_ops;
    a.exec_mode = exec_mode;
    hfq hr = hfq_from(*r), hr2 = hfq_mul(hr, hr), hr3 = hfq_mul(hr2, hr);
    memcpy(&a.tau, tau, 32);
    memcpy(&a.r, &hr, 32);
    memcpy(&a.r2, &hr2, 32);
    memcpy(&a.r3, &hr3, 32);
    SPG_LAUNCH(ctx, k_wit_rows, (unsigned)((rows + 3) / 4), 128, 0, a);
    // (pi, D) of the INPUT pair, then of the PHY and VIR pairs whose x column is the last PMC / VMC
    // entry of the row in w2 (the constant column v when the instance has no such operations)
    SPG_TRY(spg_perm_scan(ctx, rows, seg_len, n_seg, w3, 0, 8, w3, 1, 8, w3, 3, 8, w3, 2, 8));
    if (!exec_mode) {
      if (phy_ops) SPG_TRY(spg_perm_scan(ctx, rows, seg_len, n_seg, w3, 0, 8, w2, 2 * n + 2 * (phy_ops - 1) + 1, w2_width, w3, 5, 8, w3, 4, 8));
      else SPG_TRY(spg_perm_scan(ctx, rows, seg_len, n_seg, w3, 0, 8, w3, 0, 8, w3, 5, 8, w3, 4, 8));
      if (vir_ops) SPG_TRY(spg_perm_scan(ctx, rows, seg_len, n_seg, w3, 0, 8, w2, 2 * n + 2 * phy_ops + 4 * (vir_ops - 1) + 3, w2_width, w3, 7, 8, w3, 6, 8));
      else SPG_TRY(spg_perm_scan(ctx, rows, seg_len, n_seg, w3, 0, 8, w3, 0, 8, w3, 7, 8, w3, 6, 8));
    }
    return SPG_OK;
  }();
  if (rc != SPG_OK) {
    spg_vec_free(w2);
    spg_vec_free(w3);
    return rc;
  }
  *w2_out = w2;
  *w3_out = w3;
  return SPG_OK;
}

int spg_wit_mem(spg_ctx *ctx, const spg_vec *mems, size_t rows, size_t in_width, const spg_fq *tau, const spg_fq *r,
                size_t mem_width, spg_vec **w2_out, spg_vec **w3_out) {
  spg::DeviceGuard _dev(spg::ctx_of(ctx));
  SPG_CHECK(ctx && mems && tau && r && w2_out && w3_out, "spg_wit_mem: null argument");
  SPG_CHECK(rows >= 1 && in_width >= 4 && mem_width >= 4 && mems->n >= rows * in_width, "spg_wit_mem: bad shape");
  spg_vec *w2 = nullptr, *w3 = nullptr;
  SPG_TRY(vec_new(ctx, rows * mem_width, &w2));
  int rc = vec_new(ctx, rows * 8, &w3);
  if (rc == SPG_OK) rc = [&]() -> int {
    fq ft, fr;
    memcpy(&ft, tau, 32);
    memcpy(&fr, r, 32);
    SPG_LAUNCH(ctx, k_wit_mem, (unsigned)((rows + 127) / 128), 128, 0, mems->d, rows, in_width, ft, fr, mem_width, w2->d, w3->d);
    size_t seg = rows;
    SPG_TRY(spg_perm_scan(ctx, rows, &seg, 1, w3, 0, 8, w3, 1, 8, w3, 3, 8, w3, 2, 8));
    return SPG_OK;
  }();
  if (rc != SPG_OK) {
    spg_vec_free(w2);
    spg_vec_free(w3);
    return rc;
  }
  *w2_out = w2;
  *w3_out = w3;
  return SPG_OK;
}

int spg_wit_shift(spg_ctx *ctx, const spg_vec *w3, size_t rows, size_t width, const size_t *seg_len, size_t n_seg,
                  spg_vec **out) {
  spg::DeviceGuard _dev(spg::ctx_of(ctx));
  SPG_CHECK(ctx && w3 && out, "spg_wit_shift: null argument");
  SPG_CHECK(rows >= 1 && width >= 1 && w3->n >= rows * width, "spg_wit_shift: %zu rows of %zu exceed the vector", rows, width);
  unsigned char *flags = nullptr;
  SPG_TRY(seg_last_flags(ctx, rows, seg_len, n_seg, "spg_wit_shift", &flags));
  spg_vec *o = nullptr;
  int rc = vec_new(ctx, rows * width, &o);
  if (rc == SPG_OK) rc = [&]() -> int {
    SPG_LAUNCH(ctx, k_wit_shift, grid_for(ctx, rows * width, 256), 256, 0, w3->d, rows, width, flags, o->d);
    return SPG_OK;
  }();
  dev_free(ctx, flags);
  if (rc != SPG_OK) {
    spg_vec_free(o);
    return rc;
  }
  *out = o;
  return SPG_OK;
}

}  // extern "C"
