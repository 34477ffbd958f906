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

}  // extern "C"
