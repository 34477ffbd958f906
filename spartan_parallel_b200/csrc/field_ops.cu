// Elementwise field ops, EqPolynomial expansion and DensePolynomial kernels.
//   reference: src/scalar/ristretto255.rs, src/dense_mlpoly.rs:60-131, 258-367
// All of these are one-pass streams over 32-byte scalars (LDG.256 / STG.256), so
// they are HBM-bound except for the modmul-heavy reductions.
#include "common.cuh"

namespace spg {

// ---------------------------------------------------------------- elementwise
template <int OP>
__global__ void k_vec_op(const fq *__restrict__ a, const fq *__restrict__ b, fq *__restrict__ out,
                         size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x) {
    fq x = fq_load(a + i);
    fq r;
    if (OP == 0) r = fq_mul(x, fq_load(b + i));
    else if (OP == 1) r = fq_add(x, fq_load(b + i));
    else if (OP == 2) r = fq_sub(x, fq_load(b + i));
    else if (OP == 3) r = fq_neg(x);
    else if (OP == 4) r = fq_sqr(x);
    else if (OP == 5) r = fq_from_mont(x);
    else r = fq_invert(x);
    fq_store(out + i, r);
  }
}

// Scalar::from_u512: d0*R2 + d1*R3
__global__ void k_from_u512(const uint32_t *__restrict__ wide, fq *__restrict__ out, size_t n) {
  fq R2, R3;
  R2.v[0] = 0x449c0f01u; R2.v[1] = 0xa40611e3u; R2.v[2] = 0x68859347u; R2.v[3] = 0xd00e1ba7u;
  R2.v[4] = 0x17f5be65u; R2.v[5] = 0xceec73d2u; R2.v[6] = 0x7c309a3du; R2.v[7] = 0x0399411bu;
  R3.v[0] = 0x7b83a2dbu; R3.v[1] = 0x2a9e4968u; R3.v[2] = 0xaef7f3ecu; R3.v[3] = 0x278324e6u;
  R3.v[4] = 0x04ec5b65u; R3.v[5] = 0x8065dc6cu; R3.v[6] = 0x3599cec7u; R3.v[7] = 0x0e530b77u;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x) {
    fq d0, d1;
#pragma unroll
    for (int k = 0; k < 8; k++) {
      d0.v[k] = wide[i * 16 + k];
      d1.v[k] = wide[i * 16 + 8 + k];
    }
    // d0, d1 are arbitrary 256-bit values: the Montgomery product stays below 2q
    // as long as the other operand is canonical (ristretto255.rs:455-461)
    fq x = fq_canon(fq_mul_lazy(R2, d0));
    fq y = fq_canon(fq_mul_lazy(R3, d1));
    fq_store(out + i, fq_add(x, y));
  }
}

// ---------------------------------------------------------------- eq expansion
// One doubling step of EqPolynomial::evals (dense_mlpoly.rs:81-90):
//   out[2i+1] = prev[i]*r, out[2i] = prev[i] - out[2i+1]
__global__ void k_eq_expand(const fq *__restrict__ prev, fq *__restrict__ out, size_t n, fq r) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x) {
    fq s = fq_load(prev + i);
    fq hi = fq_mul(s, r);
    fq lo = fq_sub(s, hi);
    fq_store(out + 2 * i, lo);
    fq_store(out + 2 * i + 1, hi);
  }
}

// The first levels (up to 2^LV entries) in one block through shared memory.
template <int LV>
__global__ void k_eq_expand_small(const fq *__restrict__ r, int ell, fq *__restrict__ out) {
  __shared__ fq buf[2][1 << LV];
  if (threadIdx.x == 0) buf[0][0] = fq_one();
  __syncthreads();
  int cur = 0;
  for (int j = 0; j < ell; j++) {
    size_t n = (size_t)1 << j;
    fq rj = r[j];
    for (size_t i = threadIdx.x; i < n; i += blockDim.x) {
      fq s = buf[cur][i];
      fq hi = fq_mul(s, rj);
      buf[cur ^ 1][2 * i + 1] = hi;
      buf[cur ^ 1][2 * i] = fq_sub(s, hi);
    }
    __syncthreads();
    cur ^= 1;
  }
  size_t n = (size_t)1 << ell;
  for (size_t i = threadIdx.x; i < n; i += blockDim.x) out[i] = buf[cur][i];
}

// K doubling steps in one launch: a thread takes entry i of the input level and produces its 2^K
// descendants in registers (field arithmetic is exact, so the values are those of K separate steps).
// ALL = false writes only the last level to out[i 2^K + t]; ALL = true also the levels in between, each at
// the place the suffix tables of the sumcheck keep it (level m at buf + 2^m, sc1.cu): out = buf, the
// input level is m0. One launch per level above the first 2^9 entries cost ~8 us of mostly launch overhead
// for a few microseconds of work, eleven times per table and two tables per proof.
struct EqSteps {
  fq r[3];
};
template <int K, bool ALL>
__global__ void __launch_bounds__(256)
k_eq_expand_multi(const fq *__restrict__ prev, size_t n, fq *__restrict__ out, unsigned int m0, const __grid_constant__ EqSteps rs) {
  static_assert(K >= 1 && K <= 3, "k_eq_expand_multi: 1..3 steps");
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    // one array per level (statically indexed: they stay in registers)
    fq l1[2], l2[4], l3[8];
    {
      fq s = fq_load(prev + i);
      fq hi = fq_mul(s, rs.r[0]);
      l1[0] = fq_sub(s, hi);
      l1[1] = hi;
    }
    if (K == 1 || ALL) {
      fq *lv = ALL ? out + ((size_t)1 << (m0 + 1)) + (i << 1) : out + (i << 1);
#pragma unroll
      for (int t = 0; t < 2; t++) fq_store(lv + t, l1[t]);
    }
    if (K >= 2) {
#pragma unroll
      for (int t = 0; t < 2; t++) {
        fq hi = fq_mul(l1[t], rs.r[1]);
        l2[2 * t] = fq_sub(l1[t], hi);
        l2[2 * t + 1] = hi;
      }
      if (K == 2 || ALL) {
        fq *lv = ALL ? out + ((size_t)1 << (m0 + 2)) + (i << 2) : out + (i << 2);
#pragma unroll
        for (int t = 0; t < 4; t++) fq_store(lv + t, l2[t]);
      }
    }
    if (K >= 3) {
#pragma unroll
      for (int t = 0; t < 4; t++) {
        fq hi = fq_mul(l2[t], rs.r[2]);
        l3[2 * t] = fq_sub(l2[t], hi);
        l3[2 * t + 1] = hi;
      }
      fq *lv = ALL ? out + ((size_t)1 << (m0 + 3)) + (i << 3) : out + (i << 3);
#pragma unroll
      for (int t = 0; t < 8; t++) fq_store(lv + t, l3[t]);
    }
  }
}

// `steps` (1..3) doubling steps from the n-entry level at prev; see k_eq_expand_multi
int eq_expand_steps(spg_ctx *ctx, const fq *prev, size_t n, fq *out, unsigned int m0, const spg_fq *r, int steps, bool all) {
  EqSteps rs;
  memset(&rs, 0, sizeof rs);
  for (int k = 0; k < steps; k++) memcpy(&rs.r[k], &r[k], sizeof(fq));
  int grid = grid_for(ctx, n, 256);
  if (all) {
    if (steps == 3) SPG_LAUNCH(ctx, (k_eq_expand_multi<3, true>), grid, 256, 0, prev, n, out, m0, rs);
    else if (steps == 2) SPG_LAUNCH(ctx, (k_eq_expand_multi<2, true>), grid, 256, 0, prev, n, out, m0, rs);
    else SPG_LAUNCH(ctx, (k_eq_expand_multi<1, true>), grid, 256, 0, prev, n, out, m0, rs);
  } else {
    if (steps == 3) SPG_LAUNCH(ctx, (k_eq_expand_multi<3, false>), grid, 256, 0, prev, n, out, m0, rs);
    else if (steps == 2) SPG_LAUNCH(ctx, (k_eq_expand_multi<2, false>), grid, 256, 0, prev, n, out, m0, rs);
    else SPG_LAUNCH(ctx, (k_eq_expand_multi<1, false>), grid, 256, 0, prev, n, out, m0, rs);
  }
  return SPG_OK;
}

constexpr int EQ_SMALL_LV = 9;

// evals of eq(r, .) with MSB <-> r[0]; r on device (ell scalars). out has 2^ell entries.
// scratch must hold 2^(ell-1) entries when ell > EQ_SMALL_LV.
int eq_evals_device(spg_ctx *ctx, const fq *d_r, const spg_fq *h_r, size_t ell, fq *out, fq *scratch) {
  int small = (int)(ell < (size_t)EQ_SMALL_LV ? ell : EQ_SMALL_LV);
  if (ell <= (size_t)EQ_SMALL_LV) {
    SPG_LAUNCH(ctx, k_eq_expand_small<EQ_SMALL_LV>, 1, 256, 0, d_r, small, out);
    return SPG_OK;
  }
  // ping-pong, three levels per launch, so that the last level lands in `out`
  size_t remaining = ell - small;
  size_t launches = (remaining + 2) / 3;
  fq *bufs[2] = {out, scratch};
  int which = (launches % 2 == 0) ? 0 : 1;
  SPG_LAUNCH(ctx, k_eq_expand_small<EQ_SMALL_LV>, 1, 256, 0, d_r, small, bufs[which]);
  for (size_t j = small; j < ell;) {
    int steps = (int)(ell - j < 3 ? ell - j : 3);
    SPG_TRY(eq_expand_steps(ctx, bufs[which], (size_t)1 << j, bufs[which ^ 1], 0, h_r + j, steps, false));
    which ^= 1;
    j += steps;
  }
  return SPG_OK;
}

// ---------------------------------------------------------------- dense bind
// top: Z[i] += r*(Z[i+n]-Z[i]); bot: Z[i] = Z[2i] + r*(Z[2i+1]-Z[2i]) (out of place)
__global__ void k_bound_top(fq *__restrict__ Z, size_t n, fq r) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x) {
    fq lo = fq_load(Z + i), hi = fq_load(Z + i + n);
    fq_store(Z + i, fq_add(lo, fq_mul(r, fq_sub(hi, lo))));
  }
}

__global__ void k_bound_bot(const fq *__restrict__ Z, fq *__restrict__ out, size_t n, fq r) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x) {
    fq lo = fq_load(Z + 2 * i), hi = fq_load(Z + 2 * i + 1);
    fq_store(out + i, fq_add(lo, fq_mul(r, fq_sub(hi, lo))));
  }
}

// ---------------------------------------------------------------- reductions
// partial[b] = sum over the block's elements of a[i]*b[i]
__global__ void k_dot(const fq *__restrict__ a, const fq *__restrict__ b, size_t n,
                      fq *__restrict__ partials) {
  __shared__ fq sm[32];
  fq acc[1] = {fq_zero()};
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x)
    acc[0] = fq_add(acc[0], fq_mul(fq_load_stream(a + i), fq_load_stream(b + i)));
  block_sum<1>(acc, sm);
  if (threadIdx.x == 0) partials[blockIdx.x] = acc[0];
}

// DensePolynomial::evaluate without materialising the 2^ell chi table:
//   Z(r) = sum_j L[j] * sum_i R[i] * Z[j*Rs + i], L = eq(r_hi), R = eq(r_lo).
// One block handles `chunk` consecutive elements of one row j.
__global__ void k_eval_rows(const fq *__restrict__ Z, const fq *__restrict__ L,
                            const fq *__restrict__ R, size_t Rs, size_t chunk, size_t chunks_per_row,
                            fq *__restrict__ partials) {
  __shared__ fq sm[32];
  size_t j = blockIdx.x / chunks_per_row, c = blockIdx.x % chunks_per_row;
  size_t base = c * chunk;
  fq acc[1] = {fq_zero()};
  for (size_t i = base + threadIdx.x; i < base + chunk && i < Rs; i += blockDim.x)
    acc[0] = fq_add(acc[0], fq_mul(fq_load_stream(Z + j * Rs + i), fq_load(R + i)));
  block_sum<1>(acc, sm);
  if (threadIdx.x == 0) partials[blockIdx.x] = fq_mul(acc[0], L[j]);
}

// bound(L): partial[y][i] = sum_{j in slab y} L[j] * Z[j*Rs + i]
__global__ void k_bound_L_partial(const fq *__restrict__ Z, const fq *__restrict__ L, size_t Ls,
                                  size_t Rs, size_t slab, fq *__restrict__ partial) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Rs) return;
  size_t j0 = (size_t)blockIdx.y * slab, j1 = j0 + slab < Ls ? j0 + slab : Ls;
  fq acc = fq_zero();
  for (size_t j = j0; j < j1; j++) acc = fq_add(acc, fq_mul(L[j], fq_load_stream(Z + j * Rs + i)));
  fq_store(partial + (size_t)blockIdx.y * Rs + i, acc);
}

__global__ void k_sum_slabs(const fq *__restrict__ partial, size_t nslabs, size_t Rs,
                            fq *__restrict__ out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Rs) return;
  fq acc = fq_zero();
  for (size_t y = 0; y < nslabs; y++) acc = fq_add(acc, fq_load(partial + y * Rs + i));
  fq_store(out + i, acc);
}

// Modular all-reduce of one table replicated-by-address on `world` GPUs of one node, over
// peer memory: rank r owns chunk r. A thread loads element i of that chunk from every peer
// (P2P loads over NVLink), adds them, and stores the sum back into every peer's table (P2P
// stores), so each rank moves 2 (G-1)/G of the table instead of receiving G-1 whole copies
// and adding them in separate passes. The same thread reads and writes element i everywhere,
// so the update is in place; visibility is by kernel boundaries + a host barrier.
struct PeerPtrs {
  fq *p[16];
};
// scatter_only: the sum of chunk `begin` stays with its owner (P.p[0]) -- a reduce-scatter, half the traffic
__global__ void k_peer_sum(const __grid_constant__ PeerPtrs P, int world, size_t begin, size_t count, int scatter_only) {
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < count; t += (size_t)gridDim.x * blockDim.x) {
    size_t i = begin + t;
    fq acc = fq_load_cg(P.p[0] + i);
    for (int r = 1; r < world; r++) acc = fq_add(acc, fq_load_cg(P.p[r] + i));
    const int nw = scatter_only ? 1 : world;
    for (int r = 0; r < nw; r++) fq_store(P.p[r] + i, acc);
  }
}

int dense_evaluate_device(spg_ctx *ctx, const fq *Z, size_t n, const spg_fq *r, size_t ell, fq *d_out) {
  // split r = (r_hi | r_lo) like compute_factored_lens (dense_mlpoly.rs:118-120)
  size_t left = ell / 2, right = ell - left;
  size_t Ls = (size_t)1 << left, Rs = (size_t)1 << right;
  fq *tabs = nullptr, *d_r = nullptr;
  SPG_CUDA(dev_alloc(ctx, &tabs, (Ls + Rs + Rs) * sizeof(fq)));
  SPG_CUDA(dev_alloc(ctx, &d_r, (ell ? ell : 1) * sizeof(fq)));
  SPG_CUDA(cudaMemcpyAsync(d_r, r, ell * sizeof(fq), cudaMemcpyHostToDevice, ctx->stream));
  fq *dL = tabs, *dR = tabs + Ls, *scratch = tabs + Ls + Rs;
  int rc = eq_evals_device(ctx, d_r, r, left, dL, scratch);
  if (rc == SPG_OK) rc = eq_evals_device(ctx, d_r + left, r + left, right, dR, scratch);
  if (rc == SPG_OK) {
    size_t chunk = Rs < 2048 ? Rs : 2048;
    size_t cpr = (Rs + chunk - 1) / chunk;
    size_t nblocks = Ls * cpr;
    rc = ensure_partials(ctx, nblocks);
    if (rc == SPG_OK) {
      int threads = chunk >= 256 ? 256 : 64;
      k_eval_rows<<<(unsigned)nblocks, threads, 0, ctx->stream>>>(Z, dL, dR, Rs, chunk, cpr, ctx->d_partials);
      ctx->launches++;
      rc = reduce_partials(ctx, ctx->d_partials, nblocks, 1, d_out);
    }
  }
  (void)n;
  cudaStreamSynchronize(ctx->stream);
  dev_free(ctx, tabs);
  dev_free(ctx, d_r);
  return rc;
}


// Self-test of the wide-range forms of fq.cuh (fq_sub_plus2q / fq_sub_plus6q / fq_fold2q and fq_mul_lazy
// on operands beyond [0, 2q)): the bind and the two evaluation terms exactly as k_rows_rolled computes
// them, on operands in [0, 2q) that include the edges (0, 1, q - 1, q, q + 1, 2q - 1), against the same
// quantities from canonical arithmetic (fq_add / fq_sub / fq_mul on canonicalised operands).
__global__ void k_fq_wide_selftest(size_t n, uint64_t seed, unsigned int *__restrict__ bad) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  uint64_t st = seed + 0x9E3779B97F4A7C15ull * (t + 1);
  auto rnd = [&]() {
    st += 0x9E3779B97F4A7C15ull;
    uint64_t z = st;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
  };
  const uint32_t Q[8] = {SPG_Q0, SPG_Q1, SPG_Q2, SPG_Q3, 0, 0, 0, SPG_Q7};
  const uint32_t Q2[8] = {SPG_2Q0, SPG_2Q1, SPG_2Q2, SPG_2Q3, 0, 0, 0, SPG_2Q7};
  // value number `sel` of [0, 2q): edges for sel < 6, else uniform below 2q
  auto pick = [&](unsigned sel) {
    fq x = fq_zero();
    if (sel == 1) x.v[0] = 1;
    else if (sel >= 2 && sel <= 4) {  // q - 1, q, q + 1
      for (int i = 0; i < 8; i++) x.v[i] = Q[i];
      x.v[0] += sel - 3;                // limb 0 of q is far from 0 and 2^32 - 1: no carry
    } else if (sel == 5) {              // 2q - 1
      for (int i = 0; i < 8; i++) x.v[i] = Q2[i];
      x.v[0] -= 1;
    } else if (sel >= 6) {
      for (int i = 0; i < 8; i += 2) {
        uint64_t r = rnd();
        x.v[i] = (uint32_t)r;
        x.v[i + 1] = (uint32_t)(r >> 32);
      }
      x.v[7] &= 0x1fffffffu;            // < 2^253 < 2q ... then fold anything >= 2q
      x = fq_fold2q(x);
    }
    return x;
  };
  unsigned m = (unsigned)(t % 1296);  // 6^4 edge combinations for (lo, hi, other, acc), the rest random
  bool edges = t < 4 * 1296;
  fq lo = pick(edges ? m % 6 : 6), hi = pick(edges ? (m / 6) % 6 : 6), ot = pick(edges ? (m / 36) % 6 : 6),
     acc = pick(edges ? (m / 216) % 6 : 6);
  fq r = fq_canon(pick(6)), w = fq_canon(pick(edges ? 2 + (unsigned)(t / 1296) % 4 : 6));
  unsigned int err = 0;
  fq clo = fq_canon(lo), chi = fq_canon(hi), cot = fq_canon(ot), cacc = fq_canon(acc);
  // bind
  fq v = fq_fold2q(fq_raw_add(lo, fq_mul_lazy(r, fq_sub_plus2q(hi, lo))));
  fq want = fq_add(clo, fq_mul(r, fq_sub(chi, clo)));
  if (!fq_equal(fq_canon(v), want)) err |= 1;
  // t = 0 term: acc + w (a0 b0 - c0) with (a0, b0, c0) = (lo, hi, ot)
  fq e0 = fq_fold2q(fq_raw_add(acc, fq_mul_lazy(w, fq_sub_plus2q(fq_mul_lazy(lo, hi), ot))));
  fq want0 = fq_add(cacc, fq_mul(w, fq_sub(fq_mul(clo, chi), cot)));
  if (!fq_equal(fq_canon(e0), want0)) err |= 2;
  // t = 2 term with a = (lo, hi), b = (hi, ot), c = (ot, lo)
  fq a2 = fq_raw_add(hi, fq_sub_plus2q(hi, lo)), b2 = fq_raw_add(ot, fq_sub_plus2q(ot, hi)), c2 = fq_raw_add(lo, fq_sub_plus2q(lo, ot));
  fq e2 = fq_fold2q(fq_raw_add(acc, fq_mul_lazy(w, fq_sub_plus6q(fq_mul_lazy(a2, b2), c2))));
  fq ca2 = fq_sub(fq_add(chi, chi), clo), cb2 = fq_sub(fq_add(cot, cot), chi), cc2 = fq_sub(fq_add(clo, clo), cot);
  fq want2 = fq_add(cacc, fq_mul(w, fq_sub(fq_mul(ca2, cb2), cc2)));
  if (!fq_equal(fq_canon(e2), want2)) err |= 4;
  // the folded values really are below 2q (fq_canon of something >= 2q would not be canonical)
  fq q_minus_1 = pick(2);
  auto below_q = [&](const fq &x) {  // x <= q - 1
    fq d = fq_sub_plus(q_minus_1, x, 0, 0, 0, 0, 0);
    return (d.v[7] >> 31) == 0;      // no wrap: q - 1 - x >= 0 (both far below 2^255)
  };
  if (!below_q(fq_canon(v)) || !below_q(fq_canon(e0)) || !below_q(fq_canon(e2))) err |= 8;
  if (err) atomicOr(bad, err);
}

}  // namespace spg

using namespace spg;

extern "C" {

int spg_fq_vec_op(spg_ctx *ctx, int op, const spg_vec *a, const spg_vec *b, spg_vec *out) {
  spg::DeviceGuard _dev(spg::ctx_of(ctx));
  SPG_CHECK(ctx && a && out, "spg_fq_vec_op: null argument");
  SPG_CHECK(op >= 0 && op <= 6, "spg_fq_vec_op: unknown op %d", op);
  bool binary = op <= 2;
  SPG_CHECK(!binary || (b && b->n == a->n), "spg_fq_vec_op: operand length mismatch");
  SPG_CHECK(out->n == a->n, "spg_fq_vec_op: output length mismatch");
  size_t n = a->n;
  if (n == 0) return SPG_OK;
  int grid = grid_for(ctx, n, 256);
  const fq *pb = b ? b->d : a->d;
  switch (op) {
    case 0: SPG_LAUNCH(ctx, k_vec_op<0>, grid, 256, 0, a->d, pb, out->d, n); break;
    case 1: SPG_LAUNCH(ctx, k_vec_op<1>, grid, 256, 0, a->d, pb, out->d, n); break;
    case 2: SPG_LAUNCH(ctx, k_vec_op<2>, grid, 256, 0, a->d, pb, out->d, n); break;
    case 3: SPG_LAUNCH(ctx, k_vec_op<3>, grid, 256, 0, a->d, pb, out->d, n); break;
    case 4: SPG_LAUNCH(ctx, k_vec_op<4>, grid, 256, 0, a->d, pb, out->d, n); break;
    case 5: SPG_LAUNCH(ctx, k_vec_op<5>, grid, 256, 0, a->d, pb, out->d, n); break;
    default: SPG_LAUNCH(ctx, k_vec_op<6>, grid, 256, 0, a->d, pb, out->d, n); break;
  }
  return SPG_OK;
}

int spg_fq_from_u512(spg_ctx *ctx, const uint64_t *host_wide, size_t n, spg_vec **out) {
  spg::DeviceGuard _dev(spg::ctx_of(ctx));
  SPG_CHECK(ctx && out && (host_wide || n == 0), "spg_fq_from_u512: null argument");
  spg_vec *v = nullptr;
  SPG_TRY(vec_new(ctx, n, &v));
  if (n) {
    uint32_t *d_wide = nullptr;
    SPG_CUDA(dev_alloc(ctx, &d_wide, n * 64));
    SPG_CUDA(cudaMemcpyAsync(d_wide, host_wide, n * 64, cudaMemcpyHostToDevice, ctx->stream));
    SPG_LAUNCH(ctx, k_from_u512, grid_for(ctx, n, 256), 256, 0, d_wide, v->d, n);
    SPG_CUDA(cudaStreamSynchronize(ctx->stream));
    dev_free(ctx, d_wide);
  }
  *out = v;
  return SPG_OK;
}

int spg_eq_evals(spg_ctx *ctx, const spg_fq *r, size_t ell, spg_vec **out) {
  spg::DeviceGuard _dev(spg::ctx_of(ctx));
  SPG_CHECK(ctx && out && (r || ell == 0), "spg_eq_evals: null argument");
  SPG_CHECK(ell <= 34, "spg_eq_evals: ell = %zu too large", ell);
  size_t n = (size_t)1 << ell;
  spg_vec *v = nullptr;
  SPG_TRY(vec_new(ctx, n, &v));
  fq *d_r = nullptr, *scratch = nullptr;
  SPG_CUDA(dev_alloc(ctx, &d_r, (ell ? ell : 1) * sizeof(fq)));
  SPG_CUDA(cudaMemcpyAsync(d_r, r, ell * sizeof(fq), cudaMemcpyHostToDevice, ctx->stream));
  if (ell > (size_t)EQ_SMALL_LV) SPG_CUDA(dev_alloc(ctx, &scratch, (n / 2) * sizeof(fq)));
  int rc = eq_evals_device(ctx, d_r, r, ell, v->d, scratch);
  cudaStreamSynchronize(ctx->stream);
  dev_free(ctx, d_r);
  if (scratch) dev_free(ctx, scratch);
  if (rc != SPG_OK) {
    spg_vec_free(v);
    return rc;
  }
  *out = v;
  return SPG_OK;
}

int spg_dense_bound_top(spg_ctx *ctx, spg_vec *v, const spg_fq *r) {
  spg::DeviceGuard _dev(spg::ctx_of(ctx));
  SPG_CHECK(ctx && v && r, "spg_dense_bound_top: null argument");
  SPG_CHECK(v->n >= 2 && is_pow2(v->n), "spg_dense_bound_top: length %zu is not a power of two >= 2", v->n);
  size_t n = v->n / 2;
  fq rr;
  memcpy(&rr, r, sizeof rr);
  SPG_LAUNCH(ctx, k_bound_top, grid_for(ctx, n, 256), 256, 0, v->d, n, rr);
  v->n = n;
  return SPG_OK;
}

int spg_dense_bound_bot(spg_ctx *ctx, spg_vec *v, const spg_fq *r) {
  spg::DeviceGuard _dev(spg::ctx_of(ctx));
  SPG_CHECK(ctx && v && r, "spg_dense_bound_bot: null argument");
  SPG_CHECK(v->n >= 2 && is_pow2(v->n), "spg_dense_bound_bot: length %zu is not a power of two >= 2", v->n);
  SPG_CHECK(v->owned, "spg_dense_bound_bot: vector must be library-owned");
  size_t n = v->n / 2;
  fq rr;
  memcpy(&rr, r, sizeof rr);
  DevTmp tmp(ctx);
  SPG_CUDA(tmp.alloc(n * sizeof(fq)));
  SPG_LAUNCH(ctx, k_bound_bot, grid_for(ctx, n, 256), 256, 0, v->d, tmp.as<fq>(), n, rr);
  dev_free(ctx, v->d);
  v->d = tmp.as<fq>();
  tmp.p = nullptr;  // now owned by the vector
  v->n = v->cap = n;
  return SPG_OK;
}

int spg_dense_evaluate(spg_ctx *ctx, const spg_vec *v, const spg_fq *r, size_t ell, spg_fq *out) {
  spg::DeviceGuard _dev(spg::ctx_of(ctx));
  SPG_CHECK(ctx && v && out && (r || ell == 0), "spg_dense_evaluate: null argument");
  SPG_CHECK(v->n == ((size_t)1 << ell), "spg_dense_evaluate: len %zu != 2^%zu", v->n, ell);
  SPG_TRY(dense_evaluate_device(ctx, v->d, v->n, r, ell, ctx->d_result));
  return fetch_result(ctx, 1, out);
}

int spg_dot(spg_ctx *ctx, const spg_vec *a, const spg_vec *b, spg_fq *out) {
  spg::DeviceGuard _dev(spg::ctx_of(ctx));
  SPG_CHECK(ctx && a && b && out, "spg_dot: null argument");
  SPG_CHECK(a->n == b->n, "spg_dot: length mismatch");
  int grid = grid_for(ctx, a->n, 256, 4);
  SPG_TRY(ensure_partials(ctx, grid));
  SPG_LAUNCH(ctx, k_dot, grid, 256, 0, a->d, b->d, a->n, ctx->d_partials);
  SPG_TRY(reduce_partials(ctx, ctx->d_partials, grid, 1, ctx->d_result));
  return fetch_result(ctx, 1, out);
}

int spg_dense_bound_L(spg_ctx *ctx, const spg_vec *v, const spg_fq *L, size_t L_size, spg_vec **out) {
  spg::DeviceGuard _dev(spg::ctx_of(ctx));
  SPG_CHECK(ctx && v && L && out, "spg_dense_bound_L: null argument");
  SPG_CHECK(L_size && v->n % L_size == 0, "spg_dense_bound_L: L_size %zu does not divide len %zu", L_size, v->n);
  size_t Rs = v->n / L_size;
  VecOut o;
  SPG_TRY(vec_new(ctx, Rs, &o.v));
  size_t slab = 64;
  size_t nslabs = (L_size + slab - 1) / slab;
  DevTmp tL(ctx), tpart(ctx);
  SPG_CUDA(tL.alloc(L_size * sizeof(fq)));
  SPG_CUDA(tpart.alloc(nslabs * Rs * sizeof(fq)));
  SPG_CUDA(cudaMemcpyAsync(tL.p, L, L_size * sizeof(fq), cudaMemcpyHostToDevice, ctx->stream));
  dim3 grid((unsigned)((Rs + 127) / 128), (unsigned)nslabs);
  SPG_LAUNCH(ctx, k_bound_L_partial, grid, 128, 0, v->d, tL.as<fq>(), L_size, Rs, slab, tpart.as<fq>());
  SPG_LAUNCH(ctx, k_sum_slabs, (unsigned)((Rs + 127) / 128), 128, 0, tpart.as<fq>(), nslabs, Rs, o.v->d);
  SPG_CUDA(cudaStreamSynchronize(ctx->stream));
  *out = o.release();
  return SPG_OK;
}

}  // extern "C"

extern "C" {

int spg_peer_alloc(spg_ctx *ctx, size_t n, spg_vec **out, uint8_t handle[64]) {
  spg::DeviceGuard _dev(spg::ctx_of(ctx));
  SPG_CHECK(ctx && out && handle && n, "spg_peer_alloc: null argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  fq *d = nullptr;
  SPG_CUDA(cudaMalloc(&d, n * sizeof(fq)));  // IPC needs a plain allocation, not pool memory
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, d);
  if (e != cudaSuccess) {
    cudaFree(d);
    return cuda_fail(e, "cudaIpcGetMemHandle", __FILE__, __LINE__);
  }
  memcpy(handle, &h, 64);
  spg_vec *v = nullptr;
  int rc = spg_vec_wrap(ctx, d, n, &v);
  if (rc != SPG_OK) {
    cudaFree(d);
    return rc;
  }
  *out = v;
  return SPG_OK;
}

int spg_peer_free(spg_vec *v) {
  spg::DeviceGuard _dev(spg::ctx_of(v));
  if (!v) return SPG_OK;
  void *d = v->d;
  spg_vec_free(v);
  SPG_CUDA(cudaFree(d));
  return SPG_OK;
}

int spg_peer_open(spg_ctx *ctx, const uint8_t handle[64], void **ptr) {
  spg::DeviceGuard _dev(spg::ctx_of(ctx));
  SPG_CHECK(ctx && handle && ptr, "spg_peer_open: null argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, 64);
  SPG_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return SPG_OK;
}

int spg_peer_close(void *ptr) {
  if (ptr) SPG_CUDA(cudaIpcCloseMemHandle(ptr));
  return SPG_OK;
}

static int peer_sum_impl(spg_ctx *ctx, void *const *peer_ptrs, int world, int rank, size_t n, int scatter_only);
int spg_peer_sum(spg_ctx *ctx, void *const *peer_ptrs, int world, int rank, size_t n) {
  return peer_sum_impl(ctx, peer_ptrs, world, rank, n, 0);
}
int spg_peer_reduce_scatter(spg_ctx *ctx, void *const *peer_ptrs, int world, int rank, size_t n) {
  return peer_sum_impl(ctx, peer_ptrs, world, rank, n, 1);
}
static int peer_sum_impl(spg_ctx *ctx, void *const *peer_ptrs, int world, int rank, size_t n, int scatter_only) {
  spg::DeviceGuard _dev(spg::ctx_of(ctx));
  SPG_CHECK(ctx && peer_ptrs, "spg_peer_sum: null argument");
  SPG_CHECK(world >= 1 && world <= 16 && rank >= 0 && rank < world, "spg_peer_sum: bad rank %d of %d", rank, world);
  SPG_CHECK(n % (size_t)world == 0, "spg_peer_sum: %zu scalars do not split over %d ranks", n, world);
  PeerPtrs P;
  memset(&P, 0, sizeof P);
  // own table first so that at least one operand is a local load
  P.p[0] = (fq *)peer_ptrs[rank];
  for (int r = 0, k = 1; r < world; r++)
    if (r != rank) P.p[k++] = (fq *)peer_ptrs[r];
  size_t chunk = n / world;
  ctx->next_units = (scatter_only ? 32.0 : 64.0) * (double)chunk * (double)(world - 1);
  SPG_LAUNCH(ctx, k_peer_sum, grid_for(ctx, chunk, 256), 256, 0, P, world, (size_t)rank * chunk, chunk, scatter_only);
  return SPG_OK;
}

int spg_debug_fq_wide_selftest(spg_ctx *ctx, size_t n, uint64_t seed, uint32_t *out_bad) {
  spg::DeviceGuard _dev(spg::ctx_of(ctx));
  SPG_CHECK(ctx && out_bad && n >= 1, "spg_debug_fq_wide_selftest: bad argument");
  DevTmp d_bad(ctx);
  SPG_CUDA(d_bad.alloc(sizeof(unsigned int)));
  SPG_CUDA(cudaMemsetAsync(d_bad.p, 0, sizeof(unsigned int), ctx->stream));
  SPG_LAUNCH(ctx, k_fq_wide_selftest, (unsigned)((n + 127) / 128), 128, 0, n, seed, d_bad.as<unsigned int>());
  SPG_CUDA(cudaMemcpyAsync(out_bad, d_bad.p, sizeof(unsigned int), cudaMemcpyDeviceToHost, ctx->stream));
  SPG_CUDA(cudaStreamSynchronize(ctx->stream));
  return SPG_OK;
}

}  // extern "C"
