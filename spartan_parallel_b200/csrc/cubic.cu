// Grand-product circuits and the batched cubic sumcheck of the sparse-polynomial
// memory check.
//   ProductCircuit::{new, compute_layer, evaluate}      src/product_tree.rs:17-64
//   SumcheckInstanceProof::prove_cubic_batched          src/sumcheck.rs:264-434
//   Layers::build_hash_layer / deref_mem                src/sparse_mlpoly.rs:612-687, 255-264
// These tables are plain DensePolynomials bound with bound_poly_var_top: the pair
// (i, i + n/2) is two perfectly coalesced streams, so the reference order is kept.
#include "common.cuh"
#include "rounds.cuh"

namespace spg {

constexpr int CB = 128;

// value of the line through (0, lo), (1, hi) at 2 and 3
__device__ __forceinline__ void cline23(const fq &lo, const fq &hi, fq &at2, fq &at3) {
  fq d = fq_sub(hi, lo);
  at2 = fq_add(hi, d);
  at3 = fq_add(at2, d);
}

struct CubicPtrs {
  const fq *A[24];
  const fq *B[24];
  const fq *C[24];
};

// One round's evaluations of all triples (sumcheck.rs:297-371) with the random linear combination
// of the triples (:369-371) and the final reduction done on the device: a 1-D grid of ntriples * gx blocks, block b works on
// triple b / gx, scales its three sums by that triple's coefficient, and the block that draws
// the last ticket adds everything up and publishes the three combined evaluations through the
// mapped result slot (finish_block, common.cuh) -- one launch and no stream synchronisation per
// round instead of eval + reduce + copy + sync. Field arithmetic is exact, so
// sum_k c_k (sum_b partial) and sum_{k,b} c_k partial are the same canonical scalar.
struct CubicCoeffs {
  fq c[24];
};
__global__ void __launch_bounds__(CB)
k_cubic_eval_rlc(const __grid_constant__ CubicPtrs P, const __grid_constant__ CubicCoeffs K, int gx, size_t half,
                 FinishArgs fa) {
  __shared__ fq sm[3 * 32];
  const int k = blockIdx.x / gx, bx = blockIdx.x % gx;
  const fq *__restrict__ A = P.A[k];
  const fq *__restrict__ B = P.B[k];
  const fq *__restrict__ C = P.C[k];
  fq acc[3] = {fq_zero(), fq_zero(), fq_zero()};
  for (size_t i = (size_t)bx * CB + threadIdx.x; i < half; i += (size_t)gx * CB) {
    fq a0 = fq_load(A + i), a1 = fq_load(A + i + half);
    fq b0 = fq_load(B + i), b1 = fq_load(B + i + half);
    fq c0 = fq_load(C + i), c1 = fq_load(C + i + half);
    fq a2, a3, b2, b3, c2, c3;
    cline23(a0, a1, a2, a3);
    cline23(b0, b1, b2, b3);
    cline23(c0, c1, c2, c3);
    acc[0] = fq_add(acc[0], fq_mul(fq_mul(a0, b0), c0));
    acc[1] = fq_add(acc[1], fq_mul(fq_mul(a2, b2), c2));
    acc[2] = fq_add(acc[2], fq_mul(fq_mul(a3, b3), c3));
  }
  block_sum<3>(acc, sm);
  if (threadIdx.x == 0) {
    fq ck = K.c[k];
#pragma unroll
    for (int t = 0; t < 3; t++) acc[t] = fq_mul(ck, acc[t]);
  }
  finish_block<3>(fa, acc, sm);
}

// The same round on small tables, where it is a dependent chain and not throughput (rounds.cuh,
// k_quad_split): one item over four lanes, lane l = 0, 1, 2 evaluating at t = 0, 2, 3, so the three
// sums of a block are reduced by ONE butterfly, the triple's coefficient is applied by three lanes at
// once, and the final reduction over the blocks runs side by side as well. A block works on one triple.
constexpr int CUBIC_SPLIT_ITEMS = 32;  // items per 128-thread block
__global__ void __launch_bounds__(128)
k_cubic_eval_split(const __grid_constant__ CubicPtrs P, const __grid_constant__ CubicCoeffs K, int gx, size_t half, FinishArgs fa) {
  __shared__ fq sm[4 * 4];
  const unsigned int lane = threadIdx.x & 31, l = lane & 3, warp = threadIdx.x >> 5;
  const int k = blockIdx.x / gx, bx = blockIdx.x % gx;
  const fq *__restrict__ A = P.A[k];
  const fq *__restrict__ B = P.B[k];
  const fq *__restrict__ C = P.C[k];
  const size_t i = (size_t)bx * CUBIC_SPLIT_ITEMS + (threadIdx.x >> 2);
  fq acc = fq_zero();
  if (i < half && l < 3) {
    fq a = split_point(fq_load(A + i), fq_load(A + i + half), l);
    fq b = split_point(fq_load(B + i), fq_load(B + i + half), l);
    fq c = split_point(fq_load(C + i), fq_load(C + i + half), l);
    acc = fq_mul(fq_mul(a, b), c);
  }
  // lanes 4 j + l hold point l of item j
  acc = fq_add(acc, fq_shfl_down(acc, 4));
  acc = fq_add(acc, fq_shfl_down(acc, 8));
  acc = fq_add(acc, fq_shfl_down(acc, 16));
  if (lane < 3) sm[warp * 4 + lane] = acc;
  __syncthreads();
  fq mine[3] = {fq_zero(), fq_zero(), fq_zero()};
  if (warp == 0) {
    fq t = (lane < 16 && l < 3) ? sm[(lane >> 2) * 4 + l] : fq_zero();
    t = fq_add(t, fq_shfl_down(t, 4));
    t = fq_add(t, fq_shfl_down(t, 8));
    t = fq_mul(K.c[k], t);  // lanes 0 .. 2: the three sums of this block, scaled by the triple's coefficient
    mine[0] = t;
    mine[1] = fq_shfl(t, 1);
    mine[2] = fq_shfl(t, 2);
  }
  finish_block_lanes3(fa, mine, sm);
}

// Fused bind_j + eval_{j+1} (sumcheck.rs:373-407 then :297-371 of the next round): an item binds the four
// entries i, i + q, i + 2q, i + 3q (q = a quarter of the current length) of each table of its triple with
// r_j -- the bound pair (i, i + q) is exactly what round j + 1 evaluates -- stores the pair and evaluates
// it. A, B (and a sequential triple's own C) are bound in place: an item reads and writes only its own
// four / two positions. The C table SHARED by the parallel triples is read by every one of them, so its
// bound form goes to a second buffer (Cout of the first parallel triple, null for the others, which bind
// their copy of the pair in registers only); the two buffers swap roles every round. One launch per round
// instead of a bind launch and an evaluation launch.
struct CubicBindPtrs {
  fq *A[24];
  fq *B[24];
  const fq *C[24];
  fq *Cout[24];
};
__device__ __forceinline__ fq cbind(const fq &lo, const fq &hi, const fq &r) { return fq_add(lo, fq_mul(r, fq_sub(hi, lo))); }

__global__ void __launch_bounds__(CB)
k_cubic_bind_eval_rlc(const __grid_constant__ CubicBindPtrs P, const __grid_constant__ CubicCoeffs K, int gx, size_t quarter, fq r,
                      FinishArgs fa) {
  __shared__ fq sm[3 * 32];
  const int k = blockIdx.x / gx, bx = blockIdx.x % gx;
  fq *A = P.A[k], *B = P.B[k], *Co = P.Cout[k];
  const fq *C = P.C[k];
  const size_t half = 2 * quarter;
  fq acc[3] = {fq_zero(), fq_zero(), fq_zero()};
  for (size_t i = (size_t)bx * CB + threadIdx.x; i < quarter; i += (size_t)gx * CB) {
    fq a0 = cbind(fq_load(A + i), fq_load(A + i + half), r), a1 = cbind(fq_load(A + i + quarter), fq_load(A + i + quarter + half), r);
    fq b0 = cbind(fq_load(B + i), fq_load(B + i + half), r), b1 = cbind(fq_load(B + i + quarter), fq_load(B + i + quarter + half), r);
    fq c0 = cbind(fq_load(C + i), fq_load(C + i + half), r), c1 = cbind(fq_load(C + i + quarter), fq_load(C + i + quarter + half), r);
    fq_store(A + i, a0);
    fq_store(A + i + quarter, a1);
    fq_store(B + i, b0);
    fq_store(B + i + quarter, b1);
    if (Co) {
      fq_store(Co + i, c0);
      fq_store(Co + i + quarter, c1);
    }
    fq a2, a3, b2, b3, c2, c3;
    cline23(a0, a1, a2, a3);
    cline23(b0, b1, b2, b3);
    cline23(c0, c1, c2, c3);
    acc[0] = fq_add(acc[0], fq_mul(fq_mul(a0, b0), c0));
    acc[1] = fq_add(acc[1], fq_mul(fq_mul(a2, b2), c2));
    acc[2] = fq_add(acc[2], fq_mul(fq_mul(a3, b3), c3));
  }
  block_sum<3>(acc, sm);
  if (threadIdx.x == 0) {
    fq ck = K.c[k];
#pragma unroll
    for (int t = 0; t < 3; t++) acc[t] = fq_mul(ck, acc[t]);
  }
  finish_block<3>(fa, acc, sm);
}

// the same on small tables (k_cubic_eval_split's layout: four lanes per item, lanes 0..2 on t = 0, 2, 3). Each of
// the three lanes binds all six scalars itself -- the products are independent and the round is a dependent
// chain, not throughput -- and lane 0 / 1 / 2 stores the A / B / C pair after the warp has finished reading.
__global__ void __launch_bounds__(128)
k_cubic_bind_eval_split(const __grid_constant__ CubicBindPtrs P, const __grid_constant__ CubicCoeffs K, int gx, size_t quarter, fq r,
                        FinishArgs fa) {
  __shared__ fq sm[4 * 4];
  const unsigned int lane = threadIdx.x & 31, l = lane & 3, warp = threadIdx.x >> 5;
  const int k = blockIdx.x / gx, bx = blockIdx.x % gx;
  fq *A = P.A[k], *B = P.B[k], *Co = P.Cout[k];
  const fq *C = P.C[k];
  const size_t half = 2 * quarter;
  const size_t i = (size_t)bx * CUBIC_SPLIT_ITEMS + (threadIdx.x >> 2);
  const bool act = i < quarter && l < 3;
  fq acc = fq_zero(), s0 = fq_zero(), s1 = fq_zero();
  if (act) {
    fq a0 = cbind(fq_load(A + i), fq_load(A + i + half), r), a1 = cbind(fq_load(A + i + quarter), fq_load(A + i + quarter + half), r);
    fq b0 = cbind(fq_load(B + i), fq_load(B + i + half), r), b1 = cbind(fq_load(B + i + quarter), fq_load(B + i + quarter + half), r);
    fq c0 = cbind(fq_load(C + i), fq_load(C + i + half), r), c1 = cbind(fq_load(C + i + quarter), fq_load(C + i + quarter + half), r);
    acc = fq_mul(fq_mul(split_point(a0, a1, l), split_point(b0, b1, l)), split_point(c0, c1, l));
#pragma unroll
    for (int w = 0; w < 8; w++) {
      s0.v[w] = l == 0 ? a0.v[w] : (l == 1 ? b0.v[w] : c0.v[w]);
      s1.v[w] = l == 0 ? a1.v[w] : (l == 1 ? b1.v[w] : c1.v[w]);
    }
  }
  __syncwarp();  // every lane of the item has read the unbound entries before one of them overwrites them
  if (act) {
    fq *T = l == 0 ? A : (l == 1 ? B : Co);
    if (T) {
      fq_store(T + i, s0);
      fq_store(T + i + quarter, s1);
    }
  }
  // lanes 4 j + l hold point l of item j
  acc = fq_add(acc, fq_shfl_down(acc, 4));
  acc = fq_add(acc, fq_shfl_down(acc, 8));
  acc = fq_add(acc, fq_shfl_down(acc, 16));
  if (lane < 3) sm[warp * 4 + lane] = acc;
  __syncthreads();
  fq mine[3] = {fq_zero(), fq_zero(), fq_zero()};
  if (warp == 0) {
    fq t = (lane < 16 && l < 3) ? sm[(lane >> 2) * 4 + l] : fq_zero();
    t = fq_add(t, fq_shfl_down(t, 4));
    t = fq_add(t, fq_shfl_down(t, 8));
    t = fq_mul(K.c[k], t);  // lanes 0 .. 2: the three sums of this block, scaled by the triple's coefficient
    mine[0] = t;
    mine[1] = fq_shfl(t, 1);
    mine[2] = fq_shfl(t, 2);
  }
  finish_block_lanes3(fa, mine, sm);
}

struct BindPtrs {
  fq *T[64];
};

// blockIdx.y = table; T[i] += r * (T[i + half] - T[i])
__global__ void k_multi_bind_top(BindPtrs P, size_t half, fq r) {
  fq *__restrict__ T = P.T[blockIdx.y];
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < half;
       i += (size_t)gridDim.x * blockDim.x) {
    fq lo = fq_load(T + i), hi = fq_load(T + i + half);
    fq_store(T + i, fq_add(lo, fq_mul(r, fq_sub(hi, lo))));
  }
}

// one ProductCircuit layer: out[i] = in[i] * in[i + n]  (compute_layer, product_tree.rs:18-34)
__global__ void k_prod_layer(const fq *__restrict__ in, fq *__restrict__ out, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x)
    fq_store(out + i, fq_mul(fq_load_stream(in + i), fq_load_stream(in + i + n)));
}

// the last few layers in one block
__global__ void k_prod_tail(fq *__restrict__ buf, size_t n /* length of the first layer handled */) {
  // layer of length n lives at buf; the next one right behind it, and so on
  fq *in = buf;
  while (n > 1) {
    fq *out = in + n;
    size_t h = n / 2;
    for (size_t i = threadIdx.x; i < h; i += blockDim.x) out[i] = fq_mul(in[i], in[i + h]);
    __syncthreads();
    in = out;
    n = h;
  }
}

__device__ __forceinline__ fq fq_from_u64_dev(unsigned long long v) {
  // Scalar::from(u64) = [v,0,0,0] * R2  (ristretto255.rs:212-216)
  fq R2, raw = fq_zero();
  R2.v[0] = 0x449c0f01u; R2.v[1] = 0xa40611e3u; R2.v[2] = 0x68859347u; R2.v[3] = 0xd00e1ba7u;
  R2.v[4] = 0x17f5be65u; R2.v[5] = 0xceec73d2u; R2.v[6] = 0x7c309a3du; R2.v[7] = 0x0399411bu;
  raw.v[0] = (unsigned int)v;
  raw.v[1] = (unsigned int)(v >> 32);
  return fq_mul(R2, raw);
}

// hash_func(addr, val, ts) - tau = ts*gamma^2 + val*gamma + addr - tau
__global__ void k_hash_layer(const unsigned long long *__restrict__ addr, const fq *__restrict__ val,
                             const unsigned long long *__restrict__ ts, size_t n, fq gamma, fq gamma2,
                             fq tau, int ts_plus_one, fq *__restrict__ out) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x) {
    fq a = fq_from_u64_dev(addr ? addr[i] : (unsigned long long)i);
    fq h = fq_add(fq_mul(fq_load(val + i), gamma), a);
    if (ts) {
      fq t = fq_from_u64_dev(ts[i]);
      if (ts_plus_one) t = fq_add(t, fq_one());
      h = fq_add(h, fq_mul(t, gamma2));
    } else if (ts_plus_one) {
      h = fq_add(h, gamma2);
    }
    fq_store(out + i, fq_sub(h, tau));
  }
}

__global__ void k_deref(const unsigned long long *__restrict__ addr, size_t n, const fq *__restrict__ mem,
                        fq *__restrict__ out) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x)
    fq_store(out + i, fq_load(mem + addr[i]));
}

}  // namespace spg

using namespace spg;

struct spg_prodtree {
  spg_ctx *ctx = nullptr;
  fq *buf = nullptr;  // layer k (length n >> k) at offset n*2 - (n*2 >> k)
  size_t n = 0;
  size_t num_layers = 0;
  std::vector<spg_vec *> left, right;
};

struct spg_cubic {
  spg_ctx *ctx = nullptr;
  size_t npar = 0, nseq = 0;
  std::vector<spg_vec *> A_par, B_par, A_seq, B_seq, C_seq;
  spg_vec *C_par = nullptr;
  std::vector<hfq> coeffs;
  size_t len = 0;
  bool evaluated = false;
  // fused bind + evaluation (k_cubic_bind_eval_*): the shared C table alternates between C_par's own buffer
  // and C_alt; the evaluations of the round that follows a fused bind are cached here
  fq *C_alt = nullptr;
  bool c_in_alt = false, have_cached = false;
  spg_fq cached[3];
  const fq *c_cur() const { return c_in_alt ? C_alt : C_par->d; }
  fq *c_cur() { return c_in_alt ? C_alt : C_par->d; }
};

extern "C" {

int spg_prodtree_build(spg_ctx *ctx, const spg_vec *leaves, spg_prodtree **out) {
  spg::DeviceGuard _dev(spg::ctx_of(ctx));
  SPG_CHECK(ctx && leaves && out, "spg_prodtree_build: null argument");
  size_t n = leaves->n;
  SPG_CHECK(is_pow2(n) && n >= 2, "spg_prodtree_build: length %zu must be a power of two >= 2", n);
  spg_prodtree *t = new (std::nothrow) spg_prodtree();
  if (!t) return SPG_ENOMEM;
  t->ctx = ctx;
  t->n = n;
  t->num_layers = log2u(n);
  cudaError_t e = dev_alloc(ctx, &t->buf, 2 * n * sizeof(fq));
  if (e != cudaSuccess) {
    delete t;
    return cuda_fail(e, "cudaMalloc(prodtree)", __FILE__, __LINE__);
  }
  SPG_CUDA(cudaMemcpyAsync(t->buf, leaves->d, n * sizeof(fq), cudaMemcpyDeviceToDevice, ctx->stream));
  fq *in = t->buf;
  size_t len = n;
  while (len > 1024) {
    size_t h = len / 2;
    SPG_LAUNCH(ctx, k_prod_layer, grid_for(ctx, h, 256), 256, 0, in, in + len, h);
    in += len;
    len = h;
  }
  if (len > 1) SPG_LAUNCH(ctx, k_prod_tail, 1, 256, 0, in, len);
  // views: layer k's vector V_k (length n >> k); left = first half, right = second half
  size_t off = 0;
  for (size_t k = 0; k < t->num_layers; k++) {
    size_t L = n >> k;
    spg_vec *l = nullptr, *r = nullptr;
    SPG_TRY(spg_vec_wrap(ctx, t->buf + off, L / 2, &l));
    SPG_TRY(spg_vec_wrap(ctx, t->buf + off + L / 2, L / 2, &r));
    t->left.push_back(l);
    t->right.push_back(r);
    off += L;
  }
  *out = t;
  return SPG_OK;
}

size_t spg_prodtree_num_layers(const spg_prodtree *t) { return t ? t->num_layers : 0; }

int spg_prodtree_layer(spg_prodtree *t, size_t layer, spg_vec **left, spg_vec **right) {
  spg::DeviceGuard _dev(spg::ctx_of(t));
  SPG_CHECK(t && left && right, "spg_prodtree_layer: null argument");
  SPG_CHECK(layer < t->num_layers, "spg_prodtree_layer: layer %zu out of range (%zu)", layer, t->num_layers);
  *left = t->left[layer];
  *right = t->right[layer];
  return SPG_OK;
}

int spg_prodtree_evaluate(spg_ctx *ctx, spg_prodtree *t, spg_fq *out) {
  spg::DeviceGuard _dev(spg::ctx_of(ctx));
  SPG_CHECK(ctx && t && out, "spg_prodtree_evaluate: null argument");
  // left_vec[last][0] * right_vec[last][0]; the last layer (two scalars) sits at 2n - 4
  fq h[2];
  SPG_CUDA(cudaMemcpyAsync(h, t->buf + 2 * t->n - 4, 2 * sizeof(fq), cudaMemcpyDeviceToHost, ctx->stream));
  SPG_CUDA(cudaStreamSynchronize(ctx->stream));
  hfq a, b;
  memcpy(&a, &h[0], 32);
  memcpy(&b, &h[1], 32);
  *out = hfq_to(hfq_mul(a, b));
  return SPG_OK;
}

void spg_prodtree_destroy(spg_prodtree *t) {
  spg::DeviceGuard _dev(spg::ctx_of(t));
  if (!t) return;
  for (spg_vec *v : t->left) spg_vec_free(v);
  for (spg_vec *v : t->right) spg_vec_free(v);
  if (t->buf) dev_free(t->ctx, t->buf);
  delete t;
}

int spg_cubic_create(spg_ctx *ctx, size_t npar, spg_vec *const *A_par, spg_vec *const *B_par,
                     spg_vec *C_par, size_t nseq, spg_vec *const *A_seq, spg_vec *const *B_seq,
                     spg_vec *const *C_seq, const spg_fq *coeffs, spg_cubic **out) {
  spg::DeviceGuard _dev(spg::ctx_of(ctx));
  SPG_CHECK(ctx && out && coeffs, "spg_cubic_create: null argument");
  SPG_CHECK(npar + nseq >= 1 && npar + nseq <= 24, "spg_cubic_create: %zu triples (1..24 supported)", npar + nseq);
  SPG_CHECK(npar == 0 || (A_par && B_par && C_par), "spg_cubic_create: null parallel tables");
  SPG_CHECK(nseq == 0 || (A_seq && B_seq && C_seq), "spg_cubic_create: null sequential tables");
  spg_cubic *s = new (std::nothrow) spg_cubic();
  if (!s) return SPG_ENOMEM;
  s->ctx = ctx;
  s->npar = npar;
  s->nseq = nseq;
  s->C_par = C_par;
  size_t len = npar ? A_par[0]->n : A_seq[0]->n;
  bool ok = is_pow2(len);
  for (size_t i = 0; i < npar; i++) {
    s->A_par.push_back(A_par[i]);
    s->B_par.push_back(B_par[i]);
    ok = ok && A_par[i]->n == len && B_par[i]->n == len;
  }
  ok = ok && (npar == 0 || C_par->n == len);
  for (size_t i = 0; i < nseq; i++) {
    s->A_seq.push_back(A_seq[i]);
    s->B_seq.push_back(B_seq[i]);
    s->C_seq.push_back(C_seq[i]);
    ok = ok && A_seq[i]->n == len && B_seq[i]->n == len && C_seq[i]->n == len;
  }
  if (!ok) {
    delete s;
    set_error("spg_cubic_create: all tables must share one power-of-two length");
    return SPG_EINVAL;
  }
  s->len = len;
  for (size_t i = 0; i < npar + nseq; i++) s->coeffs.push_back(hfq_from(coeffs[i]));
  *out = s;
  return SPG_OK;
}

int spg_cubic_round_eval(spg_cubic *s, spg_fq e[3]) {
  spg::DeviceGuard _dev(spg::ctx_of(s));
  SPG_CHECK(s && e, "spg_cubic_round_eval: null argument");
  if (s->len < 2 || s->evaluated) {
    set_error(s->evaluated ? "spg_cubic_round_eval: round already evaluated" : "spg_cubic_round_eval: all rounds are done");
    return SPG_ESTATE;
  }
  if (s->have_cached) {  // evaluated by the fused kernel of the previous bind
    memcpy(e, s->cached, sizeof s->cached);
    s->have_cached = false;
    s->evaluated = true;
    return SPG_OK;
  }
  spg_ctx *ctx = s->ctx;
  size_t nt = s->npar + s->nseq, half = s->len / 2;
  CubicPtrs P;
  for (size_t k = 0; k < s->npar; k++) {
    P.A[k] = s->A_par[k]->d;
    P.B[k] = s->B_par[k]->d;
    P.C[k] = s->c_cur();
  }
  for (size_t k = 0; k < s->nseq; k++) {
    P.A[s->npar + k] = s->A_seq[k]->d;
    P.B[s->npar + k] = s->B_seq[k]->d;
    P.C[s->npar + k] = s->C_seq[k]->d;
  }
  int gx = grid_for(ctx, half, CB, 2);
  if (gx > 1024) gx = 1024;
  if ((size_t)gx * nt > 4096) gx = (int)(4096 / nt);  // keep the in-kernel final reduction (finish_args)
  const size_t gx_split = (half + CUBIC_SPLIT_ITEMS - 1) / CUBIC_SPLIT_ITEMS;
  const bool split = gx_split * nt <= 512;  // small tables: the lane-split kernel (latency, not throughput)
  if (split) gx = (int)gx_split;
  size_t nblocks = (size_t)gx * nt;
  SPG_TRY(ensure_partials(ctx, nblocks * 3));
  CubicCoeffs K;
  for (size_t k = 0; k < nt; k++) {
    spg_fq c = hfq_to(s->coeffs[k]);
    memcpy(&K.c[k], &c, sizeof(fq));
  }
  ctx->next_units = 192.0 * (double)half * (double)nt;
  FinishArgs fa = finish_args(ctx, nblocks);
  if (split)
    SPG_LAUNCH(ctx, k_cubic_eval_split, (unsigned)nblocks, 128, 0, P, K, gx, half, fa);
  else
    SPG_LAUNCH(ctx, k_cubic_eval_rlc, (unsigned)nblocks, CB, 0, P, K, gx, half, fa);
  SPG_TRY(finish_result(ctx, fa, nblocks, 3, e));
  s->evaluated = true;
  return SPG_OK;
}

int spg_cubic_round_bind(spg_cubic *s, const spg_fq *r) {
  spg::DeviceGuard _dev(spg::ctx_of(s));
  SPG_CHECK(s && r, "spg_cubic_round_bind: null argument");
  if (!s->evaluated) {
    set_error("spg_cubic_round_bind: round has not been evaluated");
    return SPG_ESTATE;
  }
  spg_ctx *ctx = s->ctx;
  size_t half = s->len / 2;
  fq rr;
  memcpy(&rr, r, 32);
  std::vector<spg_vec *> all;
  for (auto v : s->A_par) all.push_back(v);
  for (auto v : s->B_par) all.push_back(v);
  if (s->npar) all.push_back(s->C_par);
  for (auto v : s->A_seq) all.push_back(v);
  for (auto v : s->B_seq) all.push_back(v);
  for (auto v : s->C_seq) all.push_back(v);
  static const bool fuse = [] {
    const char *e = getenv("SPG_CUBIC_FUSE");  // SPG_CUBIC_FUSE=0: separate bind and evaluation launches (A/B switch)
    return !(e && *e == '0');
  }();
  if (fuse && s->len >= 4) {
    // another round follows: bind and evaluate it in one launch
    const size_t nt = s->npar + s->nseq, quarter = s->len / 4;
    if (s->npar && !s->C_alt) SPG_CUDA(dev_alloc(ctx, &s->C_alt, (s->len / 2) * sizeof(fq)));
    CubicBindPtrs P;
    memset(&P, 0, sizeof P);
    fq *c_next = s->npar ? (s->c_in_alt ? s->C_par->d : s->C_alt) : nullptr;
    for (size_t k = 0; k < s->npar; k++) {
      P.A[k] = s->A_par[k]->d;
      P.B[k] = s->B_par[k]->d;
      P.C[k] = s->c_cur();
      P.Cout[k] = k == 0 ? c_next : nullptr;
    }
    for (size_t k = 0; k < s->nseq; k++) {
      P.A[s->npar + k] = s->A_seq[k]->d;
      P.B[s->npar + k] = s->B_seq[k]->d;
      P.C[s->npar + k] = s->C_seq[k]->d;
      P.Cout[s->npar + k] = s->C_seq[k]->d;
    }
    int gx = grid_for(ctx, quarter, CB, 2);
    if (gx > 1024) gx = 1024;
    if ((size_t)gx * nt > 4096) gx = (int)(4096 / nt);  // keep the in-kernel final reduction (finish_args)
    const size_t gx_split = (quarter + CUBIC_SPLIT_ITEMS - 1) / CUBIC_SPLIT_ITEMS;
    const bool split = gx_split * nt <= 512;
    if (split) gx = (int)gx_split;
    size_t nblocks = (size_t)gx * nt;
    SPG_TRY(ensure_partials(ctx, nblocks * 3));
    CubicCoeffs K;
    for (size_t k = 0; k < nt; k++) {
      spg_fq c = hfq_to(s->coeffs[k]);
      memcpy(&K.c[k], &c, sizeof(fq));
    }
    ctx->next_units = 288.0 * (double)quarter * (double)nt;  // per item and triple: 12 scalars read (C shared), 6 written
    FinishArgs fa = finish_args(ctx, nblocks);
    if (split)
      SPG_LAUNCH(ctx, k_cubic_bind_eval_split, (unsigned)nblocks, 128, 0, P, K, gx, quarter, rr, fa);
    else
      SPG_LAUNCH(ctx, k_cubic_bind_eval_rlc, (unsigned)nblocks, CB, 0, P, K, gx, quarter, rr, fa);
    if (s->npar) s->c_in_alt = !s->c_in_alt;
    for (auto v : all) v->n = half;
    s->len = half;
    s->evaluated = false;
    SPG_TRY(finish_result(ctx, fa, nblocks, 3, s->cached));
    s->have_cached = true;
    return SPG_OK;
  }
  for (size_t base = 0; base < all.size(); base += 64) {
    size_t cnt = all.size() - base < 64 ? all.size() - base : 64;
    BindPtrs P;
    for (size_t i = 0; i < cnt; i++) P.T[i] = (all[base + i] == s->C_par && s->npar) ? s->c_cur() : all[base + i]->d;
    int gx = grid_for(ctx, half, 128, 2);
    dim3 grid(gx, (unsigned)cnt);
    ctx->next_units = 96.0 * (double)half * (double)cnt;
    SPG_LAUNCH(ctx, k_multi_bind_top, grid, 128, 0, P, half, rr);
  }
  for (auto v : all) v->n = half;
  s->len = half;
  s->evaluated = false;
  return SPG_OK;
}

int spg_cubic_final(spg_cubic *s, spg_fq *claims) {
  spg::DeviceGuard _dev(spg::ctx_of(s));
  SPG_CHECK(s && claims, "spg_cubic_final: null argument");
  if (s->len != 1) {
    set_error("spg_cubic_final: tables still have %zu entries", s->len);
    return SPG_ESTATE;
  }
  spg_ctx *ctx = s->ctx;
  std::vector<spg_vec *> all;
  for (auto v : s->A_par) all.push_back(v);
  for (auto v : s->B_par) all.push_back(v);
  if (s->npar) all.push_back(s->C_par);
  for (auto v : s->A_seq) all.push_back(v);
  for (auto v : s->B_seq) all.push_back(v);
  for (auto v : s->C_seq) all.push_back(v);
  if (s->npar && s->c_in_alt) {
    // the caller's vector holds the bound scalar too, like every other table
    SPG_CUDA(cudaMemcpyAsync(s->C_par->d, s->C_alt, sizeof(fq), cudaMemcpyDeviceToDevice, ctx->stream));
    s->c_in_alt = false;
  }
  std::vector<const fq *> heads;
  for (auto v : all) heads.push_back(v->d);
  return gather_heads(ctx, heads.data(), heads.size(), claims);
}

void spg_cubic_destroy(spg_cubic *s) {
  spg::DeviceGuard _dev(spg::ctx_of(s));
  if (!s) return;
  if (s->C_alt) dev_free(s->ctx, s->C_alt);
  delete s;
}

int spg_hash_layer(spg_ctx *ctx, const uint64_t *addr, const spg_vec *val, const uint64_t *ts, size_t n,
                   const spg_fq *gamma, const spg_fq *tau, int ts_plus_one, spg_vec **out) {
  spg::DeviceGuard _dev(spg::ctx_of(ctx));
  SPG_CHECK(ctx && val && gamma && tau && out, "spg_hash_layer: null argument");
  SPG_CHECK(val->n >= n, "spg_hash_layer: val has %zu entries, need %zu", val->n, n);
  VecOut o;
  SPG_TRY(vec_new(ctx, n, &o.v));
  DevTmp t_addr(ctx), t_ts(ctx);
  if (addr) {
    SPG_CUDA(t_addr.alloc(n * 8));
    SPG_CUDA(cudaMemcpyAsync(t_addr.p, addr, n * 8, cudaMemcpyHostToDevice, ctx->stream));
  }
  if (ts) {
    SPG_CUDA(t_ts.alloc(n * 8));
    SPG_CUDA(cudaMemcpyAsync(t_ts.p, ts, n * 8, cudaMemcpyHostToDevice, ctx->stream));
  }
  hfq g = hfq_from(*gamma);
  hfq g2 = hfq_mul(g, g);
  fq fg, fg2, ft;
  memcpy(&fg, &g, 32);
  memcpy(&fg2, &g2, 32);
  memcpy(&ft, tau, 32);
  ctx->next_units = (double)n * (32.0 + 32.0 + (addr ? 8 : 0) + (ts ? 8 : 0));
  SPG_LAUNCH(ctx, k_hash_layer, grid_for(ctx, n, 256), 256, 0, t_addr.as<unsigned long long>(), val->d,
             t_ts.as<unsigned long long>(), n, fg, fg2, ft, ts_plus_one, o.v->d);
  SPG_CUDA(cudaStreamSynchronize(ctx->stream));
  *out = o.release();
  return SPG_OK;
}

int spg_deref(spg_ctx *ctx, const uint64_t *addr, size_t n, const spg_vec *mem, spg_vec **out) {
  spg::DeviceGuard _dev(spg::ctx_of(ctx));
  SPG_CHECK(ctx && addr && mem && out, "spg_deref: null argument");
  for (size_t i = 0; i < n; i++)
    SPG_CHECK(addr[i] < mem->n, "spg_deref: address %llu at %zu exceeds %zu memory cells",
              (unsigned long long)addr[i], i, mem->n);
  VecOut o;
  SPG_TRY(vec_new(ctx, n, &o.v));
  DevTmp t_addr(ctx);
  SPG_CUDA(t_addr.alloc((n ? n : 1) * 8));
  SPG_CUDA(cudaMemcpyAsync(t_addr.p, addr, n * 8, cudaMemcpyHostToDevice, ctx->stream));
  SPG_LAUNCH(ctx, k_deref, grid_for(ctx, n, 256), 256, 0, t_addr.as<unsigned long long>(), n, mem->d, o.v->d);
  SPG_CUDA(cudaStreamSynchronize(ctx->stream));
  *out = o.release();
  return SPG_OK;
}

}  // extern "C"
