// Pedersen vector commitments on the device.
//   Commitments::commit            src/commitments.rs:69-92   (sum_j s_j G_j + blind h)
//   DensePolynomial::commit_inner  src/dense_mlpoly.rs:199-239 (one commitment per matrix row,
//                                                              all rows share the bases)
//   GroupElement::vartime_multiscalar_mul / compress           src/group.rs:98-117
// dalek's Straus/Pippenger choice is irrelevant for parity: the compressed ristretto
// encoding of the sum is canonical (RFC 9496).
//
// Algorithm. All rows of a polynomial commitment use the same R bases, and there are
// thousands of rows, so the bases get fixed-base window tables once:
//   T[j][w][d-1] = d * 2^(C*w) * G_j      d in 1..2^C-1, w in 0..ceil(253/C)-1
// stored in "cached" form (Y+X, Y-X, Z, 2dT). A row commitment is then a pure sum of
// table entries -- no doublings, no buckets, no atomics, zero digits are skipped:
//   C_i = sum_j sum_w T[j][w][digit_w(s_ij)]
// One thread owns (row i, chunk of bases); a warp is 32 consecutive rows of the same
// chunk, so all its table reads fall in the same few (j, w) slices (L1/L2 resident).
#include "common.cuh"
#include "ed25519.cuh"

namespace spg {

template <int C>
struct WinCfg {
  static constexpr int WINS = (253 + C - 1) / C;
  static constexpr int ENTRIES = (1 << C) - 1;
};

__global__ void k_decompress(const uint8_t *__restrict__ in, size_t n, ge *__restrict__ out,
                             int *__restrict__ bad) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint8_t b[32];
  for (int k = 0; k < 32; k++) b[k] = in[32 * i + k];
  ge p;
  if (!ristretto_decompress(b, &p)) {
    atomicExch(bad, (int)i + 1);
    p = ge_identity();
  }
  out[i] = p;
}

// MultiCommitGens::new's per-point step (src/commitments.rs:23-31): 64 uniform bytes ->
// RistrettoPoint::from_uniform_bytes (two Elligator maps and an addition)
__global__ void k_from_uniform(const uint8_t *__restrict__ in, size_t n, ge *__restrict__ out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint8_t b[64];
  for (int k = 0; k < 64; k++) b[k] = in[64 * i + k];
  out[i] = ristretto_from_uniform_bytes(b);
}

// thread (j, w): entries d = 1 .. 2^C - 1 of window w of base j
template <int C>
__global__ void k_build_table(const ge *__restrict__ bases, size_t nbases, ge_cached *__restrict__ table) {
  constexpr int WINS = WinCfg<C>::WINS, ENT = WinCfg<C>::ENTRIES;
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nbases * WINS) return;
  size_t j = t / WINS;
  int w = (int)(t % WINS);
  ge p = bases[j];
  for (int k = 0; k < C * w; k++) p = ge_double(p);
  ge_cached pc = ge_to_cached(p);
  ge_cached *dst = table + (j * WINS + w) * ENT;
  dst[0] = pc;
  ge m = p;
  for (int d = 2; d <= ENT; d++) {
    m = ge_add(m, pc);
    dst[d - 1] = ge_to_cached(m);
  }
}

template <int C>
__device__ __forceinline__ void accumulate_scalar(ge &acc, const fq &s_mont, const ge_cached *__restrict__ tj) {
  constexpr int WINS = WinCfg<C>::WINS, ENT = WinCfg<C>::ENTRIES;
  // leave Montgomery form: the group multiplies by the integer value (src/scalar/mod.rs:32-36)
  fq s = fq_from_mont(s_mont);
#pragma unroll 1
  for (int w = 0; w < WINS; w++) {
    int bit = C * w;
    uint32_t d = (s.v[bit >> 5] >> (bit & 31));
    if ((bit & 31) + C > 32 && (bit >> 5) < 7) d |= s.v[(bit >> 5) + 1] << (32 - (bit & 31));
    d &= (1u << C) - 1;
    if (d) acc = ge_add(acc, tj[(size_t)w * ENT + (d - 1)]);
  }
}

// partial[(i * nchunks + k)] = sum over bases j in chunk k of s[i][j] * G_j
template <int C>
__global__ void __launch_bounds__(128)
k_msm_partial(const fq *__restrict__ scalars, size_t L, size_t R, size_t row_stride,
              const ge_cached *__restrict__ table, size_t chunk, size_t nchunks, ge *__restrict__ partial) {
  constexpr int WINS = WinCfg<C>::WINS, ENT = WinCfg<C>::ENTRIES;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t k = blockIdx.y;
  if (i >= L) return;
  ge acc = ge_identity();
  size_t j0 = k * chunk, j1 = j0 + chunk < R ? j0 + chunk : R;
  for (size_t j = j0; j < j1; j++) {
    fq s = fq_load(scalars + i * row_stride + j);
    if (fq_is_zero(s)) continue;
    accumulate_scalar<C>(acc, s, table + j * WINS * ENT);
  }
  partial[i * nchunks + k] = acc;
}

// few rows, many bases (the L / R vectors of a bullet reduction round, Cx of an opening):
// WIDE_SPLIT threads per base, each adding a quarter of the base's windows (the additions of one
// scalar are a dependent chain: 8 instead of 32 in a row), a block sums its 128 points through
// shared memory. grid (ceil(R / WIDE_BASES), L); partial[i * gridDim.x + blockIdx.x]
constexpr int WIDE_SPLIT = 4, WIDE_BASES = 128 / WIDE_SPLIT;
template <int C>
__global__ void __launch_bounds__(128)
k_msm_wide(const fq *__restrict__ scalars, size_t R, size_t row_stride, const ge_cached *__restrict__ table,
           ge *__restrict__ partial) {
  constexpr int WINS = WinCfg<C>::WINS, ENT = WinCfg<C>::ENTRIES;
  constexpr int PER = (WINS + WIDE_SPLIT - 1) / WIDE_SPLIT;
  __shared__ ge sm[64];
  size_t i = blockIdx.y;
  size_t j = (size_t)blockIdx.x * WIDE_BASES + threadIdx.x / WIDE_SPLIT;
  const int part = threadIdx.x % WIDE_SPLIT;
  ge acc = ge_identity();
  if (j < R) {
    fq sm_ = fq_load(scalars + i * row_stride + j);
    if (!fq_is_zero(sm_)) {
      // leave Montgomery form: the group multiplies by the integer value (src/scalar/mod.rs:32-36)
      fq s = fq_from_mont(sm_);
      const ge_cached *__restrict__ tj = table + j * WINS * ENT;
#pragma unroll 1
      for (int w = part * PER; w < (part + 1) * PER && w < WINS; w++) {
        int bit = C * w;
        uint32_t d = (s.v[bit >> 5] >> (bit & 31));
        if ((bit & 31) + C > 32 && (bit >> 5) < 7) d |= s.v[(bit >> 5) + 1] << (32 - (bit & 31));
        d &= (1u << C) - 1;
        if (d) acc = ge_add(acc, tj[(size_t)w * ENT + (d - 1)]);
      }
    }
  }
  for (int half = 64; half >= 1; half >>= 1) {
    if ((int)threadIdx.x >= half && (int)threadIdx.x < 2 * half) sm[threadIdx.x - half] = acc;
    __syncthreads();
    if ((int)threadIdx.x < half) acc = ge_add(acc, ge_to_cached(sm[threadIdx.x]));
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[i * gridDim.x + blockIdx.x] = acc;
}

// out[i] = compress(sum_k partial[i][k] + blind[i] * h); h's table sits in slot `hslot`
template <int C>
__global__ void k_msm_finish(const ge *__restrict__ partial, size_t L, size_t nchunks,
                             const fq *__restrict__ blinds, const ge_cached *__restrict__ table,
                             size_t hslot, uint8_t *__restrict__ out) {
  constexpr int WINS = WinCfg<C>::WINS, ENT = WinCfg<C>::ENTRIES;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= L) return;
  ge acc = partial[i * nchunks];
  for (size_t k = 1; k < nchunks; k++) acc = ge_add(acc, ge_to_cached(partial[i * nchunks + k]));
  if (blinds) {
    fq b = blinds[i];
    if (!fq_is_zero(b)) accumulate_scalar<C>(acc, b, table + hslot * WINS * ENT);
  }
  uint8_t enc[32];
  ristretto_compress(acc, enc);
  for (int k = 0; k < 32; k++) out[32 * i + k] = enc[k];
}

// Few rows (the L / R of a bullet-reduction round, Cx, the folded generator): one block per row
// adds the row's per-block partial sums AND the blind's window points as a tree through shared
// memory -- ~8 dependent additions instead of the (nblk + 32) sequential ones a single thread of
// k_msm_finish would do, which is what a round of the opening proof used to wait for.
template <int C>
__global__ void __launch_bounds__(128)
k_msm_finish_tree(const ge *__restrict__ partial, size_t nblk, const fq *__restrict__ blinds,
                  const ge_cached *__restrict__ table, size_t hslot, uint8_t *__restrict__ out) {
  constexpr int WINS = WinCfg<C>::WINS, ENT = WinCfg<C>::ENTRIES;
  __shared__ ge sm[64];
  const size_t i = blockIdx.x;
  ge acc = ge_identity();
  for (size_t k = threadIdx.x; k < nblk; k += 128) acc = ge_add(acc, ge_to_cached(partial[i * nblk + k]));
  if (blinds && (int)threadIdx.x < WINS) {
    // window w = threadIdx.x of blind * h
    fq b = fq_from_mont(blinds[i]);
    int bit = C * (int)threadIdx.x;
    uint32_t d = (b.v[bit >> 5] >> (bit & 31));
    if ((bit & 31) + C > 32 && (bit >> 5) < 7) d |= b.v[(bit >> 5) + 1] << (32 - (bit & 31));
    d &= (1u << C) - 1;
    if (d) acc = ge_add(acc, table[(hslot * WINS + threadIdx.x) * ENT + (d - 1)]);
  }
  for (int half = 64; half >= 1; half >>= 1) {
    if ((int)threadIdx.x >= half && (int)threadIdx.x < 2 * half) sm[threadIdx.x - half] = acc;
    __syncthreads();
    if ((int)threadIdx.x < half) acc = ge_add(acc, ge_to_cached(sm[threadIdx.x]));
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    uint8_t enc[32];
    ristretto_compress(acc, enc);
    for (int k = 0; k < 32; k++) out[32 * i + k] = enc[k];
  }
}

}  // namespace spg

using namespace spg;

struct spg_gens {
  spg_ctx *ctx = nullptr;
  size_t n = 0;          // number of G's; h is bases[n]
  ge *bases = nullptr;   // n + 1 points
  // window tables for bases [0, tab_R) and h (slot tab_R)
  ge_cached *table = nullptr;
  size_t tab_R = 0;
  int tab_C = 0;
};

namespace {

template <int C>
int build_table(spg_gens *g, size_t R) {
  spg_ctx *ctx = g->ctx;
  constexpr int WINS = WinCfg<C>::WINS, ENT = WinCfg<C>::ENTRIES;
  size_t slots = R + 1;
  ge_cached *t = nullptr;
  SPG_CUDA(cudaMalloc(&t, slots * WINS * ENT * sizeof(ge_cached)));
  size_t threads = R * WINS;
  SPG_LAUNCH(ctx, k_build_table<C>, (unsigned)((threads + 63) / 64), 64, 0, g->bases, R, t);
  // h goes into slot R
  SPG_LAUNCH(ctx, k_build_table<C>, (unsigned)((WINS + 63) / 64), 64, 0, g->bases + g->n, (size_t)1,
             t + R * WINS * ENT);
  SPG_CUDA(cudaStreamSynchronize(ctx->stream));
  if (g->table) cudaFree(g->table);
  g->table = t;
  g->tab_R = R;
  g->tab_C = C;
  return SPG_OK;
}

int ensure_table(spg_gens *g, size_t R) {
  if (g->table && g->tab_R >= R) return SPG_OK;
  // 8-bit windows cost (R+1) * 32 * 255 * 160 B = 1.3 MB per base; fall back to 4-bit
  // windows (154 KB per base) when that would exceed 24 GiB
  size_t bytes8 = (R + 1) * (size_t)WinCfg<8>::WINS * WinCfg<8>::ENTRIES * sizeof(ge_cached);
  if (bytes8 <= ((size_t)24 << 30)) return build_table<8>(g, R);
  return build_table<4>(g, R);
}

template <int C>
int run_msm(spg_gens *g, const fq *scalars, size_t L, size_t R, size_t row_stride, const fq *d_blinds,
            uint8_t *d_out) {
  spg_ctx *ctx = g->ctx;
  if (L <= 16 && R >= 256) {
    size_t nblk = (R + WIDE_BASES - 1) / WIDE_BASES;
    ge *partial = nullptr;
    SPG_CUDA(dev_alloc(ctx, &partial, L * nblk * sizeof(ge)));
    dim3 grid((unsigned)nblk, (unsigned)L);
    ctx->next_units = 32.0 * (double)L * (double)R;
    SPG_LAUNCH(ctx, k_msm_wide<C>, grid, 128, 0, scalars, R, row_stride, g->table, partial);
    static_assert(WinCfg<C>::WINS <= 128, "one thread per window of the blind");
    SPG_LAUNCH(ctx, k_msm_finish_tree<C>, (unsigned)L, 128, 0, partial, nblk, d_blinds, g->table, g->tab_R, d_out);
    SPG_CUDA(cudaStreamSynchronize(ctx->stream));
    dev_free(ctx, partial);
    return SPG_OK;
  }
  size_t chunk = 1;
  while (chunk < 64 && (L * R) / (chunk * 2) >= 131072) chunk *= 2;
  size_t nchunks = (R + chunk - 1) / chunk;
  if (nchunks == 0) nchunks = 1;
  ge *partial = nullptr;
  SPG_CUDA(dev_alloc(ctx, &partial, L * nchunks * sizeof(ge)));
  dim3 grid((unsigned)((L + 127) / 128), (unsigned)nchunks);
  ctx->next_units = 32.0 * (double)L * (double)R;
  SPG_LAUNCH(ctx, k_msm_partial<C>, grid, 128, 0, scalars, L, R, row_stride, g->table, chunk, nchunks, partial);
  SPG_LAUNCH(ctx, k_msm_finish<C>, (unsigned)((L + 63) / 64), 64, 0, partial, L, nchunks, d_blinds, g->table,
             g->tab_R, d_out);
  SPG_CUDA(cudaStreamSynchronize(ctx->stream));
  dev_free(ctx, partial);
  return SPG_OK;
}

int msm_rows(spg_gens *g, const fq *scalars, size_t L, size_t R, size_t row_stride, const fq *d_blinds,
             uint8_t *host_out) {
  SPG_CHECK(R <= g->n, "commit: %zu scalars per row but only %zu generators", R, g->n);
  SPG_TRY(ensure_table(g, R));
  uint8_t *d_out = nullptr;
  spg_ctx *ctx = g->ctx;
  // a handful of points: the finish kernel writes them straight into the context's mapped result
  // page (run_msm ends with a stream synchronise), no device buffer and no copy call
  const bool mapped = L * 32 <= 48 * sizeof(fq);
  if (mapped) d_out = reinterpret_cast<uint8_t *>(ctx->d_result);
  else SPG_CUDA(dev_alloc(ctx, &d_out, L * 32));
  int rc = g->tab_C == 8 ? run_msm<8>(g, scalars, L, R, row_stride, d_blinds, d_out)
                         : run_msm<4>(g, scalars, L, R, row_stride, d_blinds, d_out);
  if (rc == SPG_OK) {
    if (mapped) {
      memcpy(host_out, ctx->h_result, L * 32);
    } else {
      cudaError_t e = cudaMemcpy(host_out, d_out, L * 32, cudaMemcpyDeviceToHost);
      if (e != cudaSuccess) rc = cuda_fail(e, "commit download", __FILE__, __LINE__);
    }
  }
  if (!mapped) dev_free(ctx, d_out);
  return rc;
}

// ---------------------------------------------------------------- bullet reduction on unfolded bases
// BulletReductionProof::prove (src/nizk/bullet.rs:72-119) with the generator fold unrolled
// (DESIGN.md section 4): s[m] = prod over the rounds so far of (u_j or u_j^-1) by the top bits
// of m stays on the device; a round's L and R are MSMs over the ORIGINAL bases with scalars
//   L: [m mod nk >= nk/2] a[(m mod nk) - nk/2] s[m],   R: [m mod nk < nk/2] a[(m mod nk) + nk/2] s[m].
__global__ void k_bullet_rows(const fq *__restrict__ a, const fq *__restrict__ s, size_t n, size_t nk,
                              fq *__restrict__ rows) {
  size_t m = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= n) return;
  size_t nh = nk >> 1, j = m & (nk - 1);
  fq sm = fq_load(s + m);
  if (j >= nh) {
    fq_store(rows + m, fq_mul(fq_load(a + j - nh), sm));
    fq_store(rows + n + m, fq_zero());
  } else {
    fq_store(rows + m, fq_zero());
    fq_store(rows + n + m, fq_mul(fq_load(a + nh + j), sm));
  }
}
// the fold of one round applied to the scalars instead of the generators (bullet.rs:113-118)
__global__ void k_bullet_fold(fq *__restrict__ s, size_t n, size_t nk, fq u, fq u_inv) {
  size_t m = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= n) return;
  fq_store(s + m, fq_mul(fq_load(s + m), (m & (nk - 1)) >= (nk >> 1) ? u : u_inv));
}
__global__ void k_fill_one(fq *__restrict__ s, size_t n) {
  size_t m = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m < n) fq_store(s + m, fq_one());
}

}  // namespace

struct spg_bullet {
  spg_ctx *ctx = nullptr;
  spg_gens *gens = nullptr;
  size_t n = 0;
  fq *s = nullptr, *rows = nullptr;
  fq *in = nullptr;       // device: [blind_L, blind_R, a_0 .. a_{nk-1}]
  fq *h_in = nullptr;     // pinned staging of the same layout: one copy per round
};

extern "C" {

int spg_bullet_create(spg_ctx *ctx, const spg_gens *gens, size_t n, spg_bullet **out) {
  SPG_CHECK(ctx && gens && out, "spg_bullet_create: null argument");
  SPG_CHECK(n >= 2 && (n & (n - 1)) == 0 && n <= gens->n, "spg_bullet_create: n = %zu must be a power of two <= %zu generators", n,
            gens->n);
  spg_bullet *b = new (std::nothrow) spg_bullet();
  if (!b) return SPG_ENOMEM;
  b->ctx = ctx;
  b->gens = const_cast<spg_gens *>(gens);
  b->n = n;
  cudaError_t e = dev_alloc(ctx, &b->s, n * sizeof(fq));
  if (e == cudaSuccess) e = dev_alloc(ctx, &b->rows, 2 * n * sizeof(fq));
  if (e == cudaSuccess) e = dev_alloc(ctx, &b->in, (n + 2) * sizeof(fq));
  if (e == cudaSuccess) e = cudaHostAlloc(&b->h_in, (n + 2) * sizeof(fq), cudaHostAllocDefault);
  if (e != cudaSuccess) {
    spg_bullet_destroy(b);
    return cuda_fail(e, "spg_bullet_create", __FILE__, __LINE__);
  }
  SPG_LAUNCH(ctx, k_fill_one, (unsigned)((n + 255) / 256), 256, 0, b->s, n);
  *out = b;
  return SPG_OK;
}

int spg_bullet_lr(spg_bullet *b, size_t nk, const spg_fq *a, const spg_fq blinds[2], uint8_t out_LR[64]) {
  SPG_CHECK(b && a && blinds && out_LR, "spg_bullet_lr: null argument");
  SPG_CHECK(nk >= 2 && nk <= b->n && (nk & (nk - 1)) == 0, "spg_bullet_lr: bad round size %zu", nk);
  spg_ctx *ctx = b->ctx;
  // (the previous round ended with a stream synchronise, so the staging buffer is free again)
  memcpy(b->h_in, blinds, 2 * sizeof(fq));
  memcpy(b->h_in + 2, a, nk * sizeof(fq));
  SPG_CUDA(cudaMemcpyAsync(b->in, b->h_in, (nk + 2) * sizeof(fq), cudaMemcpyHostToDevice, ctx->stream));
  SPG_LAUNCH(ctx, k_bullet_rows, (unsigned)((b->n + 255) / 256), 256, 0, b->in + 2, b->s, b->n, nk, b->rows);
  return msm_rows(b->gens, b->rows, 2, b->n, b->n, b->in, out_LR);
}

int spg_bullet_fold(spg_bullet *b, size_t nk, const spg_fq *u, const spg_fq *u_inv) {
  SPG_CHECK(b && u && u_inv, "spg_bullet_fold: null argument");
  SPG_CHECK(nk >= 2 && nk <= b->n && (nk & (nk - 1)) == 0, "spg_bullet_fold: bad round size %zu", nk);
  fq fu, fi;
  memcpy(&fu, u, sizeof(fq));
  memcpy(&fi, u_inv, sizeof(fq));
  SPG_LAUNCH(b->ctx, k_bullet_fold, (unsigned)((b->n + 255) / 256), 256, 0, b->s, b->n, nk, fu, fi);
  return SPG_OK;
}

int spg_bullet_final(spg_bullet *b, uint8_t out_G[32]) {
  SPG_CHECK(b && out_G, "spg_bullet_final: null argument");
  return msm_rows(b->gens, b->s, 1, b->n, b->n, nullptr, out_G);
}

void spg_bullet_destroy(spg_bullet *b) {
  if (!b) return;
  dev_free(b->ctx, b->s);
  dev_free(b->ctx, b->rows);
  dev_free(b->ctx, b->in);
  if (b->h_in) cudaFreeHost(b->h_in);
  delete b;
}

int spg_gens_upload(spg_ctx *ctx, const uint8_t *compressed, size_t n_plus_1, spg_gens **out) {
  SPG_CHECK(ctx && compressed && out, "spg_gens_upload: null argument");
  SPG_CHECK(n_plus_1 >= 2, "spg_gens_upload: need at least one generator and h");
  spg_gens *g = new (std::nothrow) spg_gens();
  if (!g) return SPG_ENOMEM;
  g->ctx = ctx;
  g->n = n_plus_1 - 1;
  uint8_t *d_in = nullptr;
  int *d_bad = nullptr;
  int bad = 0;
  cudaError_t e = cudaMalloc(&g->bases, n_plus_1 * sizeof(ge));
  if (e == cudaSuccess) e = dev_alloc(ctx, &d_in, n_plus_1 * 32);
  if (e == cudaSuccess) e = dev_alloc(ctx, &d_bad, sizeof(int));
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_in, compressed, n_plus_1 * 32, cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(d_bad, 0, sizeof(int), ctx->stream);
  if (e == cudaSuccess) {
    k_decompress<<<(unsigned)((n_plus_1 + 63) / 64), 64, 0, ctx->stream>>>(d_in, n_plus_1, g->bases, d_bad);
    ctx->launches++;
    e = cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (d_in) dev_free(ctx, d_in);
  if (d_bad) dev_free(ctx, d_bad);
  if (e != cudaSuccess) {
    spg_gens_destroy(g);
    return cuda_fail(e, "spg_gens_upload", __FILE__, __LINE__);
  }
  if (bad) {
    spg_gens_destroy(g);
    set_error("spg_gens_upload: generator %d is not a valid ristretto255 encoding", bad - 1);
    return SPG_EINVAL;
  }
  *out = g;
  return SPG_OK;
}

int spg_gens_from_uniform(spg_ctx *ctx, const uint8_t *uniform, size_t n_plus_1, spg_gens **out) {
  SPG_CHECK(ctx && uniform && out, "spg_gens_from_uniform: null argument");
  SPG_CHECK(n_plus_1 >= 2, "spg_gens_from_uniform: need at least one generator and h");
  spg_gens *g = new (std::nothrow) spg_gens();
  if (!g) return SPG_ENOMEM;
  g->ctx = ctx;
  g->n = n_plus_1 - 1;
  uint8_t *d_in = nullptr;
  cudaError_t e = cudaMalloc(&g->bases, n_plus_1 * sizeof(ge));
  if (e == cudaSuccess) e = dev_alloc(ctx, &d_in, n_plus_1 * 64);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_in, uniform, n_plus_1 * 64, cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) {
    k_from_uniform<<<(unsigned)((n_plus_1 + 63) / 64), 64, 0, ctx->stream>>>(d_in, n_plus_1, g->bases);
    ctx->launches++;
    e = cudaStreamSynchronize(ctx->stream);
  }
  if (d_in) dev_free(ctx, d_in);
  if (e != cudaSuccess) {
    spg_gens_destroy(g);
    return cuda_fail(e, "spg_gens_from_uniform", __FILE__, __LINE__);
  }
  *out = g;
  return SPG_OK;
}

void spg_gens_destroy(spg_gens *g) {
  if (!g) return;
  if (g->bases) cudaFree(g->bases);
  if (g->table) cudaFree(g->table);
  delete g;
}

int spg_poly_commit(spg_ctx *ctx, const spg_gens *gens, const spg_vec *poly, size_t L_size,
                    uint8_t *out_compressed) {
  SPG_CHECK(ctx && gens && poly && out_compressed, "spg_poly_commit: null argument");
  SPG_CHECK(L_size >= 1 && poly->n % L_size == 0, "spg_poly_commit: L_size %zu does not divide len %zu", L_size, poly->n);
  size_t R = poly->n / L_size;
  return msm_rows(const_cast<spg_gens *>(gens), poly->d, L_size, R, R, nullptr, out_compressed);
}

int spg_commit_batch(spg_ctx *ctx, const spg_gens *gens, const spg_fq *scalars, size_t len,
                     const spg_fq *blinds, size_t count, uint8_t *out_compressed) {
  SPG_CHECK(ctx && gens && scalars && out_compressed, "spg_commit_batch: null argument");
  SPG_CHECK(len >= 1 && count >= 1, "spg_commit_batch: empty batch");
  fq *d_s = nullptr, *d_b = nullptr;
  SPG_CUDA(dev_alloc(ctx, &d_s, len * count * sizeof(fq)));
  SPG_CUDA(cudaMemcpyAsync(d_s, scalars, len * count * sizeof(fq), cudaMemcpyHostToDevice, ctx->stream));
  if (blinds) {
    SPG_CUDA(dev_alloc(ctx, &d_b, count * sizeof(fq)));
    SPG_CUDA(cudaMemcpyAsync(d_b, blinds, count * sizeof(fq), cudaMemcpyHostToDevice, ctx->stream));
  }
  int rc = msm_rows(const_cast<spg_gens *>(gens), d_s, count, len, len, d_b, out_compressed);
  dev_free(ctx, d_s);
  if (d_b) dev_free(ctx, d_b);
  return rc;
}

}  // extern "C"
