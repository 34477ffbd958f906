// Phase-2 sumcheck of R1CSProof::prove:
//   ZKSumcheckInstanceProof::prove_cubic_disjoint_rounds
//   (/root/reference/src/sumcheck.rs:788-1065), comb = A*B*C with
//   A = eq(rp, p), B = ABC(p | 0, w, y), C = Z bound to rq (p, w, y).
// create also owns the three table builders that precede it in R1CSProof::prove:
//   ABC table (src/r1csproof.rs:431-465), Z_poly.bound_poly_vars_rq (:469-479) and
//   eq(rp) (:482).
// Rounds: y (low bit first, natural order = adjacent pairs) -> w (top bit first;
// tables are re-laid out once in bit-reversed w so the pairs are adjacent again)
// -> p (top bit first, tiny).
#include "r1cs.cuh"
#include "rounds.cuh"

namespace spg {

int eq_evals_device(spg_ctx *ctx, const fq *d_r, const spg_fq *h_r, size_t ell, fq *out, fq *scratch);
int build_suffix_tables(spg_ctx *ctx, const std::vector<hfq> &tau, size_t max_level, fq *buf);

constexpr int RB2 = 128;

// Z_rq[p][w][y] = sum_q E[q] * Z[p][q][w][y]; E = LSB-first eq table of rq_rev, whose
// entries q < Q_p already carry the (1 - r) factors of the rounds after instance p ran
// out of proofs (bound_poly_q, custom_dense_mlpoly.rs:222-244).
// One thread per output scalar. The Q weights are staged through shared memory 64 at a time
// and consumed four per Montgomery dot product (fq_dot4_lazy: one reduction per row of four
// products), which cuts the wide multiplies of this IMAD-bound kernel by 29 %.
constexpr int ZB = 128, ZTILE = 64;
__global__ void __launch_bounds__(ZB)
k_z_bind_rq(const SecView *__restrict__ secs, const fq *__restrict__ E, size_t Q, size_t W,
            unsigned int log_y, fq *__restrict__ out) {
  __shared__ fq Es[ZTILE];
  size_t WY = W << log_y;
  size_t t = (size_t)blockIdx.x * ZB + threadIdx.x;
  bool live = t < WY;
  size_t w = live ? t >> log_y : 0, y = t & (((size_t)1 << log_y) - 1);
  SecView v = secs[w];
  live = live && y < v.copy;
  const fq *src = v.ptr + y;
  fq acc = fq_zero();
  if (v.q_stride == 0) {
    // a short section holds one row for every proof: sum_q E[q] * z = (sum_q E[q]) * z
    if (live) {
      fq es = fq_zero();
      for (size_t q = 0; q < Q; q++) es = fq_add_lazy(es, fq_load(E + q));
      acc = fq_mul_lazy(es, fq_load(src));
    }
  }
  // uniform trip count: every thread of the block takes part in the staging barriers
  for (size_t q0 = 0; q0 < Q; q0 += ZTILE) {
    size_t nq = Q - q0 < (size_t)ZTILE ? Q - q0 : (size_t)ZTILE;
    __syncthreads();
    if (threadIdx.x < nq) Es[threadIdx.x] = fq_load(E + q0 + threadIdx.x);
    __syncthreads();
    if (!live || v.q_stride == 0) continue;
    size_t q = 0;
    for (; q + 4 <= nq; q += 4) {
      fq a[4], b[4];
#pragma unroll
      for (int k = 0; k < 4; k++) {
        a[k] = fq_load_stream(src + (q0 + q + k) * v.q_stride);
        b[k] = Es[q + k];
      }
      acc = fq_add_lazy(acc, fq_dot4_lazy(a, b));
    }
    for (; q < nq; q++)
      acc = fq_add_lazy(acc, fq_mul_lazy(Es[q], fq_load_stream(src + (q0 + q) * v.q_stride)));
  }
  if (t < WY) fq_store(out + t, fq_canon(acc));
}

// one (lo, hi) pair per item; weight = A[p] (constant in the bound variable)
__global__ void __launch_bounds__(RB2)
k2_pair_eval(const fq *__restrict__ B, const fq *__restrict__ C, const Seg *__restrict__ segs, int nseg, const __grid_constant__ SegPack pk,
             unsigned long long total_items, const fq *__restrict__ A, fq *__restrict__ partials) {
  __shared__ fq sm[3 * 32];
  fq acc[3] = {fq_zero(), fq_zero(), fq_zero()};
  for (unsigned long long item = (unsigned long long)blockIdx.x * RB2 + threadIdx.x; item < total_items;
       item += (unsigned long long)gridDim.x * RB2) {
    Seg sg = pick_seg(pk, segs, nseg, item);
    unsigned long long local = item - sg.item_start;
    fq b0, b1, c0, c1;
    if (sg.log_len >= 1) {
      unsigned long long idx = sg.in_off + 2 * local;
      b0 = fq_load_stream(B + idx); b1 = fq_load_stream(B + idx + 1);
      c0 = fq_load_stream(C + idx); c1 = fq_load_stream(C + idx + 1);
    } else {
      unsigned long long idx = sg.in_off + local;
      b0 = fq_load_stream(B + idx); c0 = fq_load_stream(C + idx);
      b1 = c1 = fq_zero();
    }
    comb_accumulate<2>(acc, fq_load(A + sg.rw_off), b0, b1, c0, c1, c0, c1);
  }
  block_sum<3>(acc, sm);
  if (threadIdx.x == 0) {
    partials[blockIdx.x * 3 + 0] = acc[0];
    partials[blockIdx.x * 3 + 1] = acc[1];
    partials[blockIdx.x * 3 + 2] = acc[2];
  }
}

__global__ void __launch_bounds__(RB2)
k2_pair_bind(const fq *__restrict__ B, const fq *__restrict__ C, fq *__restrict__ OB, fq *__restrict__ OC,
             const Seg *__restrict__ segs, int nseg, const __grid_constant__ SegPack pk, unsigned long long total_items, fq r) {
  for (unsigned long long item = (unsigned long long)blockIdx.x * RB2 + threadIdx.x; item < total_items;
       item += (unsigned long long)gridDim.x * RB2) {
    Seg sg = pick_seg(pk, segs, nseg, item);
    unsigned long long local = item - sg.item_start;
    unsigned long long o = sg.out_off + local;
    if (sg.log_len >= 1) {
      unsigned long long idx = sg.in_off + 2 * local;
      fq lo = fq_load_stream(B + idx), hi = fq_load_stream(B + idx + 1);
      fq_store(OB + o, fq_add(lo, fq_mul(r, fq_sub(hi, lo))));
      lo = fq_load_stream(C + idx); hi = fq_load_stream(C + idx + 1);
      fq_store(OC + o, fq_add(lo, fq_mul(r, fq_sub(hi, lo))));
    } else {
      unsigned long long idx = sg.in_off + local;
      fq lo = fq_load_stream(B + idx);
      fq_store(OB + o, fq_sub(lo, fq_mul(r, lo)));
      lo = fq_load_stream(C + idx);
      fq_store(OC + o, fq_sub(lo, fq_mul(r, lo)));
    }
  }
}

// fused bind_j + eval_{j+1}; needs log_len >= 2 everywhere
__global__ void __launch_bounds__(RB2)
k2_quad_bind_eval(const fq *__restrict__ B, const fq *__restrict__ C, fq *__restrict__ OB,
                  fq *__restrict__ OC, const Seg *__restrict__ segs, int nseg, const __grid_constant__ SegPack pk,
                  unsigned long long total_items, fq r, const fq *__restrict__ A,
                  FinishArgs fa) {
  __shared__ fq sm[3 * 32];
  fq acc[3] = {fq_zero(), fq_zero(), fq_zero()};
  for (unsigned long long item = (unsigned long long)blockIdx.x * RB2 + threadIdx.x; item < total_items;
       item += (unsigned long long)gridDim.x * RB2) {
    Seg sg = pick_seg(pk, segs, nseg, item);
    unsigned long long local = item - sg.item_start;
    unsigned long long idx = sg.in_off + 4 * local, o = sg.out_off + 2 * local;
    fq lo, hi, b0, b1, c0, c1;
    lo = fq_load_stream(B + idx); hi = fq_load_stream(B + idx + 1);
    b0 = fq_add(lo, fq_mul(r, fq_sub(hi, lo)));
    lo = fq_load_stream(B + idx + 2); hi = fq_load_stream(B + idx + 3);
    b1 = fq_add(lo, fq_mul(r, fq_sub(hi, lo)));
    fq_store(OB + o, b0); fq_store(OB + o + 1, b1);
    lo = fq_load_stream(C + idx); hi = fq_load_stream(C + idx + 1);
    c0 = fq_add(lo, fq_mul(r, fq_sub(hi, lo)));
    lo = fq_load_stream(C + idx + 2); hi = fq_load_stream(C + idx + 3);
    c1 = fq_add(lo, fq_mul(r, fq_sub(hi, lo)));
    fq_store(OC + o, c0); fq_store(OC + o + 1, c1);
    comb_accumulate<2>(acc, fq_load(A + sg.rw_off), b0, b1, c0, c1, c0, c1);
  }
  block_sum<3>(acc, sm);
  finish_block<3>(fa, acc, sm);
}

// [p][w] (W per instance) -> [p][bitrev(w)] with W' slots, zero padded
__global__ void k2_w_relayout(const fq *__restrict__ in, fq *__restrict__ out, size_t P, size_t W,
                              unsigned int logWp) {
  size_t Wp = (size_t)1 << logWp;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < P * Wp;
       t += (size_t)gridDim.x * blockDim.x) {
    size_t p = t >> logWp, slot = t & (Wp - 1);
    size_t w = logWp ? (__brev((unsigned int)slot) >> (32 - logWp)) : 0;
    out[t] = w < W ? in[p * W + w] : fq_zero();
  }
}

__global__ void k2_p_eval(const fq *__restrict__ A, const fq *__restrict__ B, const fq *__restrict__ C,
                          size_t half, size_t limit, fq *__restrict__ out) {
  __shared__ fq sm[3 * 32];
  fq acc[3] = {fq_zero(), fq_zero(), fq_zero()};
  for (size_t p = threadIdx.x; p < limit; p += blockDim.x) {
    fq a0 = A[p], a1 = A[p + half], b0 = B[p], b1 = B[p + half], c0 = C[p], c1 = C[p + half];
    fq a2, a3, b2, b3, c2, c3;
    line23(a0, a1, a2, a3);
    line23(b0, b1, b2, b3);
    line23(c0, c1, c2, c3);
    acc[0] = fq_add(acc[0], fq_mul(fq_mul(a0, b0), c0));
    acc[1] = fq_add(acc[1], fq_mul(fq_mul(a2, b2), c2));
    acc[2] = fq_add(acc[2], fq_mul(fq_mul(a3, b3), c3));
  }
  block_sum<3>(acc, sm);
  if (threadIdx.x == 0) {
    out[0] = acc[0];
    out[1] = acc[1];
    out[2] = acc[2];
  }
}

__global__ void k2_p_bind(fq *__restrict__ A, fq *__restrict__ B, fq *__restrict__ C, size_t half, fq r) {
  for (size_t p = threadIdx.x; p < half; p += blockDim.x) {
    fq lo = A[p], hi = A[p + half];
    A[p] = fq_add(lo, fq_mul(r, fq_sub(hi, lo)));
    lo = B[p]; hi = B[p + half];
    B[p] = fq_add(lo, fq_mul(r, fq_sub(hi, lo)));
    lo = C[p]; hi = C[p + half];
    C[p] = fq_add(lo, fq_mul(r, fq_sub(hi, lo)));
  }
}

// replicate entry 0 into [1, n)
__global__ void k_scale_vec(fq *__restrict__ v, size_t n, fq c) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    v[i] = fq_mul(v[i], c);
}

__global__ void k2_replicate(fq *__restrict__ v, size_t n) {
  fq x = v[0];
  for (size_t i = 1 + threadIdx.x; i < n; i += blockDim.x) v[i] = x;
}

}  // namespace spg

using namespace spg;

struct spg_sc2 {
  spg_ctx *ctx = nullptr;
  size_t P = 0, Pp = 1, W = 0, Wp = 1;
  size_t ny = 0, nw = 0, np = 0;
  bool single_inst = false;
  std::vector<size_t> Y;
  fq *tab[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};  // [buffer][B, C]
  size_t cap = 0;
  int cur = 0;
  fq *A = nullptr;  // eq(rp), P' entries
  Seg *d_segs = nullptr;
  std::vector<Seg> segs;
  std::vector<unsigned> loglen;
  size_t round = 0;
  bool evaluated = false, have_cached = false;
  spg_fq cached[3];
  size_t p_len = 1;
  bool w_ready = false, p_ready = false;
};

namespace {

int phase2_of(const spg_sc2 *s, size_t round) {
  if (round < s->ny) return 0;
  if (round < s->ny + s->nw) return 1;
  return 2;
}

void build_segs2(spg_sc2 *s, int phase, int quad, unsigned long long *items_out, unsigned long long *out_total) {
  unsigned long long in_off = 0, out_off = 0, items = 0;
  s->segs.resize(s->P);
  for (size_t p = 0; p < s->P; p++) {
    Seg &g = s->segs[p];
    unsigned ll = s->loglen[p];
    unsigned long long rows = phase == 0 ? s->W : 1;
    g.in_off = in_off;
    g.out_off = out_off;
    g.item_start = items;
    g.log_len = ll;
    g.n_rows = (unsigned)rows;
    g.rw_off = (unsigned)p;
    g.log_tiles = 0;
    unsigned long long in_sz = rows << ll;
    unsigned long long out_sz = ll >= 1 ? in_sz >> 1 : in_sz;
    in_off += in_sz;
    out_off += out_sz;
    items += quad ? (in_sz >> 2) : out_sz;
  }
  *items_out = items;
  *out_total = out_off;
}

// y rounds finished: [p][w] -> [p][bitrev w] padded to W'
int enter_w_phase(spg_sc2 *s) {
  if (s->w_ready) return SPG_OK;
  spg_ctx *ctx = s->ctx;
  int nxt = s->cur ^ 1;
  unsigned logWp = log2u(s->Wp);
  for (int k = 0; k < 2; k++)
    SPG_LAUNCH(ctx, k2_w_relayout, grid_for(ctx, s->P * s->Wp, 128), 128, 0, s->tab[s->cur][k],
               s->tab[nxt][k], s->P, s->W, logWp);
  s->cur = nxt;
  for (size_t p = 0; p < s->P; p++) s->loglen[p] = logWp;
  s->w_ready = true;
  return SPG_OK;
}

int enter_p_phase(spg_sc2 *s) {
  if (s->p_ready) return SPG_OK;
  spg_ctx *ctx = s->ctx;
  if (s->Pp > s->P) {
    // C (and B for distinct instances) are zero beyond the last instance; a shared
    // instance (single_inst) has the same ABC for every p of the padded cube and is
    // not bound in MODE_P (sumcheck.rs:965-967): keeping P' identical copies is equivalent.
    SPG_CUDA(cudaMemsetAsync(s->tab[s->cur][1] + s->P, 0, (s->Pp - s->P) * sizeof(fq), ctx->stream));
    if (s->single_inst) SPG_LAUNCH(ctx, k2_replicate, 1, 64, 0, s->tab[s->cur][0], s->Pp);
    else SPG_CUDA(cudaMemsetAsync(s->tab[s->cur][0] + s->P, 0, (s->Pp - s->P) * sizeof(fq), ctx->stream));
  }
  s->p_ready = true;
  return SPG_OK;
}

}  // namespace

// Z_rq[p][w][y] for every instance, written at dst + off[p]; scale multiplies the eq table
static int z_bind_rq_all(spg_ctx *ctx, const spg_zmat *z, const spg_fq *rq_rev, size_t nq, const spg_fq *scale,
                         fq *dst, const std::vector<size_t> &off) {
  fq *Sq = nullptr;
  SPG_CUDA(dev_alloc(ctx, &Sq, ((size_t)2 << nq) * sizeof(fq)));
  std::vector<hfq> tq(nq);
  for (size_t i = 0; i < nq; i++) tq[i] = hfq_from(rq_rev[i]);
  int rc = build_suffix_tables(ctx, tq, nq, Sq);
  fq *E = Sq + ((size_t)1 << nq);
  if (rc == SPG_OK && scale) {
    fq sc;
    memcpy(&sc, scale, 32);
    k_scale_vec<<<grid_for(ctx, (size_t)1 << nq, 128), 128, 0, ctx->stream>>>(E, (size_t)1 << nq, sc);
    ctx->launches++;
  }
  for (size_t p = 0; rc == SPG_OK && p < z->P; p++) {
    size_t WY = z->W * z->num_inputs[p];
    if (z->num_proofs[p] > ((size_t)1 << nq)) {
      set_error("z_bind_rq: instance %zu has %zu proofs but only %zu challenges", p, z->num_proofs[p], nq);
      rc = SPG_EINVAL;
      break;
    }
    ctx->next_units = 32.0 * (double)WY * (double)(z->num_proofs[p] + 1);
    SPG_LAUNCH(ctx, k_z_bind_rq, (unsigned)((WY + ZB - 1) / ZB), ZB, 0, z->views + p * z->W, E, z->num_proofs[p], z->W,
               log2u(z->num_inputs[p]), dst + off[p]);
  }
  if (rc == SPG_OK && cudaGetLastError() != cudaSuccess) rc = cuda_fail(cudaGetLastError(), "k_z_bind_rq", __FILE__, __LINE__);
  dev_free(ctx, Sq);
  return rc;
}

static int sc2_build(spg_ctx *ctx, const spg_r1cs *inst, const spg_zmat *z, const spg_vec *zrq,
                     size_t num_instances, const size_t *num_proofs, size_t max_num_proofs,
                     const size_t *num_inputs, size_t max_num_inputs, size_t num_witness_secs,
                     const spg_fq *rx, const spg_fq *rq_rev, const spg_fq *rp, const spg_fq *r_A,
                     const spg_fq *r_B, const spg_fq *r_C, spg_sc2 **out) {
  SPG_CHECK(ctx && inst && (z || zrq) && out && num_inputs && r_A && r_B && r_C, "spg_sc2_create: null argument");
  size_t P = num_instances;
  SPG_CHECK(P >= 1, "spg_sc2_create: need at least one instance");
  if (z) SPG_CHECK(num_proofs && z->P == P && z->W == num_witness_secs, "spg_sc2_create: z_mat shape mismatch");
  SPG_CHECK(inst->num_instances == 1 || inst->num_instances == P,
            "spg_sc2_create: instance has %zu blocks, proving %zu", inst->num_instances, P);
  SPG_CHECK(is_pow2(max_num_proofs) && is_pow2(max_num_inputs), "spg_sc2_create: maxima must be powers of two");
  // compute_eval_table_sparse_disjoint_rounds asserts W' * Y_max == num_vars (r1csinstance.rs:500)
  SPG_CHECK(next_pow2(num_witness_secs) * max_num_inputs == inst->num_vars,
            "spg_sc2_create: next_pow2(num_witness_secs) * max_num_inputs = %zu != num_vars = %zu",
            next_pow2(num_witness_secs) * max_num_inputs, inst->num_vars);
  bool single = inst->num_instances == 1 && P > 1;
  for (size_t p = 0; p < P; p++) {
    if (z)
      SPG_CHECK(z->num_proofs[p] == num_proofs[p] && z->num_inputs[p] == num_inputs[p],
                "spg_sc2_create: z_mat shape mismatch at instance %zu", p);
    SPG_CHECK(is_pow2(num_inputs[p]) && num_inputs[p] <= max_num_inputs, "spg_sc2_create: bad num_inputs[%zu]", p);
    SPG_CHECK(!single || num_inputs[p] == num_inputs[0],
              "spg_sc2_create: a shared instance requires equal num_inputs (got %zu vs %zu)", num_inputs[p], num_inputs[0]);
  }
  spg_sc2 *s = new (std::nothrow) spg_sc2();
  if (!s) return SPG_ENOMEM;
  s->ctx = ctx;
  s->P = P;
  s->Pp = next_pow2(P);
  s->W = num_witness_secs;
  s->Wp = next_pow2(num_witness_secs);
  s->ny = log2u(max_num_inputs);
  s->nw = log2u(s->Wp);
  s->np = log2u(s->Pp);
  s->single_inst = single;
  s->Y.assign(num_inputs, num_inputs + P);
  s->p_len = s->Pp;
  size_t nq = log2u(max_num_proofs), nx = log2u(inst->max_num_cons);
  SPG_CHECK((nx == 0 || rx) && (nq == 0 || rq_rev || zrq) && (s->np == 0 || rp), "spg_sc2_create: null challenge vector");
  size_t total = 0;
  std::vector<size_t> off(P);
  for (size_t p = 0; p < P; p++) {
    off[p] = total;
    total += s->W * s->Y[p];
  }
  s->cap = std::max(std::max(total, P * s->Wp), s->Pp);
  int rc = SPG_OK;
  auto fail = [&](int code) {
    spg_sc2_destroy(s);
    return code;
  };
  for (int b = 0; b < 2; b++)
    for (int k = 0; k < 2; k++)
      if (dev_alloc(ctx, &s->tab[b][k], s->cap * sizeof(fq)) != cudaSuccess) return fail(cuda_fail(cudaErrorMemoryAllocation, "sc2 tables", __FILE__, __LINE__));
  if (dev_alloc(ctx, &s->A, s->Pp * sizeof(fq)) != cudaSuccess || dev_alloc(ctx, &s->d_segs, P * sizeof(Seg)) != cudaSuccess)
    return fail(cuda_fail(cudaErrorMemoryAllocation, "sc2 aux", __FILE__, __LINE__));

  // scratch: eq(rx) table, suffix tables of rq_rev, challenge staging
  size_t X = inst->max_num_cons;
  fq *scr = nullptr;
  size_t XS = X > s->Pp ? X : s->Pp;  // eq expansion scratch
  if (zrq) SPG_CHECK(zrq->n == total, "spg_sc2_create: Z_rq has %zu entries, expected %zu", zrq->n, total);
  size_t scr_n = X + XS + ((size_t)2 << nq) + s->Pp + nx + s->np + 8;
  if (dev_alloc(ctx, &scr, scr_n * sizeof(fq)) != cudaSuccess) return fail(cuda_fail(cudaErrorMemoryAllocation, "sc2 scratch", __FILE__, __LINE__));
  fq *evals_rx = scr, *eq_scratch = scr + X, *Sq = scr + X + XS, *d_r = Sq + ((size_t)2 << nq) + s->Pp;
  do {
    // evals_rx = EqPolynomial::new(rx).evals()  (src/r1csproof.rs:433)
    if (nx && (rc = cudaMemcpyAsync(d_r, rx, nx * sizeof(fq), cudaMemcpyHostToDevice, ctx->stream)) != cudaSuccess) { rc = cuda_fail((cudaError_t)rc, "rx upload", __FILE__, __LINE__); break; }
    if ((rc = eq_evals_device(ctx, d_r, rx, nx, evals_rx, eq_scratch)) != SPG_OK) break;
    // ABC table straight into B's buffer; a shared instance is expanded to one copy per p
    {
      std::vector<size_t> ncols(inst->num_instances), ooff(inst->num_instances);
      for (size_t p = 0; p < inst->num_instances; p++) {
        ncols[p] = s->Y[p];
        ooff[p] = off[p];
      }
      if ((rc = r1cs_abc_table(ctx, inst, evals_rx, s->W, max_num_inputs, ncols.data(), ooff.data(), r_A, r_B, r_C, s->tab[0][0])) != SPG_OK) break;
      if (single)
        for (size_t p = 1; p < P; p++)
          if (cudaMemcpyAsync(s->tab[0][0] + off[p], s->tab[0][0], s->W * s->Y[0] * sizeof(fq), cudaMemcpyDeviceToDevice, ctx->stream) != cudaSuccess) { rc = cuda_fail(cudaGetLastError(), "ABC replicate", __FILE__, __LINE__); break; }
      if (rc != SPG_OK) break;
    }
    // C = Z bound to rq (computed here, or handed in already summed over the proof shards)
    if (zrq) {
      if (cudaMemcpyAsync(s->tab[0][1], zrq->d, total * sizeof(fq), cudaMemcpyDeviceToDevice, ctx->stream) != cudaSuccess) { rc = cuda_fail(cudaGetLastError(), "Z_rq copy", __FILE__, __LINE__); break; }
    } else {
      if ((rc = z_bind_rq_all(ctx, z, rq_rev, nq, nullptr, s->tab[0][1], off)) != SPG_OK) break;
    }
    // A = EqPolynomial::new(rp).evals()
    if (s->np && cudaMemcpyAsync(d_r, rp, s->np * sizeof(fq), cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) { rc = cuda_fail(cudaGetLastError(), "rp upload", __FILE__, __LINE__); break; }
    if ((rc = eq_evals_device(ctx, d_r, rp, s->np, s->A, eq_scratch)) != SPG_OK) break;
  } while (0);
  dev_free(ctx, scr);  // stream-ordered: no host synchronisation needed before the first round is queued
  if (rc != SPG_OK) return fail(rc);
  s->loglen.resize(P);
  for (size_t p = 0; p < P; p++) s->loglen[p] = log2u(s->Y[p]);
  *out = s;
  return SPG_OK;
}

extern "C" {

int spg_sc2_create(spg_ctx *ctx, const spg_r1cs *inst, const spg_zmat *z, size_t num_instances,
                   const size_t *num_proofs, size_t max_num_proofs, const size_t *num_inputs,
                   size_t max_num_inputs, size_t num_witness_secs, const spg_fq *rx,
                   const spg_fq *rq_rev, const spg_fq *rp, const spg_fq *r_A, const spg_fq *r_B,
                   const spg_fq *r_C, spg_sc2 **out) {
  spg::DeviceGuard _dev(spg::ctx_of(ctx));
  SPG_CHECK(z, "spg_sc2_create: null z_mat");
  return sc2_build(ctx, inst, z, nullptr, num_instances, num_proofs, max_num_proofs, num_inputs, max_num_inputs,
                   num_witness_secs, rx, rq_rev, rp, r_A, r_B, r_C, out);
}

int spg_sc2_create_from_zrq(spg_ctx *ctx, const spg_r1cs *inst, const spg_vec *zrq, size_t num_instances,
                            const size_t *num_inputs, size_t max_num_inputs, size_t num_witness_secs,
                            const spg_fq *rx, const spg_fq *rp, const spg_fq *r_A, const spg_fq *r_B,
                            const spg_fq *r_C, spg_sc2 **out) {
  spg::DeviceGuard _dev(spg::ctx_of(ctx));
  SPG_CHECK(zrq, "spg_sc2_create_from_zrq: null Z_rq");
  return sc2_build(ctx, inst, nullptr, zrq, num_instances, nullptr, 1, num_inputs, max_num_inputs, num_witness_secs,
                   rx, nullptr, rp, r_A, r_B, r_C, out);
}

// One rank's slice of a y-sharded phase 2 (single instance): the flat [w][y] tables are cut into
// `flat_len`-entry chunks, rank r owns entries [flat_off, flat_off + flat_len) -- one contiguous y
// range of one witness section. The summand eq_p * ABC * Z carries no weight over (w, y), so the
// rank runs the unchanged round kernels on its chunk as a (P = 1, W = 1, Y = flat_len) prover: the ABC
// slice is built here from the CSC arrays, the Z slice is read from `zrq` at flat_off (where the
// reduce-scatter over peer memory leaves this rank's sums). After log2(flat_len) rounds the
// chunk is one scalar per table; the remaining y and w rounds run on the gathered scalars
// (spg_sc2_host_tail_*).
int spg_sc2_create_slice(spg_ctx *ctx, const spg_r1cs *inst, const spg_vec *zrq, size_t max_num_inputs,
                         size_t num_witness_secs, size_t flat_off, size_t flat_len, const spg_fq *rx, const spg_fq *r_A,
                         const spg_fq *r_B, const spg_fq *r_C, spg_sc2 **out) {
  spg::DeviceGuard _dev(spg::ctx_of(ctx));
  SPG_CHECK(ctx && inst && zrq && out && r_A && r_B && r_C, "spg_sc2_create_slice: null argument");
  SPG_CHECK(inst->num_instances == 1, "spg_sc2_create_slice: one instance only");
  SPG_CHECK(is_pow2(max_num_inputs) && is_pow2(flat_len) && flat_len >= 1 && flat_len <= max_num_inputs && flat_off % flat_len == 0,
            "spg_sc2_create_slice: chunk [%zu, +%zu) must be an aligned power-of-two part of one section", flat_off, flat_len);
  size_t total = num_witness_secs * max_num_inputs;
  SPG_CHECK(next_pow2(num_witness_secs) * max_num_inputs == inst->num_vars, "spg_sc2_create_slice: W' * Y != num_vars");
  SPG_CHECK(flat_off + flat_len <= total && zrq->n >= total, "spg_sc2_create_slice: chunk exceeds the %zu-entry tables", total);
  size_t nx = log2u(inst->max_num_cons), X = inst->max_num_cons;
  SPG_CHECK(nx == 0 || rx, "spg_sc2_create_slice: null rx");
  spg_sc2 *s = new (std::nothrow) spg_sc2();
  if (!s) return SPG_ENOMEM;
  s->ctx = ctx;
  s->P = s->Pp = 1;
  s->W = s->Wp = 1;
  s->ny = log2u(flat_len);
  s->nw = s->np = 0;
  s->Y.assign(1, flat_len);
  s->p_len = 1;
  s->cap = flat_len;
  auto fail = [&](int code) {
    spg_sc2_destroy(s);
    return code;
  };
  for (int b = 0; b < 2; b++)
    for (int k = 0; k < 2; k++)
      if (dev_alloc(ctx, &s->tab[b][k], s->cap * sizeof(fq)) != cudaSuccess) return fail(cuda_fail(cudaErrorMemoryAllocation, "sc2 slice tables", __FILE__, __LINE__));
  if (dev_alloc(ctx, &s->A, sizeof(fq)) != cudaSuccess || dev_alloc(ctx, &s->d_segs, sizeof(Seg)) != cudaSuccess)
    return fail(cuda_fail(cudaErrorMemoryAllocation, "sc2 slice aux", __FILE__, __LINE__));
  DevTmp scr(ctx);
  if (scr.alloc((2 * X + nx + 8) * sizeof(fq)) != cudaSuccess) return fail(cuda_fail(cudaErrorMemoryAllocation, "sc2 slice scratch", __FILE__, __LINE__));
  fq *evals_rx = scr.as<fq>(), *eq_scratch = evals_rx + X, *d_r = eq_scratch + X;
  int rc = SPG_OK;
  if (nx && cudaMemcpyAsync(d_r, rx, nx * sizeof(fq), cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) return fail(cuda_fail(cudaGetLastError(), "rx upload", __FILE__, __LINE__));
  if ((rc = eq_evals_device(ctx, d_r, rx, nx, evals_rx, eq_scratch)) != SPG_OK) return fail(rc);
  if ((rc = r1cs_abc_slice(ctx, inst, evals_rx, num_witness_secs, max_num_inputs, flat_off, flat_len, r_A, r_B, r_C, s->tab[0][0])) != SPG_OK) return fail(rc);
  if (cudaMemcpyAsync(s->tab[0][1], zrq->d + flat_off, flat_len * sizeof(fq), cudaMemcpyDeviceToDevice, ctx->stream) != cudaSuccess) return fail(cuda_fail(cudaGetLastError(), "Z slice copy", __FILE__, __LINE__));
  if ((rc = eq_evals_device(ctx, d_r, nullptr, 0, s->A, eq_scratch)) != SPG_OK) return fail(rc);  // A = [1]
  s->loglen.assign(1, log2u(flat_len));
  *out = s;
  return SPG_OK;
}

// local rounds of a y-sharded phase 2: evaluate, exchange the 3 partial evaluations through the host
// mailbox, add, bind (the phase-2 twin of spg_sc1_run_rounds_sharded)
int spg_sc2_run_rounds_sharded(spg_sc2 *s, size_t num_rounds, const spg_fq *challenges, spg_fq *evals_out, void *mailbox,
                               size_t slot_stride, int rank, int world, uint64_t *calls) {
  spg::DeviceGuard _dev(spg::ctx_of(s));
  SPG_CHECK(s && challenges && evals_out && mailbox && calls, "spg_sc2_run_rounds_sharded: null argument");
  SPG_CHECK(world >= 1 && world <= 64 && rank >= 0 && rank < world && slot_stride >= 64 + 3 * sizeof(spg_fq),
            "spg_sc2_run_rounds_sharded: bad mailbox geometry");
  int rc = SPG_OK;
  for (size_t j = 0; j < num_rounds && rc == SPG_OK; j++) {
    spg_fq part[3], all[64 * 3];
    if ((rc = spg_sc2_round_eval(s, part)) != SPG_OK) break;
    if ((rc = mailbox_exchange((char *)mailbox, slot_stride, rank, world, ++*calls, part, sizeof part, all)) != SPG_OK) break;
    hfq acc[3] = {hfq_zero(), hfq_zero(), hfq_zero()};
    for (int r = 0; r < world; r++)
      for (int t = 0; t < 3; t++) acc[t] = hfq_add(acc[t], hfq_from(all[3 * r + t]));
    for (int t = 0; t < 3; t++) evals_out[3 * j + t] = hfq_to(acc[t]);
    rc = spg_sc2_round_bind(s, challenges + j);
  }
  if (rc != SPG_OK) spg_mailbox_poison(mailbox, slot_stride, rank, world);
  return rc;
}

// The cross-rank end of a y-sharded phase 2, on the host like spg_sc1_host_tail_*: state = [B | C], G
// scalars each in flat (w, y_high) order; e(t) = scale * sum B(t) C(t), t = 0, 2, 3. mode 0 binds the low
// bit (adjacent pairs: the remaining y rounds), mode 1 the TOP bit (pairs i, i + len/2: the w rounds, which
// the reference binds top first, src/custom_dense_mlpoly.rs:247-264).
int spg_sc2_host_tail_eval(const spg_fq *state, size_t G, size_t len, int mode, const spg_fq *scale, spg_fq e[3]) {
  SPG_CHECK(state && scale && e, "spg_sc2_host_tail_eval: null argument");
  SPG_CHECK(is_pow2(G) && G <= 64 && is_pow2(len) && len >= 2 && len <= G, "spg_sc2_host_tail_eval: bad sizes %zu / %zu", len, G);
  hfq acc[3] = {hfq_zero(), hfq_zero(), hfq_zero()};
  size_t half = len / 2;
  for (size_t i = 0; i < half; i++) {
    size_t i0 = mode ? i : 2 * i, i1 = mode ? i + half : 2 * i + 1;
    hfq lo[2], d[2];
    for (int k = 0; k < 2; k++) {
      lo[k] = hfq_from(state[k * G + i0]);
      d[k] = hfq_sub(hfq_from(state[k * G + i1]), lo[k]);
    }
    for (int pt = 0; pt < 3; pt++) {
      if (pt >= 1)
        for (int k = 0; k < 2; k++) {
          lo[k] = hfq_add(lo[k], d[k]);
          if (pt == 1) lo[k] = hfq_add(lo[k], d[k]);
        }
      acc[pt] = hfq_add(acc[pt], hfq_mul(lo[0], lo[1]));
    }
  }
  hfq sc = hfq_from(*scale);
  for (int pt = 0; pt < 3; pt++) e[pt] = hfq_to(hfq_mul(sc, acc[pt]));
  return SPG_OK;
}

int spg_sc2_host_tail_bind(spg_fq *state, size_t G, size_t len, int mode, const spg_fq *r) {
  SPG_CHECK(state && r, "spg_sc2_host_tail_bind: null argument");
  SPG_CHECK(is_pow2(G) && G <= 64 && is_pow2(len) && len >= 2 && len <= G, "spg_sc2_host_tail_bind: bad sizes %zu / %zu", len, G);
  hfq rr = hfq_from(*r);
  size_t half = len / 2;
  for (int k = 0; k < 2; k++)
    for (size_t i = 0; i < half; i++) {
      size_t i0 = mode ? i : 2 * i, i1 = mode ? i + half : 2 * i + 1;
      hfq lo = hfq_from(state[k * G + i0]), hi = hfq_from(state[k * G + i1]);
      state[k * G + i] = hfq_to(hfq_add(lo, hfq_mul(rr, hfq_sub(hi, lo))));
    }
  return SPG_OK;
}

int spg_zmat_bind_rq(spg_ctx *ctx, const spg_zmat *z, const spg_fq *rq_rev, size_t nq, const spg_fq *scale,
                     spg_vec *out) {
  spg::DeviceGuard _dev(spg::ctx_of(ctx));
  SPG_CHECK(ctx && z && out && (rq_rev || nq == 0), "spg_zmat_bind_rq: null argument");
  std::vector<size_t> off(z->P);
  size_t total = 0;
  for (size_t p = 0; p < z->P; p++) {
    off[p] = total;
    total += z->W * z->num_inputs[p];
  }
  SPG_CHECK(out->n == total, "spg_zmat_bind_rq: output has %zu entries, expected %zu", out->n, total);
  return z_bind_rq_all(ctx, z, rq_rev, nq, scale, out->d, off);
}

// The step between the two phases of a proof sharded over the proof axis, as ONE call: this rank's
// rq-bound partial Z table, scaled by the eq weight of its shard index, lands in its peer-mapped table;
// once every rank has done so (mailbox barrier) the tables are summed over NVLink peer memory
// (spg_peer_sum, or spg_peer_reduce_scatter when phase 2 is sharded the same way); a second barrier
// keeps any rank from overwriting its table while a peer still reads it. The same sequence driven
// from Python costs ~0.2 ms more per proof in interpreter time, which matters at 8 GPUs where the whole
// pass is 3.5 ms.
int spg_zmat_bind_rq_sharded(spg_ctx *ctx, const spg_zmat *z, const spg_fq *rq_rev, size_t nq_local, size_t nq_total,
                             void *const *peer_ptrs, int world, int rank, size_t n, int scatter_only, void *mailbox,
                             size_t slot_stride, uint64_t *calls) {
  spg::DeviceGuard _dev(spg::ctx_of(ctx));
  SPG_CHECK(ctx && z && peer_ptrs && mailbox && calls && (rq_rev || nq_total == 0), "spg_zmat_bind_rq_sharded: null argument");
  SPG_CHECK(world >= 1 && world <= 16 && rank >= 0 && rank < world && nq_local <= nq_total && slot_stride >= 64 + 8,
            "spg_zmat_bind_rq_sharded: bad geometry");
  SPG_CHECK(((size_t)1 << (nq_total - nq_local)) == (size_t)world, "spg_zmat_bind_rq_sharded: %zu shard challenges for %d ranks",
            nq_total - nq_local, world);
  std::vector<size_t> off(z->P);
  size_t total = 0;
  for (size_t p = 0; p < z->P; p++) {
    off[p] = total;
    total += z->W * z->num_inputs[p];
  }
  SPG_CHECK(total == n, "spg_zmat_bind_rq_sharded: the peer tables hold %zu scalars, the Z table %zu", n, total);
  int rc = [&]() -> int {
    spg_fq weight;
    SPG_TRY(spg_fq_host_eq_weight(rq_rev + nq_local, nq_total - nq_local, (uint64_t)rank, &weight));
    SPG_TRY(z_bind_rq_all(ctx, z, rq_rev, nq_local, &weight, (fq *)peer_ptrs[rank], off));
    SPG_CUDA(cudaStreamSynchronize(ctx->stream));
    uint64_t token = 0, all[16];
    SPG_TRY(mailbox_exchange((char *)mailbox, slot_stride, rank, world, ++*calls, &token, sizeof token, all));
    SPG_TRY(scatter_only ? spg_peer_reduce_scatter(ctx, peer_ptrs, world, rank, n) : spg_peer_sum(ctx, peer_ptrs, world, rank, n));
    SPG_CUDA(cudaStreamSynchronize(ctx->stream));
    return mailbox_exchange((char *)mailbox, slot_stride, rank, world, ++*calls, &token, sizeof token, all);
  }();
  if (rc != SPG_OK) spg_mailbox_poison(mailbox, slot_stride, rank, world);
  return rc;
}

// out[p][w][y] = sum_q weights[p][q] * Z[p][q][w][y] with explicit weights (sum_p Q_p of them, instance
// major): what spg_zmat_bind_rq computes when weights[p][q] = eq(rq, q), for a rank of a sharded proof
// whose rows are an arbitrary subset of the batch (it passes the global eq weights of its own rows).
int spg_zmat_bind_weights(spg_ctx *ctx, const spg_zmat *z, const spg_fq *weights, size_t n_weights, const size_t *out_off,
                          spg_vec *out) {
  spg::DeviceGuard _dev(spg::ctx_of(ctx));
  SPG_CHECK(ctx && z && weights && out, "spg_zmat_bind_weights: null argument");
  std::vector<size_t> off(z->P), woff(z->P);
  size_t total = 0, rows = 0;
  for (size_t p = 0; p < z->P; p++) {
    off[p] = out_off ? out_off[p] : total;
    woff[p] = rows;
    total += z->W * z->num_inputs[p];
    rows += z->num_proofs[p];
    SPG_CHECK(off[p] + z->W * z->num_inputs[p] <= out->n, "spg_zmat_bind_weights: instance %zu runs past the output (%zu entries)", p, out->n);
  }
  SPG_CHECK(out_off || out->n == total, "spg_zmat_bind_weights: output has %zu entries, expected %zu", out->n, total);
  SPG_CHECK(n_weights == rows, "spg_zmat_bind_weights: %zu weights for %zu rows", n_weights, rows);
  fq *dW = nullptr;
  SPG_CUDA(dev_alloc(ctx, &dW, rows * sizeof(fq)));
  int rc = [&]() -> int {
    SPG_CUDA(cudaMemcpyAsync(dW, weights, rows * sizeof(fq), cudaMemcpyHostToDevice, ctx->stream));
    for (size_t p = 0; p < z->P; p++) {
      size_t WY = z->W * z->num_inputs[p];
      ctx->next_units = 32.0 * (double)WY * (double)(z->num_proofs[p] + 1);
      SPG_LAUNCH(ctx, k_z_bind_rq, (unsigned)((WY + ZB - 1) / ZB), ZB, 0, z->views + p * z->W, dW + woff[p], z->num_proofs[p],
                 z->W, log2u(z->num_inputs[p]), out->d + off[p]);
    }
    SPG_CUDA(cudaStreamSynchronize(ctx->stream));
    return SPG_OK;
  }();
  dev_free(ctx, dW);
  return rc;
}

size_t spg_sc2_num_rounds(const spg_sc2 *s) { return s ? s->ny + s->nw + s->np : 0; }

int spg_sc2_round_eval(spg_sc2 *s, spg_fq e[3]) {
  spg::DeviceGuard _dev(spg::ctx_of(s));
  SPG_CHECK(s && e, "spg_sc2_round_eval: null argument");
  if (s->round >= spg_sc2_num_rounds(s)) {
    set_error("spg_sc2_round_eval: all rounds are done");
    return SPG_ESTATE;
  }
  if (s->evaluated) {
    set_error("spg_sc2_round_eval: round %zu already evaluated; call round_bind", s->round);
    return SPG_ESTATE;
  }
  spg_ctx *ctx = s->ctx;
  int phase = phase2_of(s, s->round);
  if (phase >= 1) SPG_TRY(enter_w_phase(s));
  if (phase == 2) {
    SPG_TRY(enter_p_phase(s));
    size_t half = s->p_len / 2;
    size_t limit = half < s->P ? half : s->P;
    SPG_LAUNCH(ctx, k2_p_eval, 1, 128, 0, s->A, s->tab[s->cur][0], s->tab[s->cur][1], half, limit, ctx->d_result);
    SPG_TRY(fetch_result(ctx, 3, e));
  } else if (s->have_cached) {
    memcpy(e, s->cached, sizeof s->cached);
    s->have_cached = false;
  } else {
    unsigned long long items = 0, out_total = 0;
    build_segs2(s, phase, 0, &items, &out_total);
    if (s->P > (size_t)SEG_INLINE) SPG_CUDA(cudaMemcpyAsync(s->d_segs, s->segs.data(), s->P * sizeof(Seg), cudaMemcpyHostToDevice, ctx->stream));
    int grid = grid_for(ctx, items, RB2, 4);
    SPG_TRY(ensure_partials(ctx, (size_t)grid * 3));
    SPG_LAUNCH(ctx, k2_pair_eval, grid, RB2, 0, s->tab[s->cur][0], s->tab[s->cur][1], s->d_segs, (int)s->P, make_pack(s->segs),
               items, s->A, ctx->d_partials);
    SPG_TRY(reduce_partials(ctx, ctx->d_partials, grid, 3, ctx->d_result));
    SPG_TRY(fetch_result(ctx, 3, e));
  }
  s->evaluated = true;
  return SPG_OK;
}

int spg_sc2_round_bind(spg_sc2 *s, const spg_fq *r) {
  spg::DeviceGuard _dev(spg::ctx_of(s));
  SPG_CHECK(s && r, "spg_sc2_round_bind: null argument");
  if (!s->evaluated) {
    set_error("spg_sc2_round_bind: round %zu has not been evaluated", s->round);
    return SPG_ESTATE;
  }
  spg_ctx *ctx = s->ctx;
  int phase = phase2_of(s, s->round);
  fq rr;
  memcpy(&rr, r, sizeof rr);
  if (phase == 2) {
    size_t half = s->p_len / 2;
    SPG_LAUNCH(ctx, k2_p_bind, 1, 128, 0, s->A, s->tab[s->cur][0], s->tab[s->cur][1], half, rr);
    s->p_len = half;
  } else {
    size_t j = phase == 0 ? s->round : s->round - s->ny;
    size_t n_phase = phase == 0 ? s->ny : s->nw;
    bool next_same = j + 1 < n_phase;
    unsigned minlen = 64;
    for (size_t p = 0; p < s->P; p++) minlen = s->loglen[p] < minlen ? s->loglen[p] : minlen;
    int nxt = s->cur ^ 1;
    unsigned long long items = 0, out_total = 0;
    if (next_same && minlen >= 2) {
      build_segs2(s, phase, 1, &items, &out_total);
      if (s->P > (size_t)SEG_INLINE) SPG_CUDA(cudaMemcpyAsync(s->d_segs, s->segs.data(), s->P * sizeof(Seg), cudaMemcpyHostToDevice, ctx->stream));
      const bool split = items <= SPLIT_MAX_ITEMS;  // a late round: latency, not throughput (rounds.cuh)
      int grid = split ? (int)((items + SPLIT_ITEMS_PER_BLOCK - 1) / SPLIT_ITEMS_PER_BLOCK) : grid_for(ctx, items, RB2, 4);
      SPG_TRY(ensure_partials(ctx, (size_t)grid * 3));
      FinishArgs fa = finish_args(ctx, grid);
      if (split) {
        SplitTabs T = {{s->tab[s->cur][0], s->tab[s->cur][1], nullptr}, {s->tab[nxt][0], s->tab[nxt][1], nullptr}};
        SPG_LAUNCH(ctx, (k_quad_split<2, 2>), grid, 128, 0, T, s->d_segs, (int)s->P, make_pack(s->segs), items, rr, s->A,
                   (const fq *)nullptr, fa);
      } else {
        SPG_LAUNCH(ctx, k2_quad_bind_eval, grid, RB2, 0, s->tab[s->cur][0], s->tab[s->cur][1], s->tab[nxt][0],
                   s->tab[nxt][1], s->d_segs, (int)s->P, make_pack(s->segs), items, rr, s->A, fa);
      }
      SPG_TRY(finish_result(ctx, fa, grid, 3, s->cached));
      s->have_cached = true;
    } else {
      build_segs2(s, phase, 0, &items, &out_total);
      if (s->P > (size_t)SEG_INLINE) SPG_CUDA(cudaMemcpyAsync(s->d_segs, s->segs.data(), s->P * sizeof(Seg), cudaMemcpyHostToDevice, ctx->stream));
      SPG_LAUNCH(ctx, k2_pair_bind, grid_for(ctx, items, RB2, 8), RB2, 0, s->tab[s->cur][0], s->tab[s->cur][1],
                 s->tab[nxt][0], s->tab[nxt][1], s->d_segs, (int)s->P, make_pack(s->segs), items, rr);
    }
    s->cur = nxt;
    for (size_t p = 0; p < s->P; p++)
      if (s->loglen[p] > 0) s->loglen[p]--;
  }
  s->round++;
  s->evaluated = false;
  return SPG_OK;
}

int spg_sc2_run_rounds(spg_sc2 *s, size_t num_rounds, const spg_fq *challenges, spg_fq *evals_out) {
  spg::DeviceGuard _dev(spg::ctx_of(s));
  SPG_CHECK(s && challenges && evals_out, "spg_sc2_run_rounds: null argument");
  for (size_t j = 0; j < num_rounds; j++) {
    SPG_TRY(spg_sc2_round_eval(s, evals_out + 3 * j));
    SPG_TRY(spg_sc2_round_bind(s, challenges + j));
  }
  return SPG_OK;
}

int spg_sc2_final(spg_sc2 *s, spg_fq claims[3]) {
  spg::DeviceGuard _dev(spg::ctx_of(s));
  SPG_CHECK(s && claims, "spg_sc2_final: null argument");
  if (s->round != spg_sc2_num_rounds(s)) {
    set_error("spg_sc2_final: %zu of %zu rounds bound", s->round, spg_sc2_num_rounds(s));
    return SPG_ESTATE;
  }
  SPG_TRY(enter_w_phase(s));
  spg_ctx *ctx = s->ctx;
  const fq *heads[3] = {s->A, s->tab[s->cur][0], s->tab[s->cur][1]};
  return gather_heads(ctx, heads, 3, claims);
}

void spg_sc2_destroy(spg_sc2 *s) {
  spg::DeviceGuard _dev(spg::ctx_of(s));
  if (!s) return;
  for (int b = 0; b < 2; b++)
    for (int k = 0; k < 2; k++) dev_free(s->ctx, s->tab[b][k]);
  dev_free(s->ctx, s->A);
  dev_free(s->ctx, s->d_segs);
  delete s;
}

}  // extern "C"
