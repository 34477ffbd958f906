// Pedersen vector commitments on the device.
//   Commitments::commit            src/commitments.rs:69-92   (sum_j s_j G_j + blind h)
//   DensePolynomial::commit_inner  src/dense_mlpoly.rs:199-239 (one commitment per matrix row,
//                                                              all rows share the bases)
//   GroupElement::vartime_multiscalar_mul / compress           src/group.rs:98-117
// dalek's Straus/Pippenger choice is irrelevant for parity: the compressed ristretto
// encoding of the sum is canonical (RFC 9496).
//
// Algorithm. All rows of a polynomial commitment use the same R bases, and there are
// thousands of rows, so the bases get fixed-base window tables once (signed digits):
//   T[j][w][d-1] = d * 2^(c*w) * G_j      d in 1..2^(c-1), w in 0..ceil(254/c)-1
// stored as AFFINE precomputed points (y+x, y-x, 2dxy) on eight saturated 32-bit limbs
// (96 bytes, csrc/fe8.cuh). A row commitment is then a pure sum of table entries -- no
// doublings, no buckets, no atomics, zero digits are skipped:
//   C_i = sum_j sum_w sign(d) T[j][w][|d|-1],   d = digit_w(s_ij) in [-(2^(c-1)-1), 2^(c-1)]
// at 7 field multiplications per entry (mixed addition). The window width c (8..13) is the
// largest whose table fits the memory budget: 8192 bases take 3.2 GB at c = 8 (32 additions
// per scalar), 35 GB at c = 12 (22 additions) and 64 GB at c = 13 (20 additions); see pick_window.
// One thread owns (row i, chunk of bases); a block is 128 consecutive rows of the same chunk and
// concurrently resident blocks share a handful of chunks, so table reads are L2 hits. The next
// table entry is fetched before the current addition is computed (its address depends only on
// the scalar's digits).
#include <atomic>
#include <chrono>

#include "common.cuh"
#include "ed25519.cuh"
#include "fe8.cuh"

namespace spg {

// window geometry of a table
struct Win {
  int c;         // digit width in bits
  int wins;      // ceil(254 / c): the top digit absorbs the last carry of the signed recoding
  uint32_t ent;  // entries per window = 2^(c-1) = the largest digit magnitude
};
static inline Win make_win(int c) {
  Win w;
  w.c = c;
  w.wins = (254 + c - 1) / c;
  w.ent = 1u << (c - 1);
  return w;
}

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

// ---------------------------------------------------------------- table construction (fe8)
__device__ __forceinline__ ge8 ge8_double(const ge8 &p) {
  fe8 A = fe8_mul(p.X, p.X), B = fe8_mul(p.Y, p.Y);
  fe8 ZZ = fe8_mul(p.Z, p.Z);
  fe8 C = fe8_add(ZZ, ZZ);
  fe8 H = fe8_add(A, B);
  fe8 XY = fe8_add(p.X, p.Y);
  fe8 E = fe8_sub(H, fe8_mul(XY, XY));
  fe8 G = fe8_sub(A, B);
  fe8 F = fe8_add(C, G);
  ge8 r;
  r.X = fe8_mul(E, F);
  r.Y = fe8_mul(G, H);
  r.Z = fe8_mul(F, G);
  r.T = fe8_mul(E, H);
  return r;
}
__device__ __noinline__ fe8 fe8_sqn(fe8 f, int n) {
#pragma unroll 1
  for (int i = 0; i < n; i++) f = fe8_mul(f, f);
  return f;
}
__device__ __noinline__ fe8 fe8_mul_call(const fe8 &a, const fe8 &b) { return fe8_mul(a, b); }
// z^(p-2) = z^(2^255 - 21) = (z^(2^252 - 3))^8 * z^3
__device__ __noinline__ fe8 fe8_invert(const fe8 &z) {
  fe8 t0 = fe8_sqn(z, 1);
  fe8 t1 = fe8_sqn(t0, 2);
  t1 = fe8_mul_call(z, t1);
  t0 = fe8_mul_call(t0, t1);
  t0 = fe8_sqn(t0, 1);
  t0 = fe8_mul_call(t1, t0);
  t1 = fe8_sqn(t0, 5);
  t0 = fe8_mul_call(t1, t0);
  t1 = fe8_sqn(t0, 10);
  t1 = fe8_mul_call(t1, t0);
  fe8 t2 = fe8_sqn(t1, 20);
  t1 = fe8_mul_call(t2, t1);
  t1 = fe8_sqn(t1, 10);
  t0 = fe8_mul_call(t1, t0);
  t1 = fe8_sqn(t0, 50);
  t1 = fe8_mul_call(t1, t0);
  t2 = fe8_sqn(t1, 100);
  t1 = fe8_mul_call(t2, t1);
  t1 = fe8_sqn(t1, 50);
  t0 = fe8_mul_call(t1, t0);
  t0 = fe8_sqn(t0, 2);
  fe8 p22523 = fe8_mul_call(t0, z);
  fe8 z3 = fe8_mul_call(fe8_sqn(z, 1), z);
  return fe8_mul_call(fe8_sqn(p22523, 3), z3);
}

__device__ inline ge8 ge8_from_ge(const ge &p) {
  ge8 r;
  r.X = fe8_from_fe(p.X);
  r.Y = fe8_from_fe(p.Y);
  r.Z = fe8_from_fe(p.Z);
  r.T = fe8_from_fe(p.T);
  return r;
}

// wb[j * wins + w] = 2^(c w) G_j
__global__ void k_window_bases(const ge *__restrict__ bases, size_t nbases, Win win, ge8 *__restrict__ wb) {
  size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nbases) return;
  ge8 p = ge8_from_ge(bases[j]);
#pragma unroll 1
  for (int w = 0; w < win.wins; w++) {
    wb[j * win.wins + w] = p;
#pragma unroll 1
    for (int k = 0; k < win.c; k++) p = ge8_double(p);
  }
}

// thread (j, w, blk): entries d = blk * TB + 1 .. blk * TB + TB of window w of base j, made affine
// with one shared inversion
constexpr int TB = 16;
__global__ void __launch_bounds__(64)
k_build_table(const ge8 *__restrict__ wb, size_t nwb, Win win, niels8 *__restrict__ table) {
  const uint32_t nblk = win.ent / TB;
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nwb * nblk) return;
  size_t jw = t / nblk;
  uint32_t blk = (uint32_t)(t % nblk);
  const ge8 P = wb[jw];
  // (blk * TB + 1) P, most significant bit first
  uint32_t d0 = blk * TB + 1;
  ge8 m = P;
  int top = 31 - __clz(d0);
#pragma unroll 1
  for (int b = top - 1; b >= 0; b--) {
    m = ge8_double(m);
    if ((d0 >> b) & 1u) m = ge8_add(m, P);
  }
  ge8 pts[TB];
  fe8 pref[TB];
  pts[0] = m;
  pref[0] = m.Z;
#pragma unroll 1
  for (int k = 1; k < TB; k++) {
    m = ge8_add(m, P);
    pts[k] = m;
    pref[k] = fe8_mul_call(pref[k - 1], m.Z);
  }
  fe8 inv = fe8_invert(pref[TB - 1]);
  niels8 *dst = table + jw * win.ent + (size_t)blk * TB;
#pragma unroll 1
  for (int k = TB - 1; k >= 0; k--) {
    fe8 zinv = k ? fe8_mul_call(inv, pref[k - 1]) : inv;
    inv = fe8_mul_call(inv, pts[k].Z);
    fe8 x = fe8_mul_call(pts[k].X, zinv), y = fe8_mul_call(pts[k].Y, zinv);
    niels8 e;
    e.ypx = fe8_add(y, x);
    e.ymx = fe8_sub(y, x);
    e.t2d = fe8_mul_call(fe8_mul_call(x, y), fe8_2d());
    dst[k] = e;
  }
}

// ---------------------------------------------------------------- signed digits
// the scalar as an integer (leave Montgomery form: the group multiplies by the integer value,
// src/scalar/mod.rs:32-36), consumed c bits at a time from the bottom
struct Recoder {
  fq s;
  uint32_t carry;
  __device__ __forceinline__ explicit Recoder(const fq &mont) : s(fq_from_mont(mont)), carry(0) {}
  // next digit: magnitude (0 = skip) and sign
  __device__ __forceinline__ uint32_t next(const Win &win, bool &neg) {
    uint32_t raw = (s.v[0] & ((1u << win.c) - 1u)) + carry;
#pragma unroll
    for (int i = 0; i < 7; i++) s.v[i] = __funnelshift_r(s.v[i], s.v[i + 1], win.c);
    s.v[7] >>= win.c;
    carry = raw > win.ent ? 1u : 0u;
    neg = carry;
    return carry ? (1u << win.c) - raw : raw;
  }
};

// partial[(i * nchunks + k)] = sum over bases j in chunk k of s[i][j] * G_j
__global__ void __launch_bounds__(128, 3)
k_msm_rows(const fq *__restrict__ scalars, size_t L, size_t R, size_t row_stride, const niels8 *__restrict__ table,
           Win win, size_t chunk, size_t nchunks, ge8 *__restrict__ partial) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t k = blockIdx.y;
  if (i >= L) return;
  ge8 acc = ge8_identity();
  niels8 cur;
  bool have = false, cur_neg = false;
  size_t j0 = k * chunk, j1 = j0 + chunk < R ? j0 + chunk : R;
  const fq *row = scalars + i * row_stride;
  fq s_next = j0 < j1 ? fq_load(row + j0) : fq_zero();
#pragma unroll 1
  for (size_t j = j0; j < j1; j++) {
    fq sm = s_next;
    if (j + 1 < j1) s_next = fq_load(row + j + 1);
    if (fq_is_zero(sm)) continue;
    Recoder rc(sm);
    const niels8 *__restrict__ tj = table + j * (size_t)win.wins * win.ent;
#pragma unroll 1
    for (int w = 0; w < win.wins; w++) {
      bool neg;
      uint32_t mag = rc.next(win, neg);
      if (mag) {
        // fetch this entry now, add the previous one while the load is in flight
        niels8 nxt = niels8_load(tj + (size_t)w * win.ent + (mag - 1));
        if (have) acc = ge8_madd(acc, cur, cur_neg);
        cur = nxt;
        cur_neg = neg;
        have = true;
      }
    }
  }
  if (have) acc = ge8_madd(acc, cur, cur_neg);
  partial[i * nchunks + k] = acc;
}

// ---------------------------------------------------------------- many rows: ONE window table per base + Horner
// A polynomial commitment has thousands of rows over the same bases, so the factor 2^(c w) of
// window w need not be folded into the table at all: with the single table H[j][d-1] = d G_j
//   C_i = sum_w 2^(c w) S_{i,w},   S_{i,w} = sum_j sign(d) H[j][|d|-1],  d = digit_w(s_ij),
// and the doublings (c per window and ROW, against R additions per window and row) are noise.
// The table then costs R * 2^(c-1) entries instead of R * wins * 2^(c-1), which buys c = 17 --
// 15 additions per scalar in 48 GiB for 8192 bases -- where the per-window table above stops at
// c = 13 (20 additions, 60 GiB). One thread owns (row i, window w, chunk of bases) and keeps one
// accumulator; the digits come from a recoding pre-pass (one Montgomery reduction per scalar, not
// one per (scalar, window)); k_hsum adds the chunks, k_hfinish runs the Horner chain of a row.
//
// Recoding: s' = s + K with K = sum_{w < wins-1} 2^(c w + c - 1); the unsigned c-bit digits u_w of
// s' give the signed digits d_w = u_w - 2^(c-1) in [-2^(c-1), 2^(c-1) - 1] for w < wins - 1, and
// the top digit (no offset) stays non-negative and below 2^(c-1) + 1 because c * wins >= 254.
// dg[(t * R) + j], t = w * L + i (window-major: the 32 tasks of a warp are 32 rows of ONE window, so
// when the scalars are small -- addresses, timestamps, most of a real witness -- the warps of the high
// windows find only zero digits and retire at once, instead of every warp running its additions with
// the two or three lanes of the low windows active): magnitude | sign << 31 (0 = nothing to add).
struct RecodeK {
  uint32_t k[8];
};
__global__ void __launch_bounds__(256)
k_hrecode(const fq *__restrict__ scalars, size_t L, size_t R, size_t row_stride, Win win, const __grid_constant__ RecodeK K,
          uint32_t *__restrict__ dg) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= L * R) return;
  size_t i = idx / R, j = idx - i * R;
  fq s = fq_from_mont(fq_load(scalars + i * row_stride + j));
  asm("{\n\t"
      "add.cc.u32  %0, %0, %8;\n\t"
      "addc.cc.u32 %1, %1, %9;\n\t"
      "addc.cc.u32 %2, %2, %10;\n\t"
      "addc.cc.u32 %3, %3, %11;\n\t"
      "addc.cc.u32 %4, %4, %12;\n\t"
      "addc.cc.u32 %5, %5, %13;\n\t"
      "addc.cc.u32 %6, %6, %14;\n\t"
      "addc.u32    %7, %7, %15;\n\t"
      "}"
      : "+r"(s.v[0]), "+r"(s.v[1]), "+r"(s.v[2]), "+r"(s.v[3]), "+r"(s.v[4]), "+r"(s.v[5]), "+r"(s.v[6]), "+r"(s.v[7])
      : "r"(K.k[0]), "r"(K.k[1]), "r"(K.k[2]), "r"(K.k[3]), "r"(K.k[4]), "r"(K.k[5]), "r"(K.k[6]), "r"(K.k[7]));
  const uint32_t mask = (1u << win.c) - 1u;
  uint32_t *out = dg + i * R + j;
  const size_t wstride = L * R;
#pragma unroll 1
  for (int w = 0; w + 1 < win.wins; w++) {
    int d = (int)(s.v[0] & mask) - (int)win.ent;
#pragma unroll
    for (int l = 0; l < 7; l++) s.v[l] = __funnelshift_r(s.v[l], s.v[l + 1], win.c);
    s.v[7] >>= win.c;
    out[(size_t)w * wstride] = d < 0 ? ((uint32_t)(-d) | 0x80000000u) : (uint32_t)d;
  }
  out[(size_t)(win.wins - 1) * wstride] = s.v[0];  // what is left: below 2^(c-1) + 1
}

// profiling only: the number of non-zero digits = the point additions k_msm_hrows performs
__global__ void k_count_nonzero_u32(const uint32_t *__restrict__ dg, size_t n, unsigned long long *__restrict__ count) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned int mine = 0;
  for (; t < n; t += (size_t)gridDim.x * blockDim.x) mine += dg[t] != 0;
  mine = __reduce_add_sync(0xffffffffu, mine);
  if ((threadIdx.x & 31) == 0 && mine) atomicAdd(count, (unsigned long long)mine);
}

// partial[t * nchunks + k] = sum over bases j in chunk k of sign * H[j][mag - 1] for task t = (window, row)
// (MINB = 4: 124 registers, 16 warps per SM; MINB = 3: 135 registers. SPG_MSM_HROWS_MINB=3 selects the latter.)
template <int MINB>
__global__ void __launch_bounds__(128, MINB)
k_msm_hrows(const uint32_t *__restrict__ dg, size_t ntasks, size_t R, const niels8 *__restrict__ htab, uint32_t ent,
            size_t chunk, size_t nchunks, ge8 *__restrict__ partial) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t k = blockIdx.y;
  if (t >= ntasks) return;
  size_t j0 = k * chunk, j1 = j0 + chunk < R ? j0 + chunk : R;  // both multiples of four
  const uint4 *__restrict__ d4 = reinterpret_cast<const uint4 *>(dg + t * R);
  ge8 acc = ge8_identity();
  niels8 cur;
  bool have = false, cur_neg = false;
  uint4 nx = j0 < j1 ? __ldg(d4 + j0 / 4) : make_uint4(0, 0, 0, 0);
#pragma unroll 1
  for (size_t j = j0; j < j1; j += 4) {
    uint4 q = nx;
    if (j + 4 < j1) nx = __ldg(d4 + j / 4 + 1);
#pragma unroll 1
    for (int u = 0; u < 4; u++) {
      uint32_t dv = q.x;
      q.x = q.y;
      q.y = q.z;
      q.z = q.w;
      uint32_t mag = dv & 0x7fffffffu;
      if (mag) {
        // fetch this entry now, add the previous one while the load is in flight
        niels8 nxt = niels8_load(htab + (j + u) * (size_t)ent + (mag - 1));
        if (have) acc = ge8_madd(acc, cur, cur_neg);
        cur = nxt;
        cur_neg = (dv >> 31) != 0;
        have = true;
      }
    }
  }
  if (have) acc = ge8_madd(acc, cur, cur_neg);
  partial[t * nchunks + k] = acc;
}

// S[i * wins + w] = sum over the chunks of task t = w * L + i
__global__ void __launch_bounds__(128)
k_hsum(const ge8 *__restrict__ partial, size_t L, size_t wins, size_t nchunks, ge8 *__restrict__ S) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= L * wins) return;
  ge8 acc = partial[t * nchunks];
#pragma unroll 1
  for (size_t k = 1; k < nchunks; k++) acc = ge8_add(acc, partial[t * nchunks + k]);
  size_t w = t / L, i = t - w * L;
  S[i * wins + w] = acc;
}

// few rows, many bases (the L / R vectors of a bullet reduction round, Cx of an opening):
// WIDE_SPLIT threads per base, each adding a quarter of the base's windows (the additions of one
// scalar are a dependent chain), a block sums its 128 points through shared memory.
// grid (ceil(R / WIDE_BASES), L); partial[i * gridDim.x + blockIdx.x]
constexpr int WIDE_SPLIT = 4, WIDE_BASES = 128 / WIDE_SPLIT;
__global__ void __launch_bounds__(128)
k_msm_wide(const fq *__restrict__ scalars, size_t R, size_t row_stride, const niels8 *__restrict__ table, Win win,
           ge8 *__restrict__ partial) {
  const int per = (win.wins + WIDE_SPLIT - 1) / WIDE_SPLIT;
  __shared__ ge8 sm[64];
  size_t i = blockIdx.y;
  size_t j = (size_t)blockIdx.x * WIDE_BASES + threadIdx.x / WIDE_SPLIT;
  const int part = threadIdx.x % WIDE_SPLIT;
  ge8 acc = ge8_identity();
  if (j < R) {
    fq sm_ = fq_load(scalars + i * row_stride + j);
    if (!fq_is_zero(sm_)) {
      Recoder rc(sm_);
      const niels8 *__restrict__ tj = table + j * (size_t)win.wins * win.ent;
      const int w0 = part * per, w1 = min(w0 + per, win.wins);
#pragma unroll 1
      for (int w = 0; w < w1; w++) {  // the recoding is sequential (carries); only [w0, w1) is added here
        bool neg;
        uint32_t mag = rc.next(win, neg);
        if (mag && w >= w0) acc = ge8_madd(acc, niels8_load(tj + (size_t)w * win.ent + (mag - 1)), neg);
      }
    }
  }
  for (int half = 64; half >= 1; half >>= 1) {
    if ((int)threadIdx.x >= half && (int)threadIdx.x < 2 * half) sm[threadIdx.x - half] = acc;
    __syncthreads();
    if ((int)threadIdx.x < half) acc = ge8_add(acc, sm[threadIdx.x]);
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[i * gridDim.x + blockIdx.x] = acc;
}

// acc += blind * h, h's table in slot `hslot`
__device__ __forceinline__ void add_blind(ge8 &acc, const fq &blind, const niels8 *__restrict__ table, size_t hslot,
                                          const Win &win) {
  if (fq_is_zero(blind)) return;
  Recoder rc(blind);
  const niels8 *__restrict__ th = table + hslot * (size_t)win.wins * win.ent;
#pragma unroll 1
  for (int w = 0; w < win.wins; w++) {
    bool neg;
    uint32_t mag = rc.next(win, neg);
    if (mag) acc = ge8_madd(acc, niels8_load(th + (size_t)w * win.ent + (mag - 1)), neg);
  }
}

__device__ __forceinline__ void store_compressed(const ge8 &acc, uint8_t *__restrict__ out) {
  uint8_t enc[32];
  ristretto_compress(ge8_to_ge(acc), enc);
  for (int k = 0; k < 32; k++) out[k] = enc[k];
}

// the point itself instead of its encoding: X, Y, Z, T as four canonical 32-byte field elements. A caller
// that goes on adding to the point (a bullet round adds c Q on the host) skips the inversion + square root
// inside the ristretto encoding here (one thread, ~0.1 ms) and the decoding on its side.
__device__ __forceinline__ void store_ext(const ge8 &acc, uint8_t *__restrict__ out) {
  const fe8 *c[4] = {&acc.X, &acc.Y, &acc.Z, &acc.T};
  for (int k = 0; k < 4; k++) {
    uint8_t enc[32];
    fe_tobytes(fe8_to_fe(*c[k]), enc);
    for (int j = 0; j < 32; j++) out[32 * k + j] = enc[j];
  }
}
__device__ __forceinline__ void store_point(const ge8 &acc, uint8_t *__restrict__ out, size_t i, int ext) {
  if (ext) store_ext(acc, out + 128 * i);
  else store_compressed(acc, out + 32 * i);
}

// out[i] = compress(sum_k partial[i][k] + blind[i] * h)
__global__ void k_msm_finish(const ge8 *__restrict__ partial, size_t L, size_t nchunks,
                             const fq *__restrict__ blinds, const niels8 *__restrict__ table, size_t hslot, Win win,
                             uint8_t *__restrict__ out, int ext) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= L) return;
  ge8 acc = partial[i * nchunks];
#pragma unroll 1
  for (size_t k = 1; k < nchunks; k++) acc = ge8_add(acc, partial[i * nchunks + k]);
  if (blinds) add_blind(acc, blinds[i], table, hslot, win);
  store_point(acc, out, i, ext);
}

// out[i] = compress(sum_w 2^(c w) S[i][w] + blind[i] * h): the Horner chain of one row, top window first
// (c doublings and one addition per window; the blind uses the per-window table of h)
__global__ void __launch_bounds__(64)
k_hfinish(const ge8 *__restrict__ S, size_t L, Win hwin, const fq *__restrict__ blinds, const niels8 *__restrict__ table,
          size_t hslot, Win win, uint8_t *__restrict__ out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= L) return;
  const ge8 *Si = S + i * (size_t)hwin.wins;
  ge8 acc = Si[hwin.wins - 1];
#pragma unroll 1
  for (int w = hwin.wins - 2; w >= 0; w--) {
#pragma unroll 1
    for (int b = 0; b < hwin.c; b++) acc = ge8_double(acc);
    acc = ge8_add(acc, Si[w]);
  }
  if (blinds) add_blind(acc, blinds[i], table, hslot, win);
  store_compressed(acc, out + 32 * i);
}

// Few rows (the L / R of a bullet-reduction round, Cx, the folded generator): one block per row
// adds the row's per-block partial sums AND the blind's window points as a tree through shared
// memory -- ~8 dependent additions instead of the (nblk + wins) sequential ones a single thread of
// k_msm_finish would do, which is what a round of the opening proof waits for.
__global__ void __launch_bounds__(128)
k_msm_finish_tree(const ge8 *__restrict__ partial, size_t nblk, const fq *__restrict__ blinds,
                  const niels8 *__restrict__ table, size_t hslot, Win win, uint8_t *__restrict__ out, int ext) {
  __shared__ ge8 sm[64];
  const size_t i = blockIdx.x;
  ge8 acc = ge8_identity();
  for (size_t k = threadIdx.x; k < nblk; k += 128) acc = ge8_add(acc, partial[i * nblk + k]);
  if (blinds && (int)threadIdx.x < win.wins && !fq_is_zero(blinds[i])) {
    // window w = threadIdx.x of blind * h (the recoding up to that window is a few shifts)
    Recoder rc(blinds[i]);
    bool neg = false;
    uint32_t mag = 0;
    for (int w = 0; w <= (int)threadIdx.x; w++) mag = rc.next(win, neg);
    if (mag) acc = ge8_madd(acc, niels8_load(table + (hslot * win.wins + threadIdx.x) * (size_t)win.ent + (mag - 1)), neg);
  }
  for (int half = 64; half >= 1; half >>= 1) {
    if ((int)threadIdx.x >= half && (int)threadIdx.x < 2 * half) sm[threadIdx.x - half] = acc;
    __syncthreads();
    if ((int)threadIdx.x < half) acc = ge8_add(acc, sm[threadIdx.x]);
    __syncthreads();
  }
  if (threadIdx.x == 0) store_point(acc, out, i, ext);
}

// profiling only: the number of non-zero scalars (zero scalars cost no additions)
__global__ void k_count_nonzero(const fq *__restrict__ scalars, size_t L, size_t R, size_t row_stride,
                                unsigned long long *__restrict__ count) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long mine = 0;
  for (; t < L * R; t += (size_t)gridDim.x * blockDim.x) mine += !fq_is_zero(fq_load(scalars + (t / R) * row_stride + t % R));
  mine = __reduce_add_sync(0xffffffffu, (unsigned)mine);
  if ((threadIdx.x & 31) == 0 && mine) atomicAdd(count, mine);
}

// ---------------------------------------------------------------- fe8 self-test
// random and edge operands through fe8_{mul,add,sub} and the mixed / full point additions,
// compared with the ten-limb code of ed25519.cuh (which the CPU tests pin to the oracle)
__device__ inline bool fe8_matches(const fe8 &a, const fe &b) { return fe_equal(fe8_to_fe(a), b); }
__global__ void k_fe8_selftest(size_t n, uint64_t seed, unsigned int *__restrict__ bad) {
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
  fe8 a, b;
  for (int i = 0; i < 8; i += 2) {
    uint64_t x = rnd(), y = rnd();
    a.v[i] = (uint32_t)x; a.v[i + 1] = (uint32_t)(x >> 32);
    b.v[i] = (uint32_t)y; b.v[i + 1] = (uint32_t)(y >> 32);
  }
  // edge patterns: all ones (2^256 - 1), values around p and 2p, zero, small
  switch (t % 11) {
    case 1: for (int i = 0; i < 8; i++) a.v[i] = 0xffffffffu; break;
    case 2: for (int i = 0; i < 8; i++) a.v[i] = b.v[i] = 0xffffffffu; break;
    case 3: a = fe8_zero(); break;
    case 4: for (int i = 0; i < 8; i++) b.v[i] = 0xffffffffu; b.v[0] = 0xffffffdau + (uint32_t)(t % 64); break;  // ~2p
    case 5: for (int i = 0; i < 8; i++) a.v[i] = 0xffffffffu; a.v[7] = 0x7fffffffu; a.v[0] = 0xffffffedu + (uint32_t)(t % 32); break;  // ~p
    case 6: b = fe8_zero(); b.v[0] = (uint32_t)(t % 40); break;
    case 7: a = fe8_zero(); a.v[0] = (uint32_t)(t % 40); for (int i = 0; i < 8; i++) b.v[i] = 0xffffffffu; break;
    default: break;
  }
  fe fa = fe8_to_fe(a), fb = fe8_to_fe(b);
  unsigned int err = 0;
  if (!fe8_matches(fe8_mul(a, b), fe_mul(fa, fb))) err |= 1;
  if (!fe8_matches(fe8_add(a, b), fe_add(fa, fb))) err |= 2;
  if (!fe8_matches(fe8_sub(a, b), fe_sub(fa, fb))) err |= 4;
  if (!fe8_matches(fe8_from_fe(fa), fa)) err |= 8;
  // points: P = a-derived multiple of the base point is overkill; use Elligator images of a, b
  uint8_t ub[64];
  for (int i = 0; i < 8; i++)
    for (int k = 0; k < 4; k++) {
      ub[4 * i + k] = (uint8_t)(a.v[i] >> (8 * k));
      ub[32 + 4 * i + k] = (uint8_t)(b.v[i] >> (8 * k));
    }
  ge p = ristretto_map(fe_frombytes(ub)), q = ristretto_map(fe_frombytes(ub + 32));
  ge want = ge_add(p, ge_to_cached(q));
  ge8 p8 = ge8_from_ge(p), q8 = ge8_from_ge(q);
  ge8 got = ge8_add(p8, q8);
  uint8_t e1[32], e2[32];
  ristretto_compress(want, e1);
  ristretto_compress(ge8_to_ge(got), e2);
  for (int i = 0; i < 32; i++) if (e1[i] != e2[i]) err |= 16;
  // mixed addition with the affine form of q, both signs
  fe8 zinv = fe8_invert(q8.Z);
  if (!fe8_matches(fe8_mul(zinv, q8.Z), fe_one())) err |= 32;
  fe8 x = fe8_mul(q8.X, zinv), y = fe8_mul(q8.Y, zinv);
  niels8 nq;
  nq.ypx = fe8_add(y, x);
  nq.ymx = fe8_sub(y, x);
  nq.t2d = fe8_mul(fe8_mul(x, y), fe8_2d());
  ristretto_compress(ge8_to_ge(ge8_madd(p8, nq, false)), e2);
  for (int i = 0; i < 32; i++) if (e1[i] != e2[i]) err |= 64;
  ge qn = q;
  qn.X = fe_neg(q.X);
  qn.T = fe_neg(q.T);
  ristretto_compress(ge_add(p, ge_to_cached(qn)), e1);
  ristretto_compress(ge8_to_ge(ge8_madd(p8, nq, true)), e2);
  for (int i = 0; i < 32; i++) if (e1[i] != e2[i]) err |= 128;
  ristretto_compress(ge_double(p), e1);
  ristretto_compress(ge8_to_ge(ge8_double(p8)), e2);
  for (int i = 0; i < 32; i++) if (e1[i] != e2[i]) err |= 256;
  if (err) atomicOr(bad, err);
}

}  // namespace spg

using namespace spg;

struct spg_gens {
  spg_ctx *ctx = nullptr;
  size_t n = 0;          // number of G's; h is bases[n]
  ge *bases = nullptr;   // n + 1 points
  // window tables for bases [0, tab_R) and h (slot tab_R)
  niels8 *table = nullptr;
  size_t tab_R = 0;
  Win win = make_win(8);
  size_t table_bytes = 0;
  // many-row commitments: one window per base, H[j][d-1] = d G_j for bases [0, htab_R) (Horner over the windows)
  niels8 *htab = nullptr;
  size_t htab_R = 0;
  Win hwin = make_win(8);
  size_t htab_bytes = 0;
  bool htab_failed = false;  // no budget / allocation failed once: do not try again for the same R
  bool htab_ahead = false;   // built by spg_gens_prepare_rows (widest window), not inside a commitment
  bool table_ahead = false;  // the same for the per-window table (spg_gens_prepare / _prepare_rows)
};

namespace {

size_t table_bytes_for(size_t slots, const Win &w) { return slots * (size_t)w.wins * w.ent * sizeof(niels8); }

// Largest window width in [8, 13] whose table fits the budget. Wider windows mean fewer additions
// per scalar (32 at c = 8, 26 at 10, 24 at 11, 22 at 12, 20 at 13) and a table that doubles with each
// bit. Budget of one per-window table: SPG_MSM_TABLE_GIB (default 20: 8193 bases at c = 11 take 18 GiB;
// the many-row commitments, where the additions per scalar decide the time, use the single-window
// table below, and the few-row MSMs this table serves wait on latency, not on their 24 additions),
// and at most 60 % of what is left of the process-wide allowance for tables (60 % of the device's
// memory) and of the memory free right now.
std::atomic<size_t> g_table_bytes{0};
// SPG_MSM_TRACE=1: one line per table build and per MSM call on stderr (which path, which window, how long)
bool msm_trace() {
  static const bool on = getenv("SPG_MSM_TRACE") != nullptr;
  return on;
}
double msm_now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
size_t clamp_table_budget(size_t budget) {
  size_t free_b = 0, total_b = 0;
  if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) {
    size_t allowance = total_b / 100 * 60, used = g_table_bytes.load();
    size_t left = allowance > used ? allowance - used : 0;
    if (budget > left / 10 * 6) budget = left / 10 * 6;
    if (budget > free_b / 10 * 6) budget = free_b / 10 * 6;
  }
  return budget;
}
// L_lazy = rows of the call that builds the table inside itself (0 = built ahead by spg_gens_prepare*, setup):
// a lazily built table has to pay for itself in that call, so its width minimises, per base,
// 43 * wins * 2^(c-1) products to build + 7 * L * wins to use (an entry costs ~43 field products, an addition 7)
// -- c = 8 for the two rows of a bullet round, where the widest table would cost 50 x the MSMs it serves.
Win pick_window(size_t slots, size_t L_lazy) {
  if (const char *e = getenv("SPG_MSM_WINDOW")) {  // development / tests: force a width
    int c = atoi(e);
    if (c >= 5 && c <= 16) return make_win(c);
  }
  double gib = 20.0;
  if (const char *e = getenv("SPG_MSM_TABLE_GIB")) gib = atof(e);
  size_t budget = clamp_table_budget((size_t)(gib * 1073741824.0));
  Win best = make_win(8);
  double best_cost = (43.0 * best.ent + 7.0 * (double)L_lazy) * best.wins;
  for (int c = 9; c <= 13; c++) {
    Win w = make_win(c);
    if (table_bytes_for(slots, w) > budget) continue;
    if (L_lazy) {
      double cost = (43.0 * w.ent + 7.0 * (double)L_lazy) * w.wins;
      if (cost < best_cost) {
        best = w;
        best_cost = cost;
      }
    } else if (w.wins < best.wins) {
      best = w;
    }
  }
  return best;
}

int build_table(spg_gens *g, size_t R, size_t L_lazy) {
  spg_ctx *ctx = g->ctx;
  size_t slots = R + 1;
  Win win = pick_window(slots, L_lazy);
  if (g->table && g->tab_R >= R && g->win.c >= win.c) return SPG_OK;  // what exists is at least as wide
  const double t_start = msm_trace() ? msm_now_ms() : 0;
  niels8 *t = nullptr;
  ge8 *wb = nullptr;
  size_t bytes = table_bytes_for(slots, win);
  SPG_CUDA(cudaMalloc(&t, bytes));
  cudaError_t e = cudaMalloc(&wb, slots * win.wins * sizeof(ge8));
  if (e != cudaSuccess) {
    cudaFree(t);
    return cuda_fail(e, "window bases", __FILE__, __LINE__);
  }
  int rc = [&]() -> int {
    SPG_LAUNCH(ctx, k_window_bases, (unsigned)((R + 63) / 64), 64, 0, g->bases, R, win, wb);
    // h goes into slot R
    SPG_LAUNCH(ctx, k_window_bases, 1, 64, 0, g->bases + g->n, (size_t)1, win, wb + R * win.wins);
    size_t threads = slots * win.wins * (win.ent / TB);
    SPG_LAUNCH(ctx, k_build_table, (unsigned)((threads + 63) / 64), 64, 0, wb, slots * win.wins, win, t);
    SPG_CUDA(cudaStreamSynchronize(ctx->stream));
    return SPG_OK;
  }();
  cudaFree(wb);
  if (rc != SPG_OK) {
    cudaFree(t);
    return rc;
  }
  if (g->table) {
    cudaFree(g->table);
    g_table_bytes -= g->table_bytes;
  }
  g->table = t;
  g->tab_R = R;
  g->win = win;
  g->table_bytes = bytes;
  g_table_bytes += bytes;
  if (msm_trace())
    fprintf(stderr, "[spg msm] per-window table: %zu bases, c = %d (%d windows), %.2f GiB, %s, %.2f ms\n", R, win.c, win.wins,
            (double)bytes / 1073741824.0, L_lazy ? "built inside a call" : "built ahead", msm_now_ms() - t_start);
  return SPG_OK;
}

// L_lazy as in pick_window; a table built ahead replaces a lazily built (narrower) one
int ensure_table(spg_gens *g, size_t R, size_t L_lazy) {
  if (g->table && g->tab_R >= R && (L_lazy || g->table_ahead)) return SPG_OK;
  SPG_TRY(build_table(g, R, L_lazy));
  if (!L_lazy) g->table_ahead = true;
  return SPG_OK;
}

// ---- the single-window table of the many-row path
// Width: the largest c in [9, 17] whose table (R * 2^(c-1) entries of 96 bytes: 48 GiB for 8192 bases at
// c = 17, 15 additions per scalar; 24 GiB at c = 16, 16 additions) fits SPG_MSM_HTABLE_GIB (default 48)
// and the same allowance rules as above; 0 if none gives fewer additions than the per-window table.
// A table built ahead (spg_gens_prepare_rows: setup, like the generators themselves) takes the widest window
// the budget allows. A table built lazily inside the first commitment that wants it has to earn its
// cost in that call: an entry costs ~43 field products to build (k_build_table), an addition 7, so the
// width then minimises 43 * 2^(c-1) + 7 * L * ceil(254 / c) per base -- c = 12 or 13 for 4096 - 8192 rows,
// where the widest table would cost several times the commitment it serves.
size_t htab_bytes_for(size_t R, int c) { return R * ((size_t)1 << (c - 1)) * sizeof(niels8); }
int pick_hwindow(size_t R, int perwindow_wins, size_t L_lazy /* 0: built ahead */) {
  if (const char *e = getenv("SPG_MSM_HWINDOW")) {  // development / tests: force a width
    int c = atoi(e);
    if (c >= 5 && c <= 17) return c;
  }
  double gib = 48.0;
  if (const char *e = getenv("SPG_MSM_HTABLE_GIB")) gib = atof(e);
  size_t budget = clamp_table_budget((size_t)(gib * 1073741824.0));
  int best = 0, best_wins = perwindow_wins;
  double best_cost = 0;
  for (int c = 9; c <= 17; c++) {
    int wins = make_win(c).wins;
    if (htab_bytes_for(R, c) > budget || wins >= perwindow_wins) continue;
    if (L_lazy) {
      double cost = 43.0 * (double)((size_t)1 << (c - 1)) + 7.0 * (double)L_lazy * wins;
      if (best == 0 || cost < best_cost) {
        best = c;
        best_cost = cost;
      }
    } else if (wins < best_wins) {
      best = c;
      best_wins = wins;
    }
  }
  return best;
}

// Is a commitment of L rows over R bases one for the many-row path? (SPG_MSM_HORNER=0 switches it off,
// SPG_MSM_HORNER_MIN sets the smallest L * R, default 2^22: below that the table build does not pay.)
bool horner_wanted(size_t L, size_t R) {
  static const bool enabled = [] {
    const char *e = getenv("SPG_MSM_HORNER");
    return !(e && *e == '0');
  }();
  if (!enabled || L < 128 || R < 16 || (R & 3) != 0) return false;
  size_t min_scalars = (size_t)1 << 22;
  if (const char *e = getenv("SPG_MSM_HORNER_MIN")) min_scalars = (size_t)strtoull(e, nullptr, 10);
  return L * R >= min_scalars;
}

int build_htab(spg_gens *g, size_t R, size_t L_lazy) {
  spg_ctx *ctx = g->ctx;
  int c = pick_hwindow(R, g->table ? g->win.wins : 1 << 30, L_lazy);
  if (c == 0 || (g->htab && g->htab_R >= R && g->hwin.c >= c)) {
    if (!g->htab) g->htab_failed = true;
    return SPG_OK;  // not an error: the per-window path (or the table that exists) serves the call
  }
  Win hw = make_win(c), one = hw;
  one.wins = 1;  // table geometry: one window per base
  const double t_start = msm_trace() ? msm_now_ms() : 0;
  niels8 *t = nullptr;
  ge8 *wb = nullptr;
  size_t bytes = htab_bytes_for(R, c);
  if (cudaMalloc(&t, bytes) != cudaSuccess) {
    cudaGetLastError();
    g->htab_failed = true;
    return SPG_OK;
  }
  cudaError_t e = cudaMalloc(&wb, R * sizeof(ge8));
  if (e != cudaSuccess) {
    cudaFree(t);
    return cuda_fail(e, "window bases", __FILE__, __LINE__);
  }
  int rc = [&]() -> int {
    SPG_LAUNCH(ctx, k_window_bases, (unsigned)((R + 63) / 64), 64, 0, g->bases, R, one, wb);
    size_t threads = R * (one.ent / TB);
    SPG_LAUNCH(ctx, k_build_table, (unsigned)((threads + 63) / 64), 64, 0, wb, R, one, t);
    SPG_CUDA(cudaStreamSynchronize(ctx->stream));
    return SPG_OK;
  }();
  cudaFree(wb);
  if (rc != SPG_OK) {
    cudaFree(t);
    return rc;
  }
  if (g->htab) {
    cudaFree(g->htab);
    g_table_bytes -= g->htab_bytes;
  }
  g->htab = t;
  g->htab_R = R;
  g->hwin = hw;
  g->htab_bytes = bytes;
  g->htab_failed = false;
  g->htab_ahead = L_lazy == 0;
  g_table_bytes += bytes;
  if (msm_trace())
    fprintf(stderr, "[spg msm] single-window table: %zu bases, c = %d (%d additions per scalar), %.2f GiB, %s, %.2f ms\n", R, hw.c,
            hw.wins, (double)bytes / 1073741824.0, L_lazy ? "built inside a call" : "built ahead", msm_now_ms() - t_start);
  return SPG_OK;
}

// L_lazy = rows of the commitment that asks (the table is built inside it), 0 = built ahead at setup time
int ensure_htab(spg_gens *g, size_t R, size_t L_lazy) {
  if (g->htab && g->htab_R >= R && (L_lazy || g->htab_ahead)) return SPG_OK;
  if (g->htab_failed && !g->htab) return SPG_OK;
  return build_htab(g, R, L_lazy);
}

// the many-row path: slabs of rows (the digits of a slab take at most ~1 GiB), see k_msm_hrows
int run_msm_horner(spg_gens *g, const fq *scalars, size_t L, size_t R, size_t row_stride, const fq *d_blinds,
                   uint8_t *d_out, double nonzero_frac) {
  spg_ctx *ctx = g->ctx;
  const Win hw = g->hwin;
  const size_t wins = (size_t)hw.wins;
  static const int minb = [] {
    const char *e = getenv("SPG_MSM_HROWS_MINB");
    return e && *e == '3' ? 3 : 4;
  }();
  // the digits of a slab take at most 4 GiB (a whole 2^26-scalar section at c = 17): every slab ends with a
  // partly filled last wave, so fewer slabs are better
  size_t slab_bytes = (size_t)4 << 30;
  if (const char *e = getenv("SPG_MSM_SLAB_BYTES")) slab_bytes = (size_t)strtoull(e, nullptr, 10);  // tests: several slabs
  size_t slab = slab_bytes / (wins * R * sizeof(uint32_t));
  if (slab < 1) slab = 1;
  if (slab > L) slab = L;
  // chunks of bases: a block's run time is proportional to its chunk (the tail of the grid is one block
  // long) and k_hsum adds one partial per chunk and task (9 products against 7 per table entry), so the
  // chunk is the largest of 256 / 128 / 64 bases that still gives ~16 waves of resident blocks
  size_t task_blocks = (slab * wins + 127) / 128;
  size_t want = (size_t)ctx->sm_count * minb * 16;
  size_t chunk = 256;
  if (const char *e = getenv("SPG_MSM_HROWS_CHUNK")) chunk = (size_t)strtoull(e, nullptr, 10) & ~(size_t)3;
  else
    while (chunk > 64 && task_blocks * ((R + chunk - 1) / chunk) < want) chunk /= 2;
  if (chunk < 4) chunk = 4;
  if (chunk > R) chunk = R;  // R is a multiple of four (horner_wanted)
  size_t nchunks = (R + chunk - 1) / chunk;
  RecodeK K;
  memset(&K, 0, sizeof K);
  for (int w = 0; w + 1 < hw.wins; w++) {
    int b = hw.c * w + hw.c - 1;
    K.k[b >> 5] |= 1u << (b & 31);
  }
  DevTmp t_dg(ctx), t_part(ctx), t_S(ctx);
  SPG_CUDA(t_dg.alloc(slab * wins * R * sizeof(uint32_t)));
  SPG_CUDA(t_part.alloc(slab * wins * nchunks * sizeof(ge8)));
  SPG_CUDA(t_S.alloc(L * wins * sizeof(ge8)));
  uint32_t *dg = t_dg.as<uint32_t>();
  ge8 *partial = t_part.as<ge8>(), *S = t_S.as<ge8>();
  for (size_t r0 = 0; r0 < L; r0 += slab) {
    size_t n = L - r0 < slab ? L - r0 : slab;
    size_t ntasks = n * wins;
    SPG_LAUNCH(ctx, k_hrecode, (unsigned)((n * R + 255) / 256), 256, 0, scalars + r0 * row_stride, n, R, row_stride, hw, K, dg);
    dim3 grid((unsigned)((ntasks + 127) / 128), (unsigned)nchunks);
    double adds = (double)wins * nonzero_frac * (double)n * (double)R;
    if (ctx->profiling) {  // the work units of the launch: one addition per non-zero digit (small scalars have few)
      DevTmp t_cnt(ctx);
      unsigned long long h_cnt = 0;
      if (t_cnt.alloc(sizeof(h_cnt)) == cudaSuccess) {
        cudaMemsetAsync(t_cnt.p, 0, sizeof(h_cnt), ctx->stream);
        k_count_nonzero_u32<<<grid_for(ctx, ntasks * R, 256), 256, 0, ctx->stream>>>(dg, ntasks * R, t_cnt.as<unsigned long long>());
        cudaMemcpyAsync(&h_cnt, t_cnt.p, sizeof(h_cnt), cudaMemcpyDeviceToHost, ctx->stream);
        if (cudaStreamSynchronize(ctx->stream) == cudaSuccess) adds = (double)h_cnt;
      }
    }
    ctx->next_units = adds;
    if (minb == 4) SPG_LAUNCH(ctx, k_msm_hrows<4>, grid, 128, 0, dg, ntasks, R, g->htab, hw.ent, chunk, nchunks, partial);
    else SPG_LAUNCH(ctx, k_msm_hrows<3>, grid, 128, 0, dg, ntasks, R, g->htab, hw.ent, chunk, nchunks, partial);
    SPG_LAUNCH(ctx, k_hsum, (unsigned)((ntasks + 127) / 128), 128, 0, partial, n, wins, nchunks, S + r0 * wins);
  }
  SPG_LAUNCH(ctx, k_hfinish, (unsigned)((L + 63) / 64), 64, 0, S, L, hw, d_blinds, g->table, g->tab_R, g->win, d_out);
  SPG_CUDA(cudaStreamSynchronize(ctx->stream));
  return SPG_OK;
}

int run_msm(spg_gens *g, const fq *scalars, size_t L, size_t R, size_t row_stride, const fq *d_blinds,
            uint8_t *d_out, int ext) {
  spg_ctx *ctx = g->ctx;
  const Win win = g->win;
  // work units of the launch = point additions: windows per scalar x non-zero scalars (counted
  // only while profiling; otherwise every scalar is assumed non-zero)
  double nonzero = 1.0;
  if (ctx->profiling && L * R >= 4096) {
    unsigned long long *d_cnt = nullptr, h_cnt = 0;
    if (dev_alloc(ctx, &d_cnt, sizeof(*d_cnt)) == cudaSuccess) {
      cudaMemsetAsync(d_cnt, 0, sizeof(*d_cnt), ctx->stream);
      k_count_nonzero<<<grid_for(ctx, L * R, 256), 256, 0, ctx->stream>>>(scalars, L, R, row_stride, d_cnt);
      cudaMemcpyAsync(&h_cnt, d_cnt, sizeof(h_cnt), cudaMemcpyDeviceToHost, ctx->stream);
      if (cudaStreamSynchronize(ctx->stream) == cudaSuccess) nonzero = (double)h_cnt / ((double)L * (double)R);
      dev_free(ctx, d_cnt);
    }
  }
  if (!ext && g->htab && g->htab_R >= R && horner_wanted(L, R))
    return run_msm_horner(g, scalars, L, R, row_stride, d_blinds, d_out, nonzero);
  double adds = (double)win.wins * nonzero;
  if (L <= 16 && R >= 256) {
    size_t nblk = (R + WIDE_BASES - 1) / WIDE_BASES;
    ge8 *partial = nullptr;
    SPG_CUDA(dev_alloc(ctx, &partial, L * nblk * sizeof(ge8)));
    dim3 grid((unsigned)nblk, (unsigned)L);
    ctx->next_units = adds * (double)L * (double)R;
    int rc = [&]() -> int {
      SPG_LAUNCH(ctx, k_msm_wide, grid, 128, 0, scalars, R, row_stride, g->table, win, partial);
      SPG_CHECK(win.wins <= 128, "one thread per window of the blind");
      SPG_LAUNCH(ctx, k_msm_finish_tree, (unsigned)L, 128, 0, partial, nblk, d_blinds, g->table, g->tab_R, win, d_out, ext);
      SPG_CUDA(cudaStreamSynchronize(ctx->stream));
      return SPG_OK;
    }();
    dev_free(ctx, partial);
    return rc;
  }
  size_t chunk = 1;
  while (chunk < 64 && (L * R) / (chunk * 2) >= 131072) chunk *= 2;
  size_t nchunks = (R + chunk - 1) / chunk;
  if (nchunks == 0) nchunks = 1;
  ge8 *partial = nullptr;
  SPG_CUDA(dev_alloc(ctx, &partial, L * nchunks * sizeof(ge8)));
  dim3 grid((unsigned)((L + 127) / 128), (unsigned)nchunks);
  ctx->next_units = adds * (double)L * (double)R;
  int rc = [&]() -> int {
    SPG_LAUNCH(ctx, k_msm_rows, grid, 128, 0, scalars, L, R, row_stride, g->table, win, chunk, nchunks, partial);
    SPG_LAUNCH(ctx, k_msm_finish, (unsigned)((L + 63) / 64), 64, 0, partial, L, nchunks, d_blinds, g->table, g->tab_R, win,
               d_out, ext);
    SPG_CUDA(cudaStreamSynchronize(ctx->stream));
    return SPG_OK;
  }();
  dev_free(ctx, partial);
  return rc;
}

int msm_rows(spg_gens *g, const fq *scalars, size_t L, size_t R, size_t row_stride, const fq *d_blinds,
             uint8_t *host_out, int ext = 0) {
  SPG_CHECK(R <= g->n, "commit: %zu scalars per row but only %zu generators", R, g->n);
  // many rows: the single-window table serves the sums; the per-window table is then only needed for blinds
  bool horner = !ext && horner_wanted(L, R);
  if (horner) {
    SPG_TRY(ensure_htab(g, R, L));
    horner = g->htab && g->htab_R >= R;
  }
  if (!horner || d_blinds) SPG_TRY(ensure_table(g, R, L));
  uint8_t *d_out = nullptr;
  spg_ctx *ctx = g->ctx;
  const size_t per = ext ? 128 : 32;  // extended coordinates (store_ext) or the ristretto encoding
  // a handful of points: the finish kernel writes them straight into the context's mapped result
  // page (run_msm ends with a stream synchronise), no device buffer and no copy call
  const bool mapped = L * per <= 48 * sizeof(fq);
  if (mapped) d_out = reinterpret_cast<uint8_t *>(ctx->d_result);
  else SPG_CUDA(dev_alloc(ctx, &d_out, L * per));
  const double t_start = msm_trace() ? msm_now_ms() : 0;
  int rc = run_msm(g, scalars, L, R, row_stride, d_blinds, d_out, ext);
  if (msm_trace())
    fprintf(stderr, "[spg msm] %zu rows x %zu bases%s: %s, %.3f ms\n", L, R, d_blinds ? " + blinds" : "",
            horner ? "single-window table + Horner" : (L <= 16 && R >= 256 ? "per-window table, few-row kernels" : "per-window table"),
            msm_now_ms() - t_start);
  if (rc == SPG_OK) {
    if (mapped) {
      memcpy(host_out, ctx->h_result, L * per);
    } else {
      cudaError_t e = cudaMemcpy(host_out, d_out, L * per, cudaMemcpyDeviceToHost);
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
// the same round with the vectors a and b resident on the device (spg_bullet_set_ab): the fold of s and,
// for the first nk / 2 entries, of a and b (bullet.rs:113-116); reads come from [nk/2, nk) and the thread's
// own entry, writes go to [0, nk/2)
__global__ void k_bullet_fold_ab(fq *__restrict__ s, fq *__restrict__ a, fq *__restrict__ b, size_t n, size_t nk, fq u,
                                 fq u_inv) {
  size_t m = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= n) return;
  size_t nh = nk >> 1;
  fq_store(s + m, fq_mul(fq_load(s + m), (m & (nk - 1)) >= nh ? u : u_inv));
  if (m < nh) {
    fq_store(a + m, fq_add(fq_mul(fq_load(a + m), u), fq_mul(u_inv, fq_load(a + nh + m))));
    fq_store(b + m, fq_add(fq_mul(fq_load(b + m), u_inv), fq_mul(u, fq_load(b + nh + m))));
  }
}
// c_L = <a_L, b_R>, c_R = <a_R, b_L> (bullet.rs:83-84) -> out[0], out[1]; one block
__global__ void __launch_bounds__(256)
k_bullet_inner(const fq *__restrict__ a, const fq *__restrict__ b, size_t nh, fq *__restrict__ out) {
  __shared__ fq sm[2 * 32];
  fq acc[2] = {fq_zero(), fq_zero()};
  for (size_t i = threadIdx.x; i < nh; i += blockDim.x) {
    acc[0] = fq_add(acc[0], fq_mul(fq_load(a + i), fq_load(b + nh + i)));
    acc[1] = fq_add(acc[1], fq_mul(fq_load(a + nh + i), fq_load(b + i)));
  }
  block_sum<2>(acc, sm);
  if (threadIdx.x == 0) {
    fq_store(out, acc[0]);
    fq_store(out + 1, acc[1]);
  }
}
__global__ void k_bullet_heads(const fq *__restrict__ a, const fq *__restrict__ b, fq *__restrict__ out) {
  if (threadIdx.x == 0) {
    fq_store(out, fq_load(a));
    fq_store(out + 1, fq_load(b));
  }
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
  fq *a = nullptr, *bv = nullptr;  // spg_bullet_set_ab: the vectors themselves, folded in place on the device
};

extern "C" {

int spg_bullet_create(spg_ctx *ctx, const spg_gens *gens, size_t n, spg_bullet **out) {
  spg::DeviceGuard _dev(spg::ctx_of(ctx));
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
  spg::DeviceGuard _dev(spg::ctx_of(b));
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
  spg::DeviceGuard _dev(spg::ctx_of(b));
  SPG_CHECK(b && u && u_inv, "spg_bullet_fold: null argument");
  SPG_CHECK(nk >= 2 && nk <= b->n && (nk & (nk - 1)) == 0, "spg_bullet_fold: bad round size %zu", nk);
  fq fu, fi;
  memcpy(&fu, u, sizeof(fq));
  memcpy(&fi, u_inv, sizeof(fq));
  if (b->a)
    SPG_LAUNCH(b->ctx, k_bullet_fold_ab, (unsigned)((b->n + 255) / 256), 256, 0, b->s, b->a, b->bv, b->n, nk, fu, fi);
  else
    SPG_LAUNCH(b->ctx, k_bullet_fold, (unsigned)((b->n + 255) / 256), 256, 0, b->s, b->n, nk, fu, fi);
  return SPG_OK;
}

int spg_bullet_set_ab(spg_bullet *b, const spg_fq *a, const spg_fq *bvec) {
  spg::DeviceGuard _dev(spg::ctx_of(b));
  SPG_CHECK(b && a && bvec, "spg_bullet_set_ab: null argument");
  spg_ctx *ctx = b->ctx;
  if (!b->a) {
    cudaError_t e = dev_alloc(ctx, &b->a, b->n * sizeof(fq));
    if (e == cudaSuccess) e = dev_alloc(ctx, &b->bv, b->n * sizeof(fq));
    if (e != cudaSuccess) {
      if (b->a) dev_free(ctx, b->a);
      b->a = b->bv = nullptr;
      return cuda_fail(e, "spg_bullet_set_ab", __FILE__, __LINE__);
    }
  }
  // through the pinned staging buffer (n + 2 scalars), one vector at a time
  memcpy(b->h_in, a, b->n * sizeof(fq));
  SPG_CUDA(cudaMemcpyAsync(b->a, b->h_in, b->n * sizeof(fq), cudaMemcpyHostToDevice, ctx->stream));
  SPG_CUDA(cudaStreamSynchronize(ctx->stream));
  memcpy(b->h_in, bvec, b->n * sizeof(fq));
  SPG_CUDA(cudaMemcpyAsync(b->bv, b->h_in, b->n * sizeof(fq), cudaMemcpyHostToDevice, ctx->stream));
  SPG_CUDA(cudaStreamSynchronize(ctx->stream));
  return SPG_OK;
}

int spg_bullet_lr_resident(spg_bullet *b, size_t nk, const spg_fq blinds[2], int ext, uint8_t *out_LR, spg_fq out_c[2]) {
  spg::DeviceGuard _dev(spg::ctx_of(b));
  SPG_CHECK(b && blinds && out_LR && out_c, "spg_bullet_lr_resident: null argument");
  SPG_CHECK(b->a, "spg_bullet_lr_resident: call spg_bullet_set_ab first");
  SPG_CHECK(nk >= 2 && nk <= b->n && (nk & (nk - 1)) == 0, "spg_bullet_lr_resident: bad round size %zu", nk);
  spg_ctx *ctx = b->ctx;
  memcpy(b->h_in, blinds, 2 * sizeof(fq));
  SPG_CUDA(cudaMemcpyAsync(b->in, b->h_in, 2 * sizeof(fq), cudaMemcpyHostToDevice, ctx->stream));
  // c_L, c_R into slots 48 and 49 of the mapped result page (56 scalars): msm_rows uses the first 48 at most
  // -- here 2 or 8 for the two points -- and ends with a stream synchronise, so one wait serves both
  SPG_LAUNCH(ctx, k_bullet_inner, 1, 256, 0, b->a, b->bv, nk / 2, ctx->d_result + 48);
  SPG_LAUNCH(ctx, k_bullet_rows, (unsigned)((b->n + 255) / 256), 256, 0, b->a, b->s, b->n, nk, b->rows);
  SPG_TRY(msm_rows(b->gens, b->rows, 2, b->n, b->n, b->in, out_LR, ext ? 1 : 0));
  memcpy(out_c, ctx->h_result + 48, 2 * sizeof(fq));
  return SPG_OK;
}

int spg_bullet_final_ab(spg_bullet *b, uint8_t out_G[32], spg_fq out_ab[2]) {
  spg::DeviceGuard _dev(spg::ctx_of(b));
  SPG_CHECK(b && out_G && out_ab, "spg_bullet_final_ab: null argument");
  SPG_CHECK(b->a, "spg_bullet_final_ab: call spg_bullet_set_ab first");
  spg_ctx *ctx = b->ctx;
  SPG_LAUNCH(ctx, k_bullet_heads, 1, 32, 0, b->a, b->bv, ctx->d_result + 48);
  SPG_TRY(msm_rows(b->gens, b->s, 1, b->n, b->n, nullptr, out_G));
  memcpy(out_ab, ctx->h_result + 48, 2 * sizeof(fq));
  return SPG_OK;
}

int spg_bullet_final(spg_bullet *b, uint8_t out_G[32]) {
  spg::DeviceGuard _dev(spg::ctx_of(b));
  SPG_CHECK(b && out_G, "spg_bullet_final: null argument");
  return msm_rows(b->gens, b->s, 1, b->n, b->n, nullptr, out_G);
}

void spg_bullet_destroy(spg_bullet *b) {
  spg::DeviceGuard _dev(spg::ctx_of(b));
  if (!b) return;
  dev_free(b->ctx, b->s);
  dev_free(b->ctx, b->rows);
  dev_free(b->ctx, b->in);
  if (b->a) dev_free(b->ctx, b->a);
  if (b->bv) dev_free(b->ctx, b->bv);
  if (b->h_in) cudaFreeHost(b->h_in);
  delete b;
}

int spg_gens_upload(spg_ctx *ctx, const uint8_t *compressed, size_t n_plus_1, spg_gens **out) {
  spg::DeviceGuard _dev(spg::ctx_of(ctx));
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
  spg::DeviceGuard _dev(spg::ctx_of(ctx));
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
  spg::DeviceGuard _dev(spg::ctx_of(g));
  if (!g) return;
  if (g->bases) cudaFree(g->bases);
  if (g->table) {
    cudaFree(g->table);
    g_table_bytes -= g->table_bytes;
  }
  if (g->htab) {
    cudaFree(g->htab);
    g_table_bytes -= g->htab_bytes;
  }
  delete g;
}

int spg_poly_commit(spg_ctx *ctx, const spg_gens *gens, const spg_vec *poly, size_t L_size,
                    uint8_t *out_compressed) {
  spg::DeviceGuard _dev(spg::ctx_of(ctx));
  return spg_poly_commit_rows(ctx, gens, poly, L_size, 0, L_size, out_compressed);
}

int spg_poly_commit_rows(spg_ctx *ctx, const spg_gens *gens, const spg_vec *poly, size_t L_size, size_t row0,
                         size_t nrows, uint8_t *out_compressed) {
  spg::DeviceGuard _dev(spg::ctx_of(ctx));
  SPG_CHECK(ctx && gens && poly && out_compressed, "spg_poly_commit: null argument");
  SPG_CHECK(L_size >= 1 && poly->n % L_size == 0, "spg_poly_commit: L_size %zu does not divide len %zu", L_size, poly->n);
  SPG_CHECK(row0 <= L_size && nrows <= L_size - row0, "spg_poly_commit_rows: rows [%zu, %zu) of %zu", row0, row0 + nrows, L_size);
  if (nrows == 0) return SPG_OK;
  size_t R = poly->n / L_size;
  return msm_rows(const_cast<spg_gens *>(gens), poly->d + row0 * R, nrows, R, R, nullptr, out_compressed);
}

int spg_gens_prepare(spg_ctx *ctx, spg_gens *gens, size_t R) {
  spg::DeviceGuard _dev(spg::ctx_of(ctx));
  SPG_CHECK(ctx && gens, "spg_gens_prepare: null argument");
  SPG_CHECK(R >= 1 && R <= gens->n, "spg_gens_prepare: %zu bases requested, %zu generators", R, gens->n);
  return ensure_table(gens, R, 0);
}

int spg_gens_prepare_rows(spg_ctx *ctx, spg_gens *gens, size_t L, size_t R) {
  spg::DeviceGuard _dev(spg::ctx_of(ctx));
  SPG_CHECK(ctx && gens, "spg_gens_prepare_rows: null argument");
  SPG_CHECK(R >= 1 && R <= gens->n, "spg_gens_prepare_rows: %zu bases requested, %zu generators", R, gens->n);
  SPG_TRY(ensure_table(gens, R, 0));
  if (horner_wanted(L, R)) SPG_TRY(ensure_htab(gens, R, 0));
  return SPG_OK;
}

int spg_gens_info_rows(const spg_gens *gens, size_t out[4]) {
  SPG_CHECK(gens && out, "spg_gens_info_rows: null argument");
  out[0] = gens->htab ? (size_t)gens->hwin.c : 0;
  out[1] = gens->htab ? (size_t)gens->hwin.wins : 0;
  out[2] = gens->htab_bytes;
  out[3] = gens->htab_R;
  return SPG_OK;
}

int spg_gens_info(const spg_gens *gens, size_t out[4]) {
  SPG_CHECK(gens && out, "spg_gens_info: null argument");
  out[0] = gens->table ? (size_t)gens->win.c : 0;
  out[1] = gens->table ? (size_t)gens->win.wins : 0;
  out[2] = gens->table_bytes;
  out[3] = gens->tab_R;
  return SPG_OK;
}

int spg_debug_fe8_selftest(spg_ctx *ctx, size_t n, uint64_t seed, uint32_t *out_bad) {
  spg::DeviceGuard _dev(spg::ctx_of(ctx));
  SPG_CHECK(ctx && out_bad && n >= 1, "spg_debug_fe8_selftest: bad argument");
  unsigned int *d_bad = nullptr;
  SPG_CUDA(dev_alloc(ctx, &d_bad, sizeof(unsigned int)));
  int rc = [&]() -> int {
    SPG_CUDA(cudaMemsetAsync(d_bad, 0, sizeof(unsigned int), ctx->stream));
    SPG_LAUNCH(ctx, k_fe8_selftest, (unsigned)((n + 63) / 64), 64, 0, n, seed, d_bad);
    SPG_CUDA(cudaMemcpyAsync(out_bad, d_bad, sizeof(unsigned int), cudaMemcpyDeviceToHost, ctx->stream));
    SPG_CUDA(cudaStreamSynchronize(ctx->stream));
    return SPG_OK;
  }();
  dev_free(ctx, d_bad);
  return rc;
}

int spg_commit_batch(spg_ctx *ctx, const spg_gens *gens, const spg_fq *scalars, size_t len,
                     const spg_fq *blinds, size_t count, uint8_t *out_compressed) {
  spg::DeviceGuard _dev(spg::ctx_of(ctx));
  SPG_CHECK(ctx && gens && scalars && out_compressed, "spg_commit_batch: null argument");
  SPG_CHECK(len >= 1 && count >= 1, "spg_commit_batch: empty batch");
  DevTmp t_s(ctx), t_b(ctx);
  SPG_CUDA(t_s.alloc(len * count * sizeof(fq)));
  SPG_CUDA(cudaMemcpyAsync(t_s.p, scalars, len * count * sizeof(fq), cudaMemcpyHostToDevice, ctx->stream));
  if (blinds) {
    SPG_CUDA(t_b.alloc(count * sizeof(fq)));
    SPG_CUDA(cudaMemcpyAsync(t_b.p, blinds, count * sizeof(fq), cudaMemcpyHostToDevice, ctx->stream));
  }
  return msm_rows(const_cast<spg_gens *>(gens), t_s.as<fq>(), count, len, len, t_b.as<fq>(), out_compressed);
}

}  // extern "C"
