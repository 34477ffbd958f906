// Host-side scalars, group elements and Pedersen generators for the C++ mirror of the
// reference's transcript-side code. Group arithmetic reuses csrc/ed25519.cuh (the same
// source the CUDA commitment kernels compile); scalars reuse csrc/host_fq.h.
//   reference: src/group.rs, src/commitments.rs, src/transcript.rs, src/random.rs
#pragma once
#include <memory>
#include <stdexcept>

#include "../csrc/ed25519.cuh"
#include "../csrc/host_fq.h"
#include "merlin.hpp"

namespace sph {

using spg::hfq;

// ---------------------------------------------------------------- Scalar
struct Scalar {
  hfq v;
  Scalar() : v(spg::hfq_zero()) {}
  explicit Scalar(const hfq &x) : v(x) {}
  static Scalar zero() { return Scalar(); }
  static Scalar one() { return Scalar(spg::hfq_one()); }
  static Scalar from_u64(uint64_t x) { return Scalar(spg::hfq_from_u64(x)); }
  static Scalar from_fq(const spg_fq &x) { return Scalar(spg::hfq_from(x)); }
  spg_fq to_fq() const { return spg::hfq_to(v); }
  Scalar operator+(const Scalar &o) const { return Scalar(spg::hfq_add(v, o.v)); }
  Scalar operator-(const Scalar &o) const { return Scalar(spg::hfq_sub(v, o.v)); }
  Scalar operator*(const Scalar &o) const { return Scalar(spg::hfq_mul(v, o.v)); }
  Scalar operator-() const { return Scalar(spg::hfq_neg(v)); }
  Scalar &operator+=(const Scalar &o) { return *this = *this + o; }
  Scalar &operator*=(const Scalar &o) { return *this = *this * o; }
  bool operator==(const Scalar &o) const { return spg::hfq_eq(v, o.v); }
  Scalar invert() const { return Scalar(spg::hfq_invert(v)); }
  void to_bytes(uint8_t out[32]) const { spg::hfq_to_bytes(v, out); }  // Scalar::to_bytes
  // Scalar::from_bytes_wide (src/scalar/ristretto255.rs:435-446)
  static Scalar from_bytes_wide(const uint8_t b[64]) {
    uint64_t w[8];
    for (int i = 0; i < 8; i++) {
      w[i] = 0;
      for (int k = 7; k >= 0; k--) w[i] = (w[i] << 8) | b[8 * i + k];
    }
    return Scalar(spg::hfq_from_u512(w));
  }
};

// ---------------------------------------------------------------- GroupElement
struct Compressed {
  uint8_t b[32];
  bool operator==(const Compressed &o) const { return memcmp(b, o.b, 32) == 0; }
};

struct Point {
  spg::ge p;
  Point() : p(spg::ge_identity()) {}
  explicit Point(const spg::ge &g) : p(g) {}
  Point operator+(const Point &o) const { return Point(spg::ge_add(p, spg::ge_to_cached(o.p))); }
  Point operator-() const {
    spg::ge n = p;
    n.X = spg::fe_neg(p.X);
    n.T = spg::fe_neg(p.T);
    return Point(n);
  }
  Point operator-(const Point &o) const { return *this + (-o); }
  // scalar * point with the canonical integer value of the Montgomery scalar
  // (Scalar::decompress_scalar, src/scalar/mod.rs:32-36); 4-bit fixed windows
  Point operator*(const Scalar &s) const {
    uint8_t k[32];
    s.to_bytes(k);
    spg::ge_cached tab[15];
    spg::ge m = p;
    tab[0] = spg::ge_to_cached(m);
    for (int i = 1; i < 15; i++) {
      m = spg::ge_add(m, tab[0]);
      tab[i] = spg::ge_to_cached(m);
    }
    spg::ge acc = spg::ge_identity();
    for (int i = 63; i >= 0; i--) {
      for (int d = 0; d < 4; d++) acc = spg::ge_double(acc);
      int nib = (k[i >> 1] >> ((i & 1) * 4)) & 15;
      if (nib) acc = spg::ge_add(acc, tab[nib - 1]);
    }
    return Point(acc);
  }
  Compressed compress() const {
    Compressed c;
    spg::ristretto_compress(p, c.b);
    return c;
  }
  // extended coordinates X, Y, Z, T as four canonical 32-byte field elements (spg_bullet_lr_resident, ext = 1)
  static Point from_ext_bytes(const uint8_t b[128]) {
    spg::ge g;
    g.X = spg::fe_frombytes(b);
    g.Y = spg::fe_frombytes(b + 32);
    g.Z = spg::fe_frombytes(b + 64);
    g.T = spg::fe_frombytes(b + 96);
    return Point(g);
  }
  static Point decompress(const Compressed &c) {
    spg::ge g;
    if (!spg::ristretto_decompress(c.b, &g)) throw std::runtime_error("invalid ristretto255 encoding");
    return Point(g);
  }
};

inline Point operator*(const Scalar &s, const Point &p) { return p * s; }

// GroupElement::vartime_multiscalar_mul (src/group.rs:98-117)
inline Point multiscalar_mul(const std::vector<Scalar> &s, const std::vector<Point> &g) {
  if (s.size() != g.size()) throw std::runtime_error("multiscalar_mul: length mismatch");
  Point acc;
  for (size_t i = 0; i < s.size(); i++)
    if (!(s[i] == Scalar::zero())) acc = acc + g[i] * s[i];
  return acc;
}

// Fixed-base scalar multiplication for the handful of generators every sigma protocol and
// every ZK-sumcheck round commits with (gens_1, gens_3, gens_4): 8-bit windows,
//   T[w][d-1] = d * 2^(8w) * P  in cached form, so  s * P = sum_w T[w][digit_w(s)]
// -- 32 additions, no doublings (the same layout the device MSM uses, csrc/msm.cu).
struct FixedBaseTable {
  std::vector<spg::ge_cached> t;  // 32 windows x 255 entries
  explicit FixedBaseTable(const Point &p) : t(32 * 255) {
    spg::ge base = p.p;
    for (int w = 0; w < 32; w++) {
      spg::ge_cached bc = spg::ge_to_cached(base);
      spg::ge m = base;
      t[w * 255] = bc;
      for (int d = 2; d <= 255; d++) {
        m = spg::ge_add(m, bc);
        t[w * 255 + d - 1] = spg::ge_to_cached(m);
      }
      for (int k = 0; k < 8; k++) base = spg::ge_double(base);
    }
  }
  Point mul(const Scalar &s) const {
    uint8_t k[32];
    s.to_bytes(k);
    spg::ge acc = spg::ge_identity();
    for (int w = 0; w < 32; w++)
      if (k[w]) acc = spg::ge_add(acc, t[w * 255 + k[w] - 1]);
    return Point(acc);
  }
};

// RISTRETTO_BASEPOINT_COMPRESSED (src/group.rs:23-24)
inline const uint8_t *basepoint_compressed() {
  static const uint8_t B[32] = {0xe2, 0xf2, 0xae, 0x0a, 0x6a, 0xbc, 0x4e, 0x71, 0xa8, 0x84, 0xa9,
                                0x61, 0xc5, 0x00, 0x51, 0x5f, 0x58, 0xe3, 0x0b, 0x6a, 0xa5, 0x82,
                                0xdd, 0x8d, 0xb6, 0xa6, 0x59, 0x45, 0xe0, 0x8d, 0x2d, 0x76};
  return B;
}

// ---------------------------------------------------------------- MultiCommitGens
struct MultiCommitGens {
  size_t n = 0;
  std::vector<Point> G;
  Point h;
  // optional fixed-base tables for G[0..n) and h (slot n); shared by copies
  std::shared_ptr<std::vector<FixedBaseTable>> tabs;
  void precompute() {
    auto v = std::make_shared<std::vector<FixedBaseTable>>();
    for (auto &g : G) v->emplace_back(g);
    v->emplace_back(h);
    tabs = v;
  }
  MultiCommitGens() {}
  // the SHAKE256 stream MultiCommitGens::new reads its points from (src/commitments.rs:16-26)
  static std::vector<uint8_t> uniform_bytes(size_t n_, const std::string &label) {
    std::vector<uint8_t> in(label.begin(), label.end());
    in.insert(in.end(), basepoint_compressed(), basepoint_compressed() + 32);
    return shake256(in, 64 * (n_ + 1));
  }
  // MultiCommitGens::new (src/commitments.rs:15-33)
  MultiCommitGens(size_t n_, const std::string &label) : n(n_) {
    std::vector<uint8_t> xof = uniform_bytes(n, label);
    for (size_t i = 0; i < n + 1; i++) {
      Point pt(spg::ristretto_from_uniform_bytes(xof.data() + 64 * i));
      if (i < n) G.push_back(pt);
      else h = pt;
    }
  }
  std::pair<MultiCommitGens, MultiCommitGens> split_at(size_t mid) const {
    MultiCommitGens a, b;
    a.n = mid;
    a.G.assign(G.begin(), G.begin() + mid);
    a.h = h;
    b.n = n - mid;
    b.G.assign(G.begin() + mid, G.end());
    b.h = h;
    return {a, b};
  }
  MultiCommitGens scale(const Scalar &s) const {
    MultiCommitGens r;
    r.n = n;
    r.h = h;
    for (auto &g : G) r.G.push_back(g * s);
    return r;
  }
  std::vector<uint8_t> compressed() const {  // G[0..n], h -- the layout spg_gens_upload expects
    std::vector<uint8_t> out;
    for (auto &g : G) {
      Compressed c = g.compress();
      out.insert(out.end(), c.b, c.b + 32);
    }
    Compressed c = h.compress();
    out.insert(out.end(), c.b, c.b + 32);
    return out;
  }
};

// Commitments for Scalar / [Scalar] (src/commitments.rs:69-92)
inline Point commit(const Scalar &v, const Scalar &blind, const MultiCommitGens &g) {
  if (g.n != 1) throw std::runtime_error("commit(scalar): gens.n != 1");
  if (g.tabs) return (*g.tabs)[0].mul(v) + (*g.tabs)[1].mul(blind);
  return g.G[0] * v + g.h * blind;
}
inline Point commit(const std::vector<Scalar> &v, const Scalar &blind, const MultiCommitGens &g) {
  if (g.n < v.size()) throw std::runtime_error("commit(vec): not enough generators");
  Point acc;
  if (g.tabs) {
    for (size_t i = 0; i < v.size(); i++) acc = acc + (*g.tabs)[i].mul(v[i]);
    return acc + (*g.tabs)[g.n].mul(blind);
  }
  for (size_t i = 0; i < v.size(); i++)
    if (!(v[i] == Scalar::zero())) acc = acc + g.G[i] * v[i];
  return acc + g.h * blind;
}

// ---------------------------------------------------------------- ProofTranscript / RandomTape
struct ProofTranscript : Transcript {
  using Transcript::Transcript;
  void append_protocol_name(const std::string &name) { append_message("protocol-name", name); }
  void append_scalar(const std::string &label, const Scalar &s) {
    uint8_t b[32];
    s.to_bytes(b);
    append_message(label, b, 32);
  }
  void append_point(const std::string &label, const Compressed &c) { append_message(label, c.b, 32); }
  void append_scalars(const std::string &label, const std::vector<Scalar> &v) {  // impl for [Scalar]
    append_message(label, "begin_append_vector");
    for (auto &s : v) append_scalar(label, s);
    append_message(label, "end_append_vector");
  }
  Scalar challenge_scalar(const std::string &label) {
    uint8_t buf[64];
    challenge_bytes(label, buf, 64);
    return Scalar::from_bytes_wide(buf);
  }
  std::vector<Scalar> challenge_vector(const std::string &label, size_t len) {
    std::vector<Scalar> v;
    for (size_t i = 0; i < len; i++) v.push_back(challenge_scalar(label));
    return v;
  }
};

// RandomTape (src/random.rs:9-28) with the seed supplied by the caller instead of OsRng:
// the deterministic hook needed for bit-exact parity.
struct RandomTape {
  ProofTranscript tape;
  RandomTape(const std::string &name, const Scalar &seed) : tape(name) { tape.append_scalar("init_randomness", seed); }
  Scalar random_scalar(const std::string &label) { return tape.challenge_scalar(label); }
  std::vector<Scalar> random_vector(const std::string &label, size_t len) { return tape.challenge_vector(label, len); }
};

}  // namespace sph
