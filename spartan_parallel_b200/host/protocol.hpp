// C++ mirror of the transcript-side code of R1CSProof::prove, driving the device through
// the C ABI (include/spgpu.h). Everything here is what the Rust host keeps doing in
// production (sigma protocols, UniPoly, ZK sumcheck glue, opening proofs, serialization);
// the table work is delegated to libspgpu.
//   reference: src/unipoly.rs, src/nizk/mod.rs, src/nizk/bullet.rs,
//              src/sumcheck.rs:788-1380 (ZK glue), src/dense_mlpoly.rs:861-960,
//              src/r1csproof.rs:210-685
#pragma once
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <map>

#include "../../include/spgpu.h"
#include "group.hpp"

namespace sph {

inline void check(int rc, const char *what) {
  if (rc != SPG_OK) throw std::runtime_error(std::string(what) + ": " + spg_last_error());
}
// Wall time of the prover's stages: always recorded (sph_timings returns the log of the last
// proof, which is how bench.py reports the reference's Timer labels); SPH_TRACE=1 also prints it.
inline std::vector<std::pair<std::string, double>> &timing_log() {
  static thread_local std::vector<std::pair<std::string, double>> log;
  return log;
}
struct Trace {
  bool on;
  std::chrono::steady_clock::time_point t0;
  Trace() : on(getenv("SPH_TRACE") != nullptr), t0(std::chrono::steady_clock::now()) {}
  void lap(const char *what) {
    auto t1 = std::chrono::steady_clock::now();
    double ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
    timing_log().emplace_back(what, ms);
    if (on) fprintf(stderr, "[sph] %-28s %8.3f ms\n", what, ms);
    t0 = t1;
  }
};
inline size_t log2z(size_t n) {
  size_t l = 0;
  while (((size_t)1 << l) < n) l++;
  return l;
}
inline size_t next_pow2(size_t n) {
  size_t p = 1;
  while (p < n) p <<= 1;
  return p;
}

// ---------------------------------------------------------------- serialization (bincode 1.x,
// default options: fixed-width little-endian integers, u64 length prefix for Vec, none for
// arrays/tuples; Scalar = its four Montgomery limbs, CompressedRistretto = 32 bytes)
struct Writer {
  std::vector<uint8_t> out;
  void u64(uint64_t x) {
    for (int i = 0; i < 8; i++) out.push_back((uint8_t)(x >> (8 * i)));
  }
  void scalar(const Scalar &s) {
    for (int i = 0; i < 4; i++) u64(s.v.l[i]);
  }
  void point(const Compressed &c) { out.insert(out.end(), c.b, c.b + 32); }
  void scalars(const std::vector<Scalar> &v) {
    u64(v.size());
    for (auto &s : v) scalar(s);
  }
  void points(const std::vector<Compressed> &v) {
    u64(v.size());
    for (auto &p : v) point(p);
  }
};

// ---------------------------------------------------------------- UniPoly (src/unipoly.rs:22-93)
struct UniPoly {
  std::vector<Scalar> coeffs;  // lowest degree first
  static UniPoly from_evals(const std::vector<Scalar> &e) {
    // (2_usize).to_scalar() sums ones (src/scalar/mod.rs:10-15): same value as from_u64
    static const Scalar two_inv = Scalar::from_u64(2).invert();
    UniPoly p;
    if (e.size() == 3) {
      Scalar c = e[0];
      Scalar a = two_inv * (e[2] - e[1] - e[1] + c);
      Scalar b = e[1] - c - a;
      p.coeffs = {c, b, a};
    } else if (e.size() == 4) {
      static const Scalar six_inv = Scalar::from_u64(6).invert();
      Scalar d = e[0];
      Scalar a = six_inv * (e[3] - e[2] - e[2] - e[2] + e[1] + e[1] + e[1] - e[0]);
      Scalar b = two_inv * (e[0] + e[0] - e[1] - e[1] - e[1] - e[1] - e[1] + e[2] + e[2] + e[2] + e[2] - e[3]);
      Scalar c = e[1] - d - a - b;
      p.coeffs = {d, c, b, a};
    } else {
      throw std::runtime_error("UniPoly::from_evals: degree must be 2 or 3");
    }
    return p;
  }
  size_t degree() const { return coeffs.size() - 1; }
  Scalar evaluate(const Scalar &r) const {
    Scalar eval = coeffs[0], power = r;
    for (size_t i = 1; i < coeffs.size(); i++) {
      eval += power * coeffs[i];
      power *= r;
    }
    return eval;
  }
  Point commit(const MultiCommitGens &g, const Scalar &blind) const { return sph::commit(coeffs, blind, g); }
};

// ---------------------------------------------------------------- sigma protocols (src/nizk/mod.rs)
struct KnowledgeProof {
  Compressed alpha;
  Scalar z1, z2;
  static std::pair<KnowledgeProof, Compressed> prove(const MultiCommitGens &g, ProofTranscript &t, RandomTape &tape,
                                                     const Scalar &x, const Scalar &r) {
    t.append_protocol_name("knowledge proof");
    Scalar t1 = tape.random_scalar("t1"), t2 = tape.random_scalar("t2");
    Compressed C = commit(x, r, g).compress();
    t.append_point("C", C);
    KnowledgeProof p;
    p.alpha = commit(t1, t2, g).compress();
    t.append_point("alpha", p.alpha);
    Scalar c = t.challenge_scalar("c");
    p.z1 = x * c + t1;
    p.z2 = r * c + t2;
    return {p, C};
  }
  void write(Writer &w) const {
    w.point(alpha);
    w.scalar(z1);
    w.scalar(z2);
  }
};

struct EqualityProof {
  Compressed alpha;
  Scalar z;
  static EqualityProof prove(const MultiCommitGens &g, ProofTranscript &t, RandomTape &tape, const Scalar &v1,
                             const Scalar &s1, const Scalar &v2, const Scalar &s2) {
    t.append_protocol_name("equality proof");
    Scalar r = tape.random_scalar("r");
    t.append_point("C1", commit(v1, s1, g).compress());
    t.append_point("C2", commit(v2, s2, g).compress());
    EqualityProof p;
    p.alpha = (g.tabs ? (*g.tabs)[g.n].mul(r) : g.h * r).compress();
    t.append_point("alpha", p.alpha);
    Scalar c = t.challenge_scalar("c");
    p.z = c * (s1 - s2) + r;
    return p;
  }
  void write(Writer &w) const {
    w.point(alpha);
    w.scalar(z);
  }
};

struct ProductProof {
  Compressed alpha, beta, delta;
  Scalar z[5];
  struct Out {
    Compressed X, Y, Z;
  };
  static std::pair<ProductProof, Out> prove(const MultiCommitGens &g, ProofTranscript &t, RandomTape &tape,
                                            const Scalar &x, const Scalar &rX, const Scalar &y, const Scalar &rY,
                                            const Scalar &z, const Scalar &rZ) {
    t.append_protocol_name("product proof");
    Scalar b1 = tape.random_scalar("b1"), b2 = tape.random_scalar("b2"), b3 = tape.random_scalar("b3"),
           b4 = tape.random_scalar("b4"), b5 = tape.random_scalar("b5");
    Out o;
    o.X = commit(x, rX, g).compress();
    t.append_point("X", o.X);
    o.Y = commit(y, rY, g).compress();
    t.append_point("Y", o.Y);
    o.Z = commit(z, rZ, g).compress();
    t.append_point("Z", o.Z);
    ProductProof p;
    p.alpha = commit(b1, b2, g).compress();
    t.append_point("alpha", p.alpha);
    p.beta = commit(b3, b4, g).compress();
    t.append_point("beta", p.beta);
    MultiCommitGens gX;
    gX.n = 1;
    gX.G = {Point::decompress(o.X)};
    gX.h = g.h;
    p.delta = commit(b3, b5, gX).compress();
    t.append_point("delta", p.delta);
    Scalar c = t.challenge_scalar("c");
    p.z[0] = b1 + c * x;
    p.z[1] = b2 + c * rX;
    p.z[2] = b3 + c * y;
    p.z[3] = b4 + c * rY;
    p.z[4] = b5 + c * (rZ - rX * y);
    return {p, o};
  }
  void write(Writer &w) const {
    w.point(alpha);
    w.point(beta);
    w.point(delta);
    for (int i = 0; i < 5; i++) w.scalar(z[i]);
  }
};

struct DotProductProof {
  Compressed delta, beta;
  std::vector<Scalar> z;
  Scalar z_delta, z_beta;
  static Scalar dot(const std::vector<Scalar> &a, const std::vector<Scalar> &b) {
    Scalar acc;
    for (size_t i = 0; i < a.size(); i++) acc += a[i] * b[i];
    return acc;
  }
  static DotProductProof prove(const MultiCommitGens &g1, const MultiCommitGens &gn, ProofTranscript &t, RandomTape &tape,
                               const std::vector<Scalar> &x, const Scalar &blind_x, const std::vector<Scalar> &a,
                               const Scalar &y, const Scalar &blind_y, const Compressed *Cx_known = nullptr) {
    t.append_protocol_name("dot product proof");
    size_t n = x.size();
    std::vector<Scalar> d = tape.random_vector("d_vec", n);
    Scalar r_delta = tape.random_scalar("r_delta"), r_beta = tape.random_scalar("r_beta");
    // Cx = commit(x, blind_x): the ZK sumcheck has just computed exactly this point as comm_poly
    // (src/sumcheck.rs:1247-1253 and :1336-1347 commit the same coefficients with the same blind)
    t.append_point("Cx", Cx_known ? *Cx_known : commit(x, blind_x, gn).compress());
    t.append_point("Cy", commit(y, blind_y, g1).compress());
    t.append_scalars("a", a);
    DotProductProof p;
    p.delta = commit(d, r_delta, gn).compress();
    t.append_point("delta", p.delta);
    p.beta = commit(dot(a, d), r_beta, g1).compress();
    t.append_point("beta", p.beta);
    Scalar c = t.challenge_scalar("c");
    for (size_t i = 0; i < n; i++) p.z.push_back(c * x[i] + d[i]);
    p.z_delta = c * blind_x + r_delta;
    p.z_beta = c * blind_y + r_beta;
    return p;
  }
  void write(Writer &w) const {
    w.point(delta);
    w.point(beta);
    w.scalars(z);
    w.scalar(z_delta);
    w.scalar(z_beta);
  }
};

// DotProductProofGens (src/nizk/mod.rs:406-418). `dev` (optional, not owned) is the same
// gens_n resident on the device with its fixed-base window tables: when present the
// n-sized multiscalar multiplications of the opening proofs run there.
struct DotProductProofGens {
  size_t n = 0;
  MultiCommitGens gens_n, gens_1;
  spg_ctx *ctx = nullptr;
  spg_gens *dev = nullptr;
  DotProductProofGens() {}
  DotProductProofGens(size_t n_, const std::string &label) : n(n_) {
    auto pr = MultiCommitGens(n + 1, label).split_at(n);
    gens_n = pr.first;
    gens_1 = pr.second;
  }
  void attach_device(spg_ctx *c, spg_gens *g) {
    ctx = c;
    dev = g;
  }
  // Same generators with the n-sized part derived ON the device (the hash-to-group of
  // thousands of points is the expensive part of MultiCommitGens::new): the host keeps only
  // gens_1 (point n) and h (point n + 1); gens_n.G stays empty. *owned receives the handle.
  static DotProductProofGens on_device(spg_ctx *c, size_t n_, const std::string &label, spg_gens **owned) {
    DotProductProofGens g;
    g.n = n_;
    std::vector<uint8_t> xof = MultiCommitGens::uniform_bytes(n_ + 1, label);  // n + 2 points
    Point gn(spg::ristretto_from_uniform_bytes(xof.data() + 64 * n_));
    Point h(spg::ristretto_from_uniform_bytes(xof.data() + 64 * (n_ + 1)));
    g.gens_n.n = n_;
    g.gens_n.h = h;
    g.gens_1.n = 1;
    g.gens_1.G = {gn};
    g.gens_1.h = h;
    std::vector<uint8_t> dev_bytes(xof.begin(), xof.begin() + 64 * n_);
    dev_bytes.insert(dev_bytes.end(), xof.begin() + 64 * (n_ + 1), xof.end());
    check(spg_gens_from_uniform(c, dev_bytes.data(), n_ + 1, owned), "spg_gens_from_uniform");
    g.attach_device(c, *owned);
    return g;
  }
};

// sum_j s[i][j] G_j + blind[i] h for `count` rows of `len` scalars on the device
inline std::vector<Point> device_msm(const DotProductProofGens &g, const std::vector<Scalar> &rows, size_t len, size_t count,
                                     const std::vector<Scalar> *blinds) {
  std::vector<spg_fq> s(rows.size()), b;
  for (size_t i = 0; i < rows.size(); i++) s[i] = rows[i].to_fq();
  if (blinds)
    for (auto &x : *blinds) b.push_back(x.to_fq());
  std::vector<Compressed> out(count);
  check(spg_commit_batch(g.ctx, g.dev, s.data(), len, blinds ? b.data() : nullptr, count, (uint8_t *)out.data()),
        "spg_commit_batch");
  std::vector<Point> pts;
  for (auto &c : out) pts.push_back(Point::decompress(c));
  return pts;
}

// ---------------------------------------------------------------- bullet reduction (src/nizk/bullet.rs:32-132)
struct BulletReductionProof {
  std::vector<Compressed> L_vec, R_vec;
  struct Out {
    Scalar a, b, blind_fin;
    Point G;
  };
  static std::pair<BulletReductionProof, Out> prove(ProofTranscript &t, const Point &Q, std::vector<Point> G, const Point &H,
                                                    std::vector<Scalar> a, std::vector<Scalar> b, const Scalar &blind,
                                                    const std::vector<std::pair<Scalar, Scalar>> &blinds) {
    size_t n = G.size();
    BulletReductionProof p;
    Scalar blind_fin = blind;
    size_t round = 0;
    while (n != 1) {
      n /= 2;
      Scalar c_L, c_R;
      for (size_t i = 0; i < n; i++) {
        c_L += a[i] * b[n + i];
        c_R += a[n + i] * b[i];
      }
      const Scalar &blind_L = blinds[round].first, &blind_R = blinds[round].second;
      round++;
      Point L = Q * c_L + H * blind_L, R = Q * c_R + H * blind_R;
      for (size_t i = 0; i < n; i++) {
        if (!(a[i] == Scalar::zero())) L = L + G[n + i] * a[i];
        if (!(a[n + i] == Scalar::zero())) R = R + G[i] * a[n + i];
      }
      Compressed Lc = L.compress(), Rc = R.compress();
      t.append_point("L", Lc);
      t.append_point("R", Rc);
      Scalar u = t.challenge_scalar("u");
      Scalar u_inv = u.invert();
      for (size_t i = 0; i < n; i++) {
        a[i] = a[i] * u + u_inv * a[n + i];
        b[i] = b[i] * u_inv + u * b[n + i];
        G[i] = G[i] * u_inv + G[n + i] * u;
      }
      blind_fin = blind_fin + blind_L * u * u + blind_R * u_inv * u_inv;
      p.L_vec.push_back(Lc);
      p.R_vec.push_back(Rc);
    }
    Out o{a[0], b[0], blind_fin, G[0]};
    return {p, o};
  }
  // Same protocol with the generators never folded on the host. After k rounds the folded
  // generator is G_k[i] = sum over original indices m with (m mod n_k) == i of s_k[m] G[m],
  // s_k[m] = prod_j (bit_j(m) ? u_j : u_j^-1) over the k top bits of m (the fold at :113-118
  // unrolled), so every L, R (and the final G) is a multiscalar multiplication over the
  // ORIGINAL bases, which the device serves from fixed-base tables with no doublings:
  //   L = sum_m [m mod n_k >= n_k/2] a[(m mod n_k) - n_k/2] s[m] G[m] + c_L Q + blind_L H
  //   R = sum_m [m mod n_k <  n_k/2] a[(m mod n_k) + n_k/2] s[m] G[m] + c_R Q + blind_R H
  // Group elements are the same, so the compressed encodings are the reference's.
  // `q_mul(c)` returns c * Q (the caller knows Q = r * G_1 and has a fixed-base table for G_1).
  template <typename QMul>
  static std::pair<BulletReductionProof, Out> prove_device(ProofTranscript &t, QMul q_mul, const DotProductProofGens &gens,
                                                           size_t n, std::vector<Scalar> a, std::vector<Scalar> b,
                                                           const Scalar &blind,
                                                           const std::vector<std::pair<Scalar, Scalar>> &blinds) {
    BulletReductionProof p;
    Scalar blind_fin = blind;
    // the per-base scalars s[m] and the L / R scalar rows (O(n) per round) live on the device
    // (spg_bullet_*), and so do a and b; the host keeps the transcript
    spg_bullet *st = nullptr;
    check(spg_bullet_create(gens.ctx, gens.dev, n, &st), "spg_bullet_create");
    struct Guard {
      spg_bullet *s;
      ~Guard() { spg_bullet_destroy(s); }
    } guard{st};
    // a and b go to the device once; a round's inner products c_L, c_R come back with its L and R, and the
    // fold of a and b rides in the kernel that folds s (spg_bullet_set_ab / _lr_resident / _final_ab)
    {
      std::vector<spg_fq> a_fq(n), b_fq(n);
      for (size_t i = 0; i < n; i++) {
        a_fq[i] = a[i].to_fq();
        b_fq[i] = b[i].to_fq();
      }
      check(spg_bullet_set_ab(st, a_fq.data(), b_fq.data()), "spg_bullet_set_ab");
    }
    size_t nk = n, round = 0;
    static const bool trace = getenv("SPH_TRACE") != nullptr;
    double t_dev = 0, t_grp = 0, t_fold = 0;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
      return std::chrono::duration<double, std::milli>(b - a).count();
    };
    while (nk != 1) {
      size_t nh = nk / 2;
      const Scalar &blind_L = blinds[round].first, &blind_R = blinds[round].second;
      round++;
      spg_fq bl[2] = {blind_L.to_fq(), blind_R.to_fq()}, cc[2];
      uint8_t lr[256];  // the two points in extended coordinates: c Q is added before anything is encoded
      auto t1 = now();
      check(spg_bullet_lr_resident(st, nk, bl, 1, lr, cc), "spg_bullet_lr_resident");
      auto t2 = now();
      Scalar c_L = Scalar::from_fq(cc[0]), c_R = Scalar::from_fq(cc[1]);
      Compressed Lc = (Point::from_ext_bytes(lr) + q_mul(c_L)).compress(), Rc = (Point::from_ext_bytes(lr + 128) + q_mul(c_R)).compress();
      t.append_point("L", Lc);
      t.append_point("R", Rc);
      Scalar u = t.challenge_scalar("u");
      Scalar u_inv = u.invert();
      auto t3 = now();
      spg_fq fu = u.to_fq(), fi = u_inv.to_fq();
      check(spg_bullet_fold(st, nk, &fu, &fi), "spg_bullet_fold");
      blind_fin = blind_fin + blind_L * u * u + blind_R * u_inv * u_inv;
      p.L_vec.push_back(Lc);
      p.R_vec.push_back(Rc);
      nk = nh;
      auto t4 = now();
      t_fold += ms(t3, t4);
      t_dev += ms(t1, t2);
      t_grp += ms(t2, t3);
    }
    Compressed gh;
    spg_fq ab[2];
    auto t5 = now();
    check(spg_bullet_final_ab(st, gh.b, ab), "spg_bullet_final_ab");
    if (trace)
      fprintf(stderr, "[sph]   bullet n=%zu: device L/R + inner products %.3f ms, host group+transcript %.3f ms, fold launches %.3f ms, G_hat %.3f ms\n",
              n, t_dev, t_grp, t_fold, ms(t5, now()));
    Out o{Scalar::from_fq(ab[0]), Scalar::from_fq(ab[1]), blind_fin, Point::decompress(gh)};
    return {p, o};
  }
  void write(Writer &w) const {
    w.points(L_vec);
    w.points(R_vec);
  }
};

// DotProductProofLog (src/nizk/mod.rs:420-523)
struct DotProductProofLog {
  BulletReductionProof bullet;
  Compressed delta, beta;
  Scalar z1, z2;
  static DotProductProofLog prove(const DotProductProofGens &gens, ProofTranscript &t, RandomTape &tape,
                                  const std::vector<Scalar> &x, const Scalar &blind_x, const std::vector<Scalar> &a,
                                  const Scalar &y, const Scalar &blind_y) {
    t.append_protocol_name("dot product proof (log)");
    size_t n = x.size();
    if (gens.n < n) throw std::runtime_error("DotProductProofLog: not enough generators");
    Scalar d = tape.random_scalar("d");
    Scalar r_delta = tape.random_scalar("r_delta");
    Scalar r_beta = tape.random_scalar("r_delta");  // sic: the reference reuses the label (nizk/mod.rs:454)
    size_t lg = log2z(n);
    std::vector<Scalar> v1 = tape.random_vector("blinds_vec_1", 2 * lg), v2 = tape.random_vector("blinds_vec_2", 2 * lg);
    std::vector<std::pair<Scalar, Scalar>> blinds;
    for (size_t i = 0; i < v1.size(); i++) blinds.push_back({v1[i], v2[i]});
    const MultiCommitGens &gn = gens.gens_n;
    bool on_device = gens.dev != nullptr && (n >= 32 || gens.gens_n.G.empty());
    if (on_device) {
      std::vector<Scalar> bl = {blind_x};
      t.append_point("Cx", device_msm(gens, x, n, 1, &bl)[0].compress());
    } else {
      t.append_point("Cx", commit(x, blind_x, gn).compress());
    }
    t.append_point("Cy", commit(y, blind_y, gens.gens_1).compress());
    t.append_scalars("a", a);
    Scalar r = t.challenge_scalar("r");
    MultiCommitGens g1s = gens.gens_1.scale(r);
    Scalar blind_Gamma = blind_x + r * blind_y;
    // Q = r * G_1: with a fixed-base table for G_1, c * Q is the table product (c r) * G_1
    const MultiCommitGens &g1 = gens.gens_1;
    auto q_mul = [&](const Scalar &c) { return g1.tabs ? (*g1.tabs)[0].mul(c * r) : g1s.G[0] * c; };
    auto br = on_device ? BulletReductionProof::prove_device(t, q_mul, gens, n, x, a, blind_Gamma, blinds)
                        : BulletReductionProof::prove(t, g1s.G[0], std::vector<Point>(gn.G.begin(), gn.G.begin() + n), gn.h,
                                                      x, a, blind_Gamma, blinds);
    const auto &o = br.second;
    Scalar y_hat = o.a * o.b;
    DotProductProofLog p;
    p.bullet = br.first;
    MultiCommitGens ghat;
    ghat.n = 1;
    ghat.G = {o.G};
    ghat.h = gens.gens_1.h;
    p.delta = commit(d, r_delta, ghat).compress();
    t.append_point("delta", p.delta);
    p.beta = commit(d, r_beta, g1s).compress();
    t.append_point("beta", p.beta);
    Scalar c = t.challenge_scalar("c");
    p.z1 = d + c * y_hat;
    p.z2 = o.b * (c * o.blind_fin + r_beta) + r_delta;
    return p;
  }
  void write(Writer &w) const {
    bullet.write(w);
    w.point(delta);
    w.point(beta);
    w.scalar(z1);
    w.scalar(z2);
  }
};

// ---------------------------------------------------------------- generators (src/r1csproof.rs:45-80)
struct R1CSGens {
  MultiCommitGens sc_gens_1, sc_gens_3, sc_gens_4;
  DotProductProofGens pc;
  spg_gens *d_pc = nullptr;
  // with a context the opening-proof generators are also uploaded (tables are built on first use)
  R1CSGens(const std::string &label, size_t num_vars, spg_ctx *ctx = nullptr) {
    size_t ell = log2z(num_vars);
    size_t right = ell - ell / 2;
    size_t n = (size_t)1 << right;
    if (ctx && n >= 256) {
      pc = DotProductProofGens::on_device(ctx, n, label, &d_pc);
    } else {
      pc = DotProductProofGens(n, label);
      if (ctx) {
        std::vector<uint8_t> c = pc.gens_n.compressed();
        check(spg_gens_upload(ctx, c.data(), pc.gens_n.n + 1, &d_pc), "spg_gens_upload");
        pc.attach_device(ctx, d_pc);
      }
    }
    sc_gens_3 = MultiCommitGens(3, label);
    sc_gens_4 = MultiCommitGens(4, label);
    pc.gens_1.precompute();
    sc_gens_1 = pc.gens_1;
    sc_gens_3.precompute();
    sc_gens_4.precompute();
  }
  ~R1CSGens() { spg_gens_destroy(d_pc); }
  R1CSGens(const R1CSGens &) = delete;
  R1CSGens &operator=(const R1CSGens &) = delete;
};

// ---------------------------------------------------------------- ZK sumcheck glue
struct ZKSumcheckProof {
  std::vector<Compressed> comm_polys, comm_evals;
  std::vector<DotProductProof> proofs;
  void write(Writer &w) const {
    w.points(comm_polys);
    w.points(comm_evals);
    w.u64(proofs.size());
    for (auto &p : proofs) p.write(w);
  }
};

// One loop for both disjoint-round provers (src/sumcheck.rs:1104-1367 == :816-1054): the only
// difference is where (e0, e2, e3) come from and what is bound, i.e. the two callbacks.
template <typename Eval, typename Bind>
ZKSumcheckProof zk_sumcheck(const Scalar &claim, const Scalar &blind_claim, size_t num_rounds, Eval eval, Bind bind,
                            const MultiCommitGens &g1, const MultiCommitGens &gn, ProofTranscript &t, RandomTape &tape,
                            std::vector<Scalar> *r_out, Scalar *blind_out) {
  std::vector<Scalar> blinds_poly = tape.random_vector("blinds_poly", num_rounds);
  std::vector<Scalar> blinds_evals = tape.random_vector("blinds_evals", num_rounds);
  Scalar claim_per_round = claim;
  Compressed comm_claim_per_round = commit(claim_per_round, blind_claim, g1).compress();
  ZKSumcheckProof pr;
  for (size_t j = 0; j < num_rounds; j++) {
    spg_fq e[3];
    eval(e);
    Scalar e0 = Scalar::from_fq(e[0]);
    UniPoly poly = UniPoly::from_evals({e0, claim_per_round - e0, Scalar::from_fq(e[1]), Scalar::from_fq(e[2])});
    Compressed comm_poly = poly.commit(gn, blinds_poly[j]).compress();
    t.append_point("comm_poly", comm_poly);
    pr.comm_polys.push_back(comm_poly);
    Scalar r_j = t.challenge_scalar("challenge_nextround");
    spg_fq rj = r_j.to_fq();
    bind(&rj);
    Scalar ev = poly.evaluate(r_j);
    Compressed comm_eval = commit(ev, blinds_evals[j], g1).compress();
    t.append_point("comm_claim_per_round", comm_claim_per_round);
    t.append_point("comm_eval", comm_eval);
    std::vector<Scalar> w = t.challenge_vector("combine_two_claims_to_one", 2);
    Scalar target = w[0] * claim_per_round + w[1] * ev;
    const Scalar &blind_sc = j == 0 ? blind_claim : blinds_evals[j - 1];
    Scalar blind = w[0] * blind_sc + w[1] * blinds_evals[j];
    // a = w0 * (2,1,1,1) + w1 * (1, r, r^2, r^3)
    size_t deg = poly.degree();
    std::vector<Scalar> a_sc(deg + 1, Scalar::one()), a_ev(deg + 1, Scalar::one()), a;
    a_sc[0] += Scalar::one();
    for (size_t k = 1; k <= deg; k++) a_ev[k] = a_ev[k - 1] * r_j;
    for (size_t k = 0; k <= deg; k++) a.push_back(w[0] * a_sc[k] + w[1] * a_ev[k]);
    pr.proofs.push_back(DotProductProof::prove(g1, gn, t, tape, poly.coeffs, blinds_poly[j], a, target, blind, &comm_poly));
    claim_per_round = ev;
    comm_claim_per_round = comm_eval;
    r_out->push_back(r_j);
    pr.comm_evals.push_back(comm_eval);
  }
  *blind_out = blinds_evals[num_rounds - 1];
  return pr;
}

// ---------------------------------------------------------------- witnesses
struct WitnessSec {  // ProverWitnessSecInfo
  std::vector<size_t> num_proofs, num_inputs;  // per instance of this section
  spg_witness *dev = nullptr;
};

// ---------------------------------------------------------------- PolyEvalProof (src/dense_mlpoly.rs:861-960)
struct PolyRef {
  spg_witness *w;
  size_t p;
  size_t num_proofs, num_inputs;
};

inline std::vector<Scalar> eq_evals_host(const std::vector<Scalar> &r) {  // EqPolynomial::evals
  std::vector<Scalar> ev((size_t)1 << r.size(), Scalar::one());
  size_t size = 1;
  for (size_t j = 0; j < r.size(); j++) {
    size *= 2;
    for (size_t i = size - 1;; i -= 2) {
      Scalar s = ev[i / 2];
      ev[i] = s * r[j];
      ev[i - 1] = s - ev[i];
      if (i == 1) break;
    }
  }
  return ev;
}

inline std::vector<DotProductProofLog> prove_batched_instances_disjoint_rounds(
    spg_ctx *ctx, const std::vector<PolyRef> &polys, const std::vector<Scalar> &rq, const std::vector<Scalar> &ry,
    const std::vector<Scalar> &Zr, const DotProductProofGens &gens, ProofTranscript &t, RandomTape &tape) {
  Trace tr;
  t.append_protocol_name("polynomial evaluation proof");
  std::map<std::pair<size_t, size_t>, size_t> index_map;
  std::vector<std::vector<Scalar>> LZ_list, L_list, R_list;
  std::vector<Scalar> Zc_list;
  Scalar c_base = t.challenge_scalar("challenge_c");
  Scalar c = Scalar::one();
  auto bound = [&](const PolyRef &pr, const std::vector<Scalar> &L) {
    spg_vec *pv = nullptr, *out = nullptr;
    check(spg_witness_poly(pr.w, pr.p, &pv), "spg_witness_poly");
    std::vector<spg_fq> Lf;
    for (auto &s : L) Lf.push_back(s.to_fq());
    check(spg_dense_bound_L(ctx, pv, Lf.data(), Lf.size(), &out), "spg_dense_bound_L");
    size_t n = spg_vec_len(out);
    std::vector<spg_fq> h(n);
    check(spg_vec_download(ctx, out, 0, n, h.data()), "spg_vec_download");
    spg_vec_free(out);
    std::vector<Scalar> r;
    for (auto &x : h) r.push_back(Scalar::from_fq(x));
    return r;
  };
  for (size_t i = 0; i < polys.size(); i++) {
    auto key = std::make_pair(polys[i].num_proofs, polys[i].num_inputs);
    auto it = index_map.find(key);
    if (it != index_map.end()) {
      c *= c_base;
      size_t idx = it->second;
      std::vector<Scalar> LZ = bound(polys[i], L_list[idx]);
      for (size_t j = 0; j < LZ.size(); j++) LZ_list[idx][j] = LZ_list[idx][j] + c * LZ[j];
      Zc_list[idx] += c * Zr[i];
    } else {
      index_map[key] = LZ_list.size();
      Zc_list.push_back(Zr[i]);
      size_t nvq = log2z(key.first), nvy = log2z(key.second);
      std::vector<Scalar> r(rq.end() - nvq, rq.end());
      if (nvy >= ry.size()) {
        r.insert(r.end(), nvy - ry.size(), Scalar::zero());
        r.insert(r.end(), ry.begin(), ry.end());
      } else {
        r.insert(r.end(), ry.end() - nvy, ry.end());
      }
      size_t left = r.size() / 2;
      std::vector<Scalar> L = eq_evals_host(std::vector<Scalar>(r.begin(), r.begin() + left));
      std::vector<Scalar> R = eq_evals_host(std::vector<Scalar>(r.begin() + left, r.end()));
      LZ_list.push_back(bound(polys[i], L));
      L_list.push_back(L);
      R_list.push_back(R);
    }
  }
  tr.lap("  openings: bound(L) + eq tables");
  std::vector<DotProductProofLog> proofs;
  for (size_t i = 0; i < LZ_list.size(); i++)
    proofs.push_back(DotProductProofLog::prove(gens, t, tape, LZ_list[i], Scalar::zero(), R_list[i], Zc_list[i], Scalar::zero()));
  tr.lap("  openings: dot-product proofs");
  return proofs;
}

// ---------------------------------------------------------------- the other opening variants
// PolyEvalProof::prove_batched_points / prove_batched_instances / prove_uni_batched_instances
// (src/dense_mlpoly.rs:531-622, 689-780, 1046-1130; called from src/lib.rs:2587, 2657, 2673), zero blinds as at
// every call site. The polynomials stay on the device: `poly.bound(&L)` is spg_dense_bound_L, the dot-product
// proofs' MSMs run there as well; the grouping, the random linear combinations and the transcript are host work.
inline std::vector<Scalar> bound_L_dev(spg_ctx *ctx, const spg_vec *pv, const std::vector<Scalar> &L) {
  spg_vec *out = nullptr;
  std::vector<spg_fq> Lf;
  for (auto &s : L) Lf.push_back(s.to_fq());
  check(spg_dense_bound_L(ctx, pv, Lf.data(), Lf.size(), &out), "spg_dense_bound_L");
  size_t n = spg_vec_len(out);
  std::vector<spg_fq> h(n);
  check(spg_vec_download(ctx, out, 0, n, h.data()), "spg_vec_download");
  spg_vec_free(out);
  std::vector<Scalar> r;
  for (auto &x : h) r.push_back(Scalar::from_fq(x));
  return r;
}
// EqPolynomial::compute_factored_evals (src/dense_mlpoly.rs:122-130)
inline std::pair<std::vector<Scalar>, std::vector<Scalar>> factored_evals(const std::vector<Scalar> &r) {
  size_t left = r.size() / 2;
  return {eq_evals_host(std::vector<Scalar>(r.begin(), r.begin() + left)), eq_evals_host(std::vector<Scalar>(r.begin() + left, r.end()))};
}
inline std::vector<uint64_t> limbs_key(const std::vector<Scalar> &v, size_t from, size_t to) {
  std::vector<uint64_t> k;
  for (size_t i = from; i < to; i++) {
    spg_fq f = v[i].to_fq();
    k.insert(k.end(), f.l, f.l + 4);
  }
  return k;
}

inline std::vector<DotProductProofLog> prove_batched_points(spg_ctx *ctx, const spg_vec *poly,
                                                            const std::vector<std::vector<Scalar>> &r_list,
                                                            const std::vector<Scalar> &Zr_list, const DotProductProofGens &gens,
                                                            ProofTranscript &t, RandomTape &tape) {
  t.append_protocol_name("polynomial evaluation proof");
  size_t left = r_list.at(0).size() / 2;
  std::map<std::vector<uint64_t>, size_t> index_map;
  std::vector<std::vector<Scalar>> L_list, R_list;
  std::vector<Scalar> Zc_list;
  Scalar c_base = t.challenge_scalar("challenge_c");
  Scalar c = Scalar::one();
  for (size_t i = 0; i < r_list.size(); i++) {
    auto LR = factored_evals(r_list[i]);
    auto key = limbs_key(r_list[i], 0, left);
    auto it = index_map.find(key);
    if (it != index_map.end()) {
      c *= c_base;
      size_t idx = it->second;
      for (size_t j = 0; j < LR.second.size(); j++) R_list[idx][j] = R_list[idx][j] + c * LR.second[j];
      Zc_list[idx] += c * Zr_list[i];
    } else {
      index_map[key] = L_list.size();
      L_list.push_back(LR.first);
      R_list.push_back(LR.second);
      Zc_list.push_back(Zr_list[i]);
    }
  }
  std::vector<DotProductProofLog> proofs;
  for (size_t i = 0; i < L_list.size(); i++) {
    std::vector<Scalar> LZ = bound_L_dev(ctx, poly, L_list[i]);
    proofs.push_back(DotProductProofLog::prove(gens, t, tape, LZ, Scalar::zero(), R_list[i], Zc_list[i], Scalar::zero()));
  }
  return proofs;
}

inline std::vector<DotProductProofLog> prove_batched_instances(spg_ctx *ctx, const std::vector<const spg_vec *> &polys,
                                                               const std::vector<std::vector<Scalar>> &r_list,
                                                               const std::vector<Scalar> &Zr_list, const DotProductProofGens &gens,
                                                               ProofTranscript &t, RandomTape &tape) {
  t.append_protocol_name("polynomial evaluation proof");
  std::map<std::pair<size_t, std::vector<uint64_t>>, size_t> index_map;
  std::vector<std::vector<Scalar>> LZ_list, R_list;
  std::vector<Scalar> Zc_list;
  Scalar c_base = t.challenge_scalar("challenge_c");
  Scalar c = Scalar::one();
  for (size_t i = 0; i < polys.size(); i++) {
    size_t nv = log2z(spg_vec_len(polys[i]));
    std::vector<Scalar> r;
    if (nv >= r_list[i].size()) {
      r.assign(nv - r_list[i].size(), Scalar::zero());
      r.insert(r.end(), r_list[i].begin(), r_list[i].end());
    } else {
      r.assign(r_list[i].end() - nv, r_list[i].end());
    }
    auto LR = factored_evals(r);
    auto key = std::make_pair(nv, limbs_key(LR.second, 0, LR.second.size()));
    std::vector<Scalar> LZ = bound_L_dev(ctx, polys[i], LR.first);
    auto it = index_map.find(key);
    if (it != index_map.end()) {
      c *= c_base;
      size_t idx = it->second;
      for (size_t j = 0; j < LZ.size(); j++) LZ_list[idx][j] = LZ_list[idx][j] + c * LZ[j];
      Zc_list[idx] += c * Zr_list[i];
    } else {
      index_map[key] = LZ_list.size();
      Zc_list.push_back(Zr_list[i]);
      R_list.push_back(LR.second);
      LZ_list.push_back(LZ);
    }
  }
  std::vector<DotProductProofLog> proofs;
  for (size_t i = 0; i < LZ_list.size(); i++)
    proofs.push_back(DotProductProofLog::prove(gens, t, tape, LZ_list[i], Scalar::zero(), R_list[i], Zc_list[i], Scalar::zero()));
  return proofs;
}

inline std::pair<DotProductProofLog, Compressed> prove_uni_batched_instances(spg_ctx *ctx, const std::vector<const spg_vec *> &polys,
                                                                            const Scalar &r, const std::vector<Scalar> &Zr,
                                                                            const DotProductProofGens &gens, ProofTranscript &t,
                                                                            RandomTape &tape) {
  t.append_protocol_name("polynomial evaluation proof");
  size_t max_nv = 0;
  for (auto *p : polys) max_nv = std::max(max_nv, log2z(spg_vec_len(p)));
  size_t R_size = (size_t)1 << (max_nv - max_nv / 2);
  std::vector<Scalar> R;
  Scalar rb = Scalar::one();
  for (size_t i = 0; i < R_size; i++) {
    R.push_back(rb);
    rb *= r;
  }
  std::map<size_t, std::vector<Scalar>> L_map;
  Scalar c_base = t.challenge_scalar("challenge_c");
  Scalar c = Scalar::one();
  std::vector<Scalar> LZ_comb(R_size, Scalar::zero());
  Scalar Zr_comb = Scalar::zero();
  for (size_t i = 0; i < polys.size(); i++) {
    size_t nv = log2z(spg_vec_len(polys[i]));
    if (!L_map.count(nv)) {
      size_t Ls = (size_t)1 << (nv / 2), Rs = (size_t)1 << (nv - nv / 2);
      Scalar r_base = Scalar::one();
      for (size_t k = 0; k < Rs; k++) r_base *= r;
      std::vector<Scalar> L;
      Scalar lb = Scalar::one();
      for (size_t k = 0; k < Ls; k++) {
        L.push_back(lb);
        lb *= r_base;
      }
      L_map[nv] = L;
    }
    std::vector<Scalar> LZ = bound_L_dev(ctx, polys[i], L_map[nv]);
    for (size_t j = 0; j < LZ.size() && j < R_size; j++) LZ_comb[j] = LZ_comb[j] + c * LZ[j];
    Zr_comb += c * Zr[i];
    c *= c_base;
  }
  DotProductProofLog pr = DotProductProofLog::prove(gens, t, tape, LZ_comb, Scalar::zero(), R, Zr_comb, Scalar::zero());
  // C_Zr_prime = the Cy the dot-product proof committed to (src/nizk/mod.rs:470)
  Compressed Cy = commit(Zr_comb, Scalar::zero(), gens.gens_1).compress();
  return {pr, Cy};
}

// ---------------------------------------------------------------- R1CSProof::prove (src/r1csproof.rs:210-685)
struct R1CSProofOut {
  std::vector<uint8_t> bytes;                 // bincode layout of R1CSProof (src/r1csproof.rs:25-43)
  std::vector<Scalar> rp, rq_rev, rx, rw_ry;  // the returned challenge vectors (:683)
};

inline R1CSProofOut r1cs_prove(spg_ctx *ctx, size_t num_instances, size_t max_num_proofs,
                               const std::vector<size_t> &num_proofs, size_t max_num_inputs,
                               const std::vector<size_t> &num_inputs, const std::vector<WitnessSec> &secs,
                               const spg_r1cs *inst, size_t inst_num_instances, size_t inst_max_num_cons,
                               const std::vector<size_t> &inst_num_cons, const R1CSGens &gens, ProofTranscript &t,
                               RandomTape &tape) {
  Trace tr;
  t.append_protocol_name("R1CS proof");
  size_t W = secs.size();
  size_t P = num_instances;
  size_t num_cons = inst_max_num_cons;
  std::vector<size_t> block_num_cons = inst_num_instances == 1 ? std::vector<size_t>(P, inst_num_cons[0]) : inst_num_cons;
  // z_mat on the device (:278-293)
  std::vector<spg_witness *> wptr;
  for (auto &s : secs) wptr.push_back(s.dev);
  spg_zmat *z = nullptr;
  check(spg_zmat_build(ctx, P, num_proofs.data(), num_inputs.data(), W, wptr.data(), &z), "spg_zmat_build");
  size_t nrp = log2z(next_pow2(P)), nrq = log2z(max_num_proofs), nrx = log2z(num_cons), nrw = log2z(next_pow2(W)),
         nry = log2z(max_num_inputs);
  std::vector<Scalar> tau_p = t.challenge_vector("challenge_tau_p", nrp);
  std::vector<Scalar> tau_q = t.challenge_vector("challenge_tau_q", nrq);
  std::vector<Scalar> tau_x = t.challenge_vector("challenge_tau_x", nrx);
  auto fqv = [](const std::vector<Scalar> &v) {
    std::vector<spg_fq> o;
    for (auto &s : v) o.push_back(s.to_fq());
    if (o.empty()) o.push_back(spg_fq{{0, 0, 0, 0}});
    return o;
  };
  // PHASE 1 (:313-343)
  spg_sc1 *sc1 = nullptr;
  check(spg_sc1_create(ctx, inst, z, P, num_proofs.data(), max_num_proofs, block_num_cons.data(), num_cons,
                       max_num_inputs, fqv(tau_p).data(), fqv(tau_q).data(), fqv(tau_x).data(), &sc1),
        "spg_sc1_create");
  // claim_phase1 = 0 (:330): the prover's witness satisfies the instance row by row, which also lets the
  // fused SpMV + first round skip the evaluation at 0
  check(spg_sc1_set_satisfied(sc1), "spg_sc1_set_satisfied");
  tr.lap("z_mat + SpMV + sc1 setup");
  std::vector<Scalar> r1;
  Scalar blind_claim_postsc1;
  ZKSumcheckProof sc_proof_phase1 = zk_sumcheck(
      Scalar::zero(), Scalar::zero(), nrx + nrq + nrp, [&](spg_fq *e) { check(spg_sc1_round_eval(sc1, e), "spg_sc1_round_eval"); },
      [&](const spg_fq *r) { check(spg_sc1_round_bind(sc1, r), "spg_sc1_round_bind"); }, gens.sc_gens_1, gens.sc_gens_4, t,
      tape, &r1, &blind_claim_postsc1);
  tr.lap("phase-1 rounds (zk)");
  spg_fq claims1[4];
  check(spg_sc1_final(sc1, claims1), "spg_sc1_final");
  spg_sc1_destroy(sc1);
  Scalar tau_claim = Scalar::from_fq(claims1[0]), Az_claim = Scalar::from_fq(claims1[1]),
         Bz_claim = Scalar::from_fq(claims1[2]), Cz_claim = Scalar::from_fq(claims1[3]);
  Scalar Az_blind = tape.random_scalar("Az_blind"), Bz_blind = tape.random_scalar("Bz_blind"),
         Cz_blind = tape.random_scalar("Cz_blind"), prod_Az_Bz_blind = tape.random_scalar("prod_Az_Bz_blind");
  auto pok = KnowledgeProof::prove(gens.sc_gens_1, t, tape, Cz_claim, Cz_blind);
  Compressed comm_Cz_claim = pok.second;
  Scalar prod = Az_claim * Bz_claim;
  auto pp = ProductProof::prove(gens.sc_gens_1, t, tape, Az_claim, Az_blind, Bz_claim, Bz_blind, prod, prod_Az_Bz_blind);
  Compressed comm_Az_claim = pp.second.X, comm_Bz_claim = pp.second.Y, comm_prod = pp.second.Z;
  t.append_point("comm_Az_claim", comm_Az_claim);
  t.append_point("comm_Bz_claim", comm_Bz_claim);
  t.append_point("comm_Cz_claim", comm_Cz_claim);
  t.append_point("comm_prod_Az_Bz_claims", comm_prod);
  Scalar blind_expected_claim_postsc1 = tau_claim * (prod_Az_Bz_blind - Cz_blind);
  Scalar claim_post_phase1 = (Az_claim * Bz_claim - Cz_claim) * tau_claim;
  EqualityProof proof_eq_sc_phase1 = EqualityProof::prove(gens.sc_gens_1, t, tape, claim_post_phase1,
                                                          blind_expected_claim_postsc1, claim_post_phase1, blind_claim_postsc1);
  // split r1 into rx_rev | rq_rev | rp (:410-416)
  std::vector<Scalar> rx_rev(r1.begin(), r1.begin() + nrx), rq_rev(r1.begin() + nrx, r1.begin() + nrx + nrq),
      rp(r1.begin() + nrx + nrq, r1.end());
  std::vector<Scalar> rx(rx_rev.rbegin(), rx_rev.rend()), rq(rq_rev.rbegin(), rq_rev.rend());
  // PHASE 2 (:421-501)
  Scalar r_A = t.challenge_scalar("challenge_Az"), r_B = t.challenge_scalar("challenge_Bz"), r_C = t.challenge_scalar("challenge_Cz");
  Scalar claim_phase2 = r_A * Az_claim + r_B * Bz_claim + r_C * Cz_claim;
  Scalar blind_claim_phase2 = r_A * Az_blind + r_B * Bz_blind + r_C * Cz_blind;
  tr.lap("sigma protocols");
  spg_sc2 *sc2 = nullptr;
  spg_fq fA = r_A.to_fq(), fB = r_B.to_fq(), fC = r_C.to_fq();
  check(spg_sc2_create(ctx, inst, z, P, num_proofs.data(), max_num_proofs, num_inputs.data(), max_num_inputs, W,
                       fqv(rx).data(), fqv(rq_rev).data(), fqv(rp).data(), &fA, &fB, &fC, &sc2),
        "spg_sc2_create");
  tr.lap("sc2 setup (ABC, Z bind)");
  std::vector<Scalar> r2;
  Scalar blind_claim_postsc2;
  ZKSumcheckProof sc_proof_phase2 = zk_sumcheck(
      claim_phase2, blind_claim_phase2, nry + nrw + nrp, [&](spg_fq *e) { check(spg_sc2_round_eval(sc2, e), "spg_sc2_round_eval"); },
      [&](const spg_fq *r) { check(spg_sc2_round_bind(sc2, r), "spg_sc2_round_bind"); }, gens.sc_gens_1, gens.sc_gens_4, t,
      tape, &r2, &blind_claim_postsc2);
  tr.lap("phase-2 rounds (zk)");
  spg_fq claims2[3];
  check(spg_sc2_final(sc2, claims2), "spg_sc2_final");
  spg_sc2_destroy(sc2);
  spg_zmat_destroy(z);
  std::vector<Scalar> ry_rev(r2.begin(), r2.begin() + nry), rw(r2.begin() + nry, r2.begin() + nry + nrw),
      rp2(r2.begin() + nry + nrw, r2.end());
  std::vector<Scalar> ry(ry_rev.rbegin(), ry_rev.rend());
  // POLYEVAL (:518-586)
  std::vector<Scalar> ry_factors(nry + 1, Scalar::one());
  for (size_t i = 0; i < nry; i++) ry_factors[i + 1] = ry_factors[i] * (Scalar::one() - ry[i]);
  std::vector<PolyRef> poly_list;
  std::vector<Scalar> Zr_list;
  std::vector<std::vector<Scalar>> eval_vars_at_ry_list(W);
  std::vector<std::vector<Compressed>> comm_vars_at_ry_list(W);
  for (size_t i = 0; i < W; i++) {
    const WitnessSec &w = secs[i];
    for (size_t p = 0; p < w.num_proofs.size(); p++) {
      size_t wq = w.num_proofs[p], wy = w.num_inputs[p];
      poly_list.push_back(PolyRef{w.dev, p, wq, wy});
      std::vector<Scalar> r(rq.end() - log2z(wq), rq.end());
      if (wy >= max_num_inputs) {
        r.insert(r.end(), log2z(wy) - log2z(max_num_inputs), Scalar::zero());
        r.insert(r.end(), ry.begin(), ry.end());
      } else {
        r.insert(r.end(), ry.end() - log2z(wy), ry.end());
      }
      spg_vec *pv = nullptr;
      check(spg_witness_poly(w.dev, p, &pv), "spg_witness_poly");
      spg_fq ev;
      check(spg_dense_evaluate(ctx, pv, fqv(r).data(), r.size(), &ev), "spg_dense_evaluate");
      Scalar e = Scalar::from_fq(ev);
      Zr_list.push_back(e);
      eval_vars_at_ry_list[i].push_back(wy >= max_num_inputs ? e : e * ry_factors[nry - log2z(wy)]);
      comm_vars_at_ry_list[i].push_back(commit(e, Scalar::zero(), gens.pc.gens_1).compress());
    }
  }
  tr.lap("witness evaluations");
  std::vector<DotProductProofLog> proof_eval_vars =
      prove_batched_instances_disjoint_rounds(ctx, poly_list, rq, ry, Zr_list, gens.pc, t, tape);
  tr.lap("opening proofs");
  // combine per instance (:588-638)
  size_t Wp = next_pow2(W);
  if (Wp > 8) throw std::runtime_error("Unsupported num_witness_secs");  // the reference panics (:629-631)
  std::vector<Scalar> prefix(Wp);
  for (size_t k = 0; k < Wp; k++) {
    Scalar acc = Scalar::one();
    for (size_t b = 0; b < nrw; b++) acc = acc * (((k >> (nrw - 1 - b)) & 1) ? rw[b] : Scalar::one() - rw[b]);
    prefix[k] = acc;
  }
  std::vector<Scalar> eval_vars_comb_list;
  for (size_t p = 0; p < P; p++) {
    Scalar comb;
    for (size_t i = 0; i < W; i++) {
      size_t wp = secs[i].num_proofs.size() == 1 ? 0 : p;
      comb = comb + prefix[i] * eval_vars_at_ry_list[i][wp];
    }
    for (size_t q = 0; q < nrq - log2z(num_proofs[p]); q++) comb *= Scalar::one() - rq[q];
    eval_vars_comb_list.push_back(comb);
  }
  // poly_vars.evaluate(rp): DensePolynomial::new pads with zeros (:641-642)
  eval_vars_comb_list.resize(next_pow2(P), Scalar::zero());
  std::vector<Scalar> eqp = eq_evals_host(rp2);
  Scalar eval_vars_at_ry;
  for (size_t i = 0; i < eqp.size(); i++) eval_vars_at_ry += eqp[i] * eval_vars_comb_list[i];
  Compressed comm_vars_at_ry = commit(eval_vars_at_ry, Scalar::zero(), gens.pc.gens_1).compress();
  Scalar claim_post_phase2 = Scalar::from_fq(claims2[0]) * Scalar::from_fq(claims2[1]) * Scalar::from_fq(claims2[2]);
  EqualityProof proof_eq_sc_phase2 =
      EqualityProof::prove(gens.pc.gens_1, t, tape, claim_post_phase2, Scalar::zero(), claim_post_phase2, blind_claim_postsc2);
  // serialize in field order (:25-43)
  Writer w;
  sc_proof_phase1.write(w);
  w.point(comm_Az_claim);
  w.point(comm_Bz_claim);
  w.point(comm_Cz_claim);
  w.point(comm_prod);
  pok.first.write(w);
  pp.first.write(w);
  proof_eq_sc_phase1.write(w);
  sc_proof_phase2.write(w);
  // the reference pre-sizes the list to W and then pushes one more empty Vec per section
  // (src/r1csproof.rs:532-538), so the serialized list has 2W entries, the last W empty
  w.u64(2 * comm_vars_at_ry_list.size());
  for (auto &v : comm_vars_at_ry_list) w.points(v);
  for (size_t i = 0; i < comm_vars_at_ry_list.size(); i++) w.u64(0);
  w.point(comm_vars_at_ry);
  w.u64(proof_eval_vars.size());
  for (auto &p : proof_eval_vars) p.write(w);
  proof_eq_sc_phase2.write(w);
  tr.lap("tail + serialization");
  R1CSProofOut out;
  out.bytes = w.out;
  out.rp = rp2;
  out.rq_rev = rq_rev;
  out.rx = rx;
  out.rw_ry = rw;
  out.rw_ry.insert(out.rw_ry.end(), ry.begin(), ry.end());
  return out;
}

}  // namespace sph
