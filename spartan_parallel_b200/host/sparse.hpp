// C++ mirror of the transcript side of the sparse-polynomial evaluation proof, with every
// table (lookups, hashed multisets, product circuits, cubic sumchecks, openings' L*Z
// products, commitments) on the device through the C ABI.
//   reference: src/sparse_mlpoly.rs:283-316 (gens), 566-586 (multi_commit), 1104-1263
//              (ProductLayerProof::prove), 805-918 (HashLayerProof::prove), 1509-1564
//              (SparseMatPolyEvalProof::prove); src/product_tree.rs:260-384;
//              src/sumcheck.rs:264-434; src/dense_mlpoly.rs:437-486
#pragma once
#include "protocol.hpp"

namespace sph {

// owns device vectors for the duration of a proof
struct VecPool {
  std::vector<spg_vec *> vs;
  spg_vec *keep(spg_vec *v) {
    vs.push_back(v);
    return v;
  }
  ~VecPool() {
    for (auto v : vs) spg_vec_free(v);
  }
};

inline std::vector<spg_fq> to_fqv(const std::vector<Scalar> &v) {
  std::vector<spg_fq> o;
  for (auto &s : v) o.push_back(s.to_fq());
  if (o.empty()) o.push_back(spg_fq{{0, 0, 0, 0}});
  return o;
}

// SparseMatPolyCommitmentGens::new (:289-316), with the bases also resident on the device
struct SparseGens {
  DotProductProofGens ops, mem, derefs;
  spg_gens *d_ops = nullptr, *d_mem = nullptr, *d_derefs = nullptr;
  size_t nv_ops = 0, nv_mem = 0, nv_derefs = 0;  // variables of the polynomials committed under each
  SparseGens(spg_ctx *ctx, const std::string &label, size_t nvx, size_t nvy, size_t num_nz, size_t batch) {
    auto pcg = [&](size_t nv, spg_gens **d) {
      size_t n = (size_t)1 << (nv - nv / 2);
      if (n >= 256) return DotProductProofGens::on_device(ctx, n, label, d);
      DotProductProofGens g(n, label);
      std::vector<uint8_t> c = g.gens_n.compressed();
      check(spg_gens_upload(ctx, c.data(), g.gens_n.n + 1, d), "spg_gens_upload");
      g.attach_device(ctx, *d);
      return g;
    };
    size_t lg = log2z(next_pow2(num_nz));
    nv_ops = lg + log2z(next_pow2(batch * 5));
    nv_mem = (nvx > nvy ? nvx : nvy) + 1;
    nv_derefs = lg + log2z(next_pow2(batch * 2));
    ops = pcg(nv_ops, &d_ops);
    mem = pcg(nv_mem, &d_mem);
    derefs = pcg(nv_derefs, &d_derefs);
    ops.gens_1.precompute();
    mem.gens_1.precompute();
    derefs.gens_1.precompute();
  }
  ~SparseGens() {
    spg_gens_destroy(d_ops);
    spg_gens_destroy(d_mem);
    spg_gens_destroy(d_derefs);
  }
  SparseGens(const SparseGens &) = delete;
  SparseGens &operator=(const SparseGens &) = delete;
};

// DensePolynomial::commit with zero blinds (src/dense_mlpoly.rs:199-239)
inline std::vector<Compressed> poly_commit_dev(spg_ctx *ctx, const spg_gens *g, const spg_vec *poly) {
  size_t ell = log2z(spg_vec_len(poly));
  size_t L = (size_t)1 << (ell / 2);
  std::vector<Compressed> out(L);
  check(spg_poly_commit(ctx, g, poly, L, (uint8_t *)out.data()), "spg_poly_commit");
  return out;
}

inline void append_poly_commitment(ProofTranscript &t, const std::string &label, const std::vector<Compressed> &C) {
  t.append_message(label, "poly_commitment_begin");
  for (auto &c : C) t.append_point("poly_commitment_share", c);
  t.append_message(label, "poly_commitment_end");
}

// SparseMatPolynomial::multi_commit (:566-586): comm_comb_ops then comm_comb_mem
struct SparseCommitment {
  size_t batch_size, num_ops, num_mem_cells;
  std::vector<Compressed> comm_comb_ops, comm_comb_mem;
};

inline SparseCommitment sparse_commit(spg_ctx *ctx, spg_sparse *sp, size_t batch, const SparseGens &gens) {
  VecPool pool;
  spg_vec *ops = nullptr, *mem = nullptr;
  check(spg_sparse_view(sp, SPG_SPARSE_COMB_OPS, 0, &ops), "spg_sparse_view");
  pool.keep(ops);
  check(spg_sparse_view(sp, SPG_SPARSE_COMB_MEM, 0, &mem), "spg_sparse_view");
  pool.keep(mem);
  SparseCommitment c;
  c.batch_size = batch;
  c.num_ops = spg_sparse_num_ops(sp);
  c.num_mem_cells = spg_sparse_num_mem_cells(sp);
  c.comm_comb_ops = poly_commit_dev(ctx, gens.d_ops, ops);
  c.comm_comb_mem = poly_commit_dev(ctx, gens.d_mem, mem);
  return c;
}

// PolyEvalProof::prove with no blinds (src/dense_mlpoly.rs:437-486)
inline DotProductProofLog polyeval_prove(spg_ctx *ctx, const spg_vec *poly, const std::vector<Scalar> &r, const Scalar &Zr,
                                         const DotProductProofGens &gens, ProofTranscript &t, RandomTape &tape) {
  t.append_protocol_name("polynomial evaluation proof");
  if (((size_t)1 << r.size()) != spg_vec_len(poly)) throw std::runtime_error("polyeval_prove: point / polynomial size mismatch");
  size_t left = r.size() / 2;
  std::vector<Scalar> L = eq_evals_host(std::vector<Scalar>(r.begin(), r.begin() + left));
  std::vector<Scalar> R = eq_evals_host(std::vector<Scalar>(r.begin() + left, r.end()));
  spg_vec *lz = nullptr;
  check(spg_dense_bound_L(ctx, poly, to_fqv(L).data(), L.size(), &lz), "spg_dense_bound_L");
  std::vector<spg_fq> h(R.size());
  int rc = spg_vec_download(ctx, lz, 0, R.size(), h.data());
  spg_vec_free(lz);
  check(rc, "spg_vec_download");
  std::vector<Scalar> LZ;
  for (auto &x : h) LZ.push_back(Scalar::from_fq(x));
  return DotProductProofLog::prove(gens, t, tape, LZ, Scalar::zero(), R, Zr, Scalar::zero());
}

// ---------------------------------------------------------------- batched product-circuit proof
struct LayerProofBatched {
  std::vector<std::vector<Scalar>> polys;  // CompressedUniPoly: coefficients without the linear term
  std::vector<Scalar> claims_prod_left, claims_prod_right;
};
struct ProductCircuitEvalProofBatched {
  std::vector<LayerProofBatched> proof;
  std::vector<Scalar> dotp_left, dotp_right, dotp_weight;
  void write(Writer &w) const {
    w.u64(proof.size());
    for (auto &l : proof) {
      w.u64(l.polys.size());
      for (auto &p : l.polys) w.scalars(p);
      w.scalars(l.claims_prod_left);
      w.scalars(l.claims_prod_right);
    }
    w.scalars(dotp_left);
    w.scalars(dotp_right);
    w.scalars(dotp_weight);
  }
};

inline void append_unipoly(ProofTranscript &t, const UniPoly &p) {  // src/unipoly.rs:112-120
  t.append_message("poly", "UniPoly_begin");
  for (auto &c : p.coeffs) t.append_scalar("coeff", c);
  t.append_message("poly", "UniPoly_end");
}

struct DotpCircuit {  // DotProductCircuit (src/product_tree.rs:66-107); tables are consumed
  spg_vec *left, *right, *weight;
  Scalar eval;
};

// ProductCircuitEvalProofBatched::prove (src/product_tree.rs:260-384)
inline ProductCircuitEvalProofBatched pcepb_prove(spg_ctx *ctx, const std::vector<spg_prodtree *> &trees,
                                                  const std::vector<Scalar> &tree_evals, const std::vector<DotpCircuit> &dotp,
                                                  ProofTranscript &t, std::vector<Scalar> *rand_out) {
  ProductCircuitEvalProofBatched out;
  size_t num_layers = spg_prodtree_num_layers(trees[0]);
  std::vector<Scalar> claims_to_verify = tree_evals, rand;
  for (size_t lid = num_layers; lid-- > 0;) {
    VecPool pool;
    spg_vec *C_par = nullptr;
    if (rand.empty()) {
      spg_fq one = Scalar::one().to_fq();
      check(spg_vec_upload(ctx, &one, 1, &C_par), "spg_vec_upload");
    } else {
      check(spg_eq_evals(ctx, to_fqv(rand).data(), rand.size(), &C_par), "spg_eq_evals");
    }
    pool.keep(C_par);
    size_t num_rounds = rand.size();
    std::vector<spg_vec *> A_par, B_par, A_seq, B_seq, C_seq;
    for (auto tr : trees) {
      spg_vec *l = nullptr, *r = nullptr;
      check(spg_prodtree_layer(tr, lid, &l, &r), "spg_prodtree_layer");
      if (spg_vec_len(l) != spg_vec_len(C_par)) throw std::runtime_error("pcepb_prove: layer / eq table length mismatch");
      A_par.push_back(l);
      B_par.push_back(r);
    }
    bool with_dotp = lid == 0 && !dotp.empty();
    if (with_dotp)
      for (auto &d : dotp) {
        claims_to_verify.push_back(d.eval);
        A_seq.push_back(d.left);
        B_seq.push_back(d.right);
        C_seq.push_back(d.weight);
      }
    std::vector<Scalar> coeff = t.challenge_vector("rand_coeffs_next_layer", claims_to_verify.size());
    Scalar claim;
    for (size_t i = 0; i < coeff.size(); i++) claim += claims_to_verify[i] * coeff[i];
    // SumcheckInstanceProof::prove_cubic_batched (src/sumcheck.rs:264-434)
    spg_cubic *cub = nullptr;
    check(spg_cubic_create(ctx, A_par.size(), A_par.data(), B_par.data(), C_par, A_seq.size(), A_seq.data(), B_seq.data(),
                           C_seq.data(), to_fqv(coeff).data(), &cub),
          "spg_cubic_create");
    LayerProofBatched lay;
    std::vector<Scalar> rand_prod;
    Scalar e = claim;
    size_t nt = A_par.size() + A_seq.size();
    std::vector<spg_fq> fin(2 * A_par.size() + 1 + 3 * A_seq.size());
    try {
      for (size_t j = 0; j < num_rounds; j++) {
        spg_fq ev[3];
        check(spg_cubic_round_eval(cub, ev), "spg_cubic_round_eval");
        Scalar e0 = Scalar::from_fq(ev[0]);
        UniPoly poly = UniPoly::from_evals({e0, e - e0, Scalar::from_fq(ev[1]), Scalar::from_fq(ev[2])});
        append_unipoly(t, poly);
        Scalar r_j = t.challenge_scalar("challenge_nextround");
        rand_prod.push_back(r_j);
        spg_fq rj = r_j.to_fq();
        check(spg_cubic_round_bind(cub, &rj), "spg_cubic_round_bind");
        e = poly.evaluate(r_j);
        lay.polys.push_back({poly.coeffs[0], poly.coeffs[2], poly.coeffs[3]});
      }
      check(spg_cubic_final(cub, fin.data()), "spg_cubic_final");
    } catch (...) {
      spg_cubic_destroy(cub);
      throw;
    }
    spg_cubic_destroy(cub);
    (void)nt;
    size_t np = A_par.size(), ns = A_seq.size();
    for (size_t i = 0; i < np; i++) {
      lay.claims_prod_left.push_back(Scalar::from_fq(fin[i]));
      lay.claims_prod_right.push_back(Scalar::from_fq(fin[np + i]));
    }
    for (size_t i = 0; i < np; i++) {
      t.append_scalar("claim_prod_left", lay.claims_prod_left[i]);
      t.append_scalar("claim_prod_right", lay.claims_prod_right[i]);
    }
    if (with_dotp) {
      size_t o = 2 * np + 1;
      for (size_t i = 0; i < ns; i++) {
        out.dotp_left.push_back(Scalar::from_fq(fin[o + i]));
        out.dotp_right.push_back(Scalar::from_fq(fin[o + ns + i]));
        out.dotp_weight.push_back(Scalar::from_fq(fin[o + 2 * ns + i]));
      }
      for (size_t i = 0; i < ns; i++) {
        t.append_scalar("claim_dotp_left", out.dotp_left[i]);
        t.append_scalar("claim_dotp_right", out.dotp_right[i]);
        t.append_scalar("claim_dotp_weight", out.dotp_weight[i]);
      }
    }
    Scalar r_layer = t.challenge_scalar("challenge_r_layer");
    claims_to_verify.clear();
    for (size_t i = 0; i < np; i++)
      claims_to_verify.push_back(lay.claims_prod_left[i] + r_layer * (lay.claims_prod_right[i] - lay.claims_prod_left[i]));
    rand.assign(1, r_layer);
    rand.insert(rand.end(), rand_prod.begin(), rand_prod.end());
    out.proof.push_back(lay);
  }
  *rand_out = rand;
  return out;
}

// ---------------------------------------------------------------- SparseMatPolyEvalProof::prove
struct TreeSet {
  std::vector<spg_prodtree *> all;
  ~TreeSet() {
    for (auto t : all) spg_prodtree_destroy(t);
  }
  spg_prodtree *build(spg_ctx *ctx, spg_vec *leaves) {  // takes ownership of the leaves
    spg_prodtree *t = nullptr;
    int rc = spg_prodtree_build(ctx, leaves, &t);
    spg_vec_free(leaves);
    check(rc, "spg_prodtree_build");
    all.push_back(t);
    return t;
  }
};

inline std::vector<uint8_t> sparse_prove(spg_ctx *ctx, spg_sparse *sp, size_t batch, const std::vector<Scalar> &rx,
                                         const std::vector<Scalar> &ry, const std::vector<Scalar> &evals,
                                         const SparseGens &gens, ProofTranscript &t, RandomTape &tape) {
  Trace tr;
  t.append_protocol_name("Sparse polynomial evaluation proof");
  if (evals.size() != batch) throw std::runtime_error("sparse_prove: one evaluation per matrix expected");
  size_t N = spg_sparse_num_ops(sp), M = spg_sparse_num_mem_cells(sp);
  // equalize (:1493-1507)
  std::vector<Scalar> rx_ext = rx, ry_ext = ry;
  if (rx.size() < ry.size()) rx_ext.insert(rx_ext.begin(), ry.size() - rx.size(), Scalar::zero());
  if (ry.size() < rx.size()) ry_ext.insert(ry_ext.begin(), rx.size() - ry.size(), Scalar::zero());
  if (((size_t)1 << rx_ext.size()) != M) throw std::runtime_error("sparse_prove: point does not match the matrix dimensions");
  VecPool pool;
  auto view = [&](int kind, size_t i) {
    spg_vec *v = nullptr;
    check(spg_sparse_view(sp, kind, i, &v), "spg_sparse_view");
    return pool.keep(v);
  };
  auto slice = [&](spg_vec *v, size_t off, size_t n) {
    spg_vec *o = nullptr;
    check(spg_vec_wrap(ctx, (char *)spg_vec_device_ptr(v) + off * 32, n, &o), "spg_vec_wrap");
    return pool.keep(o);
  };
  spg_vec *mem_rx = nullptr, *mem_ry = nullptr, *comb = nullptr;
  check(spg_eq_evals(ctx, to_fqv(rx_ext).data(), rx_ext.size(), &mem_rx), "spg_eq_evals");
  pool.keep(mem_rx);
  check(spg_eq_evals(ctx, to_fqv(ry_ext).data(), ry_ext.size(), &mem_ry), "spg_eq_evals");
  pool.keep(mem_ry);
  // Derefs (:34-62) and their commitment
  check(spg_sparse_deref(ctx, sp, mem_rx, mem_ry, &comb), "spg_sparse_deref");
  pool.keep(comb);
  std::vector<spg_vec *> row_ops_val, col_ops_val, val;
  for (size_t i = 0; i < batch; i++) {
    row_ops_val.push_back(slice(comb, i * N, N));
    col_ops_val.push_back(slice(comb, (batch + i) * N, N));
    val.push_back(view(SPG_SPARSE_VAL, i));
  }
  tr.lap("sparse: eq tables + derefs");
  std::vector<Compressed> comm_derefs = poly_commit_dev(ctx, gens.d_derefs, comb);
  tr.lap("sparse: derefs commitment");
  t.append_message("derefs_commitment", "begin_derefs_commitment");
  append_poly_commitment(t, "comm_poly_row_col_ops_val", comm_derefs);
  t.append_message("derefs_commitment", "end_derefs_commitment");
  std::vector<Scalar> r_mem_check = t.challenge_vector("challenge_r_hash", 2);
  spg_fq gamma = r_mem_check[0].to_fq(), tau = r_mem_check[1].to_fq();
  // Layers::new for rows and columns (:689-737)
  TreeSet trees;
  struct Side {
    spg_prodtree *init, *audit;
    std::vector<spg_prodtree *> read, write;
    Scalar e_init, e_audit;
    std::vector<Scalar> e_read, e_write;
  } side[2];
  auto tree_eval = [&](spg_prodtree *tr) {
    spg_fq o;
    check(spg_prodtree_evaluate(ctx, tr, &o), "spg_prodtree_evaluate");
    return Scalar::from_fq(o);
  };
  for (int s = 0; s < 2; s++) {
    spg_vec *mem = s == 0 ? mem_rx : mem_ry;
    spg_vec *audit_ts = view(s == 0 ? SPG_SPARSE_ROW_AUDIT_TS : SPG_SPARSE_COL_AUDIT_TS, 0);
    auto hash = [&](spg_vec *addr, spg_vec *v, spg_vec *ts, int plus_one) {
      spg_vec *o = nullptr;
      check(spg_hash_layer_fq(ctx, addr, v, ts, plus_one, &gamma, &tau, &o), "spg_hash_layer_fq");
      return o;
    };
    side[s].init = trees.build(ctx, hash(nullptr, mem, nullptr, 0));
    side[s].audit = trees.build(ctx, hash(nullptr, mem, audit_ts, 0));
    for (size_t i = 0; i < batch; i++) {
      spg_vec *addr = view(s == 0 ? SPG_SPARSE_ROW_ADDR : SPG_SPARSE_COL_ADDR, i);
      spg_vec *rts = view(s == 0 ? SPG_SPARSE_ROW_READ_TS : SPG_SPARSE_COL_READ_TS, i);
      spg_vec *ov = s == 0 ? row_ops_val[i] : col_ops_val[i];
      side[s].read.push_back(trees.build(ctx, hash(addr, ov, rts, 0)));
      side[s].write.push_back(trees.build(ctx, hash(addr, ov, rts, 1)));
    }
  }
  tr.lap("sparse: hash layers + trees");
  // PolyEvalNetworkProof::prove -> ProductLayerProof::prove (:1118-1263)
  t.append_protocol_name("Sparse polynomial evaluation proof");
  t.append_protocol_name("Sparse polynomial product layer proof");
  const char *names[2] = {"row", "col"};
  for (int s = 0; s < 2; s++) {
    Side &S = side[s];
    S.e_init = tree_eval(S.init);
    S.e_audit = tree_eval(S.audit);
    Scalar ws = Scalar::one(), rs = Scalar::one();
    for (size_t i = 0; i < batch; i++) {
      S.e_read.push_back(tree_eval(S.read[i]));
      S.e_write.push_back(tree_eval(S.write[i]));
      ws *= S.e_write[i];
      rs *= S.e_read[i];
    }
    if (!(S.e_init * ws == rs * S.e_audit)) throw std::runtime_error("sparse_prove: memory check does not balance");
    std::string n = names[s];
    t.append_scalar("claim_" + n + "_eval_init", S.e_init);
    t.append_scalars("claim_" + n + "_eval_read", S.e_read);
    t.append_scalars("claim_" + n + "_eval_write", S.e_write);
    t.append_scalar("claim_" + n + "_eval_audit", S.e_audit);
  }
  std::vector<DotpCircuit> dotp;
  std::vector<Scalar> eval_dotp_left, eval_dotp_right;
  for (size_t i = 0; i < batch; i++) {
    Scalar halves[2];
    for (int h = 0; h < 2; h++) {
      DotpCircuit d;
      auto clone = [&](spg_vec *v) {
        spg_vec *o = nullptr;
        check(spg_vec_clone(ctx, v, h * (N / 2), N / 2, &o), "spg_vec_clone");
        return pool.keep(o);
      };
      d.left = clone(row_ops_val[i]);
      d.right = clone(col_ops_val[i]);
      d.weight = clone(val[i]);
      spg_vec *tmp = nullptr;
      check(spg_vec_alloc(ctx, N / 2, &tmp), "spg_vec_alloc");
      pool.keep(tmp);
      check(spg_fq_vec_op(ctx, 0, d.left, d.right, tmp), "spg_fq_vec_op");
      spg_fq o;
      check(spg_dot(ctx, tmp, d.weight, &o), "spg_dot");
      d.eval = halves[h] = Scalar::from_fq(o);
      dotp.push_back(d);
    }
    t.append_scalar("claim_eval_dotp_left", halves[0]);
    t.append_scalar("claim_eval_dotp_right", halves[1]);
    if (!(halves[0] + halves[1] == evals[i])) throw std::runtime_error("sparse_prove: claimed evaluation is wrong");
    eval_dotp_left.push_back(halves[0]);
    eval_dotp_right.push_back(halves[1]);
  }
  std::vector<spg_prodtree *> ops_trees;
  std::vector<Scalar> ops_evals;
  for (int s = 0; s < 2; s++) {
    for (size_t i = 0; i < batch; i++) ops_trees.push_back(side[s].read[i]), ops_evals.push_back(side[s].e_read[i]);
    for (size_t i = 0; i < batch; i++) ops_trees.push_back(side[s].write[i]), ops_evals.push_back(side[s].e_write[i]);
  }
  tr.lap("sparse: tree evals + dotp");
  std::vector<Scalar> rand_ops, rand_mem;
  ProductCircuitEvalProofBatched proof_ops = pcepb_prove(ctx, ops_trees, ops_evals, dotp, t, &rand_ops);
  ProductCircuitEvalProofBatched proof_mem =
      pcepb_prove(ctx, {side[0].init, side[0].audit, side[1].init, side[1].audit},
                  {side[0].e_init, side[0].e_audit, side[1].e_init, side[1].e_audit}, {}, t, &rand_mem);
  tr.lap("sparse: product-circuit sumchecks");
  // HashLayerProof::prove (:805-918)
  t.append_protocol_name("Sparse polynomial hash layer proof");
  auto evaluate = [&](spg_vec *v, const std::vector<Scalar> &r) {
    spg_fq o;
    check(spg_dense_evaluate(ctx, v, to_fqv(r).data(), r.size(), &o), "spg_dense_evaluate");
    return Scalar::from_fq(o);
  };
  auto n_to_one = [&](std::vector<Scalar> ev, const char *label, std::vector<Scalar> *ch_out) {
    std::vector<Scalar> ch = t.challenge_vector(label, log2z(ev.size()));
    for (size_t k = ch.size(); k-- > 0;) {  // bound_poly_var_bot
      std::vector<Scalar> nx(ev.size() / 2);
      for (size_t i = 0; i < nx.size(); i++) nx[i] = ev[2 * i] + ch[k] * (ev[2 * i + 1] - ev[2 * i]);
      ev.swap(nx);
    }
    *ch_out = ch;
    return ev[0];
  };
  std::vector<Scalar> e_row_val, e_col_val;
  for (size_t i = 0; i < batch; i++) e_row_val.push_back(evaluate(row_ops_val[i], rand_ops));
  for (size_t i = 0; i < batch; i++) e_col_val.push_back(evaluate(col_ops_val[i], rand_ops));
  t.append_protocol_name("Derefs evaluation proof");
  std::vector<Scalar> evs = e_row_val;
  evs.insert(evs.end(), e_col_val.begin(), e_col_val.end());
  evs.resize(next_pow2(evs.size()), Scalar::zero());
  t.append_scalars("evals_ops_val", evs);
  std::vector<Scalar> r_joint;
  Scalar joint = n_to_one(evs, "challenge_combine_n_to_one", &r_joint);
  r_joint.insert(r_joint.end(), rand_ops.begin(), rand_ops.end());
  t.append_scalar("joint_claim_eval", joint);
  DotProductProofLog proof_derefs = polyeval_prove(ctx, comb, r_joint, joint, gens.derefs, t, tape);
  std::vector<Scalar> e_addr[2], e_rts[2];
  Scalar e_audit[2];
  for (int s = 0; s < 2; s++) {
    for (size_t i = 0; i < batch; i++) e_addr[s].push_back(evaluate(view(s == 0 ? SPG_SPARSE_ROW_ADDR : SPG_SPARSE_COL_ADDR, i), rand_ops));
    for (size_t i = 0; i < batch; i++)
      e_rts[s].push_back(evaluate(view(s == 0 ? SPG_SPARSE_ROW_READ_TS : SPG_SPARSE_COL_READ_TS, i), rand_ops));
    e_audit[s] = evaluate(view(s == 0 ? SPG_SPARSE_ROW_AUDIT_TS : SPG_SPARSE_COL_AUDIT_TS, 0), rand_mem);
  }
  std::vector<Scalar> e_val;
  for (size_t i = 0; i < batch; i++) e_val.push_back(evaluate(val[i], rand_ops));
  std::vector<Scalar> evals_ops;
  for (int s = 0; s < 2; s++) {
    evals_ops.insert(evals_ops.end(), e_addr[s].begin(), e_addr[s].end());
    evals_ops.insert(evals_ops.end(), e_rts[s].begin(), e_rts[s].end());
  }
  evals_ops.insert(evals_ops.end(), e_val.begin(), e_val.end());
  evals_ops.resize(next_pow2(evals_ops.size()), Scalar::zero());
  t.append_scalars("claim_evals_ops", evals_ops);
  std::vector<Scalar> r_joint_ops;
  Scalar joint_ops = n_to_one(evals_ops, "challenge_combine_n_to_one", &r_joint_ops);
  r_joint_ops.insert(r_joint_ops.end(), rand_ops.begin(), rand_ops.end());
  t.append_scalar("joint_claim_eval_ops", joint_ops);
  DotProductProofLog proof_ops_eval = polyeval_prove(ctx, view(SPG_SPARSE_COMB_OPS, 0), r_joint_ops, joint_ops, gens.ops, t, tape);
  std::vector<Scalar> evals_mem = {e_audit[0], e_audit[1]};
  t.append_scalars("claim_evals_mem", evals_mem);
  std::vector<Scalar> r_joint_mem;
  Scalar joint_mem = n_to_one(evals_mem, "challenge_combine_two_to_one", &r_joint_mem);
  r_joint_mem.insert(r_joint_mem.end(), rand_mem.begin(), rand_mem.end());
  t.append_scalar("joint_claim_eval_mem", joint_mem);
  DotProductProofLog proof_mem_eval = polyeval_prove(ctx, view(SPG_SPARSE_COMB_MEM, 0), r_joint_mem, joint_mem, gens.mem, t, tape);
  tr.lap("sparse: hash-layer evals + 3 openings");
  // bincode layout: SparseMatPolyEvalProof { comm_derefs, PolyEvalNetworkProof { ProductLayerProof, HashLayerProof } }
  Writer w;
  w.points(comm_derefs);
  for (int s = 0; s < 2; s++) {
    w.scalar(side[s].e_init);
    w.scalars(side[s].e_read);
    w.scalars(side[s].e_write);
    w.scalar(side[s].e_audit);
  }
  w.scalars(eval_dotp_left);
  w.scalars(eval_dotp_right);
  proof_mem.write(w);
  proof_ops.write(w);
  for (int s = 0; s < 2; s++) {
    w.scalars(e_addr[s]);
    w.scalars(e_rts[s]);
    w.scalar(e_audit[s]);
  }
  w.scalars(e_val);
  w.scalars(e_row_val);
  w.scalars(e_col_val);
  proof_ops_eval.write(w);
  proof_mem_eval.write(w);
  proof_derefs.write(w);
  return w.out;
}

}  // namespace sph
