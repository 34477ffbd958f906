// C entry points of libspghost.so: the C++ host mirror of the reference's transcript-side
// prover (what stays in Rust in production), driving libspgpu.so through its C ABI.
#include <cstdlib>
#include <memory>

#include "protocol.hpp"
#include "sparse.hpp"

using namespace sph;

static thread_local std::string g_err;

extern "C" {

const char *sph_last_error(void) { return g_err.c_str(); }

void sph_free(void *p) { free(p); }

// stage timings recorded since the last sph_timings_reset, one "label\tmilliseconds\n" line each;
// returns the length needed (writes at most cap - 1 bytes + NUL)
void sph_timings_reset(void) { timing_log().clear(); }
size_t sph_timings(char *out, size_t cap) {
  std::string s;
  char buf[64];
  for (auto &kv : timing_log()) {
    snprintf(buf, sizeof buf, "\t%.6f\n", kv.second);
    s += kv.first + buf;
  }
  if (out && cap) {
    size_t n = s.size() < cap - 1 ? s.size() : cap - 1;
    memcpy(out, s.data(), n);
    out[n] = 0;
  }
  return s.size() + 1;
}

// MultiCommitGens::new(n, label).compressed(): n + 1 points (G[0..n], h), 32 bytes each
int sph_gens_derive(const char *label, size_t n, uint8_t *out) {
  try {
    MultiCommitGens g(n, label);
    std::vector<uint8_t> c = g.compressed();
    memcpy(out, c.data(), c.size());
    return 0;
  } catch (const std::exception &e) {
    g_err = e.what();
    return -1;
  }
}

// merlin check vector helper: Transcript::new(label); append_message(l, m); challenge_bytes(c, n)
int sph_transcript_kat(const char *label, const char *l, const char *m, const char *c, size_t n, uint8_t *out) {
  Transcript t(label);
  t.append_message(l, m);
  t.challenge_bytes(c, out, n);
  return 0;
}

// R1CSGens::new (src/r1csproof.rs:45-80), created once by the caller like the reference's SNARKGens; with a
// context the opening-proof generators also live on the device (fixed-base tables built on first use)
void *sph_r1cs_gens_new(spg_ctx *ctx, const char *label, size_t num_vars) {
  try {
    std::unique_ptr<R1CSGens> g(new R1CSGens(label, num_vars, ctx));
    // a generator set that is created once and reused (the reference's SNARKGens): its tables are setup too --
    // the per-window table of the openings' few-row MSMs and, for a polynomial of num_vars scalars, the
    // single-window table of its row commitments (spg_gens_prepare_rows decides whether that one pays)
    if (ctx && g->d_pc) {
      size_t ell = log2z(num_vars);
      check(spg_gens_prepare_rows(ctx, g->d_pc, (size_t)1 << (ell / 2), g->pc.n), "spg_gens_prepare_rows");
    }
    return g.release();
  } catch (const std::exception &e) {
    g_err = e.what();
    return nullptr;
  }
}

void sph_r1cs_gens_free(void *g) { delete (R1CSGens *)g; }
// the device copy of gens_pc.gens.gens_n (the generators DensePolynomial::commit uses for the witness
// sections, src/dense_mlpoly.rs:214-239); owned by the R1CSGens object
spg_gens *sph_r1cs_gens_device_pc(void *g) { return g ? ((R1CSGens *)g)->d_pc : nullptr; }

// R1CSProof::prove (src/r1csproof.rs:210-685). The caller owns the device handles.
//   gens_handle: from sph_r1cs_gens_new, or NULL to derive host-only generators for this call
//   sec_*: per witness section: number of instances, then num_proofs / num_inputs per instance (flattened)
//   out_bytes / out_len: bincode layout of the proof (malloc'ed; free with sph_free)
//   out_challenges: rp | rq_rev | rx | rw ++ ry as Montgomery scalars; out_counts[4] their lengths
int sph_r1cs_prove(spg_ctx *ctx, const char *transcript_label, const char *gens_label, const uint64_t tape_seed[4],
                   size_t num_instances, size_t max_num_proofs, const size_t *num_proofs, size_t max_num_inputs,
                   const size_t *num_inputs, size_t num_witness_secs, spg_witness *const *secs,
                   const size_t *sec_num_instances, const size_t *sec_num_proofs, const size_t *sec_num_inputs,
                   const spg_r1cs *inst, size_t inst_num_instances, size_t inst_max_num_cons, const size_t *inst_num_cons,
                   size_t gens_num_vars, const void *gens_handle, uint8_t **out_bytes, size_t *out_len, spg_fq *out_challenges,
                   size_t out_counts[4]) {
  try {
    std::vector<WitnessSec> ws;
    size_t k = 0;
    for (size_t i = 0; i < num_witness_secs; i++) {
      WitnessSec w;
      w.dev = secs[i];
      for (size_t p = 0; p < sec_num_instances[i]; p++, k++) {
        w.num_proofs.push_back(sec_num_proofs[k]);
        w.num_inputs.push_back(sec_num_inputs[k]);
      }
      ws.push_back(w);
    }
    std::unique_ptr<R1CSGens> own;
    if (!gens_handle) own.reset(new R1CSGens(gens_label, gens_num_vars));  // sized for the largest committed polynomial
    const R1CSGens &gens = gens_handle ? *(const R1CSGens *)gens_handle : *own;
    ProofTranscript t(transcript_label);
    hfq seed{{tape_seed[0], tape_seed[1], tape_seed[2], tape_seed[3]}};
    RandomTape tape("proof", Scalar(seed));
    R1CSProofOut o = r1cs_prove(ctx, num_instances, max_num_proofs, std::vector<size_t>(num_proofs, num_proofs + num_instances),
                                max_num_inputs, std::vector<size_t>(num_inputs, num_inputs + num_instances), ws, inst,
                                inst_num_instances, inst_max_num_cons,
                                std::vector<size_t>(inst_num_cons, inst_num_cons + inst_num_instances), gens, t, tape);
    *out_len = o.bytes.size();
    *out_bytes = (uint8_t *)malloc(o.bytes.size());
    memcpy(*out_bytes, o.bytes.data(), o.bytes.size());
    size_t pos = 0;
    const std::vector<Scalar> *vs[4] = {&o.rp, &o.rq_rev, &o.rx, &o.rw_ry};
    for (int i = 0; i < 4; i++) {
      out_counts[i] = vs[i]->size();
      for (auto &s : *vs[i]) out_challenges[pos++] = s.to_fq();
    }
    return 0;
  } catch (const std::exception &e) {
    g_err = e.what();
    return -1;
  }
}

// The PolyEvalProof variants other than the disjoint-rounds one R1CSProof::prove uses (src/dense_mlpoly.rs:531,
// 689, 1046), over device-resident polynomials.
//   variant 0: prove_batched_points  (polys[0]; num_points points of r_len scalars each, Zr per point)
//   variant 1: prove_batched_instances (num_polys polynomials, one point of r_len scalars and one Zr each)
//   variant 2: prove_uni_batched_instances (num_polys polynomials, the single scalar r[0], one Zr each);
//              out_extra receives C_Zr_prime
// gens: DotProductProofGens::new(gens_n, gens_label) with the bases on the device.
// out: bincode of Vec<PolyEvalProof> (variants 0, 1) or of one PolyEvalProof (variant 2); free with sph_free.
int sph_polyeval_prove(spg_ctx *ctx, int variant, const char *transcript_label, const char *gens_label,
                       const uint64_t tape_seed[4], size_t gens_n, size_t num_polys, spg_vec *const *polys, size_t num_points,
                       size_t r_len, const spg_fq *r, const spg_fq *Zr, uint8_t **out, size_t *out_len, uint8_t out_extra[32]) {
  spg_gens *dev = nullptr;
  try {
    DotProductProofGens gens;
    if (gens_n >= 256) {
      gens = DotProductProofGens::on_device(ctx, gens_n, gens_label, &dev);
    } else {
      gens = DotProductProofGens(gens_n, gens_label);
      std::vector<uint8_t> c = gens.gens_n.compressed();
      check(spg_gens_upload(ctx, c.data(), gens.gens_n.n + 1, &dev), "spg_gens_upload");
      gens.attach_device(ctx, dev);
    }
    ProofTranscript t(transcript_label);
    hfq seed{{tape_seed[0], tape_seed[1], tape_seed[2], tape_seed[3]}};
    RandomTape tape("proof", Scalar(seed));
    std::vector<std::vector<Scalar>> r_list(num_points);
    std::vector<Scalar> Zr_list;
    for (size_t i = 0; i < num_points; i++) {
      for (size_t k = 0; k < r_len; k++) r_list[i].push_back(Scalar::from_fq(r[i * r_len + k]));
      Zr_list.push_back(Scalar::from_fq(Zr[i]));
    }
    std::vector<const spg_vec *> pl(polys, polys + num_polys);
    Writer w;
    if (variant == 0) {
      auto proofs = prove_batched_points(ctx, pl.at(0), r_list, Zr_list, gens, t, tape);
      w.u64(proofs.size());
      for (auto &p : proofs) p.write(w);
    } else if (variant == 1) {
      if (num_points != num_polys) throw std::runtime_error("prove_batched_instances: one point per polynomial");
      auto proofs = prove_batched_instances(ctx, pl, r_list, Zr_list, gens, t, tape);
      w.u64(proofs.size());
      for (auto &p : proofs) p.write(w);
    } else if (variant == 2) {
      if (num_points != num_polys || r_len != 1) throw std::runtime_error("prove_uni_batched_instances: one scalar point, one Zr per polynomial");
      auto pr = prove_uni_batched_instances(ctx, pl, r_list[0][0], Zr_list, gens, t, tape);
      pr.first.write(w);
      if (out_extra) memcpy(out_extra, pr.second.b, 32);
    } else {
      throw std::runtime_error("sph_polyeval_prove: unknown variant");
    }
    spg_gens_destroy(dev);
    *out_len = w.out.size();
    *out = (uint8_t *)malloc(w.out.size());
    memcpy(*out, w.out.data(), w.out.size());
    return 0;
  } catch (const std::exception &e) {
    spg_gens_destroy(dev);
    g_err = e.what();
    return -1;
  }
}

// SparseMatPolyCommitmentGens::new (src/sparse_mlpoly.rs:289-316) with the bases and their window tables
// resident on the device; created once and reused across proofs
void *sph_sparse_gens_new(spg_ctx *ctx, const char *label, size_t num_vars_x, size_t num_vars_y, size_t max_nz,
                          size_t batch) {
  try {
    std::unique_ptr<SparseGens> g(new SparseGens(ctx, label, num_vars_x, num_vars_y, max_nz, batch));
    // tables for a commitment of 2^(nv / 2) rows over the 2^(nv - nv / 2) bases (DensePolynomial::commit)
    check(spg_gens_prepare_rows(ctx, g->d_ops, (size_t)1 << (g->nv_ops / 2), g->ops.n), "spg_gens_prepare_rows");
    check(spg_gens_prepare_rows(ctx, g->d_mem, (size_t)1 << (g->nv_mem / 2), g->mem.n), "spg_gens_prepare_rows");
    check(spg_gens_prepare_rows(ctx, g->d_derefs, (size_t)1 << (g->nv_derefs / 2), g->derefs.n), "spg_gens_prepare_rows");
    return g.release();
  } catch (const std::exception &e) {
    g_err = e.what();
    return nullptr;
  }
}
void sph_sparse_gens_free(void *g) { delete (SparseGens *)g; }

// SparseMatPolynomial::multi_commit + SparseMatPolyEvalProof::prove (src/sparse_mlpoly.rs:566-586, 1509-1564)
// for `batch` matrices given as concatenated (row, col, val) entries, nnz[i] each.
//   out_comm: bincode of SparseMatPolyCommitment; out_proof: bincode of SparseMatPolyEvalProof (malloc'ed)
//   gens_handle: from sph_sparse_gens_new (created once, like the reference's SNARKGens), or NULL to
//   derive the generators inside this call
int sph_sparse_prove(spg_ctx *ctx, const char *transcript_label, const char *gens_label, const uint64_t tape_seed[4],
                     size_t batch, size_t num_vars_x, size_t num_vars_y, const size_t *nnz, const uint32_t *rows,
                     const uint32_t *cols, const spg_fq *vals, const spg_fq *rx, const spg_fq *ry, const spg_fq *evals,
                     const void *gens_handle, uint8_t **out_comm, size_t *out_comm_len, uint8_t **out_proof,
                     size_t *out_proof_len) {
  spg_sparse *sp = nullptr;
  try {
    size_t max_nz = 0;
    for (size_t i = 0; i < batch; i++) max_nz = nnz[i] > max_nz ? nnz[i] : max_nz;
    Trace tr;
    check(spg_sparse_create(ctx, batch, num_vars_x, num_vars_y, nnz, rows, cols, vals, &sp), "spg_sparse_create");
    tr.lap("sparse: dense representation");
    std::unique_ptr<SparseGens> own;
    if (!gens_handle) own.reset(new SparseGens(ctx, gens_label, num_vars_x, num_vars_y, max_nz, batch));
    const SparseGens &gens = gens_handle ? *(const SparseGens *)gens_handle : *own;
    tr.lap("sparse: generators");
    SparseCommitment c = sparse_commit(ctx, sp, batch, gens);
    tr.lap("sparse: multi_commit");
    Writer wc;
    wc.u64(c.batch_size);
    wc.u64(c.num_ops);
    wc.u64(c.num_mem_cells);
    wc.points(c.comm_comb_ops);
    wc.points(c.comm_comb_mem);
    ProofTranscript t(transcript_label);
    hfq seed{{tape_seed[0], tape_seed[1], tape_seed[2], tape_seed[3]}};
    RandomTape tape("proof", Scalar(seed));
    std::vector<Scalar> vrx, vry, vev;
    for (size_t i = 0; i < num_vars_x; i++) vrx.push_back(Scalar::from_fq(rx[i]));
    for (size_t i = 0; i < num_vars_y; i++) vry.push_back(Scalar::from_fq(ry[i]));
    for (size_t i = 0; i < batch; i++) vev.push_back(Scalar::from_fq(evals[i]));
    std::vector<uint8_t> proof = sparse_prove(ctx, sp, batch, vrx, vry, vev, gens, t, tape);
    spg_sparse_destroy(sp);
    sp = nullptr;
    *out_comm_len = wc.out.size();
    *out_comm = (uint8_t *)malloc(wc.out.size());
    memcpy(*out_comm, wc.out.data(), wc.out.size());
    *out_proof_len = proof.size();
    *out_proof = (uint8_t *)malloc(proof.size());
    memcpy(*out_proof, proof.data(), proof.size());
    return 0;
  } catch (const std::exception &e) {
    spg_sparse_destroy(sp);
    g_err = e.what();
    return -1;
  }
}

}  // extern "C"
