//! Raw bindings to `libspgpu.so`, the B200 prover backend for spartan-parallel's data-parallel
//! R1CS proving path. GENERATED from `include/spgpu.h` by `tools/gen_sys_crate.py` -- do not edit;
//! the header documents every function (reference file:line each one replaces).
#![allow(non_camel_case_types, non_snake_case)]
#![no_std]

use core::ffi::{c_char, c_int, c_void};

/// The reference's `Scalar`: four little-endian u64 limbs of a * 2^256 mod q, fully reduced
/// (`src/scalar/ristretto255.rs:193-199`). Layout-compatible with `Scalar(pub(crate) [u64; 4])`.
#[repr(C)]
#[derive(Clone, Copy, Debug, Default, PartialEq, Eq)]
pub struct spg_fq {
    pub l: [u64; 4],
}

#[repr(C)]
pub struct spg_ctx {
    _opaque: [u8; 0],
}
#[repr(C)]
pub struct spg_vec {
    _opaque: [u8; 0],
}
#[repr(C)]
pub struct spg_r1cs {
    _opaque: [u8; 0],
}
#[repr(C)]
pub struct spg_witness {
    _opaque: [u8; 0],
}
#[repr(C)]
pub struct spg_zmat {
    _opaque: [u8; 0],
}
#[repr(C)]
pub struct spg_sc1 {
    _opaque: [u8; 0],
}
#[repr(C)]
pub struct spg_sc2 {
    _opaque: [u8; 0],
}
#[repr(C)]
pub struct spg_cubic {
    _opaque: [u8; 0],
}
#[repr(C)]
pub struct spg_gens {
    _opaque: [u8; 0],
}
#[repr(C)]
pub struct spg_bullet {
    _opaque: [u8; 0],
}
#[repr(C)]
pub struct spg_prodtree {
    _opaque: [u8; 0],
}
#[repr(C)]
pub struct spg_sparse {
    _opaque: [u8; 0],
}

pub const SPG_OK: c_int = 0;
pub const SPG_EINVAL: c_int = -1;
pub const SPG_ECUDA: c_int = -2;
pub const SPG_ENOMEM: c_int = -3;
pub const SPG_ESTATE: c_int = -4;
pub const SPG_SPARSE_ROW_ADDR: c_int = 0;
pub const SPG_SPARSE_ROW_READ_TS: c_int = 1;
pub const SPG_SPARSE_COL_ADDR: c_int = 2;
pub const SPG_SPARSE_COL_READ_TS: c_int = 3;
pub const SPG_SPARSE_VAL: c_int = 4;
pub const SPG_SPARSE_ROW_AUDIT_TS: c_int = 5;
pub const SPG_SPARSE_COL_AUDIT_TS: c_int = 6;
pub const SPG_SPARSE_COMB_OPS: c_int = 7;
pub const SPG_SPARSE_COMB_MEM: c_int = 8;

extern "C" {
    pub fn spg_last_error() -> *const c_char;
    pub fn spg_version() -> c_int;
    pub fn spg_ctx_create(device: c_int, out: *mut *mut spg_ctx) -> c_int;
    pub fn spg_ctx_destroy(ctx: *mut spg_ctx);
    pub fn spg_ctx_sync(ctx: *mut spg_ctx) -> c_int;
    pub fn spg_ctx_launch_count(ctx: *const spg_ctx) -> u64;
    pub fn spg_ctx_profile_begin(ctx: *mut spg_ctx) -> c_int;
    pub fn spg_ctx_profile_end(ctx: *mut spg_ctx, out_json: *mut c_char, cap: usize) -> c_int;
    pub fn spg_ctx_stream(ctx: *const spg_ctx) -> *mut c_void;
    pub fn spg_host_alloc(bytes: usize, out: *mut *mut c_void) -> c_int;
    pub fn spg_host_free(p: *mut c_void);
    pub fn spg_vec_alloc(ctx: *mut spg_ctx, n: usize, out: *mut *mut spg_vec) -> c_int;
    pub fn spg_vec_zero(ctx: *mut spg_ctx, v: *mut spg_vec) -> c_int;
    pub fn spg_vec_upload(ctx: *mut spg_ctx, host: *const spg_fq, n: usize, out: *mut *mut spg_vec) -> c_int;
    pub fn spg_vec_wrap(
        ctx: *mut spg_ctx,
        device_ptr: *mut c_void,
        n: usize,
        out: *mut *mut spg_vec,
    ) -> c_int;
    pub fn spg_vec_download(
        ctx: *mut spg_ctx,
        v: *const spg_vec,
        offset: usize,
        n: usize,
        host: *mut spg_fq,
    ) -> c_int;
    pub fn spg_vec_len(v: *const spg_vec) -> usize;
    pub fn spg_vec_device_ptr(v: *const spg_vec) -> *mut c_void;
    pub fn spg_vec_free(v: *mut spg_vec);
    pub fn spg_fq_vec_op(
        ctx: *mut spg_ctx,
        op: c_int,
        a: *const spg_vec,
        b: *const spg_vec,
        out: *mut spg_vec,
    ) -> c_int;
    pub fn spg_fq_from_u512(
        ctx: *mut spg_ctx,
        host_wide: *const u64,
        n: usize,
        out: *mut *mut spg_vec,
    ) -> c_int;
    pub fn spg_fq_host_sum(in_: *const spg_fq, count: usize, width: usize, out: *mut spg_fq) -> c_int;
    pub fn spg_fq_host_mul(a: *const spg_fq, b: *const spg_fq, out: *mut spg_fq) -> c_int;
    pub fn spg_fq_host_eq_weight(tau: *const spg_fq, nbits: usize, index: u64, out: *mut spg_fq) -> c_int;
    pub fn spg_eq_evals(ctx: *mut spg_ctx, r: *const spg_fq, ell: usize, out: *mut *mut spg_vec) -> c_int;
    pub fn spg_dense_bound_top(ctx: *mut spg_ctx, v: *mut spg_vec, r: *const spg_fq) -> c_int;
    pub fn spg_dense_bound_bot(ctx: *mut spg_ctx, v: *mut spg_vec, r: *const spg_fq) -> c_int;
    pub fn spg_dense_evaluate(
        ctx: *mut spg_ctx,
        v: *const spg_vec,
        r: *const spg_fq,
        ell: usize,
        out: *mut spg_fq,
    ) -> c_int;
    pub fn spg_dense_bound_L(
        ctx: *mut spg_ctx,
        v: *const spg_vec,
        L: *const spg_fq,
        L_size: usize,
        out: *mut *mut spg_vec,
    ) -> c_int;
    pub fn spg_dot(ctx: *mut spg_ctx, a: *const spg_vec, b: *const spg_vec, out: *mut spg_fq) -> c_int;
    pub fn spg_r1cs_create(
        ctx: *mut spg_ctx,
        num_instances: usize,
        max_num_cons: usize,
        num_cons: *const usize,
        num_vars: usize,
        nnz: *const usize,
        rows: *const u32,
        cols: *const u32,
        vals: *const spg_fq,
        out: *mut *mut spg_r1cs,
    ) -> c_int;
    pub fn spg_r1cs_destroy(inst: *mut spg_r1cs);
    pub fn spg_r1cs_multi_evaluate(
        ctx: *mut spg_ctx,
        inst: *const spg_r1cs,
        rx: *const spg_fq,
        nrx: usize,
        ry: *const spg_fq,
        nry: usize,
        out: *mut spg_fq,
    ) -> c_int;
    pub fn spg_witness_upload(
        ctx: *mut spg_ctx,
        num_instances: usize,
        num_proofs: *const usize,
        num_inputs: *const usize,
        host_w_mat: *const spg_fq,
        out: *mut *mut spg_witness,
    ) -> c_int;
    pub fn spg_witness_upload_async(
        ctx: *mut spg_ctx,
        num_instances: usize,
        num_proofs: *const usize,
        num_inputs: *const usize,
        host_w_mat: *const spg_fq,
        out: *mut *mut spg_witness,
    ) -> c_int;
    pub fn spg_witness_destroy(w: *mut spg_witness);
    pub fn spg_witness_poly(w: *mut spg_witness, p: usize, out: *mut *mut spg_vec) -> c_int;
    pub fn spg_zmat_build(
        ctx: *mut spg_ctx,
        num_instances: usize,
        num_proofs: *const usize,
        num_inputs: *const usize,
        num_witness_secs: usize,
        witness_secs: *const *mut spg_witness,
        out: *mut *mut spg_zmat,
    ) -> c_int;
    pub fn spg_zmat_destroy(z: *mut spg_zmat);
    pub fn spg_sc1_create(
        ctx: *mut spg_ctx,
        inst: *const spg_r1cs,
        z: *const spg_zmat,
        num_instances: usize,
        num_proofs: *const usize,
        max_num_proofs: usize,
        num_cons: *const usize,
        max_num_cons: usize,
        max_num_inputs: usize,
        tau_p: *const spg_fq,
        tau_q: *const spg_fq,
        tau_x: *const spg_fq,
        out: *mut *mut spg_sc1,
    ) -> c_int;
    pub fn spg_sc1_create_from_tables(
        ctx: *mut spg_ctx,
        num_instances: usize,
        num_proofs: *const usize,
        max_num_proofs: usize,
        num_cons: *const usize,
        max_num_cons: usize,
        Az: *const spg_fq,
        Bz: *const spg_fq,
        Cz: *const spg_fq,
        tau_p: *const spg_fq,
        tau_q: *const spg_fq,
        tau_x: *const spg_fq,
        out: *mut *mut spg_sc1,
    ) -> c_int;
    pub fn spg_sc1_set_scale(s: *mut spg_sc1, c: *const spg_fq) -> c_int;
    pub fn spg_sc1_set_claim(s: *mut spg_sc1, claim: *const spg_fq) -> c_int;
    pub fn spg_sc1_set_claim_checked(s: *mut spg_sc1, claim: *const spg_fq) -> c_int;
    pub fn spg_sc1_set_satisfied(s: *mut spg_sc1) -> c_int;
    pub fn spg_sc1_num_rounds(s: *const spg_sc1) -> usize;
    pub fn spg_sc1_round_eval(s: *mut spg_sc1, e: *mut spg_fq) -> c_int;
    pub fn spg_sc1_round_bind(s: *mut spg_sc1, r: *const spg_fq) -> c_int;
    pub fn spg_sc1_run_rounds(
        s: *mut spg_sc1,
        num_rounds: usize,
        challenges: *const spg_fq,
        evals_out: *mut spg_fq,
    ) -> c_int;
    pub fn spg_sc1_run_rounds_sharded(
        s: *mut spg_sc1,
        num_rounds: usize,
        challenges: *const spg_fq,
        evals_out: *mut spg_fq,
        mailbox: *mut c_void,
        slot_stride: usize,
        rank: c_int,
        world: c_int,
        calls: *mut u64,
    ) -> c_int;
    pub fn spg_mailbox_all_gather(
        mailbox: *mut c_void,
        slot_stride: usize,
        rank: c_int,
        world: c_int,
        calls: *mut u64,
        data: *const c_void,
        nbytes: usize,
        out: *mut c_void,
    ) -> c_int;
    pub fn spg_mailbox_poison(mailbox: *mut c_void, slot_stride: usize, rank: c_int, world: c_int);
    pub fn spg_sc1_set_row_weights(s: *mut spg_sc1, weights: *const spg_fq, n_rows: usize) -> c_int;
    pub fn spg_sc1_host_tail_eval(
        state: *const spg_fq,
        G: usize,
        len: usize,
        scale: *const spg_fq,
        e: *mut spg_fq,
    ) -> c_int;
    pub fn spg_sc1_host_tail_bind(state: *mut spg_fq, G: usize, len: usize, r: *const spg_fq) -> c_int;
    pub fn spg_sc1_final(s: *mut spg_sc1, claims: *mut spg_fq) -> c_int;
    pub fn spg_sc1_debug_tables(
        s: *mut spg_sc1,
        Az: *mut spg_fq,
        Bz: *mut spg_fq,
        Cz: *mut spg_fq,
        cap: usize,
        n: *mut usize,
    ) -> c_int;
    pub fn spg_sc1_destroy(s: *mut spg_sc1);
    pub fn spg_sc2_create(
        ctx: *mut spg_ctx,
        inst: *const spg_r1cs,
        z: *const spg_zmat,
        num_instances: usize,
        num_proofs: *const usize,
        max_num_proofs: usize,
        num_inputs: *const usize,
        max_num_inputs: usize,
        num_witness_secs: usize,
        rx: *const spg_fq,
        rq_rev: *const spg_fq,
        rp: *const spg_fq,
        r_A: *const spg_fq,
        r_B: *const spg_fq,
        r_C: *const spg_fq,
        out: *mut *mut spg_sc2,
    ) -> c_int;
    pub fn spg_sc2_create_from_zrq(
        ctx: *mut spg_ctx,
        inst: *const spg_r1cs,
        zrq: *const spg_vec,
        num_instances: usize,
        num_inputs: *const usize,
        max_num_inputs: usize,
        num_witness_secs: usize,
        rx: *const spg_fq,
        rp: *const spg_fq,
        r_A: *const spg_fq,
        r_B: *const spg_fq,
        r_C: *const spg_fq,
        out: *mut *mut spg_sc2,
    ) -> c_int;
    pub fn spg_sc2_create_slice(
        ctx: *mut spg_ctx,
        inst: *const spg_r1cs,
        zrq: *const spg_vec,
        max_num_inputs: usize,
        num_witness_secs: usize,
        flat_off: usize,
        flat_len: usize,
        rx: *const spg_fq,
        r_A: *const spg_fq,
        r_B: *const spg_fq,
        r_C: *const spg_fq,
        out: *mut *mut spg_sc2,
    ) -> c_int;
    pub fn spg_sc2_run_rounds_sharded(
        s: *mut spg_sc2,
        num_rounds: usize,
        challenges: *const spg_fq,
        evals_out: *mut spg_fq,
        mailbox: *mut c_void,
        slot_stride: usize,
        rank: c_int,
        world: c_int,
        calls: *mut u64,
    ) -> c_int;
    pub fn spg_sc2_host_tail_eval(
        state: *const spg_fq,
        G: usize,
        len: usize,
        mode: c_int,
        scale: *const spg_fq,
        e: *mut spg_fq,
    ) -> c_int;
    pub fn spg_sc2_host_tail_bind(
        state: *mut spg_fq,
        G: usize,
        len: usize,
        mode: c_int,
        r: *const spg_fq,
    ) -> c_int;
    pub fn spg_zmat_bind_rq(
        ctx: *mut spg_ctx,
        z: *const spg_zmat,
        rq_rev: *const spg_fq,
        nq: usize,
        scale: *const spg_fq,
        out: *mut spg_vec,
    ) -> c_int;
    pub fn spg_zmat_bind_weights(
        ctx: *mut spg_ctx,
        z: *const spg_zmat,
        weights: *const spg_fq,
        n_weights: usize,
        out_off: *const usize,
        out: *mut spg_vec,
    ) -> c_int;
    pub fn spg_sc2_num_rounds(s: *const spg_sc2) -> usize;
    pub fn spg_sc2_round_eval(s: *mut spg_sc2, e: *mut spg_fq) -> c_int;
    pub fn spg_sc2_round_bind(s: *mut spg_sc2, r: *const spg_fq) -> c_int;
    pub fn spg_sc2_run_rounds(
        s: *mut spg_sc2,
        num_rounds: usize,
        challenges: *const spg_fq,
        evals_out: *mut spg_fq,
    ) -> c_int;
    pub fn spg_sc2_final(s: *mut spg_sc2, claims: *mut spg_fq) -> c_int;
    pub fn spg_sc2_destroy(s: *mut spg_sc2);
    pub fn spg_prodtree_build(
        ctx: *mut spg_ctx,
        leaves: *const spg_vec,
        out: *mut *mut spg_prodtree,
    ) -> c_int;
    pub fn spg_prodtree_num_layers(t: *const spg_prodtree) -> usize;
    pub fn spg_prodtree_layer(
        t: *mut spg_prodtree,
        layer: usize,
        left: *mut *mut spg_vec,
        right: *mut *mut spg_vec,
    ) -> c_int;
    pub fn spg_prodtree_evaluate(ctx: *mut spg_ctx, t: *mut spg_prodtree, out: *mut spg_fq) -> c_int;
    pub fn spg_prodtree_destroy(t: *mut spg_prodtree);
    pub fn spg_cubic_create(
        ctx: *mut spg_ctx,
        npar: usize,
        A_par: *const *mut spg_vec,
        B_par: *const *mut spg_vec,
        C_par: *mut spg_vec,
        nseq: usize,
        A_seq: *const *mut spg_vec,
        B_seq: *const *mut spg_vec,
        C_seq: *const *mut spg_vec,
        coeffs: *const spg_fq,
        out: *mut *mut spg_cubic,
    ) -> c_int;
    pub fn spg_cubic_round_eval(s: *mut spg_cubic, e: *mut spg_fq) -> c_int;
    pub fn spg_cubic_round_bind(s: *mut spg_cubic, r: *const spg_fq) -> c_int;
    pub fn spg_cubic_final(s: *mut spg_cubic, claims: *mut spg_fq) -> c_int;
    pub fn spg_cubic_destroy(s: *mut spg_cubic);
    pub fn spg_hash_layer(
        ctx: *mut spg_ctx,
        addr: *const u64,
        val: *const spg_vec,
        ts: *const u64,
        n: usize,
        gamma: *const spg_fq,
        tau: *const spg_fq,
        ts_plus_one: c_int,
        out: *mut *mut spg_vec,
    ) -> c_int;
    pub fn spg_deref(
        ctx: *mut spg_ctx,
        addr: *const u64,
        n: usize,
        mem: *const spg_vec,
        out: *mut *mut spg_vec,
    ) -> c_int;
    pub fn spg_hash_layer_fq(
        ctx: *mut spg_ctx,
        addr: *const spg_vec,
        val: *const spg_vec,
        ts: *const spg_vec,
        ts_plus_one: c_int,
        gamma: *const spg_fq,
        tau: *const spg_fq,
        out: *mut *mut spg_vec,
    ) -> c_int;
    pub fn spg_sparse_create(
        ctx: *mut spg_ctx,
        batch: usize,
        num_vars_x: usize,
        num_vars_y: usize,
        nnz: *const usize,
        rows: *const u32,
        cols: *const u32,
        vals: *const spg_fq,
        out: *mut *mut spg_sparse,
    ) -> c_int;
    pub fn spg_sparse_destroy(s: *mut spg_sparse);
    pub fn spg_sparse_num_ops(s: *const spg_sparse) -> usize;
    pub fn spg_sparse_num_mem_cells(s: *const spg_sparse) -> usize;
    pub fn spg_sparse_view(s: *mut spg_sparse, kind: c_int, i: usize, out: *mut *mut spg_vec) -> c_int;
    pub fn spg_sparse_deref(
        ctx: *mut spg_ctx,
        s: *const spg_sparse,
        mem_rx: *const spg_vec,
        mem_ry: *const spg_vec,
        out: *mut *mut spg_vec,
    ) -> c_int;
    pub fn spg_vec_clone(
        ctx: *mut spg_ctx,
        v: *const spg_vec,
        offset: usize,
        n: usize,
        out: *mut *mut spg_vec,
    ) -> c_int;
    pub fn spg_perm_scan(
        ctx: *mut spg_ctx,
        n: usize,
        seg_len: *const usize,
        n_seg: usize,
        v: *const spg_vec,
        v_off: usize,
        v_stride: usize,
        x: *const spg_vec,
        x_off: usize,
        x_stride: usize,
        D: *mut spg_vec,
        D_off: usize,
        D_stride: usize,
        pi: *mut spg_vec,
        pi_off: usize,
        pi_stride: usize,
    ) -> c_int;
    pub fn spg_peer_alloc(ctx: *mut spg_ctx, n: usize, out: *mut *mut spg_vec, handle: *mut u8) -> c_int;
    pub fn spg_peer_free(v: *mut spg_vec) -> c_int;
    pub fn spg_peer_open(ctx: *mut spg_ctx, handle: *const u8, ptr: *mut *mut c_void) -> c_int;
    pub fn spg_peer_close(ptr: *mut c_void) -> c_int;
    pub fn spg_peer_sum(
        ctx: *mut spg_ctx,
        peer_ptrs: *const *mut c_void,
        world: c_int,
        rank: c_int,
        n: usize,
    ) -> c_int;
    pub fn spg_peer_reduce_scatter(
        ctx: *mut spg_ctx,
        peer_ptrs: *const *mut c_void,
        world: c_int,
        rank: c_int,
        n: usize,
    ) -> c_int;
    pub fn spg_zmat_bind_rq_sharded(
        ctx: *mut spg_ctx,
        z: *const spg_zmat,
        rq_rev: *const spg_fq,
        nq_local: usize,
        nq_total: usize,
        peer_ptrs: *const *mut c_void,
        world: c_int,
        rank: c_int,
        n: usize,
        scatter_only: c_int,
        mailbox: *mut c_void,
        slot_stride: usize,
        calls: *mut u64,
    ) -> c_int;
    pub fn spg_wit_perm_w0(
        ctx: *mut spg_ctx,
        tau: *const spg_fq,
        r: *const spg_fq,
        used: usize,
        total: usize,
        out: *mut *mut spg_vec,
    ) -> c_int;
    pub fn spg_wit_block(
        ctx: *mut spg_ctx,
        exec_mode: c_int,
        vars: *const spg_vec,
        rows: usize,
        vars_width: usize,
        perm_w0: *const spg_vec,
        tau: *const spg_fq,
        r: *const spg_fq,
        num_inputs_unpadded: usize,
        io_width: usize,
        phy_ops: usize,
        vir_ops: usize,
        w2_width: usize,
        seg_len: *const usize,
        n_seg: usize,
        w2_out: *mut *mut spg_vec,
        w3_out: *mut *mut spg_vec,
    ) -> c_int;
    pub fn spg_wit_mem(
        ctx: *mut spg_ctx,
        mems: *const spg_vec,
        rows: usize,
        in_width: usize,
        tau: *const spg_fq,
        r: *const spg_fq,
        mem_width: usize,
        w2_out: *mut *mut spg_vec,
        w3_out: *mut *mut spg_vec,
    ) -> c_int;
    pub fn spg_wit_shift(
        ctx: *mut spg_ctx,
        w3: *const spg_vec,
        rows: usize,
        width: usize,
        seg_len: *const usize,
        n_seg: usize,
        out: *mut *mut spg_vec,
    ) -> c_int;
    pub fn spg_gens_upload(
        ctx: *mut spg_ctx,
        compressed: *const u8,
        n_plus_1: usize,
        out: *mut *mut spg_gens,
    ) -> c_int;
    pub fn spg_gens_from_uniform(
        ctx: *mut spg_ctx,
        uniform: *const u8,
        n_plus_1: usize,
        out: *mut *mut spg_gens,
    ) -> c_int;
    pub fn spg_gens_destroy(g: *mut spg_gens);
    pub fn spg_poly_commit(
        ctx: *mut spg_ctx,
        gens: *const spg_gens,
        poly: *const spg_vec,
        L_size: usize,
        out_compressed: *mut u8,
    ) -> c_int;
    pub fn spg_poly_commit_rows(
        ctx: *mut spg_ctx,
        gens: *const spg_gens,
        poly: *const spg_vec,
        L_size: usize,
        row0: usize,
        nrows: usize,
        out_compressed: *mut u8,
    ) -> c_int;
    pub fn spg_gens_prepare(ctx: *mut spg_ctx, gens: *mut spg_gens, R: usize) -> c_int;
    pub fn spg_gens_info(gens: *const spg_gens, out: *mut usize) -> c_int;
    pub fn spg_gens_prepare_rows(ctx: *mut spg_ctx, gens: *mut spg_gens, L: usize, R: usize) -> c_int;
    pub fn spg_gens_info_rows(gens: *const spg_gens, out: *mut usize) -> c_int;
    pub fn spg_debug_fe8_selftest(ctx: *mut spg_ctx, n: usize, seed: u64, out_bad: *mut u32) -> c_int;
    pub fn spg_debug_fq_wide_selftest(ctx: *mut spg_ctx, n: usize, seed: u64, out_bad: *mut u32) -> c_int;
    pub fn spg_commit_batch(
        ctx: *mut spg_ctx,
        gens: *const spg_gens,
        scalars: *const spg_fq,
        len: usize,
        blinds: *const spg_fq,
        count: usize,
        out_compressed: *mut u8,
    ) -> c_int;
    pub fn spg_bullet_create(
        ctx: *mut spg_ctx,
        gens: *const spg_gens,
        n: usize,
        out: *mut *mut spg_bullet,
    ) -> c_int;
    pub fn spg_bullet_lr(
        b: *mut spg_bullet,
        nk: usize,
        a: *const spg_fq,
        blinds: *const spg_fq,
        out_LR: *mut u8,
    ) -> c_int;
    pub fn spg_bullet_fold(b: *mut spg_bullet, nk: usize, u: *const spg_fq, u_inv: *const spg_fq) -> c_int;
    pub fn spg_bullet_final(b: *mut spg_bullet, out_G: *mut u8) -> c_int;
    pub fn spg_bullet_set_ab(b: *mut spg_bullet, a: *const spg_fq, bvec: *const spg_fq) -> c_int;
    pub fn spg_bullet_lr_resident(
        b: *mut spg_bullet,
        nk: usize,
        blinds: *const spg_fq,
        ext: c_int,
        out_LR: *mut u8,
        out_c: *mut spg_fq,
    ) -> c_int;
    pub fn spg_bullet_final_ab(b: *mut spg_bullet, out_G: *mut u8, out_ab: *mut spg_fq) -> c_int;
    pub fn spg_bullet_destroy(b: *mut spg_bullet);
}
