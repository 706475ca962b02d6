//! src/ffi.rs - the `extern "C"` block and helpers a maintainer of SpekalsG3/zk-stark-tutor adds to call libzkb200.so
//! (include/zkb200.h).  INTEGRATION.md sections 2-4 explain which function bodies of the crate call which entry point.
//!
//! NOT COMPILED IN THIS REPOSITORY'S IMAGE (no rustc / cargo): every symbol bound here is exercised through the same C ABI by
//! the ctypes stub (zk_stark_tutor_b200/_lib.py, tests/test_abi.py checks that the library exports each of them) and by the C++
//! mirror of the crate's API (include/zk_impl.hpp, tests/cpp/reference_tests.cpp).
//!
//! A field element crosses the boundary as 16 little-endian bytes = Rust's in-memory `u128` on x86-64, so `&[u128]` is passed as
//! `*const c_void`.  `FieldElement<'a>` is `{ field: &Field, value: u128 }` and not `repr(C)`: vectors are packed first.
#![allow(dead_code)]
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)] pub struct ZkbCtx { _p: [u8; 0] }
#[repr(C)] pub struct ZkbTree { _p: [u8; 0] }
#[repr(C)] pub struct ZkbFriLayers { _p: [u8; 0] }
#[repr(C)] pub struct ZkbPs { _p: [u8; 0] }
#[repr(C)] pub struct ZkbNtt4 { _p: [u8; 0] }

#[repr(C)]
pub struct ZkbFriParams {
    pub offset: [u8; 16],
    pub omega: [u8; 16],
    pub domain_length: u64,
    pub expansion_factor: u64,
    pub num_colinearity_tests: u64,
}

pub type ZkbFsCallback = extern "C" fn(user: *mut c_void, round: u32, root: *const u8, want_alpha: c_int, alpha_out: *mut u8) -> c_int;

extern "C" {
    pub fn zkb_ctx_create(device: c_int, stream: *mut c_void, out: *mut *mut ZkbCtx) -> c_int;
    pub fn zkb_ctx_destroy(ctx: *mut ZkbCtx);
    pub fn zkb_last_error(ctx: *const ZkbCtx) -> *const c_char;

    pub fn zkb_ntt(ctx: *mut ZkbCtx, root: *const u8, input: *const c_void, n_in: usize, out: *mut c_void) -> c_int;
    pub fn zkb_intt(ctx: *mut ZkbCtx, root: *const u8, input: *const c_void, n_in: usize, out: *mut c_void) -> c_int;
    pub fn zkb_poly_scale(ctx: *mut ZkbCtx, factor: *const u8, coeffs: *const c_void, n: usize, out: *mut c_void) -> c_int;
    pub fn zkb_coset_lde(ctx: *mut ZkbCtx, omega: *const u8, order: u64, offset: *const u8,
                         coeffs: *const c_void, n_coeffs: usize, out: *mut c_void) -> c_int;
    pub fn zkb_poly_mul(ctx: *mut ZkbCtx, root: *const u8, root_order: u64, lhs: *const c_void, n_lhs: usize,
                        rhs: *const c_void, n_rhs: usize, out: *mut c_void, n_out: *mut usize) -> c_int;
    pub fn zkb_coset_div(ctx: *mut ZkbCtx, root: *const u8, root_order: u64, offset: *const u8, lhs: *const c_void,
                         n_lhs: usize, rhs: *const c_void, n_rhs: usize, out: *mut c_void, n_out: *mut usize) -> c_int;

    pub fn zkb_merkle_commit(ctx: *mut ZkbCtx, vals: *const c_void, n: usize, root: *mut u8) -> c_int;
    pub fn zkb_merkle_build(ctx: *mut ZkbCtx, vals: *const c_void, n: usize, tree: *mut *mut ZkbTree) -> c_int;
    pub fn zkb_merkle_open(tree: *mut ZkbTree, idx: *const u64, k: usize, paths_out: *mut u8) -> c_int;
    pub fn zkb_merkle_free(tree: *mut ZkbTree);

    pub fn zkb_fri_commit(ctx: *mut ZkbCtx, p: *const ZkbFriParams, codeword: *const c_void, n: usize,
                          fs: ZkbFsCallback, user: *mut c_void, layers: *mut *mut ZkbFriLayers) -> c_int;
    pub fn zkb_fri_layer_count(l: *const ZkbFriLayers) -> u64;
    pub fn zkb_fri_layer_len(l: *const ZkbFriLayers, round: u64) -> u64;
    pub fn zkb_fri_layer_codeword(l: *mut ZkbFriLayers, round: u64, out: *mut c_void) -> c_int;
    pub fn zkb_fri_query(l: *mut ZkbFriLayers, round: u64, idx_c: *const u64, ncc: usize,
                         leafs_out: *mut u8, paths_out: *mut u8) -> c_int;
    pub fn zkb_fri_layers_free(l: *mut ZkbFriLayers);

    // ---- the library's own proof stream (IndependentProofStream / SignatureProofStream wire format) and the calls that run
    //      WITHOUT a host hop per round: Fiat-Shamir (SHAKE256, Field::sample) is drawn on the device
    pub fn zkb_ps_create(document: *const u8, document_len: usize, is_signature: c_int, out: *mut *mut ZkbPs) -> c_int;
    pub fn zkb_ps_free(ps: *mut ZkbPs);
    pub fn zkb_ps_push_root(ps: *mut ZkbPs, root: *const u8, len: usize) -> c_int;
    pub fn zkb_ps_push_object(ps: *mut ZkbPs, code: u8, payload: *const u8, len: usize) -> c_int;
    pub fn zkb_ps_digest(ps: *const ZkbPs, out: *mut u8, cap: usize) -> usize;
    pub fn zkb_ps_fiat_shamir(ps: *const ZkbPs, num_bytes: usize, out: *mut u8) -> c_int;
    pub fn zkb_fri_commit_ps(ctx: *mut ZkbCtx, p: *const ZkbFriParams, codeword: *const c_void, n: usize, ps: *mut ZkbPs,
                             layers: *mut *mut ZkbFriLayers) -> c_int;
    pub fn zkb_lde_fri_commit_ps(ctx: *mut ZkbCtx, p: *const ZkbFriParams, coeffs: *const c_void, n_coeffs: usize, ps: *mut ZkbPs,
                                 layers: *mut *mut ZkbFriLayers) -> c_int;
    pub fn zkb_fri_prove(ctx: *mut ZkbCtx, p: *const ZkbFriParams, codeword: *const c_void, n: usize, ps: *mut ZkbPs,
                         top_indices_out: *mut u64) -> c_int;
    pub fn zkb_merkle_open_ps(tree: *mut ZkbTree, idx: *const u64, k: usize, ps: *mut ZkbPs) -> c_int;

    // ---- one box, several GPUs: independent columns, and ONE NTT over all GPUs
    pub fn zkb_lde_commit_batch(ctxs: *const *mut ZkbCtx, n_ctx: usize, p: *const ZkbFriParams, cols: *const *const c_void,
                                n_coeffs: usize, ncols: usize, roots_out: *mut u8) -> c_int;
    pub fn zkb_ntt_4step(ctxs: *const *mut ZkbCtx, world: usize, root: *const u8, inverse: c_int, x_local: *const *const c_void,
                         n_local: usize, out: *const *mut c_void) -> c_int;
    pub fn zkb_ntt4_create(ctx: *mut ZkbCtx, rank: u32, world: u32, n_local: usize, plan: *mut *mut ZkbNtt4) -> c_int;
    pub fn zkb_ntt4_free(plan: *mut ZkbNtt4);
    pub fn zkb_ntt4_connect_local(plans: *const *mut ZkbNtt4, world: usize) -> c_int;
    pub fn zkb_ntt4_export(plan: *mut ZkbNtt4, handle: *mut u8 /* 128 bytes */) -> c_int;
    pub fn zkb_ntt4_connect_ipc(plan: *mut ZkbNtt4, all_handles: *const u8 /* world x 128 bytes */) -> c_int;
    pub fn zkb_ntt4_scatter(plan: *mut ZkbNtt4, root: *const u8, inverse: c_int, x_local: *const c_void) -> c_int;
    pub fn zkb_ntt4_finish(plan: *mut ZkbNtt4, out: *mut c_void) -> c_int;
    pub fn zkb_ntt4_run(plans: *const *mut ZkbNtt4, world: usize, root: *const u8, inverse: c_int, x_local: *const *const c_void,
                        out: *const *mut c_void) -> c_int;

    // ---- Stark::prove for a batch of instances of one AIR as one call (csrc/prover.cu); `air` from zkb_air_create
    pub fn zkb_stark_prove_batch(ctx: *mut ZkbCtx, air: *mut c_void, shape: *const c_void /* zkb_stark_shape, include/zkb200.h */, batch: usize,
                                 traces: *const c_void, boundary_values: *const c_void, randomness: *const c_void,
                                 ps: *const *mut ZkbPs, proof_len_out: *mut u64) -> c_int;
    // ---- per-context knobs
    pub fn zkb_ctx_tail_threads(ctx: *mut ZkbCtx, threads: c_int) -> c_int;
    pub fn zkb_ctx_blocking_sync(ctx: *mut ZkbCtx, enable: c_int) -> c_int;
    pub fn zkb_ctx_zero_copy_inputs(ctx: *mut ZkbCtx, enable: c_int) -> c_int;
    pub fn zkb_ctx_sync(ctx: *mut ZkbCtx) -> c_int;
}

thread_local! {
    /// One context per thread (the reference is single-threaded; a context is not Sync).
    pub static CTX: *mut ZkbCtx = unsafe {
        let mut c = std::ptr::null_mut();
        let rc = zkb_ctx_create(0, std::ptr::null_mut(), &mut c);
        assert!(rc == 0, "zkb_ctx_create failed: no CUDA device (there is no CPU fallback)");
        c
    };
}

/// Non-zero status -> panic with the library's message, which repeats the reference's own
/// panic text for the same misuse (SURVEY.md 8b "Error conventions").
pub fn check(ctx: *mut ZkbCtx, rc: c_int) {
    if rc != 0 {
        let msg = unsafe { std::ffi::CStr::from_ptr(zkb_last_error(ctx)) }.to_string_lossy().into_owned();
        panic!("{}", msg);
    }
}

pub fn pack(v: &[crate::field::field_element::FieldElement]) -> Vec<u128> { v.iter().map(|e| e.value).collect() }
pub fn unpack<'a>(field: &'a crate::field::field::Field, v: Vec<u128>) -> Vec<crate::field::field_element::FieldElement<'a>> {
    v.into_iter().map(|x| crate::field::field_element::FieldElement::new(field, x)).collect()
}
