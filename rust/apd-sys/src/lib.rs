//! Raw bindings of include/apd.h (ABI version 2).  One declaration per exported symbol;
//! the reference interface each one replaces is cited in the header.
//! tests/test_rust_sources.py checks this file against the header: every symbol declared,
//! and the #[repr(C)] struct layouts against `sizeof` / `offsetof` in a generated C program.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

pub const APD_ABI_VERSION: u32 = 2;
pub const APD_OK: c_int = 0;
pub const APD_ERR_INVALID: c_int = 1;
pub const APD_ERR_NO_DEVICE: c_int = 2;
pub const APD_ERR_CUDA: c_int = 3;
pub const APD_ERR_UNSUPPORTED: c_int = 4;
pub const APD_ERR_STATE: c_int = 5;
pub const APD_ERR_INTERNAL: c_int = 6;
pub const APD_MODE_STRICT: u32 = 0;
pub const APD_MODE_FAST: u32 = 1;
pub const APD_MAX_DIM: u32 = 32;
pub const APD_MAX_DEVICES: u32 = 8;
pub const APD_AE_MAX_BINS: u32 = 64;

#[repr(C)]
pub struct apd_ctx {
    _private: [u8; 0],
}

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct apd_params {
    pub warping_band_percentage: f32,
    pub insertion_penalty: f32,
    pub deletion_penalty: f32,
    pub match_penalty: f32,
    pub mode: u32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct apd_stats {
    pub n_sequences: u64,
    pub ordered_pairs: u64,
    pub units_total: u64,
    pub units_local: u64,
    pub cells_reference: u64,
    pub cells_computed: u64,
    pub kernel_launches: u32,
    pub kernel_ms: f32,
    pub scatter_ms: f32,
    pub h2d_ms: f32,
    pub d2h_ms: f32,
    pub h2d_bytes: u64,
    pub d2h_bytes: u64,
    pub sm_clock_mhz: f32,
    pub sm_count: u32,
    pub select_ms: f32,
    pub path_ms: f32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct apd_merge {
    pub merge_i: u32,
    pub merge_j: u32,
    pub into: u32,
    pub distance: f32,
    pub operation: u32,
    pub tie: u32,
}

extern "C" {
    pub fn apd_abi_version() -> u32;
    pub fn apd_create(device_id: c_int, out: *mut *mut apd_ctx) -> c_int;
    pub fn apd_create_multi(device_ids: *const c_int, n_dev: c_int, out: *mut *mut apd_ctx) -> c_int;
    pub fn apd_device_count(count: *mut c_int) -> c_int;
    pub fn apd_group_size(ctx: *mut apd_ctx, n_dev: *mut u32, peer_stores: *mut u32) -> c_int;
    pub fn apd_destroy(ctx: *mut apd_ctx);
    pub fn apd_last_error(ctx: *const apd_ctx) -> *const c_char;
    pub fn apd_set_sequences(ctx: *mut apd_ctx, frames: *const *const f32, lens: *const u32, n: u32, dim: u32) -> c_int;
    pub fn apd_set_sequences_flat(ctx: *mut apd_ctx, flat: *const f32, offsets: *const u64, lens: *const u32, n: u32,
                                  dim: u32) -> c_int;
    pub fn apd_set_sequences_encoded(ctx: *mut apd_ctx, cepstra: *const *const f32, lens: *const u32, n: u32, n_bins: u32,
                                     w_encode: *const f32, b_encode: *const f32, n_latent: u32) -> c_int;
    pub fn apd_get_sequence(ctx: *mut apd_ctx, index: u32, out: *mut f32, cap_floats: u64) -> c_int;
    pub fn apd_set_sequences_layout(ctx: *mut apd_ctx, lens: *const u32, n: u32, dim: u32) -> c_int;
    pub fn apd_arena_device(ctx: *mut apd_ctx, d_arena: *mut *mut c_void, n_floats: *mut u64) -> c_int;
    pub fn apd_arena_commit(ctx: *mut apd_ctx) -> c_int;
    pub fn apd_set_shard(ctx: *mut apd_ctx, rank: u32, world: u32) -> c_int;
    pub fn apd_align_all(ctx: *mut apd_ctx, p: *const apd_params, out_nxn: *mut f32) -> c_int;
    pub fn apd_packed_len(ctx: *mut apd_ctx, p: *const apd_params, n_floats: *mut u64) -> c_int;
    pub fn apd_align_packed(ctx: *mut apd_ctx, p: *const apd_params, d_packed: *mut f32, stream: *mut c_void) -> c_int;
    pub fn apd_scatter_packed(ctx: *mut apd_ctx, d_gathered: *const f32, world: u32, d_out_nxn: *mut f32,
                              stream: *mut c_void) -> c_int;
    pub fn apd_synchronize(ctx: *mut apd_ctx, stream: *mut c_void) -> c_int;
    pub fn apd_align_pair(ctx: *mut apd_ctx, p: *const apd_params, i: u32, j: u32, score: *mut f32, path_ij: *mut u32,
                          path_cap: u64, path_len: *mut u64) -> c_int;
    pub fn apd_align_pairs(ctx: *mut apd_ctx, p: *const apd_params, pairs_ij: *const u32, n_pairs: u64,
                           scores: *mut f32, paths_ij: *mut u32, path_cap: u64, path_lens: *mut u64) -> c_int;
    pub fn apd_align_pairs_band(ctx: *mut apd_ctx, p: *const apd_params, warping_band: u64, pairs_ij: *const u32,
                                n_pairs: u64, scores: *mut f32, paths_ij: *mut u32, path_cap: u64,
                                path_lens: *mut u64) -> c_int;
    pub fn apd_percentile_matrix(ctx: *mut apd_ctx, perc: f32, out: *mut f32) -> c_int;
    pub fn apd_percentile_device(ctx: *mut apd_ctx, d_x: *const f32, len: u64, perc: f32, stream: *mut c_void,
                                 out: *mut f32) -> c_int;
    pub fn apd_upgma(dist_nxn: *const f32, n: u32, perc: f32, threshold_in: *const f32, ops: *mut apd_merge,
                     n_ops: *mut u32, threshold_out: *mut f32, assignment_out: *mut u32) -> c_int;
    pub fn apd_save_matrix(stem: *const c_char, dist_nxn: *const f32, n: u32, params_json: *const c_char) -> c_int;
    pub fn apd_load_matrix(stem: *const c_char, out_nxn: *mut f32, cap_floats: u64, n_out: *mut u32, verify: c_int) -> c_int;
    pub fn apd_save_paths(stem: *const c_char, pairs_ij: *const u32, n_pairs: u64, scores: *const f32, paths_ij: *const u32,
                          path_cap: u64, path_lens: *const u64) -> c_int;
    pub fn apd_load_paths(stem: *const c_char, pairs_ij: *mut u32, scores: *mut f32, path_lens: *mut u64, cap_pairs: u64,
                          paths_ij: *mut u32, path_cap: u64, n_pairs: *mut u64) -> c_int;
    pub fn apd_get_stats(ctx: *mut apd_ctx, out: *mut apd_stats) -> c_int;
    pub fn apd_last_launch_plan(ctx: *mut apd_ctx) -> *const c_char;
}
