//! Raw `extern "C"` bindings of `include/fsgpu.h` (ABI version 4) — the B200 semantic-tier hot
//! path behind frankensearch's own seams.  Each item names the reference interface it replaces;
//! see INTEGRATION.md for the safe wrapper (`GpuVectorIndex`) and the error mapping.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)] pub struct fsgpu_index { _private: [u8; 0] }
#[repr(C)] pub struct fsgpu_potion { _private: [u8; 0] }
#[repr(C)] pub struct fsgpu_minilm { _private: [u8; 0] }
#[repr(C)] pub struct fsgpu_sharded { _private: [u8; 0] }

pub const FSGPU_ABI_VERSION: c_int = 4;
/// `fsgpu_index_options.flags`: this host replays the `<index>.wal` sidecar itself (VectorIndex::open,
/// lib.rs:1833-1878) and hands the rows to `fsgpu_index_set_wal`.
pub const FSGPU_OPEN_HOST_REPLAYS_WAL: i32 = 1;

pub const FSGPU_OK: c_int = 0;
pub const FSGPU_ERR_DIMENSION_MISMATCH: c_int = 1; // SearchError::DimensionMismatch
pub const FSGPU_ERR_INVALID_CONFIG: c_int = 2;     // SearchError::InvalidConfig
pub const FSGPU_ERR_INDEX_CORRUPTED: c_int = 3;    // SearchError::IndexCorrupted
pub const FSGPU_ERR_EMBEDDING_FAILED: c_int = 4;   // SearchError::EmbeddingFailed
pub const FSGPU_ERR_CANCELLED: c_int = 5;          // SearchError::Cancelled
pub const FSGPU_ERR_SUBSYSTEM: c_int = 6;          // SearchError::SubsystemError { subsystem: "gpu" }
pub const FSGPU_ERR_IO: c_int = 7;                 // SearchError::Io

/// `VectorHit` without the doc id (crates/frankensearch-core/src/types.rs:88-95).
#[repr(C)] #[derive(Clone, Copy, Debug, Default)]
pub struct fsgpu_hit { pub row: u32, pub score: f32 }

#[repr(C)] #[derive(Clone, Copy, Debug)]
pub struct fsgpu_index_options {
    pub device: i32,
    pub reduce_order: i32,   // lane order of wide::f32x8::reduce_add (simd.rs:439); 0 = halves pairwise
    pub tail_fma: i32,       // 1: bytes-kernel tail (simd.rs:440-444), 0: slice-kernel tail (simd.rs:298-300)
    pub slab_is_device: i32,
    pub row_base: u64,       // global row of local row 0 (row-sharded corpora)
    pub int8_codes: i32,     // 1 (default): also keep int8 codes for the int8 forms of the scan (same results)
    pub flags: i32,          // FSGPU_OPEN_* bits
}

/// `RrfConfig` (crates/frankensearch-fusion/src/rrf.rs:25-48).
#[repr(C)] #[derive(Clone, Copy, Debug)]
pub struct fsgpu_rrf_config { pub k: f64, pub lexical_weight: f64, pub semantic_weight: f64, pub tiebreak: i32, pub reserved: i32 }

/// `FusedHit` (crates/frankensearch-core/src/types.rs:3892-3925).
#[repr(C)] #[derive(Clone, Copy, Debug)]
pub struct fsgpu_fused_hit {
    pub rrf_score: f64, pub semantic_rank: i32, pub lexical_rank: i32, pub semantic_row: u32,
    pub semantic_score: f32, pub lexical_score: f32, pub in_both_sources: u32,
}

#[repr(C)] #[derive(Clone, Copy, Debug)]
pub struct fsgpu_minilm_layer_weights {
    pub qkv_w: *const f32, pub qkv_b: *const f32, pub attn_out_w: *const f32, pub attn_out_b: *const f32,
    pub attn_ln_g: *const f32, pub attn_ln_b: *const f32, pub ffn_in_w: *const f32, pub ffn_in_b: *const f32,
    pub ffn_out_w: *const f32, pub ffn_out_b: *const f32, pub ffn_ln_g: *const f32, pub ffn_ln_b: *const f32,
}

#[repr(C)] #[derive(Clone, Copy, Debug)]
pub struct fsgpu_minilm_weights {
    pub vocab_size: u32, pub max_positions: u32, pub n_layers: u32, pub hidden: u32, pub heads: u32,
    pub intermediate: u32, pub ln_eps: f32, pub reserved: u32,
    pub word_emb: *const f32, pub pos_emb: *const f32, pub type_emb: *const f32,
    pub emb_ln_g: *const f32, pub emb_ln_b: *const f32, pub layers: *const fsgpu_minilm_layer_weights,
}

extern "C" {
    pub fn fsgpu_abi_version() -> c_int;
    pub fn fsgpu_last_error() -> *const c_char;
    pub fn fsgpu_device_count(out_count: *mut c_int) -> c_int;
    pub fn fsgpu_index_options_default(opts: *mut fsgpu_index_options);

    // InMemoryVectorIndex::from_vectors (in_memory.rs:1667) / VectorIndex::open (lib.rs:819)
    pub fn fsgpu_index_create_f16(slab: *const u16, n_rows: u64, dim: u32, tombstones: *const u8,
                                  opts: *const fsgpu_index_options, out: *mut *mut fsgpu_index) -> c_int;
    pub fn fsgpu_index_create_f32(rows: *const f32, n_rows: u64, dim: u32, tombstones: *const u8,
                                  opts: *const fsgpu_index_options, out: *mut *mut fsgpu_index) -> c_int;
    pub fn fsgpu_index_open_fsvi(path: *const c_char, row_start: u64, n_rows_or_0: u64,
                                 opts: *const fsgpu_index_options, out: *mut *mut fsgpu_index) -> c_int;
    pub fn fsgpu_index_destroy(index: *mut fsgpu_index);
    pub fn fsgpu_index_rows(index: *const fsgpu_index) -> u64;
    pub fn fsgpu_index_dim(index: *const fsgpu_index) -> u32;
    pub fn fsgpu_index_set_tombstones(index: *mut fsgpu_index, bitmap_or_null: *const u8) -> c_int;
    pub fn fsgpu_index_int8_ready(index: *const fsgpu_index) -> c_int;
    pub fn fsgpu_index_read_codes_i8(index: *const fsgpu_index, row_start: u64, n: u64, out_codes: *mut i8,
                                     out_scale: *mut f32) -> c_int;
    pub fn fsgpu_index_read_tombstones(index: *const fsgpu_index, out_bitmap: *mut u8) -> c_int;
    pub fn fsgpu_index_set_doc_hashes(index: *mut fsgpu_index, hashes_or_null: *const u64) -> c_int;
    pub fn fsgpu_search_top_k_hashes(index: *const fsgpu_index, queries: *const f32, batch: u32, k: u32, dim: u32,
                                     allowed_sorted: *const u64, n_allowed: u32, wal_allow_bitmap: *const u8,
                                     out: *mut fsgpu_hit, out_counts: *mut u32, out_used_gather: *mut c_int) -> c_int;
    pub fn fsgpu_index_set_wal(index: *mut fsgpu_index, embeddings: *const f32, n_wal: u32, virtual_base: u64) -> c_int;
    pub fn fsgpu_index_wal_rows(index: *const fsgpu_index) -> u32;
    pub fn fsgpu_index_doc_id(index: *const fsgpu_index, global_row: u64, out_ptr: *mut *const u8,
                              out_len: *mut u32) -> c_int;

    // VectorIndex::search_top_k (search.rs:192-206) / InMemoryVectorIndex::search_top_k (in_memory.rs:2555)
    pub fn fsgpu_search_top_k(index: *const fsgpu_index, queries: *const f32, batch: u32, k: u32, dim: u32,
                              out: *mut fsgpu_hit, out_counts: *mut u32) -> c_int;
    pub fn fsgpu_search_top_k_filtered(index: *const fsgpu_index, queries: *const f32, batch: u32, k: u32,
                                       dim: u32, allow_bitmap: *const u8, out: *mut fsgpu_hit,
                                       out_counts: *mut u32) -> c_int;
    pub fn fsgpu_search_top_k_device(index: *const fsgpu_index, d_queries: *const f32, batch: u32, k: u32,
                                     d_out_keys: *mut u64, d_out_hits: *mut fsgpu_hit,
                                     d_out_counts: *mut u32, stream: *mut c_void) -> c_int;
    // merge_partial_heaps (search.rs:1704-1720) across shards
    pub fn fsgpu_merge_top_k_hits_device(device: c_int, d_keys: *const u64, d_hits: *const fsgpu_hit, batch: u32,
                                         n_lists: u32, k_in: u32, list_stride: u64, query_stride: u64,
                                         k_out: u32, d_out_keys: *mut u64, d_out_hits: *mut fsgpu_hit,
                                         d_out_counts: *mut u32, stream: *mut c_void) -> c_int;
    // TwoTierIndex::quality_scores_for_hits (two_tier.rs:1566-1631)
    pub fn fsgpu_scores_for_rows(index: *const fsgpu_index, query: *const f32, dim: u32, rows: *const u32,
                                 n: u32, out_scores: *mut f32, out_present: *mut u8) -> c_int;
    // rrf_fuse (rrf.rs:282) / blend_two_tier (blend.rs:107)
    pub fn fsgpu_rrf_fuse(device: c_int, config: *const fsgpu_rrf_config, batch: u32,
                          lex_ids: *const u64, lex_scores: *const f32, lex_tie: *const u32,
                          lex_counts: *const u32, n_lex_max: u32,
                          sem_rows: *const u32, sem_scores: *const f32, sem_tie: *const u32,
                          sem_counts: *const u32, n_sem_max: u32, limit: u32, offset: u32,
                          out: *mut fsgpu_fused_hit, out_counts: *mut u32) -> c_int;
    pub fn fsgpu_blend_two_tier(device: c_int, blend_factor: f32, fast_rows: *const u32,
                                fast_scores: *const f32, fast_tie: *const u32, n_fast: u32,
                                quality_rows: *const u32, quality_scores: *const f32,
                                quality_present: *const u8, quality_tie: *const u32, n_quality: u32,
                                out: *mut fsgpu_hit, out_count: *mut u32) -> c_int;
    // Model2VecEmbedder (model2vec_embedder.rs:67) / FastEmbedEmbedder (fastembed_embedder.rs:169)
    pub fn fsgpu_potion_create(table: *const f32, vocab: u64, dim: u32, device: c_int,
                               out: *mut *mut fsgpu_potion) -> c_int;
    pub fn fsgpu_potion_destroy(enc: *mut fsgpu_potion);
    pub fn fsgpu_potion_embed(enc: *const fsgpu_potion, ids: *const u32, offsets: *const u64, batch: u32,
                              out: *mut f32) -> c_int;
    pub fn fsgpu_minilm_create(weights: *const fsgpu_minilm_weights, device: c_int,
                               out: *mut *mut fsgpu_minilm) -> c_int;
    pub fn fsgpu_minilm_destroy(enc: *mut fsgpu_minilm);
    pub fn fsgpu_minilm_embed(enc: *const fsgpu_minilm, ids: *const i32, lens: *const i32, batch: u32,
                              max_len: u32, out: *mut f32) -> c_int;
    // weights straight from model.safetensors (model_manifest.rs:343-349)
    pub fn fsgpu_minilm_load(safetensors_path: *const c_char, device: c_int, out: *mut *mut fsgpu_minilm) -> c_int;

    // VectorIndex::zero_signal_state (lib.rs:2441-2459): [records, live, tombstoned, wal, usable]
    pub fn fsgpu_index_zero_signal_state(index: *const fsgpu_index, out_state: *mut u64) -> c_int;
    // status words of the last (stream-asynchronous) search on an index
    pub fn fsgpu_index_last_status(index: *const fsgpu_index, out_flags: *mut u32) -> c_int;

    // the two-tier pipeline on the device (sync_searcher.rs:616-1009): re-score a shard's own candidates,
    // pick the quality score that travelled through the cross-shard merge, blend, fuse
    pub fn fsgpu_scores_for_hits_device(index: *const fsgpu_index, d_queries: *const f32, batch: u32,
                                        d_hits: *const fsgpu_hit, n_per_query: u32, d_out_scores: *mut f32,
                                        d_out_present: *mut u8, stream: *mut c_void) -> c_int;
    pub fn fsgpu_merge_payload_device(device: c_int, d_keys: *const u64, d_payload: *const f32, batch: u32,
                                      n_lists: u32, k_in: u32, list_stride: u64, query_stride: u64,
                                      payload_list_stride: u64, payload_query_stride: u64,
                                      d_merged_keys: *const u64, k_out: u32, d_out_payload: *mut f32,
                                      d_out_present: *mut u8, stream: *mut c_void) -> c_int;
    pub fn fsgpu_blend_two_tier_device(device: c_int, blend_factor: f32, batch: u32, d_fast_hits: *const fsgpu_hit,
                                       d_fast_tie: *const u32, d_fast_counts: *const u32, n_fast_max: u32,
                                       d_quality_hits: *const fsgpu_hit, d_quality_scores: *const f32,
                                       d_quality_present: *const u8, d_quality_tie: *const u32,
                                       d_quality_counts: *const u32, n_quality_max: u32, d_out: *mut fsgpu_hit,
                                       d_out_counts: *mut u32, stream: *mut c_void) -> c_int;
    pub fn fsgpu_rrf_fuse_device(device: c_int, config: *const fsgpu_rrf_config, batch: u32, d_lex_ids: *const u64,
                                 d_lex_scores: *const f32, d_lex_tie: *const u32, d_lex_counts: *const u32,
                                 n_lex_max: u32, d_sem_hits: *const fsgpu_hit, d_sem_tie: *const u32,
                                 d_sem_counts: *const u32, n_sem_max: u32, limit: u32, offset: u32,
                                 d_out: *mut fsgpu_fused_hit, d_out_counts: *mut u32, stream: *mut c_void) -> c_int;

    // VectorIndex::search_top_k_int8_two_pass / search_top_k_4bit_two_pass with the reference's semantics
    // (search.rs:514-650, :876-946): bits = 8 or 4
    pub fn fsgpu_search_top_k_two_pass(index: *const fsgpu_index, query: *const f32, k: u32, candidate_multiplier: u32,
                                       bits: c_int, dim: u32, out: *mut fsgpu_hit, out_count: *mut u32) -> c_int;
    pub fn fsgpu_index_read_two_pass_codes(index: *const fsgpu_index, bits: c_int, out: *mut u8) -> c_int;

    // row-sharded index over the GPUs of one box, one process, no NCCL
    pub fn fsgpu_sharded_create_f16(slab: *const u16, n_rows: u64, dim: u32, tombstones: *const u8,
                                    devices: *const c_int, n_devices: c_int, opts: *const fsgpu_index_options,
                                    out: *mut *mut fsgpu_sharded) -> c_int;
    pub fn fsgpu_sharded_from_shards(shards: *const *mut fsgpu_index, n_shards: c_int, take_ownership: c_int,
                                     out: *mut *mut fsgpu_sharded) -> c_int;
    pub fn fsgpu_sharded_destroy(sharded: *mut fsgpu_sharded);
    pub fn fsgpu_sharded_shard_count(sharded: *const fsgpu_sharded) -> c_int;
    pub fn fsgpu_sharded_rows(sharded: *const fsgpu_sharded) -> u64;
    pub fn fsgpu_sharded_shard(sharded: *const fsgpu_sharded, i: c_int) -> *mut fsgpu_index;
    pub fn fsgpu_sharded_search_top_k(sharded: *mut fsgpu_sharded, queries: *const f32, batch: u32, k: u32, dim: u32,
                                      out: *mut fsgpu_hit, out_counts: *mut u32) -> c_int;
}
