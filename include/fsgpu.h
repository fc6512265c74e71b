/*
 * fsgpu.h — C ABI of the B200 (sm_100a) semantic-tier hot path for frankensearch.
 *
 * The reference (a Rust workspace) has no FFI for this path; its seams are inherent methods and
 * free functions.  Each entry point below names the reference interface it
 * replaces (paths relative to the reference checkout).  INTEGRATION.md shows the Rust
 * `extern "C"` binding a maintainer would add.
 *
 * Conventions
 *   - return 0 (FSGPU_OK) on success, otherwise an fsgpu_status that maps 1:1 onto a
 *     `SearchError` variant (crates/frankensearch-core/src/error.rs:12-247); the message is
 *     available from fsgpu_last_error() (thread-local).
 *   - plain pointers and sizes only; the caller owns every input and output buffer.
 *   - "host" entry points take host pointers and copy; "_device" entry points take device
 *     pointers on the index's GPU and enqueue on `stream` (a cudaStream_t passed as void*;
 *     NULL = the index's own stream, in which case the call synchronises before returning).
 *   - every entry point is thread-safe; calls on one index are serialised internally, and calls
 *     that arrive on DIFFERENT streams are ordered on the device (each waits for the previous call's
 *     work: the per-index workspaces are shared).  A `_device` call with a caller stream never blocks
 *     the host: queries the batched path cannot cover are re-run by the exact kernels on that stream.
 *   - there is no CPU fallback: without a usable CUDA device every call fails with
 *     FSGPU_ERR_SUBSYSTEM.
 *   - rows are GLOBAL row numbers: `row_base + local row`, so per-shard results from a
 *     row-sharded corpus merge into exactly the single-index answer.
 */
#ifndef FSGPU_H
#define FSGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FSGPU_ABI_VERSION 4

typedef enum fsgpu_status {
    FSGPU_OK = 0,
    FSGPU_ERR_DIMENSION_MISMATCH = 1, /* SearchError::DimensionMismatch  error.rs:82  */
    FSGPU_ERR_INVALID_CONFIG = 2,     /* SearchError::InvalidConfig      error.rs:169 */
    FSGPU_ERR_INDEX_CORRUPTED = 3,    /* SearchError::IndexCorrupted     error.rs:60  */
    FSGPU_ERR_EMBEDDING_FAILED = 4,   /* SearchError::EmbeddingFailed    error.rs:30  */
    FSGPU_ERR_CANCELLED = 5,          /* SearchError::Cancelled          error.rs:210 */
    FSGPU_ERR_SUBSYSTEM = 6,          /* SearchError::SubsystemError{subsystem:"gpu"} error.rs:235 */
    FSGPU_ERR_IO = 7                  /* SearchError::Io */
} fsgpu_status;

/* Lane order of the final 8-lane horizontal add of the reference dot kernel
 * (`wide::f32x8::reduce_add`, crates/frankensearch-index/src/simd.rs:439).  wide 1.6.1 is not
 * vendored in the reference checkout, so the order is a switch (SURVEY.md §7 "hard parts"). */
typedef enum fsgpu_reduce_order {
    FSGPU_REDUCE_HALVES_PAIRWISE = 0,   /* ((v0+v1)+(v2+v3)) + ((v4+v5)+(v6+v7))  default */
    FSGPU_REDUCE_AVX_TREE = 1,          /* ((v0+v4)+(v2+v6)) + ((v1+v5)+(v3+v7)) */
    FSGPU_REDUCE_HALVES_SEQUENTIAL = 2, /* (((v0+v1)+v2)+v3) + (((v4+v5)+v6)+v7) */
    FSGPU_REDUCE_HALVES_STRIDE2 = 3,    /* ((v0+v2)+(v1+v3)) + ((v4+v6)+(v5+v7)) */
    FSGPU_REDUCE_SEQUENTIAL = 4         /* ((((((v0+v1)+v2)+v3)+v4)+v5)+v6)+v7 */
} fsgpu_reduce_order;

typedef struct fsgpu_index fsgpu_index; /* opaque: one f16 slab shard resident on one GPU */

/* VectorHit without the doc-id string (crates/frankensearch-core/src/types.rs:88-95):
 * `row` is VectorHit.index, `score` the raw f32 dot (NaN preserved).  The host resolves
 * doc ids for the winners only, as the reference does (search.rs:1549-1553). */
typedef struct fsgpu_hit {
    uint32_t row;
    float score;
} fsgpu_hit;

typedef struct fsgpu_index_options {
    int32_t device;        /* CUDA device ordinal */
    int32_t reduce_order;  /* fsgpu_reduce_order */
    int32_t tail_fma;      /* 1: FSVI/bytes kernel tail (`mul_add`, simd.rs:440-444);
                              0: in-memory slice kernel tail (`+= a*b`, simd.rs:298-300).
                              Only matters when dim % 8 != 0. */
    int32_t slab_is_device;/* 1: `slab` is already a device pointer on `device`; it is used in
                              place (no copy), must outlive the index and must not change while
                              the index exists (the index keeps statistics and codes of it) */
    uint64_t row_base;     /* global row number of local row 0 (row-sharded corpora) */
    int32_t int8_codes;    /* 1 (default): when dim % 128 == 0 also keep the corpus as int8 codes
                              (+ n_rows * dim bytes of HBM) for the int8 forms of the scan — same
                              results, less time (see fsgpu_index_int8_ready); 0: f16 slab only */
    int32_t flags;         /* FSGPU_OPEN_* bits (0 by default) */
} fsgpu_index_options;

/* fsgpu_index_open_fsvi: the caller replays the `<path>.wal` sidecar itself (reads it, keeps the last
 * entry of each doc id, drops a stale one — VectorIndex::open, lib.rs:1833-1878 — and hands the rows to
 * fsgpu_index_set_wal).  Without this bit a sidecar with pending appends makes the open FAIL rather
 * than silently drop documents (their superseded main rows are already tombstoned in the file). */
#define FSGPU_OPEN_HOST_REPLAYS_WAL 1

/* ---- library ------------------------------------------------------------------------------ */
int fsgpu_abi_version(void);
const char* fsgpu_last_error(void);          /* thread-local, never NULL */
int fsgpu_device_count(int* out_count);
void fsgpu_index_options_default(fsgpu_index_options* opts);

/* ---- index lifetime ----------------------------------------------------------------------- */
/* Replaces InMemoryVectorIndex::from_vectors (crates/frankensearch-index/src/in_memory.rs:1667)
 * and the mmap'd slab of VectorIndex::open (crates/frankensearch-index/src/lib.rs:819):
 * `slab` is n_rows x dim little-endian IEEE f16, row-major, caller order.  `tombstones` is an
 * optional packed bitmap (bit r%8 of byte r/8 set = record flag 0x0001, lib.rs:172) of the
 * rows to skip (search.rs:1281). */
int fsgpu_index_create_f16(const uint16_t* slab, uint64_t n_rows, uint32_t dim,
                           const uint8_t* tombstones, const fsgpu_index_options* opts,
                           fsgpu_index** out);
/* Same from f32 rows, encoded on the device with round-to-nearest-even exactly like
 * encode_f32_to_f16_extend (crates/frankensearch-index/src/simd.rs:2245-2304). */
int fsgpu_index_create_f32(const float* rows, uint64_t n_rows, uint32_t dim,
                           const uint8_t* tombstones, const fsgpu_index_options* opts,
                           fsgpu_index** out);
/* Opens a reference-written FSVI file (v1 layout: crates/frankensearch-index/src/lib.rs:6-43;
 * header CRC lib.rs:6114) and uploads rows [row_start, row_start+n) (n_rows_or_0 == 0: to the end),
 * streaming the slab through pinned memory.  f16 slabs (quantization 1) get every scan form; f32 slabs
 * (quantization 0) are scored exactly with the reference's f32 kernel (search.rs:1300-1321) through the
 * score-every-row + select path.  Tombstone flags and the doc-id string table are kept on the host
 * side of the handle (fsgpu_index_doc_id).  FSVI v2 files (identity-complete artifacts,
 * lib.rs:4229-4520) are read too: the header LAYOUT and CRC are checked, the identity documents are
 * not interpreted — admitting an artifact's identity is the host's decision before it uploads it. */
int fsgpu_index_open_fsvi(const char* path, uint64_t row_start, uint64_t n_rows_or_0,
                          const fsgpu_index_options* opts, fsgpu_index** out);
void fsgpu_index_destroy(fsgpu_index* index);

uint64_t fsgpu_index_rows(const fsgpu_index* index);      /* VectorIndex::record_count */
uint32_t fsgpu_index_dim(const fsgpu_index* index);       /* VectorIndex::dimension    */
uint64_t fsgpu_index_row_base(const fsgpu_index* index);
int fsgpu_index_device(const fsgpu_index* index);
const void* fsgpu_index_device_slab(const fsgpu_index* index);
/* Optional host-side doc-id table (concatenated UTF-8 + offsets[n_rows+1]); FSVI-opened
 * indexes have one already.  Replaces VectorIndex::doc_id_at (lib.rs:3801-3824). */
int fsgpu_index_set_doc_ids(fsgpu_index* index, const uint8_t* bytes, const uint64_t* offsets);
int fsgpu_index_doc_id(const fsgpu_index* index, uint64_t global_row, const uint8_t** out_ptr,
                       uint32_t* out_len);
/* Replaces VectorIndex::vector_at_f16 (crates/frankensearch-index/src/lib.rs, raw slab rows):
 * copies local rows [row_start, row_start + n) back to the host as f16 bit patterns. */
int fsgpu_index_read_rows_f16(const fsgpu_index* index, uint64_t row_start, uint64_t n,
                              uint16_t* out_bits);
/* Replaces VectorIndex::soft_delete's effect on the scan (flag bit 0, search.rs:1281). */
int fsgpu_index_set_tombstones(fsgpu_index* index, const uint8_t* bitmap_or_null);
/* int8 forms of the scan (fsgpu_index_options.int8_codes, on by default; dim % 128 == 0; the
 * environment switch FSGPU_MMA_I8=0 turns them off at creation or per search).  The index also holds the corpus as int8 codes made by
 * the reference's corpus-wide quantiser (quantize_f16_slab_to_i8, crates/frankensearch-index/src/
 * simd.rs:1842-1859: scale = 127 / max|x|, code = clamp(round(x * scale), -127, 127)); batches with
 * k <= 32 (shards of >= 1 M rows) run on tcgen05.mma kind::i8 — half the bytes and twice the MMA rate — and one or two
 * queries through the host API take a dp4a pass over the codes.  Results are unchanged (exact):
 * the int8 score only selects a candidate superset under a proven error bound, the winners are
 * re-scored with the reference's f16 arithmetic.  `fsgpu_index_read_codes_i8` copies codes back
 * (parity with the reference quantiser); `out_scale` receives max|x| / 127. */
int fsgpu_index_int8_ready(const fsgpu_index* index);
int fsgpu_index_read_codes_i8(const fsgpu_index* index, uint64_t row_start, uint64_t n, int8_t* out_codes,
                              float* out_scale);
/* VectorIndex::is_deleted (lib.rs:2401-2406) in bulk: the current soft-delete bitmap over local rows
 * ((n_rows + 7) / 8 bytes; for an FSVI file, flag bit 0 of each record as read at open). */
int fsgpu_index_read_tombstones(const fsgpu_index* index, uint8_t* out_bitmap);
/* VectorIndex::zero_signal_state (crates/frankensearch-index/src/lib.rs:2441-2459): the census that
 * classifies an EMPTY result (ZeroSignalState, crates/frankensearch-core/src/config.rs:682-741).
 * out_state[5] = {record_count, live_count, tombstone_count, wal_count, usable_vector_count}, where a
 * usable vector is finite with a positive finite squared norm (vector_signal_usable, lib.rs:6133-6142).
 * One pass over the slab; only called when a well-formed search came back empty. */
int fsgpu_index_zero_signal_state(const fsgpu_index* index, uint64_t* out_state);

/* Replaces the resident WAL rows of VectorIndex (`wal_entries`: f32 embeddings appended since the
 * last compaction, crates/frankensearch-index/src/lib.rs:2532-2720, wal.rs:101-107).  `embeddings`
 * is host memory, [n_wal, dim] f32 in WAL order; n_wal = 0 clears.  Every search then also scores
 * the WAL rows with dot_product_f32_f32 (simd.rs:134-222), skips non-finite scores and merges the
 * rest into the same top-k (scan_wal, search.rs:1449-1475; search.rs:488-491).  WAL row w is
 * reported as hit row `virtual_base + w` — the reference's `record_count + wal_idx`
 * (search.rs:1583-1597) — so `virtual_base` must be >= row_base + n_rows and the sum must fit u32;
 * on equal scores main rows rank before WAL rows (wal.rs:557-569).  Doc-id work — main rows
 * shadowed by a WAL row of the same doc id, duplicate doc ids — stays with the host's resolve step
 * (resolve_sorted_entries, search.rs:1503-1558), as do durability and compaction.  With a filter,
 * the allow bitmap carries the WAL rows after the slab's: bit n_rows + w. */
int fsgpu_index_set_wal(fsgpu_index* index, const float* embeddings, uint32_t n_wal, uint64_t virtual_base);
uint32_t fsgpu_index_wal_rows(const fsgpu_index* index);

/* ---- measurement ---------------------------------------------------------------------------- */
/* Launch accounting for bench.py: counts every kernel this index launches and, while enabled,
 * brackets each fused scan launch with CUDA events on its stream.  `scan_bytes` is the
 * algorithmic traffic (n_rows * dim * 2 per scan launch). */
typedef struct fsgpu_profile {
    uint64_t scan_launches;
    uint64_t merge_launches;
    uint64_t other_launches;
    uint64_t scan_bytes;
    double scan_ms;        /* sum of event-timed scan launch durations (0 unless enabled) */
    uint64_t mma_launches; /* scan launches that ran on the tensor-core batched kernel */
    double mma_flops;      /* 2 * query slots * rows * dim summed over those launches */
    uint64_t redo_queries; /* queries of batched launches re-run on the exact CUDA-core kernel */
    uint64_t i8_launches;  /* scan launches that read the int8 codes (kind::i8 MMAs or the dp4a pass) */
    uint64_t pair_launches; /* tensor-core scan launches of the CTA-pair form (mma_scan_pair_kernel) */
    uint64_t quad_launches; /* ... of the two-blocks-per-CTA int8 form (mma_scan_quad_kernel) */
} fsgpu_profile;
int fsgpu_index_profile_enable(fsgpu_index* index, int on);
int fsgpu_index_profile_read(fsgpu_index* index, fsgpu_profile* out, int reset);

/* Tensor-pipe issue ceiling of this GPU for the MMA shape of the batched scan (tcgen05.mma
 * cta_group::2, M = 256 x N = 256; kind 0 = f16 -> f32, 1 = int8 -> s32): a bare issue loop, operands
 * in shared memory, no loads, no epilogue, timed with CUDA events for about `target_ms`.  Writes
 * operations per second (2 per multiply-add).  bench.py quotes the scan kernel against this figure. */
int fsgpu_measure_tensor_peak(int device, int kind, uint32_t target_ms, double* out_ops_per_s);

/* ---- exact scan + top-k -------------------------------------------------------------------- */
/* Replaces VectorIndex::search_top_k / InMemoryVectorIndex::search_top_k
 * (crates/frankensearch-index/src/search.rs:192-206, :426-494;
 *  crates/frankensearch-index/src/in_memory.rs:2555, :2651-2710) for `batch` queries at once.
 *   queries  [batch, dim] f32 row-major            dim != index dim -> DIMENSION_MISMATCH
 *   out      [batch, k] hits, best first           out_counts[b] = hits written (min(k, live rows))
 * Ordering is the reference's strict total order (search.rs:1655-1686): score_key (NaN -> -inf)
 * by f32::total_cmp descending, ties -> lower row first.  Scores are bit-identical to the
 * reference's dot kernel (simd.rs:398-446) for the configured reduce order.
 * k == 0 or an empty index -> all counts 0 (search.rs:438-440).
 * Batches of >= 3 queries (FSGPU_MMA_MIN_BATCH) with k <= 1024 on an all-finite slab whose dim is
 * a multiple of 64 take ONE tensor-core pass over the slab (tcgen05/TMA, mma_scan_kernels.cuh)
 * followed by an exact re-scoring of a provable superset of the top-k; results are identical
 * to the per-query path. */
int fsgpu_search_top_k(const fsgpu_index* index, const float* queries, uint32_t batch, uint32_t k,
                       uint32_t dim, fsgpu_hit* out, uint32_t* out_counts);

/* Device-resident form: d_queries [batch, dim] f32, d_out_keys [batch, k] packed 64-bit order
 * keys (larger = better; 0 = empty slot), d_out_hits [batch, k] (nullable), d_out_counts
 * [batch] (nullable).  Key layout: high 32 bits = ascending total-order image of score_key,
 * low 32 bits = ~global_row. */
int fsgpu_search_top_k_device(const fsgpu_index* index, const float* d_queries, uint32_t batch,
                              uint32_t k, uint64_t* d_out_keys, fsgpu_hit* d_out_hits,
                              uint32_t* d_out_counts, void* stream);

/* Status words of the most recent search on this index, for callers of the stream-asynchronous
 * `_device` entry points (they return before the GPU has run; the synchronous ones report through
 * their return value).  Waits for that call's flag read-back, then writes out_flags[4] =
 * {contract violation (also returned as FSGPU_ERR_SUBSYSTEM), reserved, reserved, queries that were
 * re-run by the exact kernels on the device (non-finite / overflowing queries, overflowed lists)}. */
int fsgpu_index_last_status(const fsgpu_index* index, uint32_t* out_flags);

/* The `filter: Option<&dyn SearchFilter>` argument of VectorIndex::search_top_k
 * (crates/frankensearch-index/src/search.rs:192-206; applied before heap admission,
 * search.rs:1329-1447): the host evaluates the filter per doc id / doc-id hash and passes a packed
 * bitmap over LOCAL rows (bit r%8 of byte r/8 set = row r may be returned; NULL = no filter),
 * followed by one bit per resident WAL row (bit n_rows + w; search.rs:1457-1465).
 * One filter per call, shared by every query of the batch.  Excluded rows never count towards
 * k, exactly like tombstones. */
int fsgpu_search_top_k_filtered(const fsgpu_index* index, const float* queries, uint32_t batch, uint32_t k,
                                uint32_t dim, const uint8_t* allow_bitmap, fsgpu_hit* out,
                                uint32_t* out_counts);
int fsgpu_search_top_k_filtered_device(const fsgpu_index* index, const float* d_queries, uint32_t batch,
                                       uint32_t k, const uint8_t* d_allow_bitmap, uint64_t* d_out_keys,
                                       fsgpu_hit* d_out_hits, uint32_t* d_out_counts, void* stream);

/* The reference's QUANTISED TWO-PASS searches with their own semantics — VectorIndex::search_top_k_int8_two_pass
 * (crates/frankensearch-index/src/search.rs:514-650; TwoTierIndex::search_fast calls it with multiplier 3,
 * two_tier.rs:1323-1342) and search_top_k_4bit_two_pass (search.rs:876-946).  `bits` = 8 or 4.  Pass 1 ranks every
 * live row by the INTEGER dot of corpus-wide-scaled codes (simd.rs:1842-1859 / :2201-2233) with the query's own codes
 * (search.rs:1610-1655) and keeps the best max(min(k * max(multiplier, 1), rows), min(k, rows)) by (score, lower
 * row); pass 2 re-scores exactly those rows with the f16 kernel and returns the top k.  A row outside the candidate
 * set is lost, as in the reference — the result is bit-identical to the reference's for every multiplier (all pass-1
 * quantities are exact integers), and equal to fsgpu_search_top_k whenever the reference's own recall is 1.  k == 0, an
 * empty index, resident WAL rows or an f32-quantised slab take the exact search (search.rs:578-586); so does
 * k > 4096.  One query, host pointers, synchronous.  The code slab of the chosen width is built on first use
 * (nibbles_slab / int8_slab, search.rs:840-858, :986-998); an index created with int8_codes = 1 and dim % 128 == 0
 * shares its resident int8 codes. */
int fsgpu_search_top_k_two_pass(const fsgpu_index* index, const float* query, uint32_t k, uint32_t candidate_multiplier,
                                int bits, uint32_t dim, fsgpu_hit* out, uint32_t* out_count);
/* The code slab those searches scan (tests: byte-identical to the reference quantisers). */
int fsgpu_index_read_two_pass_codes(const fsgpu_index* index, int bits, uint8_t* out);

/* Doc-id-hash filters evaluated on the device.  `fsgpu_index_set_doc_hashes` gives the index the
 * 8-byte FNV-1a doc-id hash of every local row (record-table field 0, lib.rs:130-174, :6120-6127);
 * an FSVI file supplies them at open.  `fsgpu_search_top_k_hashes` is VectorIndex::search_top_k with
 * a BitsetFilter (crates/frankensearch-core/src/filter.rs:330-383: a row passes iff its hash is in
 * the set; search.rs:1329-1447): `allowed_sorted` is the set, strictly ascending.  When
 * n_allowed * 50 < n_rows it takes the reference's selective arm, try_gather_filtered
 * (search.rs:33, :1114-1255) — only the rows carrying an allowed hash are scored — otherwise the
 * filtered scan; both give the same hits.  `wal_allow_bitmap` (nullable = all pass) holds one bit
 * per resident WAL row, evaluated by the host on the WAL doc ids (search.rs:1457-1465).
 * `out_used_gather` (nullable) reports which arm ran. */
int fsgpu_index_set_doc_hashes(fsgpu_index* index, const uint64_t* hashes_or_null);
int fsgpu_search_top_k_hashes(const fsgpu_index* index, const float* queries, uint32_t batch, uint32_t k,
                              uint32_t dim, const uint64_t* allowed_sorted, uint32_t n_allowed,
                              const uint8_t* wal_allow_bitmap, fsgpu_hit* out, uint32_t* out_counts,
                              int* out_used_gather);

/* Replaces merge_partial_heaps (search.rs:1704-1720) across shards: d_keys holds, per query,
 * `n_lists` lists of `k_in` keys ([batch, n_lists, k_in], 0 = empty; e.g. an all-gather of
 * per-rank results laid out rank-major is passed with `list_stride` = batch*k_in and
 * `query_stride` = k_in); d_scores (nullable, same layout) are the raw scores that travel with
 * the keys.  Writes the best k_out per query. */
int fsgpu_merge_top_k_device(int device, const uint64_t* d_keys, const float* d_scores,
                             uint32_t batch, uint32_t n_lists, uint32_t k_in, uint64_t list_stride,
                             uint64_t query_stride, uint32_t k_out, uint64_t* d_out_keys,
                             fsgpu_hit* d_out_hits, uint32_t* d_out_counts, void* stream);

/* Same merge, with the raw scores taken from the fsgpu_hit records that travelled with the keys
 * (same addressing as d_keys): lets a sharded caller all-gather ONE buffer [keys | hits] per rank. */
int fsgpu_merge_top_k_hits_device(int device, const uint64_t* d_keys, const fsgpu_hit* d_hits,
                                  uint32_t batch, uint32_t n_lists, uint32_t k_in, uint64_t list_stride,
                                  uint64_t query_stride, uint32_t k_out, uint64_t* d_out_keys,
                                  fsgpu_hit* d_out_hits, uint32_t* d_out_counts, void* stream);

/* Replaces TwoTierIndex::quality_scores_for_hits -> VectorIndex::dot_query_at
 * (crates/frankensearch-index/src/two_tier.rs:1566-1631, :1946-1973; lib.rs:3229-3239):
 * out_scores[i] = dot(row rows[i], query); rows outside this shard (or UINT32_MAX) give
 * out_present[i] = 0 (`None`). */
int fsgpu_scores_for_rows(const fsgpu_index* index, const float* query, uint32_t dim,
                          const uint32_t* rows, uint32_t n, float* out_scores,
                          uint8_t* out_present);
int fsgpu_scores_for_rows_device(const fsgpu_index* index, const float* d_queries, uint32_t batch,
                                 const uint32_t* d_rows, uint32_t n_per_query, float* d_out_scores,
                                 uint8_t* d_out_present, void* stream);

/* Same re-scoring with the rows taken from the hit records a device search wrote
 * (d_hits [batch, n_per_query]; row UINT32_MAX = empty slot -> present 0).  In a row-sharded two-tier
 * index every rank re-scores its OWN fast-tier candidates on its quality-tier shard before the
 * all-gather (both tiers share the row partition), so the quality score travels with
 * the candidate. */
int fsgpu_scores_for_hits_device(const fsgpu_index* index, const float* d_queries, uint32_t batch,
                                 const fsgpu_hit* d_hits, uint32_t n_per_query, float* d_out_scores,
                                 uint8_t* d_out_present, void* stream);

/* The value that travelled with each key through a cross-shard merge: for every key of
 * d_merged_keys [batch, k_out] (the output of fsgpu_merge_top_k*_device) the entry of `d_payload`
 * stored beside that key in the per-shard lists (same [list, query, slot] addressing as d_keys, with
 * its own strides in float units).  The per-shard lists must be best-first (descending keys,
 * 0-padded) — what every search entry point writes.  d_out_present[i] = 0 for empty slots. */
int fsgpu_merge_payload_device(int device, const uint64_t* d_keys, const float* d_payload, uint32_t batch,
                               uint32_t n_lists, uint32_t k_in, uint64_t list_stride, uint64_t query_stride,
                               uint64_t payload_list_stride, uint64_t payload_query_stride,
                               const uint64_t* d_merged_keys, uint32_t k_out, float* d_out_payload,
                               uint8_t* d_out_present, void* stream);

/* ---- row-sharded index over several GPUs, one process --------------------------------------- */
/* contiguous row shards [i*N/G, (i+1)*N/G), one per device, the exact search on every
 * shard and merge_partial_heaps (crates/frankensearch-index/src/search.rs:1704-1720) across them.
 * Because the key order is a strict total order on (score_key, global row) (search.rs:1669-1686) the
 * result is byte-identical to one index over all rows.  No NCCL, no second process: with peer access
 * every shard's search kernels store their top-k straight into the merge device's buffer over NVLink
 * (fsgpu_sharded_is_direct), otherwise one cudaMemcpyPeerAsync per shard carries it; each shard is
 * driven by its own host thread.  `devices` may name a device more than once (several shards on one GPU).
 * The two-tier pipeline and process-per-GPU deployments use the per-shard entry points above instead. */
typedef struct fsgpu_sharded fsgpu_sharded;
int fsgpu_sharded_create_f16(const uint16_t* slab, uint64_t n_rows, uint32_t dim, const uint8_t* tombstones,
                             const int* devices, int n_devices, const fsgpu_index_options* opts,
                             fsgpu_sharded** out);
/* Adopts existing shard indexes (contiguous row ranges, in order; e.g. each built from a device-resident
 * slab or opened from a row range of an FSVI file).  take_ownership != 0: destroy them with the handle. */
int fsgpu_sharded_from_shards(fsgpu_index* const* shards, int n_shards, int take_ownership, fsgpu_sharded** out);
void fsgpu_sharded_destroy(fsgpu_sharded* sharded);
int fsgpu_sharded_shard_count(const fsgpu_sharded* sharded);
uint64_t fsgpu_sharded_rows(const fsgpu_sharded* sharded);
fsgpu_index* fsgpu_sharded_shard(const fsgpu_sharded* sharded, int i);
int fsgpu_sharded_is_direct(const fsgpu_sharded* sharded, int i);
/* VectorIndex::search_top_k over the whole corpus: host queries [batch, dim], host hits [batch, k]. */
int fsgpu_sharded_search_top_k(fsgpu_sharded* sharded, const float* queries, uint32_t batch, uint32_t k,
                               uint32_t dim, fsgpu_hit* out, uint32_t* out_counts);

/* ---- fusion -------------------------------------------------------------------------------- */
typedef struct fsgpu_rrf_config { /* RrfConfig, crates/frankensearch-fusion/src/rrf.rs:25-48 */
    double k;               /* non-finite or < 0 -> 60 (rrf.rs:124-130) */
    double lexical_weight;  /* non-finite or <= 0 -> 1 (rrf.rs:92-98)   */
    double semantic_weight;
    int32_t tiebreak;       /* 0 LexicalThenId, 1 Hash (rrf.rs:52-65) */
    int32_t reserved;
} fsgpu_rrf_config;

typedef struct fsgpu_fused_hit { /* FusedHit, crates/frankensearch-core/src/types.rs:3892-3925 */
    double rrf_score;
    int32_t semantic_rank;  /* position in the semantic list, -1 = None */
    int32_t lexical_rank;   /* first position in the lexical list, -1 = None */
    uint32_t semantic_row;  /* semantic_index, UINT32_MAX = None */
    float semantic_score;
    float lexical_score;
    uint32_t in_both_sources;
} fsgpu_fused_hit;

/* Replaces rrf_fuse / fuse_by_strategy_for_vector_lane(Rrf) -> rrf_fuse_merge_inner
 * (crates/frankensearch-fusion/src/rrf.rs:282-320, :970-1004, :1038-1210; comparator :179-198).
 * Identity is an opaque 64-bit id per document: the host passes the vector row for documents
 * that have one and any other unique value (>= 2^32) for lexical-only documents — the same
 * doc-id -> row lookup the reference index already offers (lib.rs:3317).  `*_tie` are the
 * level-4 tie-break ranks (smaller first; for LexicalThenId the byte-wise rank of the doc id,
 * for Hash the rank of (fnv1a(doc_id), doc_id)); NULL = use the id itself.
 * One fused list per query; all arrays are [batch, n_*] row-major on the host. */
int fsgpu_rrf_fuse(int device, const fsgpu_rrf_config* config, uint32_t batch,
                   const uint64_t* lex_ids, const float* lex_scores, const uint32_t* lex_tie,
                   const uint32_t* lex_counts, uint32_t n_lex_max,
                   const uint32_t* sem_rows, const float* sem_scores, const uint32_t* sem_tie,
                   const uint32_t* sem_counts, uint32_t n_sem_max,
                   uint32_t limit, uint32_t offset, fsgpu_fused_hit* out /*[batch, limit]*/,
                   uint32_t* out_counts);
int fsgpu_rrf_fuse_device(int device, const fsgpu_rrf_config* config, uint32_t batch,
                          const uint64_t* d_lex_ids, const float* d_lex_scores,
                          const uint32_t* d_lex_tie, const uint32_t* d_lex_counts,
                          uint32_t n_lex_max, const fsgpu_hit* d_sem_hits, const uint32_t* d_sem_tie,
                          const uint32_t* d_sem_counts, uint32_t n_sem_max, uint32_t limit,
                          uint32_t offset, fsgpu_fused_hit* d_out, uint32_t* d_out_counts,
                          void* stream);

/* Replaces blend_two_tier / blend_two_tier_aligned(_unique)
 * (crates/frankensearch-fusion/src/blend.rs:107-191, :213-286, :296-338).
 * aligned form: quality_scores[i] / quality_present[i] belong to fast hit i.
 * union form (quality_rows != NULL): a separately retrieved quality list, joined by row.
 * `*_tie`: doc-id byte-order ranks for the final tie-break (NULL = row order).
 * out [n_fast + n_quality] hits (row = VectorHit.index kept from the fast list when present). */
int fsgpu_blend_two_tier(int device, float blend_factor,
                         const uint32_t* fast_rows, const float* fast_scores,
                         const uint32_t* fast_tie, uint32_t n_fast,
                         const uint32_t* quality_rows, const float* quality_scores,
                         const uint8_t* quality_present, const uint32_t* quality_tie,
                         uint32_t n_quality, fsgpu_hit* out, uint32_t* out_count);

/* Batched, device-resident form (one CTA per query; phase 2 of SyncTwoTierSearcher::search_internal,
 * crates/frankensearch-fusion/src/sync_searcher.rs:876-886, for a whole batch without leaving the
 * GPU).  d_fast_hits [batch, n_fast_max] with d_fast_counts[b] filled slots (NULL = all).
 *   aligned form (d_quality_hits == NULL): d_quality_scores / d_quality_present [batch, n_fast_max]
 *     belong to the fast hits slot by slot (blend_two_tier_aligned, blend.rs:213-286);
 *   union form: d_quality_hits [batch, n_quality_max] (+ d_quality_counts) is a separately retrieved
 *     quality list joined by row (blend_two_tier, blend.rs:107-191).
 * d_out [batch, n_fast_max (+ n_quality_max in the union form)], best first, empty slots row
 * UINT32_MAX; d_out_counts [batch]. */
int fsgpu_blend_two_tier_device(int device, float blend_factor, uint32_t batch,
                                const fsgpu_hit* d_fast_hits, const uint32_t* d_fast_tie,
                                const uint32_t* d_fast_counts, uint32_t n_fast_max,
                                const fsgpu_hit* d_quality_hits, const float* d_quality_scores,
                                const uint8_t* d_quality_present, const uint32_t* d_quality_tie,
                                const uint32_t* d_quality_counts, uint32_t n_quality_max,
                                fsgpu_hit* d_out, uint32_t* d_out_counts, void* stream);

/* ---- query encoders ------------------------------------------------------------------------ */
typedef struct fsgpu_potion fsgpu_potion;
/* Replaces Model2VecEmbedder (crates/frankensearch-embed/src/model2vec_embedder.rs:67):
 * `table` is the [vocab, dim] f32 static embedding matrix (model.safetensors "embeddings"). */
int fsgpu_potion_create(const float* table, uint64_t vocab, uint32_t dim, int device,
                        fsgpu_potion** out);
void fsgpu_potion_destroy(fsgpu_potion* enc);
/* Replaces Model2VecEmbedder::embed_token_ids + finish_mean_pool_and_normalize
 * (model2vec_embedder.rs:312-335, :435-451; embed/src/simd.rs:74-116): token ids of query b are
 * ids[offsets[b] .. offsets[b+1]); tokenisation stays on the host.  out [batch, dim]. */
int fsgpu_potion_embed(const fsgpu_potion* enc, const uint32_t* ids, const uint64_t* offsets,
                       uint32_t batch, float* out);
int fsgpu_potion_embed_device(const fsgpu_potion* enc, const uint32_t* d_ids,
                              const uint64_t* d_offsets, uint32_t batch, float* d_out,
                              void* stream);

/* MiniLM-L6-v2 (BERT) query encoder.  Replaces FastEmbedEmbedder::embed / embed_batch
 * (crates/frankensearch-embed/src/fastembed_embedder.rs:317-398, :416-426): BERT forward with the
 * architecture stated at crates/frankensearch-rerank/src/native.rs:36-45 (H = 384, 12 heads x 32,
 * erf-GELU, post-LN), attention-mask mean pool, L2 (eps 1e-12), adapter L2 with zero-vector guard
 * (crates/frankensearch-embed/src/model_manifest.rs:300-304).  Tokenisation ([CLS] ... [SEP],
 * truncation to 512, batch-longest padding; model_manifest.rs:74-80) stays on the host.
 * All weights are f32 host arrays in PyTorch layout ([out, in] row-major for linear layers);
 * qkv_w is the row-wise concatenation query | key | value. */
typedef struct fsgpu_minilm fsgpu_minilm;
typedef struct fsgpu_minilm_layer_weights {
    const float *qkv_w, *qkv_b;             /* [3H, H], [3H] */
    const float *attn_out_w, *attn_out_b;   /* [H, H], [H]   */
    const float *attn_ln_g, *attn_ln_b;     /* [H]           */
    const float *ffn_in_w, *ffn_in_b;       /* [I, H], [I]   */
    const float *ffn_out_w, *ffn_out_b;     /* [H, I], [H]   */
    const float *ffn_ln_g, *ffn_ln_b;       /* [H]           */
} fsgpu_minilm_layer_weights;
typedef struct fsgpu_minilm_weights {
    uint32_t vocab_size, max_positions, n_layers, hidden, heads, intermediate;
    float ln_eps;                            /* 1e-12 */
    uint32_t reserved;
    const float *word_emb, *pos_emb, *type_emb; /* [vocab, H], [max_positions, H], [>=1, H] (row 0 used) */
    const float *emb_ln_g, *emb_ln_b;        /* [H] */
    const fsgpu_minilm_layer_weights* layers; /* [n_layers] */
} fsgpu_minilm_weights;
typedef struct fsgpu_minilm_profile {
    uint64_t gemm_launches, other_launches;
    double gemm_flops;     /* 2*M*N*K per tensor-core product issued, summed; M = batch * max_len (the f16 form packs
                              its rows on the device, sum(len) of them: scale by sum(len) / (batch * max_len)) */
    double gemm_ms;        /* event-timed GEMM durations (0 unless enabled) */
} fsgpu_minilm_profile;
int fsgpu_minilm_create(const fsgpu_minilm_weights* weights, int device, fsgpu_minilm** out);
/* Same from a `model.safetensors` file (Hugging Face BertModel tensor names, optionally prefixed
 * "bert." / "0.auto_model."; F32, F16 or BF16): the weights file of sentence-transformers/
 * all-MiniLM-L6-v2 that the reference pins (crates/frankensearch-embed/src/model_manifest.rs:343-349).
 * heads = 12 and layer_norm_eps = 1e-12 are the architecture's (crates/frankensearch-rerank/src/
 * native.rs:36-45). */
int fsgpu_minilm_load(const char* safetensors_path, int device, fsgpu_minilm** out);
void fsgpu_minilm_destroy(fsgpu_minilm* enc);
/* ids [batch, max_len] int32 (slots >= lens[b] ignored), lens [batch] (0 = empty text -> zero
 * vector, fastembed_embedder.rs:432-434); out [batch, hidden] f32, L2-normalised. */
int fsgpu_minilm_embed(const fsgpu_minilm* enc, const int32_t* ids, const int32_t* lens, uint32_t batch,
                       uint32_t max_len, float* out);
/* Device-resident form, asynchronous on `stream` (NULL: the encoder's own stream, synchronised before returning).
 * Calls on one encoder are serialised; its activation buffers are shared, so a call on another stream first waits (on
 * the device) for the previous call's kernels.  A caller that CAPTURES `stream` into a CUDA graph must itself order
 * the replays of one encoder (the library cannot see them). */
int fsgpu_minilm_embed_device(const fsgpu_minilm* enc, const int32_t* d_ids, const int32_t* d_lens,
                              uint32_t batch, uint32_t max_len, float* d_out, void* stream);
int fsgpu_minilm_profile_enable(fsgpu_minilm* enc, int on);
int fsgpu_minilm_profile_read(fsgpu_minilm* enc, fsgpu_minilm_profile* out, int reset);

/* ---- synthetic corpora (bench / test utility) ---------------------------------------------- */
/* The reference's bench generators on the device
 * (crates/frankensearch-index/benches/fsvi_int8_two_pass.rs:199-231), bit-identical to the
 * oracle's fso_synth_rows: kind 0 uniform, 1 clustered.  Writes n_rows x dim f16 at d_out. */
int fsgpu_synth_rows_device(int device, int kind, uint64_t seed_base, uint64_t row_start,
                            uint64_t n_rows, uint32_t dim, uint32_t n_centroids, float noise,
                            uint16_t* d_out_f16, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FSGPU_H */
