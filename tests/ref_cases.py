"""Literal known-answer cases replayed from the reference's own unit tests.

Each case cites the reference test it restates (paths relative to the reference checkout).  The
same tables drive the oracle tests (CPU) and the GPU parity tests, so both sides are pinned to
the reference's expectations, not just to each other.
"""
import math

NAN = float("nan")

# crates/frankensearch-index/src/search.rs — (name, rows [(doc_id, vector)], query, k, expected doc ids or None)
SCAN_CASES = [
    # top_k_orders_by_score_descending (search.rs:2116)
    dict(name="top_k_orders_by_score_descending",
         rows=[("a", [1.0, 0.0, 0.0, 0.0]), ("b", [0.8, 0.0, 0.0, 0.0]), ("c", [0.2, 0.0, 0.0, 0.0])],
         query=[1.0, 0.0, 0.0, 0.0], k=2, expect_ids=["a", "b"]),
    # ties_are_broken_by_index (search.rs:2741)
    dict(name="ties_are_broken_by_index",
         rows=[("doc-a", [1.0, 0.0, 0.0, 0.0]), ("doc-b", [1.0, 0.0, 0.0, 0.0]), ("doc-c", [1.0, 0.0, 0.0, 0.0])],
         query=[1.0, 0.0, 0.0, 0.0], k=3, expect_rows=[0, 1, 2]),
    # nan_scores_do_not_panic_and_sort_last (search.rs:2767): all scores NaN, ordered by row
    dict(name="nan_scores_do_not_panic_and_sort_last",
         rows=[("doc-a", [1.0, 0.0, 0.0, 0.0]), ("doc-b", [0.5, 0.0, 0.0, 0.0]), ("doc-c", [0.2, 0.0, 0.0, 0.0])],
         query=[NAN, 0.0, 0.0, 0.0], k=3, expect_rows=[0, 1, 2], all_nan=True),
    # k_above_record_count_returns_all_hits (search.rs:2627)
    dict(name="k_above_record_count_returns_all_hits",
         rows=[("a", [1.0, 0.0, 0.0, 0.0]), ("b", [0.5, 0.0, 0.0, 0.0])],
         query=[1.0, 0.0, 0.0, 0.0], k=10, expect_ids=["a", "b"]),
    # limit_zero_or_empty_index_returns_no_hits (search.rs:2409)
    dict(name="limit_zero_returns_no_hits",
         rows=[("a", [1.0, 0.0, 0.0, 0.0])], query=[1.0, 0.0, 0.0, 0.0], k=0, expect_ids=[]),
    # tombstoned_records_are_excluded_from_search (search.rs:2163)
    dict(name="tombstoned_records_are_excluded_from_search",
         rows=[("a", [1.0, 0.0, 0.0, 0.0]), ("b", [0.9, 0.0, 0.0, 0.0]), ("c", [0.8, 0.0, 0.0, 0.0])],
         tombstones=[True, False, False], query=[1.0, 0.0, 0.0, 0.0], k=3, expect_ids=["b", "c"]),
    # all-zero query: every product is +-0, every sum +0.0 (accumulators start at +0.0), so all
    # rows tie and the order is by row (search.rs:1673-1678)
    dict(name="zero_query_ties_by_row",
         rows=[("neg", [-1.0, 0.0, 0.0, 0.0]), ("pos", [1.0, 0.0, 0.0, 0.0])],
         query=[0.0, 0.0, 0.0, 0.0], k=2, expect_ids=["neg", "pos"]),
]

# parallel_and_sequential_paths_match (search.rs:2233): 64 rows of [rank, 0, 0, 0]
def parallel_case():
    rows = [(f"doc-{i:02}", [float(i), 0.0, 0.0, 0.0]) for i in range(64)]
    return dict(rows=rows, query=[1.0, 0.0, 0.0, 0.0], k=10)


# crates/frankensearch-index/src/simd.rs:3044 simd_matches_scalar_f16 — literal 16-element vectors
DOT_LITERAL = dict(
    query=[0.4, -0.1, 0.6, 0.2, -0.3, 0.8, 0.7, -0.5, 0.9, -0.6, 0.11, 0.25, 0.41, -0.72, 0.55, 0.31],
    stored=[-0.8, 0.7, 0.6, -0.2, 0.3, 0.9, -0.4, 0.1, 0.12, 0.21, -0.14, 0.75, -0.22, 0.35, 0.66, -0.19],
    tol=1e-6,
)
# simd.rs:2423 avx2_f16dot_matches_generic — xorshift seed and dims
DOT_XORSHIFT = dict(seed=0x13579BDF2468ACE0, dims=[1, 7, 8, 9, 16, 17, 31, 64, 100, 256, 384, 512])
# simd.rs:3113 f16_precision_error_is_bounded_for_unit_vectors
DOT_PRECISION = dict(pattern=[0.11, -0.07, 0.19, 0.02, -0.13, 0.23, 0.31, -0.17, 0.05, -0.29, 0.37, 0.41],
                     dim=384, tol=0.01)


def L(doc, score):  # lexical_hit helper of rrf.rs tests
    return (doc, score)


def S(doc, score, index=0):  # semantic_hit helper
    return (doc, index, score)


# crates/frankensearch-fusion/src/rrf.rs tests
RRF_CASES = [
    dict(name="rrf_score_formula_k60", ref="rrf.rs:1864", lexical=[L("doc-a", 10.0)], semantic=[], limit=10,
         expect=[("doc-a", 1.0 / 61.0)]),
    dict(name="rrf_score_formula_k1", ref="rrf.rs:1881", k=1.0, lexical=[], semantic=[S("first", 0.9), S("second", 0.8)],
         limit=10, expect=[("first", 0.5), ("second", 1.0 / 3.0)]),
    dict(name="rrf_score_formula_k0_is_valid", ref="rrf.rs:1898", k=0.0, lexical=[L("doc-a", 10.0)], semantic=[],
         limit=10, expect=[("doc-a", 1.0)]),
    dict(name="invalid_k_nan", ref="rrf.rs:1912", k=NAN, lexical=[L("doc-a", 10.0)], semantic=[], limit=10,
         expect=[("doc-a", 1.0 / 61.0)]),
    dict(name="invalid_k_inf", ref="rrf.rs:1912", k=math.inf, lexical=[L("doc-a", 10.0)], semantic=[], limit=10,
         expect=[("doc-a", 1.0 / 61.0)]),
    dict(name="invalid_k_negative", ref="rrf.rs:1912", k=-100.0, lexical=[L("doc-a", 10.0)], semantic=[], limit=10,
         expect=[("doc-a", 1.0 / 61.0)]),
    dict(name="document_in_both_sources_gets_summed_score", ref="rrf.rs:1933", lexical=[L("shared", 5.0)],
         semantic=[S("shared", 0.9)], limit=10, expect=[("shared", 2.0 / 61.0)], expect_both=[True]),
    dict(name="multi_source_doc_ranks_higher_than_single_source", ref="rrf.rs:1952",
         lexical=[L("shared", 5.0), L("lex-only", 4.0)], semantic=[S("shared", 0.9), S("sem-only", 0.8)], limit=10,
         expect_order=["shared", "lex-only", "sem-only"]),
    dict(name="tier_weight_semantic_2x", ref="rrf.rs:2030", w_sem=2.0, lexical=[L("lex", 1.0)], semantic=[S("sem", 0.9)],
         limit=10, expect=[("sem", 2.0 / 61.0), ("lex", 1.0 / 61.0)]),
    dict(name="tier_weight_lexical_2x", ref="rrf.rs:2030", w_lex=2.0, lexical=[L("lex", 1.0)], semantic=[S("sem", 0.9)],
         limit=10, expect_order=["lex", "sem"]),
    dict(name="tier_weight_bad_values_are_neutral", ref="rrf.rs:2030", w_sem=NAN, w_lex=-1.0, lexical=[L("lex", 1.0)],
         semantic=[S("sem", 0.9)], limit=10, expect=[("lex", 1.0 / 61.0), ("sem", 1.0 / 61.0)]),
    dict(name="tie_breaking_lexical_score_first", ref="rrf.rs:2173", lexical=[L("only-lex", 10.0)],
         semantic=[S("only-sem", 0.9)], limit=10, expect_order=["only-lex", "only-sem"]),
    dict(name="tie_breaking_doc_id_ascending", ref="rrf.rs:2196", lexical=[L("alpha", 10.0)], semantic=[S("beta", 0.9)],
         limit=10, expect_order=["alpha", "beta"]),
    dict(name="hash_tiebreak_is_symmetric_across_tiers", ref="rrf.rs:2076", tiebreak="Hash", lexical=[L("alpha", 5.0)],
         semantic=[S("beta", 0.9)], limit=10, expect_hash_order=["alpha", "beta"]),
    dict(name="duplicate_doc_id_same_source_uses_best_rank", ref="rrf.rs:2619",
         lexical=[L("dup", 10.0), L("other", 8.0), L("dup", 5.0)], semantic=[], limit=10,
         expect_score={"dup": 1.0 / 61.0}, expect_len=2),
    dict(name="both_empty_returns_empty", ref="rrf.rs:2112", lexical=[], semantic=[], limit=10, expect=[]),
    dict(name="limit_truncates_results", ref="rrf.rs:2122", lexical=[], semantic=[S("a", 0.9), S("b", 0.8), S("c", 0.7)],
         limit=2, expect_order=["a", "b"]),
    dict(name="offset_skips_results", ref="rrf.rs:2140", lexical=[], semantic=[S("a", 0.9), S("b", 0.8), S("c", 0.7)],
         limit=10, offset=1, expect_order=["b", "c"]),
    dict(name="offset_beyond_results_is_empty", ref="rrf.rs:2158", lexical=[], semantic=[S("a", 0.9)], limit=10, offset=5,
         expect=[]),
]


def H(doc, score, index):  # `hit(doc, score, index)` helper of blend.rs tests
    return (doc, index, score)


EPS = 1.1920929e-07  # f32::EPSILON, the tolerance of the blend.rs tests

# crates/frankensearch-fusion/src/blend.rs tests
BLEND_CASES = [
    dict(name="blend_factor_point_seven_matches_weighted_formula", ref="blend.rs:708",
         fast=[H("a", 1.0, 0), H("b", 0.0, 1), H("c", 2.0, 2)], quality=[H("a", 2.0, 0), H("b", 0.0, 1), H("c", 1.0, 2)],
         alpha=0.7, expect_score={"a": 0.85}),
    dict(name="alpha_one_uses_quality_only", ref="blend.rs:724", fast=[H("a", 10.0, 0), H("b", 0.0, 1)],
         quality=[H("a", 5.0, 0), H("b", 15.0, 1)], alpha=1.0, expect_score={"a": 0.0, "b": 1.0}),
    dict(name="alpha_zero_uses_fast_only", ref="blend.rs:734", fast=[H("a", 10.0, 0), H("b", 0.0, 1)],
         quality=[H("a", 5.0, 0), H("b", 15.0, 1)], alpha=0.0, expect_score={"a": 1.0, "b": 0.0}),
    dict(name="single_source_scores_are_not_penalized", ref="blend.rs:744", fast=[H("fast-only", 10.0, 0)],
         quality=[H("quality-only", 10.0, 1)], alpha=0.7, expect_score={"fast-only": 1.0, "quality-only": 1.0}),
    dict(name="equal_scores_remain_equal", ref="blend.rs:759", fast=[H("same", 1.0, 0), H("other", 1.0, 1)],
         quality=[H("same", 2.0, 0), H("other", 2.0, 1)], alpha=0.7, expect_score={"same": 1.0}),
    dict(name="non_finite_scores_are_sanitized", ref="blend.rs:768", fast=[H("nan-doc", NAN, 0), H("ok-doc", 1.0, 1)],
         quality=[], alpha=0.3, expect_finite=True),
    dict(name="ordering_prefers_higher_blended_score", ref="blend.rs:777", fast=[H("a", 10.0, 0), H("b", 1.0, 1)],
         quality=[H("a", 1.0, 0), H("b", 10.0, 1)], alpha=0.7, expect_order=["b", "a"]),
    dict(name="narrow_negative_scores_keep_order", ref="blend.rs:787",
         fast=[H("z-best", -0.88, 0), H("a-middle", -0.89, 1), H("m-worst", -0.90, 2)], quality=[], alpha=0.7,
         expect_order=["z-best", "a-middle", "m-worst"]),
    dict(name="narrow_above_one_scores_keep_order", ref="blend.rs:787",
         fast=[H("z-best", 1.02, 0), H("a-middle", 1.01, 1), H("m-worst", 1.00, 2)], quality=[], alpha=0.7,
         expect_order=["z-best", "a-middle", "m-worst"]),
    dict(name="blend_both_empty_returns_empty", ref="blend.rs:851", fast=[], quality=[], alpha=0.7, expect_len=0),
    dict(name="blend_fast_only_returns_results", ref="blend.rs:857", fast=[H("a", 1.0, 0), H("b", 0.5, 1)], quality=[],
         alpha=0.7, expect_len=2, expect_finite=True),
    dict(name="blend_quality_only_returns_results", ref="blend.rs:865", fast=[], quality=[H("a", 1.0, 0), H("b", 0.5, 1)],
         alpha=0.7, expect_len=2, expect_finite=True),
    dict(name="blend_factor_half_weights_equally", ref="blend.rs:873", fast=[H("a", 10.0, 0), H("b", 0.0, 1)],
         quality=[H("a", 0.0, 0), H("b", 10.0, 1)], alpha=0.5, expect_equal=("a", "b")),
    dict(name="non_finite_blend_factor_falls_back_to_default", ref="blend.rs:886", fast=[H("a", 1.0, 0)],
         quality=[H("a", 1.0, 0)], alpha=NAN, same_as_alpha=0.7),
]

# blend.rs:578 aligned_blend_is_bit_identical_to_materialized
BLEND_ALIGNED = dict(
    fast=[H("a", 0.90, 0), H("b", 0.70, 1), H("c", 0.50, 2), H("d", 0.30, 3), H("a", 0.20, 9), H("e", 0.10, 4)],
    scores=[0.10, None, 0.95, 0.40, 0.99, None],
    alphas=[0.0, 0.3, 0.7, 1.0, NAN],
)


# ── resident WAL rows (crates/frankensearch-index/src/search.rs:426-494, :1449-1558) ─────────
# Each scenario: main rows, then a list of steps.  ("append", doc, vec) / ("append_batch", [(doc, vec)..]) /
# ("soft_delete", doc, expect_bool) / ("search", query, k, filter_doc_ids_or_None, checks) where checks is a
# dict of: ids (exact doc ids, best-first), n (hit count), first (doc id of hit 0), first_score_abs_lt,
# absent (doc ids that must not appear), present (doc ids that must appear), wal_count.
def _rows48():
    return [(f"doc-{i:03}", [float(48 - i), 0.0, 0.0, 0.0]) for i in range(48)]


E0 = [1.0, 0.0, 0.0, 0.0]
WAL_SCENARIOS = [
    # filter_applies_to_wal_entries (search.rs:2902)
    dict(name="filter_applies_to_wal_entries", rows=[("doc-a", [1.0, 0.0, 0.0, 0.0])],
         steps=[("append", "doc-b", [0.9, 0.0, 0.0, 0.0]), ("append", "doc-c", [0.8, 0.0, 0.0, 0.0]),
                ("search", E0, 10, ["doc-b"], dict(ids=["doc-b"]))]),
    # filter_works_with_wal_and_main_combined (search.rs:2925)
    dict(name="filter_works_with_wal_and_main_combined",
         rows=[("doc-a", [1.0, 0.0, 0.0, 0.0]), ("doc-b", [0.5, 0.0, 0.0, 0.0])],
         steps=[("append", "doc-c", [0.9, 0.0, 0.0, 0.0]),
                ("search", E0, 10, ["doc-a", "doc-c"], dict(ids=["doc-a", "doc-c"]))]),
    # wal_only_search_returns_wal_entries (search.rs:2997)
    dict(name="wal_only_search_returns_wal_entries", rows=[],
         steps=[("append", "wal-a", [1.0, 0.0, 0.0, 0.0]), ("append", "wal-b", [0.5, 0.0, 0.0, 0.0]),
                ("search", E0, 5, None, dict(ids=["wal-a", "wal-b"]))]),
    # wal_entries_can_outrank_main_index (search.rs:3024)
    dict(name="wal_entries_can_outrank_main_index",
         rows=[("main-a", [0.3, 0.0, 0.0, 0.0]), ("main-b", [0.2, 0.0, 0.0, 0.0])],
         steps=[("append", "wal-top", [1.0, 0.0, 0.0, 0.0]),
                ("search", E0, 3, None, dict(ids=["wal-top", "main-a", "main-b"]))]),
    # stale_main_entry_shadowed_by_wal (search.rs:3054): the WAL row (score 0) must be the only hit
    dict(name="stale_main_entry_shadowed_by_wal", rows=[("doc-a", [1.0, 0.0])],
         steps=[("append", "doc-a", [0.0, 1.0]),
                ("search", [1.0, 0.0], 1, None, dict(n=1, first="doc-a", first_score_abs_lt=1.1920929e-07))]),
    # wal_entries_rank_correctly_against_main (lib.rs:9880)
    dict(name="wal_entries_rank_correctly_against_main", rows=[("main-mediocre", [0.5, 0.5, 0.0, 0.0])],
         steps=[("append", "wal-perfect", [1.0, 0.0, 0.0, 0.0]),
                ("search", E0, 2, None, dict(ids=["wal-perfect", "main-mediocre"]))]),
    # append_duplicate_doc_id_both_searchable (lib.rs:9907): the WAL row shadows the main row
    dict(name="append_duplicate_doc_id_shadows_main", rows=[("doc-a", [1.0, 0.0, 0.0, 0.0])],
         steps=[("append", "doc-a", [0.0, 0.0, 0.0, 1.0]),
                ("search", E0, 10, None, dict(ids=["doc-a"], wal_count=1))]),
    # soft_delete_removes_wal_only_record_and_persists (lib.rs:10066)
    dict(name="soft_delete_removes_wal_only_record", rows=[("main-0", [1.0, 0.0, 0.0, 0.0])],
         steps=[("append", "wal-only", [0.0, 1.0, 0.0, 0.0]), ("soft_delete", "wal-only", True),
                ("search", [0.0, 1.0, 0.0, 0.0], 10, None, dict(absent=["wal-only"], wal_count=0))]),
    # soft_delete_clears_pending_wal_updates_for_same_doc_id (lib.rs:10102)
    dict(name="soft_delete_clears_pending_wal_updates", rows=[("doc-a", [1.0, 0.0, 0.0, 0.0])],
         steps=[("append", "doc-a", [0.0, 1.0, 0.0, 0.0]), ("append", "doc-b", [0.0, 0.0, 1.0, 0.0]),
                ("soft_delete", "doc-a", True),
                ("search", [0.0, 1.0, 0.0, 0.0], 10, None, dict(absent=["doc-a"], present=["doc-b"], wal_count=1))]),
    # wal_entries_are_searchable_before_compaction (tests/fsvi_roundtrip.rs:510); normalised inputs
    dict(name="wal_entries_are_searchable_before_compaction", rows=[("main-doc", [1.0, 0.0, 0.0, 0.0])],
         steps=[("append", "wal-doc", [0.0, 1.0, 0.0, 0.0]),
                ("search", [0.1104315, 0.9938837, 0.0, 0.0], 2, None, dict(first="wal-doc", n=2))]),
    # an update inside one batch keeps the LAST entry of a doc id (append_batch_impl, lib.rs:2601-2611)
    dict(name="append_batch_last_entry_of_a_doc_wins", rows=[("doc-a", [0.5, 0.0, 0.0, 0.0])],
         steps=[("append_batch", [("doc-x", [0.1, 0.0, 0.0, 0.0]), ("doc-y", [0.7, 0.0, 0.0, 0.0]),
                                  ("doc-x", [0.9, 0.0, 0.0, 0.0])]),
                ("search", E0, 10, None, dict(ids=["doc-x", "doc-y", "doc-a"], wal_count=2))]),
]


# full_recall_collect_all_matches_heap_prefix_with_wal (search.rs:2688)
def wal_full_recall_case():
    return dict(rows=_rows48(),
                wal=[("wal-top", [200.0, 0.0, 0.0, 0.0]), ("wal-mid", [24.5, 0.0, 0.0, 0.0]),
                     ("wal-tail", [-1.0, 0.0, 0.0, 0.0])],
                query=E0)


# avx2_f32slicedot_matches_generic (simd.rs:2512): xorshift seed and dims of the f32 x f32 dot
DOT_F32_XORSHIFT = dict(seed=0x7C6D5E4F3A2B1908, dims=[1, 7, 8, 9, 16, 31, 32, 33, 64, 100, 256, 384, 512])


# ── `wide::f32x8::reduce_add` probe (SURVEY.md section 7 "hard parts") ─────────────────
# One 8-lane vector whose five candidate summation orders give five DIFFERENT f32 results.  A row of
# eight f16 ones dotted with it is exactly reduce_add(probe) (products are exact, the four-accumulator
# tree only adds zeros), so `dot_product_f16_f32(&[f16::ONE; 8], &REDUCE_PROBE).to_bits()` in the
# reference build names the order `fsgpu_index_options.reduce_order` must be set to.
REDUCE_PROBE = [-95027.75, -1704061.0, -10354.58203125, -14751.4765625, 35482.0, -4012.865234375, 2076639.5,
                237570.6875]
REDUCE_PROBE_BITS = {
    0: 0x48FEA198,  # FSGPU_REDUCE_HALVES_PAIRWISE    ((v0+v1)+(v2+v3)) + ((v4+v5)+(v6+v7))   521484.75
    1: 0x48FEA190,  # FSGPU_REDUCE_AVX_TREE           ((v0+v4)+(v2+v6)) + ((v1+v5)+(v3+v7))   521484.5
    2: 0x48FEA194,  # FSGPU_REDUCE_HALVES_SEQUENTIAL  (((v0+v1)+v2)+v3) + (((v4+v5)+v6)+v7)   521484.625
    3: 0x48FEA18C,  # FSGPU_REDUCE_HALVES_STRIDE2     ((v0+v2)+(v1+v3)) + ((v4+v6)+(v5+v7))   521484.375
    4: 0x48FEA18E,  # FSGPU_REDUCE_SEQUENTIAL         ((((((v0+v1)+v2)+v3)+v4)+v5)+v6)+v7     521484.4375
}
