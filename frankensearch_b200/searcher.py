"""Host-side mirror of `SyncTwoTierSearcher` (crates/frankensearch-fusion/src/sync_searcher.rs:281,
`search_collect` :527, `search_internal` :616-1009): pre-embedded tiered queries in, ranked results
out.  The orchestration is the reference's, statement for statement; every numeric stage (exact
scan + top-k, quality re-scoring, blend, RRF) runs in libfsgpu.so.

Honoured but not re-implemented: query admission / identity bundles, NQC adaptive
down-weighting (the caller passes the effective `semantic_weight`, i.e. the behaviour of
`with_nqc_dense_downweight_disabled`, sync_searcher.rs:474-481), explanations.  Phase 2 also reports
the rank-change diagnostics of the async searcher (searcher.rs:2405-2412): Kendall tau and
promoted / demoted / stable counts between the fast pool and the blended pool.
"""
from __future__ import annotations

import time
from dataclasses import dataclass, field
from typing import Callable, List, Optional, Sequence

import numpy as np

from .fusion import blend_two_tier, blend_two_tier_aligned, compute_rank_changes, kendall_tau, rrf_fuse
from .index import GpuVectorIndex
from .types import RrfConfig, ScoredResult, VectorHit, candidate_count


@dataclass
class TwoTierConfig:
    """The fields of `TwoTierConfig` this path reads (crates/frankensearch-core/src/config.rs:66-194)."""
    candidate_multiplier: int = 3
    rrf_k: float = 60.0
    quality_weight: float = 0.7
    fast_only: bool = False


@dataclass
class SyncSearchOutcome:
    """sync_searcher.rs: `SyncSearchOutcome { phases, final_results, metrics }`."""
    final_results: List[ScoredResult]
    initial_results: List[ScoredResult] = field(default_factory=list)
    refined: bool = False
    metrics: dict = field(default_factory=dict)


def _fused_to_scored(fused, k: int) -> List[ScoredResult]:
    # fused_hits_to_scored_results: score = rrf_score as f32 (rrf.rs:200-213)
    return [ScoredResult(f.doc_id, float(np.float32(f.rrf_score)), index=f.semantic_index,
                         fast_score=f.semantic_score, lexical_score=f.lexical_score) for f in fused[:k]]


def _vector_hits_to_scored(hits: Sequence[VectorHit], k: int, quality: bool = False) -> List[ScoredResult]:
    out, seen = [], set()
    for h in hits:  # unique_vector_hits_to_scored_results: first (= best) occurrence of a doc id wins
        if h.doc_id in seen:
            continue
        seen.add(h.doc_id)
        out.append(ScoredResult(h.doc_id, float(h.score), index=h.index,
                                fast_score=None if quality else float(h.score)))
        if len(out) == k:
            break
    return out


class GpuSyncTwoTierSearcher:
    """`lexical(fast_query, fetch) -> list[ScoredResult]` plays `SyncLexicalSearch::search_sync`
    (sync_searcher.rs:101-110).  `quality_attested` selects independent retrieval from the quality
    tier (attested FSVI v2 space identity, sync_searcher.rs:810-813) instead of re-scoring the fast
    pool (`quality_scores_for_hits`, :814-818)."""

    def __init__(self, fast_index: GpuVectorIndex, quality_index: Optional[GpuVectorIndex] = None,
                 config: Optional[TwoTierConfig] = None, *,
                 lexical: Optional[Callable[[np.ndarray, int], List[ScoredResult]]] = None,
                 rrf_lexical_weight: float = 1.0, rrf_semantic_weight: float = 1.0,
                 rrf_tiebreak: str = "LexicalThenId", quality_attested: bool = False,
                 alignment=None):
        self.fast_index = fast_index
        self.quality_index = quality_index
        self.config = config or TwoTierConfig()
        self.lexical = lexical
        self.rrf_lexical_weight = rrf_lexical_weight
        self.rrf_semantic_weight = rrf_semantic_weight
        self.rrf_tiebreak = rrf_tiebreak
        self.quality_attested = quality_attested
        self.alignment = alignment  # fast row -> quality row (two_tier.rs:404-409); None = same order

    def _rrf_config(self) -> RrfConfig:
        return RrfConfig(self.config.rrf_k, self.rrf_lexical_weight, self.rrf_semantic_weight, self.rrf_tiebreak)

    def search_collect(self, fast_query, quality_query, k: int, filter=None) -> SyncSearchOutcome:
        """sync_searcher.rs:527 / :616-1009."""
        ms = lambda t0: (time.perf_counter() - t0) * 1e3  # noqa: E731
        metrics: dict = {}
        if k == 0:  # :639-645
            return SyncSearchOutcome([], metrics=metrics)
        fq = np.ascontiguousarray(fast_query, dtype=np.float32).reshape(-1)
        if not fq.any():  # all-zero fast query (:646-653)
            metrics["zero_signal"] = "ZeroNormQuery"
            return SyncSearchOutcome([], metrics=metrics)
        fetch = max(candidate_count(k, 0, max(self.config.candidate_multiplier, 1)), k)  # :654

        t0 = time.perf_counter()
        fast_hits = self.fast_index.search_top_k(fq, fetch, filter=filter)  # exact (:1018-1026)
        metrics["phase1_vectors_searched"] = self.fast_index.record_count()
        t1 = time.perf_counter()
        lexical_hits = self.lexical(fq, fetch) if self.lexical is not None else None
        if lexical_hits is not None and filter is not None and callable(filter):
            lexical_hits = [r for r in lexical_hits if filter(r.doc_id)]  # filter_lexical_hits
        metrics["lexical_search_ms"] = ms(t1)
        metrics["lexical_candidates"] = len(lexical_hits) if lexical_hits is not None else 0
        t2 = time.perf_counter()
        if lexical_hits is None:
            initial = _vector_hits_to_scored(fast_hits, k)  # :698-710
        else:
            initial = _fused_to_scored(rrf_fuse(lexical_hits, fast_hits, k, 0, self._rrf_config()), k)  # :713-731
        metrics["rrf_fusion_ms"] = ms(t2)
        metrics["vector_search_ms"] = metrics["phase1_total_ms"] = ms(t0)

        # phase 2 needs: not fast_only, a quality tier, a quality-bound query (:765-786)
        if self.config.fast_only or self.quality_index is None or quality_query is None:
            metrics["skip_reason"] = ("fast_only" if self.config.fast_only else
                                      "quality_query_embedding_absent" if self.quality_index is not None else
                                      "quality_index_unavailable")
            return SyncSearchOutcome(initial, initial, False, metrics)
        qq = np.ascontiguousarray(quality_query, dtype=np.float32).reshape(-1)
        t3 = time.perf_counter()
        fast_scores_by_doc = {h.doc_id: h.score for h in fast_hits}
        if self.quality_attested:  # independent retrieval (:810-813)
            quality_hits = self.quality_index.search_top_k(qq, fetch, filter=filter)
            quality_scores_by_doc = {h.doc_id: h.score for h in quality_hits}
            blended = blend_two_tier(fast_hits, quality_hits, self.config.quality_weight)
            fast_index_of = {h.doc_id: h.index for h in fast_hits}
            for h in blended:  # VectorHit.index is a FAST-tier ordinal or the "no index" sentinel (:876-886)
                h.index = fast_index_of.get(h.doc_id, 0xFFFFFFFF)
        else:  # re-score the fast pool (:814-818)
            scores = self.quality_index.quality_scores_for_hits(qq, fast_hits, alignment=self.alignment)
            quality_scores_by_doc = {h.doc_id: s for h, s in zip(fast_hits, scores) if s is not None}
            blended = blend_two_tier_aligned(fast_hits, scores, self.config.quality_weight)
        metrics["phase2_vectors_searched"] = len(quality_scores_by_doc)
        metrics["quality_search_ms"] = ms(t3)
        metrics["kendall_tau"] = kendall_tau(fast_hits, blended)             # searcher.rs:2405-2412
        metrics["rank_changes"] = compute_rank_changes(fast_hits, blended)
        if lexical_hits is not None:
            refined = _fused_to_scored(rrf_fuse(lexical_hits, blended, k, 0, self._rrf_config()), k)  # :898-918
        else:
            refined = _vector_hits_to_scored(blended, k, quality=True)  # :920-928
        for r in refined:  # evidence fields keep the raw per-tier scores (:939-942)
            r.fast_score = fast_scores_by_doc.get(r.doc_id)
            r.quality_score = quality_scores_by_doc.get(r.doc_id)
        metrics["phase2_total_ms"] = ms(t3)
        return SyncSearchOutcome(refined, initial, True, metrics)
