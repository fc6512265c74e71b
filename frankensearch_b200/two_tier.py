"""Host-side mirror of `TwoTierIndex` (crates/frankensearch-index/src/two_tier.rs:505): a fast tier and
an optional quality tier over the same documents, with the method set the searchers call —
`search_fast` / `search_fast_classified` (two_tier.rs:1262-1343, :1358-1390) and
`quality_scores_for_hits` (:1566-1631).

The reference's default `search_fast` is `search_top_k_int8_two_pass(query, k, 3)` — an int8 pass that
nominates `3k` rows for an exact f16 re-score — "candidate-lossless" by measurement
(two_tier.rs:1323-1342).  By default the GPU fast tier returns the EXACT top-k (its own int8 forms are exact by a
proven bound), i.e. what that two-pass returns whenever its recall is 1, and what an
explicit `SearchParams` (exact scan) returns always; `reference_two_pass=True` runs the reference's two-pass itself
(fsgpu_search_top_k_two_pass).  ANN and MRL dispatch are out of scope.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np

from ._ffi import SearchError
from .index import GpuVectorIndex
from .types import ClassifiedHits, VectorHit, ZeroSignalReason


class GpuTwoTierIndex:
    def __init__(self, fast: GpuVectorIndex, quality: Optional[GpuVectorIndex] = None, *, alignment=None,
                 reference_two_pass: bool = False):
        self._fast = fast
        self._quality = quality
        # True: `search_fast` without params is the reference's literal default, the int8 two-pass with multiplier 3
        # (two_tier.rs:1323-1342) — its candidate set bit for bit, rows lost to the multiplier included.  False
        # (default): the exact top-k, i.e. what that two-pass returns whenever its recall is 1.
        self._reference_two_pass = reference_two_pass
        self._alignment = alignment  # fast row -> quality row (two_tier.rs:404-409); None = `Aligned`
        self._last_zero_signal: Optional[str] = None

    def fast_index(self) -> GpuVectorIndex:
        return self._fast

    def quality_index(self) -> Optional[GpuVectorIndex]:
        return self._quality

    def has_quality_index(self) -> bool:
        return self._quality is not None

    def doc_count(self) -> int:
        return self._fast.record_count()

    FAST_TIER_MULT = 3  # two_tier.rs:1333

    def search_fast(self, query_vec, k: int) -> List[VectorHit]:
        """two_tier.rs:1262-1264."""
        return self.search_fast_with_params(query_vec, k, None)

    def search_fast_with_params(self, query_vec, k: int, params=None) -> List[VectorHit]:
        """two_tier.rs:1275-1343: explicit `params` select the exact scan; without them the reference runs
        `search_top_k_int8_two_pass(query, k, 3)` — reproduced literally when `reference_two_pass` is set."""
        if params is None and self._reference_two_pass:
            return self._fast.search_top_k_int8_two_pass(query_vec, k, self.FAST_TIER_MULT)
        return self._fast.search_top_k(query_vec, k)

    def search_fast_classified(self, query_vec, k: int) -> ClassifiedHits:
        """two_tier.rs:1358-1390: dimension check, k == 0, non-finite rejection, zero-norm short circuit,
        then search_fast; an empty result carries the census verdict of the fast tier."""
        q = np.ascontiguousarray(query_vec, dtype=np.float32).reshape(-1)
        if q.size != self._fast.dimension():
            raise SearchError("DimensionMismatch", f"expected {self._fast.dimension()}, found {q.size}")
        if k == 0:
            out = ClassifiedHits([], ZeroSignalReason.CALLER_REQUESTED_ZERO_K)
        elif not np.isfinite(q).all():
            raise SearchError("InvalidConfig", "query: <contains non-finite values>: query vector must be finite")
        elif not q.any():
            out = ClassifiedHits([], ZeroSignalReason.ZERO_NORM_QUERY)
        else:
            hits = self.search_fast(q, k)
            out = ClassifiedHits(hits, None if hits else self._fast.zero_signal_state().empty_result_reason(False))
        self._last_zero_signal = out.zero_signal  # note_zero_signal (two_tier.rs:1423): state, not a log line per query
        return out

    def quality_scores_for_hits(self, query_vec, hits: Sequence[VectorHit]) -> List[Optional[float]]:
        """two_tier.rs:1566-1631."""
        if self._quality is None:
            raise SearchError("InvalidConfig", "quality index is not available")
        return self._quality.quality_scores_for_hits(query_vec, hits, alignment=self._alignment)
