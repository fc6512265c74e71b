"""Boundary types, named after the reference's (crates/frankensearch-core/src/types.rs)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional


@dataclass
class VectorHit:
    """types.rs:88-95 — raw hit of the vector index: row, raw f32 dot, resolved doc id."""
    index: int
    score: float
    doc_id: Optional[str] = None


@dataclass
class ScoredResult:
    """types.rs:4004-4035 — only the fields the fusion path reads."""
    doc_id: str
    score: float
    index: Optional[int] = None
    fast_score: Optional[float] = None
    quality_score: Optional[float] = None
    lexical_score: Optional[float] = None


@dataclass
class FusedHit:
    """types.rs:3892-3925."""
    doc_id: str
    rrf_score: float
    lexical_rank: Optional[int]
    semantic_rank: Optional[int]
    semantic_index: Optional[int]
    lexical_score: Optional[float]
    semantic_score: Optional[float]
    in_both_sources: bool


@dataclass
class RrfConfig:
    """crates/frankensearch-fusion/src/rrf.rs:25-48, defaults :77-86."""
    k: float = 60.0
    lexical_weight: float = 1.0
    semantic_weight: float = 1.0
    tiebreak: str = "LexicalThenId"  # or "Hash" (rrf.rs:52-65)


def candidate_count(limit: int, offset: int, multiplier: int) -> int:
    """rrf.rs:113-115."""
    return (limit + offset) * multiplier


def fnv1a_hash(data: bytes) -> int:
    """FNV-1a 64 of the doc id: FSVI row order key (lib.rs:6120-6127) and the RRF `Hash`
    tie-break (rrf.rs:68-75).  Host-side id bookkeeping, not scan arithmetic."""
    h = 0xCBF29CE484222325
    for b in data:
        h ^= b
        h = (h * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
    return h


# ─── candidate budgets (host arithmetic of the searcher) ─────────────────────────────────────
def scaled_budget(base_candidates: int, multiplier: float) -> int:
    """crates/frankensearch-fusion/src/searcher.rs:120-126: ceil(base * multiplier) in f32, at least 1;
    0 for an empty base or a non-finite / non-positive multiplier."""
    import math

    import numpy as np

    m = np.float32(multiplier)
    if base_candidates == 0 or not math.isfinite(float(m)) or m <= 0:
        return 0
    return max(int(math.ceil(float(np.float32(np.float32(base_candidates) * m)))), 1)


class QueryClass:
    """crates/frankensearch-core/src/query_class.rs:24-215: heuristic query classes and the candidate
    budget multipliers they imply (identifiers lean lexical, natural language leans semantic)."""
    EMPTY, IDENTIFIER, SHORT_KEYWORD, NATURAL_LANGUAGE = "empty", "identifier", "short_keyword", "natural_language"
    _LEXICAL = {EMPTY: 0.0, IDENTIFIER: 2.0, SHORT_KEYWORD: 1.0, NATURAL_LANGUAGE: 0.5}   # :197-205
    _SEMANTIC = {EMPTY: 0.0, IDENTIFIER: 0.5, SHORT_KEYWORD: 1.0, NATURAL_LANGUAGE: 2.0}  # :208-215

    @staticmethod
    def lexical_budget_multiplier(cls_: str) -> float:
        return QueryClass._LEXICAL[cls_]

    @staticmethod
    def semantic_budget_multiplier(cls_: str) -> float:
        return QueryClass._SEMANTIC[cls_]

    @staticmethod
    def _looks_like_identifier(s: str) -> bool:  # query_class.rs:68-190
        if not any(c.isspace() for c in s):
            if "/" in s or "\\" in s or "." in s or "::" in s or "_" in s:
                return True
            has_lower = any(c.islower() for c in s)
            has_upper = any(c.isupper() for c in s)
            first_upper = s[0].isupper()
            rest_lower = all(c.islower() for c in s[1:])
            if has_lower and has_upper and not (first_upper and rest_lower):
                return True
            if "-" in s:
                prefix, suffix = s.rsplit("-", 1)
                if prefix and suffix and all(c in "0123456789" for c in suffix) and \
                        all((c.isascii() and c.isalnum()) or c in "-_" for c in prefix):
                    return True
        return s.startswith("fn ") or s.startswith("struct ") or s.startswith("impl ")

    @staticmethod
    def classify(query: str) -> str:  # query_class.rs:47-66
        trimmed = query.strip()
        if not trimmed:
            return QueryClass.EMPTY
        if QueryClass._looks_like_identifier(trimmed):
            return QueryClass.IDENTIFIER
        return QueryClass.SHORT_KEYWORD if len(trimmed.split()) <= 3 else QueryClass.NATURAL_LANGUAGE


def phase1_budgets(k: int, candidate_multiplier: int, query_class: str):
    """crates/frankensearch-fusion/src/searcher.rs:1599-1608: (base, semantic fetch, lexical fetch) —
    base = candidate_count(k, 0, max(multiplier, 1)); each lane fetches scaled_budget(base, class multiplier)."""
    base = candidate_count(k, 0, max(int(candidate_multiplier), 1))
    return (base, scaled_budget(base, QueryClass.semantic_budget_multiplier(query_class)),
            scaled_budget(base, QueryClass.lexical_budget_multiplier(query_class)))


# ─── typed empty results (crates/frankensearch-core/src/config.rs:579-741) ───────────────────
class ZeroSignalReason:
    CALLER_REQUESTED_ZERO_K = "caller_requested_zero_k"
    FILTER_ELIMINATED_ALL = "filter_eliminated_all"
    NON_FINITE_QUERY = "non_finite_query"
    ZERO_NORM_QUERY = "zero_norm_query"
    NEWLY_CREATED_EMPTY = "newly_created_empty"
    ALL_TOMBSTONED = "all_tombstoned"
    WAL_ONLY_NO_LIVE_RECORDS = "wal_only_no_live_records"
    NO_USABLE_VECTORS = "no_usable_vectors"
    ANN_RETURNED_EMPTY_DESPITE_USABLE_VECTORS = "ann_returned_empty_despite_usable_vectors"


@dataclass
class ZeroSignalState:
    """config.rs:682-741."""
    record_count: int = 0
    live_count: int = 0
    tombstone_count: int = 0
    wal_count: int = 0
    usable_vector_count: int = 0

    def state_reason(self) -> Optional[str]:
        if self.record_count == 0 and self.wal_count == 0:
            return ZeroSignalReason.NEWLY_CREATED_EMPTY
        if self.live_count == 0 and self.wal_count == 0:
            return ZeroSignalReason.ALL_TOMBSTONED
        if self.live_count > 0 and self.usable_vector_count == 0:
            return ZeroSignalReason.NO_USABLE_VECTORS
        return None

    def is_wal_only(self) -> bool:
        return self.live_count == 0 and self.wal_count > 0

    def empty_result_reason(self, had_filter: bool) -> str:
        reason = self.state_reason()
        if reason is not None:
            return reason
        if had_filter:
            return ZeroSignalReason.FILTER_ELIMINATED_ALL
        if self.is_wal_only():
            return ZeroSignalReason.WAL_ONLY_NO_LIVE_RECORDS
        return ZeroSignalReason.NO_USABLE_VECTORS


@dataclass
class ClassifiedHits:
    """crates/frankensearch-index/src/search.rs:63-86: `zero_signal is not None` iff `hits` is empty."""
    hits: list
    zero_signal: Optional[str] = None
