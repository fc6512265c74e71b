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
