"""Host-side mirror of the reference fusion free functions over the C ABI.

`rrf_fuse` (crates/frankensearch-fusion/src/rrf.rs:282-320 -> :1038-1210) and
`blend_two_tier*` (crates/frankensearch-fusion/src/blend.rs:107-191, :213-286, :296-338).
The f64/f32 arithmetic and the ranking sorts run in libfsgpu.so on the GPU; this file only turns
doc-id strings into the integer ids / tie-break ranks the kernels join and order on.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import _ffi
from ._ffi import check, ptr
from .types import FusedHit, RrfConfig, ScoredResult, VectorHit, fnv1a_hash

_HIT_DT = np.dtype([("row", np.uint32), ("score", np.float32)])
_FUSED_DT = np.dtype([("rrf_score", np.float64), ("semantic_rank", np.int32), ("lexical_rank", np.int32),
                      ("semantic_row", np.uint32), ("semantic_score", np.float32),
                      ("lexical_score", np.float32), ("in_both_sources", np.uint32)])


def _dense_ids(doc_ids: Sequence[str]) -> Dict[str, int]:
    ids: Dict[str, int] = {}
    for d in doc_ids:
        if d not in ids:
            ids[d] = len(ids)
    return ids


def _tie_ranks(ids: Dict[str, int], tiebreak: str) -> Dict[str, int]:
    if tiebreak == "Hash":  # rrf.rs:191-196: fnv1a(doc_id), then doc_id
        order = sorted(ids, key=lambda d: (fnv1a_hash(d.encode("utf-8")), d.encode("utf-8")))
    else:                   # byte-wise doc_id order (Rust str::cmp)
        order = sorted(ids, key=lambda d: d.encode("utf-8"))
    return {d: r for r, d in enumerate(order)}


def rrf_fuse(lexical: Sequence[ScoredResult], semantic: Sequence[VectorHit], limit: int, offset: int = 0,
             config: Optional[RrfConfig] = None, *, device: int = 0) -> List[FusedHit]:
    """rrf_fuse(lexical, semantic, limit, offset, config) -> Vec<FusedHit> (rrf.rs:282)."""
    config = config or RrfConfig()
    if limit + offset == 0 or (not lexical and not semantic):
        return []
    ids = _dense_ids([r.doc_id for r in lexical] + [h.doc_id for h in semantic])
    tie = _tie_ranks(ids, config.tiebreak)
    n_lex, n_sem = len(lexical), len(semantic)
    lex_ids = np.array([ids[r.doc_id] for r in lexical], dtype=np.uint64)
    lex_scores = np.array([r.score for r in lexical], dtype=np.float32)
    lex_tie = np.array([tie[r.doc_id] for r in lexical], dtype=np.uint32)
    sem_rows = np.array([ids[h.doc_id] for h in semantic], dtype=np.uint32)
    sem_scores = np.array([h.score for h in semantic], dtype=np.float32)
    sem_tie = np.array([tie[h.doc_id] for h in semantic], dtype=np.uint32)
    lex_counts = np.array([n_lex], dtype=np.uint32)
    sem_counts = np.array([n_sem], dtype=np.uint32)
    cfg = _ffi.RrfConfigC(float(config.k), float(config.lexical_weight), float(config.semantic_weight),
                          1 if config.tiebreak == "Hash" else 0, 0)
    out = np.zeros(max(limit, 1), dtype=_FUSED_DT)
    out_counts = np.zeros(1, dtype=np.uint32)
    check(_ffi.lib().fsgpu_rrf_fuse(device, C.byref(cfg), 1, ptr(lex_ids), ptr(lex_scores), ptr(lex_tie),
                                    ptr(lex_counts), max(n_lex, 1) if n_lex else 0, ptr(sem_rows),
                                    ptr(sem_scores), ptr(sem_tie), ptr(sem_counts), n_sem, limit, offset,
                                    ptr(out), ptr(out_counts)))
    fused = []
    for i in range(int(out_counts[0])):
        o = out[i]
        sr, lr = int(o["semantic_rank"]), int(o["lexical_rank"])
        doc = semantic[sr].doc_id if sr >= 0 else lexical[lr].doc_id
        fused.append(FusedHit(
            doc_id=doc, rrf_score=float(o["rrf_score"]),
            lexical_rank=lr if lr >= 0 else None, semantic_rank=sr if sr >= 0 else None,
            semantic_index=semantic[sr].index if sr >= 0 else None,
            lexical_score=float(np.float32(lexical[lr].score)) if lr >= 0 else None,
            semantic_score=float(np.float32(semantic[sr].score)) if sr >= 0 else None,
            in_both_sources=bool(o["in_both_sources"])))
    return fused


def _blend(fast: Sequence[VectorHit], quality_rows, quality_scores, quality_present, quality_tie,
           n_quality: int, blend_factor: float, ids: Dict[str, int], tie: Dict[str, int], device: int):
    fast_rows = np.array([ids[h.doc_id] for h in fast], dtype=np.uint32)
    fast_scores = np.array([h.score for h in fast], dtype=np.float32)
    fast_tie = np.array([tie[h.doc_id] for h in fast], dtype=np.uint32)
    out = np.zeros(max(len(fast) + (n_quality if quality_rows is not None else 0), 1), dtype=_HIT_DT)
    cnt = C.c_uint32(0)
    check(_ffi.lib().fsgpu_blend_two_tier(device, float(blend_factor), ptr(fast_rows), ptr(fast_scores),
                                          ptr(fast_tie), len(fast), ptr(quality_rows), ptr(quality_scores),
                                          ptr(quality_present), ptr(quality_tie), n_quality, ptr(out),
                                          C.byref(cnt)))
    return out[: cnt.value]


def blend_two_tier(fast_results: Sequence[VectorHit], quality_results: Sequence[VectorHit],
                   blend_factor: float, *, device: int = 0) -> List[VectorHit]:
    """blend_two_tier (blend.rs:107-191): union of both tiers by doc id."""
    if not fast_results and not quality_results:
        return []
    ids = _dense_ids([h.doc_id for h in fast_results] + [h.doc_id for h in quality_results])
    tie = _tie_ranks(ids, "LexicalThenId")
    q_rows = np.array([ids[h.doc_id] for h in quality_results], dtype=np.uint32)
    q_scores = np.array([h.score for h in quality_results], dtype=np.float32)
    q_tie = np.array([tie[h.doc_id] for h in quality_results], dtype=np.uint32)
    if len(quality_results) == 0:  # union form with an empty quality list == aligned form, no scores
        out = _blend(fast_results, None, None, None, None, 0, blend_factor, ids, tie, device)
    else:
        out = _blend(fast_results, q_rows, q_scores, None, q_tie, len(quality_results), blend_factor, ids,
                     tie, device)
    first_index: Dict[str, int] = {}
    for h in list(fast_results) + list(quality_results):  # pair.index: first fast occurrence, else quality
        first_index.setdefault(h.doc_id, h.index)
    names = list(ids)
    return [VectorHit(first_index[names[int(o["row"])]], float(o["score"]), names[int(o["row"])]) for o in out]


def blend_two_tier_aligned(fast_hits: Sequence[VectorHit], quality_scores: Sequence[Optional[float]],
                           blend_factor: float, *, device: int = 0) -> List[VectorHit]:
    """blend_two_tier_aligned / _aligned_unique (blend.rs:213-286, :296-338): `quality_scores[i]`
    is the optional quality-tier score of `fast_hits[i]`."""
    if not fast_hits:
        return []
    ids = _dense_ids([h.doc_id for h in fast_hits])
    tie = _tie_ranks(ids, "LexicalThenId")
    present = np.array([i < len(quality_scores) and quality_scores[i] is not None
                        for i in range(len(fast_hits))], dtype=np.uint8)
    qs = np.array([quality_scores[i] if present[i] else 0.0 for i in range(len(fast_hits))], dtype=np.float32)
    out = _blend(fast_hits, None, qs, present, None, len(fast_hits), blend_factor, ids, tie, device)
    first_index: Dict[str, int] = {}
    for h in fast_hits:
        first_index.setdefault(h.doc_id, h.index)
    names = list(ids)
    return [VectorHit(first_index[names[int(o["row"])]], float(o["score"]), names[int(o["row"])]) for o in out]


# ── phase-2 diagnostics: how much the quality tier reordered the fast tier ──────────────────
# Host-side integer work on the two result lists (crates/frankensearch-fusion/src/blend.rs:365-544);
# the searcher reports them with the refined results (searcher.rs:2560-2617).
class RankChanges:
    """blend.rs RankChanges: promoted / demoted / stable counts."""

    __slots__ = ("promoted", "demoted", "stable")

    def __init__(self, promoted: int = 0, demoted: int = 0, stable: int = 0):
        self.promoted, self.demoted, self.stable = promoted, demoted, stable

    def __eq__(self, other):
        return (self.promoted, self.demoted, self.stable) == (other.promoted, other.demoted, other.stable)

    def __repr__(self):
        return f"RankChanges(promoted={self.promoted}, demoted={self.demoted}, stable={self.stable})"


def build_rank_map(hits: Sequence[VectorHit]) -> Dict[str, int]:
    """build_borrowed_rank_map (blend.rs:538-544): doc id -> rank, first occurrence wins."""
    ranks: Dict[str, int] = {}
    for rank, h in enumerate(hits):
        ranks.setdefault(h.doc_id, rank)
    return ranks


def compute_rank_changes(initial: Sequence[VectorHit], refined: Sequence[VectorHit]) -> RankChanges:
    """compute_rank_changes (blend.rs:365-409): promoted = rank improved or new in `refined`,
    demoted = rank worsened or dropped from `refined`, stable = unchanged."""
    a, b = build_rank_map(initial), build_rank_map(refined)
    out = RankChanges()
    for doc, old in a.items():
        new = b.get(doc)
        if new is None or new > old:
            out.demoted += 1
        elif new < old:
            out.promoted += 1
        else:
            out.stable += 1
    out.promoted += sum(1 for doc in b if doc not in a)
    return out


def _count_inversions(values: List[int]) -> int:
    """merge_sort_inversions (blend.rs:461-516): pairs i < j with values[i] > values[j]."""
    n = len(values)
    if n <= 1:
        return 0
    mid = n // 2
    left, right = values[:mid], values[mid:]
    count = _count_inversions(left) + _count_inversions(right)
    i = j = 0
    merged = []
    while i < len(left) and j < len(right):
        if left[i] <= right[j]:
            merged.append(left[i])
            i += 1
        else:
            merged.append(right[j])
            count += len(left) - i
            j += 1
    merged.extend(left[i:])
    merged.extend(right[j:])
    values[:] = merged
    return count


def kendall_tau(initial: Sequence[VectorHit], refined: Sequence[VectorHit]) -> Optional[float]:
    """kendall_tau (blend.rs:411-459): rank correlation over the doc ids common to both lists, in
    `initial` order (first occurrence of a doc id only); None with fewer than two common docs."""
    refined_rank = build_rank_map(refined)
    seen, ranks = set(), []
    for h in initial:
        r = refined_rank.get(h.doc_id)
        if r is not None and h.doc_id not in seen:
            seen.add(h.doc_id)
            ranks.append(r)
    n = len(ranks)
    if n < 2:
        return None
    total = n * (n - 1) // 2
    discordant = _count_inversions(ranks)
    concordant = max(total - discordant, 0)
    return (float(concordant) - float(discordant)) / float(total)
