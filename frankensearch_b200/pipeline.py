"""Batched, device-resident two-tier search: `SyncTwoTierSearcher::search_internal`
(crates/frankensearch-fusion/src/sync_searcher.rs:616-1009) for a whole batch of pre-embedded
queries without leaving the GPU, on one GPU or row-sharded over the ranks of a process group.

Per query (statement order of the reference):

  fetch   = candidate_count(k, 0, multiplier).max(k)                      sync_searcher.rs:654
  fast    = fast_index.search_top_k(fast_query, fetch)                    :656, :1011-1039
  initial = rrf_fuse(lexical, fast, k, 0) | fast[:k]                      :698-731
  scores  = quality_scores_for_hits(quality_query, fast)                  :814-818 (two_tier.rs:1566)
  blended = blend_two_tier_aligned(fast, scores, quality_weight)          :876
  refined = rrf_fuse(lexical, blended, k, 0) | blended[:k]                :898-928

Row-sharded form (SURVEY.md 8e "Two-tier"): both tiers use the SAME contiguous row partition, so
each rank re-scores its own fast candidates on its quality shard BEFORE the exchange and ONE
all-gather per search carries `[keys | hits | quality score]`; the merge keeps the global top-`fetch`
by fast key and `fsgpu_merge_payload_device` picks up the quality score that travelled with each
survivor.  Blend and RRF then run on every rank on identical inputs (<= a few thousand entries).

Identity on the device is the global row.  Lexical ids are rows for documents that have a vector
row and any unique value >= 2^32 otherwise (include/fsgpu.h, fsgpu_rrf_fuse).  Level-4 doc-id
tie-breaks use the optional `*_tie` rank arrays; without them the row / list position decides, which
is the reference order whenever doc-id byte order equals row order.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional

import numpy as np

from . import _ffi
from ._ffi import SearchError, check
from .types import RrfConfig, candidate_count

FUSED_HIT_DTYPE = np.dtype([("rrf_score", np.float64), ("semantic_rank", np.int32), ("lexical_rank", np.int32),
                            ("semantic_row", np.uint32), ("semantic_score", np.float32),
                            ("lexical_score", np.float32), ("in_both_sources", np.uint32)])
HIT_DTYPE = np.dtype([("row", np.uint32), ("score", np.float32)])


@dataclass
class DeviceLexical:
    """Ranked lexical candidates of a batch, on the device: ids int64 [B, n_max] (row, or >= 2^32 for
    documents without a vector row), scores f32 [B, n_max], counts int32 [B] (None = all n_max),
    optional level-4 tie ranks int32 [B, n_max]."""
    ids: "object"
    scores: "object"
    counts: "object" = None
    tie: "object" = None


@dataclass
class TwoTierDeviceResult:
    fetch: int
    fast_hits: "object"        # int32 view [B, fetch, 2] of fsgpu_hit
    fast_counts: "object"      # int32 [B]
    quality_scores: "object"   # f32 [B, fetch] (None without a quality tier / query)
    quality_present: "object"  # uint8 [B, fetch]
    initial: "object"          # phase 1: uint8 [B, k, 32] fsgpu_fused_hit (lexical) or hits [B, k, 2]
    initial_counts: "object"
    blended: "object"          # int32 view [B, fetch, 2]
    blended_counts: "object"
    refined: "object"          # phase 2, same layout as `initial`; None when phase 2 did not run
    refined_counts: "object"
    fused: bool                # True: initial/refined are fsgpu_fused_hit records


def _rrf_c(cfg: Optional[RrfConfig]):
    cfg = cfg or RrfConfig()
    return _ffi.RrfConfigC(cfg.k, cfg.lexical_weight, cfg.semantic_weight, 1 if cfg.tiebreak == "Hash" else 0, 0)


class DeviceTwoTierSearcher:
    """`fast_index` / `quality_index`: GpuVectorIndex shards of this rank (same row partition, the
    `Aligned` case of two_tier.rs:404-409).  `group`: torch.distributed process group (None with an
    initialised default group = WORLD; not initialised = single GPU)."""

    def __init__(self, fast_index, quality_index=None, *, candidate_multiplier: int = 3,
                 quality_weight: float = 0.7, rrf: Optional[RrfConfig] = None, group=None):
        import torch.distributed as dist

        self.fast = fast_index
        self.quality = quality_index
        self.multiplier = max(int(candidate_multiplier), 1)
        self.quality_weight = float(quality_weight)
        self.rrf = rrf or RrfConfig()
        self._dist = dist
        self._group = group
        self._world = dist.get_world_size(group) if dist.is_initialized() else 1
        self._L = _ffi.lib()
        self._bufs = {}
        if quality_index is not None and (quality_index.row_base() != fast_index.row_base() or
                                          quality_index.record_count() != fast_index.record_count()):
            raise SearchError("InvalidConfig", "fast and quality shards must cover the same rows")

    def _buf(self, name, shape, dtype, dev):
        import torch

        key = (name, tuple(shape), dtype, dev)
        t = self._bufs.get(name)
        if t is None or t[0] != key:
            t = (key, torch.empty(shape, dtype=dtype, device=dev))
            self._bufs[name] = t
        return t[1]

    @staticmethod
    def _stream(dev):
        """The current torch stream of `dev` as a cudaStream_t (CPU tensors — the gloo plumbing test with a
        stand-in library — have none)."""
        import torch

        return torch.cuda.current_stream(dev).cuda_stream if dev.type == "cuda" else None

    def fetch_for(self, k: int) -> int:
        return max(candidate_count(k, 0, self.multiplier), k)  # sync_searcher.rs:654

    # ── phase 1 candidates (+ the quality score of each, sharded or not) ─────────────────────
    def _candidates(self, d_fast_q, d_quality_q, fetch: int):
        import torch

        L, dev = self._L, d_fast_q.device
        b, g = d_fast_q.shape[0], self._world
        s = self._stream(dev)
        want_q = self.quality is not None and d_quality_q is not None
        # one packed buffer per rank: keys [B, fetch] u64 | hits [B, fetch] (row, score) | quality f32 [B, fetch] (+pad)
        n = b * fetch
        words = 4 * n + n + (n & 1)  # int32 words, 8-byte multiple
        packed = self._buf("packed", (words,), torch.int32, dev)
        keys_p, hits_p, qual_p = packed.data_ptr(), packed.data_ptr() + 8 * n, packed.data_ptr() + 16 * n
        counts = self._buf("local_counts", (b,), torch.int32, dev)
        check(L.fsgpu_search_top_k_device(self.fast.handle, d_fast_q.data_ptr(), b, fetch, keys_p, hits_p,
                                          counts.data_ptr(), s))
        present_local = self._buf("present_local", (b, fetch), torch.uint8, dev)
        if want_q:
            check(L.fsgpu_scores_for_hits_device(self.quality.handle, d_quality_q.data_ptr(), b, hits_p, fetch, qual_p,
                                                 present_local.data_ptr(), s))
        if g == 1:
            hits = packed[2 * n:4 * n].view(b, fetch, 2)
            if not want_q:
                return hits, counts, None, None
            return hits, counts, packed[4 * n:5 * n].view(torch.float32).view(b, fetch), present_local
        gathered = self._buf("gathered", (g, words), torch.int32, dev)
        self._dist.all_gather_into_tensor(gathered.view(-1), packed, group=self._group)
        out_keys = self._buf("m_keys", (b, fetch), torch.int64, dev)
        out_hits = self._buf("m_hits", (b, fetch, 2), torch.int32, dev)
        out_counts = self._buf("m_counts", (b,), torch.int32, dev)
        base = gathered.data_ptr()
        check(L.fsgpu_merge_top_k_hits_device(dev.index or 0, base, base + 8 * n, b, g, fetch, words // 2, fetch, fetch,
                                              out_keys.data_ptr(), out_hits.data_ptr(), out_counts.data_ptr(), s))
        if not want_q:
            return out_hits, out_counts, None, None
        q_scores = self._buf("m_quality", (b, fetch), torch.float32, dev)
        q_present = self._buf("m_present", (b, fetch), torch.uint8, dev)
        check(L.fsgpu_merge_payload_device(dev.index or 0, base, base + 16 * n, b, g, fetch, words // 2, fetch, words,
                                           fetch, out_keys.data_ptr(), fetch, q_scores.data_ptr(),
                                           q_present.data_ptr(), s))
        return out_hits, out_counts, q_scores, q_present

    def _rrf(self, lexical: DeviceLexical, sem_hits, sem_counts, n_sem: int, k: int, name: str):
        import torch

        dev = sem_hits.device
        b = sem_hits.shape[0]
        out = self._buf(name, (b, k, FUSED_HIT_DTYPE.itemsize), torch.uint8, dev)
        out_counts = self._buf(name + "_counts", (b,), torch.int32, dev)
        cfg = _rrf_c(self.rrf)
        s = self._stream(dev)
        check(self._L.fsgpu_rrf_fuse_device(dev.index or 0, C.byref(cfg), b, lexical.ids.data_ptr(),
                                            lexical.scores.data_ptr(),
                                            lexical.tie.data_ptr() if lexical.tie is not None else None,
                                            lexical.counts.data_ptr() if lexical.counts is not None else None,
                                            lexical.ids.shape[1], sem_hits.data_ptr(), None, sem_counts.data_ptr(), n_sem,
                                            k, 0, out.data_ptr(), out_counts.data_ptr(), s))
        return out, out_counts

    def search_device(self, d_fast_q, d_quality_q, k: int, lexical: Optional[DeviceLexical] = None,
                      fast_only: bool = False) -> TwoTierDeviceResult:
        """All inputs are CUDA tensors on this rank's GPU; every rank passes the same queries and
        lexical lists and gets the same result.  Work is enqueued on the current torch stream."""
        import torch

        if d_fast_q.dtype != torch.float32 or not d_fast_q.is_contiguous() or d_fast_q.dim() != 2:
            raise SearchError("InvalidConfig", "fast queries must be a contiguous float32 tensor [B, dim] on this rank's GPU")
        if d_fast_q.shape[1] != self.fast.dimension():
            raise SearchError("DimensionMismatch", f"expected {self.fast.dimension()}, found {d_fast_q.shape[1]}")
        phase2 = not fast_only and self.quality is not None and d_quality_q is not None
        if phase2 and d_quality_q.shape[1] != self.quality.dimension():
            raise SearchError("DimensionMismatch", f"expected {self.quality.dimension()}, found {d_quality_q.shape[1]}")
        if k <= 0:
            raise SearchError("InvalidConfig", "k must be positive (k == 0 is answered by the host without a search)")
        dev, b = d_fast_q.device, d_fast_q.shape[0]
        fetch = self.fetch_for(k)
        hits, counts, q_scores, q_present = self._candidates(d_fast_q, d_quality_q if phase2 else None, fetch)
        fused = lexical is not None
        if fused:
            initial, initial_counts = self._rrf(lexical, hits, counts, fetch, k, "initial")
        else:
            initial, initial_counts = hits[:, :k], torch.clamp(counts, max=k)
        if not phase2:
            return TwoTierDeviceResult(fetch, hits, counts, None, None, initial, initial_counts, None, None, None, None,
                                       fused)
        blended = self._buf("blended", (b, fetch, 2), torch.int32, dev)
        blended_counts = self._buf("blended_counts", (b,), torch.int32, dev)
        s = self._stream(dev)
        check(self._L.fsgpu_blend_two_tier_device(dev.index or 0, self.quality_weight, b, hits.data_ptr(), None,
                                                  counts.data_ptr(), fetch, None, q_scores.data_ptr(),
                                                  q_present.data_ptr(), None, None, 0, blended.data_ptr(),
                                                  blended_counts.data_ptr(), s))
        if fused:
            refined, refined_counts = self._rrf(lexical, blended, blended_counts, fetch, k, "refined")
        else:
            refined, refined_counts = blended[:, :k], torch.clamp(blended_counts, max=k)
        return TwoTierDeviceResult(fetch, hits, counts, q_scores, q_present, initial, initial_counts, blended,
                                   blended_counts, refined, refined_counts, fused)


def fused_to_numpy(t) -> np.ndarray:
    """uint8 [B, k, 32] device tensor of fsgpu_fused_hit -> structured numpy array [B, k]."""
    a = t.cpu().numpy()
    return a.view(FUSED_HIT_DTYPE).reshape(a.shape[0], a.shape[1])


def hits_to_numpy(t) -> np.ndarray:
    """int32 view [B, n, 2] of fsgpu_hit -> structured numpy array [B, n]."""
    a = np.ascontiguousarray(t.cpu().numpy())
    return a.view(HIT_DTYPE).reshape(a.shape[0], a.shape[1])
