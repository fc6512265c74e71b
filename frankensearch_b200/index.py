"""Host-side mirror of the reference vector-index seam over the C ABI.

`GpuVectorIndex` presents the method set the reference searchers call on
`VectorIndex` / `InMemoryVectorIndex` (crates/frankensearch-index/src/search.rs:192-206,
crates/frankensearch-index/src/in_memory.rs:1667, :2555) and
`TwoTierIndex::quality_scores_for_hits` (two_tier.rs:1566-1631).  All arithmetic happens in
libfsgpu.so on the GPU; this file is plumbing (buffers, doc-id strings, error mapping).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import numpy as np

from . import _ffi
from ._ffi import SearchError, check, ptr
from .types import ClassifiedHits, VectorHit, ZeroSignalReason, ZeroSignalState, fnv1a_hash

REDUCE_ORDERS = {
    "halves_pairwise": 0, "avx_tree": 1, "halves_sequential": 2, "halves_stride2": 3, "sequential": 4,
}


def _options(device: int, reduce_order, tail_fma: bool, slab_is_device: bool, row_base: int, int8_codes=None):
    o = _ffi.IndexOptions()
    _ffi.lib().fsgpu_index_options_default(C.byref(o))
    o.device = device
    o.reduce_order = REDUCE_ORDERS[reduce_order] if isinstance(reduce_order, str) else int(reduce_order)
    o.tail_fma = 1 if tail_fma else 0
    o.slab_is_device = 1 if slab_is_device else 0
    o.row_base = row_base
    if int8_codes is not None:
        o.int8_codes = 1 if int8_codes else 0
    return o


def _bitmap(flags) -> Optional[np.ndarray]:
    if flags is None:
        return None
    flags = np.asarray(flags, dtype=bool)
    return np.packbits(flags, bitorder="little") if flags.any() else None


class GpuVectorIndex:
    """One f16 slab shard resident on one B200."""

    def __init__(self, handle: int, doc_ids: Optional[Sequence[str]] = None, dedup_doc_ids: bool = False,
                 keepalive=None):
        self._h = C.c_void_p(handle)
        self._doc_ids = list(doc_ids) if doc_ids is not None else None
        self._dedup = dedup_doc_ids
        self._keepalive = keepalive  # e.g. the torch tensor that owns an adopted device slab
        self._L = _ffi.lib()
        # mutable state of VectorIndex mirrored on the host: resident WAL rows and soft-delete flags
        self._wal: List[tuple] = []          # [(doc_id, float32[dim])] in WAL order (wal.rs:101-107)
        self._tomb: Optional[np.ndarray] = None
        self._rows_of: Optional[dict] = None  # doc id -> local rows, built on first mutation
        self._hashes_on_device = False        # record-table hashes uploaded (FSVI files bring their own)
        self.last_filter_arm: Optional[str] = None  # "gather" | "scan" | "bitmap" for the last filtered call

    # ── construction ────────────────────────────────────────────────────────────────────────
    @classmethod
    def from_vectors(cls, doc_ids: Optional[Sequence[str]], vectors, *, device: int = 0,
                     reduce_order="halves_pairwise", tail_fma: bool = False, row_base: int = 0,
                     tombstones=None) -> "GpuVectorIndex":
        """InMemoryVectorIndex::from_vectors (in_memory.rs:1667-1730): f32 rows in caller order,
        encoded to f16 with round-to-nearest-even on the device (simd.rs:2245-2304)."""
        v = np.ascontiguousarray(vectors, dtype=np.float32)
        if v.ndim != 2:
            raise SearchError("InvalidConfig", "vectors must be [n, dim]")
        if doc_ids is not None and len(doc_ids) != v.shape[0]:
            raise SearchError("InvalidConfig", "doc_ids and vectors differ in length")
        if v.size and not np.isfinite(v).all():  # write_record (lib.rs:3647-3653)
            raise SearchError("InvalidConfig", "all embedding values must be finite")
        h = C.c_void_p()
        o = _options(device, reduce_order, tail_fma, False, row_base)
        bm = _bitmap(tombstones)
        check(_ffi.lib().fsgpu_index_create_f32(ptr(v), v.shape[0], v.shape[1], ptr(bm), C.byref(o), C.byref(h)))
        ix = cls(h.value, doc_ids)
        ix._tomb = None if tombstones is None else np.asarray(tombstones, dtype=bool).copy()
        return ix

    @classmethod
    def from_f16_bits(cls, doc_ids: Optional[Sequence[str]], slab_bits, *, device: int = 0,
                      reduce_order="halves_pairwise", tail_fma: bool = True, row_base: int = 0,
                      tombstones=None) -> "GpuVectorIndex":
        """A ready f16 slab (uint16 bit patterns, [n, dim]) — the FSVI vector slab layout
        (lib.rs:6-43)."""
        s = np.ascontiguousarray(slab_bits, dtype=np.uint16)
        if s.ndim != 2:
            raise SearchError("InvalidConfig", "slab must be [n, dim]")
        h = C.c_void_p()
        o = _options(device, reduce_order, tail_fma, False, row_base)
        bm = _bitmap(tombstones)
        check(_ffi.lib().fsgpu_index_create_f16(ptr(s), s.shape[0], s.shape[1], ptr(bm), C.byref(o), C.byref(h)))
        ix = cls(h.value, doc_ids)
        ix._tomb = None if tombstones is None else np.asarray(tombstones, dtype=bool).copy()
        return ix

    @classmethod
    def from_device_tensor(cls, slab, *, doc_ids: Optional[Sequence[str]] = None,
                           reduce_order="halves_pairwise", tail_fma: bool = True, row_base: int = 0,
                           tombstones=None) -> "GpuVectorIndex":
        """Adopt a CUDA tensor [n, dim] of f16 (torch.float16 / int16 / uint16) without copying."""
        if not slab.is_cuda or slab.dim() != 2 or not slab.is_contiguous() or slab.element_size() != 2:
            raise SearchError("InvalidConfig", "slab must be a contiguous 2-byte CUDA tensor [n, dim]")
        h = C.c_void_p()
        o = _options(slab.device.index or 0, reduce_order, tail_fma, True, row_base)
        bm = _bitmap(tombstones)
        check(_ffi.lib().fsgpu_index_create_f16(slab.data_ptr(), slab.shape[0], slab.shape[1], ptr(bm),
                                                C.byref(o), C.byref(h)))
        ix = cls(h.value, doc_ids, keepalive=slab)
        ix._tomb = None if tombstones is None else np.asarray(tombstones, dtype=bool).copy()
        return ix

    @classmethod
    def open(cls, path: str, *, device: int = 0, reduce_order="halves_pairwise", row_start: int = 0,
             n_rows: int = 0) -> "GpuVectorIndex":
        """VectorIndex::open (lib.rs:819) for an FSVI v1 / f16 file; doc ids come from the file."""
        from . import fsvi

        h = C.c_void_p()
        o = _options(device, reduce_order, True, False, row_start)
        o.flags = 1  # FSGPU_OPEN_HOST_REPLAYS_WAL: this host replays the WAL sidecar itself (below); without the flag a pending sidecar is an error
        check(_ffi.lib().fsgpu_index_open_fsvi(path.encode(), row_start, n_rows, C.byref(o), C.byref(h)))
        ix = cls(h.value, None, dedup_doc_ids=True)
        ix._doc_ids_from_handle = True
        ix._hashes_on_device = True
        n = ix.record_count()
        if n:  # record flag bit 0 as stored in the file (lib.rs:172)
            bm = np.zeros((n + 7) // 8, dtype=np.uint8)
            check(ix._L.fsgpu_index_read_tombstones(ix._h, ptr(bm)))
            flags = np.unpackbits(bm, bitorder="little")[:n].astype(bool)
            ix._tomb = flags if flags.any() else None
        # pending appends: the sidecar's rows become resident WAL rows again (VectorIndex::open,
        # lib.rs:1833-1878: last entry of a doc id wins, a stale sidecar is ignored).  append_batch already
        # tombstoned the main rows they supersede in the record table, and resolve_hits shadows the rest.
        # A row-range shard only takes the sidecar when it holds the END of the file (WAL rows are numbered
        # after record_count).
        header = fsvi.read_fsvi_header(path)
        if row_start + n == header["record_count"]:
            try:
                pending = fsvi.replay_wal_for(path, header)
            except SearchError:
                ix.close()
                raise
            if pending:
                ix._wal = [(d, np.ascontiguousarray(v, dtype=np.float32)) for d, v in pending]
                ix._upload_wal()
        return ix

    def close(self) -> None:
        if self._h:
            self._L.fsgpu_index_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    # ── accessors ───────────────────────────────────────────────────────────────────────────
    def record_count(self) -> int:
        return int(self._L.fsgpu_index_rows(self._h))

    def dimension(self) -> int:
        return int(self._L.fsgpu_index_dim(self._h))

    def row_base(self) -> int:
        return int(self._L.fsgpu_index_row_base(self._h))

    def device(self) -> int:
        return int(self._L.fsgpu_index_device(self._h))

    @property
    def handle(self) -> C.c_void_p:
        return self._h

    def doc_id_at(self, row: int) -> Optional[str]:
        """VectorIndex::doc_id_at (lib.rs:3801-3824) for a GLOBAL row; rows at or after
        `record_count` are the resident WAL rows (resolve_wal_hit, search.rs:1560-1598)."""
        w = row - self._wal_base()
        if 0 <= w < len(self._wal):
            return self._wal[w][0]
        if self._doc_ids is not None:
            return self._doc_ids[row - self.row_base()]
        p, n = C.c_void_p(), C.c_uint32()
        rc = self._L.fsgpu_index_doc_id(self._h, row, C.byref(p), C.byref(n))
        if rc != 0:
            return None
        return C.string_at(p, n.value).decode("utf-8")

    def read_rows_f16(self, row_start: int, n: int) -> np.ndarray:
        """Raw f16 bit patterns of local rows [row_start, row_start+n) (VectorIndex::vector_at_f16)."""
        out = np.zeros((n, self.dimension()), dtype=np.uint16)
        check(self._L.fsgpu_index_read_rows_f16(self._h, row_start, n, ptr(out)))
        return out

    def profile_enable(self, on: bool = True) -> None:
        check(self._L.fsgpu_index_profile_enable(self._h, 1 if on else 0))

    def profile_read(self, reset: bool = True) -> dict:
        p = _ffi.Profile()
        check(self._L.fsgpu_index_profile_read(self._h, C.byref(p), 1 if reset else 0))
        return dict(scan_launches=int(p.scan_launches), merge_launches=int(p.merge_launches),
                    other_launches=int(p.other_launches), scan_bytes=int(p.scan_bytes), scan_ms=float(p.scan_ms),
                    mma_launches=int(p.mma_launches), mma_flops=float(p.mma_flops),
                    redo_queries=int(p.redo_queries), i8_launches=int(p.i8_launches),
                    pair_launches=int(p.pair_launches), quad_launches=int(p.quad_launches))

    def set_tombstones(self, flags) -> None:
        """Soft-delete flags (record flag bit 0, lib.rs:172; honoured by the scan, search.rs:1281)."""
        bm = _bitmap(flags)
        check(self._L.fsgpu_index_set_tombstones(self._h, ptr(bm)))
        self._tomb = None if flags is None else np.asarray(flags, dtype=bool).copy()

    def is_deleted(self, row: int) -> bool:
        """VectorIndex::is_deleted (lib.rs:2401-2406) for a GLOBAL main row."""
        r = row - self.row_base()
        return bool(self._tomb is not None and 0 <= r < self._tomb.size and self._tomb[r])

    # ── WAL: rows appended since the last compaction, searchable at once ────────────────────
    def _wal_base(self) -> int:
        return self.row_base() + self.record_count()

    def wal_record_count(self) -> int:
        return len(self._wal)

    def wal_records(self):
        """VectorIndex::wal_records (lib.rs:2259): (doc_id, embedding) in WAL order."""
        return [(d, v.copy()) for d, v in self._wal]

    def _main_rows_of(self, doc_id: str) -> List[int]:
        if self._rows_of is None:
            n, base = self.record_count(), self.row_base()
            table: dict = {}
            for r in range(n):
                table.setdefault(self.doc_id_at(base + r), []).append(r)
            self._rows_of = table
        return self._rows_of.get(doc_id, [])

    def _tombstone_rows(self, rows: Sequence[int]) -> int:
        if self._tomb is None:
            self._tomb = np.zeros(self.record_count(), dtype=bool)
        changed = 0
        for r in rows:
            if not self._tomb[r]:
                self._tomb[r] = True
                changed += 1
        if changed:
            bm = _bitmap(self._tomb)  # named: the buffer must outlive the call
            check(self._L.fsgpu_index_set_tombstones(self._h, ptr(bm)))
        return changed

    def _upload_wal(self) -> None:
        dim = self.dimension()
        emb = np.ascontiguousarray(np.stack([v for _, v in self._wal]) if self._wal
                                   else np.zeros((0, dim)), dtype=np.float32)
        check(self._L.fsgpu_index_set_wal(self._h, ptr(emb) if self._wal else None, len(self._wal),
                                          self._wal_base()))

    def append(self, doc_id: str, vector) -> None:
        """VectorIndex::append (lib.rs:2532)."""
        self.append_batch([(doc_id, vector)])

    def append_batch(self, entries) -> None:
        """VectorIndex::append_batch (lib.rs:2546-2720) without the durability half: validate every
        entry first, keep the LAST entry of a doc id within the batch, supersede resident WAL rows
        of the same doc ids, admit the new rows (searchable at once, scored as f32), tombstone the
        main rows they replace."""
        entries = list(entries)
        if not entries:
            return
        dim = self.dimension()
        checked = []
        for doc_id, vector in entries:
            v = np.ascontiguousarray(vector, dtype=np.float32).reshape(-1)
            if v.size != dim:
                raise SearchError("DimensionMismatch", f"expected {dim}, found {v.size}")
            if not np.isfinite(v).all():
                raise SearchError("InvalidConfig", "all embedding values must be finite")
            norm_sq = np.float32(0.0)
            with np.errstate(over="ignore"):
                for x in v:  # vector_signal_usable (lib.rs:6133-6142): sequential f32 sum of squares
                    norm_sq = np.float32(norm_sq + np.float32(x * x))
            if not (norm_sq > 0.0 and np.isfinite(norm_sq)):
                raise SearchError("InvalidConfig", "embedding norm must be non-zero and finite")
            if len(doc_id.encode("utf-8")) > 0xFFFF:
                raise SearchError("InvalidConfig", "doc_id byte length must fit in u16")
            checked.append((doc_id, v))
        seen, fresh = set(), []
        for doc_id, v in reversed(checked):
            if doc_id not in seen:
                seen.add(doc_id)
                fresh.append((doc_id, v))
        fresh.reverse()
        self._wal = [e for e in self._wal if e[0] not in seen] + fresh
        self._upload_wal()
        self._tombstone_rows([r for doc_id, _ in fresh for r in self._main_rows_of(doc_id)])

    def soft_delete(self, doc_id: str) -> bool:
        """VectorIndex::soft_delete (lib.rs:2303)."""
        return self.soft_delete_batch([doc_id]) > 0

    def soft_delete_batch(self, doc_ids: Sequence[str]) -> int:
        """VectorIndex::soft_delete_batch (lib.rs:2314-2396): tombstone the live main rows of each
        doc id and drop its resident WAL rows; returns how many records went live -> deleted."""
        ids = set(doc_ids)
        deleted = self._tombstone_rows([r for d in doc_ids for r in self._main_rows_of(d)])
        kept = [e for e in self._wal if e[0] not in ids]
        if len(kept) < len(self._wal):
            deleted += len(self._wal) - len(kept)
            self._wal = kept
            self._upload_wal()
        return deleted

    def _ensure_doc_hashes(self) -> None:
        if self._hashes_on_device or self.record_count() == 0:
            return
        if self._doc_ids is None:
            raise SearchError("InvalidConfig", "a hash filter needs doc ids")
        hashes = np.array([fnv1a_hash(d.encode("utf-8")) for d in self._doc_ids], dtype=np.uint64)
        check(self._L.fsgpu_index_set_doc_hashes(self._h, ptr(hashes)))
        self._hashes_on_device = True

    # ── exact search ────────────────────────────────────────────────────────────────────────
    def _allow_bitmap(self, filter) -> Optional[np.ndarray]:
        """`filter` is the reference's `SearchFilter` seen from the host: a callable
        `doc_id -> bool` (SearchFilter::matches, crates/frankensearch-core/src/filter.rs:19) or a
        boolean array over local rows.  Returns the packed allow-bitmap the C ABI takes."""
        if filter is None:
            return None
        n = self.record_count()
        n_wal = len(self._wal)
        if callable(filter):  # WAL rows are filtered by doc id too (search.rs:1457-1465)
            base = self.row_base()
            mask = np.fromiter((bool(filter(self.doc_id_at(base + r))) for r in range(n + n_wal)), dtype=bool,
                               count=n + n_wal)
        else:
            mask = np.asarray(filter, dtype=bool).reshape(-1)
            if mask.size == n and n_wal:
                mask = np.concatenate([mask, np.ones(n_wal, dtype=bool)])
            if mask.size != n + n_wal:
                raise SearchError("InvalidConfig", f"filter mask has {mask.size} entries, index has {n} rows")
        return np.packbits(mask, bitorder="little")

    def search_top_k_batch(self, queries, limit: int, filter=None):
        """`batch` queries in one call.  Returns (rows u32 [B,k], scores f32 [B,k], counts u32 [B])."""
        q = np.ascontiguousarray(queries, dtype=np.float32)
        if q.ndim == 1:
            q = q[None, :]
        b, dim = q.shape
        k = int(limit)
        hits = np.zeros((b, max(k, 1)), dtype=np.dtype([("row", np.uint32), ("score", np.float32)]))
        counts = np.zeros(b, dtype=np.uint32)
        if filter is None:
            check(self._L.fsgpu_search_top_k(self._h, ptr(q), b, k, dim, ptr(hits), ptr(counts)))
        elif callable(getattr(filter, "candidate_hashes", None)) and filter.candidate_hashes() is not None:
            # BitsetFilter: decided by the record-table hash, on the device (search.rs:1329-1447),
            # selective sets through the gather arm (search.rs:1114-1161)
            self._ensure_doc_hashes()
            allowed = np.array(sorted(filter.candidate_hashes()), dtype=np.uint64)
            wal_allow = None
            if self._wal:
                wal_allow = np.packbits(np.array([bool(filter.matches(d)) for d, _ in self._wal], dtype=bool),
                                        bitorder="little")
            used = C.c_int(0)
            check(self._L.fsgpu_search_top_k_hashes(self._h, ptr(q), b, k, dim, ptr(allowed) if allowed.size else None,
                                                    allowed.size, ptr(wal_allow), ptr(hits), ptr(counts),
                                                    C.byref(used)))
            self.last_filter_arm = "gather" if used.value else "scan"
        else:
            self.last_filter_arm = "bitmap"
            bm = self._allow_bitmap(filter)
            check(self._L.fsgpu_search_top_k_filtered(self._h, ptr(q), b, k, dim, ptr(bm), ptr(hits), ptr(counts)))
        return hits["row"][:, :k].copy(), hits["score"][:, :k].copy(), counts

    def search_top_k(self, query, limit: int, filter=None) -> List[VectorHit]:
        """VectorIndex::search_top_k (search.rs:192-206) / InMemoryVectorIndex::search_top_k
        (in_memory.rs:2555): best-first hits, `limit == 0` or empty index -> []."""
        q = np.ascontiguousarray(query, dtype=np.float32).reshape(-1)
        rows, scores, counts = self.search_top_k_batch(q[None, :], limit, filter=filter)
        n = int(counts[0])
        hits = [VectorHit(int(rows[0, i]), float(scores[0, i]), self.doc_id_at(int(rows[0, i]))) for i in range(n)]
        if self._dedup or self._wal:
            # resolve_sorted_entries (search.rs:1503-1558): the first (= best) occurrence of a doc id
            # wins, and a main row whose doc id has a resident WAL row is shadowed by it
            wal_ids = {d for d, _ in self._wal}
            wal_base = self._wal_base()
            seen, out = set(), []
            for h in hits:
                if h.index < wal_base and (self.is_deleted(h.index) or h.doc_id in wal_ids):
                    continue
                if h.doc_id in seen:
                    continue
                seen.add(h.doc_id)
                out.append(h)
            hits = out
        return hits

    def _search_two_pass(self, query, k: int, candidate_multiplier: int, bits: int) -> List[VectorHit]:
        q = np.ascontiguousarray(query, dtype=np.float32).reshape(-1)
        if self._wal:  # the reference's gate (search.rs:578-586): resident WAL rows take the exact search
            return self.search_top_k(q, k)
        hits = np.zeros((max(int(k), 1), 2), dtype=np.uint32)
        count = np.zeros(1, dtype=np.uint32)
        check(self._L.fsgpu_search_top_k_two_pass(self._h, ptr(q), int(k), int(candidate_multiplier), bits, q.size, ptr(hits),
                                                 ptr(count)))
        n = int(count[0])
        scores = hits[:, 1].view(np.float32)
        return [VectorHit(int(hits[i, 0]), float(scores[i]), self.doc_id_at(int(hits[i, 0]))) for i in range(n)]

    def search_top_k_int8_two_pass(self, query, k: int, candidate_multiplier: int) -> List[VectorHit]:
        """VectorIndex::search_top_k_int8_two_pass (search.rs:514-650): integer pass 1 over the corpus-wide int8 codes keeps
        `k * candidate_multiplier` rows, exact f16 pass 2 keeps k — the reference's result for every multiplier."""
        return self._search_two_pass(query, k, candidate_multiplier, 8)

    def search_top_k_4bit_two_pass(self, query, k: int, candidate_multiplier: int) -> List[VectorHit]:
        """VectorIndex::search_top_k_4bit_two_pass (search.rs:876-946): the same over signed 4-bit nibbles."""
        return self._search_two_pass(query, k, candidate_multiplier, 4)

    def two_pass_codes(self, bits: int) -> np.ndarray:
        """The code slab those searches scan: [rows, dim] int8 or [rows, ceil(dim / 2)] packed nibbles."""
        n, d = self.record_count(), self.dimension()
        out = np.zeros((n, d if bits == 8 else (d + 1) // 2), dtype=np.uint8)
        check(self._L.fsgpu_index_read_two_pass_codes(self._h, bits, ptr(out)))
        return out.view(np.int8) if bits == 8 else out

    def zero_signal_state(self) -> ZeroSignalState:
        """VectorIndex::zero_signal_state (lib.rs:2441-2459): one census pass on the device."""
        out = np.zeros(5, dtype=np.uint64)
        check(self._L.fsgpu_index_zero_signal_state(self._h, ptr(out)))
        return ZeroSignalState(*(int(x) for x in out))

    def search_top_k_classified(self, query, limit: int, filter=None) -> ClassifiedHits:
        """VectorIndex::search_top_k_classified (search.rs:206-260): like search_top_k, but a non-finite
        query is rejected (InvalidConfig on `query`) and an empty result always says why."""
        q = np.ascontiguousarray(query, dtype=np.float32).reshape(-1)
        if q.size != self.dimension():  # ensure_query_dimension comes first (search.rs:233)
            raise SearchError("DimensionMismatch", f"expected {self.dimension()}, found {q.size}")
        if limit == 0:
            return ClassifiedHits([], ZeroSignalReason.CALLER_REQUESTED_ZERO_K)
        if not np.isfinite(q).all():
            raise SearchError("InvalidConfig", "query: <contains non-finite values>: query vector must be finite")
        if not q.any():
            return ClassifiedHits([], ZeroSignalReason.ZERO_NORM_QUERY)
        hits = self.search_top_k(q, limit, filter=filter)
        if not hits:
            return ClassifiedHits(hits, self.zero_signal_state().empty_result_reason(filter is not None))
        return ClassifiedHits(hits, None)

    def search_top_k_device(self, d_queries, limit: int, *, want_hits: bool = True, stream=None):
        """Device-resident form: `d_queries` is a CUDA float32 tensor [B, dim].  Returns torch
        tensors (keys int64 [B,k], hits int32-view [B,k,2] or None, counts int32 [B])."""
        import torch

        if d_queries.dtype != torch.float32 or not d_queries.is_cuda or not d_queries.is_contiguous():
            raise SearchError("InvalidConfig", "d_queries must be a contiguous CUDA float32 tensor")
        if d_queries.dim() != 2 or d_queries.shape[1] != self.dimension():
            raise SearchError("DimensionMismatch", f"expected {self.dimension()}, found {d_queries.shape[-1]}")
        b, k = d_queries.shape[0], int(limit)
        dev = d_queries.device
        keys = torch.zeros((b, max(k, 1)), dtype=torch.int64, device=dev)
        hits = torch.zeros((b, max(k, 1), 2), dtype=torch.int32, device=dev) if want_hits else None
        counts = torch.zeros(b, dtype=torch.int32, device=dev)
        s = torch.cuda.current_stream(dev).cuda_stream if stream is None else stream
        check(self._L.fsgpu_search_top_k_device(self._h, d_queries.data_ptr(), b, k, keys.data_ptr(),
                                                hits.data_ptr() if want_hits else None, counts.data_ptr(), s))
        return keys, hits, counts

    # ── two-tier rescoring ──────────────────────────────────────────────────────────────────
    def scores_for_rows(self, query, rows):
        q = np.ascontiguousarray(query, dtype=np.float32).reshape(-1)
        r = np.ascontiguousarray(rows, dtype=np.uint32)
        out = np.zeros(r.size, dtype=np.float32)
        present = np.zeros(r.size, dtype=np.uint8)
        check(self._L.fsgpu_scores_for_rows(self._h, ptr(q), q.size, ptr(r), r.size, ptr(out), ptr(present)))
        return out, present.astype(bool)

    def quality_scores_for_hits(self, query, hits: Sequence[VectorHit], alignment=None) -> List[Optional[float]]:
        """TwoTierIndex::quality_scores_for_hits (two_tier.rs:1566-1631): quality-tier dot at the
        row aligned with each fast hit.  `alignment[fast_row]` maps to the quality row (None =
        identical row order, the `Aligned` case of two_tier.rs:404-409); a missing row gives None."""
        rows = []
        for h in hits:
            r = h.index if alignment is None else alignment.get(h.index, 0xFFFFFFFF) \
                if isinstance(alignment, dict) else int(alignment[h.index])
            rows.append(0xFFFFFFFF if r is None or r < 0 else r)
        scores, present = self.scores_for_rows(query, np.asarray(rows, dtype=np.uint32))
        return [float(s) if p else None for s, p in zip(scores, present)]
