"""Row-sharded exact search across the GPUs of one box.

One process per GPU (torch.distributed, NCCL over NVLink).  Rank r owns the contiguous global
rows [r*N/G, (r+1)*N/G); a search is: local fused scan+top-k on every rank -> ONE all-gather of
the per-rank `batch x k` packed order keys (+ raw scores) -> the same k-way merge kernel on every
rank.  Because the key order is a strict total order on (score_key, global_row)
(crates/frankensearch-index/src/search.rs:1669-1686) the merged result is identical to the
single-index result — this is merge_partial_heaps (search.rs:1704-1720) lifted across devices.

Two-dimensional layout (`query_groups` = Q > 1): the G ranks form R = G/Q row shards x Q query groups.
Rank g*R + r holds rows [r*N/R, (r+1)*N/R) and searches only the g-th block of ceil(B/Q) queries; the
all-gather and the merge are the same (one merge launch per query group over that group's R lists).
Every rank does the same share of the contraction as with G row shards, but the per-call work that
does not shrink with the shard (sample cascade, exact gate, refine: one CTA per QUERY) is halved with
the batch — at 8 GPUs and 1024 queries over 10 M rows that is the difference the bench records.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

from . import _ffi
from ._ffi import SearchError, check


def grid_position(world_size: int, rank: int, query_groups: int = 1) -> Tuple[int, int, int]:
    """(row_shards R, row shard r, query group g) of `rank` in the R x Q layout: rank = g*R + r."""
    q = int(query_groups)
    if q <= 0 or world_size <= 0 or world_size % q != 0:
        raise SearchError("InvalidConfig", f"{world_size} ranks do not split into {query_groups} query groups")
    if not (0 <= rank < world_size):
        raise SearchError("InvalidConfig", f"rank {rank} outside world of {world_size}")
    r_shards = world_size // q
    return r_shards, rank % r_shards, rank // r_shards


def query_block(batch: int, query_groups: int, group: int) -> Tuple[int, int]:
    """[lo, hi) of the queries group `group` searches: blocks of ceil(batch / Q), the last ones may be short or empty."""
    bq = (batch + query_groups - 1) // query_groups
    return min(batch, group * bq), min(batch, (group + 1) * bq)


def shard_bounds(n_rows: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous row range of `rank`: [rank*N/G, (rank+1)*N/G) (integer division as written)."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise SearchError("InvalidConfig", f"rank {rank} outside world of {world_size}")
    return (rank * n_rows) // world_size, ((rank + 1) * n_rows) // world_size


class ShardedGpuIndex:
    """`local_search(d_queries, k) -> (keys int64 [B,k], scores f32 [B,k])` and
    `merge(keys [G,B,k], scores [G,B,k], k) -> (keys [B,k], hits, counts)` default to the CUDA
    kernels; tests inject CPU stand-ins to exercise the plumbing over gloo."""

    def __init__(self, local_index, *, group=None, local_search: Optional[Callable] = None,
                 merge: Optional[Callable] = None, query_groups: int = 1):
        import torch.distributed as dist

        self._dist = dist
        self._ix = local_index
        self._group = group
        self._world = dist.get_world_size(group) if dist.is_initialized() else 1
        self._rank = dist.get_rank(group) if dist.is_initialized() else 0
        # R row shards x Q query groups; the local index must hold row shard `row_shard` of `row_shards`
        self.query_groups = int(query_groups)
        self.row_shards, self.row_shard, self.query_group = grid_position(self._world, self._rank, self.query_groups)
        self._fast = local_search is None and merge is None  # both stages are the CUDA kernels
        self._local_search = local_search or self._cuda_local_search
        self._merge = merge or self._cuda_merge
        self._packed_shape = None  # (batch, k, device) of the cached all-gather buffers

    # CUDA stages -----------------------------------------------------------------------------
    def _cuda_local_search(self, d_queries, k: int):
        import torch

        keys, hits, _counts = self._ix.search_top_k_device(d_queries, k, want_hits=True)
        scores = hits[..., 1].contiguous().view(torch.float32)
        return keys, scores

    def _cuda_merge(self, keys, scores, k: int):
        import torch

        g, b, k_in = keys.shape
        dev = keys.device
        out_keys = torch.zeros((b, k), dtype=torch.int64, device=dev)
        out_hits = torch.zeros((b, k, 2), dtype=torch.int32, device=dev)
        counts = torch.zeros(b, dtype=torch.int32, device=dev)
        s = torch.cuda.current_stream(dev).cuda_stream
        check(_ffi.lib().fsgpu_merge_top_k_device(dev.index or 0, keys.data_ptr(), scores.data_ptr(), b, g, k_in,
                                                  b * k_in, k_in, k, out_keys.data_ptr(), out_hits.data_ptr(),
                                                  counts.data_ptr(), s))
        return out_keys, out_hits, counts

    def _cuda_search_packed(self, d_queries, k: int):
        """CUDA path with ONE packed buffer per rank: [keys | hits] as int64 [2, B, k] -> one
        all-gather -> the merge kernel reads keys and the hits' raw scores in place (no copies).
        The local search writes straight into the packed buffer; the packed and gathered buffers are
        kept between calls (they are internal, and reuse is ordered by the stream)."""
        import torch

        if self._world == 1:
            return self._ix.search_top_k_device(d_queries, k, want_hits=True)
        if d_queries.dtype != torch.float32 or not d_queries.is_cuda or not d_queries.is_contiguous():
            raise SearchError("InvalidConfig", "d_queries must be a contiguous CUDA float32 tensor")
        if d_queries.dim() != 2 or d_queries.shape[1] != self._ix.dimension():
            raise SearchError("DimensionMismatch", f"expected {self._ix.dimension()}, found {d_queries.shape[-1]}")
        b, g, dev = d_queries.shape[0], self._world, d_queries.device
        nq, r_shards = self.query_groups, self.row_shards
        bq = (b + nq - 1) // nq  # queries per group (the all-gather needs equal blocks; short groups leave a tail unused)
        shape = (b, k, dev)
        if self._packed_shape != shape:
            self._packed = torch.empty((2, bq, k), dtype=torch.int64, device=dev)
            self._gathered = torch.empty((g, 2, bq, k), dtype=torch.int64, device=dev)
            self._local_counts = torch.empty(bq, dtype=torch.int32, device=dev)
            self._packed_shape = shape
        packed, flat = self._packed, self._gathered
        s = torch.cuda.current_stream(dev).cuda_stream
        L = _ffi.lib()
        lo, hi = query_block(b, nq, self.query_group)
        if hi > lo:
            check(L.fsgpu_search_top_k_device(self._ix.handle, d_queries.data_ptr() + lo * d_queries.shape[1] * 4, hi - lo, k,
                                              packed[0].data_ptr(), packed[1].data_ptr(), self._local_counts.data_ptr(), s))
        self._dist.all_gather_into_tensor(flat, packed, group=self._group)
        out_keys = torch.empty((b, k), dtype=torch.int64, device=dev)
        out_hits = torch.empty((b, k, 2), dtype=torch.int32, device=dev)
        out_counts = torch.empty(b, dtype=torch.int32, device=dev)
        rank_bytes = 2 * bq * k * 8
        for grp in range(nq):  # group grp's R lists are the blocks of ranks grp*R .. grp*R + R - 1
            lo, hi = query_block(b, nq, grp)
            if hi <= lo:
                continue
            base = flat.data_ptr() + grp * r_shards * rank_bytes
            check(L.fsgpu_merge_top_k_hits_device(dev.index or 0, base, base + bq * k * 8, hi - lo, r_shards,
                                                  k, 2 * bq * k, k, k, out_keys.data_ptr() + lo * k * 8,
                                                  out_hits.data_ptr() + lo * k * 8, out_counts.data_ptr() + lo * 4, s))
        return out_keys, out_hits, out_counts

    # the sharded search ------------------------------------------------------------------------
    def search_top_k_device(self, d_queries, k: int):
        """Every rank passes the same queries; every rank returns the same merged result."""
        import torch

        if self._fast and k > 0:
            return self._cuda_search_packed(d_queries, k)
        if self.query_groups > 1:
            return self._search_grid(d_queries, k)
        keys, scores = self._local_search(d_queries, k)
        if self._world == 1:
            return self._merge(keys.unsqueeze(0), scores.unsqueeze(0), k)
        g = self._world
        all_keys = torch.empty((g,) + tuple(keys.shape), dtype=keys.dtype, device=keys.device)
        all_scores = torch.empty((g,) + tuple(scores.shape), dtype=scores.dtype, device=scores.device)
        # one collective: keys and scores travel in a single packed buffer
        packed = torch.cat([keys.reshape(-1).view(torch.int32), scores.reshape(-1).view(torch.int32)])
        flat = torch.empty(g * packed.numel(), dtype=torch.int32, device=packed.device)
        self._dist.all_gather_into_tensor(flat, packed, group=self._group)
        gathered = flat.view(g, packed.numel())
        nk = keys.numel() * 2
        all_keys.copy_(gathered[:, :nk].contiguous().view(torch.int64).view(all_keys.shape))
        all_scores.copy_(gathered[:, nk:].contiguous().view(torch.float32).view(all_scores.shape))
        return self._merge(all_keys, all_scores, k)

    def _search_grid(self, d_queries, k: int):
        """R x Q layout through the injectable stages (tests over gloo; k = 0): this rank's query block ->
        all-gather of equal, zero-padded blocks -> one merge per query group, results concatenated."""
        import torch

        b, nq, r_shards, g = d_queries.shape[0], self.query_groups, self.row_shards, self._world
        bq = (b + nq - 1) // nq
        lo, hi = query_block(b, nq, self.query_group)
        keys = torch.zeros((bq, k), dtype=torch.int64, device=d_queries.device)
        scores = torch.zeros((bq, k), dtype=torch.float32, device=d_queries.device)
        if hi > lo:
            lk, ls = self._local_search(d_queries[lo:hi], k)
            keys[: hi - lo], scores[: hi - lo] = lk, ls
        packed = torch.cat([keys.reshape(-1).view(torch.int32), scores.reshape(-1).view(torch.int32)])
        flat = torch.empty(g * packed.numel(), dtype=torch.int32, device=packed.device)
        self._dist.all_gather_into_tensor(flat, packed, group=self._group)
        gathered = flat.view(g, packed.numel())
        nk = keys.numel() * 2
        all_keys = gathered[:, :nk].contiguous().view(torch.int64).view(g, bq, k)
        all_scores = gathered[:, nk:].contiguous().view(torch.float32).view(g, bq, k)
        outs = []
        for grp in range(nq):
            lo, hi = query_block(b, nq, grp)
            if hi > lo:
                sl = slice(grp * r_shards, (grp + 1) * r_shards)
                outs.append(self._merge(all_keys[sl, : hi - lo].contiguous(), all_scores[sl, : hi - lo].contiguous(), k))
        cat = lambda i: None if outs[0][i] is None else torch.cat([o[i] for o in outs])  # noqa: E731
        return cat(0), cat(1), cat(2)


class GpuShardedIndex:
    """Single-process form (no torch.distributed, no NCCL): one `fsgpu_sharded` handle over several GPUs
    of the box — what a Rust host binds (`fsgpu_sharded_*`, include/fsgpu.h).  Shard i holds the contiguous
    rows [i*N/G, (i+1)*N/G) on `devices[i]`; a search runs on every shard concurrently (one host thread
    each) and, with peer access, the shards' kernels store their top-k straight into the merge device's
    buffer over NVLink.  The result is byte-identical to one index over all rows."""

    def __init__(self, handle, keepalive=None):
        import ctypes as C

        self._h = C.c_void_p(handle)
        self._L = _ffi.lib()
        self._keepalive = keepalive

    @classmethod
    def from_f16_bits(cls, slab_bits, devices, *, tombstones=None, reduce_order=0, tail_fma=True):
        import ctypes as C

        import numpy as np

        s = np.ascontiguousarray(slab_bits, dtype=np.uint16)
        o = _ffi.IndexOptions()
        _ffi.lib().fsgpu_index_options_default(C.byref(o))
        o.reduce_order, o.tail_fma = int(reduce_order), 1 if tail_fma else 0
        dev = (C.c_int * len(devices))(*devices)
        bm = None if tombstones is None else np.packbits(np.asarray(tombstones, dtype=bool), bitorder="little")
        h = C.c_void_p()
        check(_ffi.lib().fsgpu_sharded_create_f16(_ffi.ptr(s), s.shape[0], s.shape[1], _ffi.ptr(bm), dev, len(devices),
                                                  C.byref(o), C.byref(h)))
        return cls(h.value)

    @classmethod
    def from_indexes(cls, indexes):
        """Adopt per-device GpuVectorIndex shards (contiguous row ranges, in order); they stay owned by the caller."""
        import ctypes as C

        arr = (C.c_void_p * len(indexes))(*[ix.handle for ix in indexes])
        h = C.c_void_p()
        check(_ffi.lib().fsgpu_sharded_from_shards(arr, len(indexes), 0, C.byref(h)))
        return cls(h.value, keepalive=list(indexes))

    def close(self):
        if self._h:
            self._L.fsgpu_sharded_destroy(self._h)
            self._h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def shard_count(self) -> int:
        return int(self._L.fsgpu_sharded_shard_count(self._h))

    def record_count(self) -> int:
        return int(self._L.fsgpu_sharded_rows(self._h))

    def direct_shards(self):
        return [bool(self._L.fsgpu_sharded_is_direct(self._h, i)) for i in range(self.shard_count())]

    def search_top_k_batch(self, queries, limit: int):
        import numpy as np

        q = np.ascontiguousarray(queries, dtype=np.float32)
        if q.ndim == 1:
            q = q[None, :]
        b, dim = q.shape
        k = int(limit)
        hits = np.zeros((b, max(k, 1)), dtype=np.dtype([("row", np.uint32), ("score", np.float32)]))
        counts = np.zeros(b, dtype=np.uint32)
        check(self._L.fsgpu_sharded_search_top_k(self._h, _ffi.ptr(q), b, k, dim, _ffi.ptr(hits), _ffi.ptr(counts)))
        return hits["row"][:, :k].copy(), hits["score"][:, :k].copy(), counts

