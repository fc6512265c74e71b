"""A stand-in for libfsgpu.so's index entry points, backed by the CPU oracle, so that the HOST logic
of frankensearch_b200/index.py (WAL bookkeeping, tombstones, filters -> bitmaps, doc-id resolve)
can be exercised on a box without a GPU.  TEST INFRASTRUCTURE ONLY: the product never sees this;
it replaces `GpuVectorIndex._L` on an object built around a dummy handle."""
import ctypes as C

import numpy as np

from oracle import fs_oracle as fo


def _bytes_at(addr, n):
    return np.frombuffer((C.c_uint8 * n).from_address(addr), dtype=np.uint8).copy() if n else np.zeros(0, np.uint8)


class FakeIndexLib:
    """One index: `slab_bits` [n, dim] uint16.  Implements exactly the calls index.py makes."""

    def __init__(self, slab_bits, row_base=0, tail_fma=False):
        self.slab = np.ascontiguousarray(slab_bits, dtype=np.uint16)
        self.n, self.dim = self.slab.shape
        self.row_base = row_base
        self.tail_fma = tail_fma
        self.tomb = np.zeros(self.n, dtype=bool)
        self.wal = np.zeros((0, self.dim), dtype=np.float32)
        self.wal_base = row_base + self.n
        self.calls = []

    # accessors ------------------------------------------------------------------------------
    def fsgpu_index_rows(self, h):
        return self.n

    def fsgpu_index_dim(self, h):
        return self.dim

    def fsgpu_index_row_base(self, h):
        return self.row_base

    def fsgpu_index_destroy(self, h):
        self.calls.append("destroy")

    # mutable state --------------------------------------------------------------------------
    def fsgpu_index_set_tombstones(self, h, addr):
        self.calls.append("set_tombstones")
        if addr is None:
            self.tomb = np.zeros(self.n, dtype=bool)
        else:
            bm = _bytes_at(addr, (self.n + 7) // 8)
            self.tomb = np.unpackbits(bm, bitorder="little")[: self.n].astype(bool)
        return 0

    def fsgpu_index_set_wal(self, h, addr, n_wal, base):
        self.calls.append(("set_wal", int(n_wal), int(base)))
        if n_wal:
            raw = _bytes_at(addr, int(n_wal) * self.dim * 4)
            self.wal = raw.view(np.float32).reshape(int(n_wal), self.dim).copy()
        else:
            self.wal = np.zeros((0, self.dim), dtype=np.float32)
        self.wal_base = int(base)
        return 0

    # search ---------------------------------------------------------------------------------
    def _search(self, q_addr, batch, k, dim, allow_addr, hits_addr, counts_addr):
        assert dim == self.dim
        q = _bytes_at(q_addr, batch * dim * 4).view(np.float32).reshape(batch, dim)
        n_wal = self.wal.shape[0]
        excl = self.tomb.copy()
        wal_allow = None
        if allow_addr is not None:
            bm = _bytes_at(allow_addr, (self.n + n_wal + 7) // 8)
            allow = np.unpackbits(bm, bitorder="little")[: self.n + n_wal].astype(bool)
            excl |= ~allow[: self.n]
            wal_allow = fo.pack_bitmap(allow[self.n:]) if n_wal else None
        hits = np.frombuffer((C.c_uint8 * (batch * max(k, 1) * 8)).from_address(hits_addr), dtype=np.dtype(
            [("row", np.uint32), ("score", np.float32)])).reshape(batch, max(k, 1))
        counts = np.frombuffer((C.c_uint8 * (batch * 4)).from_address(counts_addr), dtype=np.uint32)
        for b in range(batch):
            rows, scores = fo.search_top_k_wal(self.slab, self.wal, q[b], k, fo.pack_bitmap(excl) if self.n else None,
                                               wal_allow, 1, 0, self.tail_fma)
            # local rows -> global rows; WAL rows -> wal_base + w
            g = np.where(rows >= self.n, rows - self.n + self.wal_base, rows + self.row_base)
            counts[b] = len(rows)
            hits["row"][b, : len(rows)] = g.astype(np.uint32)
            hits["score"][b, : len(rows)] = scores
        return 0

    def fsgpu_search_top_k(self, h, q_addr, batch, k, dim, hits_addr, counts_addr):
        self.calls.append(("search", int(batch), int(k)))
        return self._search(q_addr, batch, k, dim, None, hits_addr, counts_addr)

    def fsgpu_search_top_k_filtered(self, h, q_addr, batch, k, dim, allow_addr, hits_addr, counts_addr):
        self.calls.append(("search_filtered", int(batch), int(k)))
        return self._search(q_addr, batch, k, dim, allow_addr, hits_addr, counts_addr)


def make_cpu_index(doc_ids, vectors, dim=None):
    """A GpuVectorIndex whose library calls go to FakeIndexLib (no GPU, no libfsgpu compute)."""
    from frankensearch_b200.index import GpuVectorIndex

    v = np.asarray(vectors, dtype=np.float32)
    dim = int(dim if dim is not None else v.shape[1])
    slab = fo.encode_f16(v.reshape(len(doc_ids), dim))
    ix = GpuVectorIndex.__new__(GpuVectorIndex)
    ix._h = C.c_void_p(1)
    ix._doc_ids = list(doc_ids)
    ix._dedup = False
    ix._keepalive = None
    ix._L = FakeIndexLib(slab)
    ix._wal = []
    ix._tomb = None
    ix._rows_of = None
    ix._hashes_on_device = False
    ix.last_filter_arm = None
    return ix
