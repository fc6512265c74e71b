"""Oracle-backed stand-ins for the libfsgpu.so entry points that frankensearch_b200/pipeline.py calls,
operating on HOST memory addresses (torch CPU tensors), so the sharded two-tier plumbing — packed
buffer layout, strides, the single all-gather, merge + payload pickup, blend and RRF wiring — can run
over gloo on a box without a GPU.  TEST INFRASTRUCTURE ONLY; the product never sees this."""
import ctypes as C

import numpy as np

from oracle import fs_oracle as fo
from oracle import np_oracle as no

HIT = np.dtype([("row", np.uint32), ("score", np.float32)])
FUSED = np.dtype([("rrf_score", np.float64), ("semantic_rank", np.int32), ("lexical_rank", np.int32),
                  ("semantic_row", np.uint32), ("semantic_score", np.float32), ("lexical_score", np.float32),
                  ("in_both_sources", np.uint32)])


def view(addr, dtype, count):
    dtype = np.dtype(dtype)
    buf = (C.c_uint8 * (dtype.itemsize * count)).from_address(int(addr))
    return np.frombuffer(buf, dtype=dtype, count=count)


def doc(i):
    i = int(i)
    return f"doc-{i:08}" if i < (1 << 32) else f"lexonly-{i - (1 << 32):08}"


class FakeShard:
    """Quacks like GpuVectorIndex for DeviceTwoTierSearcher: a row range of a host slab."""

    def __init__(self, handle, slab_bits, row_base):
        self.handle, self.slab, self.base = handle, np.ascontiguousarray(slab_bits, dtype=np.uint16), row_base

    def dimension(self):
        return self.slab.shape[1]

    def row_base(self):
        return self.base

    def record_count(self):
        return self.slab.shape[0]


class FakePipelineLib:
    def __init__(self, shards):
        self.shards = {s.handle: s for s in shards}

    def fsgpu_search_top_k_device(self, h, q_addr, b, k, keys_addr, hits_addr, counts_addr, stream):
        s = self.shards[h]
        q = view(q_addr, np.float32, b * s.dimension()).reshape(b, -1)
        keys, hits, counts = view(keys_addr, np.uint64, b * k).reshape(b, k), view(hits_addr, HIT, b * k).reshape(b, k), view(counts_addr, np.uint32, b)
        keys[:] = 0
        hits["row"][:] = 0xFFFFFFFF
        hits["score"][:] = 0
        for i in range(b):
            rows, scores = fo.search_top_k(s.slab, q[i], k, threads=1)
            g = rows.astype(np.uint64) + np.uint64(s.base)
            n = len(rows)
            keys[i, :n] = ~no.order_keys(scores, g)
            hits["row"][i, :n], hits["score"][i, :n] = g.astype(np.uint32), scores
            counts[i] = n
        return 0

    def fsgpu_scores_for_hits_device(self, h, q_addr, b, hits_addr, n, out_addr, present_addr, stream):
        s = self.shards[h]
        q = view(q_addr, np.float32, b * s.dimension()).reshape(b, -1)
        hits = view(hits_addr, HIT, b * n).reshape(b, n)
        out, present = view(out_addr, np.float32, b * n).reshape(b, n), view(present_addr, np.uint8, b * n).reshape(b, n)
        for i in range(b):
            rows = hits["row"][i].astype(np.int64)
            ok = (rows != 0xFFFFFFFF) & (rows >= s.base) & (rows < s.base + s.record_count())
            sc, _ = fo.scores_for_rows(s.slab, q[i], np.where(ok, rows - s.base, 0).astype(np.uint64))
            out[i] = np.where(ok, sc, 0)
            present[i] = ok
        return 0

    def fsgpu_merge_top_k_hits_device(self, dev, keys_addr, hits_addr, b, g, k_in, list_stride, query_stride, k_out,
                                      out_keys_addr, out_hits_addr, out_counts_addr, stream):
        span = (g - 1) * list_stride + (b - 1) * query_stride + k_in
        keys, hits = view(keys_addr, np.uint64, span), view(hits_addr, HIT, span)
        ok_, oh, oc = view(out_keys_addr, np.uint64, b * k_out).reshape(b, k_out), view(out_hits_addr, HIT, b * k_out).reshape(b, k_out), view(out_counts_addr, np.uint32, b)
        for q in range(b):
            idx = np.concatenate([np.arange(k_in) + s * list_stride + q * query_stride for s in range(g)])
            kk = keys[idx]
            order = np.argsort(kk, kind="stable")[::-1]
            order = order[kk[order] != 0][:k_out]
            n = len(order)
            ok_[q] = 0
            ok_[q, :n] = kk[order]
            oh["row"][q], oh["score"][q] = 0xFFFFFFFF, 0
            oh["row"][q, :n], oh["score"][q, :n] = hits["row"][idx[order]], hits["score"][idx[order]]
            oc[q] = n
        return 0

    def fsgpu_merge_payload_device(self, dev, keys_addr, pl_addr, b, g, k_in, list_stride, query_stride, pl_list_stride,
                                   pl_query_stride, merged_addr, k_out, out_addr, present_addr, stream):
        keys = view(keys_addr, np.uint64, (g - 1) * list_stride + (b - 1) * query_stride + k_in)
        pl = view(pl_addr, np.float32, (g - 1) * pl_list_stride + (b - 1) * pl_query_stride + k_in)
        merged = view(merged_addr, np.uint64, b * k_out).reshape(b, k_out)
        out, present = view(out_addr, np.float32, b * k_out).reshape(b, k_out), view(present_addr, np.uint8, b * k_out).reshape(b, k_out)
        for q in range(b):
            table = {}
            for s in range(g):
                for i in range(k_in):
                    kk = int(keys[s * list_stride + q * query_stride + i])
                    if kk:
                        table[kk] = pl[s * pl_list_stride + q * pl_query_stride + i]
            for i in range(k_out):
                kk = int(merged[q, i])
                present[q, i] = 1 if kk in table else 0
                out[q, i] = table.get(kk, 0.0)
        return 0

    def fsgpu_rrf_fuse_device(self, dev, cfg, b, lex_ids_addr, lex_scores_addr, lex_tie, lex_counts_addr, n_lex, sem_addr,
                              sem_tie, sem_counts_addr, n_sem, limit, offset, out_addr, out_counts_addr, stream):
        cfg = cfg._obj
        lex_ids, lex_scores = view(lex_ids_addr, np.uint64, b * n_lex).reshape(b, n_lex), view(lex_scores_addr, np.float32, b * n_lex).reshape(b, n_lex)
        sem = view(sem_addr, HIT, b * n_sem).reshape(b, n_sem)
        sem_counts = view(sem_counts_addr, np.uint32, b)
        lex_counts = view(lex_counts_addr, np.uint32, b) if lex_counts_addr else np.full(b, n_lex, np.uint32)
        out, oc = view(out_addr, FUSED, b * limit).reshape(b, limit), view(out_counts_addr, np.uint32, b)
        for q in range(b):
            lex = [(doc(i), float(s)) for i, s in zip(lex_ids[q, :lex_counts[q]], lex_scores[q, :lex_counts[q]])]
            se = [(doc(r), int(r), np.float32(s)) for r, s in zip(sem["row"][q, :sem_counts[q]], sem["score"][q, :sem_counts[q]])]
            fused = fo.rrf_fuse(lex, se, limit, offset, cfg.k, cfg.lexical_weight, cfg.semantic_weight, cfg.tiebreak)
            oc[q] = len(fused)
            for i, f in enumerate(fused):
                out[q, i] = (f.rrf_score, -1 if f.semantic_rank is None else f.semantic_rank,
                             -1 if f.lexical_rank is None else f.lexical_rank,
                             0xFFFFFFFF if f.semantic_index is None else f.semantic_index,
                             0.0 if f.semantic_score is None else f.semantic_score,
                             0.0 if f.lexical_score is None else f.lexical_score, int(f.in_both_sources))
        return 0

    def fsgpu_blend_two_tier_device(self, dev, alpha, b, fast_addr, fast_tie, fast_counts_addr, n_fast, q_hits, q_scores_addr,
                                    q_present_addr, q_tie, q_counts, n_q, out_addr, out_counts_addr, stream):
        assert not q_hits, "the plumbing test uses the aligned form"
        fast = view(fast_addr, HIT, b * n_fast).reshape(b, n_fast)
        fc = view(fast_counts_addr, np.uint32, b)
        qs, qp = view(q_scores_addr, np.float32, b * n_fast).reshape(b, n_fast), view(q_present_addr, np.uint8, b * n_fast).reshape(b, n_fast)
        out, oc = view(out_addr, HIT, b * n_fast).reshape(b, n_fast), view(out_counts_addr, np.uint32, b)
        for q in range(b):
            n = int(fc[q])
            f = [(doc(r), int(r), np.float32(s)) for r, s in zip(fast["row"][q, :n], fast["score"][q, :n])]
            blended = fo.blend_two_tier_aligned(f, [float(s) if p else None for s, p in zip(qs[q, :n], qp[q, :n])], alpha)
            out["row"][q], out["score"][q] = 0xFFFFFFFF, 0
            for i, (_, r, s) in enumerate(blended):
                out[q, i] = (r, s)
            oc[q] = len(blended)
        return 0
