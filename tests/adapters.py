"""Two implementations behind one test-facing interface: the CPU oracle and the GPU product."""
import numpy as np


class OracleImpl:
    name = "oracle"

    def __init__(self):
        from oracle import fs_oracle as fo

        self.fo = fo

    # scan ------------------------------------------------------------------------------------
    def search(self, rows_f32, query, k, tombstones=None, reduce_order=0, tail_fma=True, threads=None):
        slab = self.fo.encode_f16(np.asarray(rows_f32, dtype=np.float32))
        bm = None if tombstones is None else self.fo.pack_bitmap(tombstones)
        r, s = self.fo.search_top_k(slab, np.asarray(query, dtype=np.float32), k, bm, threads, reduce_order, tail_fma)
        return [int(x) for x in r], s

    def search_bits(self, slab_bits, query, k, tombstones=None, reduce_order=0, tail_fma=True, threads=None):
        bm = None if tombstones is None else self.fo.pack_bitmap(tombstones)
        r, s = self.fo.search_top_k(slab_bits, np.asarray(query, dtype=np.float32), k, bm, threads, reduce_order,
                                    tail_fma)
        return [int(x) for x in r], s

    # fusion ----------------------------------------------------------------------------------
    def rrf(self, lexical, semantic, limit, offset=0, k=60.0, w_lex=1.0, w_sem=1.0, tiebreak="LexicalThenId"):
        out = self.fo.rrf_fuse(lexical, semantic, limit, offset, k, w_lex, w_sem, 1 if tiebreak == "Hash" else 0)
        return [(h.doc_id, h.rrf_score, h.lexical_rank, h.semantic_rank, h.in_both_sources) for h in out]

    def blend(self, fast, quality, alpha):
        return self.fo.blend_two_tier(fast, quality, alpha)

    def blend_aligned(self, fast, scores, alpha):
        return self.fo.blend_two_tier_aligned(fast, scores, alpha)

    def potion(self, table, ids):
        return self.fo.potion_embed(table, ids)


class GpuImpl:
    name = "gpu"

    def __init__(self):
        import frankensearch_b200 as fs

        self.fs = fs

    def search(self, rows_f32, query, k, tombstones=None, reduce_order=0, tail_fma=True, threads=None):
        ix = self.fs.GpuVectorIndex.from_vectors(None, np.asarray(rows_f32, dtype=np.float32), reduce_order=reduce_order,
                                                 tail_fma=tail_fma, tombstones=tombstones)
        try:
            rows, scores, counts = ix.search_top_k_batch(np.asarray(query, dtype=np.float32), k)
        finally:
            ix.close()
        n = int(counts[0])
        return [int(x) for x in rows[0, :n]], scores[0, :n]

    def search_bits(self, slab_bits, query, k, tombstones=None, reduce_order=0, tail_fma=True, threads=None):
        ix = self.fs.GpuVectorIndex.from_f16_bits(None, slab_bits, reduce_order=reduce_order, tail_fma=tail_fma,
                                                  tombstones=tombstones)
        try:
            rows, scores, counts = ix.search_top_k_batch(np.asarray(query, dtype=np.float32), k)
        finally:
            ix.close()
        n = int(counts[0])
        return [int(x) for x in rows[0, :n]], scores[0, :n]

    def rrf(self, lexical, semantic, limit, offset=0, k=60.0, w_lex=1.0, w_sem=1.0, tiebreak="LexicalThenId"):
        fs = self.fs
        lex = [fs.ScoredResult(d, s) for d, s in lexical]
        sem = [fs.VectorHit(i, s, d) for d, i, s in semantic]
        out = fs.rrf_fuse(lex, sem, limit, offset, fs.RrfConfig(k, w_lex, w_sem, tiebreak))
        return [(h.doc_id, h.rrf_score, h.lexical_rank, h.semantic_rank, h.in_both_sources) for h in out]

    def blend(self, fast, quality, alpha):
        fs = self.fs
        out = fs.blend_two_tier([fs.VectorHit(i, s, d) for d, i, s in fast],
                                [fs.VectorHit(i, s, d) for d, i, s in quality], alpha)
        return [(h.doc_id, h.index, np.float32(h.score)) for h in out]

    def blend_aligned(self, fast, scores, alpha):
        fs = self.fs
        out = fs.blend_two_tier_aligned([fs.VectorHit(i, s, d) for d, i, s in fast], scores, alpha)
        return [(h.doc_id, h.index, np.float32(h.score)) for h in out]

    def potion(self, table, ids):
        enc = self.fs.Model2VecEmbedder(table)
        try:
            return enc.embed_token_ids(ids)
        finally:
            enc.close()
