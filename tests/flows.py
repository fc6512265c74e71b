"""The reference's two-tier search flow restated on the CPU oracle (TEST INFRASTRUCTURE ONLY):
SyncTwoTierSearcher::search_internal, crates/frankensearch-fusion/src/sync_searcher.rs:616-1009,
composed from the oracle's scan, re-score, blend and RRF restatements.  The GPU pipeline
(frankensearch_b200/pipeline.py) is compared with this, query by query."""
import numpy as np

from oracle import fs_oracle as fo


def doc_id(row: int) -> str:
    return f"doc-{int(row):06}"  # fsvi_int8_two_pass.rs:272-281


def synthetic_lexical(fast_rows, n_docs, fetch, seed):
    """The synthetic BM25 list of frankensearch/benches/search_bench.rs:236-251 adapted to rows: about
    half of the entries are semantic candidates (every other one), the rest other documents (some of
    them without a vector row: ids >= 2^32); scores (n - i) as f32, rank order."""
    rng = np.random.default_rng(seed)
    pool = [int(r) for r in fast_rows[::2]]
    while len(pool) < fetch:
        r = int(rng.integers(0, n_docs + n_docs // 8))
        pool.append(r if r < n_docs else (1 << 32) + r)
    pool = pool[:fetch]
    order = rng.permutation(len(pool))
    ids = np.array([pool[i] for i in order], dtype=np.uint64)
    # duplicates keep their FIRST rank (rrf.rs:1059-1064): leave them in, the fusion handles it
    scores = np.array([np.float32(len(ids) - i) for i in range(len(ids))], dtype=np.float32)
    return ids, scores


def lexical_doc_id(i: int) -> str:
    return doc_id(i) if i < (1 << 32) else f"lexonly-{i - (1 << 32):08}"


def oracle_two_tier(fast_slab, quality_slab, fast_q, quality_q, k, lexical=None, multiplier=3, alpha=0.7,
                    rrf_k=60.0, w_lex=1.0, w_sem=1.0):
    """Returns dict(fast=(rows, scores), quality=scores|None, initial=[...], blended=[...], refined=[...]);
    fused entries are oracle FusedHit objects, plain ones (doc_id, row, score)."""
    fetch = max(k * max(multiplier, 1), k)
    rows, scores = fo.search_top_k(fast_slab, fast_q, fetch)
    fast = [(doc_id(r), int(r), np.float32(s)) for r, s in zip(rows, scores)]
    out = {"fast": (rows, scores), "fetch": fetch}
    lex = None
    if lexical is not None:
        lex = [(lexical_doc_id(int(i)), float(s)) for i, s in zip(*lexical)]
        out["initial"] = fo.rrf_fuse(lex, fast, k, 0, rrf_k, w_lex, w_sem)
    else:
        out["initial"] = fast[:k]
    if quality_slab is None or quality_q is None:
        return out
    q_scores, present = fo.scores_for_rows(quality_slab, quality_q, rows)
    out["quality"] = q_scores
    blended = fo.blend_two_tier_aligned(fast, [float(s) if p else None for s, p in zip(q_scores, present)], alpha)
    out["blended"] = blended
    if lex is not None:
        out["refined"] = fo.rrf_fuse(lex, blended, k, 0, rrf_k, w_lex, w_sem)
    else:
        out["refined"] = blended[:k]
    return out


def assert_fused_equal(got, want, what):
    """got: structured numpy row array of fsgpu_fused_hit; want: oracle FusedHit list."""
    assert len(got) == len(want), (what, len(got), len(want))
    for i, (g, w) in enumerate(zip(got, want)):
        assert np.float64(g["rrf_score"]).view(np.uint64) == np.float64(w.rrf_score).view(np.uint64), (what, i)
        assert (int(g["semantic_rank"]) if g["semantic_rank"] >= 0 else None) == w.semantic_rank, (what, i)
        assert (int(g["lexical_rank"]) if g["lexical_rank"] >= 0 else None) == w.lexical_rank, (what, i)
        assert bool(g["in_both_sources"]) == w.in_both_sources, (what, i)
        if w.semantic_rank is not None:
            assert int(g["semantic_row"]) == w.semantic_index, (what, i)
            assert np.float32(g["semantic_score"]).view(np.uint32) == np.float32(w.semantic_score).view(np.uint32), (what, i)
        if w.lexical_rank is not None:
            assert np.float32(g["lexical_score"]).view(np.uint32) == np.float32(w.lexical_score).view(np.uint32), (what, i)


def assert_hits_equal(got, want, what):
    """got: structured numpy array of fsgpu_hit; want: [(doc_id, row, score)]."""
    assert len(got) == len(want), (what, len(got), len(want))
    assert [int(r) for r in got["row"]] == [int(w[1]) for w in want], what
    assert np.array_equal(got["score"].view(np.uint32),
                          np.array([w[2] for w in want], dtype=np.float32).view(np.uint32)), what
