"""Pipeline-level parity of `GpuSyncTwoTierSearcher` (sync_searcher.rs:616-1009) against the same
flow composed from the CPU oracle's stages: BASELINE configs[1]/[3] shape (fast tier -> RRF with a
precomputed BM25 list -> quality refinement -> blend -> re-fusion), small enough for the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fs(cuda_ok):
    assert cuda_ok, "no usable CUDA device: the product has no CPU fallback"
    import frankensearch_b200 as fs

    return fs


def oracle_flow(fo, fast_slab, qual_slab, ids, fq, qq, k, lexical, attested, cfg):
    fetch = max(k * cfg.candidate_multiplier, k)
    rows, scores = fo.search_top_k(fast_slab, fq, fetch)
    fast = [(ids[int(r)], int(r), float(s)) for r, s in zip(rows, scores)]
    if lexical is None:
        initial = [(d, np.float32(s)) for d, _, s in fast[:k]]
    else:
        fused = fo.rrf_fuse(lexical, fast, k, 0, cfg.rrf_k, 1.0, 1.0, 0)
        initial = [(f.doc_id, np.float32(f.rrf_score)) for f in fused]
    if attested:
        qrows, qscores = fo.search_top_k(qual_slab, qq, fetch)
        quality = [(ids[int(r)], int(r), float(s)) for r, s in zip(qrows, qscores)]
        blended = fo.blend_two_tier(fast, quality, cfg.quality_weight)
        qmap = {d: s for d, _, s in quality}
    else:
        qs, _ = fo.scores_for_rows(qual_slab, qq, rows)
        blended = fo.blend_two_tier_aligned(fast, [float(x) for x in qs], cfg.quality_weight)
        qmap = {d: float(s) for (d, _, _), s in zip(fast, qs)}
    if lexical is None:
        refined = [(d, np.float32(s)) for d, _, s in blended[:k]]
    else:
        fused = fo.rrf_fuse(lexical, [(d, i, float(s)) for d, i, s in blended], k, 0, cfg.rrf_k, 1.0, 1.0, 0)
        refined = [(f.doc_id, np.float32(f.rrf_score)) for f in fused]
    fmap = {d: s for d, _, s in fast}
    return initial, refined, fmap, qmap


@pytest.mark.parametrize("attested", [False, True])
@pytest.mark.parametrize("with_lexical", [False, True])
def test_sync_two_tier_searcher_matches_oracle_flow(fs, fo, attested, with_lexical):
    n, k = 30000, 15
    fast_slab, _ = fo.synth_rows(1, 1, 0, n, 256)
    qual_slab, _ = fo.synth_rows(1, 7, 0, n, 384)
    ids = [f"doc-{i:06}" for i in range(n)]
    fast_ix = fs.GpuVectorIndex.from_f16_bits(ids, fast_slab)
    qual_ix = fs.GpuVectorIndex.from_f16_bits(ids, qual_slab)
    cfg = fs.TwoTierConfig()
    rng = np.random.default_rng(11)
    for qi in range(3):
        fq, qq = fo.clustered_query(qi, 256), fo.clustered_query(qi, 384)
        lexical = None
        if with_lexical:
            sem_rows, _ = fo.search_top_k(fast_slab, fq, 3 * k)
            docs = [ids[int(r)] for r in rng.permutation(sem_rows)[: (3 * k) // 2]] + \
                   [ids[int(i)] for i in rng.integers(0, n, (3 * k) // 2)]
            lexical = [(d, float(len(docs) - i)) for i, d in enumerate(docs)]
        searcher = fs.GpuSyncTwoTierSearcher(
            fast_ix, qual_ix, cfg, quality_attested=attested,
            lexical=(lambda q, fetch, lx=lexical: [fs.ScoredResult(d, s) for d, s in lx]) if lexical else None)
        out = searcher.search_collect(fq, qq, k)
        want_initial, want_refined, fmap, qmap = oracle_flow(fo, fast_slab, qual_slab, ids, fq, qq, k, lexical,
                                                              attested, cfg)
        assert out.refined
        rc_ = out.metrics["rank_changes"]  # searcher.rs:2405-2412: every fast-pool doc is counted once
        assert rc_.demoted + rc_.stable + rc_.promoted >= 3 * k
        tau = out.metrics["kendall_tau"]
        assert tau is not None and -1.0 <= tau <= 1.0
        assert [(r.doc_id, np.float32(r.score).view(np.uint32)) for r in out.initial_results] == \
               [(d, np.float32(s).view(np.uint32)) for d, s in want_initial]
        assert [(r.doc_id, np.float32(r.score).view(np.uint32)) for r in out.final_results] == \
               [(d, np.float32(s).view(np.uint32)) for d, s in want_refined]
        for r in out.final_results:  # evidence fields are the raw per-tier scores
            assert (r.fast_score is None) == (r.doc_id not in fmap)
            if r.fast_score is not None:
                assert np.float32(r.fast_score) == np.float32(fmap[r.doc_id])
            if r.doc_id in qmap:
                assert np.float32(r.quality_score) == np.float32(qmap[r.doc_id])
    fast_ix.close()
    qual_ix.close()


def test_sync_searcher_short_circuits(fs, fo):
    slab, _ = fo.synth_rows(1, 1, 0, 2000, 128)
    ix = fs.GpuVectorIndex.from_f16_bits([f"d{i}" for i in range(2000)], slab)
    s = fs.GpuSyncTwoTierSearcher(ix)
    q = fo.clustered_query(0, 128)
    assert s.search_collect(q, None, 0).final_results == []                       # k == 0
    z = s.search_collect(np.zeros(128, np.float32), None, 5)
    assert z.final_results == [] and z.metrics["zero_signal"] == "ZeroNormQuery"   # all-zero query
    out = s.search_collect(q, None, 5)
    assert len(out.final_results) == 5 and not out.refined
    assert out.metrics["skip_reason"] == "quality_index_unavailable"
    fast_only = fs.GpuSyncTwoTierSearcher(ix, ix, fs.TwoTierConfig(fast_only=True))
    assert fast_only.search_collect(q, q, 5).metrics["skip_reason"] == "fast_only"
    ix.close()
