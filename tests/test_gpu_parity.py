"""GPU parity tests (run with `-m gpu` on a B200): the CUDA path, called through the C ABI,
against (1) the reference's known-answer tables, (2) the committed golden fixtures, (3) the CPU
oracle on the same seeded inputs, (4) size-independent properties at BASELINE.json's full size.

Bars: bit-exact rows/ranks and bit-exact f32 scores for the exact scan (north_star asks for
<= 1e-3 relative on scores; the kernel reproduces the reference accumulation tree so the test
demands 0 ULP); bit-exact f64 rrf_score; bit-exact blend scores.
"""
import json
import math
import os

import numpy as np
import pytest

import ref_cases as rc
from adapters import GpuImpl, OracleImpl
from checks import check_blend_aligned, check_blend_case, check_rrf_case, check_scan_case
from oracle import np_oracle as no

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def gpu(cuda_ok):
    assert cuda_ok, "no usable CUDA device: the product has no CPU fallback, -m gpu must run on a B200"
    return GpuImpl()


@pytest.fixture(scope="module")
def cpu():
    return OracleImpl()


def bits(x):
    return np.asarray(x, dtype=np.float32).view(np.uint32)


def assert_same_hits(got, want, ctx=""):
    assert list(got[0]) == [int(r) for r in want[0]], f"rows differ {ctx}"
    g, w = np.asarray(got[1], dtype=np.float32), np.asarray(want[1], dtype=np.float32)
    nan = np.isnan(w)
    assert np.array_equal(np.isnan(g), nan), f"NaN pattern differs {ctx}"
    assert np.array_equal(bits(g)[~nan], bits(w)[~nan]), f"score bits differ {ctx}"


# ── reference known-answer tables ───────────────────────────────────────────────────────────
@pytest.mark.parametrize("case", rc.SCAN_CASES, ids=[c["name"] for c in rc.SCAN_CASES])
def test_scan_known_answers(gpu, case):
    check_scan_case(gpu, case)


@pytest.mark.parametrize("case", rc.RRF_CASES, ids=[c["name"] for c in rc.RRF_CASES])
def test_rrf_known_answers(gpu, case):
    check_rrf_case(gpu, case)


@pytest.mark.parametrize("case", rc.BLEND_CASES, ids=[c["name"] for c in rc.BLEND_CASES])
def test_blend_known_answers(gpu, case):
    check_blend_case(gpu, case)


def test_aligned_blend_is_bit_identical_to_materialized(gpu):
    check_blend_aligned(gpu, rc.BLEND_ALIGNED)


def test_parallel_and_sequential_paths_match(gpu):
    c = rc.parallel_case()
    rows = np.array([v for _, v in c["rows"]], dtype=np.float32)
    got = gpu.search(rows, c["query"], c["k"])
    assert got[0] == list(range(63, 53, -1))


# ── golden fixtures ─────────────────────────────────────────────────────────────────────────
def test_golden_dot_vectors(gpu, fo):
    import frankensearch_b200 as fs

    with open(os.path.join(GOLDEN, "dot_f16_f32.json")) as f:
        g = json.load(f)
    for case in g["cases"]:
        row = np.array(case["row_bits"], dtype=np.uint16)[None, :]
        q = np.array(case["query_bits"], dtype=np.uint32).view(np.float32)
        for order, want in enumerate(case["score_bits_by_order"]):
            ix = fs.GpuVectorIndex.from_f16_bits(None, row, reduce_order=order, tail_fma=True)
            s, present = ix.scores_for_rows(q, [0, 1, 0xFFFFFFFF])
            rows, scores, counts = ix.search_top_k_batch(q, 1)
            ix.close()
            assert present.tolist() == [True, False, False]
            assert int(bits(s[0])) == want, (case["dim"], order)
            assert int(counts[0]) == 1 and int(bits(scores[0, 0])) == want


def test_golden_clustered_scan(gpu, fo):
    with open(os.path.join(GOLDEN, "scan_clustered_3000x384.json")) as f:
        g = json.load(f)
    slab, _ = fo.synth_rows(1, 1, 0, 3000, 384)
    for res in g["results"]:
        q = fo.clustered_query(res["query"], 384)
        rows, scores = gpu.search_bits(slab, q, 10)
        assert rows == res["rows"]
        assert [int(b) for b in bits(scores)] == res["score_bits"]


# ── oracle parity on seeded inputs ──────────────────────────────────────────────────────────
@pytest.mark.parametrize("dim", [4, 7, 8, 20, 31, 32, 100, 128, 256, 384, 512])
def test_scan_parity_dims(gpu, cpu, fo, dim):
    """Every dim class: pure tail (<8), left-over chunks, tails (dim%8), fast path (128/256/384)."""
    rng = np.random.default_rng(dim)
    n = 777
    rows = rng.uniform(-1, 1, (n, dim)).astype(np.float32)
    rows /= np.linalg.norm(rows, axis=1, keepdims=True)
    slab = fo.encode_f16(rows)
    q = rng.uniform(-1, 1, dim).astype(np.float32)
    for tail_fma in (True, False):
        for k in (1, 10, 100):
            want = cpu.search_bits(slab, q, k, tail_fma=tail_fma)
            got = gpu.search_bits(slab, q, k, tail_fma=tail_fma)
            assert_same_hits(got, want, f"dim={dim} k={k} tail_fma={tail_fma}")


@pytest.mark.parametrize("order", [0, 1, 2, 3, 4])
def test_scan_parity_reduce_orders(gpu, cpu, fo, order):
    slab, _ = fo.synth_rows(0, 11, 0, 3000, 384)
    q = fo.normalize(fo.raw_vector(0xBEEF, 384))
    want = cpu.search_bits(slab, q, 50, reduce_order=order)
    got = gpu.search_bits(slab, q, 50, reduce_order=order)
    assert_same_hits(got, want, f"order={order}")


@pytest.mark.parametrize("n", [1, 7, 63, 64, 65, 127, 129, 1000, 20011])
def test_scan_parity_row_counts(gpu, cpu, fo, n):
    """N below / at / just past the tile size and not a multiple of it; k around N."""
    slab, _ = fo.synth_rows(1, 5, 0, n, 384)
    q = fo.clustered_query(1, 384)
    for k in sorted({1, 10, min(n, 100), n, n + 5}):
        want = cpu.search_bits(slab, q, k)
        got = gpu.search_bits(slab, q, k)
        assert_same_hits(got, want, f"n={n} k={k}")


def test_scan_parity_large_k_paths(gpu, cpu, fo):
    """k = 1000 (fused path, cap 2048), k = 1024 (limit of the fused path), k = 1500 and
    k >= n (score-all + radix sort arm, search.rs:449-473)."""
    n = 6000
    slab, _ = fo.synth_rows(1, 9, 0, n, 256)
    q = fo.clustered_query(2, 256)
    tomb = np.zeros(n, dtype=bool)
    tomb[5::11] = True
    for k in (1000, 1024, 1025, 1500, n, n + 100):
        for tb in (None, tomb):
            want = cpu.search_bits(slab, q, k, tombstones=tb)
            got = gpu.search_bits(slab, q, k, tombstones=tb)
            assert_same_hits(got, want, f"k={k} tomb={tb is not None}")


def test_scan_parity_ties_duplicates_and_tombstones(gpu, cpu, fo):
    """Duplicate vectors tie exactly -> lower row wins (search.rs:2741); tombstoned rows never
    appear (search.rs:2163), including tombstoned duplicates of the best row."""
    slab, _ = fo.synth_rows(1, 3, 0, 5000, 384)
    q = fo.clustered_query(0, 384)
    best = cpu.search_bits(slab, q, 1)[0][0]
    for r in (17, 2500, 4999, 64, 63):
        slab[r] = slab[best]
    tomb = np.zeros(5000, dtype=bool)
    tomb[[17, best]] = True
    for tb in (None, tomb):
        want = cpu.search_bits(slab, q, 20, tombstones=tb)
        got = gpu.search_bits(slab, q, 20, tombstones=tb)
        assert_same_hits(got, want)
    assert len(set(bits(want[1])[:3])) == 1  # the planted duplicates really tie


def test_scan_parity_special_values(gpu, cpu, fo):
    """+-inf scores, NaN scores mixed with real ones (NaN -> -inf class, ordered by row)."""
    rng = np.random.default_rng(5)
    rows = rng.uniform(-1, 1, (300, 128)).astype(np.float32)
    slab = fo.encode_f16(rows)
    slab[10, 0] = 0x7C00   # +inf element
    slab[20, 0] = 0xFC00   # -inf element
    slab[30, 0] = 0x7E00   # NaN element
    slab[40, 1] = 0x7E00
    q = rng.uniform(0.1, 1, 128).astype(np.float32)
    for k in (5, 300):
        want = cpu.search_bits(slab, q, k)
        got = gpu.search_bits(slab, q, k)
        assert_same_hits(got, want, f"k={k}")


def test_packed_path_precondition_and_scalar_fallback(gpu, cpu, fo):
    """The f32x2 (FMUL2 + FADD2.FTZ) inner loop is only taken when every non-zero |q_i| >= 2^-76
    (no subnormal can then arise, so FTZ is a no-op); tiny-component and subnormal queries take
    the scalar mul.rn/add.rn loop.  Both must match the oracle bit for bit."""
    slab, _ = fo.synth_rows(0, 5, 0, 4000, 384)
    slab[7, :] = 0x0001          # smallest f16 subnormal everywhere
    slab[9, :] = 0x8001
    base = fo.normalize(fo.raw_vector(0x51, 384))
    for scale, tweak in ((1.0, None), (1.0, 1e-25), (1.0, 1e-39), (1e-30, None), (1e-38, None), (1.0, 2.0 ** -76)):
        q = (base * np.float32(scale)).astype(np.float32)
        if tweak is not None:
            q[::5] = np.float32(tweak)
            q[1::7] = -np.float32(tweak)
        for k in (10, 64):
            want = cpu.search_bits(slab, q, k)
            got = gpu.search_bits(slab, q, k)
            assert_same_hits(got, want, f"scale={scale} tweak={tweak} k={k}")


def test_batched_queries_match_single_queries(gpu, cpu, fo):
    """Batch sizes that exercise every QB grouping (8/4/2/1 remainders)."""
    import frankensearch_b200 as fs

    slab, _ = fo.synth_rows(1, 21, 0, 9000, 384)
    ix = fs.GpuVectorIndex.from_f16_bits(None, slab)
    for b in (1, 2, 3, 5, 8, 11):
        qs = np.stack([fo.clustered_query(100 + i, 384) for i in range(b)])
        rows, scores, counts = ix.search_top_k_batch(qs, 10)
        for i in range(b):
            want = cpu.search_bits(slab, qs[i], 10)
            assert_same_hits((rows[i, :counts[i]].tolist(), scores[i, :counts[i]]), want, f"b={b} i={i}")
    ix.close()


def test_errors_match_reference_contract(gpu):
    import frankensearch_b200 as fs

    ix = fs.GpuVectorIndex.from_vectors(["a", "b"], np.eye(2, 8, dtype=np.float32))
    with pytest.raises(fs.SearchError) as e:      # ensure_query_dimension (search.rs:1602-1610)
        ix.search_top_k(np.ones(4, dtype=np.float32), 1)
    assert e.value.kind == "DimensionMismatch"
    assert ix.search_top_k(np.ones(8, dtype=np.float32), 0) == []          # search.rs:438-440
    hits = ix.search_top_k(np.eye(1, 8, dtype=np.float32)[0], 5)
    assert [h.doc_id for h in hits] == ["a", "b"] and hits[0].index == 0
    ix.close()
    empty = fs.GpuVectorIndex.from_vectors([], np.zeros((0, 8), dtype=np.float32))
    assert empty.search_top_k(np.ones(8, dtype=np.float32), 3) == []
    empty.close()
    with pytest.raises(fs.SearchError):           # write_record rejects non-finite (lib.rs:3647)
        fs.GpuVectorIndex.from_vectors(["x"], np.full((1, 8), np.nan, dtype=np.float32))


def test_device_encode_matches_reference_rne(gpu, fo):
    """from_vectors encodes f32->f16 on the device: must equal vcvtps2ph RNE (simd.rs:2245-2304)."""
    import frankensearch_b200 as fs

    rng = np.random.default_rng(2)
    x = np.concatenate([rng.standard_normal(4096).astype(np.float32) * s for s in (1e-7, 1e-4, 1, 300)])
    x = np.concatenate([x, np.array([65504, 65519.9, 2.9802322e-8, 5.96e-8, 6.1e-5, 1.00048828125], np.float32)])
    x = x[: (x.size // 8) * 8].reshape(-1, 8)
    ix = fs.GpuVectorIndex.from_vectors(None, x)
    got = ix.read_rows_f16(0, x.shape[0])
    ix.close()
    assert np.array_equal(got, fo.encode_f16(x))


# ── two-tier + fusion pipeline parity (BASELINE config 2 shape, scaled to oracle seconds) ───
def test_two_tier_rrf_pipeline_parity(gpu, cpu, fo):
    """sync_searcher.rs:616-1009 flow on synthetic tiers: fast top-3k -> quality rescoring
    (quality_scores_for_hits) -> blend 0.7 -> RRF with a synthetic BM25 list; every stage is
    compared with the oracle (rows, f32 score bits, f64 rrf bits)."""
    import frankensearch_b200 as fs

    n, k = 20000, 20
    fast_slab, _ = fo.synth_rows(1, 1, 0, n, 256)
    qual_slab, _ = fo.synth_rows(1, 7, 0, n, 384)
    ids = [f"doc-{i:06}" for i in range(n)]
    fast_ix = fs.GpuVectorIndex.from_f16_bits(ids, fast_slab)
    qual_ix = fs.GpuVectorIndex.from_f16_bits(ids, qual_slab)
    fetch = max(fs.candidate_count(k, 0, 3), k)
    rng = np.random.default_rng(0)
    for qi in range(3):
        fq, qq = fo.clustered_query(qi, 256), fo.clustered_query(qi, 384)
        fast_hits = fast_ix.search_top_k(fq, fetch)
        o_rows, o_scores = fo.search_top_k(fast_slab, fq, fetch)
        assert [h.index for h in fast_hits] == [int(r) for r in o_rows]
        assert np.array_equal(bits([h.score for h in fast_hits]), bits(o_scores))
        # quality rescoring
        qs = qual_ix.quality_scores_for_hits(qq, fast_hits)
        o_qs, _ = fo.scores_for_rows(qual_slab, qq, o_rows)
        assert np.array_equal(bits(qs), bits(o_qs))
        # blend
        blended = fs.blend_two_tier_aligned(fast_hits, qs, 0.7)
        o_blend = fo.blend_two_tier_aligned([(h.doc_id, h.index, h.score) for h in fast_hits], list(o_qs), 0.7)
        assert [(h.doc_id, h.index) for h in blended] == [(d, i) for d, i, _ in o_blend]
        assert np.array_equal(bits([h.score for h in blended]), bits([s for _, _, s in o_blend]))
        # synthetic BM25 list: ~50% overlap with the semantic list, descending scores
        sem_docs = [h.doc_id for h in fast_hits]
        lex_docs = list(rng.permutation(sem_docs)[: fetch // 2]) + [f"doc-{int(i):06}" for i in
                                                                    rng.integers(0, n, fetch // 2)]
        lexical = [fs.ScoredResult(d, float(len(lex_docs) - i)) for i, d in enumerate(lex_docs)]
        for cfg in (fs.RrfConfig(), fs.RrfConfig(k=10.0, semantic_weight=0.6, tiebreak="Hash")):
            fused = fs.rrf_fuse(lexical, blended, k, 0, cfg)
            o_fused = fo.rrf_fuse([(r.doc_id, r.score) for r in lexical],
                                  [(h.doc_id, h.index, h.score) for h in blended], k, 0, cfg.k,
                                  cfg.lexical_weight, cfg.semantic_weight, 1 if cfg.tiebreak == "Hash" else 0)
            assert [f.doc_id for f in fused] == [f.doc_id for f in o_fused]
            assert [np.float64(f.rrf_score).view(np.uint64) for f in fused] == \
                   [np.float64(f.rrf_score).view(np.uint64) for f in o_fused]
            assert [(f.lexical_rank, f.semantic_rank, f.in_both_sources) for f in fused] == \
                   [(f.lexical_rank, f.semantic_rank, f.in_both_sources) for f in o_fused]
    fast_ix.close()
    qual_ix.close()


def test_rrf_random_parity(gpu, cpu):
    """Random lists with duplicates, overlaps, equal scores: device RRF == oracle, bit for bit."""
    rng = np.random.default_rng(42)
    for trial in range(12):
        n_lex, n_sem = int(rng.integers(0, 400)), int(rng.integers(0, 400))
        pool = [f"d{int(i)}" for i in rng.integers(0, 500, 1000)]
        lexical = [(pool[int(rng.integers(0, 1000))], float(np.float32(rng.integers(0, 50)))) for _ in range(n_lex)]
        sem_docs = list(dict.fromkeys(pool[int(rng.integers(0, 1000))] for _ in range(n_sem)))
        semantic = [(d, i, float(np.float32(1.0 - i * 1e-3))) for i, d in enumerate(sem_docs)]
        for tiebreak in ("LexicalThenId", "Hash"):
            kw = dict(limit=int(rng.integers(1, 120)), offset=int(rng.integers(0, 5)), k=float(rng.choice([0, 1, 60])),
                      w_lex=float(rng.choice([1.0, 0.5, 2.0])), w_sem=float(rng.choice([1.0, 0.3])), tiebreak=tiebreak)
            want = cpu.rrf(lexical, semantic, **kw)
            got = gpu.rrf(lexical, semantic, **kw)
            assert [g[0] for g in got] == [w[0] for w in want], (trial, kw)
            assert [np.float64(g[1]).view(np.uint64) for g in got] == [np.float64(w[1]).view(np.uint64) for w in want]
            assert [g[2:] for g in got] == [w[2:] for w in want]


def test_blend_random_parity(gpu, cpu):
    rng = np.random.default_rng(43)
    for trial in range(10):
        nf, nq = int(rng.integers(1, 300)), int(rng.integers(0, 300))
        fast = [(f"d{int(i)}", int(i), float(np.float32(rng.uniform(-1, 1)))) for i in rng.integers(0, 400, nf)]
        qual = [(f"d{int(i)}", int(i), float(np.float32(rng.uniform(-1, 1)))) for i in rng.integers(0, 400, nq)]
        for alpha in (0.7, 0.25):
            want = cpu.blend(fast, qual, alpha)
            got = gpu.blend(fast, qual, alpha)
            assert [(d, i) for d, i, _ in got] == [(d, i) for d, i, _ in want], trial
            assert np.array_equal(bits([s for _, _, s in got]), bits([s for _, _, s in want]))
        scores = [None if rng.random() < 0.3 else float(np.float32(rng.uniform(-1, 1))) for _ in fast]
        want = cpu.blend_aligned(fast, scores, 0.7)
        got = gpu.blend_aligned(fast, scores, 0.7)
        assert [(d, i) for d, i, _ in got] == [(d, i) for d, i, _ in want]
        assert np.array_equal(bits([s for _, _, s in got]), bits([s for _, _, s in want]))


def test_potion_parity(gpu, cpu):
    rng = np.random.default_rng(8)
    table = rng.standard_normal((3000, 256)).astype(np.float32)
    import frankensearch_b200 as fs

    enc = fs.Model2VecEmbedder(table)
    batches = [list(rng.integers(0, 3300, int(rng.integers(0, 40)))) for _ in range(17)] + [[], [5000, 6000]]
    out = enc.embed_token_ids_batch(batches)
    for ids, got in zip(batches, out):
        want = cpu.potion(table, np.asarray(ids, dtype=np.uint32))
        assert np.array_equal(bits(got), bits(want))
    assert not enc.embed_sync("").any()
    enc.close()


def test_fsvi_file_roundtrip(gpu, cpu, fo, tmp_path):
    """index/tests/fsvi_roundtrip.rs in spirit: write an FSVI v1 file with the reference layout,
    open it on the device, search, resolve doc ids, honour tombstone flags and duplicate doc ids."""
    import frankensearch_b200 as fs
    from frankensearch_b200.fsvi import write_fsvi_v1

    n, dim = 1200, 128
    _, vec = fo.synth_rows(1, 31, 0, n, dim, want_f32=True)
    ids = [f"doc-{i:06}" for i in range(n)]
    ids[700] = ids[100]  # duplicate doc id (soft-delete + rewrite shape): best occurrence wins
    tomb = [i % 97 == 0 for i in range(n)]
    path = str(tmp_path / "idx.fsvi")
    perm = write_fsvi_v1(path, "bench-128", dim, ids, vec, tombstones=tomb)
    ix = fs.GpuVectorIndex.open(path)
    assert ix.record_count() == n and ix.dimension() == dim
    slab = fo.encode_f16(vec[perm])
    q = fo.clustered_query(4, dim)
    hits = ix.search_top_k(q, 50)
    rows, scores = fo.search_top_k(slab, q, 50, fo.pack_bitmap(np.array(tomb)[perm]))
    want, seen = [], set()
    for r, s in zip(rows, scores):
        d = ids[perm[int(r)]]
        if d in seen:
            continue
        seen.add(d)
        want.append((int(r), d, s))
    assert [(h.index, h.doc_id) for h in hits] == [(r, d) for r, d, _ in want]
    assert np.array_equal(bits([h.score for h in hits]), bits([s for _, _, s in want]))
    ix.close()
    with open(path, "r+b") as f:  # header CRC is checked (lib.rs header_crc_detects_*)
        f.seek(9)
        f.write(b"X")
    with pytest.raises(fs.SearchError) as e:
        fs.GpuVectorIndex.open(path)
    assert e.value.kind == "IndexCorrupted"


def test_sharded_merge_on_one_gpu(gpu, cpu, fo):
    """Two row shards (row_base) searched separately + fsgpu_merge_top_k_device == one index."""
    import torch

    import frankensearch_b200 as fs
    from frankensearch_b200.sharded import ShardedGpuIndex, shard_bounds

    n, k = 30001, 64
    slab, _ = fo.synth_rows(1, 77, 0, n, 384)
    slab[20000] = slab[5]  # cross-shard exact tie
    qs = np.stack([fo.clustered_query(50 + i, 384) for i in range(5)])
    dq = torch.from_numpy(qs).cuda()
    shards = []
    for r in range(3):
        lo, hi = shard_bounds(n, 3, r)
        shards.append(fs.GpuVectorIndex.from_f16_bits(None, slab[lo:hi], row_base=lo))
    keys = torch.stack([s.search_top_k_device(dq, k)[0] for s in shards])
    scores = torch.stack([s.search_top_k_device(dq, k)[1][..., 1].contiguous().view(torch.float32) for s in shards])
    merged_keys, merged_hits, counts = ShardedGpuIndex(shards[0])._cuda_merge(keys, scores, k)
    torch.cuda.synchronize()
    hits = merged_hits.cpu().numpy()
    for b in range(5):
        want = cpu.search_bits(slab, qs[b], k)
        got_rows = hits[b, :, 0].view(np.uint32).tolist()
        got_scores = hits[b, :, 1].copy().view(np.float32)
        assert int(counts[b]) == k
        assert_same_hits((got_rows, got_scores), want, f"b={b}")
    for s in shards:
        s.close()


# ── BASELINE.json sizes ─────────────────────────────────────────────────────────────────────
def test_config2_1m_x_384_full_oracle_parity(gpu, cpu, fo):
    """BASELINE config 2: 1 M x 384, top-100, corpus generated ON THE DEVICE by the reference's
    bench generator and compared with the full CPU oracle scan of the oracle-generated corpus."""
    import torch

    import frankensearch_b200 as fs

    n, dim, k = 1_000_000, 384, 100
    slab_gpu = torch.empty((n, dim), dtype=torch.int16, device="cuda")
    fs._ffi.check(fs._ffi.lib().fsgpu_synth_rows_device(0, 1, 1, 0, n, dim, 64, 0.30, slab_gpu.data_ptr(), None))
    slab_cpu, _ = fo.synth_rows(1, 1, 0, n, dim)
    sample = np.r_[0:64, n // 2:n // 2 + 64, n - 64:n]
    assert np.array_equal(slab_gpu[torch.from_numpy(sample).cuda()].cpu().numpy().view(np.uint16), slab_cpu[sample])
    ix = fs.GpuVectorIndex.from_device_tensor(slab_gpu)
    qs = np.stack([fo.clustered_query(i, dim) for i in range(6)])
    rows, scores, counts = ix.search_top_k_batch(qs, k)
    for b in range(6):
        want = cpu.search_bits(slab_cpu, qs[b], k)
        assert_same_hits((rows[b, :counts[b]].tolist(), scores[b, :counts[b]]), want, f"b={b}")
    ix.close()


def test_config3_10m_x_384_properties_and_full_oracle(gpu, cpu, fo):
    """BASELINE config 3: 10 M x 384 (7.68 GB).  Size-independent properties: (a) device corpus ==
    oracle generator on sampled rows, (b) results sorted by the reference order, (c) returned
    scores are bit-exact oracle dots of the returned rows, (d) planted needles surface in tie
    order, (e) two half-corpus searches merged == the full search; plus one full CPU oracle scan
    of the downloaded slab for exact id parity at full size."""
    import torch

    import frankensearch_b200 as fs

    n, dim, k = 10_000_000, 384, 100
    slab_gpu = torch.empty((n, dim), dtype=torch.int16, device="cuda")
    fs._ffi.check(fs._ffi.lib().fsgpu_synth_rows_device(0, 1, 1, 0, n, dim, 64, 0.30, slab_gpu.data_ptr(), None))
    sample = np.unique(np.r_[0:16, n - 16:n, np.random.default_rng(0).integers(0, n, 200)])
    ref_rows = np.concatenate([fo.synth_rows(1, 1, int(r), 1, dim)[0] for r in sample])
    assert np.array_equal(slab_gpu[torch.from_numpy(sample).cuda()].cpu().numpy().view(np.uint16), ref_rows)
    q = fo.clustered_query(7, dim)
    # (d) plant needles: exact copies of one strong row at a tile edge, the first and the last row
    # (an adopted slab is immutable while an index uses it — the index keeps statistics and int8
    # codes of it — so the needles are planted between two indexes)
    ix = fs.GpuVectorIndex.from_device_tensor(slab_gpu)
    best = int(ix.search_top_k_batch(q, 1)[0][0, 0])
    ix.close()
    for r in (0, 63, 64, 5_000_000, n - 1):
        slab_gpu[r] = slab_gpu[best]
    torch.cuda.synchronize()
    ix = fs.GpuVectorIndex.from_device_tensor(slab_gpu)
    rows, scores, counts = ix.search_top_k_batch(np.stack([q, fo.clustered_query(8, dim)]), k)
    assert counts.tolist() == [k, k]
    planted = sorted({0, 63, 64, 5_000_000, n - 1, best})
    assert rows[0, :len(planted)].tolist() == planted
    # (b) + (c)
    for b, qq in enumerate((q, fo.clustered_query(8, dim))):
        got = slab_gpu[torch.from_numpy(rows[b].astype(np.int64)).cuda()].cpu().numpy().view(np.uint16)
        exact, _ = fo.scores_for_rows(got, qq, np.arange(k, dtype=np.uint64))
        assert np.array_equal(bits(exact), bits(scores[b]))
        keys = no.order_keys(scores[b], rows[b])
        assert np.all(keys[:-1] < keys[1:])
    # (e) halves merged == full
    from frankensearch_b200.sharded import ShardedGpuIndex
    dq = torch.from_numpy(np.stack([q])).cuda()
    lo = fs.GpuVectorIndex.from_device_tensor(slab_gpu[: n // 2])
    hi = fs.GpuVectorIndex.from_device_tensor(slab_gpu[n // 2:], row_base=n // 2)
    parts = [s.search_top_k_device(dq, k) for s in (lo, hi)]
    keys = torch.stack([p[0] for p in parts])
    sc = torch.stack([p[1][..., 1].contiguous().view(torch.float32) for p in parts])
    _, mh, _ = ShardedGpuIndex(lo)._cuda_merge(keys, sc, k)
    assert mh.cpu().numpy()[0, :, 0].view(np.uint32).tolist() == rows[0].tolist()
    lo.close()
    hi.close()
    # full-size oracle: download the slab once (7.68 GB) and scan it on the host cores
    host = np.empty((n, dim), dtype=np.uint16)
    step = 1_000_000
    for s in range(0, n, step):
        host[s:s + step] = slab_gpu[s:s + step].cpu().numpy().view(np.uint16)
    want = cpu.search_bits(host, q, k)
    assert_same_hits((rows[0].tolist(), scores[0]), want, "10M full oracle")
    ix.close()


def test_concurrent_callers_share_one_index(gpu, cpu, fo):
    """The reference's index methods take `&self` and are called from many threads at once
    (`Arc<TwoTierIndex>`, searcher.rs:256); calls on one handle are serialised inside the library.
    Eight threads mix single queries, batches (tensor-core pass), filtered searches and re-scoring
    on one index; every result must equal the serial one."""
    import threading

    import frankensearch_b200 as fs

    n, dim = 50000, 128
    slab, _ = fo.synth_rows(1, 5, 0, n, dim)
    ix = fs.GpuVectorIndex.from_f16_bits(None, slab)
    queries = np.stack([fo.clustered_query(q, dim) for q in range(40)])
    allow = np.arange(n) % 3 != 0
    jobs = []
    for t in range(8):
        lo = t * 5
        if t % 4 == 0:
            jobs.append(("single", queries[lo], 10, None))
        elif t % 4 == 1:
            jobs.append(("batch", queries[lo:lo + 5], 100, None))
        elif t % 4 == 2:
            jobs.append(("batch", queries[lo:lo + 5], 10, allow))
        else:
            jobs.append(("rescore", queries[lo], np.arange(lo, lo + 64, dtype=np.uint32), None))

    def run(job):
        kind, q, k, mask = job
        if kind == "rescore":
            s, p = ix.scores_for_rows(q, k)
            return s.view(np.uint32).copy(), p.copy()
        r, s, c = ix.search_top_k_batch(q, k, filter=mask)
        return r.copy(), s.view(np.uint32).copy(), c.copy()

    serial = [run(j) for j in jobs]
    errors, results = [], [[None] * 6 for _ in jobs]

    def worker(i):
        try:
            for rep in range(6):
                results[i][rep] = run(jobs[i])
        except Exception as e:  # pragma: no cover
            errors.append(e)

    threads = [threading.Thread(target=worker, args=(i,)) for i in range(len(jobs))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    for i, want in enumerate(serial):
        for rep in range(6):
            for a, b in zip(results[i][rep], want):
                assert np.array_equal(a, b), (i, rep)
    ix.close()


@pytest.mark.parametrize("dim", [128, 100])
def test_fsvi_f32_quantised_file(gpu, fo, tmp_path, dim):
    """An f32-quantised FSVI v1 file (quantization byte 0, lib.rs:6-43) is scored with the reference's f32
    kernel (dot_product_f32_bytes_f32, simd.rs:581-760 — `mul_add` scalar tail — search.rs:1300-1321):
    rows and score bits equal a NumPy restatement of that kernel for small, large and all-rows limits,
    with tombstones; re-scoring (dot_query_at) uses the same kernel."""
    import frankensearch_b200 as fs
    from frankensearch_b200.fsvi import QUANT_F32, write_fsvi_v1

    n = 2500
    _, vec = fo.synth_rows(1, 71, 0, n, dim, want_f32=True)
    ids = [f"doc-{i:06}" for i in range(n)]
    tomb = [i % 53 == 7 for i in range(n)]
    path = str(tmp_path / "f32.fsvi")
    perm = write_fsvi_v1(path, "bench-f32", dim, ids, vec, tombstones=tomb, quantization=QUANT_F32)
    ix = fs.GpuVectorIndex.open(path)
    rows_f32 = vec[perm]
    live = ~np.array(tomb)[perm]
    for qi in (0, 5):
        q = fo.clustered_query(qi, dim)
        exact = np.array([no.dot_f32_bytes_f32(rows_f32[r], q) for r in range(n)], dtype=np.float32)
        for k in (10, 300, 5000):
            want_rows, want_scores = no.top_k(exact, k, live)
            hits = ix.search_top_k(q, k)
            assert [h.index for h in hits] == [int(r) for r in want_rows], (dim, qi, k)
            assert np.array_equal(bits([h.score for h in hits]), bits(want_scores)), (dim, qi, k)
        got, present = ix.scores_for_rows(q, np.arange(0, n, 97, dtype=np.uint32))
        assert present.all() and np.array_equal(bits(got), bits(exact[::97]))
    with pytest.raises(fs.SearchError):
        ix.read_rows_f16(0, 1)
    ix.close()


def test_classified_lane_replays_reference_cases(gpu, fo):
    """search.rs:2430-2600 (classified_*): k = 0 vs a never-populated index, non-finite query rejected
    (the unclassified lane keeps scoring it), zero-norm query, a filter that rejects everything, an
    all-tombstoned index, a non-empty result carries no reason; plus the census itself."""
    import frankensearch_b200 as fs
    from frankensearch_b200.types import ZeroSignalReason as R

    e0 = [1.0, 0.0, 0.0, 0.0]
    empty = fs.GpuVectorIndex.from_vectors([], np.zeros((0, 4), dtype=np.float32))
    c = empty.search_top_k_classified(e0, 0)
    assert c.hits == [] and c.zero_signal == R.CALLER_REQUESTED_ZERO_K
    c = empty.search_top_k_classified(e0, 5)
    assert c.hits == [] and c.zero_signal == R.NEWLY_CREATED_EMPTY
    empty.close()

    one = fs.GpuVectorIndex.from_vectors(["doc-a"], np.array([[0.1, 0, 0, 0]], dtype=np.float32))
    with pytest.raises(fs.SearchError) as err:
        one.search_top_k_classified([float("nan"), 0, 0, 0], 5)
    assert err.value.kind == "InvalidConfig" and "query" in err.value.message
    with pytest.raises(fs.SearchError):
        one.search_top_k_classified([float("inf"), 0, 0, 0], 5)
    assert len(one.search_top_k([float("nan"), 0, 0, 0], 5)) == 1  # legacy lane is unchanged (search.rs:2474)
    c = one.search_top_k_classified([0.0, 0, 0, 0], 5)
    assert c.hits == [] and c.zero_signal == R.ZERO_NORM_QUERY
    c = one.search_top_k_classified(e0, 5)
    assert len(c.hits) == 1 and c.zero_signal is None
    with pytest.raises(fs.SearchError) as err:
        one.search_top_k_classified([1.0, 0, 0], 5)
    assert err.value.kind == "DimensionMismatch"
    one.close()

    two = fs.GpuVectorIndex.from_vectors(["doc-a", "doc-b"], np.array([[0.1, 0, 0, 0], [0.2, 0, 0, 0]], dtype=np.float32))
    c = two.search_top_k_classified(e0, 5, filter=fs.PredicateFilter("reject-all", lambda _d: False))
    assert c.hits == [] and c.zero_signal == R.FILTER_ELIMINATED_ALL
    st = two.zero_signal_state()
    assert (st.record_count, st.live_count, st.tombstone_count, st.wal_count, st.usable_vector_count) == (2, 2, 0, 0, 2)
    two.soft_delete("doc-a")
    two.soft_delete("doc-b")
    c = two.search_top_k_classified(e0, 5)
    assert c.hits == [] and c.zero_signal == R.ALL_TOMBSTONED
    two.close()

    # TwoTierIndex::search_fast_classified (two_tier.rs:1358-1390) + quality re-scoring through the pair
    slab, vec = fo.synth_rows(1, 81, 0, 500, 128, want_f32=True)
    ids = [f"doc-{i:06}" for i in range(500)]
    fast = fs.GpuVectorIndex.from_vectors(ids, vec)
    quality = fs.GpuVectorIndex.from_vectors(ids, vec[:, ::-1].copy())
    tt = fs.GpuTwoTierIndex(fast, quality)
    q = fo.clustered_query(1, 128)
    c = tt.search_fast_classified(q, 7)
    assert c.zero_signal is None and [h.index for h in c.hits] == [h.index for h in tt.search_fast(q, 7)]
    assert tt.search_fast_classified(q, 0).zero_signal == R.CALLER_REQUESTED_ZERO_K
    assert tt.search_fast_classified(np.zeros(128, dtype=np.float32), 7).zero_signal == R.ZERO_NORM_QUERY
    with pytest.raises(fs.SearchError):
        tt.search_fast_classified(np.full(128, np.nan, dtype=np.float32), 7)
    qs = tt.quality_scores_for_hits(q, c.hits)
    want, _ = fo.scores_for_rows(fo.encode_f16(vec[:, ::-1].copy()), q, [h.index for h in c.hits], tail_fma=False)
    assert np.array_equal(bits(qs), bits(want))
    assert tt.has_quality_index() and tt.doc_count() == 500
    fast.close()
    quality.close()


def test_reduce_order_probe_on_the_device(gpu):
    """The probe of INTEGRATION.md: a row of eight f16 ones against REDUCE_PROBE returns, for each
    `reduce_order`, exactly the bits listed for that order (every kernel family that scores a row:
    per-query scan, re-scoring)."""
    import frankensearch_b200 as fs

    q = np.array(rc.REDUCE_PROBE, dtype=np.float32)
    for order, name in enumerate(("halves_pairwise", "avx_tree", "halves_sequential", "halves_stride2", "sequential")):
        ix = fs.GpuVectorIndex.from_vectors(None, np.ones((3, 8), dtype=np.float32), reduce_order=name)
        rows, scores, counts = ix.search_top_k_batch(q, 1)
        assert int(scores[0, 0].view(np.uint32)) == rc.REDUCE_PROBE_BITS[order], name
        s, _ = ix.scores_for_rows(q, np.array([2], dtype=np.uint32))
        assert int(s[0].view(np.uint32)) == rc.REDUCE_PROBE_BITS[order], name
        ix.close()


def _hand_assembled_fsvi(version, dim, docs, quant=1, flags=None):
    """An FSVI file assembled BYTE BY BYTE from the layout comment of the reference (lib.rs:6-43 for v1,
    parse_v2_header lib.rs:4229-4447 for v2) — independently of frankensearch_b200/fsvi.py's writer — so
    the device reader is pinned by a second, literal statement of the format."""
    import struct
    import zlib

    from frankensearch_b200.types import fnv1a_hash

    order = sorted(range(len(docs)), key=lambda i: (fnv1a_hash(docs[i][0].encode()), docs[i][0].encode()))
    strings = b"".join(docs[i][0].encode() for i in order)
    n = len(docs)
    if version == 1:
        eid, rev = b"hand-made", b"r1"
        fixed = 4 + 2 + 2 + len(eid) + 2 + len(rev) + 4 + 1 + 3 + 8 + 8 + 4
    else:
        blobs = [b'{"bundle":1}', b'{"space":"x"}', b'{"storage":"f16"}']
        fixed = 332 + sum(len(b) for b in blobs) + 4
    voff = (fixed + 16 * n + len(strings) + 63) // 64 * 64
    if version == 1:
        head = (b"FSVI" + struct.pack("<H", 1) + struct.pack("<H", len(eid)) + eid + struct.pack("<H", len(rev)) + rev +
                struct.pack("<I", dim) + bytes([quant]) + bytes(3) + struct.pack("<Q", n) + struct.pack("<Q", voff))
    else:
        head = b"FSVI" + struct.pack("<H", 2) + struct.pack("<I", fixed) + struct.pack("<H", 1) + bytes([quant, 0])
        head += struct.pack("<H", 7) + struct.pack("<I", dim) + struct.pack("<Q", n) + struct.pack("<Q", voff)
        head += struct.pack("<HHQ", 1, 0, 42) + bytes(range(1, 17))
        head += struct.pack("<III", *(len(b) for b in blobs))
        head += b"".join(bytes([i + 1]) * 32 for i in range(8))
        assert len(head) == 332
        head += b"".join(blobs)
    head += struct.pack("<I", zlib.crc32(head) & 0xFFFFFFFF)
    assert len(head) == fixed
    records, off = b"", 0
    for rank, i in enumerate(order):
        b = docs[i][0].encode()
        records += struct.pack("<QIHH", fnv1a_hash(b), off, len(b), (flags or [0] * n)[i])
        off += len(b)
    body = np.stack([np.asarray(docs[i][1], dtype=np.float32) for i in order])
    slab = body.astype(np.float16).tobytes() if quant == 1 else body.tobytes()
    pad = bytes(voff - fixed - len(records) - len(strings))
    return head + records + strings + pad + slab, order


@pytest.mark.parametrize("version", [1, 2])
def test_reader_against_hand_assembled_fsvi_bytes(gpu, tmp_path, version):
    """index/tests/fsvi_roundtrip.rs replayed against literal bytes: single f16 record round trip (:35),
    multiple records found by doc id in hash order (:96), search_returns_closest_vector (:401),
    search_respects_limit (:443), tombstone flag honoured, corrupted header (:543: byte 6 flipped) and
    truncated file (:574: 8 bytes) detected — for the v1 layout and the v2 identity-header layout."""
    import frankensearch_b200 as fs

    def norm(v):
        v = np.asarray(v, dtype=np.float32)
        return v / np.float32(np.sqrt((v * v).sum()))

    docs = [("north", norm([1, 0, 0, 0])), ("east", norm([0, 1, 0, 0])), ("northeast", norm([1, 1, 0, 0])),
            ("gone", norm([0.9, 0.1, 0, 0]))]
    data, order = _hand_assembled_fsvi(version, 4, docs, flags=[0, 0, 0, 1])
    path = str(tmp_path / f"v{version}.fsvi")
    open(path, "wb").write(data)
    ix = fs.GpuVectorIndex.open(path)
    assert ix.record_count() == 4 and ix.dimension() == 4
    assert sorted(ix.doc_id_at(r) for r in range(4)) == sorted(d for d, _ in docs)
    assert [ix.doc_id_at(r) for r in range(4)] == [docs[i][0] for i in order]
    got = ix.read_rows_f16(0, 4).view(np.float16).astype(np.float32)
    assert np.abs(got - np.stack([docs[i][1] for i in order])).max() < 0.01
    assert ix.search_top_k(norm([0.9, 0.1, 0, 0]), 3)[0].doc_id == "north"  # "gone" is tombstoned
    assert ix.search_top_k(norm([0.1, 0.9, 0, 0]), 3)[0].doc_id == "east"
    assert len(ix.search_top_k([1.0, 0, 0, 0], 2)) == 2 and len(ix.search_top_k([1.0, 0, 0, 0], 100)) == 3
    ix.close()
    bad = bytearray(data)
    bad[6] ^= 0xFF
    open(path, "wb").write(bytes(bad))
    with pytest.raises(fs.SearchError) as e:
        fs.GpuVectorIndex.open(path)
    assert e.value.kind == "IndexCorrupted"
    open(path, "wb").write(data[:8])
    with pytest.raises(fs.SearchError):
        fs.GpuVectorIndex.open(path)


@pytest.mark.parametrize("dim", [1, 255, 256, 257])
def test_potion_reference_edge_inputs(gpu, fo, dim):
    """The edge inputs of the reference's own potion tests (model2vec_embedder.rs:1255-1330): a token row
    of signed zeros, rows below (1e-20) and above (1e-4) the normalisation guard, a NaN row, the guard
    boundary [MIN_POSITIVE, +/-sqrt(EPSILON)], non-finite sums, token counts 1, 2, 3, 4, 511, 512, 513,
    OOV ids dropped, the empty text — bit for bit against the oracle (NaN positions equal)."""
    import frankensearch_b200 as fs

    rng = np.random.default_rng(dim)
    table = rng.standard_normal((12, dim)).astype(np.float32)
    eps, tiny = np.float32(np.finfo(np.float32).eps), np.float32(np.finfo(np.float32).tiny)
    table[1] = np.float32(-0.0)
    table[2] = np.float32(1.0e-20)
    table[3] = np.float32(1.0e-4)
    table[4] = np.float32("nan")
    table[5, : min(3, dim)] = [tiny, np.sqrt(eps), -np.sqrt(eps)][: min(3, dim)]
    table[5, 3:] = 0
    table[6, : min(4, dim)] = [np.float32("nan"), np.float32("inf"), np.float32("-inf"), np.finfo(np.float32).max][: min(4, dim)]
    table[7] = np.finfo(np.float32).max  # sums overflow to +inf for counts > 1
    for d in range(dim):  # arbitrary finite values with alternating sign and consecutive bit patterns
        table[8, d] = np.uint32(0x3F800000 + (d % 256)).view(np.float32) * (1 if d % 2 == 0 else -1)
    cases = [[], [0, 9999, 0]]
    for count in (1, 2, 3, 4, 511, 512, 513):
        cases.append([0] * count)
        cases.append([8] * count)
        cases.append([5] * count)
    cases += [[1, 1], [2], [3], [0, 4, 0], [6], [6, 6, 0], [7], [7, 7], [1, 9, 1, 10], [11, 10, 9, 8, 3, 2]]
    enc = fs.Model2VecEmbedder(table)
    got = enc.embed_token_ids_batch(cases)
    for i, ids in enumerate(cases):
        want = fo.potion_embed(table, np.asarray(ids, dtype=np.uint32))
        nan = np.isnan(want)
        assert np.array_equal(np.isnan(got[i]), nan), (dim, ids[:4], len(ids))
        assert np.array_equal(bits(got[i])[~nan], bits(want)[~nan]), (dim, ids[:4], len(ids))
    enc.close()
