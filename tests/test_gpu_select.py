"""Large-k searches (128 < k <= 4096) through the radix-select path (select_kernels.cuh): rows and
score bits equal the oracle (= the reference's sort-all-live-rows order, search.rs:1655-1686) for
int8-coded and f16-only indexes, tie bands across the k-th place, tombstones, filters, WAL rows,
non-finite and zero queries, k above the live row count; plus limit >= record_count > 8192 on an index
with WAL rows (search.rs:449-493)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def bits(x):
    return np.asarray(x, dtype=np.float32).view(np.uint32)


def _check(fo, ix, slab, queries, k, tomb=None, what=""):
    rows, scores, counts = ix.search_top_k_batch(queries, k)
    bm = None if tomb is None else fo.pack_bitmap(tomb)
    for b in range(queries.shape[0]):
        o_rows, o_scores = fo.search_top_k(slab, queries[b], k, bm)
        c = int(counts[b])
        assert c == len(o_rows), (what, b, c, len(o_rows))
        assert rows[b, :c].tolist() == [int(r) for r in o_rows], (what, b)
        assert np.array_equal(bits(scores[b, :c]), bits(o_scores)), (what, b)


@pytest.mark.parametrize("n,dim", [(70001, 384), (300000, 128), (5000, 256), (40000, 100)])
@pytest.mark.parametrize("k", [129, 1000, 3000, 4096])
def test_select_path_matches_oracle(fo, n, dim, k):
    import frankensearch_b200 as fs

    slab, _ = fo.synth_rows(1, 31, 0, n, dim)
    # a tie band across every k-th place under test: 200 exact copies of a strong row per query cluster
    q0 = fo.clustered_query(0, dim)
    best = int(fo.search_top_k(slab, q0, 1)[0][0])
    rng = np.random.default_rng(n + k)
    for r in rng.choice(n, size=200, replace=False):
        slab[r] = slab[best]
    tomb = np.arange(n) % 13 == 5
    ix = fs.GpuVectorIndex.from_f16_bits(None, slab, tombstones=tomb)
    assert (ix._L.fsgpu_index_int8_ready(ix._h) == 1) == (dim % 128 == 0)
    queries = np.stack([q0, fo.clustered_query(1, dim), np.zeros(dim, dtype=np.float32)])
    _check(fo, ix, slab, queries[:1], k, tomb, f"single n={n} dim={dim} k={k}")
    _check(fo, ix, slab, queries, k, tomb, f"batch n={n} dim={dim} k={k}")
    ix.close()


def test_select_path_special_queries_and_k_above_live(fo):
    import frankensearch_b200 as fs

    n, dim = 3000, 128
    slab, _ = fo.synth_rows(0, 7, 0, n, dim)
    tomb = np.arange(n) % 2 == 0  # 1500 live rows
    ix = fs.GpuVectorIndex.from_f16_bits(None, slab, tombstones=tomb)
    q = fo.clustered_query(3, dim)
    nanq = q.copy(); nanq[5] = np.float32("nan")
    infq = q.copy(); infq[9] = np.float32("inf")
    hugeq = q * np.float32(3e37)
    tinyq = q * np.float32(1e-30)
    for k in (200, 1499, 1500, 1501, 4000):
        _check(fo, ix, slab, np.stack([q, hugeq, tinyq]), k, tomb, f"k={k}")
        rows, scores, counts = ix.search_top_k_batch(np.stack([nanq, infq]), k)
        for b, qq in enumerate((nanq, infq)):
            o_rows, o_scores = fo.search_top_k(slab, qq, k, fo.pack_bitmap(tomb))
            c = int(counts[b])
            assert c == len(o_rows)
            assert rows[b, :c].tolist() == [int(r) for r in o_rows], (k, b)
            # NaN scores sort last in row order (search.rs:2767); the payload bits of a NaN are the
            # platform's (x86 default NaN vs the GPU's canonical NaN), everything else is bit-equal
            got_nan, want_nan = np.isnan(scores[b, :c]), np.isnan(o_scores)
            assert np.array_equal(got_nan, want_nan), (k, b)
            assert np.array_equal(bits(scores[b, :c])[~got_nan], bits(o_scores)[~want_nan]), (k, b)
    ix.close()


def test_select_path_device_batches_and_forms(fo, monkeypatch):
    """Device API, batches past the tensor-core path's k ceiling (k = 3000 = the fetch of a top-1000
    search, sync_searcher.rs:654): int8 form == f16 form == oracle."""
    import torch

    import frankensearch_b200 as fs

    n, dim, k = 200_000, 384, 3000
    slab, _ = fo.synth_rows(1, 41, 0, n, dim)
    ix = fs.GpuVectorIndex.from_f16_bits(None, slab)
    q_np = np.stack([fo.clustered_query(i, dim) for i in range(5)])
    q = torch.from_numpy(q_np).cuda()
    keys, hits, counts = ix.search_top_k_device(q, k)
    torch.cuda.synchronize()
    monkeypatch.setenv("FSGPU_SELECT_I8", "0")
    fkeys, fhits, fcounts = ix.search_top_k_device(q, k)
    torch.cuda.synchronize()
    assert torch.equal(keys, fkeys) and torch.equal(hits, fhits) and torch.equal(counts, fcounts)
    h = hits.cpu().numpy()
    for b in range(5):
        o_rows, o_scores = fo.search_top_k(slab, q_np[b], k)
        assert h[b, :, 0].view(np.uint32).tolist() == [int(r) for r in o_rows]
        assert np.array_equal(h[b, :, 1].view(np.uint32), bits(o_scores))
    ix.close()


@pytest.mark.parametrize("k", [3000, 9000, 20000])
def test_large_limit_with_wal_rows_and_filter(fo, k):
    """limit >= record_count > 8192 on an index holding WAL rows (search.rs:449-493, :1449-1475), and a
    large limit with a filter: the WAL merge must not depend on a shared-memory window."""
    import frankensearch_b200 as fs

    n, dim = 12000, 128
    slab, f32 = fo.synth_rows(1, 51, 0, n, dim, want_f32=True)
    ids = [f"doc-{i:06}" for i in range(n)]
    ix = fs.GpuVectorIndex.from_f16_bits(ids, slab)
    rng = np.random.default_rng(3)
    wal = []
    for j in range(40):
        v = rng.standard_normal(dim).astype(np.float32)
        v /= np.linalg.norm(v)
        wal.append((f"new-{j:03}", v))
    ix.append_batch(wal)
    ix.append("doc-000007", wal[0][1])  # an update: main row 7 is tombstoned and shadowed
    q = fo.clustered_query(2, dim)
    mask = np.arange(n) % 3 != 1
    for flt in (None, mask):
        hits = ix.search_top_k(q, k, filter=flt)
        excl = np.zeros(n, dtype=bool)
        excl[7] = True
        if flt is not None:
            excl |= ~flt
        wal_rows = np.stack([v for _, v in ix.wal_records()])
        o_rows, o_scores = fo.search_top_k_wal(slab, wal_rows, q, k, fo.pack_bitmap(excl), None, 1, 0, True)
        assert [h.index for h in hits] == [int(r) for r in o_rows], (k, flt is not None)
        assert np.array_equal(bits([h.score for h in hits]), bits(o_scores))
    ix.close()
