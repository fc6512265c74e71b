"""Doc-id-hash filters evaluated on the device and the selective gather arm:
BitsetFilter (crates/frankensearch-core/src/filter.rs:330-383), the filtered scan
(search.rs:1329-1447) and try_gather_filtered (search.rs:1114-1255).  Both arms must equal the
oracle's scan over the rows whose hash is allowed — bit-exact rows and scores."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def bits(x):
    return np.asarray(x, dtype=np.float32).view(np.uint32)


def test_selective_bitset_filter_uses_gather_and_equals_scan(fo):
    """search.rs:2375 selective_bitset_filter_uses_file_backed_gather: 256 rows, 3 allowed doc ids."""
    import frankensearch_b200 as fs

    rows = [(f"doc-{i:03}", [(i % 31) / 31.0, (i // 31) / 10.0, 0.25, 0.5]) for i in range(256)]
    ids = [d for d, _ in rows]
    vec = np.array([v for _, v in rows], dtype=np.float32)
    ix = fs.GpuVectorIndex.from_vectors(ids, vec)
    filt = fs.BitsetFilter.from_doc_ids(["doc-003", "doc-097", "doc-203"])
    query = [0.7, 0.3, 0.0, 0.0]
    gather = ix.search_top_k(query, 3, filter=filt)
    assert ix.last_filter_arm == "gather"
    scan = ix.search_top_k(query, 3, filter=lambda d: filt.matches(d))  # host-evaluated bitmap -> filtered scan
    assert ix.last_filter_arm == "bitmap"
    assert [h.doc_id for h in gather] == [h.doc_id for h in scan] and len(gather) == 3
    assert np.array_equal(bits([h.score for h in gather]), bits([h.score for h in scan]))
    slab = fo.encode_f16(vec)
    allow = np.array([filt.matches(d) for d in ids])
    want_rows, want_scores = fo.search_top_k(slab, np.array(query, dtype=np.float32), 3, fo.pack_bitmap(~allow), 1, 0, False)
    assert [h.index for h in gather] == [int(r) for r in want_rows]
    assert np.array_equal(bits([h.score for h in gather]), bits(want_scores))
    ix.close()


def test_bitset_filter_single_doc():
    """search.rs:2343 bitset_filter_skips_doc_id_decode_for_non_matching_records (the filter half)."""
    import frankensearch_b200 as fs

    ix = fs.GpuVectorIndex.from_vectors(["doc-a", "doc-b"], np.array([[1.0, 0.0], [0.0, 1.0]], dtype=np.float32))
    hits = ix.search_top_k([1.0, 0.0], 10, filter=fs.BitsetFilter.from_doc_ids(["doc-a"]))
    assert [h.doc_id for h in hits] == ["doc-a"]
    assert ix.search_top_k([1.0, 0.0], 10, filter=fs.BitsetFilter.from_doc_ids(["nope"])) == []
    assert ix.search_top_k([1.0, 0.0], 10, filter=fs.BitsetFilter.from_doc_ids([])) == []
    ix.close()


@pytest.mark.parametrize("dim,n", [(128, 20000), (100, 3000)])
def test_hash_filter_matches_oracle_both_arms(fo, dim, n):
    import frankensearch_b200 as fs

    rng = np.random.default_rng(n)
    slab, _ = fo.synth_rows(1, 11, 0, n, dim)
    vec = fo.decode_f16(slab)
    ids = [f"doc-{i:06}" for i in range(n)]
    tomb = rng.random(n) < 0.05
    ix = fs.GpuVectorIndex.from_vectors(ids, vec, tombstones=tomb)
    queries = np.stack([fo.clustered_query(q, dim) for q in range(70)])
    for n_allowed in (1, 5, n // 50 - 1, n // 50, n // 4):
        chosen = rng.choice(n, size=n_allowed, replace=False)
        filt = fs.BitsetFilter.from_doc_ids([ids[i] for i in chosen] + ["not-in-the-index"])
        allow = np.zeros(n, dtype=bool)
        allow[chosen] = True
        excl = fo.pack_bitmap(tomb | ~allow)
        want_arm = "gather" if (n_allowed + 1) * 50 < n else "scan"
        for k in (10, 100):
            for batch in (1, 5, 70):
                rows, scores, counts = ix.search_top_k_batch(queries[:batch], k, filter=filt)
                assert ix.last_filter_arm == want_arm, (n_allowed, ix.last_filter_arm)
                for b in range(batch):
                    wr, ws = fo.search_top_k(slab, queries[b], k, excl, 2, 0, False)
                    c = int(counts[b])
                    assert c == len(wr), (n_allowed, k, batch, b)
                    assert np.array_equal(rows[b, :c].astype(np.uint64), wr), (n_allowed, k, batch, b)
                    assert np.array_equal(bits(scores[b, :c]), bits(ws)), (n_allowed, k, batch, b)
    ix.close()


def test_hash_filter_with_wal_rows(fo):
    """filter_works_with_wal_and_main_combined (search.rs:2925) through BitsetFilter, plus a larger
    randomized case: WAL rows are filtered by doc id on the host, main rows by hash on the device."""
    import frankensearch_b200 as fs
    from wal_model import OracleWalIndex

    ix = fs.GpuVectorIndex.from_vectors(["doc-a", "doc-b"], np.array([[1.0, 0, 0, 0], [0.5, 0, 0, 0]], dtype=np.float32))
    ix.append("doc-c", [0.9, 0.0, 0.0, 0.0])
    hits = ix.search_top_k([1.0, 0.0, 0.0, 0.0], 10, filter=fs.BitsetFilter.from_doc_ids(["doc-a", "doc-c"]))
    assert [h.doc_id for h in hits] == ["doc-a", "doc-c"]
    ix.close()

    n, dim = 6000, 128
    rng = np.random.default_rng(5)
    slab, _ = fo.synth_rows(1, 3, 0, n, dim)
    vec = fo.decode_f16(slab)
    ids = [f"doc-{i:06}" for i in range(n)]
    ix = fs.GpuVectorIndex.from_vectors(ids, vec)
    model = OracleWalIndex(ids, vec, dim)
    wal = [(ids[rng.integers(0, n)] if w % 2 else f"new-{w:03}", vec[rng.integers(0, n)] * np.float32(0.99)) for w in range(40)]
    ix.append_batch(wal)
    model.append_batch(wal)
    allowed_ids = [ids[i] for i in rng.choice(n, size=60, replace=False)] + [d for d, _ in wal[::3]]
    filt = fs.BitsetFilter.from_doc_ids(allowed_ids)
    allow = np.array([filt.matches(d) for d in ids] + [filt.matches(d) for d, _ in model.wal], dtype=bool)
    queries = np.stack([fo.clustered_query(q, dim) for q in range(6)])
    rows, scores, counts = ix.search_top_k_batch(queries, 20, filter=filt)
    assert ix.last_filter_arm == "gather"
    for b in range(6):
        wr, ws = model.raw_search(queries[b], 20, allow)
        c = int(counts[b])
        assert np.array_equal(rows[b, :c].astype(np.uint64), wr) and np.array_equal(bits(scores[b, :c]), bits(ws))
    got = [(h.index, h.doc_id) for h in ix.search_top_k(queries[0], 20, filter=filt)]
    assert got == [(r, d) for r, _, d in model.search_top_k(queries[0], 20, allowed_ids)]
    ix.close()


def test_many_rows_sharing_one_hash_fall_back_to_the_scan(fo):
    """More rows carry an allowed hash than the gather list holds: the call re-runs as the filtered
    scan and returns the same hits a bitmap filter gives."""
    import frankensearch_b200 as fs

    n, dim = 8000, 64
    slab, _ = fo.synth_rows(0, 9, 0, n, dim)
    vec = fo.decode_f16(slab)
    ix = fs.GpuVectorIndex.from_vectors([f"d{i}" for i in range(n)], vec)
    hashes = np.arange(n, dtype=np.uint64) + np.uint64(1000)
    hashes[1000:2000] = np.uint64(7)  # 1000 rows collide on one hash
    fs._ffi.check(ix._L.fsgpu_index_set_doc_hashes(ix._h, fs._ffi.ptr(hashes)))
    ix._hashes_on_device = True
    q = fo.clustered_query(1, dim)
    rows, scores, counts = ix.search_top_k_batch(q, 50, filter=fs.BitsetFilter.from_hashes([7, 1003]))
    assert ix.last_filter_arm == "scan"
    allow = np.zeros(n, dtype=bool)
    allow[1000:2000] = True
    allow[3] = True
    wr, ws = fo.search_top_k(slab, q, 50, fo.pack_bitmap(~allow), 2, 0, False)
    assert np.array_equal(rows[0, :int(counts[0])].astype(np.uint64), wr) and np.array_equal(bits(scores[0, :50]), bits(ws))
    ix.close()


def test_hash_filter_on_an_fsvi_file(fo, tmp_path):
    """An FSVI v1 file brings its record-table hashes (lib.rs:130-174): no upload needed."""
    import frankensearch_b200 as fs
    from frankensearch_b200.fsvi import write_fsvi_v1

    n, dim = 3000, 128
    _, vec = fo.synth_rows(1, 31, 0, n, dim, want_f32=True)
    ids = [f"doc-{i:06}" for i in range(n)]
    path = str(tmp_path / "idx.fsvi")
    perm = write_fsvi_v1(path, "bench-128", dim, ids, vec)
    ix = fs.GpuVectorIndex.open(path)
    chosen = [ids[i] for i in range(0, n, 101)]
    filt = fs.BitsetFilter.from_doc_ids(chosen)
    q = fo.clustered_query(2, dim)
    hits = ix.search_top_k(q, 10, filter=filt)
    assert ix.last_filter_arm == "gather"
    slab = fo.encode_f16(vec[perm])
    allow = np.array([ids[perm[r]] in set(chosen) for r in range(n)])
    wr, ws = fo.search_top_k(slab, q, 10, fo.pack_bitmap(~allow))
    assert [(h.index, h.doc_id) for h in hits] == [(int(r), ids[perm[int(r)]]) for r in wr]
    assert np.array_equal(bits([h.score for h in hits]), bits(ws))
    ix.close()
