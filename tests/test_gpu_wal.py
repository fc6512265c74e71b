"""Resident WAL rows in the GPU top-k: the slab's top-k and the f32 WAL rows are
merged on the device (search.rs:476-493, :1449-1475), doc-id shadowing / dedup on the host
(search.rs:1503-1558).  Checked against the reference's own WAL tests (tests/ref_cases.py
WAL_SCENARIOS) and bit-for-bit against the oracle model (tests/wal_model.py)."""
import os

import numpy as np
import pytest

import ref_cases as rc
from wal_model import OracleWalIndex, run_scenario

pytestmark = pytest.mark.gpu


def bits(x):
    return np.asarray(x, dtype=np.float32).view(np.uint32)


class GpuWalIndex:
    """GpuVectorIndex behind the (row, score, doc_id) interface of the scenarios."""

    def __init__(self, doc_ids, vectors, dim, reduce_order=0):
        import frankensearch_b200 as fs

        v = np.asarray(vectors, dtype=np.float32).reshape(len(doc_ids), dim)
        self.ix = fs.GpuVectorIndex.from_vectors(list(doc_ids), v, reduce_order=reduce_order)

    def append(self, d, v):
        self.ix.append(d, v)

    def append_batch(self, e):
        self.ix.append_batch(e)

    def soft_delete(self, d):
        return self.ix.soft_delete(d)

    def wal_record_count(self):
        return self.ix.wal_record_count()

    def search_top_k(self, query, k, filter_ids=None):
        f = None if filter_ids is None else (lambda d, ids=set(filter_ids): d in ids)
        return [(h.index, np.float32(h.score), h.doc_id) for h in self.ix.search_top_k(query, k, filter=f)]


@pytest.mark.parametrize("scenario", rc.WAL_SCENARIOS, ids=[s["name"] for s in rc.WAL_SCENARIOS])
def test_wal_known_answers_and_oracle_parity(scenario):
    seen = []
    run_scenario(lambda ids, vecs, dim: GpuWalIndex(ids, vecs, dim), scenario, on_search=lambda s, h: seen.append(h))
    want = []
    run_scenario(lambda ids, vecs, dim: OracleWalIndex(ids, vecs, dim), scenario, on_search=lambda s, h: want.append(h))
    assert len(seen) == len(want)
    for g, o in zip(seen, want):
        assert [(r, d) for r, _, d in g] == [(r, d) for r, _, d in o]
        assert np.array_equal(bits([s for _, s, _ in g]), bits([s for _, s, _ in o]))


def test_wal_full_recall_collect_all_matches_heap_prefix():
    """search.rs:2688."""
    c = rc.wal_full_recall_case()
    g = GpuWalIndex([d for d, _ in c["rows"]], [v for _, v in c["rows"]], 4)
    g.append_batch(c["wal"])
    total = len(c["rows"]) + len(c["wal"])
    full = g.search_top_k(c["query"], total + 10)
    heap = g.search_top_k(c["query"], total - 5)
    assert len(full) == total and full[0][2] == "wal-top" and full[0][0] == 48 and len(heap) == total - 5
    for h, f in zip(heap, full):
        assert h[2] == f[2] and h[0] == f[0] and bits(h[1]) == bits(f[1])
    g.append("wal-tie", [24.0, 0.0, 0.0, 0.0])  # equal score: main row first (wal.rs:557-569)
    ids = [h[2] for h in g.search_top_k(c["query"], total + 10)]
    assert ids.index("doc-024") + 1 == ids.index("wal-tie")


def test_f32_dot_of_wal_rows_is_bit_exact(fo):
    """simd.rs:2512 (same seed, same dims, tails included): a WAL-only index returns the f32 x f32
    dot of its single row — bit-identical to the oracle for every reduce order."""
    import frankensearch_b200 as fs
    from test_oracle_golden import _xorshift_stream

    nxt = _xorshift_stream(rc.DOT_F32_XORSHIFT["seed"])
    for dim in rc.DOT_F32_XORSHIFT["dims"]:
        a = np.array([nxt() for _ in range(dim)], dtype=np.float32)
        b = np.array([nxt() for _ in range(dim)], dtype=np.float32)
        if not np.any(a):
            continue
        for order in (0, 1, 4):
            ix = fs.GpuVectorIndex.from_vectors([], np.zeros((0, dim), dtype=np.float32), reduce_order=order)
            ix.append("w", a)
            hits = ix.search_top_k(b, 1)
            ix.close()
            assert len(hits) == 1 and hits[0].index == 0 and hits[0].doc_id == "w"
            assert bits(hits[0].score) == bits(fo.dot_f32_f32(a, b, order)), (dim, order)


@pytest.mark.parametrize("dim,n,n_wal", [(128, 20000, 300), (384, 6000, 37), (100, 3000, 64)])
def test_random_wal_matches_oracle(fo, dim, n, n_wal):
    """Main slab + WAL rows, single queries (CUDA-core scan), small and large batches (tensor-core
    scan where dim allows), k below and above the WAL size, with and without a filter: raw rows and
    score bits equal the oracle's heap; resolved hits equal the oracle's resolve."""
    import frankensearch_b200 as fs
    from oracle import np_oracle as no

    rng = np.random.default_rng(dim + n)
    slab, _ = fo.synth_rows(1, 7, 0, n, dim)
    vecs = fo.decode_f16(slab)
    doc_ids = [f"doc-{i:06}" for i in range(n)]
    ix = fs.GpuVectorIndex.from_vectors(doc_ids, vecs)
    model = OracleWalIndex(doc_ids, vecs, dim)
    # WAL: a third updates existing docs (shadowing + tombstones), the rest are new; near the queries' centroids
    wal = []
    for w in range(n_wal):
        base = vecs[rng.integers(0, n)] + 0.05 * rng.standard_normal(dim).astype(np.float32)
        v = (base / np.linalg.norm(base)).astype(np.float32)
        wal.append((doc_ids[rng.integers(0, n)] if w % 3 == 0 else f"new-{w:04}", v))
    ix.append_batch(wal[: n_wal // 2])
    model.append_batch(wal[: n_wal // 2])
    for e in wal[n_wal // 2:]:
        ix.append(*e)
        model.append(*e)
    assert ix.soft_delete(wal[1][0]) == model.soft_delete(wal[1][0])
    assert ix.wal_record_count() == model.wal_record_count()
    n_w = model.wal_record_count()
    queries = np.stack([fo.clustered_query(q, dim) for q in range(70)])
    allow = rng.random(n + n_w) < 0.6
    for k in (10, 100, 1000):
        for batch in (1, 5, 70):
            for mask in (None, allow):
                rows, scores, counts = ix.search_top_k_batch(queries[:batch], k, filter=mask)
                for b in range(batch):
                    want_rows, want_scores = model.raw_search(queries[b], k, mask)
                    c = int(counts[b])
                    assert c == len(want_rows), (k, batch, b)
                    assert np.array_equal(rows[b, :c].astype(np.uint64), want_rows), (k, batch, b)
                    assert np.array_equal(bits(scores[b, :c]), bits(want_scores)), (k, batch, b)
    # resolved single-query hits
    for q in range(4):
        got = [(h.index, h.doc_id) for h in ix.search_top_k(queries[q], 50)]
        want = [(r, d) for r, _, d in model.search_top_k(queries[q], 50)]
        assert got == want
    ix.close()


def test_wal_errors():
    import frankensearch_b200 as fs

    ix = fs.GpuVectorIndex.from_vectors(["a"], np.array([[1.0, 0.0, 0.0, 0.0]], dtype=np.float32))
    with pytest.raises(fs.SearchError) as e:  # wal_append_dimension_mismatch (tests/fsvi_roundtrip.rs:299)
        ix.append("b", [1.0, 0.0])
    assert e.value.kind == "DimensionMismatch"
    with pytest.raises(fs.SearchError) as e:  # append_batch_impl validation (lib.rs:2596-2608)
        ix.append("b", [float("nan"), 0.0, 0.0, 0.0])
    assert e.value.kind == "InvalidConfig"
    with pytest.raises(fs.SearchError) as e:
        ix.append("b", [0.0, 0.0, 0.0, 0.0])
    assert e.value.kind == "InvalidConfig"
    assert ix.wal_record_count() == 0
    ix.close()


def test_open_replays_the_wal_sidecar(fo, tmp_path):
    """VectorIndex::open (lib.rs:1833-1878): pending appends in `<index>.wal` are searchable after a
    reopen — updated documents (their main rows were tombstoned by append_batch, lib.rs:2546-2720), new
    documents, last-entry-wins inside the sidecar, a torn tail ignored.  The C entry point refuses to
    open such a file for a host that has not declared it replays the sidecar."""
    import ctypes as C

    import frankensearch_b200 as fs
    from frankensearch_b200 import fsvi

    n, dim = 600, 128
    _, vec = fo.synth_rows(1, 61, 0, n, dim, want_f32=True)
    ids = [f"doc-{i:06}" for i in range(n)]
    rng = np.random.default_rng(2)
    fresh = fo.normalize(rng.standard_normal(dim).astype(np.float32))
    upd = fo.normalize((vec[10] + 0.5 * rng.standard_normal(dim)).astype(np.float32))
    upd2 = fo.normalize((vec[10] + 0.1 * rng.standard_normal(dim)).astype(np.float32))
    path = str(tmp_path / "idx.fsvi")
    tomb = [d in ("doc-000010", "doc-000333") for d in ids]  # rows superseded by the pending appends
    perm = fsvi.write_fsvi_v1(path, "bench-128", dim, ids, vec, tombstones=tomb)
    wal = fsvi.wal_path_for(path)
    fsvi.append_wal_batch(wal, [("doc-000010", upd), ("new-a", fresh)], dim, compaction_gen=1)
    fsvi.append_wal_batch(wal, [("doc-000333", vec[5]), ("doc-000010", upd2)], dim, compaction_gen=1)
    with open(wal, "ab") as f:
        f.write(b"FWB1\x09\x00\x00\x00torn")
    # a host that does not replay the sidecar must not get a silently incomplete index
    h, o = C.c_void_p(), fs._ffi.IndexOptions()
    fs._ffi.lib().fsgpu_index_options_default(C.byref(o))
    rc = fs._ffi.lib().fsgpu_index_open_fsvi(path.encode(), 0, 0, C.byref(o), C.byref(h))
    assert rc == 2 and b"WAL sidecar" in fs._ffi.lib().fsgpu_last_error()
    ix = fs.GpuVectorIndex.open(path)
    assert [d for d, _ in ix.wal_records()] == ["new-a", "doc-000333", "doc-000010"]
    slab = fo.encode_f16(vec[perm])
    wal_rows = np.stack([v for _, v in ix.wal_records()])
    # WAL vectors are stored in the index quantisation (f16) and widened back exactly (wal.rs:1132-1160)
    assert np.array_equal(wal_rows[0], fo.decode_f16(fo.encode_f16(fresh)))
    for q in (fresh, upd2, fo.clustered_query(3, dim)):
        hits = ix.search_top_k(q, 20)
        rows, scores = fo.search_top_k_wal(slab, wal_rows, q, 20, fo.pack_bitmap(np.array(tomb)[perm]))
        assert [h.index for h in hits] == [int(r) for r in rows]
        assert np.array_equal(bits([h.score for h in hits]), bits(scores))
    assert ix.search_top_k(fresh, 1)[0].doc_id == "new-a"
    assert ix.search_top_k(upd2, 1)[0].doc_id == "doc-000010"
    ix.close()
    os.remove(wal)
    fsvi.append_wal_batch(wal, [("new-b", fresh)], dim, compaction_gen=9)  # stale generation: ignored (lib.rs:1856-1878)
    ix = fs.GpuVectorIndex.open(path)
    assert ix.wal_record_count() == 0
    ix.close()
