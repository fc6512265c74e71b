"""int8 forms of the scan (tcgen05.mma kind::i8 batches, dp4a single-query pass 1): corpus codes equal the
reference's quantiser byte for byte, and search results stay EXACT — identical rows and score bits
to the per-query path and the oracle — because the int8 score only selects a candidate superset
under a proven error bound (mma_scan_kernels.cuh) and winners are re-scored in f16."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def bits(x):
    return np.asarray(x, dtype=np.float32).view(np.uint32)


@pytest.fixture()
def i8_env(monkeypatch):
    monkeypatch.setenv("FSGPU_MMA_I8", "1")
    # the defaults keep the batched int8 form for k <= 16 on shards of >= 2.5 M rows (where it pays);
    # these tests exercise it on small corpora and large k too
    monkeypatch.setenv("FSGPU_I8_MIN_ROWS", "0")
    monkeypatch.setenv("FSGPU_I8_MAX_K", "1024")
    yield


def _index(fs, slab, **kw):
    ix = fs.GpuVectorIndex.from_f16_bits(None, slab, **kw)
    assert ix._L.fsgpu_index_int8_ready(ix._h) == 1
    return ix


def test_codes_match_the_reference_quantiser(i8_env, fo):
    """quantize_f16_slab_to_i8 (simd.rs:1842-1859), values as in avx2_quantize_i8_matches_generic
    (simd.rs:2545-2565: xorshift * 3.0) plus exact .5 ties and the clamp edge."""
    import frankensearch_b200 as fs
    from oracle import np_oracle as no
    from test_oracle_golden import _xorshift_stream

    nxt = _xorshift_stream(0x51ED270B9C4DA3F8)
    vals = np.array([nxt() * np.float32(3.0) for _ in range(40 * 128)], dtype=np.float32).reshape(40, 128)
    vals[0, :8] = [3.0, -3.0, 1.5, -1.5, 0.0, 2.9999, 1e-4, -1e-4]
    slab = fo.encode_f16(vals)
    ix = _index(fs, slab)
    codes = np.zeros((40, 128), dtype=np.int8)
    scale = fs._ffi.C.c_float(0.0)
    fs._ffi.check(ix._L.fsgpu_index_read_codes_i8(ix._h, 0, 40, fs._ffi.ptr(codes), fs._ffi.C.byref(scale)))
    want, max_abs = no.quantize_f16_slab_to_i8(slab)
    assert np.array_equal(codes, want)
    assert np.float32(scale.value) == np.float32(max_abs / np.float32(127.0))
    assert codes.max() == 127 and codes.min() == -127
    ix.close()


@pytest.mark.parametrize("n,dim", [(50000, 128), (30000, 256), (20011, 384), (300, 128)])
def test_int8_batches_equal_exact_path(i8_env, fo, n, dim, monkeypatch):
    import frankensearch_b200 as fs

    slab, _ = fo.synth_rows(1, 3, 0, n, dim)
    tomb = np.arange(n) % 11 == 0
    ix = _index(fs, slab, tombstones=tomb)
    queries = np.stack([fo.clustered_query(q, dim) for q in range(300)])
    for k in (1, 10, 100):
        for batch in (3, 64, 129, 300):
            monkeypatch.setenv("FSGPU_MMA_MIN_BATCH", "3")
            ix.profile_read(reset=True)
            rows, scores, counts = ix.search_top_k_batch(queries[:batch], k)
            p = ix.profile_read(reset=True)
            assert p["mma_launches"] >= 1
            monkeypatch.setenv("FSGPU_MMA_MIN_BATCH", "0")
            erows, escores, ecounts = ix.search_top_k_batch(queries[:batch], k)
            assert np.array_equal(counts, ecounts), (k, batch)
            assert np.array_equal(rows, erows), (k, batch)
            assert np.array_equal(bits(scores), bits(escores)), (k, batch)
    monkeypatch.setenv("FSGPU_MMA_MIN_BATCH", "3")
    rows, scores, counts = ix.search_top_k_batch(queries[:8], 20)
    for b in range(8):  # and against the oracle
        wr, ws = fo.search_top_k(slab, queries[b], 20, fo.pack_bitmap(tomb))
        assert np.array_equal(rows[b, :int(counts[b])].astype(np.uint64), wr)
        assert np.array_equal(bits(scores[b, :int(counts[b])]), bits(ws))
    ix.close()


def test_int8_adversarial_inputs_stay_exact(i8_env, fo, monkeypatch):
    """Queries the int8 bound handles badly must still come out exact (wide margins -> bigger
    candidate lists, or the redo path): unnormalised and huge queries, a zero query, a one-hot query,
    a query dominated by one component, near-duplicate rows (dense tie band), an outlier element that
    inflates the corpus scale, non-finite queries."""
    import frankensearch_b200 as fs

    n, dim = 40000, 128
    rng = np.random.default_rng(3)
    slab, _ = fo.synth_rows(1, 9, 0, n, dim)
    vec = fo.decode_f16(slab)
    vec[1000:1400] = vec[1000] + 1e-3 * rng.standard_normal((400, dim)).astype(np.float32)  # near duplicates
    vec[5000:5064] = vec[5000]                                                              # exact duplicates
    vec[77, 5] = 30.0                                                                       # outlier -> coarse scale
    slab = fo.encode_f16(vec)
    ix = _index(fs, slab)
    qs = [fo.clustered_query(i, dim) for i in range(24)]
    qs[1] = qs[1] * np.float32(1000.0)
    qs[2] = qs[2] * np.float32(1e-6)
    qs[3] = np.zeros(dim, dtype=np.float32)
    qs[4] = np.eye(dim, dtype=np.float32)[7]
    qs[5] = qs[5].copy(); qs[5][3] = 500.0
    qs[6] = vec[1000].copy()
    qs[7] = vec[5000].copy()
    qs[8] = qs[8].copy(); qs[8][0] = np.float32("nan")
    qs[9] = qs[9].copy(); qs[9][1] = np.float32("inf")
    qs[10] = qs[10] * np.float32(3e37)
    q = np.stack(qs).astype(np.float32)
    for k in (10, 100):
        monkeypatch.setenv("FSGPU_MMA_MIN_BATCH", "3")
        rows, scores, counts = ix.search_top_k_batch(q, k)
        monkeypatch.setenv("FSGPU_MMA_MIN_BATCH", "0")
        erows, escores, ecounts = ix.search_top_k_batch(q, k)
        assert np.array_equal(counts, ecounts)
        assert np.array_equal(rows, erows)
        assert np.array_equal(bits(scores), bits(escores))
    ix.close()


def test_int8_one_million_rows_batch_1024(i8_env, fo, monkeypatch):
    import torch

    import frankensearch_b200 as fs

    n, dim = 1_000_000, 384
    dev = torch.device("cuda", 0)
    slab = torch.empty((n, dim), dtype=torch.int16, device=dev)
    fs._ffi.check(fs._ffi.lib().fsgpu_synth_rows_device(0, 1, 1, 0, n, dim, 64, 0.30, slab.data_ptr(), None))
    ix = fs.GpuVectorIndex.from_device_tensor(slab)
    assert ix._L.fsgpu_index_int8_ready(ix._h) == 1
    q = torch.from_numpy(np.stack([fo.clustered_query(i, dim) for i in range(1024)])).to(dev)
    monkeypatch.setenv("FSGPU_MMA_MIN_BATCH", "3")
    ix.profile_read(reset=True)
    keys, hits, counts = ix.search_top_k_device(q, 10)
    torch.cuda.synchronize()
    p = ix.profile_read(reset=True)
    monkeypatch.setenv("FSGPU_MMA_I8", "0")  # same index, f16 tensor-core form
    fkeys, fhits, fcounts = ix.search_top_k_device(q, 10)
    torch.cuda.synchronize()
    assert torch.equal(keys, fkeys) and torch.equal(hits, fhits) and torch.equal(counts, fcounts)
    monkeypatch.setenv("FSGPU_MMA_MIN_BATCH", "0")
    ekeys, ehits, ecounts = ix.search_top_k_device(q[:16].contiguous(), 10)
    torch.cuda.synchronize()
    assert torch.equal(keys[:16], ekeys) and torch.equal(hits[:16], ehits)
    assert p["redo_queries"] == 0, p
    # the quad full pass keeps the second sample level's lists and skips its tiles (FSGPU_MMA_CARRY, default on): the
    # same answer as scanning every tile again, and fewer bytes in the full pass's record
    monkeypatch.setenv("FSGPU_MMA_MIN_BATCH", "3")
    monkeypatch.setenv("FSGPU_MMA_I8", "1")
    bytes_of = {}
    for carry in ("1", "0"):
        monkeypatch.setenv("FSGPU_MMA_CARRY", carry)
        ix.profile_read(reset=True)
        ckeys, chits, ccounts = ix.search_top_k_device(q, 10)
        torch.cuda.synchronize()
        pc = ix.profile_read(reset=True)
        assert pc["quad_launches"] == 1 and pc["redo_queries"] == 0, pc
        assert torch.equal(keys, ckeys) and torch.equal(hits, chits) and torch.equal(counts, ccounts)
        bytes_of[carry] = pc["scan_bytes"]
    assert bytes_of["1"] < bytes_of["0"] == n * dim
    ix.close()


def test_int8_quad_10m_batch_1024_against_the_oracle(fo, monkeypatch):
    """The headline configuration (BASELINE configs[2]: 10 M x 384, batch 1024, top-10) on the headline
    kernel (mma_scan_quad_kernel, default switches), compared DIRECTLY with the CPU oracle for nine
    queries spread over every 128-query block position of a quad (sub-block x CTA rank) and the
    batch edges: rows equal, f32 score bits equal."""
    import torch

    import frankensearch_b200 as fs

    for v in ("FSGPU_MMA_I8", "FSGPU_I8_MIN_ROWS", "FSGPU_I8_MAX_K", "FSGPU_MMA_MIN_BATCH", "FSGPU_MMA_QUAD"):
        monkeypatch.delenv(v, raising=False)
    n, dim, batch, k = 10_000_000, 384, 1024, 10
    dev = torch.device("cuda", 0)
    slab = torch.empty((n, dim), dtype=torch.int16, device=dev)
    fs._ffi.check(fs._ffi.lib().fsgpu_synth_rows_device(0, 1, 1, 0, n, dim, 64, 0.30, slab.data_ptr(), None))
    ix = fs.GpuVectorIndex.from_device_tensor(slab)
    assert ix._L.fsgpu_index_int8_ready(ix._h) == 1
    q_np = np.stack([fo.clustered_query(i, dim) for i in range(batch)])
    q = torch.from_numpy(q_np).to(dev)
    ix.profile_read(reset=True)
    keys, hits, counts = ix.search_top_k_device(q, k)
    torch.cuda.synchronize()
    p = ix.profile_read(reset=True)
    assert p["mma_launches"] == 1 and p["i8_launches"] == 1 and p["redo_queries"] == 0, p
    h = hits.cpu().numpy()
    rows, scores = h[..., 0].view(np.uint32), h[..., 1].view(np.float32)
    assert counts.cpu().numpy().tolist() == [k] * batch
    host, _ = fo.synth_rows(1, 1, 0, n, dim)  # the oracle's own copy of the corpus (same generator)
    sample = np.r_[0:4096:97, n - 64:n]
    assert np.array_equal(slab[torch.from_numpy(sample).to(dev)].cpu().numpy().view(np.uint16), host[sample])
    for b in (0, 1, 127, 128, 300, 511, 512, 777, 1023):
        o_rows, o_scores = fo.search_top_k(host, q_np[b], k)
        assert rows[b].tolist() == [int(r) for r in o_rows], f"query {b}: rows differ from the oracle"
        assert np.array_equal(bits(scores[b]), bits(o_scores)), f"query {b}: score bits differ from the oracle"
    ix.close()


@pytest.mark.parametrize("n,dim", [(60000, 128), (25000, 384), (200, 256)])
def test_int8_single_query_path_is_exact(i8_env, fo, n, dim, monkeypatch):
    """One or two queries through the host API: int8 pass 1 (half the bytes), exact gate from the
    re-scored approximate top-k, exact re-score of the listed rows.  Rows and score bits equal the
    oracle, with tombstones, a filter, resident WAL rows, k up to 1000 and k > live rows."""
    import frankensearch_b200 as fs

    rng = np.random.default_rng(n)
    slab, _ = fo.synth_rows(1, 13, 0, n, dim)
    tomb = rng.random(n) < 0.1
    ids = [f"doc-{i:06}" for i in range(n)]
    ix = fs.GpuVectorIndex.from_f16_bits(ids, slab, tombstones=tomb)
    assert ix._L.fsgpu_index_int8_ready(ix._h) == 1
    allow = rng.random(n) < 0.5
    for qi in range(6):
        q = fo.clustered_query(qi, dim)
        if qi == 4:
            q = q * np.float32(250.0)
        if qi == 5:
            q = np.zeros(dim, dtype=np.float32)
        for k in (1, 10, 100, 1000):
            for mask in (None, allow):
                ix.profile_read(reset=True)
                rows, scores, counts = ix.search_top_k_batch(q, k, filter=mask)
                p = ix.profile_read(reset=True)
                if qi < 4 and k <= 100:  # the int8 codes were scanned, and nothing else
                    assert p["scan_launches"] == 1 and p["scan_bytes"] == n * dim, p
                else:  # a huge tie band (zero query) or k = 1000 may overflow the list: + one f16 scan
                    assert p["scan_launches"] in (1, 2), p
                excl = tomb if mask is None else (tomb | ~mask)
                wr, ws = fo.search_top_k(slab, q, k, fo.pack_bitmap(excl))
                c = int(counts[0])
                assert c == len(wr), (qi, k)
                assert np.array_equal(rows[0, :c].astype(np.uint64), wr), (qi, k)
                assert np.array_equal(bits(scores[0, :c]), bits(ws)), (qi, k)
    # two queries per call, and a non-finite query (served by the f16 scan)
    q2 = np.stack([fo.clustered_query(7, dim), fo.clustered_query(8, dim)])
    rows, scores, counts = ix.search_top_k_batch(q2, 10)
    for b in range(2):
        wr, ws = fo.search_top_k(slab, q2[b], 10, fo.pack_bitmap(tomb))
        assert np.array_equal(rows[b, :int(counts[b])].astype(np.uint64), wr)
    qn = fo.clustered_query(9, dim).copy()
    qn[0] = np.float32("nan")
    monkeypatch.setenv("FSGPU_MMA_I8", "0")
    want = ix.search_top_k_batch(qn, 10)
    monkeypatch.setenv("FSGPU_MMA_I8", "1")
    got = ix.search_top_k_batch(qn, 10)
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[2], want[2])
    # resident WAL rows ride along
    from wal_model import OracleWalIndex

    model = OracleWalIndex(ids, fo.decode_f16(slab), dim)
    model.tomb = tomb.copy()
    wal = [(f"new-{w}", fo.decode_f16(slab)[rng.integers(0, n)] * np.float32(1.01)) for w in range(5)]
    ix.append_batch(wal)
    model.append_batch(wal)
    q = fo.clustered_query(3, dim)
    rows, scores, counts = ix.search_top_k_batch(q, 10)
    wr, ws = model.raw_search(q, 10)
    assert np.array_equal(rows[0, :int(counts[0])].astype(np.uint64), wr) and np.array_equal(bits(scores[0, :int(counts[0])]), bits(ws))
    ix.close()


def test_int8_single_query_list_overflow_falls_back(i8_env, fo):
    """A corpus of identical rows: every row clears the gate, the position list overflows, and the
    call re-runs on the f16 scan — same hits."""
    import frankensearch_b200 as fs

    n, dim = 70000, 128
    row = fo.clustered_query(1, dim)
    slab = fo.encode_f16(np.repeat(row[None, :], n, axis=0))
    ix = fs.GpuVectorIndex.from_f16_bits(None, slab)
    assert ix._L.fsgpu_index_int8_ready(ix._h) == 1
    rows, scores, counts = ix.search_top_k_batch(row, 10)
    assert int(counts[0]) == 10 and list(rows[0]) == list(range(10))
    wr, ws = fo.search_top_k(slab, row, 10)
    assert np.array_equal(bits(scores[0]), bits(ws))
    ix.close()


def test_int8_small_batches_share_one_pass(i8_env, fo, monkeypatch):
    """Batches of 2..7 through the host API with the multi-query pass switched on (FSGPU_I8_QB=4,
    off by default: it is ALU-bound): groups of 4 / 2 / 1 queries per int8 pass; every query's hits
    equal the oracle's."""
    import frankensearch_b200 as fs

    monkeypatch.setenv("FSGPU_I8_QB", "4")
    monkeypatch.setenv("FSGPU_I8_MAX_BATCH", "8")

    n, dim = 80000, 384
    slab, _ = fo.synth_rows(1, 21, 0, n, dim)
    tomb = np.arange(n) % 13 == 0
    ix = fs.GpuVectorIndex.from_f16_bits(None, slab, tombstones=tomb)
    assert ix._L.fsgpu_index_int8_ready(ix._h) == 1
    queries = np.stack([fo.clustered_query(q, dim) for q in range(8)])
    for batch, passes in ((2, 1), (3, 2), (4, 1), (5, 2), (7, 3)):
        for k in (10, 100):
            ix.profile_read(reset=True)
            rows, scores, counts = ix.search_top_k_batch(queries[:batch], k)
            p = ix.profile_read(reset=True)
            assert p["scan_launches"] == passes and p["scan_bytes"] == passes * n * dim, (batch, p)
            for b in range(batch):
                wr, ws = fo.search_top_k(slab, queries[b], k, fo.pack_bitmap(tomb))
                c = int(counts[b])
                assert np.array_equal(rows[b, :c].astype(np.uint64), wr), (batch, k, b)
                assert np.array_equal(bits(scores[b, :c]), bits(ws)), (batch, k, b)
    ix.close()


def test_randomised_soak_on_the_int8_form(i8_env):
    """tools/soak_batched.py with the int8 batched form forced on for every size and k: adversarial
    corpora (ascending-by-score order, duplicates, extreme norms, exact ties), tombstones, filters —
    rows, order and f32 score bits must equal the per-query path."""
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "soak_batched.py"), "35", "11"],
                       capture_output=True, text=True, timeout=600, env=dict(os.environ))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "soak ok" in r.stdout
