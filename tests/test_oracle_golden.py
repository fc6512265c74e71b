"""Pins the CPU oracle (oracle/fs_oracle.cpp) to the reference's own known-answer tests and
cross-checks it against the independent NumPy mirror.  CPU only."""
import json
import math
import os

import numpy as np
import pytest

import ref_cases as rc
from adapters import OracleImpl
from checks import check_blend_aligned, check_blend_case, check_rrf_case, check_scan_case
from oracle import np_oracle as no

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def impl():
    return OracleImpl()


def bits(x):
    return np.asarray(x, dtype=np.float32).view(np.uint32)


# ── f16 conversions ─────────────────────────────────────────────────────────────────────────
def test_f16_widen_is_bit_exact_for_all_patterns(fo):
    """simd.rs:2711 simd_f16_widen_is_bit_exact: all 65 536 patterns (NaN payloads aside: the
    F16C arm quiets signalling NaNs, the portable arm keeps the payload, like half::f16)."""
    pat = np.arange(65536, dtype=np.uint16)
    sw, hw, ref = fo.decode_f16(pat, hw=False), fo.decode_f16(pat, hw=True), no.decode_f16(pat)
    finite = ~np.isnan(ref)
    assert np.array_equal(bits(sw)[finite], bits(ref)[finite])
    assert np.array_equal(bits(hw)[finite], bits(ref)[finite])
    assert np.isnan(sw[~finite]).all() and np.isnan(hw[~finite]).all()


def test_f16_encode_matches_generic(fo):
    """simd.rs:2669 avx2_f16encode_matches_generic + RNE edge cases (ties, subnormals, overflow)."""
    rng = np.random.default_rng(7)
    x = np.concatenate([rng.standard_normal(50000).astype(np.float32) * s for s in (1e-8, 1e-5, 1e-3, 1, 100, 7e4)])
    edge = np.array([65504, 65519.996, 65520, 65536, 2.9802322e-8, 2.98e-8, 5.96e-8, 8.9e-8, 6.1e-5, 6.097e-5,
                     0.0, -0.0, np.inf, -np.inf, 1.0009765625, 1.00048828125, 1.0014648437], dtype=np.float32)
    x = np.concatenate([x, edge, no.decode_f16(np.arange(65536, dtype=np.uint16))[:31744]])
    sw, hw, ref = fo.encode_f16(x, hw=False), fo.encode_f16(x, hw=True), no.encode_f16(x)
    assert np.array_equal(sw, hw) and np.array_equal(sw, ref)


# ── dot kernel ──────────────────────────────────────────────────────────────────────────────
def _xorshift_stream(seed):
    s = seed

    def nxt():
        nonlocal s
        s ^= (s << 13) & 0xFFFFFFFFFFFFFFFF
        s ^= s >> 7
        s ^= (s << 17) & 0xFFFFFFFFFFFFFFFF
        return np.float32(np.float32(s >> 40) / np.float32(1 << 23) - np.float32(1.0))

    return nxt


def test_avx2_dot_matches_scalar_and_numpy(fo):
    """simd.rs:2423 avx2_f16dot_matches_generic (same seed, same dims) — and the NumPy mirror."""
    nxt = _xorshift_stream(rc.DOT_XORSHIFT["seed"])
    for dim in rc.DOT_XORSHIFT["dims"]:
        q = np.array([nxt() for _ in range(dim)], dtype=np.float32)
        row = fo.encode_f16(np.array([nxt() for _ in range(dim)], dtype=np.float32))
        for order in range(5):
            for tail_fma in (False, True):
                a = fo.dot_f16_f32(row, q, order, tail_fma, impl=0)
                b = fo.dot_f16_f32(row, q, order, tail_fma, impl=1)
                c = no.dot_f16_f32(row, q, order, tail_fma)
                assert bits(a) == bits(b) == bits(c), (dim, order, tail_fma)


def test_simd_matches_scalar_f16_literal(fo):
    """simd.rs:3044 simd_matches_scalar_f16 (tolerance 1e-6 against a plain sequential dot)."""
    q = np.array(rc.DOT_LITERAL["query"], dtype=np.float32)
    stored = fo.encode_f16(np.array(rc.DOT_LITERAL["stored"], dtype=np.float32))
    scalar = np.float32(0)
    for x, y in zip(no.decode_f16(stored), q):
        scalar = np.float32(scalar + np.float32(x * y))
    for order in range(5):
        assert abs(float(fo.dot_f16_f32(stored, q, order)) - float(scalar)) < rc.DOT_LITERAL["tol"]


def test_f16_precision_error_is_bounded(fo):
    """simd.rs:3113 f16_precision_error_is_bounded_for_unit_vectors."""
    p = rc.DOT_PRECISION["pattern"]
    stored = np.array([p[i % len(p)] for i in range(384)], dtype=np.float32)
    query = np.array([p[(i + 3) % len(p)] for i in range(384)], dtype=np.float32)
    stored /= np.linalg.norm(stored)
    query /= np.linalg.norm(query)
    f32_dot = float(np.dot(stored.astype(np.float64), query.astype(np.float64)))
    assert abs(f32_dot - float(fo.dot_f16_f32(fo.encode_f16(stored), query))) < rc.DOT_PRECISION["tol"]


def test_dot_golden_fixture(fo):
    """tests/golden/dot_f16_f32.json — produced by tests/golden/make_golden.py from the two
    independent restatements (C++ and NumPy) agreeing bit for bit."""
    with open(os.path.join(GOLDEN, "dot_f16_f32.json")) as f:
        g = json.load(f)
    for case in g["cases"]:
        row = np.array(case["row_bits"], dtype=np.uint16)
        q = np.array(case["query_bits"], dtype=np.uint32).view(np.float32)
        for order, want in enumerate(case["score_bits_by_order"]):
            assert int(bits(fo.dot_f16_f32(row, q, order, True))) == want


# ── scan / top-k ────────────────────────────────────────────────────────────────────────────
@pytest.mark.parametrize("case", rc.SCAN_CASES, ids=[c["name"] for c in rc.SCAN_CASES])
def test_scan_known_answers(impl, case):
    check_scan_case(impl, case)


def test_parallel_and_sequential_paths_match(impl):
    """search.rs:2233."""
    c = rc.parallel_case()
    rows = np.array([v for _, v in c["rows"]], dtype=np.float32)
    a = impl.search(rows, c["query"], c["k"], threads=1)
    b = impl.search(rows, c["query"], c["k"], threads=8)
    assert a[0] == b[0] == list(range(63, 53, -1))
    assert np.array_equal(bits(a[1]), bits(b[1]))


def test_scan_matches_numpy_mirror_and_is_chunking_invariant(fo):
    """merge_partial_heaps_preserves_top_k (search.rs:3229) / collect-all == heap prefix
    (search.rs:2646): the 1024-row chunked heap path equals a plain sort of all keys."""
    slab, _ = fo.synth_rows(1, 1, 0, 5000, 128)
    q = fo.clustered_query(5, 128)
    tomb = np.zeros(5000, dtype=bool)
    tomb[::7] = True
    for order in (0, 1):
        scores = no.dot_rows(slab, q, order)
        for k in (1, 10, 100, 4999, 5000, 6000):
            for tb in (None, tomb):
                r, s = fo.search_top_k(slab, q, k, None if tb is None else fo.pack_bitmap(tb), 4, order)
                r2, s2 = no.top_k(scores, k, None if tb is None else ~tb)
                assert np.array_equal(r, r2) and np.array_equal(bits(s), bits(s2)), (order, k)


def test_score_key_maps_nan_to_neg_infinity(fo):
    """search.rs:3470 score_key_maps_nan_to_neg_infinity + heap_entry ordering (search.rs:3094-3110):
    NaN ties with -inf and is ordered by row."""
    rows = np.array([[np.inf, 0, 0, 0, 0, 0, 0, 0], [1, 0, 0, 0, 0, 0, 0, 0], [-np.inf, 0, 0, 0, 0, 0, 0, 0]],
                    dtype=np.float32)
    slab = no.encode_f16(rows)
    q = np.array([1, 0, 0, 0, 0, 0, 0, 0], dtype=np.float32)
    r, s = fo.search_top_k(slab, q, 3)
    assert list(r) == [0, 1, 2] and s[0] == np.inf and s[2] == -np.inf
    qn = np.array([0, np.nan, 0, 0, 0, 0, 0, 0], dtype=np.float32)  # every score NaN -> row order
    r, s = fo.search_top_k(slab, qn, 3)
    assert list(r) == [0, 1, 2] and np.isnan(s).all()


# ── RRF / blend ─────────────────────────────────────────────────────────────────────────────
@pytest.mark.parametrize("case", rc.RRF_CASES, ids=[c["name"] for c in rc.RRF_CASES])
def test_rrf_known_answers(impl, case):
    check_rrf_case(impl, case)


@pytest.mark.parametrize("case", rc.BLEND_CASES, ids=[c["name"] for c in rc.BLEND_CASES])
def test_blend_known_answers(impl, case):
    check_blend_case(impl, case)


def test_aligned_blend_is_bit_identical_to_materialized(impl):
    check_blend_aligned(impl, rc.BLEND_ALIGNED)


# ── misc ────────────────────────────────────────────────────────────────────────────────────
def test_fnv1a_known_values(fo):
    """FNV-1a-64 published test vectors (lib.rs:6120-6127 is the textbook function)."""
    assert fo.fnv1a64(b"") == 0xCBF29CE484222325
    assert fo.fnv1a64(b"a") == 0xAF63DC4C8601EC8C
    assert fo.fnv1a64(b"foobar") == 0x85944171F73967E8
    from frankensearch_b200.types import fnv1a_hash
    for s in (b"", b"a", b"foobar", b"doc-000123"):
        assert fnv1a_hash(s) == fo.fnv1a64(s)


def test_potion_matches_numpy(fo):
    """model2vec_embedder.rs:312-335, :435-451: token-order sum, mean, sequential norm."""
    rng = np.random.default_rng(3)
    table = rng.standard_normal((500, 256)).astype(np.float32)
    ids = np.array([3, 499, 7, 1000, 3, 42], dtype=np.uint32)  # 1000 is out of vocabulary
    got = fo.potion_embed(table, ids)
    s = np.zeros(256, dtype=np.float32)
    cnt = 0
    for t in ids:
        if t < 500:
            s = s + table[t]
            cnt += 1
    s = s * np.float32(np.float32(1.0) / np.float32(cnt))
    nsq = np.float32(0)
    for v in s:
        nsq = np.float32(nsq + np.float32(v * v))
    want = s * np.float32(np.float32(1.0) / np.sqrt(nsq))
    assert np.array_equal(bits(got), bits(want))
    assert not fo.potion_embed(table, np.array([1000, 2000], dtype=np.uint32)).any()
    assert not fo.potion_embed(table, np.zeros(0, dtype=np.uint32)).any()


def test_synth_generator_matches_numpy(fo):
    """fsvi_int8_two_pass.rs:199-231 restated twice (C++ and NumPy scalar loops)."""
    def raw(seed, dim):
        s = seed | 1
        out = np.empty(dim, dtype=np.float32)
        for d in range(dim):
            s ^= (s << 13) & 0xFFFFFFFFFFFFFFFF
            s ^= s >> 7
            s ^= (s << 17) & 0xFFFFFFFFFFFFFFFF
            out[d] = np.float32(np.float32(s >> 40) / np.float32(8388608.0) - np.float32(1.0))
        return out

    def norm(v):
        acc = np.float32(0)
        for x in v:
            acc = np.float32(acc + np.float32(x * x))
        n = np.sqrt(acc)
        return (v / n).astype(np.float32) if n > 1e-12 else v

    dim = 64
    cents = [norm(raw(0xC0000000 + c, dim)) for c in range(4)]
    f16, f32 = fo.synth_rows(1, 1, 10, 6, dim, n_centroids=4, noise=0.30, want_f32=True)
    for r in range(6):
        i = 10 + r
        want = norm((cents[i % 4] + np.float32(0.30) * raw(1 + i, dim)).astype(np.float32))
        assert np.array_equal(bits(f32[r]), bits(want))
        assert np.array_equal(f16[r], no.encode_f16(want))
    u16, u32 = fo.synth_rows(0, 100, 0, 3, dim, want_f32=True)
    assert np.array_equal(bits(u32[2]), bits(norm(raw(102, dim))))


# ── resident WAL rows ───────────────────────────────────────────────────────────────────────
def test_f32_dot_matches_numpy_mirror(fo):
    """simd.rs:2512 avx2_f32slicedot_matches_generic (same seed, same dims): the C++ statement of
    dot_product_f32_f32 against the NumPy one, every reduce order."""
    nxt = _xorshift_stream(rc.DOT_F32_XORSHIFT["seed"])
    for dim in rc.DOT_F32_XORSHIFT["dims"]:
        a = np.array([nxt() for _ in range(dim)], dtype=np.float32)
        b = np.array([nxt() for _ in range(dim)], dtype=np.float32)
        for order in range(5):
            assert bits(fo.dot_f32_f32(a, b, order)) == bits(no.dot_f32_f32(a, b, order)), (dim, order)
        # the value is a dot product (f64 check with a loose bound: orders differ only in rounding)
        assert abs(float(fo.dot_f32_f32(a, b)) - float(np.dot(a.astype(np.float64), b.astype(np.float64)))) < 1e-4


@pytest.mark.parametrize("scenario", rc.WAL_SCENARIOS, ids=[s["name"] for s in rc.WAL_SCENARIOS])
def test_wal_known_answers(scenario):
    from wal_model import OracleWalIndex, run_scenario

    run_scenario(lambda ids, vecs, dim: OracleWalIndex(ids, vecs, dim), scenario)


def test_wal_full_recall_collect_all_matches_heap_prefix():
    """search.rs:2688 full_recall_collect_all_matches_heap_prefix_with_wal."""
    from wal_model import OracleWalIndex

    c = rc.wal_full_recall_case()
    ix = OracleWalIndex([d for d, _ in c["rows"]], [v for _, v in c["rows"]], 4)
    ix.append_batch(c["wal"])
    total = len(c["rows"]) + len(c["wal"])
    full = ix.search_top_k(c["query"], total + 10)
    heap = ix.search_top_k(c["query"], total - 5)
    assert len(full) == total and full[0][2] == "wal-top" and len(heap) == total - 5
    for h, f in zip(heap, full):
        assert h[2] == f[2] and h[0] == f[0] and bits(h[1]) == bits(f[1])
    # WAL rows are numbered after the main rows (resolve_wal_hit, search.rs:1583-1597)
    assert full[0][0] == 48 and [h[0] for h in full if h[2] == "wal-mid"] == [49]
    # equal scores: the main row ranks before the WAL row (tagged index, wal.rs:557-569)
    ix.append("wal-tie", [24.0, 0.0, 0.0, 0.0])
    ids = [h[2] for h in ix.search_top_k(c["query"], total + 10)]
    assert ids.index("doc-024") + 1 == ids.index("wal-tie")


def test_wal_nonfinite_scores_are_skipped(fo):
    """scan_wal (search.rs:1466-1470): a WAL row whose score is not finite never enters the heap."""
    slab = fo.encode_f16(np.array([[1.0, 0.0, 0.0, 0.0]], dtype=np.float32))
    wal = np.array([[3.0e38, 3.0e38, 0.0, 0.0], [0.5, 0.0, 0.0, 0.0]], dtype=np.float32)
    rows, scores = fo.search_top_k_wal(slab, wal, np.array([2.0, 2.0, 0.0, 0.0], dtype=np.float32), 10)
    assert list(rows) == [0, 2] and np.isfinite(scores).all()


def test_reduce_order_probe_separates_all_five_orders(fo):
    """INTEGRATION.md "which reduce_add order does your build use": the probe vector gives five distinct
    bit patterns, one per candidate order, in the C++ oracle and in the NumPy restatement."""
    from oracle import np_oracle as no

    ones = fo.encode_f16(np.ones(8, dtype=np.float32))
    q = np.array(rc.REDUCE_PROBE, dtype=np.float32)
    got = {o: int(np.float32(fo.dot_f16_f32(ones, q, o, True)).view(np.uint32)) for o in range(5)}
    assert got == rc.REDUCE_PROBE_BITS
    assert len(set(got.values())) == 5
    for o in range(5):
        assert int(np.float32(no.dot_f16_f32(ones, q, o)).view(np.uint32)) == rc.REDUCE_PROBE_BITS[o]
