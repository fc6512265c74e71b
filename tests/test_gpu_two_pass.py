"""fsgpu_search_top_k_two_pass against the oracle restatement of VectorIndex::search_top_k_int8_two_pass /
search_top_k_4bit_two_pass (crates/frankensearch-index/src/search.rs:514-650, :876-946): bit-identical rows and
scores for every multiplier (lossy ones included), byte-identical code slabs, the reference's fall-backs."""
import numpy as np
import pytest

from oracle import fs_oracle as fo
from test_two_pass_oracle import keep_all_fixture

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fs(cuda_ok):
    assert cuda_ok, "no usable CUDA device: the product has no CPU fallback"
    import frankensearch_b200 as fs

    return fs


def run(ix, q, k, mult, bits):
    hits = ix.search_top_k_int8_two_pass(q, k, mult) if bits == 8 else ix.search_top_k_4bit_two_pass(q, k, mult)
    return (np.array([h.index for h in hits], dtype=np.uint64), np.array([h.score for h in hits], dtype=np.float32))


def same(got, want):
    assert np.array_equal(got[0], want[0]), (got[0][:12], want[0][:12])
    assert np.array_equal(got[1].view(np.uint32), want[1].view(np.uint32))


@pytest.mark.parametrize("bits,dim", [(8, 8), (4, 70)])
def test_reference_keep_all_fixtures(fs, bits, dim):
    slab, queries = keep_all_fixture(dim)
    ix = fs.GpuVectorIndex.from_f16_bits(None, slab)
    for q in queries:
        same(run(ix, q, 10, 50, bits), fo.search_top_k(slab, q, 10))
    ix.close()


@pytest.mark.parametrize("dim", [384, 128, 96, 70, 33])
def test_two_pass_matches_oracle_for_every_multiplier(fs, dim):
    rng = np.random.default_rng(dim)
    n = 20_000
    x = rng.standard_normal((n, dim)).astype(np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    x[rng.integers(0, n, 5)] *= 3.0  # a few long rows set the corpus-wide scale
    slab = fo.encode_f16(x)
    tomb = rng.random(n) < 0.05
    ix = fs.GpuVectorIndex.from_f16_bits(None, slab, tombstones=tomb)
    assert np.array_equal(ix.two_pass_codes(8), fo.quantize_slab_i8(slab))
    assert np.array_equal(ix.two_pass_codes(4), fo.pack_slab_4bit(slab))
    lost = 0
    for qi in range(6):
        q = rng.standard_normal(dim).astype(np.float32)
        exact = fo.search_top_k(slab, q, 10, tombstones=tomb)[0]
        for bits in (8, 4):
            for k, mult in ((10, 1), (10, 3), (10, 0), (1, 5), (100, 2), (10, 5000)):
                want = fo.search_two_pass(slab, q, k, mult, bits, tombstones=tomb)
                same(run(ix, q, k, mult, bits), want)
            lost += 10 - len(set(fo.search_two_pass(slab, q, 10, 1, bits, tombstones=tomb)[0]) & set(exact))
    assert lost > 0  # the lossy regime was exercised: identical to the reference there too
    ix.close()


def test_two_pass_edges_and_fallbacks(fs):
    rng = np.random.default_rng(5)
    x = (rng.standard_normal((500, 64)) * 0.1).astype(np.float32)
    slab = fo.encode_f16(x)
    ix = fs.GpuVectorIndex.from_f16_bits([f"d{i}" for i in range(500)], slab)
    q = rng.standard_normal(64).astype(np.float32)
    assert ix.search_top_k_int8_two_pass(q, 0, 3) == []
    with pytest.raises(fs.SearchError):
        ix.search_top_k_4bit_two_pass(q[:32], 5, 3)
    same(run(ix, q, 700, 3, 8), fo.search_two_pass(slab, q, 700, 3, 8))  # k > rows: every row, best first
    zq = np.zeros(64, dtype=np.float32)  # a zero query: every integer score is 0, candidates = lowest rows
    same(run(ix, zq, 5, 2, 4), fo.search_two_pass(slab, zq, 5, 2, 4))
    same(run(ix, zq, 5, 2, 8), fo.search_two_pass(slab, zq, 5, 2, 8))
    # resident WAL rows: the reference falls back to the exact search (search.rs:578-586)
    ix.append("new-doc", (q / np.linalg.norm(q)).astype(np.float32))
    hits = ix.search_top_k_int8_two_pass(q, 3, 1)
    assert [h.doc_id for h in hits] == [h.doc_id for h in ix.search_top_k(q, 3)]
    assert hits[0].doc_id == "new-doc"
    ix.close()
    zero = fs.GpuVectorIndex.from_f16_bits(None, np.zeros((40, 16), dtype=np.uint16))
    same(run(zero, q[:16], 4, 2, 8), fo.search_two_pass(np.zeros((40, 16), dtype=np.uint16), q[:16], 4, 2, 8))
    same(run(zero, q[:16], 4, 2, 4), fo.search_two_pass(np.zeros((40, 16), dtype=np.uint16), q[:16], 4, 2, 4))
    zero.close()


def test_two_pass_one_million_rows(fs):
    """1 M x 384 clustered corpus (the reference bench generator): multiplier 3 (search_fast's default) and 5."""
    n, dim = 1_000_000, 384
    slab = fo.synth_rows(1, 1, 0, n, dim)[0]
    ix = fs.GpuVectorIndex.from_f16_bits(None, slab)
    for qi in range(3):
        q = fo.clustered_query(qi, dim)
        for bits, mult in ((8, 3), (4, 5)):
            same(run(ix, q, 10, mult, bits), fo.search_two_pass(slab, q, 10, mult, bits))
    ix.close()


def test_two_tier_search_fast_reference_default(fs):
    """TwoTierIndex::search_fast without params = search_top_k_int8_two_pass(query, k, 3) (two_tier.rs:1323-1342):
    reproduced literally with reference_two_pass=True; the default mirror returns the exact top-k."""
    rng = np.random.default_rng(21)
    x = rng.standard_normal((5000, 64)).astype(np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    slab = fo.encode_f16(x)
    fast = fs.GpuVectorIndex.from_f16_bits(None, slab)
    q = rng.standard_normal(64).astype(np.float32)
    lit = fs.GpuTwoTierIndex(fast, reference_two_pass=True).search_fast(q, 10)
    want = fo.search_two_pass(slab, q, 10, 3, 8)
    assert [h.index for h in lit] == [int(r) for r in want[0]]
    exact = fs.GpuTwoTierIndex(fast).search_fast(q, 10)
    assert [h.index for h in exact] == [int(r) for r in fo.search_top_k(slab, q, 10)[0]]
    fast.close()
