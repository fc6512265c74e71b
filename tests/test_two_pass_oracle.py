"""The oracle's restatement of the reference's quantised two-pass searches (oracle/fs_oracle.cpp
fso_search_two_pass, fso_quantize_slab_i8, fso_pack_slab_4bit) pinned by the reference's own fixtures:
`int8_two_pass_keep_all_matches_exact` (crates/frankensearch-index/src/search.rs:1815-1860),
`four_bit_two_pass_keep_all_matches_exact` (:2010-2052) and the nibble algebra of `dot_packed_4bit_matches_scalar`
(simd.rs:2799-2832)."""
import numpy as np
import pytest

from oracle import fs_oracle as fo
from oracle import np_oracle as npo


def keep_all_fixture(dim, count=300):
    """vectors / queries of the two keep-all tests (search.rs:1823-1834, :1843-1846)."""
    i = np.arange(count, dtype=np.uint64)[:, None]
    j = np.arange(dim, dtype=np.uint64)[None, :]
    s = (i * np.uint64(2_654_435_761)) ^ (j * np.uint64(40_503))
    s ^= s >> np.uint64(13)
    vectors = ((s & np.uint64(0xFFFF)).astype(np.float32) / np.float32(65_535.0)) - np.float32(0.5)
    queries = [np.array([np.float32(((qi * 7 + jj * 3) % 11)) / np.float32(11.0) - np.float32(0.5) for jj in range(dim)],
                        dtype=np.float32) for qi in range(8)]
    return fo.encode_f16(vectors), queries


@pytest.mark.parametrize("bits,dim", [(8, 8), (4, 70)])
def test_keep_all_two_pass_matches_exact_search(bits, dim):
    slab, queries = keep_all_fixture(dim)
    for q in queries:
        rows, scores = fo.search_top_k(slab, q, 10)
        # mult = 50 -> candidate_count clamps to the record count -> pass 1 retains every row
        r2, s2 = fo.search_two_pass(slab, q, 10, 50, bits)
        assert np.array_equal(rows, r2)
        assert np.array_equal(scores.view(np.uint32), s2.view(np.uint32))


def test_int8_slab_quantiser_matches_the_numpy_restatement():
    rng = np.random.default_rng(3)
    slab = fo.encode_f16((rng.standard_normal((257, 48)) * 0.2).astype(np.float32))
    want = npo.quantize_f16_slab_to_i8(slab)
    want = want[0] if isinstance(want, tuple) else want
    assert np.array_equal(fo.quantize_slab_i8(slab), np.asarray(want).reshape(slab.shape))
    assert not fo.quantize_slab_i8(np.zeros((3, 8), dtype=np.uint16)).any()  # max|x| <= 0 -> all zero (simd.rs:1847-1849)


def test_4bit_packing_layout_and_levels():
    """Byte j = dims 2j (low nibble) | 2j + 1 (high nibble); 15 levels -7..7 in two's complement; an odd dimension
    leaves the last high nibble empty (simd.rs:2201-2233, :1892-1896)."""
    vals = np.array([[0.7, -0.7, 0.35, -0.35, 0.0, 0.1, -0.1]], dtype=np.float32)  # max|x| = 0.7 -> scale 10
    packed = fo.pack_slab_4bit(fo.encode_f16(vals))
    assert packed.shape == (1, 4)
    lo = (packed & 0x0F).astype(np.int8)
    hi = (packed >> 4).astype(np.int8)
    sext = lambda n: np.where(n >= 8, n - 16, n)
    got = np.stack([sext(lo), sext(hi)], axis=2).reshape(-1)[:7]
    want = np.clip(np.round(fo.decode_f16(fo.encode_f16(vals)).astype(np.float32) * (np.float32(7.0) / np.float32(fo.decode_f16(fo.encode_f16(vals)).max()))), -7, 7)
    # round-half-away: 0.35 * 10 = 3.5 (after f16 rounding 0.3501 -> 3.501) -> 4
    assert got.tolist() == [7, -7, 4, -4, 0, 1, -1], (got, want)
    assert (packed[0, 3] >> 4) == 0
    assert not fo.pack_slab_4bit(np.zeros((2, 6), dtype=np.uint16)).any()  # max|x| <= 1e-9 -> scale 0


def test_two_pass_candidate_count_rule_and_recall_tradeoff():
    """candidate_count = max(min(k * max(mult, 1), n), min(k, n)) (search.rs:596-599): multiplier 0 behaves as 1,
    k > n returns every live row, and a small multiplier may lose true top-k rows (the reference's trade)."""
    rng = np.random.default_rng(11)
    x = rng.standard_normal((4000, 32)).astype(np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    slab = fo.encode_f16(x)
    q = rng.standard_normal(32).astype(np.float32)
    exact = fo.search_top_k(slab, q, 10)[0]
    for bits in (8, 4):
        a = fo.search_two_pass(slab, q, 10, 0, bits)[0]
        b = fo.search_two_pass(slab, q, 10, 1, bits)[0]
        assert np.array_equal(a, b)
        full = fo.search_two_pass(slab, q, 10, 400, bits)[0]
        assert np.array_equal(full, exact)
    lossy = fo.search_two_pass(slab, q, 10, 1, 4)[0]
    assert len(lossy) == 10 and len(set(lossy) & set(exact)) < 10  # 15 levels, no slack: rows are lost
    tomb = np.zeros(4000, dtype=bool)
    tomb[exact[:3]] = True
    r = fo.search_two_pass(slab[:20], q, 50, 3, 8, tombstones=tomb[:20])[0]
    assert len(r) == 20 - int(tomb[:20].sum())
