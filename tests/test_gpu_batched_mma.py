"""GPU parity tests of the batched tensor-core scan (mma_scan_kernels.cuh) through the C ABI.

The batched path must return EXACTLY what the per-query path and the CPU oracle return: same rows,
same order, bit-identical f32 scores (the tensor-core scores only nominate candidates; winners are
re-scored with the reference accumulation tree).  North-star tolerance for scores is 1e-3 relative;
these tests demand 0 ULP.
"""
import os

import numpy as np
import pytest

from adapters import OracleImpl

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fs(cuda_ok):
    assert cuda_ok, "no usable CUDA device: the product has no CPU fallback"
    import frankensearch_b200 as fs

    return fs


@pytest.fixture(scope="module")
def cpu():
    return OracleImpl()


def bits(x):
    return np.asarray(x, dtype=np.float32).view(np.uint32)


def assert_batch_matches_oracle(cpu, slab, queries, k, got, tombstones=None, which=None, ctx=""):
    rows, scores, counts = got
    for b in (range(len(queries)) if which is None else which):
        want_rows, want_scores = cpu.search_bits(slab, queries[b], k, tombstones=tombstones)
        n = int(counts[b])
        assert rows[b, :n].tolist() == want_rows, f"rows differ {ctx} b={b}"
        w = np.asarray(want_scores, dtype=np.float32)
        nan = np.isnan(w)
        assert np.array_equal(np.isnan(scores[b, :n]), nan), f"NaN pattern {ctx} b={b}"
        assert np.array_equal(bits(scores[b, :n])[~nan], bits(w)[~nan]), f"score bits differ {ctx} b={b}"


def search_with_profile(ix, queries, k):
    ix.profile_read(reset=True)
    out = ix.search_top_k_batch(queries, k)
    return out, ix.profile_read(reset=True)


@pytest.mark.parametrize("dim", [64, 128, 256, 384, 512])
def test_batched_parity_dims(fs, cpu, fo, dim):
    rng = np.random.default_rng(dim)
    n = 5003
    rows = rng.normal(size=(n, dim)).astype(np.float32)
    rows /= np.linalg.norm(rows, axis=1, keepdims=True)
    slab = fo.encode_f16(rows)
    qs = rng.normal(size=(37, dim)).astype(np.float32)
    qs /= np.linalg.norm(qs, axis=1, keepdims=True)
    ix = fs.GpuVectorIndex.from_f16_bits(None, slab)
    for k in (1, 10, 100):
        got, prof = search_with_profile(ix, qs, k)
        assert prof["mma_launches"] >= 1, "the batched tensor-core path did not run"
        assert_batch_matches_oracle(cpu, slab, qs, k, got, ctx=f"dim={dim} k={k}")
    ix.close()


@pytest.mark.parametrize("batch", [8, 9, 127, 128, 129, 300])
def test_batched_parity_batch_sizes(fs, cpu, fo, batch):
    """Query-block boundaries: one partial block, exactly one, one + 1, several."""
    slab, _ = fo.synth_rows(1, 5, 0, 20011, 384)
    qs = np.stack([fo.clustered_query(1000 + i, 384) for i in range(batch)])
    ix = fs.GpuVectorIndex.from_f16_bits(None, slab)
    got, prof = search_with_profile(ix, qs, 10)
    assert prof["mma_launches"] >= 1
    which = sorted(set(range(0, batch, max(1, batch // 24))) | {batch - 1})
    assert_batch_matches_oracle(cpu, slab, qs, 10, got, which=which, ctx=f"batch={batch}")
    ix.close()


@pytest.mark.parametrize("n", [1, 5, 127, 128, 129, 1000, 4097])
def test_batched_parity_row_counts(fs, cpu, fo, n):
    """Fewer rows than k, partial tiles, fewer tiles than CTAs."""
    slab, _ = fo.synth_rows(0, 9, 0, n, 128)
    rng = np.random.default_rng(n)
    qs = rng.uniform(-1, 1, (16, 128)).astype(np.float32)
    ix = fs.GpuVectorIndex.from_f16_bits(None, slab)
    for k in (1, 10, 200):
        got, prof = search_with_profile(ix, qs, k)
        assert prof["mma_launches"] >= 1
        assert got[2].tolist() == [min(k, n)] * 16
        assert_batch_matches_oracle(cpu, slab, qs, k, got, ctx=f"n={n} k={k}")
    ix.close()


def test_batched_equals_per_query_path(fs, fo):
    """Same index, same queries: tensor-core path vs the CUDA-core path, bit for bit."""
    slab, _ = fo.synth_rows(1, 3, 0, 200_000, 384)
    qs = np.stack([fo.clustered_query(i, 384) for i in range(256)])
    ix = fs.GpuVectorIndex.from_f16_bits(None, slab)
    for k in (10, 100, 256):
        (r1, s1, c1), prof = search_with_profile(ix, qs, k)
        assert prof["mma_launches"] >= 1
        os.environ["FSGPU_MMA_MIN_BATCH"] = "0"
        try:
            (r2, s2, c2), prof2 = search_with_profile(ix, qs, k)
        finally:
            del os.environ["FSGPU_MMA_MIN_BATCH"]
        assert prof2["mma_launches"] == 0
        assert np.array_equal(c1, c2) and np.array_equal(r1, r2) and np.array_equal(bits(s1), bits(s2)), f"k={k}"
    ix.close()


def test_batched_tombstones_and_row_base(fs, cpu, fo):
    slab, _ = fo.synth_rows(1, 11, 0, 30000, 256)
    rng = np.random.default_rng(1)
    tomb = rng.random(30000) < 0.3
    qs = np.stack([fo.clustered_query(i, 256) for i in range(40)])
    ix = fs.GpuVectorIndex.from_f16_bits(None, slab, tombstones=tomb, row_base=123456)
    (rows, scores, counts), prof = search_with_profile(ix, qs, 50)
    assert prof["mma_launches"] >= 1
    assert_batch_matches_oracle(cpu, slab, qs, 50, (rows - 123456, scores, counts), tombstones=tomb)
    # the strongest rows are deleted afterwards: results must change accordingly
    tomb2 = tomb.copy()
    tomb2[(rows[:, 0] - 123456).astype(np.int64)] = True
    ix.set_tombstones(tomb2)
    got = ix.search_top_k_batch(qs, 50)
    assert_batch_matches_oracle(cpu, slab, qs, 50, (got[0] - 123456, got[1], got[2]), tombstones=tomb2)
    ix.close()


def test_batched_tie_band_overflow_is_redone_exactly(fs, cpu, fo):
    """49 000 identical rows: every CTA sees several full tiles of exact ties, the tie band cannot
    fit in a candidate list, so those queries are re-run on the exact kernel; ties must still
    resolve to the lowest rows (search.rs:2741).  A 3 000-row tie band, by contrast, fits and is
    resolved by the refine kernel alone."""
    rng = np.random.default_rng(5)
    base = rng.normal(size=(60000, 64)).astype(np.float32)
    base /= np.linalg.norm(base, axis=1, keepdims=True)
    base[1000:50000] = base[7]
    slab = fo.encode_f16(base)
    qs = rng.normal(size=(12, 64)).astype(np.float32)
    qs[0] = base[7]  # the duplicated row is this query's best hit
    qs[5] = base[7] * 0.5
    ix = fs.GpuVectorIndex.from_f16_bits(None, slab)
    got, prof = search_with_profile(ix, qs, 20)
    assert prof["mma_launches"] >= 1 and prof["redo_queries"] >= 2
    assert got[0][0, :20].tolist() == [7] + list(range(1000, 1019))
    assert_batch_matches_oracle(cpu, slab, qs, 20, got)
    ix.close()
    base[4000:50000] = rng.normal(size=(46000, 64)).astype(np.float32) * 0.1
    slab = fo.encode_f16(base)
    ix = fs.GpuVectorIndex.from_f16_bits(None, slab)
    got, prof = search_with_profile(ix, qs, 20)
    assert prof["mma_launches"] >= 1
    assert got[0][0, :20].tolist() == [7] + list(range(1000, 1019))
    assert_batch_matches_oracle(cpu, slab, qs, 20, got)
    ix.close()


def test_batched_unsafe_queries_are_redone_exactly(fs, cpu, fo):
    """NaN / inf / f16-overflowing / all-zero / subnormal query components inside a batch."""
    slab, _ = fo.synth_rows(0, 2, 0, 6000, 128)
    rng = np.random.default_rng(9)
    qs = rng.uniform(-1, 1, (16, 128)).astype(np.float32)
    qs[1, 3] = np.nan
    qs[2, 0] = np.inf
    qs[3, :] *= np.float32(1e6)      # beyond f16 range
    qs[4, :] = 0.0                   # all-zero: every score +0.0 or -0.0 -> order by total_cmp then row
    qs[5, :] *= np.float32(1e-30)    # products underflow to subnormals
    qs[6, :] *= np.float32(300.0)    # large but representable
    qs[7, 1::2] = 0.0
    ix = fs.GpuVectorIndex.from_f16_bits(None, slab)
    got, prof = search_with_profile(ix, qs, 10)
    # NaN and inf always go to the exact kernel; the f16 form also sends the 1e6-scaled query there
    # (f16 overflow), the int8 form quantises it with its own scale
    assert prof["mma_launches"] >= 1 and prof["redo_queries"] >= 2
    assert_batch_matches_oracle(cpu, slab, qs, 10, got)
    ix.close()


def test_batched_unnormalised_rows_and_queries(fs, cpu, fo):
    """The error bound scales with ||row|| and ||q||: nothing assumes unit vectors."""
    rng = np.random.default_rng(13)
    rows = (rng.normal(size=(7000, 192)) * rng.uniform(0.01, 30.0, (7000, 1))).astype(np.float32)
    slab = fo.encode_f16(rows)
    qs = (rng.normal(size=(24, 192)) * rng.uniform(0.001, 50.0, (24, 1))).astype(np.float32)
    ix = fs.GpuVectorIndex.from_f16_bits(None, slab)
    got, prof = search_with_profile(ix, qs, 25)
    assert prof["mma_launches"] >= 1
    assert_batch_matches_oracle(cpu, slab, qs, 25, got)
    ix.close()


def test_nonfinite_slab_never_takes_the_batched_path(fs, cpu, fo):
    rng = np.random.default_rng(17)
    slab = fo.encode_f16(rng.normal(size=(3000, 128)).astype(np.float32))
    slab[17, 5] = 0x7C00  # +inf
    slab[99, 0] = 0x7E00  # NaN
    qs = rng.normal(size=(9, 128)).astype(np.float32)
    ix = fs.GpuVectorIndex.from_f16_bits(None, slab)
    got, prof = search_with_profile(ix, qs, 10)
    assert prof["mma_launches"] == 0
    assert_batch_matches_oracle(cpu, slab, qs, 10, got)
    ix.close()


def test_batched_1m_x_384_batch_1024(fs, cpu, fo):
    """BASELINE config 2/3 shape at 1 M rows: the full 1024-query batch in one pass; every query
    checked against the per-query GPU path, a sample against the CPU oracle."""
    import torch

    n, dim, k, batch = 1_000_000, 384, 10, 1024
    slab_gpu = torch.empty((n, dim), dtype=torch.int16, device="cuda")
    fs._ffi.check(fs._ffi.lib().fsgpu_synth_rows_device(0, 1, 1, 0, n, dim, 64, 0.30, slab_gpu.data_ptr(), None))
    ix = fs.GpuVectorIndex.from_device_tensor(slab_gpu)
    qs = np.stack([fo.clustered_query(i, dim) for i in range(batch)])
    (r1, s1, c1), prof = search_with_profile(ix, qs, k)
    assert prof["mma_launches"] == 1 and prof["scan_launches"] == 1 + prof["redo_queries"]
    os.environ["FSGPU_MMA_MIN_BATCH"] = "0"
    try:
        r2, s2, c2 = ix.search_top_k_batch(qs, k)
    finally:
        del os.environ["FSGPU_MMA_MIN_BATCH"]
    assert np.array_equal(c1, c2) and np.array_equal(r1, r2) and np.array_equal(bits(s1), bits(s2))
    slab_cpu = slab_gpu.cpu().numpy().view(np.uint16)
    assert_batch_matches_oracle(cpu, slab_cpu, qs, k, (r1, s1, c1), which=[0, 1, 511, 1023])
    ix.close()


def test_filtered_search_matches_oracle_with_exclusions(fs, cpu, fo):
    """SearchFilter (search.rs:192-206, :1329-1447) as an allow-bitmap: filtered-out rows behave
    like tombstones on both the per-query and the batched path, and compose with tombstones."""
    slab, _ = fo.synth_rows(1, 31, 0, 25000, 128)
    ids = [f"doc-{i:06}" for i in range(25000)]
    rng = np.random.default_rng(3)
    tomb = rng.random(25000) < 0.1
    allow = rng.random(25000) < 0.4
    ix = fs.GpuVectorIndex.from_f16_bits(ids, slab, tombstones=tomb)
    qs = np.stack([fo.clustered_query(i, 128) for i in range(20)])
    excluded = tomb | ~allow
    for batch in (1, 20):  # per-query kernel, batched tensor-core kernel
        got = ix.search_top_k_batch(qs[:batch], 30, filter=allow)
        assert_batch_matches_oracle(cpu, slab, qs[:batch], 30, got, tombstones=excluded, ctx=f"batch={batch}")
    hits = ix.search_top_k(qs[0], 10, filter=lambda d: int(d[4:]) % 3 == 0)
    want_rows, _ = cpu.search_bits(slab, qs[0], 10, tombstones=tomb | (np.arange(25000) % 3 != 0))
    assert [h.index for h in hits] == want_rows and all(int(h.doc_id[4:]) % 3 == 0 for h in hits)
    # the filter is per call: the next unfiltered search sees every live row again
    got = ix.search_top_k_batch(qs[:3], 30)
    assert_batch_matches_oracle(cpu, slab, qs[:3], 30, got, tombstones=tomb)
    ix.close()


def test_batched_super_batches_beyond_one_launch(fs, cpu, fo):
    """More queries than one launch holds (148 SMs x 128): the call loops over super-batches."""
    slab, _ = fo.synth_rows(0, 41, 0, 700, 64)
    rng = np.random.default_rng(8)
    qs = rng.uniform(-1, 1, (19500, 64)).astype(np.float32)
    ix = fs.GpuVectorIndex.from_f16_bits(None, slab)
    got, prof = search_with_profile(ix, qs, 5)
    assert prof["mma_launches"] >= 2
    assert_batch_matches_oracle(cpu, slab, qs, 5, got, which=[0, 127, 128, 9471, 18943, 18944, 19499])
    ix.close()


def test_batched_large_k_up_to_1024(fs, cpu, fo):
    """k = 1000 (BASELINE configs[4] "rerank top-1000") stays on the tensor-core path: long
    candidate lists take the kernels' unstaged selection path."""
    slab, _ = fo.synth_rows(1, 51, 0, 60000, 128)
    qs = np.stack([fo.clustered_query(i, 128) for i in range(12)])
    ix = fs.GpuVectorIndex.from_f16_bits(None, slab)
    for k in (600, 1000, 1024):
        got, prof = search_with_profile(ix, qs, k)
        assert prof["mma_launches"] >= 1
        assert_batch_matches_oracle(cpu, slab, qs, k, got, which=[0, 5, 11], ctx=f"k={k}")
    got, prof = search_with_profile(ix, qs, 1025)  # beyond the fused limit: score-all + sort arm
    assert prof["mma_launches"] == 0
    assert_batch_matches_oracle(cpu, slab, qs, 1025, got, which=[3])
    ix.close()


def test_small_batches_take_the_tensor_core_path_from_three_queries(fs, cpu, fo):
    slab, _ = fo.synth_rows(1, 61, 0, 30000, 384)
    qs = np.stack([fo.clustered_query(i, 384) for i in range(7)])
    ix = fs.GpuVectorIndex.from_f16_bits(None, slab)
    for b, want_mma in ((1, False), (2, False), (3, True), (7, True)):
        got, prof = search_with_profile(ix, qs[:b], 10)
        assert (prof["mma_launches"] >= 1) == want_mma
        assert_batch_matches_oracle(cpu, slab, qs[:b], 10, got, ctx=f"b={b}")
    ix.close()


def test_randomised_soak_batched_equals_per_query_path(fs):
    """A short run of tools/soak_batched.py (210 cases on the B200 box: profiles/r01_soak_batched_210_cases.txt):
    adversarial corpora (ascending-by-score order, duplicates, extreme norms, exact ties), tombstones
    and filters; rows, order and f32 score bits of the batched path must equal the per-query path."""
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "soak_batched.py"), "35", "7"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "soak ok" in r.stdout


def test_batched_f16_subnormal_rows(fs, cpu, fo):
    """Rows made only of f16 subnormals (|x| < 2^-14): the tensor cores must treat them exactly
    (no flush-to-zero) for the error bound to hold without help; either way the answer is exact."""
    rng = np.random.default_rng(23)
    slab = rng.integers(1, 1024, (9000, 128)).astype(np.uint16)           # subnormal magnitudes
    slab |= (rng.integers(0, 2, (9000, 128)).astype(np.uint16) << 15)     # random signs
    qs = rng.normal(size=(20, 128)).astype(np.float32)
    ix = fs.GpuVectorIndex.from_f16_bits(None, slab)
    got, prof = search_with_profile(ix, qs, 10)
    assert prof["mma_launches"] >= 1
    assert_batch_matches_oracle(cpu, slab, qs, 10, got)
    assert prof["redo_queries"] == 0, "subnormal f16 inputs were not handled exactly by the MMA path"
    ix.close()


def test_async_device_calls_redo_on_the_device(fs, cpu, fo, monkeypatch):
    """The stream-asynchronous `_device` entry point never blocks the host: queries the batched path
    cannot cover (NaN / inf / f16-overflow, and an overflowing tie band) are re-run by the exact kernels
    on the caller's stream.  Results equal the oracle, `fsgpu_index_last_status` reports how many were
    redone, and a call on ANOTHER stream right behind it is ordered after it (shared workspaces)."""
    import ctypes as C

    import torch

    rng = np.random.default_rng(5)
    base = rng.normal(size=(60000, 64)).astype(np.float32)
    base /= np.linalg.norm(base, axis=1, keepdims=True)
    base[1000:50000] = base[7]
    slab = fo.encode_f16(base)
    qs = rng.normal(size=(40, 64)).astype(np.float32)
    qs[0] = base[7]
    qs[5] = base[7] * 0.5
    qs[9, 3] = np.nan
    qs[17, 0] = np.inf
    qs[33, :] *= np.float32(1e6)
    ix = fs.GpuVectorIndex.from_f16_bits(None, slab)
    d_q = torch.from_numpy(qs).cuda()
    side = torch.cuda.Stream()
    for k in (20, 300):
        keys, hits, counts = ix.search_top_k_device(d_q, k)  # torch's current stream, asynchronous
        with torch.cuda.stream(side):  # a second call on another stream, enqueued at once
            keys2, hits2, counts2 = ix.search_top_k_device(d_q[:7].contiguous(), k)
        torch.cuda.synchronize()
        flags = (C.c_uint32 * 4)()
        fs._ffi.check(ix._L.fsgpu_index_last_status(ix._h, flags))
        h = hits.cpu().numpy()
        got = (h[..., 0].view(np.uint32), h[..., 1].view(np.float32), counts.cpu().numpy())
        assert_batch_matches_oracle(cpu, slab, qs, k, got, ctx=f"async k={k}")
        h2 = hits2.cpu().numpy()
        got2 = (h2[..., 0].view(np.uint32), h2[..., 1].view(np.float32), counts2.cpu().numpy())
        assert_batch_matches_oracle(cpu, slab, qs[:7], k, got2, ctx=f"async side stream k={k}")
        assert flags[0] == 0
    p = ix.profile_read(reset=True)
    assert p["mma_launches"] >= 2 and p["redo_queries"] >= 4, p
    ix.close()


def test_int8_overflow_falls_back_to_f16_form_sync_and_async(fs, cpu, fo, monkeypatch):
    """A corpus whose score spread is far narrower than the int8 bound (every row a small perturbation
    of one vector): the int8 lists overflow for every query.  A synchronous caller re-runs the batch in
    the f16 form; an asynchronous one gets the device redo, then the following calls avoid the int8 form."""
    import torch

    monkeypatch.setenv("FSGPU_I8_MIN_ROWS", "0")
    rng = np.random.default_rng(3)
    n, dim = 40000, 128
    centre = rng.normal(size=dim).astype(np.float32)
    centre /= np.linalg.norm(centre)
    rows = centre[None, :] + 1e-3 * rng.normal(size=(n, dim)).astype(np.float32)
    slab = fo.encode_f16(rows)
    qs = (centre[None, :] + 0.05 * rng.normal(size=(8, dim))).astype(np.float32)
    ix = fs.GpuVectorIndex.from_f16_bits(None, slab)
    assert ix._L.fsgpu_index_int8_ready(ix._h) == 1
    got = ix.search_top_k_batch(qs, 10)  # synchronous host API
    assert_batch_matches_oracle(cpu, slab, qs, 10, got, ctx="sync")
    d_q = torch.from_numpy(qs).cuda()
    for rep in range(3):
        keys, hits, counts = ix.search_top_k_device(d_q, 10)
        torch.cuda.synchronize()
        h = hits.cpu().numpy()
        assert_batch_matches_oracle(cpu, slab, qs, 10, (h[..., 0].view(np.uint32), h[..., 1].view(np.float32),
                                                        counts.cpu().numpy()), ctx=f"async rep={rep}")
    ix.close()
