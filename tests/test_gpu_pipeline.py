"""Batched, device-resident two-tier pipeline (frankensearch_b200/pipeline.py) against the oracle
flow (tests/flows.py = SyncTwoTierSearcher::search_internal, sync_searcher.rs:616-1009): fast
candidates, quality re-scores, blend and both RRF phases bit for bit; plus the new device entry
points on their own (batched blend in both forms, payload of merged keys)."""
import numpy as np
import pytest

import flows

pytestmark = pytest.mark.gpu


def bits(x):
    return np.asarray(x, dtype=np.float32).view(np.uint32)


def _dev_lexical(torch, fs, lists, dev):
    n_max = max(len(i) for i, _ in lists)
    ids = np.zeros((len(lists), n_max), dtype=np.int64)
    scores = np.zeros((len(lists), n_max), dtype=np.float32)
    counts = np.zeros(len(lists), dtype=np.int32)
    for b, (i, s) in enumerate(lists):
        ids[b, :len(i)] = i.astype(np.int64)
        scores[b, :len(i)] = s
        counts[b] = len(i)
    from frankensearch_b200.pipeline import DeviceLexical

    return DeviceLexical(torch.from_numpy(ids).to(dev), torch.from_numpy(scores).to(dev), torch.from_numpy(counts).to(dev))


@pytest.mark.parametrize("k,with_lexical", [(10, True), (10, False), (100, True)])
def test_two_tier_pipeline_matches_oracle_flow(fo, k, with_lexical):
    import torch

    import frankensearch_b200 as fs
    from frankensearch_b200.pipeline import DeviceTwoTierSearcher, fused_to_numpy, hits_to_numpy

    n, batch = 120_000, 21
    dev = torch.device("cuda", 0)
    fast_slab, _ = fo.synth_rows(1, 21, 0, n, 256)
    quality_slab, _ = fo.synth_rows(1, 22, 0, n, 384)
    fast_ix = fs.GpuVectorIndex.from_f16_bits(None, fast_slab)
    quality_ix = fs.GpuVectorIndex.from_f16_bits(None, quality_slab)
    fq = np.stack([fo.clustered_query(q, 256) for q in range(batch)])
    qq = np.stack([fo.clustered_query(500 + q, 384) for q in range(batch)])
    searcher = DeviceTwoTierSearcher(fast_ix, quality_ix)
    fetch = searcher.fetch_for(k)
    lists = None
    if with_lexical:
        lists = []
        for b in range(batch):
            rows, _ = fo.search_top_k(fast_slab, fq[b], fetch)
            lists.append(flows.synthetic_lexical(rows, n, fetch if b % 3 else fetch // 2, seed=b))
    lex = _dev_lexical(torch, fs, lists, dev) if with_lexical else None
    res = searcher.search_device(torch.from_numpy(fq).to(dev), torch.from_numpy(qq).to(dev), k, lex)
    torch.cuda.synchronize()
    fast = hits_to_numpy(res.fast_hits)
    blended = hits_to_numpy(res.blended)
    q_scores = res.quality_scores.cpu().numpy()
    for b in range(batch):
        want = flows.oracle_two_tier(fast_slab, quality_slab, fq[b], qq[b], k, lists[b] if with_lexical else None)
        assert int(res.fast_counts[b]) == fetch
        assert fast["row"][b].tolist() == [int(r) for r in want["fast"][0]], b
        assert np.array_equal(bits(fast["score"][b]), bits(want["fast"][1])), b
        assert np.array_equal(bits(q_scores[b]), bits(want["quality"])), b
        nb = int(res.blended_counts[b])
        flows.assert_hits_equal(blended[b, :nb], want["blended"], f"blended b={b}")
        if with_lexical:
            ini, ref = fused_to_numpy(res.initial), fused_to_numpy(res.refined)
            flows.assert_fused_equal(ini[b, :int(res.initial_counts[b])], want["initial"], f"initial b={b}")
            flows.assert_fused_equal(ref[b, :int(res.refined_counts[b])], want["refined"], f"refined b={b}")
        else:
            flows.assert_hits_equal(hits_to_numpy(res.initial)[b, :int(res.initial_counts[b])], want["initial"], f"initial b={b}")
            flows.assert_hits_equal(hits_to_numpy(res.refined)[b, :int(res.refined_counts[b])], want["refined"], f"refined b={b}")
    fast_ix.close()
    quality_ix.close()


def test_blend_device_union_form_matches_oracle(fo):
    """fsgpu_blend_two_tier_device, union form (blend_two_tier, blend.rs:107-191) for a ragged batch."""
    import torch

    import frankensearch_b200 as fs
    from frankensearch_b200.pipeline import HIT_DTYPE

    rng = np.random.default_rng(5)
    dev = torch.device("cuda", 0)
    batch, nf, nq = 9, 40, 28
    fast = np.zeros((batch, nf), dtype=HIT_DTYPE)
    qual = np.zeros((batch, nq), dtype=HIT_DTYPE)
    fc = np.zeros(batch, dtype=np.int32)
    qc = np.zeros(batch, dtype=np.int32)
    want = []
    for b in range(batch):
        a, c = int(rng.integers(0, nf + 1)), int(rng.integers(0, nq + 1))
        if b == 0:
            a, c = 0, nq
        if b == 1:
            a, c = nf, 0
        fr = rng.choice(400, size=a, replace=False)
        qr = rng.choice(400, size=c, replace=False)
        fs_ = np.sort(rng.standard_normal(a).astype(np.float32))[::-1]
        qs_ = np.sort(rng.standard_normal(c).astype(np.float32))[::-1]
        if b == 2 and a > 3:
            fs_[:] = fs_[0]  # degenerate range -> every normalised score is 1.0 (blend.rs:41-77)
        fast["row"][b, :a], fast["score"][b, :a] = fr, fs_
        qual["row"][b, :c], qual["score"][b, :c] = qr, qs_
        fast["row"][b, a:] = 0xFFFFFFFF
        qual["row"][b, c:] = 0xFFFFFFFF
        fc[b], qc[b] = a, c
        want.append(fo.blend_two_tier([(flows.doc_id(r), int(r), s) for r, s in zip(fr, fs_)],
                                      [(flows.doc_id(r), int(r), s) for r, s in zip(qr, qs_)], 0.7))
    t = lambda a: torch.from_numpy(a.view(np.int32).reshape(a.shape + (2,)) if a.dtype == HIT_DTYPE else a).to(dev)  # noqa: E731
    d_fast, d_qual, d_fc, d_qc = t(fast), t(qual), t(fc), t(qc)
    out = torch.zeros((batch, nf + nq, 2), dtype=torch.int32, device=dev)
    cnt = torch.zeros(batch, dtype=torch.int32, device=dev)
    fs._ffi.check(fs._ffi.lib().fsgpu_blend_two_tier_device(0, 0.7, batch, d_fast.data_ptr(), None, d_fc.data_ptr(), nf,
                                                            d_qual.data_ptr(), None, None, None, d_qc.data_ptr(), nq,
                                                            out.data_ptr(), cnt.data_ptr(), None))
    got = np.ascontiguousarray(out.cpu().numpy()).view(HIT_DTYPE).reshape(batch, nf + nq)
    for b in range(batch):
        flows.assert_hits_equal(got[b, :int(cnt[b])], want[b], f"b={b}")


def test_merge_payload_follows_the_keys(fo):
    """fsgpu_merge_payload_device: after a 3-shard merge every surviving key carries the payload its
    shard stored beside it; empty slots report present = 0."""
    import torch

    import frankensearch_b200 as fs
    from oracle import np_oracle as no

    rng = np.random.default_rng(11)
    dev = torch.device("cuda", 0)
    g, batch, k = 3, 5, 16
    keys = np.zeros((g, batch, k), dtype=np.uint64)
    payload = np.zeros((g, batch, k), dtype=np.float32)
    for s in range(g):
        for b in range(batch):
            n = k if (s + b) % 4 else k - 5  # ragged: 0-padded tails
            sc = rng.standard_normal(n).astype(np.float32)
            rows = (rng.choice(1000, size=n, replace=False) + 1000 * s).astype(np.uint64)
            kk = np.sort(~no.order_keys(sc, rows))[::-1]
            keys[s, b, :n] = kk
            payload[s, b, :n] = (kk & np.uint64(0xFFFF)).astype(np.float32)  # a function of the key
    d_keys = torch.from_numpy(keys.view(np.int64)).to(dev)
    d_pl = torch.from_numpy(payload).to(dev)
    out_keys = torch.zeros((batch, k), dtype=torch.int64, device=dev)
    L = fs._ffi.lib()
    fs._ffi.check(L.fsgpu_merge_top_k_device(0, d_keys.data_ptr(), None, batch, g, k, batch * k, k, k, out_keys.data_ptr(),
                                             None, None, None))
    out_pl = torch.zeros((batch, k), dtype=torch.float32, device=dev)
    out_pr = torch.zeros((batch, k), dtype=torch.uint8, device=dev)
    fs._ffi.check(L.fsgpu_merge_payload_device(0, d_keys.data_ptr(), d_pl.data_ptr(), batch, g, k, batch * k, k, batch * k, k,
                                               out_keys.data_ptr(), k, out_pl.data_ptr(), out_pr.data_ptr(), None))
    mk = out_keys.cpu().numpy().view(np.uint64)
    for b in range(batch):
        allk = np.sort(keys[:, b, :].reshape(-1))[::-1]
        allk = allk[allk != 0][:k]
        assert np.array_equal(mk[b, :len(allk)], allk)
        assert np.array_equal(out_pl[b, :len(allk)].cpu().numpy(), (allk & np.uint64(0xFFFF)).astype(np.float32))
        assert out_pr[b, :len(allk)].cpu().numpy().all()


@pytest.mark.parametrize("shards", [1, 3, 8])
def test_single_process_sharded_index_equals_one_index(fo, shards):
    """fsgpu_sharded_* (the multi-GPU path behind the C ABI, one process): per-shard exact search with the
    results stored into the merge device's buffer, merge == merge_partial_heaps across shards
    (search.rs:1704-1720).  Rows and score bits equal the oracle over the whole corpus, ties across a shard
    boundary go to the lower global row, tombstones are re-packed per shard, k above the live row count
    returns everything.  On a one-GPU box every shard lives on device 0; on a multi-GPU box they spread out."""
    import torch

    import frankensearch_b200 as fs

    n, dim = 90_001, 128
    slab, _ = fo.synth_rows(1, 91, 0, n, dim)
    lo = n // shards if shards > 1 else 100
    slab[lo] = slab[lo - 1]            # an exact tie across the first shard boundary
    slab[n - 1] = slab[3]
    tomb = np.arange(n) % 17 == 3
    devices = [i % torch.cuda.device_count() for i in range(shards)]
    sh = fs.GpuShardedIndex.from_f16_bits(slab, devices, tombstones=tomb)
    assert sh.shard_count() == shards and sh.record_count() == n
    queries = np.stack([fo.clustered_query(q, dim) for q in range(9)] + [slab[lo].view(np.float16).astype(np.float32)])
    bm = fo.pack_bitmap(tomb)
    for k in (1, 10, 300):
        for batch in (1, 2, 10):
            rows, scores, counts = sh.search_top_k_batch(queries[-batch:], k)
            for b in range(batch):
                o_rows, o_scores = fo.search_top_k(slab, queries[-batch:][b], k, bm)
                c = int(counts[b])
                assert c == len(o_rows) and rows[b, :c].tolist() == [int(r) for r in o_rows], (shards, k, batch, b)
                assert np.array_equal(bits(scores[b, :c]), bits(o_scores)), (shards, k, batch, b)
    small = fs.GpuShardedIndex.from_f16_bits(slab[:50], devices[: min(shards, 3)])
    rows, scores, counts = small.search_top_k_batch(queries[:2], 200)
    assert counts.tolist() == [50, 50]
    small.close()
    sh.close()
