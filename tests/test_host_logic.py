"""CPU: host-side logic of the product package (no kernels): FSVI writer layout, shard bounds,
id/tie-rank bookkeeping, the sharded search plumbing over gloo (world_size 2)."""
import os
import struct
import sys
import zlib

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from frankensearch_b200 import fusion, sharded
from frankensearch_b200.fsvi import write_fsvi_v1
from frankensearch_b200.types import candidate_count, fnv1a_hash
from oracle import np_oracle as no


def test_candidate_count():
    """rrf.rs:2652 candidate_count_multiplier_one and the default multiplier 3 (config.rs:174)."""
    assert candidate_count(10, 5, 1) == 15
    assert candidate_count(0, 10, 1) == 10
    assert candidate_count(10, 0, 3) == 30


def test_fsvi_writer_layout(tmp_path):
    """lib.rs:6-43 layout; rows sorted by (fnv1a(doc_id), doc_id) (lib.rs:3758-3762)."""
    rng = np.random.default_rng(1)
    ids = [f"doc-{i:06}" for i in range(50)]
    vec = rng.standard_normal((50, 16)).astype(np.float32)
    p = str(tmp_path / "x.fsvi")
    perm = write_fsvi_v1(p, "bench-16", 16, ids, vec, tombstones=[i == 3 for i in range(50)])
    data = open(p, "rb").read()
    assert data[:4] == b"FSVI" and struct.unpack_from("<H", data, 4)[0] == 1
    cur = 6
    n = struct.unpack_from("<H", data, cur)[0]; cur += 2 + n
    n = struct.unpack_from("<H", data, cur)[0]; cur += 2 + n
    dim, quant = struct.unpack_from("<IB", data, cur); cur += 8
    count, voff = struct.unpack_from("<QQ", data, cur); cur += 16
    assert (dim, quant, count) == (16, 1, 50) and voff % 64 == 0
    assert struct.unpack_from("<I", data, cur)[0] == zlib.crc32(data[:cur]) & 0xFFFFFFFF
    cur += 4
    recs = [struct.unpack_from("<QIHH", data, cur + 16 * i) for i in range(50)]
    hashes = [r[0] for r in recs]
    assert hashes == sorted(hashes)
    strings = data[cur + 800:]
    for row, (h, off, ln, flags) in enumerate(recs):
        doc = strings[off:off + ln].decode()
        assert doc == ids[perm[row]] and h == fnv1a_hash(doc.encode())
        assert flags == (1 if perm[row] == 3 else 0)
    slab = np.frombuffer(data, dtype=np.uint16, count=50 * 16, offset=voff).reshape(50, 16)
    assert np.array_equal(slab, no.encode_f16(vec[perm]))


def test_shard_bounds_cover_and_are_contiguous():
    for n in (0, 1, 7, 1000, 10_000_001):
        for g in (1, 2, 3, 8):
            spans = [sharded.shard_bounds(n, g, r) for r in range(g)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(g - 1))


def test_tie_ranks_follow_reference_orders():
    ids = fusion._dense_ids(["b", "a", "b", "ä", "B"])
    assert ids == {"b": 0, "a": 1, "ä": 2, "B": 3}
    tie = fusion._tie_ranks(ids, "LexicalThenId")
    assert sorted(tie, key=tie.get) == ["B", "a", "b", "ä"]  # byte order, like Rust str::cmp
    tie_h = fusion._tie_ranks(ids, "Hash")
    assert sorted(tie_h, key=tie_h.get) == sorted(ids, key=lambda d: (fnv1a_hash(d.encode()), d.encode()))


# ── sharded search plumbing over gloo, world_size 2 ──────────────────────────────────────────
def _np_local_search(slab_bits, base):
    def run(queries, k):
        keys = np.zeros((queries.shape[0], k), dtype=np.uint64)
        scores = np.zeros((queries.shape[0], k), dtype=np.float32)
        for b in range(queries.shape[0]):
            s = no.dot_rows(slab_bits, queries[b].numpy(), 0)
            rows, sc = no.top_k(s, k)
            kk = no.order_keys(sc, rows + np.uint64(base))
            # product keys are "larger is better": complement of the ascending sort key
            keys[b, :len(kk)] = ~kk
            scores[b, :len(kk)] = sc
        return torch.from_numpy(keys.view(np.int64)), torch.from_numpy(scores)
    return run


def _np_merge(keys, scores, k):
    g, b, k_in = keys.shape
    out = np.zeros((b, k), dtype=np.uint64)
    u = keys.numpy().view(np.uint64)
    for q in range(b):
        allk = np.sort(u[:, q, :].reshape(-1))[::-1]
        allk = allk[allk != 0][:k]
        out[q, :len(allk)] = allk
    return torch.from_numpy(out.view(np.int64)), None, None


def _worker(rank, world, port, slab, queries, k, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = sharded.shard_bounds(slab.shape[0], world, rank)
        ix = sharded.ShardedGpuIndex(None, local_search=_np_local_search(slab[lo:hi], lo), merge=_np_merge)
        keys, _, _ = ix.search_top_k_device(torch.from_numpy(queries), k)
        ret[rank] = keys.numpy().view(np.uint64).copy()
    finally:
        dist.destroy_process_group()


def test_sharded_search_equals_single_index_over_gloo():
    """per-shard top-k + one all-gather + merge == the unsharded answer."""
    from oracle import fs_oracle as fo

    slab, _ = fo.synth_rows(1, 1, 0, 4001, 128)
    slab[1000] = slab[3000]  # an exact cross-shard tie: lower global row must win
    queries = np.stack([fo.clustered_query(q, 128) for q in range(3)])
    k = 25
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, slab, queries, k, ret), nprocs=2, join=True)
    assert np.array_equal(ret[0], ret[1])
    for b in range(3):
        rows, scores = fo.search_top_k(slab, queries[b], k)
        want = ~no.order_keys(scores, rows)
        assert np.array_equal(ret[0][b], want)


def _grid_worker(rank, world, port, slab, queries, k, query_groups, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        r_shards, r, _g = sharded.grid_position(world, rank, query_groups)
        lo, hi = sharded.shard_bounds(slab.shape[0], r_shards, r)
        ix = sharded.ShardedGpuIndex(None, local_search=_np_local_search(slab[lo:hi], lo), merge=_np_merge,
                                     query_groups=query_groups)
        out = []
        for b in (queries.shape[0], 1):  # a ragged split, then a batch that leaves the second group empty
            keys, _, _ = ix.search_top_k_device(torch.from_numpy(queries[:b]), k)
            out.append(keys.numpy().view(np.uint64).copy())
        ret[rank] = out
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,query_groups", [(2, 2), (4, 2)])
def test_row_shards_x_query_groups_equal_single_index_over_gloo(world, query_groups):
    """R row shards x Q query groups (rank = g*R + r searches its query block over its row shard): one all-gather,
    one merge per group == the unsharded answer, for a ragged and for a one-query batch."""
    from oracle import fs_oracle as fo

    slab, _ = fo.synth_rows(1, 1, 0, 3001, 128)
    slab[700] = slab[2500]  # an exact cross-shard tie
    queries = np.stack([fo.clustered_query(q, 128) for q in range(5)])
    k = 12
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 31500 + (os.getpid() + 7 * world) % 2000
    mp.spawn(_grid_worker, args=(world, port, slab, queries, k, query_groups, ret), nprocs=world, join=True)
    for i, b in enumerate((5, 1)):
        for r in range(1, world):
            assert np.array_equal(ret[0][i], ret[r][i])
        assert ret[0][i].shape == (b, k)
        for q in range(b):
            rows, scores = fo.search_top_k(slab, queries[q], k)
            assert np.array_equal(ret[0][i][q], ~no.order_keys(scores, rows))


def test_grid_position_and_query_blocks():
    assert [sharded.grid_position(8, r, 2) for r in (0, 3, 4, 7)] == [(4, 0, 0), (4, 3, 0), (4, 0, 1), (4, 3, 1)]
    assert sharded.grid_position(2, 1, 2) == (1, 0, 1) and sharded.grid_position(4, 3, 1) == (4, 3, 0)
    with pytest.raises(Exception):
        sharded.grid_position(6, 0, 4)
    assert [sharded.query_block(1024, 2, g) for g in range(2)] == [(0, 512), (512, 1024)]
    assert [sharded.query_block(5, 2, g) for g in range(2)] == [(0, 3), (3, 5)]
    assert [sharded.query_block(1, 2, g) for g in range(2)] == [(0, 1), (1, 1)]


# ── sharded two-tier pipeline plumbing over gloo, world_size 2 ───────────────────────────────
def _pipeline_worker(rank, world, port, fast_slab, quality_slab, fq, qq, k, lex_ids, lex_scores, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
        from fake_pipeline_lib import FakePipelineLib, FakeShard
        from frankensearch_b200.pipeline import DeviceLexical, DeviceTwoTierSearcher

        lo, hi = sharded.shard_bounds(fast_slab.shape[0], world, rank)
        fast, quality = FakeShard(1, fast_slab[lo:hi], lo), FakeShard(2, quality_slab[lo:hi], lo)
        s = DeviceTwoTierSearcher(fast, quality)
        s._L = FakePipelineLib([fast, quality])
        lex = DeviceLexical(torch.from_numpy(lex_ids.view(np.int64)), torch.from_numpy(lex_scores)) if lex_ids is not None else None
        r = s.search_device(torch.from_numpy(fq), torch.from_numpy(qq), k, lex)
        ret[rank] = {n: getattr(r, n).numpy().copy() for n in
                     ("fast_hits", "fast_counts", "quality_scores", "quality_present", "initial", "initial_counts",
                      "blended", "blended_counts", "refined", "refined_counts")}
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("with_lexical", [True, False])
def test_sharded_two_tier_pipeline_equals_single_process_flow_over_gloo(with_lexical):
    """SURVEY 8e "Two-tier": per-rank fast top-fetch + local quality re-score, ONE all-gather of
    [keys | hits | quality], merge + payload pickup, blend, RRF == the unsharded oracle flow
    (sync_searcher.rs:616-1009); both ranks hold the same answer."""
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import flows
    from fake_pipeline_lib import FUSED, HIT
    from oracle import fs_oracle as fo

    n, k, batch = 3001, 7, 3
    fast_slab, _ = fo.synth_rows(1, 21, 0, n, 128)
    quality_slab, _ = fo.synth_rows(1, 22, 0, n, 64)
    fast_slab[700] = fast_slab[2400]  # an exact cross-shard tie: the lower global row wins
    fq = np.stack([fo.clustered_query(q, 128) for q in range(batch)])
    qq = np.stack([fo.clustered_query(50 + q, 64) for q in range(batch)])
    fetch = 3 * k
    lex_ids = lex_scores = None
    lists = []
    if with_lexical:
        for b in range(batch):
            rows, _ = fo.search_top_k(fast_slab, fq[b], fetch)
            lists.append(flows.synthetic_lexical(rows, n, fetch, seed=b))
        lex_ids = np.stack([i for i, _ in lists])
        lex_scores = np.stack([s for _, s in lists])
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 31500 + os.getpid() % 2000 + (1 if with_lexical else 0)
    mp.spawn(_pipeline_worker, args=(2, port, fast_slab, quality_slab, fq, qq, k, lex_ids, lex_scores, ret), nprocs=2, join=True)
    for name in ret[0]:
        assert np.array_equal(ret[0][name], ret[1][name]), name
    got = ret[0]
    fast = np.ascontiguousarray(got["fast_hits"]).view(HIT).reshape(batch, fetch)
    blended = np.ascontiguousarray(got["blended"]).view(HIT).reshape(batch, fetch)
    for b in range(batch):
        want = flows.oracle_two_tier(fast_slab, quality_slab, fq[b], qq[b], k, lists[b] if with_lexical else None)
        assert fast["row"][b].tolist() == [int(r) for r in want["fast"][0]]
        assert np.array_equal(got["quality_scores"][b].view(np.uint32), want["quality"].view(np.uint32))
        flows.assert_hits_equal(blended[b, :int(got["blended_counts"][b])], want["blended"], f"blended {b}")
        if with_lexical:
            for phase in ("initial", "refined"):
                f = np.ascontiguousarray(got[phase]).view(FUSED).reshape(batch, k)
                flows.assert_fused_equal(f[b, :int(got[phase + "_counts"][b])], want[phase], f"{phase} {b}")
        else:
            for phase in ("initial", "refined"):
                h = np.ascontiguousarray(got[phase]).view(HIT).reshape(batch, k)
                flows.assert_hits_equal(h[b, :int(got[phase + "_counts"][b])], want[phase], f"{phase} {b}")


# ── phase-2 diagnostics (crates/frankensearch-fusion/src/blend.rs:365-544 and its tests) ─────
def _hit(doc, score, index):
    from frankensearch_b200.types import VectorHit

    return VectorHit(index, score, doc)


def test_rank_changes_known_answers():
    from frankensearch_b200.fusion import RankChanges, build_rank_map, compute_rank_changes

    initial = [_hit("a", 1.0, 0), _hit("b", 0.9, 1), _hit("c", 0.8, 2)]
    refined = [_hit("b", 1.0, 1), _hit("a", 0.9, 0), _hit("d", 0.7, 3)]
    # compute_rank_changes_tracks_promoted_demoted_stable (blend.rs:817): b up + d new, a down + c dropped
    assert compute_rank_changes(initial, refined) == RankChanges(2, 2, 0)
    # compute_rank_changes_identical_lists_are_all_stable (blend.rs:898), _empty_lists (:907)
    assert compute_rank_changes(initial[:2], initial[:2]) == RankChanges(0, 0, 2)
    assert compute_rank_changes([], []) == RankChanges(0, 0, 0)
    # compute_rank_changes_with_maps_all_new (blend.rs:1166)
    assert compute_rank_changes([], refined) == RankChanges(3, 0, 0)
    # build_borrowed_rank_map_first_occurrence_wins (blend.rs:1131)
    assert build_rank_map([_hit("dup", 1.0, 0), _hit("other", 0.9, 1), _hit("dup", 0.5, 2)]) == {"dup": 0, "other": 1}


def test_kendall_tau_known_answers():
    from frankensearch_b200.fusion import kendall_tau

    abc = [_hit("a", 1.0, 0), _hit("b", 0.9, 1), _hit("c", 0.8, 2)]
    assert kendall_tau(abc, [_hit("a", 0.7, 0), _hit("b", 0.6, 1), _hit("c", 0.5, 2)]) == 1.0     # blend.rs:828
    assert kendall_tau(abc, [_hit("c", 0.7, 2), _hit("b", 0.6, 1), _hit("a", 0.5, 0)]) == -1.0    # blend.rs:836
    assert kendall_tau([_hit("a", 1.0, 0)], [_hit("b", 0.9, 1)]) is None                          # blend.rs:844
    # kendall_tau_partial_overlap (blend.rs:915): common a, c, d -> refined ranks [2, 0, 3] -> 1/3
    initial = [_hit("a", 1.0, 0), _hit("b", 0.9, 1), _hit("c", 0.8, 2), _hit("d", 0.7, 3)]
    refined = [_hit("c", 1.0, 2), _hit("x", 0.9, 4), _hit("a", 0.8, 0), _hit("d", 0.7, 3)]
    assert abs(kendall_tau(initial, refined) - 1.0 / 3.0) < 1e-10
    assert kendall_tau([_hit("a", 1.0, 0), _hit("b", 0.5, 1)], [_hit("b", 1.0, 1), _hit("a", 0.5, 0)]) == -1.0  # :940
    docs = [f"doc-{i:04}" for i in range(100)]
    fwd = [_hit(d, 1.0, 0) for d in docs]
    assert kendall_tau(fwd, [_hit(d, 0.5, 0) for d in docs]) == 1.0                                # blend.rs:951
    assert kendall_tau(fwd, [_hit(d, 0.5, 0) for d in reversed(docs)]) == -1.0                     # blend.rs:968


def test_kendall_tau_matches_naive_for_deterministic_permutations():
    """blend.rs:1021: xorshift-shuffled rankings, merge-sort inversion count == O(n^2) pair count."""
    from frankensearch_b200.fusion import kendall_tau

    def shuffle(values, seed):
        state = (seed + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
        for i in range(len(values) - 1, 0, -1):
            state ^= (state << 13) & 0xFFFFFFFFFFFFFFFF
            state ^= state >> 7
            state ^= (state << 17) & 0xFFFFFFFFFFFFFFFF
            j = state % (i + 1)
            values[i], values[j] = values[j], values[i]

    for n in (2, 3, 5, 17, 64, 257):
        for seed in (1, 7, 1234567):
            order = list(range(n))
            shuffle(order, seed)
            initial = [_hit(f"d{i}", 1.0, i) for i in range(n)]
            refined = [_hit(f"d{i}", 1.0, i) for i in order]
            rank = {f"d{i}": r for r, i in enumerate(order)}
            ranks = [rank[f"d{i}"] for i in range(n)]
            disc = sum(1 for i in range(n) for j in range(i + 1, n) if ranks[i] > ranks[j])
            total = n * (n - 1) // 2
            assert kendall_tau(initial, refined) == ((total - disc) - disc) / total


# ── host half of the WAL / soft-delete / filter path (frankensearch_b200/index.py) on the CPU ─
class _CpuWalIndex:
    """GpuVectorIndex over tests/fake_lib.py behind the scenario interface of tests/wal_model.py."""

    def __init__(self, doc_ids, vectors, dim):
        from fake_lib import make_cpu_index

        self.ix = make_cpu_index(doc_ids, vectors, dim)

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


def _wal_scenarios():
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import ref_cases as rc

    return rc.WAL_SCENARIOS


@pytest.mark.parametrize("scenario", _wal_scenarios(), ids=[s["name"] for s in _wal_scenarios()])
def test_index_host_logic_replays_the_reference_wal_tests(scenario):
    """The reference's own WAL / filter / soft-delete tests (tests/ref_cases.py WAL_SCENARIOS) through
    the product's host code — append dedup and supersede, tombstoning of replaced main rows, filter
    -> allow bitmap incl. the WAL bits, doc-id resolve with shadowing — with the device calls served
    by the CPU oracle (tests/fake_lib.py).  Hits must also equal the oracle model's, bit for bit."""
    from wal_model import OracleWalIndex, run_scenario

    got, want = [], []
    run_scenario(lambda ids, vecs, dim: _CpuWalIndex(ids, vecs, dim), scenario, on_search=lambda s, h: got.append(h))
    run_scenario(lambda ids, vecs, dim: OracleWalIndex(ids, vecs, dim), scenario, on_search=lambda s, h: want.append(h))
    assert len(got) == len(want)
    for g, w in zip(got, want):
        assert [(r, d) for r, _, d in g] == [(r, d) for r, _, d in w]
        assert np.array_equal(np.array([s for _, s, _ in g], dtype=np.float32).view(np.uint32),
                              np.array([s for _, s, _ in w], dtype=np.float32).view(np.uint32))


def test_index_host_logic_append_validation_and_bookkeeping():
    """append_batch_impl (lib.rs:2581-2720) and soft_delete_batch (lib.rs:2314-2396) on the host:
    validation order and error kinds, last-entry-wins inside a batch, superseding resident rows,
    tombstones pushed to the device exactly when they change, WAL rows numbered after the slab."""
    from fake_lib import make_cpu_index
    from frankensearch_b200 import SearchError

    ids = [f"doc-{i}" for i in range(6)]
    vec = np.eye(6, 4, dtype=np.float32) + 0.1
    ix = make_cpu_index(ids, vec, 4)
    lib = ix._L
    for bad, kind in (([1.0, 0.0], "DimensionMismatch"), ([float("nan"), 0, 0, 0], "InvalidConfig"),
                      ([0.0, 0.0, 0.0, 0.0], "InvalidConfig"), ([float("inf"), 0, 0, 0], "InvalidConfig")):
        with pytest.raises(SearchError) as e:
            ix.append_batch([("ok", [1.0, 0, 0, 0]), ("bad", bad)])  # nothing is admitted when one entry is bad
        assert e.value.kind == kind
    assert ix.wal_record_count() == 0 and lib.calls == []
    ix.append_batch([("new-a", [1, 0, 0, 0]), ("doc-2", [0, 1, 0, 0]), ("new-a", [0, 0, 1, 0])])
    assert [d for d, _ in ix.wal_records()] == ["doc-2", "new-a"]            # last entry of new-a, batch order kept
    assert np.array_equal(ix.wal_records()[1][1], np.array([0, 0, 1, 0], dtype=np.float32))
    assert lib.calls == [("set_wal", 2, 6), "set_tombstones"] and lib.tomb.tolist() == [False, False, True, False, False, False]
    assert ix.is_deleted(2) and not ix.is_deleted(3)
    ix.append("new-a", [0, 0, 0, 2])                                          # supersedes the resident row
    assert [d for d, _ in ix.wal_records()] == ["doc-2", "new-a"] and lib.calls[-1] == ("set_wal", 2, 6)
    assert ix.doc_id_at(6) == "doc-2" and ix.doc_id_at(7) == "new-a" and ix.doc_id_at(1) == "doc-1"
    hits = ix.search_top_k([0.0, 0.0, 0.0, 1.0], 3)
    assert hits[0].doc_id == "new-a" and hits[0].index == 7                   # record_count + wal_idx
    assert ix.soft_delete_batch(["doc-2", "doc-3", "missing"]) == 2           # doc-2: WAL row (main already dead), doc-3: main
    assert ix.wal_record_count() == 1 and lib.tomb.tolist() == [False, False, True, True, False, False]
    assert ix.soft_delete("doc-3") is False                                   # already tombstoned
    n_calls = len(lib.calls)
    assert ix.soft_delete("nobody") is False and len(lib.calls) == n_calls    # nothing changed, nothing uploaded
    # a boolean-array filter over the main rows is extended with "allowed" for the WAL rows
    mask = np.zeros(6, dtype=bool)
    mask[0] = True
    got = [h.doc_id for h in ix.search_top_k([1.0, 1.0, 1.0, 1.0], 10, filter=mask)]
    assert sorted(got) == ["doc-0", "new-a"]
    ix.close()
    assert lib.calls[-1] == "destroy"


def test_bitset_and_predicate_filters_follow_the_reference_seam(fo):
    """BitsetFilter / PredicateFilter (crates/frankensearch-core/src/filter.rs:330-383 and its tests):
    a bitset filter decides on the FNV-1a hash of the doc id alone and can enumerate its hashes (the
    gather arm needs that); a predicate filter decides on the doc id and cannot."""
    from frankensearch_b200 import BitsetFilter, PredicateFilter

    ids = ["doc-a", "doc-b", "ünïcode-id", ""]
    f = BitsetFilter.from_doc_ids(ids[:2])
    assert f.matches("doc-a") and f.matches("doc-b") and not f.matches("doc-c") and not f.matches("")
    assert f.candidate_hashes() == frozenset(fo.fnv1a64(d.encode("utf-8")) for d in ids[:2])
    for d in ids:
        h = fo.fnv1a64(d.encode("utf-8"))
        assert fnv1a_hash(d.encode("utf-8")) == h                      # host hash == oracle hash (lib.rs:6120-6127)
        assert f.matches_doc_id_hash(h) == (d in ids[:2])              # Some(set.contains(hash))
    g = BitsetFilter.from_hashes(f.candidate_hashes())
    assert g.matches("doc-a") and not g.matches("nope") and g.name == "bitset_filter"
    assert BitsetFilter.from_doc_ids([]).candidate_hashes() == frozenset()
    p = PredicateFilter("only-b", lambda d: d == "doc-b")
    assert p.matches("doc-b") and not p.matches("doc-a") and p("doc-b") and p.name == "only-b"
    assert p.matches_doc_id_hash(123) is None and p.candidate_hashes() is None


# ── WAL sidecar (crates/frankensearch-index/src/wal.rs) and the WAL half of VectorIndex::open ──
def test_wal_sidecar_layout_roundtrip_and_crash_tolerance(tmp_path):
    """wal.rs:1-27 layout, :852-866 read_wal, :980-1130 parse: header CRC, batch CRC, a corrupt or
    truncated final batch ends the replay without an error, a bad header IS an error."""
    from frankensearch_b200 import fsvi
    from frankensearch_b200._ffi import SearchError

    p = str(tmp_path / "x.fsvi.wal")
    assert fsvi.wal_path_for("/a/b/data.fsvi") == "/a/b/data.fsvi.wal"  # wal.rs:2493-2499
    assert fsvi.read_wal(p, 4) == ([], 0, 0)
    v = lambda *x: np.array(x, dtype=np.float32)  # noqa: E731
    fsvi.append_wal_batch(p, [("a", v(1, 0, 0, 0)), ("b", v(0, 1, 0, 0))], 4, compaction_gen=1)
    fsvi.append_wal_batch(p, [("a", v(0, 0, 1, 0))], 4, compaction_gen=1)
    data = open(p, "rb").read()
    assert data[:4] == b"FWAL" and struct.unpack_from("<HIBB", data, 4) == (1, 4, 1, 1)
    assert struct.unpack_from("<I", data, 16)[0] == zlib.crc32(data[:16]) & 0xFFFFFFFF
    assert data[20:24] == b"FWB1" and struct.unpack_from("<I", data, 24)[0] == 2
    entries, gen, valid = fsvi.read_wal(p, 4)
    assert [d for d, _ in entries] == ["a", "b", "a"] and gen == 1 and valid == len(data)
    assert entries[2][1].tolist() == [0, 0, 1, 0]
    with open(p, "ab") as f:  # a torn third batch: replay stops before it
        f.write(b"FWB1" + struct.pack("<I", 5) + b"\x01\x00z")
    entries2, _, valid2 = fsvi.read_wal(p, 4)
    assert len(entries2) == 3 and valid2 == valid
    with pytest.raises(SearchError) as e:  # dimension mismatch in the header
        fsvi.read_wal(p, 8)
    assert e.value.kind == "IndexCorrupted"
    bad = bytearray(data)
    bad[7] ^= 1
    open(p, "wb").write(bytes(bad))
    with pytest.raises(SearchError):  # header CRC
        fsvi.read_wal(p, 4)
    open(p, "wb").write(b"FWAL\x01")  # shorter than the header: ignored (wal.rs:861-864)
    assert fsvi.read_wal(p, 4) == ([], 0, 0)


def test_wal_replay_last_wins_and_stale_generation(tmp_path):
    """VectorIndex::open, lib.rs:1845-1878: the last entry of a doc id wins (order of the last entries);
    a sidecar whose generation is not the successor of the main file's is discarded."""
    from frankensearch_b200 import fsvi

    path = str(tmp_path / "idx.fsvi")
    vec = np.eye(4, dtype=np.float32)
    fsvi.write_fsvi_v1(path, "e", 4, ["m0", "m1", "m2", "m3"], vec)
    h = fsvi.read_fsvi_header(path)
    assert (h["dimension"], h["quantization"], h["record_count"], h["compaction_gen"]) == (4, 1, 4, 0)
    wal = fsvi.wal_path_for(path)
    fsvi.append_wal_batch(wal, [("x", vec[0]), ("y", vec[1]), ("x", vec[2])], 4, compaction_gen=1)
    fsvi.append_wal_batch(wal, [("z", vec[3]), ("y", vec[0])], 4, compaction_gen=1)
    kept = fsvi.replay_wal_for(path)
    assert [d for d, _ in kept] == ["x", "z", "y"]
    assert kept[0][1].tolist() == vec[2].tolist() and kept[2][1].tolist() == vec[0].tolist()
    os.remove(wal)
    fsvi.append_wal_batch(wal, [("x", vec[0])], 4, compaction_gen=3)  # not next_generation(0) == 1: stale
    assert fsvi.replay_wal_for(path) == []
    os.remove(wal)
    fsvi.append_wal_batch(wal, [("x", vec[0])], 4, compaction_gen=0)  # legacy generation 0 on a generation-0 file
    assert [d for d, _ in fsvi.replay_wal_for(path)] == ["x"]
    assert fsvi.next_generation(255) == 1 and fsvi.next_generation(7) == 8


# ── budgets, query classes, zero-signal classification (host arithmetic of the searcher) ─────
def test_scaled_budget_and_class_multipliers():
    """searcher.rs:120-126, :1599-1608 and query_class.rs:197-215 (identifier_leans_lexical :375,
    natural_language_leans_semantic :383, short_keyword_is_balanced :391, empty_has_zero_budgets :401)."""
    from frankensearch_b200.types import QueryClass, phase1_budgets, scaled_budget

    assert scaled_budget(30, 0.5) == 15 and scaled_budget(30, 2.0) == 60 and scaled_budget(30, 1.0) == 30
    assert scaled_budget(3, 0.5) == 2 and scaled_budget(1, 0.5) == 1  # ceil, at least 1
    assert scaled_budget(0, 2.0) == 0 and scaled_budget(30, 0.0) == 0 and scaled_budget(30, float("nan")) == 0
    assert scaled_budget(30, float("inf")) == 0 and scaled_budget(30, -1.0) == 0
    qc = QueryClass
    assert qc.lexical_budget_multiplier(qc.IDENTIFIER) > qc.semantic_budget_multiplier(qc.IDENTIFIER)
    assert qc.semantic_budget_multiplier(qc.NATURAL_LANGUAGE) > qc.lexical_budget_multiplier(qc.NATURAL_LANGUAGE)
    assert qc.lexical_budget_multiplier(qc.SHORT_KEYWORD) == qc.semantic_budget_multiplier(qc.SHORT_KEYWORD) == 1.0
    assert qc.lexical_budget_multiplier(qc.EMPTY) == qc.semantic_budget_multiplier(qc.EMPTY) == 0.0
    assert phase1_budgets(10, 3, qc.IDENTIFIER) == (30, 15, 60)
    assert phase1_budgets(10, 3, qc.NATURAL_LANGUAGE) == (30, 60, 15)
    assert phase1_budgets(10, 0, qc.SHORT_KEYWORD) == (10, 10, 10)  # multiplier.max(1)


def test_query_class_classify_reference_cases():
    """query_class.rs:238-372, literal for literal."""
    from frankensearch_b200.types import QueryClass as Q

    table = {
        "": Q.EMPTY, "   ": Q.EMPTY, "\\t\\n".encode().decode("unicode_escape"): Q.EMPTY,
        "src/main.rs": Q.IDENTIFIER, "path/to/file.txt": Q.IDENTIFIER,
        "how should we handle HTTP status 404/500 errors": Q.NATURAL_LANGUAGE,
        "http 404/500": Q.SHORT_KEYWORD, "bd-123": Q.IDENTIFIER, "JIRA-456": Q.IDENTIFIER,
        "my-project-123": Q.IDENTIFIER, "repo_name-789": Q.IDENTIFIER,
        "error-handling": Q.SHORT_KEYWORD, "load-balancer": Q.SHORT_KEYWORD, "bd-ab": Q.SHORT_KEYWORD,
        "std::collections::HashMap": Q.IDENTIFIER, "config.toml": Q.IDENTIFIER,
        "fn search_query": Q.IDENTIFIER, "struct TwoTierConfig": Q.IDENTIFIER,
        "search": Q.SHORT_KEYWORD, "error handling": Q.SHORT_KEYWORD, "vector index search": Q.SHORT_KEYWORD,
        "how does the search pipeline work?": Q.NATURAL_LANGUAGE,
        "find all documents about distributed consensus": Q.NATURAL_LANGUAGE,
    }
    for text, want in table.items():
        assert Q.classify(text) == want, text
        assert Q.classify("  " + text + " ") == want, text  # classify_is_trim_invariant (:436)


def test_zero_signal_state_classification_table():
    """config.rs:1080-1153 zero_signal_state_classification_table + empty_result_reason (:729-740)."""
    from frankensearch_b200.types import ZeroSignalReason as R
    from frankensearch_b200.types import ZeroSignalState as S

    cases = [(S(), R.NEWLY_CREATED_EMPTY), (S(5, 0, 5, 0, 0), R.ALL_TOMBSTONED), (S(5, 3, 2, 0, 0), R.NO_USABLE_VECTORS),
             (S(0, 0, 0, 4, 0), None), (S(5, 0, 5, 2, 0), None), (S(5, 5, 0, 0, 5), None)]
    for state, want in cases:
        assert state.state_reason() == want, state
    assert S(0, 0, 0, 4, 0).is_wal_only()
    assert S(5, 5, 0, 0, 5).empty_result_reason(True) == R.FILTER_ELIMINATED_ALL
    assert S(5, 0, 5, 2, 0).empty_result_reason(False) == R.WAL_ONLY_NO_LIVE_RECORDS
    assert S(5, 5, 0, 0, 5).empty_result_reason(False) == R.NO_USABLE_VECTORS
    assert S(5, 0, 5, 0, 0).empty_result_reason(True) == R.ALL_TOMBSTONED  # state reasons take precedence


def test_minilm_host_sequence_policy():
    """model_manifest.rs:74-80, :300-304: special tokens added, truncation to 512 INCLUDING them,
    batch-longest padding, the empty text never reaches the tokenizer."""
    from frankensearch_b200.embed import minilm_pad_batch, minilm_sequence, minilm_token_ids

    assert minilm_sequence([7, 8, 9]) == [101, 7, 8, 9, 102]
    assert minilm_sequence([]) == [101, 102]
    long = list(range(1000, 1700))
    s = minilm_sequence(long)
    assert len(s) == 512 and s[0] == 101 and s[-1] == 102 and s[1:-1] == long[:510]
    ids, lens = minilm_pad_batch([[101, 5, 102], [101, 102], []])
    assert ids.tolist() == [[101, 5, 102], [101, 102, 0], [0, 0, 0]] and lens.tolist() == [3, 2, 0]

    from tokenizers import Tokenizer, models, pre_tokenizers

    vocab = {"[PAD]": 0, "[UNK]": 1, "[CLS]": 2, "[SEP]": 3, "hello": 4, "world": 5, "##s": 6}
    tok = Tokenizer(models.WordPiece(vocab, unk_token="[UNK]"))
    tok.pre_tokenizer = pre_tokenizers.Whitespace()
    assert minilm_token_ids(tok, "hello worlds") == [2, 4, 5, 6, 3]  # this tokenizer's own [CLS]/[SEP] ids
    assert minilm_token_ids(tok, "") == []
