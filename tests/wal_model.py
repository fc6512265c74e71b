"""Oracle-side model of VectorIndex's mutable state (main rows + tombstones + resident WAL rows),
used by the tests as the checker for GpuVectorIndex's append / soft_delete / search behaviour.
TEST INFRASTRUCTURE ONLY.  Follows crates/frankensearch-index/src/lib.rs:2314-2396 (soft_delete_batch),
:2581-2720 (append_batch_impl) and search.rs:426-494, :1449-1558 (search + resolve)."""
import numpy as np

from oracle import fs_oracle as fo
from oracle import np_oracle as no


class OracleWalIndex:
    def __init__(self, doc_ids, vectors, dim=None, reduce_order=0, tail_fma=False):
        v = np.asarray(vectors, dtype=np.float32)
        self.dim = int(dim if dim is not None else v.shape[1])
        self.doc_ids = list(doc_ids)
        self.slab = fo.encode_f16(v.reshape(len(self.doc_ids), self.dim))
        self.tomb = np.zeros(len(self.doc_ids), dtype=bool)
        self.wal = []  # [(doc_id, f32[dim])]
        self.reduce_order, self.tail_fma = reduce_order, tail_fma

    def wal_record_count(self):
        return len(self.wal)

    def append(self, doc_id, vector):
        self.append_batch([(doc_id, vector)])

    def append_batch(self, entries):
        seen, fresh = set(), []
        for doc_id, v in reversed(list(entries)):
            if doc_id not in seen:
                seen.add(doc_id)
                fresh.append((doc_id, np.asarray(v, dtype=np.float32)))
        fresh.reverse()
        self.wal = [e for e in self.wal if e[0] not in seen] + fresh
        for r, d in enumerate(self.doc_ids):
            if d in seen:
                self.tomb[r] = True

    def soft_delete(self, doc_id):
        return self.soft_delete_batch([doc_id]) > 0

    def soft_delete_batch(self, doc_ids):
        ids, deleted = set(doc_ids), 0
        for r, d in enumerate(self.doc_ids):
            if d in ids and not self.tomb[r]:
                self.tomb[r] = True
                deleted += 1
        kept = [e for e in self.wal if e[0] not in ids]
        deleted += len(self.wal) - len(kept)
        self.wal = kept
        return deleted

    def raw_search(self, query, k, allow=None):
        """(rows, scores) before the doc-id resolve step; allow = bool[n + n_wal] or None."""
        n, n_wal = len(self.doc_ids), len(self.wal)
        excl = self.tomb.copy()
        wal_allow = None
        if allow is not None:
            allow = np.asarray(allow, dtype=bool)
            excl |= ~allow[:n]
            wal_allow = fo.pack_bitmap(allow[n:]) if n_wal else None
        wal = np.stack([v for _, v in self.wal]) if n_wal else np.zeros((0, self.dim), dtype=np.float32)
        slab = self.slab if n else np.zeros((0, self.dim), dtype=np.uint16)
        return fo.search_top_k_wal(slab, wal, np.asarray(query, dtype=np.float32), k,
                                   fo.pack_bitmap(excl) if n else None, wal_allow, 2,
                                   self.reduce_order, self.tail_fma)

    def search_top_k(self, query, k, filter_ids=None):
        allow = None
        if filter_ids is not None:
            ids = set(filter_ids)
            allow = np.array([d in ids for d in self.doc_ids] + [d in ids for d, _ in self.wal], dtype=bool)
        rows, scores = self.raw_search(query, k, allow)
        return no.resolve_sorted_entries(rows, scores, self.doc_ids, [d for d, _ in self.wal], self.tomb)


def run_scenario(make_index, scenario, on_search=None):
    """Replays a WAL_SCENARIOS entry.  `make_index(doc_ids, vectors, dim)` returns an object with the
    VectorIndex method set; search results are [(row, score, doc_id)]."""
    rows = scenario["rows"]
    first_vec = rows[0][1] if rows else next(s[2] if s[0] == "append" else s[1][0][1]
                                             for s in scenario["steps"] if s[0] in ("append", "append_batch"))
    dim = len(first_vec)
    ix = make_index([d for d, _ in rows], np.asarray([v for _, v in rows], dtype=np.float32).reshape(len(rows), dim), dim)
    for step in scenario["steps"]:
        if step[0] == "append":
            ix.append(step[1], step[2])
        elif step[0] == "append_batch":
            ix.append_batch(step[1])
        elif step[0] == "soft_delete":
            assert ix.soft_delete(step[1]) == step[2], scenario["name"]
        else:
            _, query, k, filt, checks = step
            hits = ix.search_top_k(query, k, filt)
            ids = [h[2] for h in hits]
            if "ids" in checks:
                assert ids == checks["ids"], (scenario["name"], ids)
            if "n" in checks:
                assert len(hits) == checks["n"], (scenario["name"], ids)
            if "first" in checks:
                assert ids[0] == checks["first"], (scenario["name"], ids)
            if "first_score_abs_lt" in checks:
                assert abs(float(hits[0][1])) < checks["first_score_abs_lt"], (scenario["name"], hits[0])
            for d in checks.get("absent", []):
                assert d not in ids, (scenario["name"], ids)
            for d in checks.get("present", []):
                assert d in ids, (scenario["name"], ids)
            if "wal_count" in checks:
                assert ix.wal_record_count() == checks["wal_count"], scenario["name"]
            for a, b in zip(hits, hits[1:]):
                assert not (float(a[1]) < float(b[1])), (scenario["name"], hits)
            if on_search is not None:
                on_search(step, hits)
    return ix
