"""Regenerates tests/golden/*.json.  The reference (Rust) cannot be executed in this image, so
these vectors are produced by the two independent restatements of its arithmetic — the C++
oracle (oracle/fs_oracle.cpp) and the NumPy mirror (oracle/np_oracle.py) — and are only written
when both agree bit for bit.  Run from the repo root:  python tests/golden/make_golden.py"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import fs_oracle as fo  # noqa: E402
from oracle import np_oracle as no  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    rng = np.random.default_rng(20261017)
    cases = []
    for dim in (4, 8, 31, 32, 100, 128, 256, 384):
        for _ in range(2):
            q = rng.uniform(-1, 1, dim).astype(np.float32)
            row = fo.encode_f16(rng.uniform(-1, 1, dim).astype(np.float32))
            by_order = []
            for order in range(5):
                a = fo.dot_f16_f32(row, q, order, True, impl=1)
                b = no.dot_f16_f32(row, q, order, True)
                assert a.view(np.uint32) == b.view(np.uint32)
                by_order.append(int(a.view(np.uint32)))
            cases.append(dict(dim=dim, row_bits=[int(x) for x in row], query_bits=[int(x) for x in q.view(np.uint32)],
                              score_bits_by_order=by_order))
    with open(os.path.join(HERE, "dot_f16_f32.json"), "w") as f:
        json.dump(dict(source="oracle/fs_oracle.cpp == oracle/np_oracle.py", cases=cases), f)

    # a small clustered corpus + queries with the expected top-10 (rows, score bits)
    slab, _ = fo.synth_rows(1, 1, 0, 3000, 384)
    out = []
    for qi in range(4):
        q = fo.clustered_query(qi, 384)
        rows, scores = fo.search_top_k(slab, q, 10)
        r2, s2 = no.top_k(no.dot_rows(slab, q, 0), 10)
        assert np.array_equal(rows, r2) and np.array_equal(scores.view(np.uint32), s2.view(np.uint32))
        out.append(dict(query=qi, rows=[int(r) for r in rows], score_bits=[int(s) for s in scores.view(np.uint32)]))
    with open(os.path.join(HERE, "scan_clustered_3000x384.json"), "w") as f:
        json.dump(dict(generator="fso_synth_rows(kind=1, seed_base=1, rows 0..3000, dim 384); clustered_query(q)",
                       reduce_order=0, results=out), f)
    print("golden vectors written")


if __name__ == "__main__":
    main()
