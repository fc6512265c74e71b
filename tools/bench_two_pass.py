"""Latency of the reference-semantics two-pass searches on the GPU box: python tools/bench_two_pass.py [rows] [dim]
(10 M x 384 clustered corpus by default; k = 10; int8 multiplier 3 = TwoTierIndex::search_fast's default, 4-bit 5)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import frankensearch_b200 as fs  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
dim = int(sys.argv[2]) if len(sys.argv) > 2 else 384
dev = torch.device("cuda", 0)
slab = torch.empty((rows, dim), dtype=torch.int16, device=dev)
fs._ffi.check(fs._ffi.lib().fsgpu_synth_rows_device(0, 1, 1, 0, rows, dim, 64, 0.30, slab.data_ptr(), None))
ix = fs.GpuVectorIndex.from_device_tensor(slab)
rng = np.random.default_rng(0)
qs = rng.standard_normal((32, dim)).astype(np.float32)
qs /= np.linalg.norm(qs, axis=1, keepdims=True)
exact = [[h.index for h in ix.search_top_k(q, 10)] for q in qs]
print(f"# rows={rows} dim={dim} k=10, 32 queries, host API (H2D query, D2H hits inside)")
for name, fn, mult in (("exact search_top_k", None, 0), ("int8 two-pass x3", ix.search_top_k_int8_two_pass, 3),
                       ("int8 two-pass x10", ix.search_top_k_int8_two_pass, 10), ("4-bit two-pass x5", ix.search_top_k_4bit_two_pass, 5),
                       ("4-bit two-pass x20", ix.search_top_k_4bit_two_pass, 20)):
    call = (lambda q: ix.search_top_k(q, 10)) if fn is None else (lambda q: fn(q, 10, mult))
    call(qs[0])
    torch.cuda.synchronize()
    t = time.perf_counter()
    got = [[h.index for h in call(q)] for q in qs]
    ms = (time.perf_counter() - t) / len(qs) * 1e3
    recall = np.mean([len(set(a) & set(b)) / 10.0 for a, b in zip(got, exact)])
    print(f"{name:22s} {ms:7.3f} ms per query   recall@10 vs exact {recall:.3f}")
ix.close()
