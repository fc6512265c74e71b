"""The single-process multi-GPU path behind the C ABI (fsgpu_sharded_*): rows sharded over every visible
GPU, host queries in, host hits out.  python tools/bench_sharded_abi.py [rows] [batch] [k] [steps]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import frankensearch_b200 as fs  # noqa: E402
from frankensearch_b200.sharded import shard_bounds  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
k = int(sys.argv[3]) if len(sys.argv) > 3 else 10
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 20
g = torch.cuda.device_count()
dim = 384
shards, slabs = [], []
for r in range(g):
    lo, hi = shard_bounds(rows, g, r)
    slab = torch.empty((hi - lo, dim), dtype=torch.int16, device=torch.device("cuda", r))
    fs._ffi.check(fs._ffi.lib().fsgpu_synth_rows_device(r, 1, 1, lo, hi - lo, dim, 64, 0.30, slab.data_ptr(), None))
    shards.append(fs.GpuVectorIndex.from_device_tensor(slab, row_base=lo))
sh = fs.GpuShardedIndex.from_indexes(shards)
q = torch.randn((batch, dim))
q = (q / q.norm(dim=1, keepdim=True)).numpy().astype(np.float32)
for _ in range(3):
    out = sh.search_top_k_batch(q, k)
t = time.perf_counter()
for _ in range(steps):
    out = sh.search_top_k_batch(q, k)
dt = (time.perf_counter() - t) / steps
# parity against one shard-by-shard host merge of the same per-shard answers
keys = []
for ix in shards:
    r_, s_, c_ = ix.search_top_k_batch(q[:4], k)
    keys.append((r_, s_))
for b in range(4):
    cand = sorted(((-float(s_[b, i]), int(r_[b, i])) for r_, s_ in keys for i in range(k)))[:k]
    assert [c[1] for c in cand] == out[0][b].tolist(), "sharded result differs from the merged per-shard answers"
print(f"gpus={g} rows={rows} batch={batch} k={k}: {dt * 1e3:.3f} ms per call = {batch / dt:.0f} queries/s "
      f"(host queries -> host hits); direct peer stores: {sh.direct_shards()}")
sh.close()
