"""Driver for launch lists / captures of the large-k single-query path (select_kernels.cuh) + RRF:
python tools/profile_select.py [rows] [k] [iters] — one query at a time, fetch = 3k, RRF against a
synthetic lexical list (the per-query step of BASELINE configs[4])."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import frankensearch_b200 as fs  # noqa: E402
from frankensearch_b200.pipeline import DeviceLexical, DeviceTwoTierSearcher  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 6_250_000
k = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dev = torch.device("cuda", 0)
slab = torch.empty((rows, 384), dtype=torch.int16, device=dev)
fs._ffi.check(fs._ffi.lib().fsgpu_synth_rows_device(0, 1, 3, 0, rows, 384, 64, 0.30, slab.data_ptr(), None))
ix = fs.GpuVectorIndex.from_device_tensor(slab)
s = DeviceTwoTierSearcher(ix, None)
fetch = s.fetch_for(k)
q = torch.randn((1, 384), device=dev)
q = (q / q.norm(dim=1, keepdim=True)).contiguous()
ids = torch.randint(0, rows, (1, fetch), device=dev, dtype=torch.int64)
sc = torch.arange(fetch, 0, -1, dtype=torch.float32, device=dev).repeat(1, 1).contiguous()
lex = DeviceLexical(ids, sc)
for _ in range(iters):
    s.search_device(q, None, k, lex)
torch.cuda.synchronize()
e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
t = time.perf_counter()
e0.record()
ix.search_top_k_device(q, fetch)
e1.record()
s.search_device(q, None, k, lex)
e2.record()
torch.cuda.synchronize()
print(f"rows={rows} k={k} fetch={fetch}: search alone {e0.elapsed_time(e1) * 1e3:.0f} us, search + rrf {e1.elapsed_time(e2) * 1e3:.0f} us",
      file=sys.stderr)
ix.close()
