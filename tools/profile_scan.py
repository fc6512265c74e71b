"""Tiny driver for ncu captures of the scan kernel: python tools/profile_scan.py QB R [rows] [k] [iters]."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import frankensearch_b200 as fs  # noqa: E402

qb, r = int(sys.argv[1]), int(sys.argv[2])
rows = int(sys.argv[3]) if len(sys.argv) > 3 else 10_000_000
k = int(sys.argv[4]) if len(sys.argv) > 4 else 10
iters = int(sys.argv[5]) if len(sys.argv) > 5 else 3
os.environ["FSGPU_SCAN_QB"] = str(qb)
os.environ["FSGPU_SCAN_R"] = str(r)
dev = torch.device("cuda", 0)
slab = torch.empty((rows, 384), dtype=torch.int16, device=dev)
fs._ffi.check(fs._ffi.lib().fsgpu_synth_rows_device(0, 1, 1, 0, rows, 384, 64, 0.30, slab.data_ptr(), None))
ix = fs.GpuVectorIndex.from_device_tensor(slab)
q = torch.randn((qb, 384), device=dev)
q = (q / q.norm(dim=1, keepdim=True)).contiguous()
for _ in range(iters):
    ix.search_top_k_device(q, k)
torch.cuda.synchronize()
ix.close()
