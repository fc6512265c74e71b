"""Driver for ncu captures of the batched scan: python tools/profile_mma.py [batch] [k] [rows] [iters]."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import frankensearch_b200 as fs  # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
k = int(sys.argv[2]) if len(sys.argv) > 2 else 10
rows = int(sys.argv[3]) if len(sys.argv) > 3 else 10_000_000
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 3
dev = torch.device("cuda", 0)
slab = torch.empty((rows, 384), dtype=torch.int16, device=dev)
fs._ffi.check(fs._ffi.lib().fsgpu_synth_rows_device(0, 1, 1, 0, rows, 384, 64, 0.30, slab.data_ptr(), None))
ix = fs.GpuVectorIndex.from_device_tensor(slab)
q = torch.randn((batch, 384), device=dev)
q = (q / q.norm(dim=1, keepdim=True)).contiguous()
for _ in range(iters):
    ix.search_top_k_device(q, k)
torch.cuda.synchronize()
ix.close()
