"""Per-query (exact CUDA-core path) and batched latency across k at one corpus size.
    python tools/sweep_k.py [rows]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import frankensearch_b200 as fs  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
dev = torch.device("cuda", 0)
slab = torch.empty((rows, 384), dtype=torch.int16, device=dev)
fs._ffi.check(fs._ffi.lib().fsgpu_synth_rows_device(0, 1, 1, 0, rows, 384, 64, 0.30, slab.data_ptr(), None))
ix = fs.GpuVectorIndex.from_device_tensor(slab)
q = torch.randn((1024, 384), device=dev)
q = (q / q.norm(dim=1, keepdim=True)).contiguous()
print(f"# rows={rows} dim=384")
print("batch     k   ms/call   queries/s   path")
for batch in (1, 4, 64, 1024):
    for k in (1, 10, 100, 256, 1000, 2000):
        for _ in range(2):
            ix.search_top_k_device(q[:batch], k)
        torch.cuda.synchronize()
        ix.profile_read(reset=True)
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 3
        t0.record()
        for _ in range(n):
            ix.search_top_k_device(q[:batch], k)
        t1.record()
        torch.cuda.synchronize()
        p = ix.profile_read(reset=True)
        ms = t0.elapsed_time(t1) / n
        path = "tensor-core" if p["mma_launches"] else ("score-all+sort" if k > 1024 else "cuda-core")
        print(f"{batch:5d} {k:5d} {ms:9.3f} {batch / (ms * 1e-3):11.0f}   {path}", flush=True)
ix.close()
