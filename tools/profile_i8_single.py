"""Driver for ncu captures of the single-query int8 pass: python tools/profile_i8_single.py [rows] [k] [iters]."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["FSGPU_MMA_I8"] = "1"
import numpy as np  # noqa: E402
import torch  # noqa: E402

import frankensearch_b200 as fs  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
k = int(sys.argv[2]) if len(sys.argv) > 2 else 10
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dev = torch.device("cuda", 0)
slab = torch.empty((rows, 384), dtype=torch.int16, device=dev)
fs._ffi.check(fs._ffi.lib().fsgpu_synth_rows_device(0, 1, 1, 0, rows, 384, 64, 0.30, slab.data_ptr(), None))
ix = fs.GpuVectorIndex.from_device_tensor(slab)
rng = np.random.default_rng(0)
for _ in range(iters):
    q = rng.standard_normal(384).astype(np.float32)
    ix.search_top_k_batch(q / np.linalg.norm(q), k)
ix.close()
