"""GPU tuning sweep for the fused scan kernel: per-launch time and achieved HBM GB/s for every
(QB, R, CTAs/SM) variant at the bench shape.  Run on the GPU box:
    python tools/sweep_scan.py [rows] [dim] [k] > gpurun_out/sweep.txt
"""
import itertools
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import frankensearch_b200 as fs  # noqa: E402


def main():
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
    dim = int(sys.argv[2]) if len(sys.argv) > 2 else 384
    k = int(sys.argv[3]) if len(sys.argv) > 3 else 10
    dev = torch.device("cuda", 0)
    slab = torch.empty((rows, dim), dtype=torch.int16, device=dev)
    fs._ffi.check(fs._ffi.lib().fsgpu_synth_rows_device(0, 1, 1, 0, rows, dim, 64, 0.30, slab.data_ptr(), None))
    ix = fs.GpuVectorIndex.from_device_tensor(slab)
    q = torch.randn((64, dim), device=dev)
    q = (q / q.norm(dim=1, keepdim=True)).contiguous()
    print(f"# rows={rows} dim={dim} k={k} bytes/pass={rows * dim * 2 / 1e9:.3f} GB")
    print("packed qb r ctas/sm  ms/launch  GB/s   queries/s")
    for packed, qb, r, ctas in itertools.product((1, 0), (1, 2, 4, 8), (1, 2), (0, 2)):
        if qb == 8 and r == 2:
            continue
        os.environ["FSGPU_SCAN_PACKED"] = str(packed)
        os.environ["FSGPU_SCAN_QB"] = str(qb)
        os.environ["FSGPU_SCAN_R"] = str(r)
        os.environ["FSGPU_SCAN_CTAS_PER_SM"] = str(ctas)
        batch = qb * 4
        try:
            ix.search_top_k_device(q[:batch], k)
            torch.cuda.synchronize()
            ix.profile_read(reset=True)
            ix.profile_enable(True)
            for _ in range(3):
                ix.search_top_k_device(q[:batch], k)
            torch.cuda.synchronize()
            p = ix.profile_read(reset=True)
            ix.profile_enable(False)
        except fs.SearchError as e:
            print(qb, r, ctas, "ERROR", e)
            continue
        ms = p["scan_ms"] / max(p["scan_launches"], 1)
        gbs = rows * dim * 2 / (ms * 1e-3) / 1e9
        print(f"{packed} {qb:2d} {r} {ctas:7d}  {ms:9.4f}  {gbs:6.0f}  {qb / (ms * 1e-3):9.0f}", flush=True)
    ix.close()


if __name__ == "__main__":
    main()
