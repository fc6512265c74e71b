"""GPU sweep of the batched tensor-core scan: per-launch time of mma_scan_kernel by batch, k and
TMA ring depth.  Run on the GPU box:  python tools/sweep_mma.py [rows] [dim] > gpurun_out/sweep_mma.txt
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import frankensearch_b200 as fs  # noqa: E402


def main():
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
    dim = int(sys.argv[2]) if len(sys.argv) > 2 else 384
    batches = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [16, 128, 256, 512, 1024, 2048]
    stages = [int(x) for x in sys.argv[4].split(",")] if len(sys.argv) > 4 else [0]
    ks = [int(x) for x in sys.argv[5].split(",")] if len(sys.argv) > 5 else [10, 100]
    dev = torch.device("cuda", 0)
    slab = torch.empty((rows, dim), dtype=torch.int16, device=dev)
    fs._ffi.check(fs._ffi.lib().fsgpu_synth_rows_device(0, 1, 1, 0, rows, dim, 64, 0.30, slab.data_ptr(), None))
    ix = fs.GpuVectorIndex.from_device_tensor(slab)
    q = torch.randn((max(batches), dim), device=dev)
    q = (q / q.norm(dim=1, keepdim=True)).contiguous()
    print(f"# rows={rows} dim={dim} slab={rows * dim * 2 / 1e9:.3f} GB")
    print("batch    k stages  scan_ms   step_ms  slab GB/s  TFLOP/s(slots)   queries/s  redo")
    for st in stages:
        os.environ["FSGPU_MMA_STAGES"] = str(st)
        for k in ks:
            for b in batches:
                for _ in range(2):
                    ix.search_top_k_device(q[:b], k)
                torch.cuda.synchronize()
                ix.profile_read(reset=True)
                ix.profile_enable(True)
                n = 3
                t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                t0.record()
                for _ in range(n):
                    ix.search_top_k_device(q[:b], k)
                t1.record()
                torch.cuda.synchronize()
                p = ix.profile_read(reset=True)
                ix.profile_enable(False)
                ms = p["scan_ms"] / max(p["scan_launches"], 1)
                step = t0.elapsed_time(t1) / n
                tf = p["mma_flops"] / max(p["mma_launches"], 1) / (ms * 1e-3) / 1e12 if p["mma_launches"] else 0.0
                print(f"{b:5d} {k:4d} {st:6d} {ms:8.3f} {step:9.3f} {rows * dim * 2 / (ms * 1e-3) / 1e9:10.0f} "
                      f"{tf:15.1f} {b / (step * 1e-3):11.0f} {p['redo_queries']:5d}", flush=True)
    ix.close()


if __name__ == "__main__":
    main()
