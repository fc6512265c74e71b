"""Per-phase device timeline of one batched search (FSGPU_MMA_TRACE=1): where the fixed per-call
time goes.  Run on the GPU box:  python tools/trace_phases.py [rows,rows,...] [batches] [k] 2> gpurun_out/trace.txt
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import frankensearch_b200 as fs  # noqa: E402


def main():
    rows_list = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [1_250_000, 10_000_000]
    batches = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [128, 1024]
    k = int(sys.argv[3]) if len(sys.argv) > 3 else 10
    dim = 384
    dev = torch.device("cuda", 0)
    for rows in rows_list:
        slab = torch.empty((rows, dim), dtype=torch.int16, device=dev)
        fs._ffi.check(fs._ffi.lib().fsgpu_synth_rows_device(0, 1, 1, 0, rows, dim, 64, 0.30, slab.data_ptr(), None))
        ix = fs.GpuVectorIndex.from_device_tensor(slab)
        q = torch.randn((max(batches), dim), device=dev)
        q = (q / q.norm(dim=1, keepdim=True)).contiguous()
        for b in batches:
            os.environ["FSGPU_MMA_TRACE"] = "0"
            for _ in range(3):
                ix.search_top_k_device(q[:b], k)
            torch.cuda.synchronize()
            os.environ["FSGPU_MMA_TRACE"] = "1"
            for _ in range(3):
                ix.search_top_k_device(q[:b], k)
            torch.cuda.synchronize()
            os.environ["FSGPU_MMA_TRACE"] = "0"
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
            for _ in range(10):
                ix.search_top_k_device(q[:b], k)
            t1.record()
            torch.cuda.synchronize()
            print(f"[untraced] rows={rows} batch={b} k={k}: {t0.elapsed_time(t1) / 10 * 1e3:.1f} us per call", file=sys.stderr)
        ix.close()
        del slab


if __name__ == "__main__":
    main()
