"""End-to-end latency of the semantic tier on one GPU (the shape of BASELINE.json configs[4]):
token ids -> MiniLM-L6 encode -> exact f16 cosine scan -> top-k, single queries back to back and
1024-query batches, everything device-resident between the stages.
    python tools/bench_pipeline.py [rows] [k] [n_single]
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

import frankensearch_b200 as fs  # noqa: E402
import minilm_ref as mr  # noqa: E402


def main():
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
    k = int(sys.argv[2]) if len(sys.argv) > 2 else 100
    n_single = int(sys.argv[3]) if len(sys.argv) > 3 else 300
    dev = torch.device("cuda", 0)
    slab = torch.empty((rows, 384), dtype=torch.int16, device=dev)
    fs._ffi.check(fs._ffi.lib().fsgpu_synth_rows_device(0, 1, 1, 0, rows, 384, 64, 0.30, slab.data_ptr(), None))
    ix = fs.GpuVectorIndex.from_device_tensor(slab)
    enc = fs.MiniLmEmbedder(mr.state_dict_numpy(mr.make_bert(seed=3, vocab=30522)))
    rng = np.random.default_rng(0)

    def make(batch):
        lens = rng.integers(4, 33, batch).astype(np.int32)
        ids = rng.integers(1, 30522, (batch, int(lens.max()))).astype(np.int32)
        return torch.from_numpy(ids).pin_memory(), torch.from_numpy(lens).pin_memory()

    def run(ids_h, lens_h):
        d_ids, d_lens = ids_h.to(dev, non_blocking=True), lens_h.to(dev, non_blocking=True)
        q = enc.embed_device(d_ids, d_lens)
        keys, hits, counts = ix.search_top_k_device(q, k)
        out = hits.cpu()  # the caller's result: device -> host
        return out

    print(f"# rows={rows} dim=384 k={k}; host ids -> MiniLM encode -> exact scan -> host hits")
    singles = [make(1) for _ in range(n_single)]
    for s in singles[:10]:
        run(*s)
    torch.cuda.synchronize()
    lat = []
    for s in singles:
        t = time.perf_counter()
        run(*s)
        lat.append((time.perf_counter() - t) * 1e3)
    lat = np.sort(np.array(lat))
    print(f"single query  : p50 {np.percentile(lat, 50):.3f} ms  p99 {np.percentile(lat, 99):.3f} ms  "
          f"mean {lat.mean():.3f} ms  ({n_single} queries back to back)")
    for batch in (64, 1024):
        b = make(batch)
        for _ in range(3):
            run(*b)
        torch.cuda.synchronize()
        ts = []
        for _ in range(10):
            t = time.perf_counter()
            run(*b)
            ts.append((time.perf_counter() - t) * 1e3)
        ts = np.array(ts)
        print(f"batch {batch:5d}   : median {np.median(ts):.3f} ms per batch -> {batch / np.median(ts) * 1e3:.0f} queries/s "
              f"(p99 {np.percentile(ts, 99):.3f} ms)")
    ix.close()
    enc.close()


if __name__ == "__main__":
    main()
