"""Single-query latency through the host API (pinned-free numpy query in, hits out), f16 scan vs
the int8 pass-1 form.  python tools/bench_single_query.py [rows] [dim] [k] [batch]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import frankensearch_b200 as fs  # noqa: E402


def main():
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
    dim = int(sys.argv[2]) if len(sys.argv) > 2 else 384
    k = int(sys.argv[3]) if len(sys.argv) > 3 else 10
    batch = int(sys.argv[4]) if len(sys.argv) > 4 else 1
    os.environ["FSGPU_MMA_I8"] = "1"
    dev = torch.device("cuda", 0)
    slab = torch.empty((rows, dim), dtype=torch.int16, device=dev)
    fs._ffi.check(fs._ffi.lib().fsgpu_synth_rows_device(0, 1, 1, 0, rows, dim, 64, 0.30, slab.data_ptr(), None))
    ix = fs.GpuVectorIndex.from_device_tensor(slab)
    rng = np.random.default_rng(1)
    qs = rng.standard_normal((64 * batch, dim)).astype(np.float32)
    qs /= np.linalg.norm(qs, axis=1, keepdims=True)
    print(f"# rows={rows} dim={dim} k={k} batch={batch} int8_ready={ix._L.fsgpu_index_int8_ready(ix._h)}")
    ref = None
    for mode in ("0", "1"):
        os.environ["FSGPU_MMA_I8"] = mode
        for i in range(5):
            ix.search_top_k_batch(qs[i * batch:(i + 1) * batch], k)
        ix.profile_read(reset=True)
        ix.profile_enable(True)
        lat = []
        out = []
        for i in range(64):
            t0 = time.perf_counter()
            r = ix.search_top_k_batch(qs[i * batch:(i + 1) * batch], k)
            lat.append((time.perf_counter() - t0) * 1e3)
            out.append(r)
        p = ix.profile_read(reset=True)
        ix.profile_enable(False)
        lat = np.array(lat)
        print(f"FSGPU_MMA_I8={mode}: p50 {np.percentile(lat, 50):.3f} ms  p99 {np.percentile(lat, 99):.3f} ms  "
              f"scan kernel {p['scan_ms'] / max(p['scan_launches'], 1):.3f} ms  "
              f"({p['scan_bytes'] / max(p['scan_launches'], 1) / (p['scan_ms'] / max(p['scan_launches'], 1) * 1e-3) / 1e9:.0f} GB/s)")
        if ref is None:
            ref = out
        else:
            same = all(np.array_equal(a[0], b[0]) and np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32))
                       for a, b in zip(ref, out))
            print("identical results:", same)
    ix.close()


if __name__ == "__main__":
    main()
