"""MiniLM-L6 encoder throughput on the GPU box: python tools/bench_minilm.py [batch] [tokens] [iters]
Prints queries/s, GEMM time share and achieved tensor TFLOP/s (3-product split and 1-product modes)."""
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
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    tokens = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    iters = int(sys.argv[3]) if len(sys.argv) > 3 else 10
    model = mr.make_bert(seed=3, vocab=30522)
    enc = fs.MiniLmEmbedder(mr.state_dict_numpy(model))
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(0)
    lens = rng.integers(4, tokens + 1, batch).astype(np.int32)
    ids = rng.integers(1, 30522, (batch, tokens)).astype(np.int32)
    d_ids, d_lens = torch.from_numpy(ids).to(dev), torch.from_numpy(lens).to(dev)
    print(f"# batch={batch} t_pad={tokens} rows={batch * tokens}")
    print("products  ms/batch   queries/s   gemm_ms  gemm_share  TFLOP/s(issued)  TFLOP/s(useful)")
    for products in (0, 3, 1):  # 0 = the f16 form (default)
        os.environ["FSGPU_MINILM_PRODUCTS"] = str(products)
        for _ in range(3):
            enc.embed_device(d_ids, d_lens)
        torch.cuda.synchronize()
        enc.profile_read(reset=True)
        enc.profile_enable(True)
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(iters):
            enc.embed_device(d_ids, d_lens)
        t1.record()
        torch.cuda.synchronize()
        p = enc.profile_read(reset=True)
        enc.profile_enable(False)
        ms = t0.elapsed_time(t1) / iters
        gemm_ms = p["gemm_ms"] / iters
        # the library counts flops on batch * t_pad rows; the f16 form (products 0) works on the packed rows only
        packed = products == 0 and os.environ.get("FSGPU_MINILM_PACKED", "1") != "0" and batch * tokens >= 256
        issued = p["gemm_flops"] / iters / (gemm_ms * 1e-3) / 1e12 * (float(lens.sum()) / (batch * tokens) if packed else 1.0)
        print(f"{products:8d} {ms:9.3f} {batch / (ms * 1e-3):11.0f} {gemm_ms:9.3f} {gemm_ms / ms:10.2f} "
              f"{issued:16.1f} {issued / max(products, 1):16.1f}", flush=True)
    # single-query latency (the reference quotes ~128 ms per query on one CPU core, README.md:527)
    os.environ["FSGPU_MINILM_PRODUCTS"] = "0"
    one_ids, one_len = d_ids[:1, :16].contiguous(), torch.tensor([16], dtype=torch.int32, device=dev)
    for _ in range(5):
        enc.embed_device(one_ids, one_len)
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(50):
        enc.embed_device(one_ids, one_len)
    torch.cuda.synchronize()
    print(f"single 16-token query: {(time.perf_counter() - t) / 50 * 1e3:.3f} ms")
    enc.close()


if __name__ == "__main__":
    main()
