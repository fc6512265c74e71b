"""Randomised parity soak of the batched tensor-core scan against the per-query CUDA-core scan
(both through the C ABI; the per-query path is itself oracle-pinned by tests/).  Adversarial corpora:
ascending-by-score row order, heavy duplicates, tiny / huge / mixed norms, tombstones, filters.
    python tools/soak_batched.py [n_cases] [seed]   -> exit 1 on the first mismatch
SOAK_I8=1: the int8 forms instead (FSGPU_I8_MIN_ROWS=0, k <= 32, batches up to 1024 so that the quad kernel, its
carried sample lists and skipped tiles are exercised on every corpus kind; larger corpora).
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

import frankensearch_b200 as fs  # noqa: E402


def f16_bits(a):
    return np.ascontiguousarray(a.astype(np.float16)).view(np.uint16)


def make_corpus(rng, kind, n, dim):
    x = rng.normal(size=(n, dim)).astype(np.float32)
    if kind == "unit":
        x /= np.linalg.norm(x, axis=1, keepdims=True)
    elif kind == "clustered":
        c = rng.normal(size=(8, dim)).astype(np.float32)
        x = c[rng.integers(0, 8, n)] + 0.3 * x
        x /= np.linalg.norm(x, axis=1, keepdims=True)
    elif kind == "dups":
        base = x[: max(4, n // 50)]
        x = base[rng.integers(0, len(base), n)].copy()
        x /= np.linalg.norm(x, axis=1, keepdims=True)
    elif kind == "tiny":
        x *= 1e-3
    elif kind == "huge":
        x *= 40.0
    elif kind == "mixed":
        x *= rng.uniform(1e-3, 30.0, (n, 1)).astype(np.float32)
    elif kind == "quantised":
        x = np.round(x * 4) / 4  # many exact ties
    return x


def main():
    n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 150
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    rng = np.random.default_rng(seed)
    kinds = ["unit", "clustered", "dups", "tiny", "huge", "mixed", "quantised"]
    redo_total = 0
    i8 = os.environ.get("SOAK_I8", "0") != "0"
    if i8:
        os.environ["FSGPU_I8_MIN_ROWS"] = "0"
    quad_total = 0
    for case in range(n_cases):
        kind = kinds[case % len(kinds)]
        dim = int(rng.choice([64, 128, 192, 256, 384, 512]))
        n = int(rng.choice([50, 300, 1000, 5000, 20000, 70000, 200000]))
        batch = int(rng.choice([3, 5, 17, 64, 129, 300, 700]))
        k = int(rng.choice([1, 3, 10, 16, 17, 50, 100, 200, 600]))
        if i8:
            dim = int(rng.choice([128, 256, 384, 512]))
            n = int(rng.choice([20000, 70000, 200000, 600000, 1500000]))
            batch = int(rng.choice([64, 257, 300, 512, 700, 1024]))
            k = int(rng.choice([1, 3, 10, 17, 32]))
        x = make_corpus(rng, kind, n, dim)
        q = rng.normal(size=(batch, dim)).astype(np.float32)
        if case % 3 == 0:
            q[: batch // 2] = x[rng.integers(0, n, batch // 2)] + 0.05 * q[: batch // 2]  # queries near rows
        if case % 5 == 1:  # ascending row order for query 0: the worst case for a streaming threshold
            order = np.argsort(x @ q[0])
            x = x[order]
        tomb = (rng.random(n) < 0.2) if case % 4 == 2 else None
        allow = (rng.random(n) < 0.5) if case % 6 == 3 else None
        ix = fs.GpuVectorIndex.from_f16_bits(None, f16_bits(x), tombstones=tomb)
        ix.profile_read(reset=True)
        r1, s1, c1 = ix.search_top_k_batch(q, k, filter=allow)
        prof = ix.profile_read(reset=True)
        os.environ["FSGPU_MMA_MIN_BATCH"] = "0"
        try:
            r2, s2, c2 = ix.search_top_k_batch(q, k, filter=allow)
        finally:
            del os.environ["FSGPU_MMA_MIN_BATCH"]
        ix.close()
        ok = np.array_equal(c1, c2)
        for b in range(batch):
            m = int(c2[b])
            ok = ok and np.array_equal(r1[b, :m], r2[b, :m]) and \
                np.array_equal(s1[b, :m].view(np.uint32), s2[b, :m].view(np.uint32))
        redo_total += prof["redo_queries"]
        quad_total += prof.get("quad_launches", 0)
        tag = f"case {case:3d} {kind:9s} n={n:6d} dim={dim:3d} batch={batch:3d} k={k:3d} " \
              f"mma={prof['mma_launches']} i8={prof.get('i8_launches', 0)} quad={prof.get('quad_launches', 0)} redo={prof['redo_queries']}"
        if not ok or prof["mma_launches"] < 1:
            print("MISMATCH", tag, flush=True)
            sys.exit(1)
        if case % 10 == 0 or i8:
            print("ok", tag, flush=True)
    print(f"soak ok: {n_cases} cases, {redo_total} queries re-run on the exact path, {quad_total} quad-kernel full passes")


if __name__ == "__main__":
    main()
