"""Multi-GPU parity of the row-sharded search, run under torchrun with NCCL:
every rank owns a contiguous row shard; the merged result on EVERY rank must be bit-identical to
one index over all rows (built on each rank's own GPU for the comparison).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \\
      --master-port 29617 tools/check_sharded_nccl.py [rows] [dim]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import frankensearch_b200 as fs  # noqa: E402
from frankensearch_b200.sharded import ShardedGpuIndex, shard_bounds  # noqa: E402


def synth(dev_index, lo, n, dim):
    slab = torch.empty((n, dim), dtype=torch.int16, device=torch.device("cuda", dev_index))
    fs._ffi.check(fs._ffi.lib().fsgpu_synth_rows_device(dev_index, 1, 1, lo, n, dim, 64, 0.30, slab.data_ptr(), None))
    return slab


def check_two_tier(rank, world, local, dev, rows):
    """The sharded two-tier pipeline (256-d fast tier -> 384-d quality re-score -> blend -> RRF, one
    all-gather of [keys | hits | quality]) must equal the same pipeline over ONE pair of indexes."""
    from frankensearch_b200.pipeline import DeviceLexical, DeviceTwoTierSearcher

    lo, hi = shard_bounds(rows, world, rank)
    mk = lambda seed, a, n, dim: fs.GpuVectorIndex.from_device_tensor(synth_seed(local, seed, a, n, dim), row_base=a)  # noqa: E731
    sharded = DeviceTwoTierSearcher(mk(21, lo, hi - lo, 256), mk(22, lo, hi - lo, 384))
    whole = DeviceTwoTierSearcher(mk(21, 0, rows, 256), mk(22, 0, rows, 384))
    whole._world = 1  # the reference arm of the comparison never exchanges anything
    g = torch.Generator(device="cpu").manual_seed(11)
    bad = 0
    for k, batch in ((10, 1), (10, 33), (100, 7), (1000, 2)):
        fq = torch.nn.functional.normalize(torch.randn((batch, 256), generator=g), dim=1).to(dev).contiguous()
        qq = torch.nn.functional.normalize(torch.randn((batch, 384), generator=g), dim=1).to(dev).contiguous()
        fetch = sharded.fetch_for(k)
        ids = torch.randint(0, rows + rows // 8, (batch, fetch), generator=g).to(torch.int64)
        ids = torch.where(ids < rows, ids, ids + (1 << 32)).to(dev)
        sc = torch.arange(fetch, 0, -1, dtype=torch.float32).repeat(batch, 1).to(dev).contiguous()
        for lex in (None, DeviceLexical(ids, sc)):
            a = sharded.search_device(fq, qq, k, lex)
            got = [x.clone() for x in (a.fast_hits, a.fast_counts, a.quality_scores, a.quality_present, a.initial,
                                       a.initial_counts, a.blended, a.blended_counts, a.refined, a.refined_counts)]
            w = whole.search_device(fq, qq, k, lex)
            want = (w.fast_hits, w.fast_counts, w.quality_scores, w.quality_present, w.initial, w.initial_counts,
                    w.blended, w.blended_counts, w.refined, w.refined_counts)
            torch.cuda.synchronize()
            names = "fast_hits fast_counts quality_scores quality_present initial initial_counts blended blended_counts refined refined_counts".split()
            for nm, x, y in zip(names, got, want):
                if nm == "quality_scores":  # slots past the count are unspecified in the single-index arm
                    m = torch.arange(fetch, device=dev)[None, :] < a.fast_counts[:, None]
                    x, y = torch.where(m, x, 0), torch.where(m, y, 0)
                if not torch.equal(x.view(torch.uint8) if x.dtype == torch.float32 else x,
                                   y.view(torch.uint8) if y.dtype == torch.float32 else y):
                    bad += 1
                    print(f"rank {rank}: TWO-TIER MISMATCH {nm} k={k} batch={batch} lexical={lex is not None}", flush=True)
    return bad


def synth_seed(dev_index, seed, lo, n, dim):
    slab = torch.empty((n, dim), dtype=torch.int16, device=torch.device("cuda", dev_index))
    fs._ffi.check(fs._ffi.lib().fsgpu_synth_rows_device(dev_index, 1, seed, lo, n, dim, 64, 0.30, slab.data_ptr(), None))
    return slab


def main():
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 400_000
    dim = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    lo, hi = shard_bounds(rows, world, rank)
    shard = fs.GpuVectorIndex.from_device_tensor(synth(local, lo, hi - lo, dim), row_base=lo)
    whole = fs.GpuVectorIndex.from_device_tensor(synth(local, 0, rows, dim))
    sharded = ShardedGpuIndex(shard)
    g = torch.Generator(device="cpu").manual_seed(7)  # same queries on every rank
    q_all = torch.randn((300, dim), generator=g)
    q_all = (q_all / q_all.norm(dim=1, keepdim=True)).to(dev).contiguous()
    bad = 0
    for k in (1, 10, 100):
        for batch in (1, 2, 5, 64, 300):
            q = q_all[:batch].contiguous()
            for rep in range(2):  # the second call reuses the cached all-gather buffers
                keys, hits, counts = sharded.search_top_k_device(q, k)
                wkeys, whits, wcounts = whole.search_top_k_device(q, k)
                torch.cuda.synchronize()
                ok = torch.equal(keys, wkeys) and torch.equal(hits, whits) and torch.equal(counts, wcounts)
                if not ok:
                    bad += 1
                    print(f"rank {rank}: MISMATCH k={k} batch={batch} rep={rep}", flush=True)
    # R row shards x 2 query groups (rank = g*R + r): same answer, ragged and one-query batches included
    if world % 2 == 0:
        from frankensearch_b200.sharded import grid_position

        r_shards, r, _g = grid_position(world, rank, 2)
        glo, ghi = shard_bounds(rows, r_shards, r)
        gshard = fs.GpuVectorIndex.from_device_tensor(synth(local, glo, ghi - glo, dim), row_base=glo)
        grid = ShardedGpuIndex(gshard, query_groups=2)
        for k in (1, 10, 100):
            for batch in (1, 2, 5, 64, 300):
                q = q_all[:batch].contiguous()
                for rep in range(2):
                    keys, hits, counts = grid.search_top_k_device(q, k)
                    wkeys, whits, wcounts = whole.search_top_k_device(q, k)
                    torch.cuda.synchronize()
                    if not (torch.equal(keys, wkeys) and torch.equal(hits, whits) and torch.equal(counts, wcounts)):
                        bad += 1
                        print(f"rank {rank}: GRID MISMATCH k={k} batch={batch} rep={rep}", flush=True)
    bad += check_two_tier(rank, world, local, dev, rows)
    t = torch.tensor([bad], device=dev)
    dist.all_reduce(t)
    if rank == 0:
        print("SHARDED_NCCL_PARITY", "OK" if t.item() == 0 else f"FAILED ({t.item()} mismatches)", f"world={world}")
    dist.destroy_process_group()
    sys.exit(0 if t.item() == 0 else 1)


if __name__ == "__main__":
    main()
