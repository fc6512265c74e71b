#!/usr/bin/env python
"""bench.py — headline benchmark of the semantic-tier hot path (BASELINE.json configs[2]).

Workload: 10 M docs x 384-dim f16 slab, a batch of 1024 queries, exact fused cosine + top-k.
A "step" is one pass of the hot path over one 1024-query batch.  At N > 1 the corpus is
row-sharded across the ranks (strong scaling, BASELINE configs[3]); every rank scans its shard
for all queries and one NCCL all-gather of the per-rank top-k keys feeds the merge kernel.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for how each field is computed.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "queries_per_sec_f16_cosine_topk_10Mx384"
UNIT = "queries/s"


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=20)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--rows", type=int, default=10_000_000)
    p.add_argument("--dim", type=int, default=384)
    p.add_argument("--batch", type=int, default=1024)
    p.add_argument("--k", type=int, default=10)
    p.add_argument("--cpu-rows", type=int, default=0, help="rows of the CPU sample (0 = same as --rows)")
    p.add_argument("--cpu-queries", type=int, default=4, help="queries per CPU step (a sample of the batch)")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--config", default="all", choices=["all", "3", "2", "4", "5"],
                   help="3 = the headline only (configs[2]); all = headline + sub-records for configs[1], [3], [4]")
    p.add_argument("--rows5", type=int, default=50_000_000, help="rows of configs[4] (50 M)")
    p.add_argument("--queries5", type=int, default=1000, help="queries of the configs[4] latency run")
    p.add_argument("--query-groups", type=int, default=0,
                   help="N > 1: ranks = (N / Q) row shards x Q query groups (frankensearch_b200/sharded.py); default 1 = "
                        "plain row shards: at 8 GPUs 2 x 4 measured 0.658 ms per step against 0.640 for 8 row shards "
                        "(profiles/r02_layout_rows_x_query_groups.md)")
    p.add_argument("--hbm-batch", type=int, default=128,
                   help="also time the HBM-bound regime with this many queries per pass (0 = skip)")
    return p.parse_args()


def query_groups_for(a, n_gpus):
    if n_gpus < 2:
        return 1
    q = a.query_groups if a.query_groups > 0 else 1
    return q if n_gpus % q == 0 else 1


def workload_config(a, n_gpus):
    qg = query_groups_for(a, n_gpus)
    return {
        "workload": f"configs[2]: {a.rows} docs x {a.dim}-dim f16, batch {a.batch} queries, exact cosine top-{a.k}",
        "rows": a.rows, "dim": a.dim, "batch": a.batch, "k": a.k,
        "corpus": "clustered (64 centroids, noise 0.30) — reference bench generator fsvi_int8_two_pass.rs:199-231",
        "storage": "f16 slab (+ int8 codes of it for the candidate pass), f32 query; tensor-core candidate pass (int8 x int8 -> s32, or f16 x f16 -> f32 with FSGPU_MMA_I8=0) + exact f16 re-scoring with the reference accumulation tree (bit-exact results)",
        "sharding": ("single GPU" if n_gpus == 1 else
                     f"rows sharded over {n_gpus} ranks, one all-gather of top-k keys" if qg == 1 else
                     f"{n_gpus // qg} row shards x {qg} query groups (rank g*R + r: rows [r*N/R, (r+1)*N/R), queries "
                     f"[g*B/Q, (g+1)*B/Q)), one all-gather of top-k keys, one merge per group; the HBM-regime "
                     f"sub-records and `row_shards_only` use {n_gpus} plain row shards"),
        "l2": "inputs larger than L2 (slab per GPU >> 126 MB); no explicit flush",
    }


# ── CPU arm: the oracle restatement of the reference scan on the host cores ──────────────────
def cpu_arm(a, steps, warmup, slab=None, queries=None):
    """Times oracle.fs_oracle.search_top_k (reference restatement: AVX2+F16C dot, 1024-row chunks,
    per-chunk heaps, serial merge; all host threads) on a bounded sample: the full-size corpus,
    `cpu_queries` queries of the batch per step, queries back to back (the reference's
    production model, benches/batched_query_scan.rs:108-126)."""
    from oracle import fs_oracle as fo

    rows = a.cpu_rows or a.rows
    threads = fo.host_threads()
    t0 = time.perf_counter()
    if slab is None or slab.shape[0] != rows:
        slab, _ = fo.synth_rows(1, 1, 0, rows, a.dim, threads=threads)
    gen_s = time.perf_counter() - t0
    if queries is None:
        queries = [fo.clustered_query(q, a.dim) for q in range(a.cpu_queries)]

    def step():
        for q in queries:
            fo.search_top_k(slab, q, a.k, threads=threads)

    for _ in range(warmup):
        step()
    times = []
    for _ in range(steps):
        t = time.perf_counter()
        step()
        times.append(time.perf_counter() - t)
    per_query = sum(times) / (steps * len(queries))
    per_query_full = per_query * (a.rows / rows)  # the scan is linear in rows
    return dict(value=1.0 / per_query_full, per_query_ms=per_query_full * 1e3, threads=threads, rows=rows,
                gen_s=gen_s, ms_per_step=1e3 * sum(times) / steps,
                sample=f"{rows} rows x {a.dim} (of {a.rows}), {len(queries)} of {a.batch} queries per step, "
                       f"{steps} steps after {warmup} warm-ups" + ("" if rows == a.rows else ", scaled linearly in rows"))


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_arm(a, a.steps, a.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(a, a.gpus),
        "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["threads"], "kind": "port",
                         "sample": r["sample"], "cpu": cpu_model(), "per_query_ms": r["per_query_ms"]},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference = the oracle port of the reference's SIMD CPU scan (the Rust workspace cannot be "
                "built in this image: no cargo/rustc); a step scans the corpus for a sample of the batch",
    }
    print(json.dumps(line), flush=True)


# ── clocks ────────────────────────────────────────────────────────────────────────────────────
class ClockSampler:
    """SM clock + throttle reasons sampled every ~5 ms through NVML while a timed region runs
    (nvidia-smi -lms cannot sample a 40 ms region); falls back to nvidia-smi if NVML is missing."""
    REASONS = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20,
               "hw_thermal_slowdown": 0x40, "hw_power_brake_slowdown": 0x80}

    def __init__(self, gpu_index: int):
        self.sm, self.mask, self.max_mhz = [], 0, None
        self.stop_flag = threading.Event()
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index(gpu_index))
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nvml = None
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    @staticmethod
    def _physical_index(i):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            parts = [p.strip() for p in vis.split(",") if p.strip()]
            if i < len(parts) and parts[i].isdigit():
                return int(parts[i])
        return i

    def _run(self):
        if self.nvml is not None:
            n = self.nvml
            while not self.stop_flag.is_set():
                try:
                    self.sm.append(float(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)))
                    self.mask |= int(n.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    try:
                        self.mask |= int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                    except Exception:
                        pass
                time.sleep(0.004)
            return
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm,"
                                      "clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_power_cap,"
                                      "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.hw_thermal_slowdown",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [x.strip() for x in out.splitlines()[0].split(",")]
                self.sm.append(float(parts[0]))
                self.max_mhz = float(parts[1])
                for name, v in zip(("hw_slowdown", "sw_power_cap", "sw_thermal_slowdown", "hw_thermal_slowdown"), parts[2:6]):
                    if v.lower().startswith("active"):
                        self.mask |= self.REASONS[name]
            except Exception:
                return

    def stop(self):
        self.stop_flag.set()
        self.thread.join(timeout=6)
        reasons = sorted(name for name, bit in self.REASONS.items() if self.mask & bit)
        return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_min_mhz": min(self.sm) if self.sm else None,
                "sm_max_mhz": self.max_mhz, "samples": len(self.sm), "reasons": reasons,
                "reasons_mask": hex(self.mask),  # raw NVML event-reason bits OR-ed over the samples (0x1 idle, 0x2 app clocks, 0x4 sw power cap)
                "source": "nvml" if self.nvml is not None else "nvidia-smi"}


# ── our arm ───────────────────────────────────────────────────────────────────────────────────
class Ctx:
    """Per-process state shared by the headline run and the extra configurations."""

    def __init__(self, a):
        import torch
        import torch.distributed as dist

        self.a = a
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device — the product has no CPU fallback (use --impl reference "
                             "for the CPU arm)")
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.peaks = {}
        try:
            self.peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except (OSError, ValueError):
            pass
        self.peak_gbs = float(self.peaks.get("hbm_gbs", 6650.0))
        self.peak_src = ("measured (MEASURED_PEAKS.json)" if "hbm_gbs" in self.peaks
                         else "fallback (B200_PROFILING.md: 6650 GB/s, 1590 TFLOP/s)")
        self.host_slabs = {}  # (seed, rows, dim) -> oracle-generated host copy (rank 0, parity + CPU leg)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def max_over_ranks(self, x: float) -> float:
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def synth_shard(self, seed, rows, dim, shards=None, shard=None):
        """This rank's contiguous row shard (default: shard `rank` of `world`) of a clustered corpus, generated
        on the device by the reference's bench generator (bit-identical to the oracle's fso_synth_rows)."""
        import frankensearch_b200 as fs
        from frankensearch_b200.sharded import shard_bounds

        lo, hi = shard_bounds(rows, self.world if shards is None else shards, self.rank if shard is None else shard)
        slab = self.torch.empty((hi - lo, dim), dtype=self.torch.int16, device=self.dev)
        fs._ffi.check(fs._ffi.lib().fsgpu_synth_rows_device(self.local_rank, 1, seed, lo, hi - lo, dim, 64, 0.30,
                                                            slab.data_ptr(), None))
        return fs.GpuVectorIndex.from_device_tensor(slab, row_base=lo), lo, hi

    def host_slab(self, seed, rows, dim):
        from oracle import fs_oracle as fo

        key = (seed, rows, dim)
        if key not in self.host_slabs:
            self.host_slabs[key] = fo.synth_rows(1, seed, 0, rows, dim, threads=fo.host_threads())[0]
        return self.host_slabs[key]


def _time_steps(ctx, step, steps, warmup):
    """`warmup` untimed + `steps` timed calls; device time (CUDA events, max over ranks) per step."""
    torch = ctx.torch
    for _ in range(warmup):
        step()
    ctx.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        out = step()
    e1.record()
    ctx.barrier()
    return ctx.max_over_ranks(e0.elapsed_time(e1)) / steps, out


def _time_wall(ctx, step, steps, warmup):
    """End to end through the host: wall clock around `steps` calls that each end with the result in
    host memory (max over ranks)."""
    for _ in range(warmup):
        step()
    ctx.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    ctx.torch.cuda.synchronize(ctx.dev)
    return ctx.max_over_ranks(time.perf_counter() - t0) / steps * 1e3


def _hits_parity(got_rows, got_scores, o_rows, o_scores):
    rows_equal = [int(r) for r in got_rows] == [int(r) for r in o_rows]
    bits_equal = np.array_equal(np.asarray(got_scores, dtype=np.float32).view(np.uint32),
                                np.asarray(o_scores, dtype=np.float32).view(np.uint32))
    return rows_equal, bits_equal


def run_ours(a):
    ctx = Ctx(a)
    line = run_headline(ctx)
    if a.config in ("all", "2", "4", "5"):
        extras = {}
        for name, fn in (("2", bench_config2), ("4", bench_config4), ("5", bench_config5)):
            if a.config not in ("all", name):
                continue
            try:
                rec = fn(ctx)
            except Exception as e:  # an extra configuration must never take the headline line down
                rec = {"error": f"{type(e).__name__}: {e}"}
            ctx.torch.cuda.empty_cache()
            if rec is not None:
                extras["config" + name] = rec
        if line is not None:
            line["configs"] = extras
    if ctx.rank == 0 and line is not None:
        print(json.dumps(line), flush=True)
    if ctx.world > 1:
        ctx.dist.destroy_process_group()


def run_headline(ctx):
    a, torch, dev, world, rank, local_rank = ctx.a, ctx.torch, ctx.dev, ctx.world, ctx.rank, ctx.local_rank
    import frankensearch_b200 as fs
    from frankensearch_b200.sharded import ShardedGpuIndex, grid_position, query_block

    # layout at N > 1: R row shards x Q query groups for the headline batch (frankensearch_b200/sharded.py); the
    # HBM-bound sub-records (1 and 128 queries per pass) run on N plain row shards
    qg = query_groups_for(a, world)
    r_shards, r_shard, q_group = grid_position(world, rank, qg)
    ix, lo, hi = ctx.synth_shard(1, a.rows, a.dim)
    sharded = ShardedGpuIndex(ix) if world > 1 else None

    # queries: the reference generator's clustered queries, built on the host (tiny)
    q_host = torch.empty((a.batch, a.dim), dtype=torch.float32).pin_memory()
    q_np = q_host.numpy()
    _fill_queries(q_np, a.dim)
    d_queries = q_host.to(dev, non_blocking=False)

    def search(q):
        if sharded is not None:
            return sharded.search_top_k_device(q, a.k)
        return ix.search_top_k_device(q, a.k, want_hits=True)

    def profiled(q, env=None):
        """Average event-timed scan launch for repeated searches of `q` (optionally under a switch)."""
        old = {}
        for k_, v_ in (env or {}).items():
            old[k_] = os.environ.get(k_)
            os.environ[k_] = v_
        try:
            for _ in range(3):
                search(q)
            ctx.barrier()
            ix.profile_read(reset=True)
            ix.profile_enable(True)
            ms, _ = _time_steps(ctx, lambda: search(q), a.steps, 0)
            p = ix.profile_read(reset=True)
            ix.profile_enable(False)
        finally:
            for k_, v_ in old.items():
                if v_ is None:
                    os.environ.pop(k_, None)
                else:
                    os.environ[k_] = v_
        n = max(p["scan_launches"], 1)
        lm = p["scan_ms"] / n
        bpl = p["scan_bytes"] / n
        kern = ("mma_scan_quad_kernel" if p.get("quad_launches") else
                "mma_scan_pair_kernel" if p.get("pair_launches") else
                "mma_scan_kernel" if p["mma_launches"] else
                "scan_i8_all_kernel" if p["i8_launches"] else "scan_topk_fast_kernel")
        if p["mma_launches"]:
            kern += "<int8>" if p["i8_launches"] else "<f16>"
        return {"batch": int(q.shape[0]), "kernel": kern, "avg_launch_ms": lm, "bytes_per_launch": bpl,
                "achieved_gbs": bpl / (lm * 1e-3) / 1e9 if lm else 0.0, "ms_per_step": ms,
                "frac": (bpl / (lm * 1e-3) / 1e9 / ctx.peak_gbs) if lm else None, "peak_gbs": ctx.peak_gbs}

    # the north-star's HBM-bound regimes, every N: exact f16 scan of one query (CUDA cores), f16 and int8
    # forms of the batched kernel with one 128-query block per corpus pass
    hbm = {}

    def run_hbm():
        if a.hbm_batch <= 0:
            return
        q1 = d_queries[:1].contiguous()
        qb = d_queries[: min(a.hbm_batch, a.batch)].contiguous()
        hbm["f16_single_query"] = profiled(q1)
        hbm["f16_batch128"] = profiled(qb, {"FSGPU_MMA_I8": "0"})
        hbm["int8_batch128"] = profiled(qb)
        for rec in hbm.values():
            rec["algorithmic_bytes"] = "rows*dim*2 (f16 slab)" if "int8" not in rec["kernel"] else "rows*dim (int8 codes)"

    # the sub-records run AFTER the headline's timed region (the headline is measured first, from the state the index
    # build leaves the GPU in); only the rows x query-groups layout needs them before its index replaces the row shards
    if qg > 1:
        run_hbm()

    # the int8 issue-rate peak of this GPU, before anything has heated it (and again after the run: the larger one is
    # the denominator of `frac`)
    peak_i8_cold = measure_tensor_peak(fs, local_rank, 1) if rank == 0 else None
    row_shards_only = None
    if qg > 1:
        # the same batch on N plain row shards, for the record; then the R x Q layout takes over
        for _ in range(3):
            search(d_queries)
        ms_rows, _ = _time_steps(ctx, lambda: search(d_queries), a.steps, 0)
        row_shards_only = {"layout": f"{world} row shards", "ms_per_step": ms_rows, "queries_per_s": a.batch / (ms_rows / 1e3)}
        ix.close()
        ix, lo, hi = ctx.synth_shard(1, a.rows, a.dim, shards=r_shards, shard=r_shard)
        sharded = ShardedGpuIndex(ix, query_groups=qg)
    my_q = query_block(a.batch, qg, q_group)
    for _ in range(max(a.warmup, 3)):
        search(d_queries)
    ctx.barrier()
    ix.profile_read(reset=True)
    ix.profile_enable(True)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ms_per_step, out = _time_steps(ctx, lambda: search(d_queries), a.steps, 0)
    prof = ix.profile_read(reset=True)
    ix.profile_enable(False)
    clocks = sampler.stop() if sampler else None
    elapsed_ms = ms_per_step * a.steps
    value = a.batch / (ms_per_step / 1e3)

    # e2e: the host-buffer C-ABI call (fsgpu_search_top_k): pinned host queries -> device, search,
    # hits -> host, every step inside the timed region
    hits_host = torch.empty((a.batch, a.k, 2), dtype=torch.int32).pin_memory()
    counts_host = torch.empty(a.batch, dtype=torch.int32).pin_memory()

    def step_e2e():
        if sharded is None:
            fs._ffi.check(fs._ffi.lib().fsgpu_search_top_k(ix.handle, q_host.data_ptr(), a.batch, a.k, a.dim,
                                                           hits_host.data_ptr(), counts_host.data_ptr()))
        else:
            dq = q_host.to(dev, non_blocking=True)
            _, mh, mc = sharded.search_top_k_device(dq, a.k)
            hits_host.copy_(mh, non_blocking=True)
            counts_host.copy_(mc, non_blocking=True)
            torch.cuda.synchronize(dev)

    e2e_ms = _time_wall(ctx, step_e2e, a.steps, 1)
    e2e_value = a.batch / (e2e_ms / 1e3)

    # sanity: the timed result is a real answer (sorted keys, k hits per query)
    keys = out[0]
    assert bool((keys[:, :-1] > keys[:, 1:]).all().item()) if a.k > 1 else True, "result keys are not sorted"
    got_hits = out[1].cpu().numpy()
    e2e_hits = hits_host.numpy().copy()
    if qg == 1:
        run_hbm()  # (after the timed results have been copied out)

    line = None
    if rank == 0:
        peak_tf = float(ctx.peaks.get("bf16_tflops", 1590.0))
        peak_tf_sus = float(ctx.peaks.get("bf16_tflops_sustained", 1400.0))
        peak_src = ctx.peak_src
        scan_launches = max(prof["scan_launches"], 1)
        avg_ms = prof["scan_ms"] / scan_launches
        bytes_per_launch = prof["scan_bytes"] / scan_launches
        hbm_achieved = bytes_per_launch / (avg_ms * 1e-3) / 1e9 if avg_ms > 0 else 0.0
        i8 = prof.get("i8_launches", 0) > 0
        # DRAM traffic of one launch comes from an ncu capture (not measurable in-process): only quoted
        # for the configuration that capture was taken on (one GPU, 10 M x 384, batch 1024)
        traffic, traffic_src = None, None
        if world == 1 and a.rows == 10_000_000 and a.dim == 384 and a.batch == 1024:
            tp = os.path.join(ROOT, "profiles", "scan_kernel_traffic_i8.json" if i8 else "scan_kernel_traffic.json")
            try:
                tj = json.load(open(tp))
                traffic, traffic_src = tj.get("dram_bytes_per_launch"), tj.get("source", os.path.basename(tp))
            except (OSError, ValueError):
                pass
        if prof["mma_launches"]:
            # a 1024-query batch is a dense [B,D]x[D,N] contraction: tensor-bound
            flops = prof["mma_flops"] / prof["mma_launches"]
            achieved_tf = flops / (avg_ms * 1e-3) / 1e12 if avg_ms > 0 else 0.0
            form = "quad" if prof.get("quad_launches") else "pair" if prof.get("pair_launches") else "single-CTA"
            kernel = {"quad": "mma_scan_quad_kernel (tcgen05.mma.cta_group::2 kind::i8, two 128-query blocks per CTA, "
                              "M=256 queries x N=256 rows per MMA, K=dim; exact: f16 re-score of the candidate superset)",
                      "pair": f"mma_scan_pair_kernel<{'int8' if i8 else 'f16'}> (tcgen05.mma.cta_group::2, M=256 queries x "
                              "N=256 rows per CTA pair, K=dim)",
                      "single-CTA": f"mma_scan_kernel<{'int8' if i8 else 'f16'}> (tcgen05.mma.cta_group::1, M=128 x N=128)"}[form]
            tensor_peak, tensor_peak_src = peak_tf, peak_src + " bf16 cuBLAS burst"
            if i8:
                # the int8 tensor peak is MEASURED on this GPU by a bare tcgen05.mma kind::i8 issue loop
                # (fsgpu_measure_tensor_peak: no loads, no epilogue); no int8 figure is in MEASURED_PEAKS.json
                m_end = measure_tensor_peak(fs, local_rank, 1)
                m = max(x for x in (peak_i8_cold, m_end, 0.0) if x is not None)
                if m:
                    tensor_peak, tensor_peak_src = m, ("measured here: bare tcgen05.mma.cta_group::2 kind::i8 issue loop (fsgpu_measure_tensor_peak), "
                                                       f"the larger of before the run ({peak_i8_cold and round(peak_i8_cold, 1)}) and after it "
                                                       f"({m_end and round(m_end, 1)} TOP/s)")
                else:
                    tensor_peak, tensor_peak_src = 2.0 * peak_tf, peak_src + "; int8 peak unmeasured: 2 x bf16 figure"
            roofline = {"bound": "tensor", "achieved": achieved_tf, "peak": tensor_peak, "unit": "TFLOP/s",
                        "frac": achieved_tf / tensor_peak if tensor_peak else None, "traffic": traffic,
                        "traffic_source": traffic_src, "kernel": kernel, "flops_per_launch": flops,
                        "ops": "int8 multiply-add = 2 ops" if i8 else "f16 multiply-add = 2 flops",
                        "peak_source": tensor_peak_src,
                        "bf16_cublas_peak": peak_tf, "bf16_cublas_sustained": peak_tf_sus,
                        "ncu_pipe_tensor_active_pct": _ncu_tensor_active(i8),
                        "hbm_gbs_same_launch": hbm_achieved, "hbm_frac_same_launch": hbm_achieved / ctx.peak_gbs}
        else:
            roofline = {"bound": "hbm", "achieved": hbm_achieved, "peak": ctx.peak_gbs, "unit": "GB/s",
                        "frac": hbm_achieved / ctx.peak_gbs if ctx.peak_gbs else None, "traffic": traffic,
                        "traffic_source": traffic_src, "kernel": "scan_topk_fast_kernel", "peak_source": peak_src}
        roofline.update({"launches_timed": prof["scan_launches"], "avg_launch_ms": avg_ms,
                         "bytes_per_launch": bytes_per_launch,
                         "queries_per_launch": (my_q[1] - my_q[0]) * a.steps / scan_launches,
                         "scan_share_of_step": prof["scan_ms"] / elapsed_ms if elapsed_ms else None,
                         "redo_queries": prof["redo_queries"]})
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps,
            "warmup": max(a.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(a, world),
            "roofline": roofline,
            "roofline_hbm": hbm,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": a.batch * a.dim * 4,
                    "d2h_bytes_per_step": a.batch * a.k * 8 + a.batch * 4},
            "gpu_launches": prof["scan_launches"] + prof["merge_launches"] + prof["other_launches"],
            "clocks": clocks,
        }
        if row_shards_only is not None:
            line["row_shards_only"] = row_shards_only
    # CPU leg + oracle parity of the timed result.  The CPU leg scans the SAME corpus with the SAME first
    # queries, so the comparison is free at N = 1; at N > 1 rank 0 runs the oracle once for the parity alone.
    if rank == 0 and not a.no_cpu_baseline:
        from oracle import fs_oracle as fo

        nq = max(1, min(a.cpu_queries, a.batch))
        # the first and the last queries of the batch: both query groups of the R x Q layout are checked
        par_idx = sorted(set(list(range((nq + 1) // 2)) + [a.batch - 1 - i for i in range(nq // 2)]))
        host = ctx.host_slab(1, a.rows, a.dim)
        if world == 1:
            r = cpu_arm(a, steps=2, warmup=1, slab=host, queries=[q_np[i] for i in par_idx])
            line["cpu_baseline"] = {"value": r["value"], "unit": UNIT, "cores": r["threads"], "kind": "port",
                                    "sample": r["sample"], "cpu": cpu_model(), "per_query_ms": r["per_query_ms"]}
        rows_ok = bits_ok = e2e_ok = True
        for b in par_idx:
            o_rows, o_scores = fo.search_top_k(host, q_np[b], a.k, threads=fo.host_threads())
            re_, be_ = _hits_parity(got_hits[b, :, 0].view(np.uint32), got_hits[b, :, 1].view(np.float32), o_rows, o_scores)
            r2, b2 = _hits_parity(e2e_hits[b, :, 0].view(np.uint32), e2e_hits[b, :, 1].view(np.float32), o_rows, o_scores)
            rows_ok, bits_ok, e2e_ok = rows_ok and re_, bits_ok and be_, e2e_ok and r2 and b2
        line["parity"] = {"queries": len(par_idx), "query_indices": par_idx, "rows_equal": rows_ok, "score_bits_equal": bits_ok,
                          "e2e_result_equal": e2e_ok, "against": "oracle fs_oracle.search_top_k on the identical corpus",
                          "checked": "the merged result of the last timed step" if world > 1 else "the result of the last timed step"}
    ctx.barrier()
    ix.close()
    return line


def _ncu_tensor_active(i8):
    """sm__pipe_tensor_cycles_active of the committed ncu capture of the headline kernel (context for
    `frac`: a profiler number is never a bench value)."""
    for name in (("r02_mma_quad_i8_b1024_v3_carry_ncu.json", "r02_mma_quad_i8_b1024_v2_ncu.json") if i8 else
                 ("r02_mma_pair_f16_b1024_paced_ncu.json", "r01_mma_pair_b1024_ncu.json")):
        try:
            j = json.load(open(os.path.join(ROOT, "profiles", name)))
            best = None
            for l in j["launches"]:
                m = l.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")
                d = l.get("gpu__time_duration.sum")
                if m and d and (best is None or float(d["value"].replace(",", "")) > best[0]):
                    best = (float(d["value"].replace(",", "")), float(m["value"].replace(",", "")))
            if best:
                return {"value": best[1], "source": "profiles/" + name}
        except (OSError, ValueError, KeyError):
            continue
    return None


def measure_tensor_peak(fs, device, kind):
    """TOP/s of a bare tcgen05.mma issue loop on this GPU (kind 0 = f16, 1 = int8); None if unavailable."""
    import ctypes as C

    L = fs._ffi.lib()
    if not hasattr(L, "fsgpu_measure_tensor_peak"):
        return None
    out = C.c_double(0.0)
    if L.fsgpu_measure_tensor_peak(device, kind, 50, C.byref(out)) != 0:
        return None
    return out.value / 1e12


# ── extra configurations of BASELINE.json (sub-records of the same JSON line) ──────────────────
def _lexical_from_hits(ctx, hits, n_docs, n_lex, seed):
    """Precomputed BM25 lists in the style of frankensearch/benches/search_bench.rs:236-251 (scores
    (n - i) as f32, rank order): about half of each list are the query's own semantic candidates (every
    other one), the rest other documents, ~1 in 9 of those without a vector row (id >= 2^32)."""
    torch = ctx.torch
    from frankensearch_b200.pipeline import DeviceLexical

    g = torch.Generator(device=ctx.dev).manual_seed(seed)
    rows = hits[..., 0].to(torch.int64) & 0xFFFFFFFF
    b = rows.shape[0]
    sem = rows[:, ::2][:, : n_lex // 2]
    rnd = torch.randint(0, n_docs + n_docs // 8, (b, n_lex - sem.shape[1]), generator=g, device=ctx.dev)
    rnd = torch.where(rnd < n_docs, rnd, rnd + (1 << 32))
    ids = torch.cat([sem, rnd], dim=1)
    perm = torch.argsort(torch.rand((b, n_lex), generator=g, device=ctx.dev), dim=1)
    ids = torch.gather(ids, 1, perm).contiguous()
    scores = torch.arange(n_lex, 0, -1, dtype=torch.float32, device=ctx.dev).repeat(b, 1).contiguous()
    return DeviceLexical(ids, scores)


def _doc(i):
    i = int(i)
    return f"doc-{i:08}" if i < (1 << 32) else f"lexonly-{i - (1 << 32):08}"


def _oracle_flow(fast_slab, quality_slab, fq, qq, k, lex_ids, lex_scores, multiplier=3, alpha=0.7):
    """SyncTwoTierSearcher::search_internal (sync_searcher.rs:616-1009) on the CPU oracle."""
    from oracle import fs_oracle as fo

    th = fo.host_threads()
    fetch = max(k * multiplier, k)
    rows, scores = fo.search_top_k(fast_slab, fq, fetch, threads=th)
    fast = [(_doc(r), int(r), np.float32(s)) for r, s in zip(rows, scores)]
    lex = [(_doc(i), float(s)) for i, s in zip(lex_ids, lex_scores)]
    out = {"initial": fo.rrf_fuse(lex, fast, k, 0)}
    if quality_slab is not None:
        qs, present = fo.scores_for_rows(quality_slab, qq, rows)
        blended = fo.blend_two_tier_aligned(fast, [float(s) if p else None for s, p in zip(qs, present)], alpha)
        out["refined"] = fo.rrf_fuse(lex, blended, k, 0)
    return out


def _fused_equal(got, want):
    if len(got) != len(want):
        return False
    for g, w in zip(got, want):
        if np.float64(g["rrf_score"]).view(np.uint64) != np.float64(w.rrf_score).view(np.uint64):
            return False
        if (int(g["semantic_rank"]) if g["semantic_rank"] >= 0 else None) != w.semantic_rank:
            return False
        if (int(g["lexical_rank"]) if g["lexical_rank"] >= 0 else None) != w.lexical_rank:
            return False
        if w.semantic_rank is not None and int(g["semantic_row"]) != w.semantic_index:
            return False
    return True


def _pct(x, p):
    return float(np.percentile(np.asarray(x), p))


def bench_config2(ctx):
    """configs[1]: 1 M x 384, cosine top-100 + RRF with precomputed BM25 ranks, one GPU; B = 1 and 32."""
    if ctx.world > 1:
        return None
    torch, dev = ctx.torch, ctx.dev
    from frankensearch_b200.pipeline import DeviceTwoTierSearcher, fused_to_numpy

    n, dim, k = 1_000_000, 384, 100
    ix, _, _ = ctx.synth_shard(1, n, dim)
    searcher = DeviceTwoTierSearcher(ix, None)
    fetch = searcher.fetch_for(k)
    q_np = np.empty((32, dim), dtype=np.float32)
    _fill_queries(q_np, dim)
    q_host = torch.from_numpy(q_np).pin_memory()
    rec = {"workload": f"configs[1]: {n} docs x {dim}-dim f16, exact cosine top-{k} (fetch {fetch}) + RRF with a "
                       f"precomputed {fetch}-entry BM25 list (k=60, weights 1), 1 GPU", "rows": n, "dim": dim, "k": k}
    out_host = torch.empty((32, k, 32), dtype=torch.uint8).pin_memory()
    for b in (1, 32):
        dq = q_host[:b].to(dev)
        warm = searcher.search_device(dq, None, k, None)
        lex = _lexical_from_hits(ctx, warm.fast_hits, n, fetch, seed=7)
        ms, res = _time_steps(ctx, lambda: searcher.search_device(dq, None, k, lex), max(ctx.a.steps, 20), 3)
        fused_snap, counts_snap = fused_to_numpy(res.initial), res.initial_counts.cpu().numpy()

        def e2e():
            d = q_host[:b].to(dev, non_blocking=True)
            r = searcher.search_device(d, None, k, lex)
            out_host[:b].copy_(r.initial, non_blocking=True)
            torch.cuda.synchronize(dev)

        lat = []
        for i in range(3 + 50):
            t = time.perf_counter()
            e2e()
            if i >= 3:
                lat.append((time.perf_counter() - t) * 1e3)
        rec[f"batch{b}"] = {"ms_per_step": ms, "queries_per_s": b / ms * 1e3, "e2e_p50_ms": _pct(lat, 50),
                            "e2e_p99_ms": _pct(lat, 99), "e2e_queries_per_s": b / _pct(lat, 50) * 1e3,
                            "h2d_bytes_per_step": b * dim * 4, "d2h_bytes_per_step": b * k * 32}
        if b == 32 and not ctx.a.no_cpu_baseline:
            host = ctx.host_slab(1, n, dim)
            li, ls = lex.ids.cpu().numpy(), lex.scores.cpu().numpy()
            ok = True
            t0 = time.perf_counter()
            for qi in range(4):
                want = _oracle_flow(host, None, q_np[qi], None, k, li[qi], ls[qi])
                ok = ok and _fused_equal(fused_snap[qi, : int(counts_snap[qi])], want["initial"])
            rec["parity"] = {"queries": 4, "fused_equal": ok, "against": "oracle scan + oracle rrf_fuse (rrf_score bits, ranks, rows)"}
            rec["cpu_baseline_ms_per_query"] = (time.perf_counter() - t0) / 4 * 1e3
    ix.close()
    return rec


def _synthetic_minilm_weights(seed=3, vocab=30522, hidden=384, layers=6, inter=1536, max_pos=512):
    """Random-init all-MiniLM-L6-v2 geometry (BERT init: N(0, 0.02), LayerNorm 1/0) under Hugging Face
    state-dict names — there are no model files on the box."""
    rng = np.random.default_rng(seed)
    w = {}
    nrm = lambda *s: (rng.standard_normal(s) * 0.02).astype(np.float32)  # noqa: E731
    w["embeddings.word_embeddings.weight"] = nrm(vocab, hidden)
    w["embeddings.position_embeddings.weight"] = nrm(max_pos, hidden)
    w["embeddings.token_type_embeddings.weight"] = nrm(2, hidden)
    w["embeddings.LayerNorm.weight"] = np.ones(hidden, np.float32)
    w["embeddings.LayerNorm.bias"] = np.zeros(hidden, np.float32)
    for i in range(layers):
        p = f"encoder.layer.{i}."
        for nme in ("query", "key", "value"):
            w[p + f"attention.self.{nme}.weight"] = nrm(hidden, hidden)
            w[p + f"attention.self.{nme}.bias"] = nrm(hidden)
        w[p + "attention.output.dense.weight"] = nrm(hidden, hidden)
        w[p + "attention.output.dense.bias"] = nrm(hidden)
        w[p + "attention.output.LayerNorm.weight"] = np.ones(hidden, np.float32)
        w[p + "attention.output.LayerNorm.bias"] = np.zeros(hidden, np.float32)
        w[p + "intermediate.dense.weight"] = nrm(inter, hidden)
        w[p + "intermediate.dense.bias"] = nrm(inter)
        w[p + "output.dense.weight"] = nrm(hidden, inter)
        w[p + "output.dense.bias"] = nrm(hidden)
        w[p + "output.LayerNorm.weight"] = np.ones(hidden, np.float32)
        w[p + "output.LayerNorm.bias"] = np.zeros(hidden, np.float32)
    return w


def _token_batch(ctx, batch, seed, vocab=30522):
    """Seeded token-id sequences, length ~ U[4, 32]: pinned host ids/lens."""
    rng = np.random.default_rng(seed)
    lens = rng.integers(4, 33, batch).astype(np.int32)
    ids = rng.integers(1000, vocab, (batch, 32)).astype(np.int32)
    ids[:, 0] = 101  # [CLS] ... [SEP]
    for i, l in enumerate(lens):
        ids[i, l - 1] = 102
        ids[i, l:] = 0
    return ctx.torch.from_numpy(ids).pin_memory(), ctx.torch.from_numpy(lens).pin_memory()


def bench_config4(ctx):
    """configs[3]: 10 M docs row-sharded over the ranks, two-tier: 256-d fast tier -> 384-d quality
    re-score -> blend 0.7 -> RRF with BM25 ranks, batch 1024, k = 10 (fetch 30); ONE all-gather of
    [keys | hits | quality] per search."""
    torch, dev, a = ctx.torch, ctx.dev, ctx.a
    import frankensearch_b200 as fs
    from frankensearch_b200.pipeline import DeviceTwoTierSearcher, fused_to_numpy

    n, k, batch = a.rows, 10, a.batch
    fast_ix, _, _ = ctx.synth_shard(2, n, 256)
    quality_ix, _, _ = ctx.synth_shard(1, n, 384)
    searcher = DeviceTwoTierSearcher(fast_ix, quality_ix)
    fetch = searcher.fetch_for(k)
    fq_np = np.empty((batch, 256), dtype=np.float32)
    qq_np = np.empty((batch, 384), dtype=np.float32)
    _fill_queries(fq_np, 256)
    _fill_queries(qq_np, 384)
    fq_host, qq_host = torch.from_numpy(fq_np).pin_memory(), torch.from_numpy(qq_np).pin_memory()
    fq, qq = fq_host.to(dev), qq_host.to(dev)
    warm = searcher.search_device(fq, qq, k, None)
    lex = _lexical_from_hits(ctx, warm.fast_hits, n, fetch, seed=11)
    ms, res = _time_steps(ctx, lambda: searcher.search_device(fq, qq, k, lex), a.steps, 3)
    # the searcher reuses its output buffers: snapshot the timed result before anything else runs
    snap = {"initial": fused_to_numpy(res.initial), "refined": fused_to_numpy(res.refined),
            "initial_counts": res.initial_counts.cpu().numpy(), "refined_counts": res.refined_counts.cpu().numpy()}
    out_host = torch.empty((batch, k, 32), dtype=torch.uint8).pin_memory()

    def e2e():
        f, q = fq_host.to(dev, non_blocking=True), qq_host.to(dev, non_blocking=True)
        r = searcher.search_device(f, q, k, lex)
        out_host.copy_(r.refined, non_blocking=True)
        torch.cuda.synchronize(dev)

    e2e_ms = _time_wall(ctx, e2e, a.steps, 2)
    # with the query encoders inside the step: potion (static table gather) for the fast tier (a replica on every rank:
    # it costs microseconds), MiniLM-L6 for the quality tier — at N > 1 every rank encodes its block of B/N queries and
    # ONE all-gather hands every rank all B embeddings (1.5 MB); synthetic weights and token ids
    rng = np.random.default_rng(5)
    potion = fs.Model2VecEmbedder((rng.standard_normal((65536, 256)) * 0.1).astype(np.float32), device=ctx.local_rank)
    minilm = fs.MiniLmEmbedder(_synthetic_minilm_weights(), device=ctx.local_rank)
    ids_h, lens_h = _token_batch(ctx, batch, seed=9)
    lens64 = lens_h.numpy().astype(np.uint64)
    off_h = torch.from_numpy(np.concatenate([[0], np.cumsum(lens64)]).astype(np.int64)).pin_memory()
    flat_h = torch.from_numpy(np.concatenate([ids_h.numpy()[i, :l] for i, l in enumerate(lens_h.numpy())]).astype(np.int32)).pin_memory()
    pf = torch.empty((batch, 256), dtype=torch.float32, device=dev)
    bq = (batch + ctx.world - 1) // ctx.world           # queries per rank (equal blocks for the all-gather)
    q_lo, q_hi = min(batch, ctx.rank * bq), min(batch, (ctx.rank + 1) * bq)
    q_block = torch.zeros((bq, 384), dtype=torch.float32, device=dev)
    q_all = torch.empty((ctx.world * bq, 384), dtype=torch.float32, device=dev)

    def with_encoders():
        d_ids, d_lens = ids_h.to(dev, non_blocking=True), lens_h.to(dev, non_blocking=True)
        d_flat, d_off = flat_h.to(dev, non_blocking=True), off_h.to(dev, non_blocking=True)
        s = torch.cuda.current_stream(dev).cuda_stream
        fs._ffi.check(fs._ffi.lib().fsgpu_potion_embed_device(potion._h, d_flat.data_ptr(), d_off.data_ptr(), batch,
                                                              pf.data_ptr(), s))
        if ctx.world == 1:
            q = minilm.embed_device(d_ids, d_lens)
        else:
            if q_hi > q_lo:
                q_block[: q_hi - q_lo] = minilm.embed_device(d_ids[q_lo:q_hi].contiguous(), d_lens[q_lo:q_hi].contiguous())
            ctx.dist.all_gather_into_tensor(q_all, q_block)
            q = q_all[:batch]
        r = searcher.search_device(pf, q, k, lex)
        out_host.copy_(r.refined, non_blocking=True)
        torch.cuda.synchronize(dev)

    enc_ms = _time_wall(ctx, with_encoders, max(3, a.steps // 2), 2)

    # the encoder alone (device-resident token ids): ms per batch, GEMM time and useful TFLOP/s from the library's own
    # events, for the default f16 form and the split-f16 form (FSGPU_MINILM_PRODUCTS=3)
    encoder = None
    if ctx.rank == 0:
        d_ids, d_lens = ids_h.to(dev), lens_h.to(dev)
        encoder = {"batch": batch, "tokens_per_query": "4-32 (t_pad 32)", "weights": "synthetic (seeded), all-MiniLM-L6-v2 geometry"}
        flop_per_row = 2.0 * 6 * (4 * 384 * 384 + 2 * 384 * 1536)
        token_rows = int(lens_h.sum().item())  # the f16 form packs rows (no padding rows); the split form pads to 32
        encoder["token_rows"] = {"packed (f16 form)": token_rows, "padded (split form)": batch * 32}
        for name, mode in (("f16_form", "0"), ("split_f16_form", "3")):
            useful_flop = flop_per_row * (token_rows if mode == "0" and os.environ.get("FSGPU_MINILM_PACKED", "1") != "0" else batch * 32)
            os.environ["FSGPU_MINILM_PRODUCTS"] = mode
            for _ in range(3):
                minilm.embed_device(d_ids, d_lens)
            torch.cuda.synchronize(dev)
            minilm.profile_read(reset=True)
            minilm.profile_enable(True)
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
            for _ in range(10):
                minilm.embed_device(d_ids, d_lens)
            t1.record()
            torch.cuda.synchronize(dev)
            pr = minilm.profile_read(reset=True)
            minilm.profile_enable(False)
            enc_only = t0.elapsed_time(t1) / 10
            encoder[name] = {"ms_per_batch": enc_only, "queries_per_s": batch / enc_only * 1e3, "gemm_ms": pr["gemm_ms"] / 10,
                             "gemm_launches_per_batch": pr["gemm_launches"] // 10,
                             "useful_tflops_whole_encoder": useful_flop / (enc_only * 1e-3) / 1e12,
                             "useful_tflops_in_gemms": useful_flop / (pr["gemm_ms"] / 10 * 1e-3) / 1e12}
        del os.environ["FSGPU_MINILM_PRODUCTS"]
        encoder["accuracy"] = ("f16 form: |component error| <= 5e-4, cosine >= 1 - 2e-6 vs PyTorch f32 (tests/test_gpu_minilm.py); "
                               "split form: <= 2e-4 (measured ~2e-6)")
        encoder["ncu"] = "profiles/r02_minilm_f16_form.md"
    rec = None
    if ctx.rank == 0:
        rec = {"workload": f"configs[3]: {n} docs, 256-d fast tier + 384-d quality tier, f16, row-sharded over {ctx.world} "
                           f"GPU(s); batch {batch}, k={k} (fetch {fetch}): exact fast top-{fetch} -> local quality re-score -> "
                           f"one all-gather [keys|hits|quality] -> blend 0.7 -> RRF (k=60) with {fetch}-entry BM25 lists",
               "n_gpus": ctx.world, "ms_per_step": ms, "queries_per_s": batch / ms * 1e3,
               "e2e_ms_per_step": e2e_ms, "e2e_queries_per_s": batch / e2e_ms * 1e3,
               "h2d_bytes_per_step": batch * (256 + 384) * 4, "d2h_bytes_per_step": batch * k * 32,
               "with_encoders": {"ms_per_step": enc_ms, "queries_per_s": batch / enc_ms * 1e3,
                                 "what": "host token ids -> potion gather-pool + MiniLM-L6 forward (synthetic weights, "
                                         "4-32 tokens; N > 1: each rank encodes B/N queries, one all-gather of the embeddings) -> the same "
                                         "search -> host results"},
               "encoder": encoder,
               "flow": "SyncTwoTierSearcher::search_internal, sync_searcher.rs:616-1009, pre-embedded queries"}
        if not a.no_cpu_baseline:
            fused_i, fused_r = snap["initial"], snap["refined"]
            li, ls = lex.ids.cpu().numpy(), lex.scores.cpu().numpy()
            fast_host, quality_host = ctx.host_slab(2, n, 256), ctx.host_slab(1, n, 384)
            ok_i = ok_r = True
            for qi in (0, batch // 2 + 1):
                want = _oracle_flow(fast_host, quality_host, fq_np[qi], qq_np[qi], k, li[qi], ls[qi])
                ok_i = ok_i and _fused_equal(fused_i[qi, : int(snap["initial_counts"][qi])], want["initial"])
                ok_r = ok_r and _fused_equal(fused_r[qi, : int(snap["refined_counts"][qi])], want["refined"])
            rec["parity"] = {"queries": 2, "initial_equal": ok_i, "refined_equal": ok_r,
                             "against": "oracle flow (scan, scores_for_rows, blend_two_tier_aligned, rrf_fuse) on the identical corpora"}
    ctx.barrier()
    potion.close()
    minilm.close()
    fast_ix.close()
    quality_ix.close()
    return rec


def bench_config5(ctx):
    """configs[4]: 50 M x 384 row-sharded over the ranks, per query: host token ids -> MiniLM-L6 encode ->
    exact scan, fetch 3000 (top-1000 x multiplier 3) -> cross-shard merge -> RRF with a 3000-entry BM25
    list -> top-1000 on the host.  p50 / p99 over >= 1000 queries, one at a time."""
    torch, dev, a = ctx.torch, ctx.dev, ctx.a
    import frankensearch_b200 as fs
    from frankensearch_b200.pipeline import DeviceTwoTierSearcher, fused_to_numpy

    n, dim, k = a.rows5, 384, 1000
    ix, lo, hi = ctx.synth_shard(3, n, dim)
    searcher = DeviceTwoTierSearcher(ix, None)
    fetch = searcher.fetch_for(k)
    minilm = fs.MiniLmEmbedder(_synthetic_minilm_weights(), device=ctx.local_rank)
    nq = a.queries5
    ids_h, lens_h = _token_batch(ctx, nq, seed=13)
    # one precomputed BM25 list per query slot (8 distinct lists reused round-robin), built from real candidates
    d_ids8, d_lens8 = ids_h[:8].to(dev), lens_h[:8].to(dev)
    warm = searcher.search_device(minilm.embed_device(d_ids8, d_lens8), None, k, None)
    lex8 = _lexical_from_hits(ctx, warm.fast_hits, n, fetch, seed=17)
    from frankensearch_b200.pipeline import DeviceLexical
    lex = [DeviceLexical(lex8.ids[i:i + 1].contiguous(), lex8.scores[i:i + 1].contiguous()) for i in range(8)]
    out_host = torch.empty((1, k, 32), dtype=torch.uint8).pin_memory()
    cnt_host = torch.empty(1, dtype=torch.int32).pin_memory()

    def one(i):
        d_ids, d_lens = ids_h[i:i + 1].to(dev, non_blocking=True), lens_h[i:i + 1].to(dev, non_blocking=True)
        q = minilm.embed_device(d_ids, d_lens)
        r = searcher.search_device(q, None, k, lex[i % 8])
        out_host.copy_(r.initial, non_blocking=True)
        cnt_host.copy_(r.initial_counts, non_blocking=True)
        torch.cuda.synchronize(dev)
        return r

    for i in range(5):
        one(i)
    ctx.barrier()
    lat = []
    stage = {"encode": [], "search": []}
    for i in range(nq):
        t = time.perf_counter()
        r = one(i)
        lat.append((time.perf_counter() - t) * 1e3)
    last_out, last_cnt = out_host.numpy().copy(), int(cnt_host[0])  # the answer of query nq - 1, on the host
    ctx.barrier()
    # device-side split of one query (events), for the record
    for i in range(20):
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        d_ids, d_lens = ids_h[i:i + 1].to(dev), lens_h[i:i + 1].to(dev)
        e0.record()
        q = minilm.embed_device(d_ids, d_lens)
        e1.record()
        searcher.search_device(q, None, k, lex[i % 8])
        e2.record()
        torch.cuda.synchronize(dev)
        stage["encode"].append(e0.elapsed_time(e1))
        stage["search"].append(e1.elapsed_time(e2))
    lat_all = lat
    if ctx.world > 1:  # a query is done when the slowest rank is done
        t = torch.tensor(lat, dtype=torch.float64, device=dev)
        ctx.dist.all_reduce(t, op=ctx.dist.ReduceOp.MAX)
        lat_all = t.cpu().numpy().tolist()
    rec = None
    if ctx.rank == 0:
        rec = {"workload": f"configs[4]: {n} docs x {dim}-dim f16 row-sharded over {ctx.world} GPU(s); per query: host token ids "
                           f"(4-32) -> MiniLM-L6 encode -> exact scan, fetch {fetch} -> merge -> RRF with a {fetch}-entry BM25 "
                           f"list -> top-{k} fused hits on the host; {nq} queries, one at a time",
               "n_gpus": ctx.world, "queries": nq, "p50_ms": _pct(lat_all, 50), "p99_ms": _pct(lat_all, 99),
               "mean_ms": float(np.mean(lat_all)), "queries_per_s": 1e3 / float(np.mean(lat_all)),
               "device_ms": {"encode_p50": _pct(stage["encode"], 50), "search_fuse_p50": _pct(stage["search"], 50)},
               "h2d_bytes_per_query": 32 * 4 + 4, "d2h_bytes_per_query": k * 32 + 4,
               "rows_per_gpu": hi - lo, "slab_gb_per_gpu": (hi - lo) * dim * 2 / 1e9}
        # size-independent properties of the last answer (a 38 GB host oracle scan is out of the bench's budget):
        # k fused hits, rrf scores non-increasing, every semantic score is the oracle's exact dot of that row
        from frankensearch_b200.pipeline import FUSED_HIT_DTYPE
        fused = last_out.view(FUSED_HIT_DTYPE).reshape(-1)[: last_cnt]
        ok_sorted = bool(np.all(np.diff(fused["rrf_score"]) <= 0)) and len(fused) == k
        rec["properties"] = {"count_is_k": len(fused) == k, "rrf_non_increasing": ok_sorted}
        if not a.no_cpu_baseline:
            from oracle import fs_oracle as fo

            sem = fused[fused["semantic_rank"] >= 0]
            mine = sem[(sem["semantic_row"] >= lo) & (sem["semantic_row"] < hi)][:64]
            if len(mine):
                rows_t = torch.from_numpy((mine["semantic_row"].astype(np.int64) - lo)).to(dev)
                slab_rows = ix._keepalive[rows_t].cpu().numpy().view(np.uint16)
                qv = minilm.embed_device(ids_h[nq - 1:nq].to(dev), lens_h[nq - 1:nq].to(dev)).cpu().numpy()[0]
                exact, _ = fo.scores_for_rows(slab_rows, qv, np.arange(len(mine), dtype=np.uint64))
                rec["properties"]["semantic_scores_are_exact_oracle_dots"] = bool(
                    np.array_equal(exact.view(np.uint32), mine["semantic_score"].view(np.uint32)))
    ctx.barrier()
    minilm.close()
    ix.close()
    return rec


def _fill_queries(q_np, dim):
    """clustered queries of the reference bench (fsvi_int8_two_pass.rs:285-287), NumPy restatement
    (host-side input preparation; ~1 ms per query, outside every timed region)."""
    def raw(seed):
        s = np.uint64(seed | 1)
        out = np.empty(dim, dtype=np.float32)
        s = int(s)
        for d in range(dim):
            s ^= (s << 13) & 0xFFFFFFFFFFFFFFFF
            s ^= s >> 7
            s ^= (s << 17) & 0xFFFFFFFFFFFFFFFF
            out[d] = np.float32(np.float32(s >> 40) / np.float32(8388608.0) - np.float32(1.0))
        return out

    def norm(v):
        acc = np.float32(0)
        for x in v:
            acc = np.float32(acc + np.float32(x * x))
        n = np.sqrt(acc)
        return (v / n).astype(np.float32) if n > 1e-12 else v

    cents = {}
    for q in range(q_np.shape[0]):
        c = q % 64
        if c not in cents:
            cents[c] = norm(raw(0xC0000000 + c))
        q_np[q] = norm((cents[c] + np.float32(0.30) * raw(0xDEAD0000 + q)).astype(np.float32))


def main():
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
