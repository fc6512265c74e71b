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
    p.add_argument("--hbm-batch", type=int, default=128,
                   help="also time the HBM-bound regime with this many queries per pass (0 = skip)")
    return p.parse_args()


def workload_config(a, n_gpus):
    return {
        "workload": f"configs[2]: {a.rows} docs x {a.dim}-dim f16, batch {a.batch} queries, exact cosine top-{a.k}",
        "rows": a.rows, "dim": a.dim, "batch": a.batch, "k": a.k,
        "corpus": "clustered (64 centroids, noise 0.30) — reference bench generator fsvi_int8_two_pass.rs:199-231",
        "storage": "f16 slab (+ int8 codes of it for the candidate pass), f32 query; tensor-core candidate pass (int8 x int8 -> s32, or f16 x f16 -> f32 with FSGPU_MMA_I8=0) + exact f16 re-scoring with the reference accumulation tree (bit-exact results)",
        "sharding": "single GPU" if n_gpus == 1 else f"rows sharded over {n_gpus} ranks, one all-gather of top-k keys",
        "l2": "inputs larger than L2 (slab per GPU >> 126 MB); no explicit flush",
    }


# ── CPU arm: the oracle restatement of the reference scan on the host cores ──────────────────
def cpu_arm(a, steps, warmup):
    """Times oracle.fs_oracle.search_top_k (reference restatement: AVX2+F16C dot, 1024-row chunks,
    per-chunk heaps, serial merge; all host threads) on a bounded sample: the full-size corpus,
    `cpu_queries` queries of the batch per step, queries back to back (the reference's
    production model, benches/batched_query_scan.rs:108-126)."""
    from oracle import fs_oracle as fo

    rows = a.cpu_rows or a.rows
    threads = fo.host_threads()
    t0 = time.perf_counter()
    slab, _ = fo.synth_rows(1, 1, 0, rows, a.dim, threads=threads)
    gen_s = time.perf_counter() - t0
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
    per_query = sum(times) / (steps * a.cpu_queries)
    per_query_full = per_query * (a.rows / rows)  # the scan is linear in rows
    return dict(value=1.0 / per_query_full, per_query_ms=per_query_full * 1e3, threads=threads, rows=rows,
                gen_s=gen_s, ms_per_step=1e3 * sum(times) / steps,
                sample=f"{rows} rows x {a.dim} (of {a.rows}), {a.cpu_queries} of {a.batch} queries per step, "
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
                "source": "nvml" if self.nvml is not None else "nvidia-smi"}


# ── our arm ───────────────────────────────────────────────────────────────────────────────────
def run_ours(a):
    import torch
    import torch.distributed as dist

    import frankensearch_b200 as fs
    from frankensearch_b200.sharded import ShardedGpuIndex, shard_bounds

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product has no CPU fallback (use --impl reference "
                         "for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # corpus shard, generated on the device by the reference's bench generator
    lo, hi = shard_bounds(a.rows, world, rank)
    slab = torch.empty((hi - lo, a.dim), dtype=torch.int16, device=dev)
    fs._ffi.check(fs._ffi.lib().fsgpu_synth_rows_device(local_rank, 1, 1, lo, hi - lo, a.dim, 64, 0.30,
                                                        slab.data_ptr(), None))
    ix = fs.GpuVectorIndex.from_device_tensor(slab, row_base=lo)
    sharded = ShardedGpuIndex(ix) if world > 1 else None

    # queries: the reference generator's clustered queries, built on the host (tiny)
    from frankensearch_b200 import _ffi  # noqa: F401
    q_host = torch.empty((a.batch, a.dim), dtype=torch.float32).pin_memory()
    q_np = q_host.numpy()
    _fill_queries(q_np, a.dim)
    d_queries = q_host.to(dev, non_blocking=False)

    def step_device():
        if sharded is not None:
            return sharded.search_top_k_device(d_queries, a.k)
        return ix.search_top_k_device(d_queries, a.k, want_hits=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # the HBM-bound regime of the same kernel: one query block (<= 128 queries) per corpus pass
    hbm = None
    if a.hbm_batch > 0:
        qb = d_queries[: min(a.hbm_batch, a.batch)].contiguous()

        def step_hbm():
            if sharded is not None:
                return sharded.search_top_k_device(qb, a.k)
            return ix.search_top_k_device(qb, a.k, want_hits=True)

        for _ in range(3):
            step_hbm()
        barrier()
        ix.profile_read(reset=True)
        ix.profile_enable(True)
        h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        h0.record()
        for _ in range(a.steps):
            step_hbm()
        h1.record()
        barrier()
        hp = ix.profile_read(reset=True)
        ix.profile_enable(False)
        h_ms = hp["scan_ms"] / max(hp["scan_launches"], 1)
        hbm = {"batch": int(qb.shape[0]), "avg_launch_ms": h_ms,
               "bytes_per_launch": hp["scan_bytes"] / max(hp["scan_launches"], 1),
               "achieved_gbs": hp["scan_bytes"] / max(hp["scan_launches"], 1) / (h_ms * 1e-3) / 1e9 if h_ms else 0.0,
               "ms_per_step": h0.elapsed_time(h1) / a.steps,
               "kernel": "mma_scan_kernel" if hp["mma_launches"] else "scan_topk_fast_kernel"}

    for _ in range(max(a.warmup, 3)):
        step_device()
    barrier()
    ix.profile_read(reset=True)
    ix.profile_enable(True)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    start.record()
    for _ in range(a.steps):
        out = step_device()
    end.record()
    barrier()
    elapsed_ms = start.elapsed_time(end)
    prof = ix.profile_read(reset=True)
    ix.profile_enable(False)
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t.item())
    ms_per_step = elapsed_ms / a.steps
    value = a.batch * a.steps / (elapsed_ms / 1e3)

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

    step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        step_e2e()
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = a.batch * a.steps / float(t.item())

    # sanity: the timed result is a real answer (sorted keys, k hits per query)
    keys = out[0]
    assert bool((keys[:, :-1] > keys[:, 1:]).all().item()) if a.k > 1 else True, "result keys are not sorted"

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except (OSError, ValueError):
            pass
        peak_gbs = float(peaks.get("hbm_gbs", 6650.0))
        peak_tf = float(peaks.get("bf16_tflops", 1590.0))
        peak_tf_sus = float(peaks.get("bf16_tflops_sustained", 1400.0))
        peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md: 6650 GB/s, 1590 TFLOP/s)"
        scan_launches = max(prof["scan_launches"], 1)
        avg_ms = prof["scan_ms"] / scan_launches
        bytes_per_launch = prof["scan_bytes"] / scan_launches
        hbm_achieved = bytes_per_launch / (avg_ms * 1e-3) / 1e9 if avg_ms > 0 else 0.0
        traffic = None
        tp = os.path.join(ROOT, "profiles", "scan_kernel_traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get("dram_bytes_per_launch")
            except (OSError, ValueError):
                traffic = None
        if prof["mma_launches"]:
            # a 1024-query batch is a dense [B,D]x[D,N] contraction: tensor-bound
            flops = prof["mma_flops"] / prof["mma_launches"]
            achieved_tf = flops / (avg_ms * 1e-3) / 1e12 if avg_ms > 0 else 0.0
            i8 = prof.get("i8_launches", 0) > 0
            if i8:
                # the int8 form: kind::i8 MMAs retire twice the K per instruction of kind::f16/bf16 at the
                # same instruction rate, so the tensor peak is 2x the measured bf16 figure (no int8 figure
                # is in MEASURED_PEAKS.json); ops = 2 per int8 multiply-add
                peak_tf, peak_tf_sus = 2.0 * peak_tf, 2.0 * peak_tf_sus
                peak_src += "; int8 tensor peak = 2 x measured bf16"
                traffic = None  # the figure read above belongs to the f16 form
                tp8 = os.path.join(ROOT, "profiles", "scan_kernel_traffic_i8.json")
                if os.path.exists(tp8):
                    try:
                        traffic = json.load(open(tp8)).get("dram_bytes_per_launch")
                    except (OSError, ValueError):
                        traffic = None
            roofline = {"bound": "tensor", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                        "frac": achieved_tf / peak_tf if peak_tf else None, "traffic": traffic,
                        "kernel": ("mma_scan_pair_kernel<int8> (tcgen05.mma.cta_group::2 kind::i8, M=256 queries x N=256 "
                                   "rows per CTA pair, K=dim; exact: f16 re-score of the candidate superset)") if i8 else
                                  "mma_scan_pair_kernel (tcgen05.mma.cta_group::2 kind::f16, M=256 queries x N=256 rows per CTA pair, K=dim)",
                        "flops_per_launch": flops, "frac_of_sustained_peak": achieved_tf / peak_tf_sus,
                        "peak_sustained": peak_tf_sus, "ops": "int8 multiply-add = 2 ops" if i8 else "f16 multiply-add = 2 flops",
                        "hbm_gbs_same_launch": hbm_achieved, "hbm_frac_same_launch": hbm_achieved / peak_gbs}
        else:
            roofline = {"bound": "hbm", "achieved": hbm_achieved, "peak": peak_gbs, "unit": "GB/s",
                        "frac": hbm_achieved / peak_gbs if peak_gbs else None, "traffic": traffic,
                        "kernel": "scan_topk_fast_kernel"}
        roofline.update({"launches_timed": prof["scan_launches"], "avg_launch_ms": avg_ms,
                         "bytes_per_launch": bytes_per_launch,
                         "queries_per_launch": a.batch * a.steps / scan_launches, "peak_source": peak_src,
                         "scan_share_of_step": prof["scan_ms"] / elapsed_ms if elapsed_ms else None,
                         "redo_queries": prof["redo_queries"]})
        if hbm is not None:
            hbm["peak_gbs"] = peak_gbs
            hbm["frac"] = hbm["achieved_gbs"] / peak_gbs if peak_gbs else None
            hbm["note"] = ("same kernel, one 128-query block per corpus pass: the HBM-bound regime "
                           "(algorithmic bytes per launch = rows*dim*2 for the f16 form, rows*dim for the int8 form: "
                           "see bytes_per_launch)")
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps,
            "warmup": max(a.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(a, world),
            "roofline": roofline,
            "roofline_hbm_regime": hbm,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": a.batch * a.dim * 4,
                    "d2h_bytes_per_step": a.batch * a.k * 8 + a.batch * 4},
            "gpu_launches": prof["scan_launches"] + prof["merge_launches"] + prof["other_launches"],
            "clocks": clocks,
        }
        if not a.no_cpu_baseline and world == 1:
            aa = argparse.Namespace(**vars(a))
            r = cpu_arm(aa, steps=2, warmup=1)
            line["cpu_baseline"] = {"value": r["value"], "unit": UNIT, "cores": r["threads"], "kind": "port",
                                    "sample": r["sample"], "cpu": cpu_model(), "per_query_ms": r["per_query_ms"]}
        print(json.dumps(line), flush=True)
    ix.close()
    if world > 1:
        dist.destroy_process_group()


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
