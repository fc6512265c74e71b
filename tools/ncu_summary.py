"""Summarise ncu outputs brought back in gpurun_out/ into small tracked files under profiles/.

  python tools/ncu_summary.py rep   gpurun_out/prof.ncu-rep  profiles/r01_scan.json
  python tools/ncu_summary.py list  gpurun_out/launches.csv  profiles/r01_launches.md
"""
import csv
import io
import json
import subprocess
import sys
from collections import OrderedDict

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "sm__cycles_elapsed.avg", "sm__cycles_active.avg",
    "smsp__cycles_active.avg", "sm__cycles_elapsed.avg.per_second",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "lts__t_sector_hit_rate.pct",
    "lts__t_bytes.sum", "l1tex__t_bytes.sum",
]


def rep(path, out):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], check=True, capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = OrderedDict(kernel=r[hdr.index("Kernel Name")])
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                d[w] = {"value": r[i], "unit": units[i]}
        # any tensor / tmem metric present in the capture
        for i, h in enumerate(hdr):
            if ("tensor" in h or "tmem" in h) and h not in d and r[i] not in ("", "0", "n/a"):
                d[h] = {"value": r[i], "unit": units[i]}
        res.append(d)
    json.dump({"source": path, "launches": res}, open(out, "w"), indent=1)
    print("wrote", out, len(res), "launches")


def launch_list(path, out):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 14 and r[12] == "gpu__time_duration.sum"]
    agg = OrderedDict()
    for r in rows:
        name = r[4].split("(")[0]
        a = agg.setdefault(name, [0, 0.0, r[7], r[8]])
        a[0] += 1
        a[1] += float(r[14].replace(",", "")) / 1e6  # ns -> ms
    total = sum(a[1] for a in agg.values())
    with open(out, "w") as f:
        f.write(f"ncu launch list ({path}): {len(rows)} launches, {total:.3f} ms total device time "
                f"(cold-cache, serialised: compare SHARES)\n\n")
        f.write("| kernel | launches | total ms | avg ms | share | block | grid |\n|---|---:|---:|---:|---:|---|---|\n")
        for name, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{name}` | {a[0]} | {a[1]:.3f} | {a[1] / a[0]:.4f} | {100 * a[1] / total:.1f}% | {a[2]} | {a[3]} |\n")
    print("wrote", out)


if __name__ == "__main__":
    {"rep": rep, "list": launch_list}[sys.argv[1]](sys.argv[2], sys.argv[3])
